"""Tree topology in PAML's node numbering (tips 0..ns-1, internal nodes ns.. in order of '(' in the Newick
string, src/treesub.c:3098-3118): a small Newick reader and a seeded random-tree generator."""
import numpy as np


class Tree:
    def __init__(self, ns, father, sons_list, branch, root):
        self.ns = ns
        self.nnode = len(father)
        self.root = root
        self.father = np.array(father, np.int32)
        self.nson = np.array([len(s) for s in sons_list], np.int32)
        self.maxsons = max(2, int(self.nson.max()))
        self.sons = -np.ones((self.nnode, self.maxsons), np.int32)
        for i, s in enumerate(sons_list):
            self.sons[i, :len(s)] = s
        self.branch = np.array(branch, float)

    def postorder_internal(self):
        out, st = [], [(self.root, 0)]
        while st:
            v, k = st.pop()
            if k < self.nson[v]:
                st.append((v, k + 1))
                c = self.sons[v, k]
                if self.nson[c] > 0:
                    st.append((c, 0))
            else:
                out.append(v)
        return out


def parse_newick(s, names=None):
    """Parse a Newick string with tips labelled 1..ns (or by `names`) into a Tree."""
    s = s.strip().rstrip(";")
    ntips = sum(1 for tok in _tokens(s) if tok not in "(),:" and not tok.startswith(":"))
    # first pass counts tips properly
    toks = list(_tokens(s))
    tips = [t for i, t in enumerate(toks) if t not in ("(", ")", ",") and not t.startswith(":") and (i == 0 or toks[i - 1] in ("(", ","))]
    ns = len(tips)
    father, sons, branch = [-1] * ns, [[] for _ in range(ns)], [0.0] * ns
    stack, cur_closed = [], None
    root = -1
    i = 0
    last = None
    while i < len(toks):
        t = toks[i]
        if t == "(":
            v = len(father)
            father.append(-1); sons.append([]); branch.append(0.0)
            if stack:
                father[v] = stack[-1]; sons[stack[-1]].append(v)
            else:
                root = v
            stack.append(v); last = None
        elif t == ")":
            last = stack.pop()
        elif t == ",":
            last = None
        elif t.startswith(":"):
            branch[last] = float(t[1:])
        else:
            if i > 0 and toks[i - 1] == ")":
                pass                    # internal label: ignored
            else:
                v = (names.index(t) if names else int(t) - 1)
                father[v] = stack[-1]; sons[stack[-1]].append(v); last = v
        i += 1
    return Tree(ns, father, sons, branch, root)


def _tokens(s):
    i = 0
    while i < len(s):
        c = s[i]
        if c in "(),":
            yield c; i += 1
        elif c == ":":
            j = i + 1
            while j < len(s) and s[j] not in "(),": j += 1
            yield ":" + s[i + 1:j].strip(); i = j
        elif c.isspace():
            i += 1
        else:
            j = i
            while j < len(s) and s[j] not in "(),:": j += 1
            yield s[i:j].strip(); i = j


def random_tree(ns, seed, bl_lo=0.02, bl_hi=0.2, shape="random"):
    """Seeded random unrooted binary tree (trifurcating root) as a Newick string -> Tree.
    shape: "random" (random joins), "balanced", or "caterpillar"."""
    rng = np.random.default_rng(seed)
    def bl(): return ":%.6f" % rng.uniform(bl_lo, bl_hi)
    items = [str(i + 1) for i in range(ns)]
    if shape == "caterpillar":
        cur = items[0]
        rest = items[1:]
        while len(rest) > 2:
            cur = "(%s%s,%s%s)" % (cur, bl(), rest.pop(0), bl())
        return parse_newick("(%s%s,%s%s,%s%s);" % (cur, bl(), rest[0], bl(), rest[1], bl()))
    while len(items) > 3:
        if shape == "balanced":
            nxt = []
            k = 0
            while k + 1 < len(items) and len(nxt) + (len(items) - k) > 3:
                nxt.append("(%s%s,%s%s)" % (items[k], bl(), items[k + 1], bl())); k += 2
            items = nxt + items[k:]
        else:
            i, j = sorted(rng.choice(len(items), 2, replace=False))
            a, b = items[i], items[j]
            items = [x for k, x in enumerate(items) if k not in (i, j)] + ["(%s%s,%s%s)" % (a, bl(), b, bl())]
    return parse_newick("(%s%s,%s%s,%s%s);" % (items[0], bl(), items[1], bl(), items[2], bl()))
