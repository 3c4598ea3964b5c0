"""Conditional vectors resident in HBM (pb200_cond_update / pb200_branch_lnl, csrc/cond.cuh -- SURVEY.md 8f rank 2, the
kernels under `method = 1`).  Checkers:
  * node vectors: the oracle's conP (oracle/paml_oracle.c restates ConditionalPNode + NodeScale and is pinned
    bit-exact on the reference's fixtures): every internal node, every class, every pattern, after a full update;
  * branch function: rooting the tree at ANY node leaves the likelihood of a reversible model unchanged, so the
    value of lfunt at the branch's own length, computed from the vectors of the tree re-rooted at that branch
    (ReRootTree + updateconP semantics, restated in the test), must equal the reference's -lnL of the fixture; its
    first and second derivatives are checked against central differences of the value in t."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_ids
from oracle import oracle
from paml_b200.fixture import load_fixture
from paml_b200.likelihood import TreeLikelihood

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tl():
    t = TreeLikelihood(0)
    yield t
    t.close()


def _postorder(p, top, sons_of):
    out = []

    def rec(v):
        for c in sons_of[v]:
            if c >= p.ns:
                rec(c)
        out.append((v, list(sons_of[v])))
    rec(top)
    return out


def _sons(p):
    return {v: [int(c) for c in p.sons[v, :p.nson[v]]] for v in range(p.nnode)}


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_node_vectors_match_oracle(tl, path):
    p = load_fixture(path)
    tl.set_problem(p)
    ref = oracle.lnl(p, want_conP=True)
    flags = p.scale_flags if p.nscale else None
    tl.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, _postorder(p, p.root, _sons(p)), flags)
    for node in range(p.ns, p.nnode):
        got, _ = tl.cond_get(node)
        exp = ref["conP"][:, node - p.ns]
        # entries built from P(t) elements that are themselves cancellation residues (codons three substitutions apart at a
        # short t: P_ij ~ 1e-9 out of O(1) terms) carry ~1e-8 relative noise in the reference as much as here
        scale = np.maximum(np.abs(exp).max(axis=2, keepdims=True), 1e-300)
        assert (np.abs(got - exp) <= 1e-12 * scale + 1e-6 * np.abs(exp)).all(), (node, float(np.abs(got - exp).max()))


def _reroot(p, newroot):
    """ReRootTree (src/treespace.c:236-279) restated: new son lists, and the branch entries (set, t) moved along the path."""
    father = {v: int(p.father[v]) for v in range(p.nnode)}
    sons = _sons(p)
    bset, bt = p.branch_set.copy(), p.branch_t.copy()
    path = [newroot]
    while path[-1] != p.root:
        path.append(father[path[-1]])
    for x, y in zip(path[:-1], path[1:]):       # edge x (child) - y (old father): y becomes a son of x
        sons[y].remove(x)
        sons[x].append(y)
    for x, y in reversed(list(zip(path[:-1], path[1:]))):   # from the old root down: y takes the branch that was above x
        bset[:, y] = p.branch_set[:, x]
        bt[:, y] = p.branch_t[:, x]
    bset[:, newroot] = -1
    return sons, bset, bt, path


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_branch_function_rerooted(tl, path):
    p = load_fixture(path)
    if p.set_kind not in (0, 1):
        pytest.skip("pb200_branch_lnl needs UVROOT or CIJK sets")
    if "nhomo" in path:
        pytest.skip("nonhomogeneous model: the likelihood depends on where the root is, no pulley principle to test against")
    tl.set_problem(p)
    flags = p.scale_flags if p.nscale else None
    sem = 1 if p.semantics == 1 else 0
    rng = np.random.default_rng(5)
    internal = [v for v in range(p.ns, p.nnode) if v != p.root]
    tips = list(range(p.ns))
    picks = [int(rng.choice(internal))] if internal else []
    picks += [int(rng.choice(tips)), int(p.sons[p.root, 0])]
    for b in picks:
        a = int(p.father[b])
        if a < p.ns:
            continue
        sons, bset, bt, _ = _reroot(p, b)
        upd = []
        if b >= p.ns:                               # b's own side: its original subtree (without a)
            own = dict(sons); own[b] = [c for c in sons[b] if c != a]
            upd += _postorder(p, b, own)
        upd += _postorder(p, a, sons)               # a's side: everything else
        tl.cond_update(p.set_kind, p.nr, p.set_data, bset, bt, upd, flags)
        t = float(bt[:, a].max())                    # parametrise by the fastest class, so that a step in t is a sane step for all
        mult = bt[:, a] / t
        out = tl.branch_lnl(p.freqK, a, b, t, mult, bset[:, a], p.pi[:p.n], sem, True)
        assert abs(out[0] - p.ref_neg_lnL) <= 2e-10 * abs(p.ref_neg_lnL), (b, out[0], p.ref_neg_lnL)
        val = tl.branch_lnl(p.freqK, a, b, t, mult, bset[:, a], p.pi[:p.n], sem, False)
        assert abs(val[0] - out[0]) <= 1e-12 * abs(out[0])
        # derivatives against central differences at two step sizes; their disagreement bounds the truncation error
        e = 1e-4 * max(t, 1e-2)
        f = {k: tl.branch_lnl(p.freqK, a, b, t + k * e, mult, bset[:, a], p.pi[:p.n], sem, True) for k in (-2, -1, 1, 2)}
        d1, d1w = (f[1][0] - f[-1][0]) / (2 * e), (f[2][0] - f[-2][0]) / (4 * e)
        d2, d2w = (f[1][1] - f[-1][1]) / (2 * e), (f[2][1] - f[-2][1]) / (4 * e)
        tol1 = 1e-6 * max(1.0, abs(d1)) + 2 * abs(d1 - d1w) + 1e-12 * abs(out[0]) / e
        tol2 = 1e-6 * max(1.0, abs(d2)) + 2 * abs(d2 - d2w) + 1e-12 * max(1.0, abs(out[1])) / e
        assert abs(out[1] - d1) <= tol1, (b, out[1], d1, tol1)
        assert abs(out[2] - d2) <= tol2, (b, out[2], d2, tol2)
