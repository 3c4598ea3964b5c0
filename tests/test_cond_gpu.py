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


def test_eigen_batch_growth_between_branch_calls():
    """Regression (round-1 ADVICE, high): pb200_eigen_qrev growing its pinned buffer used to free the pinned staging blocks of
    pb200_cond_update / pb200_branch_lnl.  cond_update -> branch_lnl -> a larger eigen batch -> branch_lnl / cond_update on the
    SAME context must keep giving the same numbers, and pb200_destroy must not double-free."""
    p = load_fixture([f for f in GOLDEN if f.endswith("lysozyme_m0.pbfx")][0])
    t = TreeLikelihood(0)
    t.set_problem(p)
    upd = _postorder(p, p.root, _sons(p))
    t.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, upd, None)
    b = int(p.sons[p.root, 0]); a = p.root
    if b < p.ns:
        b = [int(c) for c in p.sons[p.root, :p.nson[p.root]] if c >= p.ns][0]
    # the branch b -> root from the vectors as they are: not the tree's likelihood (no rerooting), just a fixed function to repeat
    args = (p.freqK, a, b, float(p.branch_t[0, b]), np.ones(1), p.branch_set[:, b], p.pi[:p.n], 0, True)
    first = t.branch_lnl(*args)
    n = p.n
    s = p.set_data[0]
    Root, U, V = s[:n], s[n:n + n * n].reshape(n, n), s[n + n * n:n + 2 * n * n].reshape(n, n)
    Q = (U * Root[None, :]) @ V
    pi = np.asarray(p.pi[:n], np.float64)
    for njob in (1, 3, 17):                            # every call grows the solver's pinned mirror
        R2, _, _, sw = t.eigen_qrev(np.tile(Q, (njob, 1, 1)), np.tile(pi, (njob, 1)))
        assert (sw > 0).all() and np.abs(np.sort(R2[0]) - np.sort(Root)).max() < 1e-10
        again = t.branch_lnl(*args)
        assert (again == first).all(), (njob, first, again)
        t.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, upd, None)
        assert (t.branch_lnl(*args) == first).all()
    t.close()


def test_set_tree_invalidates_resident_vectors_and_begin_checks_pending():
    """pb200_set_tree always drops the resident conditional vectors (they were computed for the previous topology, even one
    with the same node count); pb200_set_data with another tip count requires a new pb200_set_tree; a second
    pb200_eval_begin before pb200_eval_end is refused (the pinned staging block is still being read)."""
    import ctypes as C
    from paml_b200.likelihood import PB200Error, _ptr
    p = load_fixture([f for f in GOLDEN if f.endswith("lysozyme_m0.pbfx")][0])
    t = TreeLikelihood(0)
    t.set_problem(p)
    t.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, _postorder(p, p.root, _sons(p)), None)
    t.cond_get(p.root)
    t.set_tree(p.root, p.nson, p.sons, p.scale_flags if p.nscale else None)      # same tree again: still an invalidation
    with pytest.raises(PB200Error):
        t.cond_get(p.root)
    b = int([c for c in p.sons[p.root, :p.nson[p.root]] if c >= p.ns][0])
    with pytest.raises(PB200Error):
        t.branch_lnl(p.freqK, p.root, b, 0.1, np.ones(1), p.branch_set[:, b], p.pi[:p.n], 0, True)
    # begin twice
    L = t._L
    fk = np.ascontiguousarray(p.freqK, np.float64); sd = np.ascontiguousarray(p.set_data, np.float64)
    bs = np.ascontiguousarray(p.branch_set, np.int32); bt = np.ascontiguousarray(p.branch_t, np.float64)
    pi = np.ascontiguousarray(p.pi, np.float64)
    args = (t._ctx, C.c_int(p.semantics), C.c_int(fk.shape[0]), _ptr(fk, C.c_double), C.c_int(p.set_kind), C.c_int(p.nr),
            C.c_int(sd.shape[0]), C.c_longlong(sd.shape[1]), _ptr(sd, C.c_double), _ptr(bs, C.c_int), _ptr(bt, C.c_double),
            _ptr(pi, C.c_double))
    assert L.pb200_eval_begin(*args) == 0
    assert L.pb200_eval_begin(*args) != 0 and b"not been collected" in L.pb200_last_error()
    out = C.c_double(0)
    assert L.pb200_eval_end(t._ctx, C.byref(out)) == 0
    assert abs(out.value - p.ref_neg_lnL) <= 1e-11 * abs(p.ref_neg_lnL)
    # other tip count without a new tree
    z2 = np.ascontiguousarray(p.z[:p.ns - 1]); t.set_data(p.n, z2, p.fpatt, p.cleandata, p.nchara, p.charamap)
    with pytest.raises(PB200Error):
        t.eval_problem(p)
    t.close()


import glob as _glob
import os as _os
BRANCH_FIXTURES = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "branch", "*.pbfx")))


@pytest.mark.parametrize("path", BRANCH_FIXTURES, ids=[_os.path.basename(f)[:-5] for f in BRANCH_FIXTURES])
@pytest.mark.parametrize("dmma", ["1", "0"])
def test_branch_derivatives_against_reference_lfuntdd(path, dmma, monkeypatch):
    """pb200_branch_lnl against the REFERENCE's own lfuntdd / lfuntdd_SiteClass (src/treesub.c:8204-8300, 8405-8543): fixtures
    dumped by oracle/ref_hook.inc from `method = 1` runs of the unmodified reference, stopped inside minbranches -- the tree as
    rerooted at that moment (root = b), the parameters and branch lengths of that moment, and the reference's -lnL(t), d/dt and
    d2/dt2 for the branch b -> a.  Both kernel families are checked: the tensor path (n > 4) and the FP64-vector kernels."""
    monkeypatch.setenv("PB200_COND_DMMA", dmma)
    p = load_fixture(path)
    r = p.branch_rec
    a, b = r["a"], r["b"]
    assert b == p.root and a >= p.ns
    if dmma == "1" and p.n <= 4:
        pytest.skip("4-state alphabets use the vector kernels only")
    t_ = TreeLikelihood(0)
    t_.set_problem(p)
    sons = _sons(p)
    own = dict(sons); own[b] = [c for c in sons[b] if c != a]          # b keeps the vector of its side only (src/treesub.c:8080-8083)
    upd = _postorder(p, a, sons) + _postorder(p, b, own)
    flags = p.scale_flags if p.nscale else None
    t_.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, upd, flags)
    mult = p.branch_t[:, a] / r["branch_len"]
    out = t_.branch_lnl(p.freqK, a, b, r["t"], mult, p.branch_set[:, a], p.pi[:p.n], 1 if r["siteclass"] else 0, True)
    t_.close()
    assert abs(out[0] - r["l"]) <= 1e-10 * abs(r["l"]), (out[0], r["l"])
    assert abs(out[1] - r["dl"]) <= 1e-8 * max(1.0, abs(r["dl"])) + 1e-10 * abs(r["ddl"]), (out[1], r["dl"])
    assert abs(out[2] - r["ddl"]) <= 1e-8 * abs(r["ddl"]), (out[2], r["ddl"])


def _host_branch_iteration(t_, p, r, mult, sem, l0, first_cycle, e, tb=(1e-8, 50.0)):
    """The reference's "iterating a branch" loop (src/treesub.c:8087-8108) driven from the host over pb200_branch_lnl."""
    a, b = r["a"], r["b"]
    t = t0 = r["t"]
    l = l0
    dl, ddl, nev = 0.0, -1.0, 0
    small = 1e-20
    for _ in range(10):
        o = t_.branch_lnl(p.freqK, a, b, t, mult, p.branch_set[:, a], p.pi[:p.n], sem, True); nev += 1
        dl, ddl = o[1], o[2]
        pp = -dl / abs(ddl)
        if abs(pp) < small:
            step = 0.0
        elif pp < 0:
            step = min(1.0, (tb[0] - t0) / pp)
        else:
            step = min(1.0, (tb[1] - t0) / pp)
        if first_cycle and step != 1 and step != 0:
            step *= 0.99
        while step > small:
            t = t0 + step * pp
            l = t_.branch_lnl(p.freqK, a, b, t, mult, p.branch_set[:, a], p.pi[:p.n], sem, False)[0]; nev += 1
            if l < l0:
                break
            step /= 4
        if step <= small:
            t, l = t0, l0
            break
        if abs(t - t0) < e * abs(1 + t) and abs(l - l0) < e:
            break
        t0, l0 = t, l
    return t, l, dl, ddl, nev


@pytest.mark.parametrize("path", BRANCH_FIXTURES, ids=[_os.path.basename(f)[:-5] for f in BRANCH_FIXTURES])
@pytest.mark.parametrize("first_cycle", [True, False])
def test_branch_newton_on_device_equals_host_driven_loop(path, first_cycle):
    """pb200_branch_newton (the iteration of one branch as a device-side WHILE graph) must take exactly the steps of the
    reference's loop driven from the host over single evaluations: same t, -lnL, dl, ddl to the bit, same evaluation count.
    Run twice on the same context (the captured graph is reused) and from a displaced start (several Newton rounds)."""
    p = load_fixture(path)
    r = dict(p.branch_rec)
    a, b = r["a"], r["b"]
    t_ = TreeLikelihood(0)
    t_.set_problem(p)
    sons = _sons(p)
    own = dict(sons); own[b] = [c for c in sons[b] if c != a]
    t_.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, _postorder(p, a, sons) + _postorder(p, b, own),
                   p.scale_flags if p.nscale else None)
    mult = p.branch_t[:, a] / r["branch_len"]
    sem = 1 if r["siteclass"] else 0
    for start in (r["t"], 3.0 * r["t"] + 0.05):
        r["t"] = start
        l0 = t_.branch_lnl(p.freqK, a, b, start, mult, p.branch_set[:, a], p.pi[:p.n], sem, False)[0]
        want = _host_branch_iteration(t_, p, r, mult, sem, l0, first_cycle, 1e-6)
        got = t_.branch_newton(p.freqK, a, b, start, l0, mult, p.branch_set[:, a], p.pi[:p.n], sem, first_cycle, 1e-6)
        assert got[4] == want[4] and got[4] >= 2, (got, want)
        assert got[:4] == want[:4], (got, want)
    t_.close()
