"""Parity of the CUDA path (through the C ABI) with the reference and the oracle, on a real B200.

Gates (north_star): total lnL within 1e-6 relative of the reference -- we hold 1e-11; per-pattern log f_h to
1e-9 absolute; P(t) to 1e-13.  Integer work (pattern indexing, ambiguity code compaction) must be exact, which
the per-pattern comparison in pattern order checks."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_ids
from paml_b200.fixture import load_fixture
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood, PB200Error

pytestmark = pytest.mark.gpu

REL_LNL = 1e-11      # tolerance on total lnL (relative); north_star asks for 1e-6
ABS_LNF = 1e-9       # per-pattern log f_h (absolute)


@pytest.fixture(scope="module")
def tl():
    t = TreeLikelihood(0)
    yield t
    t.close()


def _check_fhK(got, ref, is_log):
    """com.fhK per class.  A class under which a pattern is (nearly) impossible has f(x|class) at round-off
    level of the dominant class (P(t) entries ~1e-17 from cancellation in I + sum_k ...), in the reference as
    much as here, so the tolerance is relative to the pattern's largest class term, not to each entry."""
    if is_log:
        top = ref.max(0, keepdims=True)
        np.testing.assert_allclose(np.exp(got - top), np.exp(ref - top), rtol=0, atol=1e-10)
    else:
        top = ref.max(0, keepdims=True)
        assert (np.abs(got - ref) <= 1e-10 * top).all()


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_golden_fixture(tl, path):
    p = load_fixture(path)
    tl.set_problem(p)
    v = tl.eval_problem(p)
    assert abs(v - p.ref_neg_lnL) <= REL_LNL * abs(p.ref_neg_lnL), (v, p.ref_neg_lnL)
    np.testing.assert_allclose(tl.lnf(), p.ref_lnf, rtol=0, atol=ABS_LNF)
    if p.ref_fhK is not None:
        _check_fhK(tl.fhK(), p.ref_fhK, p.nscale > 0)
    # evaluation is deterministic call to call (the optimiser differentiates it numerically)
    assert tl.eval_problem(p) == v


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_pmat_kernel_vs_reference(tl, path):
    p = load_fixture(path)
    if not p.ref_P:
        pytest.skip("no P(t) in fixture")
    mat_set = [p.branch_set[c, nd] for c, nd, _ in p.ref_P]
    mat_t = [p.branch_t[c, nd] for c, nd, _ in p.ref_P]
    P = tl.pmat(p.n, p.set_kind, p.nr, p.set_data, mat_set, mat_t)
    for i, (_, _, Pref) in enumerate(p.ref_P):
        np.testing.assert_allclose(P[i], Pref, rtol=0, atol=2e-14)


def test_pmat_edge_cases(tl):
    from oracle import oracle
    p = load_fixture([g for g in GOLDEN if g.endswith("lysozyme_m0.pbfx")][0])
    ts = np.array([0.0, 1e-200, 1e-9, 1e-3, 0.5, 5.0, 50.0])
    P = tl.pmat(p.n, p.set_kind, p.nr, p.set_data, np.zeros(len(ts), np.int32), ts)
    for i, t in enumerate(ts):
        np.testing.assert_allclose(P[i], oracle.pmat(t, p.n, p.set_kind, p.nr, p.set_data[0]), rtol=0, atol=2e-14)
    assert np.array_equal(P[0], np.eye(61)) and np.array_equal(P[1], np.eye(61))
    assert (P >= 0).all()


def _check_against_oracle(tl, p, rel=REL_LNL):
    from oracle import oracle
    r = oracle.lnl(p)
    tl.set_problem(p)
    v = tl.eval_problem(p)
    assert abs(v - r["neg_lnL"]) <= rel * abs(r["neg_lnL"]), (v, r["neg_lnL"])
    np.testing.assert_allclose(tl.lnf(), r["lnf"], rtol=0, atol=ABS_LNF)
    return v


@pytest.mark.parametrize("npatt", [1, 7, 8, 9, 16, 17, 255, 1000, 4097])
def test_codon_ragged_sizes(tl, npatt):
    p = simulate.codon_m0_problem(npatt=max(npatt, 64), seed=5)
    p.z = p.z[:, :npatt].copy(); p.fpatt = p.fpatt[:npatt].copy(); p.npatt = npatt
    _check_against_oracle(tl, p)


def test_codon_sites_classes(tl):
    _check_against_oracle(tl, simulate.codon_sites_problem(ns=16, npatt=700, seed=3))


def test_codon_deep_trees_with_scaling(tl):
    for shape in ("caterpillar", "balanced", "random"):
        from paml_b200.tree import random_tree
        p = simulate.codon_m0_problem(ns=40, npatt=300, seed=9)
        assert p.nscale > 0
        _check_against_oracle(tl, p)


def test_aa_gamma(tl):
    _check_against_oracle(tl, simulate.aa_gamma_problem(ns=24, npatt=3000, seed=4))


def test_nuc_gtr_gamma(tl):
    _check_against_oracle(tl, simulate.nuc_gtr_gamma_problem(ns=120, npatt=5000, seed=4))


@pytest.mark.parametrize("maker", [simulate.codon_m0_problem, simulate.aa_gamma_problem, simulate.nuc_gtr_gamma_problem])
def test_ambiguity_codes(tl, maker):
    p = simulate.add_ambiguity(maker(ns=10, npatt=500, seed=6), 0.05)
    _check_against_oracle(tl, p)


def test_zero_weight_patterns_are_skipped(tl):
    p = simulate.codon_m0_problem(npatt=300, seed=8)
    p.fpatt[::3] = 0
    _check_against_oracle(tl, p)


def test_underflow_guard(tl):
    """A pattern whose likelihood underflows to 0: lfun clamps f_h to 1e-80, fx_r to 1e-300."""
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=50, seed=2, K=2)
    p.branch_t[:] = 0.0                        # P = I: any variable pattern has f_h = 0
    _check_against_oracle(tl, p)


def test_shards_add_up(tl):
    """Patterns are independent: the -lnL of pattern shards sums to the whole."""
    p = simulate.codon_m0_problem(npatt=2000, seed=12)
    tl.set_problem(p)
    whole = tl.eval_problem(p)
    parts = 0.0
    for a, b in [(0, 700), (700, 1301), (1301, p.npatt)]:
        tl.set_problem(p, a, b)
        parts += tl.eval_problem(p)
    assert abs(parts - whole) <= 1e-12 * abs(whole)


def _full_size(cfg):
    if cfg == "cfg2":
        return simulate.codon_m0_problem(npatt=100_000, seed=13147)
    if cfg == "cfg3_K4":
        return simulate.codon_sites_problem(ns=32, npatt=50_000, seed=13149, freqs=(0.4, 0.3, 0.2, 0.1), omegas=(0.05, 0.3, 1.0, 2.5))
    if cfg == "cfg3_K3":
        return simulate.codon_sites_problem(ns=32, npatt=50_000, seed=13149, freqs=(0.6, 0.3, 0.1), omegas=(0.1, 1.0, 2.5))
    if cfg == "cfg4":
        return simulate.aa_gamma_problem(ns=64, npatt=200_000, seed=13151, wag=True)
    if cfg == "cfg5":
        return simulate.nuc_gtr_gamma_problem(ns=128, npatt=1_000_000, seed=1234567)
    raise ValueError(cfg)


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3_K3", "cfg3_K4", "cfg4", "cfg5"])
def test_full_size_properties(tl, cfg):
    """Every BASELINE configuration AT ITS STATED SIZE (cfg2: 8 taxa x 1e5 codon patterns; cfg3: 32 taxa x 5e4 codon patterns,
    3 and 4 site classes; cfg4: 64 taxa x 2e5 amino-acid patterns, WAG + 4 gamma classes; cfg5: 128 taxa x 1e6 nucleotide
    patterns, GTR + 8 gamma classes).  The scalar oracle is too slow for a whole evaluation at these sizes in a unit test, so
    the checks are size-independent properties: (i) -lnL equals the weighted sum of the per-pattern values, (ii) doubling every
    weight doubles -lnL exactly, (iii) a random 3 000-pattern subsample agrees with the oracle per pattern AND per class,
    (iv) three ragged shards sum to the whole."""
    from oracle import oracle
    p = _full_size(cfg)
    want = dict(cfg2=100_000, cfg3_K3=50_000, cfg3_K4=50_000, cfg4=200_000, cfg5=1_000_000)[cfg]
    assert p.npatt == want, (cfg, p.npatt)
    tl.set_problem(p)
    v = tl.eval_problem(p)
    lnf = tl.lnf()
    fhK = tl.fhK() if p.nclass > 1 else None
    assert np.isfinite(v) and (lnf < 0).all()
    assert abs(-(lnf * p.fpatt).sum() - v) <= 1e-12 * abs(v)
    q = simulate.Problem(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    q.fpatt = p.fpatt * 2
    tl.set_problem(q)
    assert abs(tl.eval_problem(q) - 2 * v) <= 1e-13 * abs(v)
    idx = np.sort(np.random.default_rng(0).choice(p.npatt, 3000, replace=False))
    s = simulate.Problem(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    s.z = p.z[:, idx].copy(); s.fpatt = p.fpatt[idx].copy(); s.npatt = len(idx)
    r = oracle.lnl(s)
    np.testing.assert_allclose(lnf[idx], r["lnf"], rtol=0, atol=ABS_LNF)
    if fhK is not None:
        _check_fhK(fhK[:, idx], r["fhK"], p.nscale > 0)
    # the subsample's own total against the oracle's: the north-star tolerance is 1e-6 relative, the bar here 1e-11
    tl.set_problem(s)
    vs = tl.eval_problem(s)
    assert abs(vs - r["neg_lnL"]) <= REL_LNL * abs(r["neg_lnL"])
    # ragged shards (boundaries not multiples of the tile sizes) add up to the whole
    cuts = [0, p.npatt // 3 + 5, 2 * p.npatt // 3 - 3, p.npatt]
    parts = 0.0
    for a, b in zip(cuts[:-1], cuts[1:]):
        tl.set_problem(p, a, b)
        parts += tl.eval_problem(p)
    assert abs(parts - v) <= 1e-12 * abs(v)


def test_errors_are_loud(tl):
    p = simulate.codon_m0_problem(npatt=100, seed=1)
    tl.set_problem(p)
    bad = p.branch_set.copy(); bad[0, 0] = 5
    with pytest.raises(PB200Error):
        tl.eval(p.semantics, p.freqK, p.set_kind, p.nr, p.set_data, bad, p.branch_t, p.pi)


# ---- the P-stationary, warp-specialised kernel (csrc/prune_ps.cuh; PB200_KERNEL=ps, read when the tree is set) ----
@pytest.fixture()
def tl_ps(monkeypatch):
    monkeypatch.setenv("PB200_KERNEL", "ps")
    t = TreeLikelihood(0)
    yield t
    t.close()


@pytest.mark.parametrize("path", [g for g in GOLDEN if any(k in g for k in ("lysozyme", "mhc"))], ids=lambda g: g.split("/")[-1])
def test_ps_kernel_on_codon_fixtures(tl_ps, path):
    p = load_fixture(path)
    tl_ps.set_problem(p)
    v = tl_ps.eval_problem(p)
    assert abs(v - p.ref_neg_lnL) <= REL_LNL * abs(p.ref_neg_lnL), (v, p.ref_neg_lnL)
    np.testing.assert_allclose(tl_ps.lnf(), p.ref_lnf, rtol=0, atol=ABS_LNF)
    if p.ref_fhK is not None:
        _check_fhK(tl_ps.fhK(), p.ref_fhK, p.nscale > 0)
    assert tl_ps.eval_problem(p) == v


@pytest.mark.parametrize("npatt", [1, 9, 31, 32, 33, 255, 4097, 20011])
def test_ps_kernel_ragged_sizes(tl_ps, npatt):
    p = simulate.codon_m0_problem(npatt=max(npatt, 64), seed=5)
    p.z = p.z[:, :npatt].copy(); p.fpatt = p.fpatt[:npatt].copy(); p.npatt = npatt
    _check_against_oracle(tl_ps, p) if npatt < 5000 else None
    if npatt >= 5000:
        tl_ps.set_problem(p)
        v = tl_ps.eval_problem(p)
        assert "prune_ps_kernel" in tl_ps.last_kernel()
        ref = TreeLikelihood(0)
        try:
            import os
            os.environ["PB200_KERNEL"] = "warp"
            ref.set_problem(p)
            os.environ["PB200_KERNEL"] = "ps"
            assert abs(ref.eval_problem(p) - v) <= 1e-12 * abs(v)
            np.testing.assert_allclose(tl_ps.lnf(), ref.lnf(), rtol=0, atol=1e-10)
        finally:
            ref.close()


def test_ps_kernel_classes_scaling_ambiguity(tl_ps):
    _check_against_oracle(tl_ps, simulate.codon_sites_problem(ns=16, npatt=700, seed=3))
    p = simulate.codon_m0_problem(ns=40, npatt=300, seed=9)
    assert p.nscale > 0
    _check_against_oracle(tl_ps, p)
    _check_against_oracle(tl_ps, simulate.add_ambiguity(simulate.codon_m0_problem(ns=10, npatt=500, seed=6), 0.05))
    assert "prune_ps_kernel" in tl_ps.last_kernel()


def test_program_in_global_memory(monkeypatch):
    """Programs that do not fit the constant-memory buffer are read from global memory (the kernels' second instantiation):
    forced here for every kernel family."""
    monkeypatch.setenv("PB200_NO_CONST_PROG", "1")
    t = TreeLikelihood(0)
    try:
        _check_against_oracle(t, simulate.codon_m0_problem(npatt=300, seed=5))
        _check_against_oracle(t, simulate.nuc_gtr_gamma_problem(ns=30, npatt=900, seed=4, K=3))
        _check_against_oracle(t, simulate.aa_gamma_problem(ns=12, npatt=300, seed=4))
    finally:
        t.close()
    monkeypatch.setenv("PB200_KERNEL", "ps")
    t = TreeLikelihood(0)
    try:
        _check_against_oracle(t, simulate.codon_sites_problem(ns=12, npatt=150, seed=3))
        assert "prune_ps_kernel<4,false>" in t.last_kernel()
    finally:
        t.close()


def test_many_contexts_take_turns():
    """Six contexts on one device evaluated in turn, as the gene units of a partitioned model or mcmctree's loci are: more contexts
    than constant-memory program slots (least-recently-used eviction), evaluation graphs captured per context and dropped when a
    program changes slot, the shared constant matrix buffer of the 4-state kernel claimed in turn.  Every value is checked against
    the oracle, and every context must return the same bits each round."""
    from oracle import oracle
    probs = [simulate.nuc_gtr_gamma_problem(ns=8 + 3 * i, npatt=400 + 50 * i, seed=40 + i, K=5) for i in range(4)]
    probs += [simulate.codon_m0_problem(npatt=200, seed=8), simulate.aa_gamma_problem(ns=10, npatt=250, seed=9)]
    ctxs = [TreeLikelihood(0) for _ in probs]
    try:
        ref = [oracle.lnl(p)["neg_lnL"] for p in probs]
        for t, p in zip(ctxs, probs):
            t.set_problem(p)
        first = None
        for rnd in range(6):
            vals = [t.eval_problem(p) for t, p in zip(ctxs, probs)]
            for v, r in zip(vals, ref):
                assert abs(v - r) <= REL_LNL * abs(r), (rnd, v, r)
            if first is None:
                first = vals
            assert vals == first, (rnd, vals, first)
        # enqueued on all contexts before any is collected (one host thread driving several contexts, as the shim does)
        for rnd in range(3):
            for t, p in zip(ctxs, probs):
                t.eval_begin_problem(p)
            vals = [t.eval_end() for t in ctxs]
            assert vals == first, (rnd, vals, first)
    finally:
        for t in ctxs:
            t.close()


@pytest.mark.parametrize("ns", [3, 4, 5, 6])
def test_tiny_trees_default_dispatch(tl, ns):
    """Three to six taxa (a star tree at three: the root's children are all tips -- the P-stationary kernel's three-tip gather op and
    its step without a contraction), through the default dispatch, which picks the P-stationary kernel for alignments this small."""
    p = simulate.codon_m0_problem(ns=ns, npatt=150, seed=60 + ns)
    _check_against_oracle(tl, p)
    assert "prune_ps_kernel" in tl.last_kernel()
    _check_against_oracle(tl, simulate.add_ambiguity(simulate.codon_m0_problem(ns=ns, npatt=90, seed=70 + ns), 0.1))
