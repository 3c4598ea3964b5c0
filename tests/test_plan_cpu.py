"""Host logic without a GPU: the planner's program interpreted in numpy must reproduce the oracle."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_ids
from paml_b200.fixture import load_fixture
from paml_b200 import likelihood, simulate
from oracle import oracle
from simprog import run_program, run_program4


def _pmat_fn(p):
    cache = {}
    def f(cls, node):
        key = (cls, node)
        if key not in cache:
            cache[key] = oracle.pmat(p.branch_t[cls, node], p.n, p.set_kind, p.nr, p.set_data[p.branch_set[cls, node]])
        return cache[key]
    return f


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_program_reproduces_reference(path):
    p = load_fixture(path)
    prog, nslots = likelihood.plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
    v, lnf = run_program(prog, nslots, p, _pmat_fn(p))
    assert abs(v - p.ref_neg_lnL) <= 1e-10 * abs(p.ref_neg_lnL)
    np.testing.assert_allclose(lnf, p.ref_lnf, rtol=0, atol=1e-9)


@pytest.mark.parametrize("shape,ns", [("random", 40), ("balanced", 64), ("caterpillar", 30), ("random", 5)])
def test_program_on_synthetic_trees(shape, ns):
    from paml_b200.tree import random_tree
    p = simulate.nuc_gtr_gamma_problem(ns=ns, npatt=300, seed=11 + ns, K=3)
    t = random_tree(ns, 5 + ns, shape=shape)
    p.father, p.nson, p.sons, p.root, p.nnode, p.maxsons = t.father, t.nson, t.sons, t.root, t.nnode, t.maxsons
    p.branch_t = np.array([1.0, 0.3, 2.0])[:, None] * t.branch[None, :]
    p.branch_set = np.zeros((3, t.nnode), np.int32)
    p.nscale, p.scale_flags = likelihood.default_scale_flags(t.nnode, t.root, t.nson, t.sons, 7)
    prog, nslots = likelihood.plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
    # slot need is logarithmic for balanced trees and 1-2 for caterpillars
    assert nslots <= int(np.ceil(np.log2(ns)))
    if shape == "caterpillar":
        assert nslots <= 1
    r = oracle.lnl(p)
    v, lnf = run_program(prog, nslots, p, _pmat_fn(p))
    assert abs(v - r["neg_lnL"]) <= 1e-10 * abs(r["neg_lnL"])


@pytest.mark.parametrize("path", [g for g in GOLDEN if "brown" in g], ids=[i for i in golden_ids() if "brown" in i])
def test_program4_reproduces_reference(path):
    # the 4-state kernel's macro-op program on the nucleotide fixtures (incl. ambiguity codes and the clock tree)
    p = load_fixture(path)
    ops, tips, nslots = likelihood.plan_tree4(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
    v, lnf = run_program4(ops, tips, nslots, p, _pmat_fn(p))
    assert abs(v - p.ref_neg_lnL) <= 1e-10 * abs(p.ref_neg_lnL)
    np.testing.assert_allclose(lnf, p.ref_lnf, rtol=0, atol=1e-9)


@pytest.mark.parametrize("shape,ns", [("random", 40), ("balanced", 64), ("caterpillar", 30), ("random", 5), ("random", 150)])
def test_program4_on_synthetic_trees(shape, ns):
    from paml_b200.tree import random_tree
    p = simulate.nuc_gtr_gamma_problem(ns=ns, npatt=200, seed=21 + ns, K=2)
    t = random_tree(ns, 9 + ns, shape=shape)
    p.father, p.nson, p.sons, p.root, p.nnode, p.maxsons = t.father, t.nson, t.sons, t.root, t.nnode, t.maxsons
    p.branch_t = np.array([0.4, 1.7])[:, None] * t.branch[None, :]
    p.branch_set = np.zeros((2, t.nnode), np.int32)
    p.nscale, p.scale_flags = likelihood.default_scale_flags(t.nnode, t.root, t.nson, t.sons, 7)
    ops, tips, nslots = likelihood.plan_tree4(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
    assert sorted(tips.tolist()) == list(range(ns))
    assert nslots <= max(1, int(np.ceil(np.log2(ns))))
    if shape == "caterpillar":
        assert nslots <= 1
    # one macro-op per internal node, plus the parks, plus extra tips of multifurcations, plus ROOT
    assert len(ops) <= 2 * (t.nnode - ns) + 2
    r = oracle.lnl(p)
    v, lnf = run_program4(ops, tips, nslots, p, _pmat_fn(p))
    assert abs(v - r["neg_lnL"]) <= 1e-10 * abs(r["neg_lnL"])


def test_program4_multifurcations():
    from paml_b200.tree import parse_newick
    for nwk in ["((1:.1,2:.1):.1,(3:.1,4:.1):.1,(5:.1,6:.1):.1,(7:.1,8:.1):.1,(9:.1,10:.1):.1,11:.2);",
                "(1:.1,2:.2,3:.3,4:.1,(5:.2,6:.1,7:.3):.2);", "(1:.1,2:.2);", "((1:.1,2:.2):.1,3:.2,(4:.1,(5:.1,6:.2):.1):.3,7:.1);"]:
        t = parse_newick(nwk)
        p = simulate.nuc_gtr_gamma_problem(ns=max(t.ns, 4), npatt=120, seed=5, K=2)
        p.z, p.ns = p.z[:t.ns], t.ns
        p.father, p.nson, p.sons, p.root, p.nnode, p.maxsons = t.father, t.nson, t.sons, t.root, t.nnode, t.maxsons
        p.branch_t = np.array([0.4, 1.7])[:, None] * t.branch[None, :]
        p.branch_set = np.zeros((2, t.nnode), np.int32)
        p.nscale, p.scale_flags = 0, np.zeros(t.nnode, np.uint8)
        ops, tips, nslots = likelihood.plan_tree4(p.nnode, p.ns, p.root, p.nson, p.sons, None)
        assert nslots == 1
        r = oracle.lnl(p)
        v, lnf = run_program4(ops, tips, nslots, p, _pmat_fn(p))
        assert abs(v - r["neg_lnL"]) <= 1e-10 * abs(r["neg_lnL"])


def test_multifurcation_is_folded():
    # star-like root with 5 internal children must not need 5 slots
    from paml_b200.tree import parse_newick
    t = parse_newick("((1:.1,2:.1):.1,(3:.1,4:.1):.1,(5:.1,6:.1):.1,(7:.1,8:.1):.1,(9:.1,10:.1):.1,11:.2);")
    prog, nslots = likelihood.plan_tree(t.nnode, t.ns, t.root, t.nson, t.sons, None)
    assert nslots == 1


def test_plan_rejects_bad_trees():
    nson = np.array([0, 0, 2], np.int32)
    sons = np.array([[-1, -1], [-1, -1], [0, 0]], np.int32)
    with pytest.raises(likelihood.PB200Error):
        likelihood.plan_tree(3, 2, 2, nson, sons, None)
    with pytest.raises(likelihood.PB200Error):
        likelihood.plan_tree(3, 2, 0, nson, sons, None)      # root must be internal


def test_default_scale_flags_match_reference():
    for g in GOLDEN:
        p = load_fixture(g)
        every = {0: 100, 1: 15, 2: 50}[p.seqtype]
        k, flags = likelihood.default_scale_flags(p.nnode, p.root, p.nson, p.sons, every)
        assert k == p.nscale
        np.testing.assert_array_equal(flags, p.scale_flags)


@pytest.mark.parametrize("path", [g for g in GOLDEN if any(k in g for k in ("lysozyme_m0", "lysozyme_m2a", "mhc_m0", "lysozyme_freeratio"))],
                         ids=lambda g: g.split("/")[-1])
def test_program_ps_reproduces_reference(path):
    # the P-stationary kernel's gather-op list + step table (codon fixtures incl. the 192-taxon MHC tree with scaling nodes)
    from simprog import run_program_ps
    p = load_fixture(path)
    try:
        prog, _ = likelihood.plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags, ps=True)
    except likelihood.PB200Error as e:
        assert "scaling node whose children are all tips" in str(e)
        return
    v, lnf = run_program_ps(prog, p, _pmat_fn(p))
    assert abs(v - p.ref_neg_lnL) <= 1e-10 * abs(p.ref_neg_lnL)
    np.testing.assert_allclose(lnf, p.ref_lnf, rtol=0, atol=1e-9)


@pytest.mark.parametrize("shape,ns", [("random", 40), ("balanced", 64), ("caterpillar", 30), ("random", 5), ("star", 6)])
def test_program_ps_on_synthetic_trees(shape, ns):
    from paml_b200.tree import random_tree
    from simprog import run_program_ps
    p = simulate.nuc_gtr_gamma_problem(ns=ns, npatt=200, seed=21 + ns, K=2)
    if shape == "star":
        nnode = ns + 1
        p.nson = np.zeros(nnode, np.int32); p.nson[ns] = ns
        p.sons = np.zeros((nnode, ns), np.int32); p.sons[ns, :] = np.arange(ns)
        p.father = np.full(nnode, ns, np.int32); p.father[ns] = -1
        p.root, p.nnode, p.maxsons = ns, nnode, ns
        branch = np.linspace(0.05, 0.4, nnode)
        p.nscale, p.scale_flags = 0, np.zeros(nnode, np.uint8)
    else:
        t = random_tree(ns, 7 + ns, shape=shape)
        p.father, p.nson, p.sons, p.root, p.nnode, p.maxsons = t.father, t.nson, t.sons, t.root, t.nnode, t.maxsons
        branch = t.branch
        p.nscale, p.scale_flags = likelihood.default_scale_flags(t.nnode, t.root, t.nson, t.sons, 7)
    p.branch_t = np.array([1.0, 0.4])[:, None] * branch[None, :]
    p.branch_set = np.zeros((2, p.nnode), np.int32)
    prog, _ = likelihood.plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags, ps=True)
    r = oracle.lnl(p)
    v, lnf = run_program_ps(prog, p, _pmat_fn(p))
    assert abs(v - r["neg_lnL"]) <= 1e-10 * abs(r["neg_lnL"])


def _random_multifurcating_newick(ns, rng, pmulti=0.35):
    """Random rooted tree as a Newick string in which a join takes 2, 3 or 4 lineages (multifurcations at any depth)."""
    items = ["%d:%.5f" % (i + 1, rng.uniform(0.02, 0.3)) for i in range(ns)]
    while len(items) > 4:
        k = 2 + (rng.random() < pmulti) + (rng.random() < pmulti / 2)
        idx = sorted(rng.choice(len(items), min(k, len(items) - 1), replace=False))
        grp = [items[i] for i in idx]
        items = [x for i, x in enumerate(items) if i not in idx] + ["(%s):%.5f" % (",".join(grp), rng.uniform(0.02, 0.3))]
    return "(" + ",".join(items) + ");"


@pytest.mark.parametrize("seed", range(12))
def test_programs_on_random_multifurcating_trees(seed):
    """Both codon planners (the warp-independent kernel's op stream with folded multifurcations, and the P-stationary kernel's
    gather ops + step table) on random trees with multifurcations at any depth and random scaling nodes: interpreted in numpy,
    they must give the oracle's value, and the interpreters' liveness assertions (no slot overwritten while live, G slots rewritten
    only after they were read, k fields = writes per chunk) must hold."""
    from paml_b200.tree import parse_newick
    from simprog import run_program_ps
    rng = np.random.default_rng(1000 + seed)
    ns = int(rng.integers(4, 40))
    p = simulate.nuc_gtr_gamma_problem(ns=ns, npatt=120, seed=50 + seed, K=2)
    t = parse_newick(_random_multifurcating_newick(ns, rng))
    p.father, p.nson, p.sons, p.root, p.nnode, p.maxsons = t.father, t.nson, t.sons, t.root, t.nnode, t.maxsons
    p.branch_t = np.array([1.0, 0.5])[:, None] * t.branch[None, :]
    p.branch_set = np.zeros((2, t.nnode), np.int32)
    flags = np.zeros(t.nnode, np.uint8)
    internal = [v for v in range(ns, t.nnode) if v != t.root and any(t.nson[c] > 0 for c in t.sons[v, :t.nson[v]])]
    for v in internal:
        if rng.random() < 0.3:
            flags[v] = 1                       # (nodes whose children are all tips stay unflagged: the P-stationary planner refuses those)
    p.scale_flags, p.nscale = flags, int(flags.sum())
    r = oracle.lnl(p)
    prog, nslots = likelihood.plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
    v, _ = run_program(prog, nslots, p, _pmat_fn(p))
    assert abs(v - r["neg_lnL"]) <= 1e-10 * abs(r["neg_lnL"])
    progps, _ = likelihood.plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags, ps=True)
    v2, _ = run_program_ps(progps, p, _pmat_fn(p))
    assert abs(v2 - r["neg_lnL"]) <= 1e-10 * abs(r["neg_lnL"])
