"""The CPU oracle (oracle/paml_oracle.c) against the golden fixtures dumped from the unmodified reference:
total -lnL, per-pattern log f_h, fhK and P(t) must agree to the last bit or two (same algorithm, same
summation order).  This is what pins the oracle; everything GPU is then checked against oracle + fixtures."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_ids
from paml_b200.fixture import load_fixture
from oracle import oracle

KNOWN = {  # documented / reproduced lnL of the reference examples
    "brown_hky85": -2665.422858, "brown_rev_g8": -2619.320596, "lysozyme_m0": -906.017440,
    "stewart_wag_g4": -1028.924417, "mhc_m0": -8225.154790, "mhc_m2a": -7231.154540,
}


@pytest.mark.parametrize("path", GOLDEN, ids=golden_ids())
def test_oracle_matches_reference(path):
    p = load_fixture(path)
    r = oracle.lnl(p)
    assert abs(r["neg_lnL"] - p.ref_neg_lnL) <= 1e-12 * abs(p.ref_neg_lnL)
    np.testing.assert_allclose(r["lnf"], p.ref_lnf, rtol=0, atol=1e-11)
    if p.ref_fhK is not None:
        np.testing.assert_allclose(r["fhK"], p.ref_fhK, rtol=1e-13, atol=1e-300)
    for cls, node, P in p.ref_P:
        Po = oracle.pmat(p.branch_t[cls, node], p.n, p.set_kind, p.nr, p.set_data[p.branch_set[cls, node]])
        np.testing.assert_allclose(Po, P, rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_fixture_is_the_known_answer(name):
    path = [g for g in GOLDEN if g.endswith("/" + name + ".pbfx")]
    assert path, "fixture %s missing" % name
    p = load_fixture(path[0])
    assert abs(-p.ref_neg_lnL - KNOWN[name]) < 2e-6 * abs(KNOWN[name]) + 1e-5


def test_golden_covers_the_cases():
    ps = [load_fixture(g) for g in GOLDEN]
    assert {p.n for p in ps} >= {4, 20, 61}
    assert any(p.cleandata == 0 for p in ps), "an ambiguity fixture is required"
    assert any(p.nscale > 0 for p in ps), "a fixture with scaling nodes is required"
    assert any(p.nset > 1 for p in ps) and any(p.nclass > 1 for p in ps)
    assert {p.set_kind for p in ps} >= {0, 1, 2, 3}


def test_scale_flags_restatement():
    for g in GOLDEN:
        p = load_fixture(g)
        every = {0: 100, 1: 15, 2: 50}[p.seqtype]
        k, flags = oracle.set_node_scale(p, every)
        assert k == p.nscale
        np.testing.assert_array_equal(flags, p.scale_flags)


def test_pattern_order_is_ascending():
    """Reference patterns are in ascending strcmp order of the raw column; for clean nucleotide
    data coded TCAG that is ascending order of the letters A<C<G<T, i.e. of codes mapped 2,1,3,0 -> rank."""
    p = load_fixture([g for g in GOLDEN if g.endswith("brown_hky85.pbfx")][0])
    rank = np.array([3, 1, 0, 2])          # T C A G -> rank of the letter
    cols = rank[p.z].T
    keys = [tuple(c) for c in cols]
    assert keys == sorted(keys)
