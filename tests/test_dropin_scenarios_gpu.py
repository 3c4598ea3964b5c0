"""The drop-in scenarios of tests/test_host_logic_cpu.py (written for the CPU mock of the C ABI) run with the REAL library on
the GPU: the callers around the hot path -- standard errors (Hessian), rates for sites, marginal / joint ancestral
reconstruction, branch and branch-site models with both optimisers, clocks, amino-acid models, ambiguity data with two trees,
two data sets, and the direct `ConditionalPNode(tree.root, ..)` callers (branch-site BEB, AncestralSeqs, PostProbNode) that go
through pb200_cond_update / pb200_cond_get (on by default since round 2; PB200_CPN=0 would keep the reference's CPU routine).
Every scenario runs the drop-in binary beside the unmodified reference binary on the same files."""
import os

import pytest

import test_dropin_models_gpu as T
import test_host_logic_cpu as H

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(REPO, "paml_b200", "bin")


@pytest.fixture()
def shim():
    for exe in ("codeml_b200", "baseml_b200"):
        if not os.path.exists(os.path.join(BIN, exe)):
            pytest.skip("prebuilt drop-in binary missing: " + exe)
    for exe in ("codeml", "baseml"):
        if not os.path.exists(os.path.join(REPO, "oracle", "_ref", exe)):
            pytest.skip("reference binary missing: oracle/_ref/" + exe)
    return BIN


test_standard_errors_hessian = H.test_standard_errors_hessian
test_rates_for_sites_and_ancestors_method0 = H.test_rates_for_sites_and_ancestors_method0
test_labelled_branches = H.test_labelled_branches
test_global_clock = H.test_global_clock
test_amino_acid_models = H.test_amino_acid_models
test_ambiguous_data_and_two_trees = H.test_ambiguous_data_and_two_trees
test_two_data_sets = H.test_two_data_sets
test_whole_tree_conditional_p_node_through_the_abi = H.test_whole_tree_conditional_p_node_through_the_abi


@pytest.mark.parametrize("which", ["beb_branch_site_A", "m0_ancestors_joint"])
def test_cpn_routing_is_really_used(shim, tmp_path, monkeypatch, which):
    """The same BEB / ancestral-reconstruction scenario with the routing switched off (PB200_CPN=0: the reference's CPU
    ConditionalPNode) must print the same tables -- and with it on, the shim must say that the calls went to the device."""
    import subprocess
    H.test_whole_tree_conditional_p_node_through_the_abi(shim, tmp_path, monkeypatch, which)
    d = str(tmp_path / "gpu")
    on = T._numbers(d + "/rst")
    r = subprocess.run([os.path.join(BIN, "codeml_b200"), "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900, env=dict(T._same_start_env(os.path.join(BIN, "codeml_b200")), PB200_VERBOSE="1"))
    assert r.returncode == 0 and "whole-tree ConditionalPNode calls go through pb200_cond_update" in r.stdout, r.stdout[-1500:]
    r = subprocess.run([os.path.join(BIN, "codeml_b200"), "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900, env=dict(T._same_start_env(os.path.join(BIN, "codeml_b200")), PB200_CPN="0"))
    assert r.returncode == 0, r.stdout[-1500:]
    off = T._numbers(d + "/rst")
    assert len(on) == len(off)
    bad = [(a, b) for (a, da), (b, db) in zip(on, off) if abs(a - b) > 1e-2 + 2.5 * 10.0 ** (-min(da, db))]
    assert not bad, bad[:10]
