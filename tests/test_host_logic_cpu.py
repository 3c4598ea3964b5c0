"""The HOST side of the drop-in without a GPU.

paml_b200/host/plfun_b200.inc + pb200_gather.inc (the C shim that re-defines lfun, lfundG, fx_r, PatternWeight, eigenQREV and
the minbranches family on top of the C ABI) is ~1000 lines of model mapping: which spectral set and t every (class, branch)
uses, gene units, the eigen queue and memo, the bookkeeping PatternWeight leaves behind, the method = 1 dispatch.  Here the
SAME wrapper translation units are linked against tests/mock/pb200_mock.c -- the C ABI answered on the CPU by the oracle --
and every drop-in scenario of tests/test_dropin_gpu.py / tests/test_dropin_models_gpu.py is run with those binaries beside
the unmodified reference.  This checks the shim (and the ABI contract it relies on); the CUDA kernels behind the real library
are checked by the same scenarios on the GPU.  Needs the reference sources (authoring container); skipped elsewhere."""
import os
import subprocess

import pytest

import test_dropin_gpu as D
import test_dropin_models_gpu as T

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(REPO, "tests", "mock", "_build")
REFSRC = os.environ.get("PAML_REFERENCE", "/root/reference") + "/src"


@pytest.fixture(scope="module")
def mock_bin():
    if not os.path.exists(os.path.join(REFSRC, "codeml.c")):
        pytest.skip("reference sources not present: the mock drop-in cannot be built here")
    for exe in ("codeml", "baseml"):
        if not os.path.exists(os.path.join(REPO, "oracle", "_ref", exe)):
            pytest.skip("reference binary not built (make -C oracle ref)")
    r = subprocess.run(["make", "-C", os.path.join(REPO, "tests", "mock"), "REF=" + REFSRC], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    for prog in ("codeml", "baseml"):                       # the scenarios look for <BIN>/<prog>_b200
        link = os.path.join(MOCK, prog + "_b200")
        if not os.path.exists(link):
            os.symlink(prog + "_mock", link)
    return MOCK


@pytest.fixture()
def shim(mock_bin, monkeypatch):
    monkeypatch.setattr(T, "BIN", mock_bin)
    monkeypatch.setattr(D, "BIN", mock_bin)
    return mock_bin


def test_m0_fit(shim, tmp_path):
    D.test_codeml_m0_fit(tmp_path)


def test_m2a_fixed_point(shim, tmp_path):
    D.test_codeml_sites_m2a_fixed_point(tmp_path)


def test_gtr_gamma_fit(shim, tmp_path):
    D.test_baseml_gtr_gamma_fit(tmp_path)


@pytest.mark.parametrize("mgene,malpha", [(0, 0), (2, 0), (3, 0), (4, 1)])
def test_baseml_genes(shim, tmp_path, mgene, malpha):
    T.test_baseml_genes(tmp_path, mgene, malpha)


@pytest.mark.parametrize("nhomo", [2, 3, 4, 5])
def test_baseml_nhomo(shim, tmp_path, nhomo):
    T.test_baseml_nhomo(tmp_path, nhomo)


def test_baseml_unrest(shim, tmp_path):
    T.test_baseml_unrest(tmp_path)


def test_auto_discrete_gamma(shim, tmp_path):
    T.test_baseml_auto_discrete_gamma(tmp_path)


@pytest.mark.parametrize("mgene", [0, 3])
def test_codeml_genes(shim, tmp_path, mgene):
    T.test_codeml_genes(tmp_path, mgene)


def test_free_ratio(shim, tmp_path):
    T.test_codeml_free_ratio_fixed_point(tmp_path)


def test_m3_five_classes_eigen_queue(shim, tmp_path):
    T.test_codeml_m3_five_classes_fixed_point(tmp_path)


def test_m2a_neb_beb(shim, tmp_path):
    T.test_codeml_m2a_neb_beb_fixed_point(tmp_path)


def test_method1_m0(shim, tmp_path):
    T.test_method1_codeml_m0(tmp_path)


def test_method1_m2a(shim, tmp_path):
    T.test_method1_codeml_m2a_site_classes(tmp_path)


def test_method1_hky_gamma(shim, tmp_path):
    T.test_method1_baseml_hky_gamma(tmp_path)


@pytest.mark.parametrize("model,nhomo", [(1, 0), (0, 0), (4, 2), (1, 2)])
def test_method1_k80_and_kappa_per_branch(shim, tmp_path, model, nhomo):
    T.test_method1_baseml_k80_and_kappa_per_branch(tmp_path, model, nhomo)


def test_method1_free_ratio(shim, tmp_path):
    T.test_method1_codeml_free_ratio(tmp_path)


def test_method1_cpu_minbranches_paths(shim, tmp_path):
    T.test_method1_cpu_minbranches_paths_get_fresh_host_vectors(tmp_path)


def test_method1_ancestral(shim, tmp_path):
    T.test_method1_ancestral_reconstruction_with_site_classes(tmp_path)


def test_too_many_genes_fails_loudly(shim, tmp_path):
    T.test_too_many_genes_fails_loudly(tmp_path)
