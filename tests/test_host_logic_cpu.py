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


@pytest.mark.parametrize("mgene,malpha,gamma", [(0, 0, False), (2, 0, False), (3, 0, True), (4, 1, True)])
def test_method1_baseml_genes(shim, tmp_path, mgene, malpha, gamma):
    T.test_method1_baseml_genes(tmp_path, mgene, malpha, gamma)


def test_method1_codeml_genes(shim, tmp_path):
    T.test_method1_codeml_genes(tmp_path)


def test_method1_aaml_poisson(shim, tmp_path):
    T.test_method1_aaml_poisson(tmp_path)


def test_method1_cpu_minbranches_paths(shim, tmp_path):
    T.test_method1_cpu_minbranches_paths_get_fresh_host_vectors(tmp_path)


def test_method1_ancestral(shim, tmp_path):
    T.test_method1_ancestral_reconstruction_with_site_classes(tmp_path)


def test_too_many_genes_fails_loudly(shim, tmp_path):
    T.test_too_many_genes_fails_loudly(tmp_path)


# ---- scenarios that exist only here (callers of com.plfun / fx_r whose GPU side is nothing but the evaluations above) ----

def _run_pair(shim, tmp_path, prog, write):
    out = {}
    for tag, exe in (("gpu", os.path.join(shim, prog + "_b200")), ("ref", os.path.join(REPO, "oracle", "_ref", prog))):
        d = tmp_path / tag
        d.mkdir()
        write(str(d))
        T._run(exe, str(d))
        out[tag] = str(d)
    return out


def test_standard_errors_hessian(shim, tmp_path):
    """getSE = 1: the Hessian by second differences of com.plfun (HessianSKT2004 / Hessian, src/treesub.c:7241, src/tools.c);
    the standard errors printed in the main result file must agree with the reference's."""
    import re
    from paml_b200 import paml_io, simulate
    p = simulate.codon_m0_problem(ns=6, npatt=250, seed=81)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, getSE=1))
    dirs = _run_pair(shim, tmp_path, "codeml", write)
    se = {}
    for tag, d in dirs.items():
        txt = open(d + "/mlc").read()
        m = re.search(r"SEs for parameters:\s*\n([^\n]+)\n", txt)
        assert m, "no SE block in " + d
        se[tag] = [float(v) for v in m.group(1).split()]
    assert len(se["gpu"]) == len(se["ref"]) == 11
    for a, b in zip(se["gpu"], se["ref"]):
        assert abs(a - b) <= 1e-2 * max(abs(b), 1e-2), (se["gpu"], se["ref"])     # two fits from random starts


def test_rates_for_sites_and_ancestors_method0(shim, tmp_path):
    """RateAncestor = 1 with gamma rates under method = 0: lfunRates reads com.fhK as left by fx_r (the GPU path), the
    marginal reconstruction runs the reference's CPU ConditionalPNode; `rates` and `rst` must agree with the reference's."""
    from paml_b200 import paml_io, simulate
    p = simulate.nuc_gtr_gamma_problem(ns=8, npatt=300, seed=25, K=4)

    def write(d):
        paml_io.write_sequence_files(p, d + "/s")
        with open(d + "/run.ctl", "w") as f:
            f.write((T.BASEML_CTL % dict(Mgene=0, Malpha=0, model=4, fix_alpha=0, alpha="0.5", ncatG=4, nhomo=0, more=""))
                    .replace("RateAncestor = 0", "RateAncestor = 1"))
    dirs = _run_pair(shim, tmp_path, "baseml", write)
    assert abs(T._lnl(dirs["gpu"] + "/mlb") - T._lnl(dirs["ref"] + "/mlb")) <= 2e-6 * abs(T._lnl(dirs["ref"] + "/mlb"))
    for fn, least in (("rates", 300), ("rst", 500)):
        g, r = T._numbers(dirs["gpu"] + "/" + fn), T._numbers(dirs["ref"] + "/" + fn)
        assert len(g) == len(r) and len(r) > least, (fn, len(g), len(r))
        bad = [(a, b) for (a, da), (b, db) in zip(g, r) if abs(a - b) > 1e-2 + 2.5 * 10.0 ** (-min(da, db))]
        assert not bad, (fn, bad[:10])


@pytest.mark.parametrize("tag,extra,method", [("branch", dict(model=2), 0), ("branch", dict(model=2), 1),
                                             ("branch_site_A", dict(model=2, NSsites=2, ncatG=3), 0),
                                             ("branch_site_A", dict(model=2, NSsites=2, ncatG=3), 1)])
def test_labelled_branches(shim, tmp_path, tag, extra, method):
    """Branch model (model = 2: omega per branch label, the _UU[] sets) and branch-site model A (Set_UVR_BranchSite, Qfactor per
    label), with the simultaneous optimiser and with method = 1 (where lfuntdd[_SiteClass] pick the set by the label of the
    branch under iteration): a foreground clade marked #1 in the tree file."""
    from paml_b200 import paml_io, simulate
    p = simulate.codon_m0_problem(ns=8, npatt=(60 if "NSsites" in extra else 200), seed=45)

    def write(d, m=method):
        paml_io.write_pattern_files(p, d + "/s")
        tre = open(d + "/s.tre").read()
        i = tre.index("):")
        open(d + "/s.tre", "w").write(tre[:i + 1] + " #1 " + tre[i + 1:])
        fixed_point = (tag == "branch_site_A" and m == 0)     # one evaluation + NEB + BEB; the fit itself is what method = 1 covers
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=(2 if fixed_point else 1), method=m, **extra))
        if fixed_point:
            with open(d + "/in.codeml", "w") as f:
                f.write("-1  2.0  0.5 0.3  0.1 3.0\n")                 # kappa, p0 p1, w0 w2
    res = T._both(tmp_path, "codeml", write, "mlc")
    if method == 1:
        T._close_method1(res, "codeml", lambda d: write(d, 0), "mlc", tmp_path, rel=5e-6)
    else:
        T._close(res, 5e-6)


@pytest.mark.parametrize("prog", ["baseml", "codeml"])
def test_global_clock(shim, tmp_path, prog):
    """clock = 1 on a rooted tree: node ages are the parameters, t of every branch comes from SetBranch + GetBranchRate
    (src/treesub.c:3682, 3770-3811) -- the shim reads nodes[].branch after the reference's SetParameters."""
    from paml_b200 import paml_io, simulate
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=200, seed=33, K=1) if prog == "baseml" else simulate.codon_m0_problem(ns=6, npatt=150, seed=41)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s")
        open(d + "/s.tre", "w").write("6 1\n((1,2),((3,4),(5,6)));\n")
        if prog == "baseml":
            with open(d + "/run.ctl", "w") as f:
                f.write((T.BASEML_CTL % dict(Mgene=0, Malpha=0, model=4, fix_alpha=1, alpha="0", ncatG=1, nhomo=0, more=""))
                        .replace("clock = 0", "clock = 1").replace("fix_blength = 1", "fix_blength = 0"))
        else:
            paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                     extra=dict(fix_kappa=0, fix_omega=0, fix_blength=0, clock=1))
    T._close(T._both(tmp_path, prog, write, "mlb" if prog == "baseml" else "mlc"), 5e-6)


def test_amino_acid_models(shim, tmp_path):
    """aaml: Poisson (closed form, JC69-like) and an empirical matrix + gamma (eigenQaa once, src/codeml.c:2382-2383)."""
    from paml_b200 import paml_io, simulate
    wag = os.path.join(os.path.dirname(REFSRC), "dat", "wag.dat")
    if not os.path.exists(wag):                          # GPU box: the copy that travels with the reference binaries
        wag = os.path.join(REPO, "oracle", "_ref", "dat", "wag.dat")
    p = simulate.aa_gamma_problem(ns=7, npatt=150, seed=51, K=4)
    for tag, extra in (("poisson", dict(model=0, fix_alpha=1, alpha=0, ncatG=1)),
                       ("wag_gamma", dict(model=2, aaRatefile=wag, fix_alpha=0, alpha=0.5, ncatG=4))):
        if tag == "wag_gamma" and not os.path.exists(wag):
            continue

        def write(d):
            paml_io.write_pattern_files(p, d + "/s")
            paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", seqtype=2, extra=dict(fix_blength=1, **extra))
        (tmp_path / tag).mkdir()
        T._close(T._both(tmp_path / tag, "codeml", write, "mlc"), 2e-6)


def test_ambiguous_data_and_two_trees(shim, tmp_path):
    """cleandata = 0 with gaps and ambiguity characters in the alignment (nChara / CharaMap), and a tree file with two trees:
    the second tree must trigger a new upload of the topology (the shim's fingerprint of the static data)."""
    import numpy as np
    from paml_b200 import paml_io, simulate
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=200, seed=35, K=4)

    def write(d):
        ls = paml_io.write_sequence_files(p, d + "/s")
        rng = np.random.default_rng(9)
        lines = open(d + "/s.phy").read().splitlines()
        out = [lines[0]]
        for ln in lines[1:]:
            name, seq = ln.split(None, 1)
            seq = list(seq.strip())
            for i in rng.choice(len(seq), len(seq) // 25, replace=False):
                seq[i] = "-NRY?"[int(rng.integers(0, 5))]
            out.append("%s  %s" % (name, "".join(seq)))
        open(d + "/s.phy", "w").write("\n".join(out) + "\n")
        open(d + "/s.tre", "w").write("6 2\n((1,2),(3,4),(5,6));\n((1,3),(2,4),(5,6));\n")
        with open(d + "/run.ctl", "w") as f:
            f.write((T.BASEML_CTL % dict(Mgene=0, Malpha=0, model=4, fix_alpha=1, alpha="0", ncatG=1, nhomo=0, more=""))
                    .replace("cleandata = 1", "cleandata = 0").replace("fix_blength = 1", "fix_blength = 0"))
    import re
    vals = {}
    for tag, exe in T._pair("baseml"):
        d = tmp_path / tag
        d.mkdir()
        write(str(d))
        T._run(exe, str(d))
        vals[tag] = [float(v) for v in re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(d / "mlb").read())]
    assert len(vals["gpu"]) == len(vals["ref"]) == 2, vals
    for a, b in zip(vals["gpu"], vals["ref"]):
        assert abs(a - b) <= 2e-6 * abs(b), vals


def test_two_data_sets(shim, tmp_path):
    """ndata = 2: the driver re-reads a second alignment into the same globals; the shim must notice and upload it."""
    import re
    from paml_b200 import paml_io, simulate
    p1 = simulate.codon_m0_problem(ns=6, npatt=120, seed=61)
    p2 = simulate.codon_m0_problem(ns=6, npatt=120, seed=62)
    vals = {}
    for tag, exe in T._pair("codeml"):
        d = tmp_path / tag
        d.mkdir()
        paml_io.write_sequence_files(p1, str(d / "a"))
        paml_io.write_sequence_files(p2, str(d / "b"))
        open(d / "s.phy", "w").write(open(d / "a.phy").read() + "\n" + open(d / "b.phy").read())
        paml_io.write_codeml_ctl(str(d / "run.ctl"), "s.phy", "a.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, ndata=2))
        T._run(exe, str(d))
        vals[tag] = [float(v) for v in re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(d / "mlc").read())]
    assert len(vals["gpu"]) == len(vals["ref"]) == 2 and vals["ref"][0] != vals["ref"][1], vals
    for a, b in zip(vals["gpu"], vals["ref"]):
        assert abs(a - b) <= 2e-6 * abs(b), vals


@pytest.mark.parametrize("which", ["beb_branch_site_A", "ancestors_method0", "ancestors_method1", "cpu_minbranches_refresh", "m0_ancestors_joint"])
def test_whole_tree_conditional_p_node_through_the_abi(shim, tmp_path, monkeypatch, which):
    """PB200_CPN=1: `ConditionalPNode(tree.root, ..)` called directly by BEB for branch-site models, by the ancestral
    reconstruction and by the reference's own lfun / fx_r (host-vector refresh) is answered through pb200_cond_update +
    pb200_cond_get, the vectors and scale factors copied back to nodes[].conP / com.nodeScaleF."""
    monkeypatch.setenv("PB200_CPN", "1")
    if which == "beb_branch_site_A":
        test_labelled_branches(shim, tmp_path, "branch_site_A", dict(model=2, NSsites=2, ncatG=3), 0)
        assert "Bayes Empirical Bayes" in open(tmp_path / "gpu" / "rst").read()
        g, r = T._numbers(str(tmp_path / "gpu" / "rst")), T._numbers(str(tmp_path / "ref" / "rst"))
        assert len(g) == len(r) and len(r) > 300
        bad = [(a, b) for (a, da), (b, db) in zip(g, r) if abs(a - b) > 1e-2 + 2.5 * 10.0 ** (-min(da, db))]
        assert not bad, bad[:10]
    elif which == "ancestors_method0":
        test_rates_for_sites_and_ancestors_method0(shim, tmp_path)
    elif which == "ancestors_method1":
        T.test_method1_ancestral_reconstruction_with_site_classes(tmp_path)
    elif which == "cpu_minbranches_refresh":
        T.test_method1_cpu_minbranches_paths_get_fresh_host_vectors(tmp_path)
    else:
        from paml_b200 import paml_io, simulate
        p = simulate.codon_m0_problem(ns=6, npatt=150, seed=83)

        def write(d):
            paml_io.write_sequence_files(p, d + "/s")
            paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                     extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, RateAncestor=1))
        dirs = _run_pair(shim, tmp_path, "codeml", write)
        g, r = T._numbers(dirs["gpu"] + "/rst"), T._numbers(dirs["ref"] + "/rst")
        assert len(g) == len(r) and len(r) > 300, (len(g), len(r))
        bad = [(a, b) for (a, da), (b, db) in zip(g, r) if abs(a - b) > 1e-2 + 2.5 * 10.0 ** (-min(da, db))]
        assert not bad, bad[:10]


ZOO = [
    ("codeml", "M7_beta", dict(NSsites=7, ncatG=5), False),
    ("codeml", "M8_beta_w_BEB", dict(NSsites=8, ncatG=5), False),
    ("codeml", "clade_model_C", dict(model=3, NSsites=2, ncatG=3), True),
    ("codeml", "hkyREV", dict(hkyREV=1), False),
    ("codeml", "codonfreq_F1x4", dict(CodonFreq=1), False),
    ("codeml", "codonfreq_table", dict(CodonFreq=3), False),
    ("codeml", "equal_codon_freqs", dict(CodonFreq=0), False),
    ("codeml", "mito_code_60_states", dict(icode=1), False),
    ("codeml", "m0_gamma_rates", dict(fix_alpha=0, alpha=0.7, ncatG=3), False),
    ("baseml", "F81", dict(model=2), False),
    ("baseml", "F84", dict(model=3), False),
    ("baseml", "T92", dict(model=5), False),
    ("baseml", "TN93_gamma", dict(model=6, fix_alpha=0, alpha="0.5", ncatG=4), False),
    ("baseml", "JC69_gamma", dict(model=0, fix_alpha=0, alpha="0.5", ncatG=4), False),
    ("baseml", "REV_estimated_pi", dict(model=7, nhomo=1), False),
]


@pytest.mark.parametrize("prog,tag,extra,labelled", ZOO, ids=[z[1] for z in ZOO])
def test_model_zoo(shim, tmp_path, prog, tag, extra, labelled):
    """Short fits over the model switches that change what SetParameters leaves in the globals the shim reads (codon frequency
    models, genetic codes with n != 61, kappa parameterisations, beta / clade site classes, the TN93 family): the drop-in and the
    reference must end at the same lnL."""
    from paml_b200 import paml_io, simulate
    if prog == "codeml":
        p = simulate.codon_m0_problem(ns=5, npatt=60, seed=91)
    else:
        p = simulate.nuc_gtr_gamma_problem(ns=7, npatt=200, seed=92, K=4)

    def write(d):
        paml_io.write_sequence_files(p, d + "/s")
        if labelled:
            tre = open(d + "/s.tre").read()
            i = tre.index("):")
            open(d + "/s.tre", "w").write(tre[:i + 1] + " #1 " + tre[i + 1:])
        if prog == "codeml":
            paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                     extra=dict(dict(fix_kappa=0, fix_omega=0, fix_blength=1), **extra))
        else:
            kw = dict(Mgene=0, Malpha=0, model=4, fix_alpha=1, alpha="0", ncatG=1, nhomo=0, more="")
            kw.update(extra)
            with open(d + "/run.ctl", "w") as f:
                f.write(T.BASEML_CTL % kw)
    T._close(T._both(tmp_path, prog, write, "mlc" if prog == "codeml" else "mlb"), 1e-5)


@pytest.mark.parametrize("which", ["m0_fit", "m2a_neb_beb", "method1_m0", "genes", "whole_tree_cpn"])
def test_pattern_shards_over_several_contexts(shim, tmp_path, monkeypatch, which):
    """PB200_NGPU=3: the shim splits the patterns into contiguous shards (boundaries on multiples of 8), one context each, adds the
    shards' -lnL (and, under method = 1, the derivatives of every Newton step), and reassembles per-pattern outputs (lnf, fhK,
    resident vectors) in pattern order.  The mock ignores the device ordinal, so the bookkeeping is what is tested here."""
    monkeypatch.setenv("PB200_NGPU", "3")
    if which == "m0_fit":
        D.test_codeml_m0_fit(tmp_path)
    elif which == "m2a_neb_beb":
        T.test_codeml_m2a_neb_beb_fixed_point(tmp_path)
    elif which == "method1_m0":
        T.test_method1_codeml_m0(tmp_path)
    elif which == "genes":
        T.test_baseml_genes(tmp_path, 4, 1)
    else:
        monkeypatch.setenv("PB200_CPN", "1")
        test_rates_for_sites_and_ancestors_method0(shim, tmp_path)
