"""The drop-in on the model families beyond the BASELINE configurations (SURVEY.md 8a "model coverage" table and
8f rank 4): several genes / partitions (Mgene = 0, 2, 3, 4, per-gene alpha), nonhomogeneous baseml models (nhomo =
2..5: spectral sets per branch), UNREST (Taylor + squaring matexp on the device), free-ratio codon model (more omega
types than the reference's _UU[] holds: eigen-solution per branch), and the consumers of fx_r: auto-discrete-gamma
(lfunAdG), NEB (lfunNSsites_rate) and the M2a BEB grid.  Each case runs paml_b200/bin/*_b200 next to the unmodified
reference binary (oracle/_ref) on the same files and compares what they print."""
import os
import re
import subprocess

import pytest

from paml_b200 import paml_io, simulate

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(REPO, "paml_b200", "bin")
REF = os.path.join(REPO, "oracle", "_ref")

BASEML_CTL = ("seqfile = s.phy\ntreefile = s.tre\noutfile = mlb\nnoisy = 0\nverbose = 0\nrunmode = 0\nMgene = %(Mgene)d\n"
              "clock = 0\nMalpha = %(Malpha)d\ngetSE = 0\nRateAncestor = 0\nSmall_Diff = 7e-6\ncleandata = 1\nmethod = 0\n"
              "model = %(model)d\nfix_kappa = 0\nkappa = 2\nfix_alpha = %(fix_alpha)d\nalpha = %(alpha)s\nncatG = %(ncatG)d\n"
              "nhomo = %(nhomo)d\nfix_blength = 1\n%(more)s")


def _baseml_ctl(path, **kw):
    d = dict(Mgene=0, Malpha=0, model=4, fix_alpha=1, alpha="0", ncatG=1, nhomo=0, more="")
    d.update(kw)
    with open(path, "w") as f:
        f.write(BASEML_CTL % d)


def _same_start_env(exe=""):
    """Both programs perturb their initial values with a generator seeded from /dev/urandom (SetSeed, src/tools.c:1378).  The
    unmodified reference gets oracle/_ref/libfixedseed.so preloaded (test infrastructure, oracle/fixed_seed.c: fopen of
    /dev/urandom answers with bytes derived from PB200_TEST_SEED); the drop-in derives the same seed from that variable in
    its own SetSeed wrapper (plfun_b200.inc) and needs no preload.  So both start every fit at the same point."""
    lib = os.path.join(REF, "libfixedseed.so")
    if not os.path.exists(lib):
        return dict(os.environ)
    env = dict(os.environ, PB200_TEST_SEED=os.environ.get("PB200_TEST_SEED", "20261017"))
    if os.path.dirname(os.path.abspath(exe)) == os.path.abspath(REF):
        env["LD_PRELOAD"] = lib
    return env


def _run(exe, cwd, ok=True):
    r = subprocess.run([exe, "run.ctl"], cwd=cwd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900, env=_same_start_env(exe))
    if ok:
        assert r.returncode == 0, r.stdout[-2000:]
    return r


def _lnl(path):
    m = re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(path).read())
    assert m, "no lnL line in " + path
    return float(m[-1])


def _pair(prog):
    gpu, ref = os.path.join(BIN, prog + "_b200"), os.path.join(REF, prog)
    for p in (gpu, ref):
        if not os.path.exists(p):
            pytest.skip("prebuilt binary missing: " + p)
    return (("gpu", gpu), ("ref", ref))


def _both(tmp_path, prog, write, outfile):
    """write(dir) prepares one run directory; returns {tag: (lnL, dir)}."""
    res = {}
    for tag, exe in _pair(prog):
        d = tmp_path / tag
        d.mkdir(parents=True)
        write(str(d))
        _run(exe, str(d))
        res[tag] = (_lnl(str(d / outfile)), str(d))
    return res


def _close(res, rel):
    g, r = res["gpu"][0], res["ref"][0]
    assert abs(g - r) <= rel * abs(r), (g, r)


@pytest.mark.parametrize("mgene,malpha", [(0, 0), (2, 0), (3, 0), (4, 1)])
def test_baseml_genes(tmp_path, mgene, malpha):
    """Two partitions (option G): shared kappa/pi with gene rates (Mgene 0), per-gene pi (2), kappa (3), both + per-gene
    alpha (4, Malpha 1); every gene is evaluated with its own SetPGene parameters (src/treesub.c:7713-7716, 7773-7775)."""
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=200, seed=33, K=4)
    n = int(p.npatt)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s", genes=[n - n // 3, n // 3])
        _baseml_ctl(d + "/run.ctl", Mgene=mgene, Malpha=malpha, fix_alpha=0, alpha="0.5", ncatG=4)
    _close(_both(tmp_path, "baseml", write, "mlb"), 2e-6)


@pytest.mark.parametrize("nhomo", [2, 3, 4, 5])
def test_baseml_nhomo(tmp_path, nhomo):
    """Nonhomogeneous models: kappa per branch (2), N1 (3), N2 = pi per node (4), labelled pi sets (5):
    GetPMatBranch re-derives Root/Cijk for every branch (src/treesub.c:7514-7519) -> one spectral set per branch."""
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=150, seed=32, K=1)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s")
        _baseml_ctl(d + "/run.ctl", nhomo=nhomo)
    _close(_both(tmp_path, "baseml", write, "mlb"), 5e-6)


def test_baseml_unrest(tmp_path):
    """UNREST (model 8): P(t) by the Taylor + squaring matexp on the device (src/treesub.c:7523-7527)."""
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=150, seed=31, K=1)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s")
        _baseml_ctl(d + "/run.ctl", model=8)
    _close(_both(tmp_path, "baseml", write, "mlb"), 2e-6)


def test_baseml_auto_discrete_gamma(tmp_path):
    """fix_rho = 0: com.plfun = lfunAdG (src/treesub.c:7447-7494), whose per-class values come from fx_r -> the GPU;
    its hidden-Markov recursion over the site order stays the reference's host C.  Ordinary (non-pattern) alignment."""
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=200, seed=44, K=4)

    def write(d):
        paml_io.write_sequence_files(p, d + "/s")
        _baseml_ctl(d + "/run.ctl", fix_alpha=0, alpha="0.5", ncatG=4, more="fix_rho = 0\nrho = 0.2\n")
    _close(_both(tmp_path, "baseml", write, "mlb"), 2e-6)


@pytest.mark.parametrize("mgene", [0, 3])
def test_codeml_genes(tmp_path, mgene):
    p = simulate.codon_m0_problem(ns=6, npatt=200, seed=41)
    n = int(p.npatt)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s", genes=[n - n // 3, n // 3])
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, Mgene=mgene))
    _close(_both(tmp_path, "codeml", write, "mlc"), 2e-6)


def test_codeml_free_ratio_fixed_point(tmp_path):
    """model = 1 with 13 branches: more omega types than _UU[] holds, the reference solves an eigen-problem per branch
    inside GetPMatBranch (src/treesub.c:7574-7583); one evaluation at a fixed point, compared digit for digit."""
    p = simulate.codon_m0_problem(ns=8, npatt=200, seed=43)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", model=1, kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=2))
        with open(d + "/in.codeml", "w") as f:
            f.write("-1  2.0 " + " ".join("%.2f" % (0.1 + 0.07 * i) for i in range(13)) + "\n")
    _close(_both(tmp_path, "codeml", write, "mlc"), 1e-9)


def test_codeml_m3_five_classes_fixed_point(tmp_path):
    """NSsites = 3 with K = 5: five eigenQcodon calls per evaluation (SetPSiteClass, src/treesub.c:7677-7686), which the
    drop-in queues and solves in ONE device launch (pb200_eigen_qrev, five thread blocks); also run with the reference's
    host solver (PB200_EIGEN=host) and with every solution forced to the device (PB200_EIGEN=device)."""
    p = simulate.codon_sites_problem(ns=7, npatt=150, seed=42)

    def write(d):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", NSsites=3, kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=2, ncatG=5))
        with open(d + "/in.codeml", "w") as f:
            f.write("-1  2.0  0.3 0.25 0.2 0.15  0.05 0.3 0.8 1.5 3.0\n")
    res = _both(tmp_path, "codeml", write, "mlc")
    _close(res, 1e-9)
    exe = os.path.join(BIN, "codeml_b200")
    for mode in ("host", "device"):
        d = tmp_path / mode
        d.mkdir()
        write(str(d))
        r = subprocess.run([exe, "run.ctl"], cwd=str(d), stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=900, env=dict(_same_start_env(exe), PB200_EIGEN=mode))
        assert r.returncode == 0, r.stdout[-2000:]
        v = _lnl(str(d / "mlc"))
        assert abs(v - res["ref"][0]) <= 1e-9 * abs(res["ref"][0]), (mode, v, res["ref"][0])


def _numbers(path):
    """Every decimal number of a report, with the number of decimals it was printed with."""
    out = []
    for tok in re.findall(r"-?\d+\.\d+", open(path).read()):
        out.append((float(tok), len(tok.split(".")[1])))
    return out


def test_codeml_m2a_neb_beb_fixed_point(tmp_path):
    """M2a at a fixed parameter point: lfundG for the lnL, then NEB (lfunNSsites_rate, src/codeml.c:5241) and the BEB grid
    of 21 omega classes (src/codeml.c:6280) -- both read com.fhK[] as left by fx_r, which the drop-in computes on the
    GPU.  The posterior tables in `rst` must agree to the printed precision."""
    p = simulate.codon_sites_problem(ns=7, npatt=150, seed=42)

    def write(d):
        paml_io.write_sequence_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", NSsites=2, kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=2, ncatG=3))
        with open(d + "/in.codeml", "w") as f:
            f.write("-1  2.0  0.5 0.45  0.1 5.0\n")
    res = _both(tmp_path, "codeml", write, "mlc")
    _close(res, 1e-9)
    g, r = _numbers(res["gpu"][1] + "/rst"), _numbers(res["ref"][1] + "/rst")
    assert len(g) == len(r) and len(r) > 1000, (len(g), len(r))
    assert "Bayes Empirical Bayes" in open(res["gpu"][1] + "/rst").read()
    bad = [(a, b) for (a, da), (b, db) in zip(g, r) if abs(a - b) > 2.5 * 10.0 ** (-min(da, db))]
    assert not bad, bad[:10]


# ---- method = 1: minbranches on conditional vectors resident in HBM ----

def _close_method1(res, prog, write0, outfile, tmp_path, rel=2e-6):
    """method = 1 alternates two optimisers (ming2 on the substitution parameters, Newton steps on one branch at a time).  Both
    programs start every fit from the same values (PB200_TEST_SEED / oracle/fixed_seed.c), every evaluation agrees with the
    reference to ~1e-13, so the drop-in must follow the reference's own method = 1 trajectory and end where it ends -- also when
    that end point is a poor one (e.g. HKY+G4 from some starts: alpha at its bound).  No second chance against method = 0."""
    g, r = res["gpu"][0], res["ref"][0]
    assert abs(g - r) <= rel * abs(r), dict(gpu_method1=g, ref_method1=r)


def _fixed_start(path, ntime, tail):
    """Both programs perturb their initial values with a generator seeded from the clock / urandom (src/baseml.c:157, 1253-1263);
    with gamma rates method = 1 then ends, from some starts, at a poor point with alpha at its bound -- in the reference as much as
    here.  An initial-values file (readx, src/treesub.c:4033-4066: np numbers, branch lengths first) makes both start at the same point."""
    with open(path, "w") as f:
        f.write(" ".join(["0.1"] * ntime + ["%g" % v for v in tail]) + "\n")


def _verbose_says_gpu(d, exe):
    """With PB200_VERBOSE the shim reports when minbranches stays on the reference's CPU routines: it must not."""
    r = subprocess.run([exe, "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900, env=dict(_same_start_env(exe), PB200_VERBOSE="1"))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "minbranches stays on the reference" not in r.stdout


def test_method1_codeml_m0(tmp_path):
    p = simulate.codon_m0_problem(ns=8, npatt=300, seed=71)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, method=method))
    res = _both(tmp_path, "codeml", write, "mlc")
    _close_method1(res, "codeml", lambda d: write(d, 0), "mlc", tmp_path)
    _verbose_says_gpu(res["gpu"][1], os.path.join(BIN, "codeml_b200"))


def test_method1_codeml_m2a_site_classes(tmp_path):
    """NSsites = 2 with method = 1: lfunt_SiteClass / lfuntdd_SiteClass, vectors per site class."""
    p = simulate.codon_sites_problem(ns=7, npatt=150, seed=42)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", NSsites=2, kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, ncatG=3, method=method))
    _close_method1(_both(tmp_path, "codeml", write, "mlc"), "codeml", lambda d: write(d, 0), "mlc", tmp_path)


def test_method1_baseml_hky_gamma(tmp_path):
    p = simulate.nuc_gtr_gamma_problem(ns=12, npatt=600, seed=23, K=4)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        with open(d + "/run.ctl", "w") as f:
            f.write((BASEML_CTL % dict(Mgene=0, Malpha=0, model=4, fix_alpha=0, alpha="0.5", ncatG=4, nhomo=0, more=""))
                    .replace("method = 0", "method = %d" % method))
        _fixed_start(d + "/in.baseml", 2 * p.ns - 3, [2.0, 0.5])
    _close_method1(_both(tmp_path, "baseml", write, "mlb"), "baseml", lambda d: write(d, 0), "mlb", tmp_path)


def test_method1_codeml_40_taxa_with_scaling_nodes(tmp_path):
    """40 taxa: the reference scales at nodes with more than 15 tips below (NodeScale); rerooting moves those nodes around."""
    p = simulate.codon_m0_problem(ns=40, npatt=150, seed=72)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, method=method))
    _close_method1(_both(tmp_path, "codeml", write, "mlc"), "codeml", lambda d: write(d, 0), "mlc", tmp_path)


@pytest.mark.parametrize("model,nhomo", [(1, 0), (0, 0), (4, 2), (1, 2)])
def test_method1_baseml_k80_and_kappa_per_branch(tmp_path, model, nhomo):
    """K80 / JC69 (closed-form P(t) in the evaluation path, their Cijk form under minbranches, as in lfunt) and nhomo = 2
    (kappa per branch: one Cijk set per branch, lfunt's per-branch eigenTN93, src/treesub.c:8149-8152)."""
    p = simulate.nuc_gtr_gamma_problem(ns=8, npatt=300, seed=26, K=1)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        with open(d + "/run.ctl", "w") as f:
            f.write((BASEML_CTL % dict(Mgene=0, Malpha=0, model=model, fix_alpha=1, alpha="0", ncatG=1, nhomo=nhomo, more=""))
                    .replace("method = 0", "method = %d" % method))
    res = _both(tmp_path, "baseml", write, "mlb")
    _close_method1(res, "baseml", lambda d: write(d, 0), "mlb", tmp_path)
    _verbose_says_gpu(res["gpu"][1], os.path.join(BIN, "baseml_b200"))


def test_method1_codeml_free_ratio(tmp_path):
    """model = 1 under method = 1: lfunt / lfuntdd solve an eigen-problem for the branch at every call (src/treesub.c:8137-8146);
    the drop-in keeps one spectral set per branch (memoised) resident for pb200_branch_lnl."""
    p = simulate.codon_m0_problem(ns=8, npatt=200, seed=43)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", model=1, kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, method=method))
    res = _both(tmp_path, "codeml", write, "mlc")
    _close_method1(res, "codeml", lambda d: write(d, 0), "mlc", tmp_path, rel=5e-6)
    _verbose_says_gpu(res["gpu"][1], os.path.join(BIN, "codeml_b200"))


@pytest.mark.parametrize("mgene,malpha,gamma", [(0, 0, False), (2, 0, False), (3, 0, True), (4, 1, True)])
def test_method1_baseml_genes(tmp_path, mgene, malpha, gamma):
    """method = 1 with several genes / partitions (round 2; round 1 ran these on the reference's CPU routines): every gene's
    conditional vectors live in that gene's context and are updated with that gene's parameters (SetPGene: kappa, pi, rgene,
    alpha), lfuntdd / lfuntdd_SiteClass sum over the genes (src/treesub.c:8239-8241, 8472-8476)."""
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=240, seed=37, K=4)
    n = int(p.npatt)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s", genes=[n - n // 3, n // 3])
        _baseml_ctl(d + "/run.ctl", Mgene=mgene, Malpha=malpha, model=4, fix_alpha=(0 if gamma else 1), alpha=("0.5" if gamma else "0"),
                    ncatG=(4 if gamma else 1))
        txt = open(d + "/run.ctl").read().replace("method = 0", "method = %d" % method)
        open(d + "/run.ctl", "w").write(txt)
    res = _both(tmp_path, "baseml", write, "mlb")
    _close_method1(res, "baseml", lambda d: write(d, 0), "mlb", tmp_path)
    _verbose_says_gpu(res["gpu"][1], os.path.join(BIN, "baseml_b200"))


def test_method1_codeml_genes(tmp_path):
    """Codon partitions with their own kappa, omega and codon frequencies (Mgene = 4) under method = 1."""
    p = simulate.codon_m0_problem(ns=6, npatt=180, seed=39)
    n = int(p.npatt)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s", genes=[n // 2, n - n // 2])
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, Mgene=4, method=method))
    res = _both(tmp_path, "codeml", write, "mlc")
    _close_method1(res, "codeml", lambda d: write(d, 0), "mlc", tmp_path, rel=5e-6)
    _verbose_says_gpu(res["gpu"][1], os.path.join(BIN, "codeml_b200"))


def test_method1_aaml_poisson(tmp_path):
    """aaml Poisson under method = 1: the evaluation uses the closed form (PMatJC69like), minbranches the spectral form eigenQaa
    left in Root / U / V once for all -- on the GPU since round 2."""
    p = simulate.aa_gamma_problem(ns=7, npatt=150, seed=51, K=1)

    def write(d, method=1):
        paml_io.write_pattern_files(p, d + "/s")
        paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", seqtype=2,
                                 extra=dict(model=0, fix_alpha=1, alpha=0, ncatG=1, fix_blength=1, method=method))
    res = _both(tmp_path, "codeml", write, "mlc")
    _close_method1(res, "codeml", lambda d: write(d, 0), "mlc", tmp_path)
    _verbose_says_gpu(res["gpu"][1], os.path.join(BIN, "codeml_b200"))


def test_method1_cpu_minbranches_paths_get_fresh_host_vectors(tmp_path):
    """Where minbranches is NOT mapped to the GPU (PB200_MINB=0 forces that here) the reference's CPU routines run on
    nodes[].conP, which the GPU likelihood never fills: the shim refreshes them with the reference's own CPU evaluation
    first.  (Without that the fit converges to a wrong optimum.)"""
    q = simulate.codon_m0_problem(ns=8, npatt=300, seed=71)
    ref = None
    for tag, exe, env in (("ref", os.path.join(REF, "codeml"), {}), ("cpu_minb", os.path.join(BIN, "codeml_b200"), {"PB200_MINB": "0"})):
        d = tmp_path / tag
        d.mkdir()
        paml_io.write_pattern_files(q, str(d / "s"))
        paml_io.write_codeml_ctl(str(d / "run.ctl"), "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, method=1))
        r = subprocess.run([exe, "run.ctl"], cwd=str(d), stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=900, env=dict(_same_start_env(exe), **env))
        assert r.returncode == 0, r.stdout[-2000:]
        v = _lnl(str(d / "mlc"))
        if ref is None:
            ref = v
        assert abs(v - ref) <= 2e-6 * abs(ref), (tag, v, ref)


def test_method1_ancestral_reconstruction_with_site_classes(tmp_path):
    """RateAncestor = 1 under method = 1 with gamma rates: PostProbNode (src/treesub.c:6217-6220) reroots and calls updateconP
    on the HOST vectors, which the shim refreshes with the reference's CPU evaluation (every node, at the final parameters);
    the marginal reconstruction in `rst` must agree with the reference's.  (The reference itself reconstructs from host vectors
    left by an earlier state of its iteration -- its per-node lnL line differs from its own fit -- hence the 2e-3 allowance.)"""
    p = simulate.nuc_gtr_gamma_problem(ns=8, npatt=300, seed=25, K=4)

    def write(d, method=1):
        paml_io.write_sequence_files(p, d + "/s")
        with open(d + "/run.ctl", "w") as f:
            f.write((BASEML_CTL % dict(Mgene=0, Malpha=0, model=4, fix_alpha=0, alpha="0.5", ncatG=4, nhomo=0, more=""))
                    .replace("method = 0", "method = %d" % method).replace("RateAncestor = 0", "RateAncestor = 1"))
        _fixed_start(d + "/in.baseml", 2 * p.ns - 3, [2.0, 0.5])
    res = _both(tmp_path, "baseml", write, "mlb")
    _close_method1(res, "baseml", lambda d: write(d, 0), "mlb", tmp_path)
    g, r = _numbers(res["gpu"][1] + "/rst"), _numbers(res["ref"][1] + "/rst")
    assert len(g) == len(r) and len(r) > 500, (len(g), len(r))
    # both programs start their iteration from randomly perturbed initial values (src/baseml.c:1253-1263): the two fits agree to
    # the optimiser's convergence tolerance, not to the last printed digit
    bad = [(a, b) for (a, da), (b, db) in zip(g, r) if abs(a - b) > 1e-2 + 2.5 * 10.0 ** (-min(da, db))]
    assert not bad, bad[:10]


def test_too_many_genes_fails_loudly(tmp_path):
    """No CPU fallback: what the GPU path does not implement stops the program with a message naming it."""
    exe = os.path.join(BIN, "baseml_b200")
    if not os.path.exists(exe):
        pytest.skip("prebuilt binary missing: " + exe)
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=200, seed=33, K=1)
    n = int(p.npatt)
    assert n > 64
    paml_io.write_pattern_files(p, str(tmp_path / "s"), genes=[1] * n)
    _baseml_ctl(str(tmp_path / "run.ctl"))
    r = _run(exe, str(tmp_path), ok=False)
    assert r.returncode != 0 and "not implemented on the GPU path" in r.stdout, r.stdout[-500:]
