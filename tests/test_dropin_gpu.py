"""The drop-in itself: the reference's own driver, .ctl parser, readers and BFGS optimiser with only lfun/lfundG
replaced (paml_b200/bin/codeml_b200, baseml_b200), run next to the unmodified reference binary (oracle/_ref)
on the same files.  Both fit free parameters, so this exercises hundreds of GPU evaluations under ming2's
numerical differentiation; the converged lnL and estimates must agree."""
import os
import re
import subprocess

import numpy as np
import pytest

from paml_b200 import paml_io, simulate

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(REPO, "paml_b200", "bin")
REF = os.path.join(REPO, "oracle", "_ref")


def _run(exe, ctl, cwd, env=None):
    base = dict(os.environ)
    lib = os.path.join(REF, "libfixedseed.so")          # same random starting values for both programs (oracle/fixed_seed.c)
    if os.path.exists(lib):
        base.update(PB200_TEST_SEED=os.environ.get("PB200_TEST_SEED", "20261017"))          # the drop-in derives the same seed itself (SetSeed wrapper in the shim)
        if os.path.dirname(os.path.abspath(exe)) == os.path.abspath(REF):
            base.update(LD_PRELOAD=lib)
    r = subprocess.run([exe, ctl], cwd=cwd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600, env=dict(base, **(env or {})))
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout


def _lnl(path):
    m = re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(path).read())
    assert m, "no lnL line in " + path
    return float(m[-1])


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip("prebuilt binary missing: " + p)


def test_codeml_m0_fit(tmp_path):
    _need(os.path.join(BIN, "codeml_b200"), os.path.join(REF, "codeml"))
    p = simulate.codon_m0_problem(npatt=400, seed=21)
    out = {}
    for tag, exe in (("gpu", os.path.join(BIN, "codeml_b200")), ("ref", os.path.join(REF, "codeml"))):
        d = tmp_path / tag
        d.mkdir()
        paml_io.write_pattern_files(p, str(d / "s"))
        paml_io.write_codeml_ctl(str(d / "run.ctl"), "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1))
        _run(exe, "run.ctl", str(d))
        txt = open(d / "mlc").read()
        out[tag] = dict(lnL=_lnl(str(d / "mlc")),
                        kappa=float(re.search(r"kappa \(ts/tv\) =\s*([\d.]+)", txt).group(1)),
                        omega=float(re.search(r"omega \(dN/dS\) =\s*([\d.]+)", txt).group(1)))
    assert abs(out["gpu"]["lnL"] - out["ref"]["lnL"]) <= 1e-6 * abs(out["ref"]["lnL"]), out
    assert abs(out["gpu"]["kappa"] - out["ref"]["kappa"]) < 2e-3 and abs(out["gpu"]["omega"] - out["ref"]["omega"]) < 2e-3, out


def test_codeml_sites_m2a_fixed_point(tmp_path):
    """NSsites = 2 (M2a), branch lengths and all parameters fixed via in.codeml: a single lfundG evaluation, compared
    digit for digit with the reference's."""
    _need(os.path.join(BIN, "codeml_b200"), os.path.join(REF, "codeml"))
    p = simulate.codon_m0_problem(npatt=300, seed=22)
    vals = {}
    for tag, exe in (("gpu", os.path.join(BIN, "codeml_b200")), ("ref", os.path.join(REF, "codeml"))):
        d = tmp_path / tag
        d.mkdir()
        paml_io.write_pattern_files(p, str(d / "s"))
        paml_io.write_codeml_ctl(str(d / "run.ctl"), "s.phy", "s.tre", NSsites=2, kappa=2.0, omega=1.7,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=2, ncatG=3))
        open(d / "in.codeml", "w").write("-1  2.0  0.6 0.3  0.1 2.5\n")   # kappa, p0 p1, w0 w2 : evaluate only
        _run(exe, "run.ctl", str(d))
        vals[tag] = _lnl(str(d / "mlc"))
    assert abs(vals["gpu"] - vals["ref"]) <= 1e-9 * abs(vals["ref"]), vals


def test_baseml_gtr_gamma_fit(tmp_path):
    _need(os.path.join(BIN, "baseml_b200"), os.path.join(REF, "baseml"))
    p = simulate.nuc_gtr_gamma_problem(ns=12, npatt=600, seed=23, K=4)
    out = {}
    for tag, exe in (("gpu", os.path.join(BIN, "baseml_b200")), ("ref", os.path.join(REF, "baseml"))):
        d = tmp_path / tag
        d.mkdir()
        paml_io.write_pattern_files(p, str(d / "s"))
        with open(d / "run.ctl", "w") as f:
            f.write("seqfile = s.phy\ntreefile = s.tre\noutfile = mlb\nnoisy = 0\nverbose = 0\nrunmode = 0\nmodel = 7\n"
                    "Mgene = 0\nclock = 0\nfix_kappa = 0\nkappa = 2\nfix_alpha = 0\nalpha = 0.7\nMalpha = 0\nncatG = 4\n"
                    "nhomo = 0\ngetSE = 0\nRateAncestor = 0\nSmall_Diff = 7e-6\ncleandata = 1\nfix_blength = 1\nmethod = 0\n")
        _run(exe, "run.ctl", str(d))
        txt = open(d / "mlb").read()
        out[tag] = dict(lnL=_lnl(str(d / "mlb")), alpha=float(re.search(r"alpha \(gamma, K=\s*4\) =\s*([\d.]+)", txt).group(1)))
    assert abs(out["gpu"]["lnL"] - out["ref"]["lnL"]) <= 1e-6 * abs(out["ref"]["lnL"]), out
    assert abs(out["gpu"]["alpha"] - out["ref"]["alpha"]) < 5e-3, out
