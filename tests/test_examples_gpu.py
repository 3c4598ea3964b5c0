"""The reference's SHIPPED examples through the drop-in programs on the GPU, against the reference's documented answers.

Data and control files: tests/golden/examples/ (copies of the reference distribution's examples/ and dat/wag.dat -- test
fixtures; the reference tree does not exist on the GPU box).  Every run is the complete program -- the reference's .ctl parser,
sequence / tree readers, initial values, ming2 / minbranches -- with the likelihood path on the GPU; the lnL it prints must be
the value the reference documents (README values) to 1e-6 relative (north_star's tolerance):

  brown.nuc  HKY85 (BASELINE config 1)          -2665.422858, kappa 9.38962
  brown.nuc  REV + G8                            -2619.320596
  lysozymeSmall  codon M0                         -906.017440
  stewart.aa  WAG + G4                           -1028.924417
  MHC (Swanson 2002), the stock codeml.ctl (method = 1, NSsites = 0 1 2 7 8, fixed branch lengths):
      M0 -8225.154790  M1a -7490.993363  M2a -7231.154540  M7 -7502.792534  M8 -7238.014961
                                                  (examples/MHC.Swanson2002MBE/README.txt:26-30)
  MHC M0 with the branch lengths ESTIMATED by `method = 1` (minbranches on HBM-resident vectors, 192 taxa, scaling nodes):
      the tree file holds the M0 estimates, so the fit must come back to -8225.1548.
"""
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(REPO, "tests", "golden", "examples")
BIN = os.path.join(REPO, "paml_b200", "bin")
REL = 1e-6


def _ctl(src, dst, overrides):
    """Copy a reference control file, overriding / appending `key = value` lines (first uncommented occurrence)."""
    out, seen = [], set()
    for ln in open(src).read().splitlines():
        m = re.match(r"\s*([A-Za-z_]+)\s*=", ln)
        if m and m.group(1) in overrides and m.group(1) not in seen:
            out.append("%s = %s" % (m.group(1), overrides[m.group(1)]))
            seen.add(m.group(1))
        else:
            out.append(ln)
    out += ["%s = %s" % (k, v) for k, v in overrides.items() if k not in seen]
    open(dst, "w").write("\n".join(out) + "\n")


def _run(prog, d, files, ctl, overrides, timeout=1500, env=None):
    exe = os.path.join(BIN, prog + "_b200")
    if not os.path.exists(exe):
        pytest.skip("prebuilt drop-in binary missing: " + exe)
    for f in files:
        shutil.copy(os.path.join(EX, f), str(d))
    _ctl(os.path.join(EX, ctl), str(d / "run.ctl"), overrides)
    r = subprocess.run([exe, "run.ctl"], cwd=str(d), stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=timeout, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def _lnls(path):
    return [float(v) for v in re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(path).read())]


def _check(got, want):
    assert abs(got - want) <= REL * abs(want), (got, want)


def test_brown_hky85_baseline_config_1(tmp_path):
    _run("baseml", tmp_path, ["brown.nuc", "brown.trees"], "baseml.ctl", dict(model=4, fix_alpha=1, alpha=0, noisy=0))
    _check(_lnls(str(tmp_path / "mlb"))[-1], -2665.422858)
    txt = open(tmp_path / "mlb").read()
    kappa = float(re.search(r"kappa under HKY85:\s*([\d.]+)", txt).group(1)) if re.search(r"kappa under HKY85", txt) else None
    if kappa is None:
        m = re.search(r"Parameters \(kappa\) in the rate matrix.*?\n\s*([\d.]+)", txt, re.S)
        kappa = float(m.group(1))
    assert abs(kappa - 9.38962) < 2e-3, kappa


def test_brown_rev_gamma8(tmp_path):
    _run("baseml", tmp_path, ["brown.nuc", "brown.trees"], "baseml.ctl", dict(model=7, fix_alpha=0, alpha=0.5, ncatG=8, noisy=0))
    _check(_lnls(str(tmp_path / "mlb"))[-1], -2619.320596)


def test_lysozyme_small_m0(tmp_path):
    _run("codeml", tmp_path, ["lysozymeSmall.nuc", "lysozymeSmall.trees"], "lysozymeSmall.ctl",
         dict(seqfile="lysozymeSmall.nuc", treefile="lysozymeSmall.trees", outfile="mlc", model=0, NSsites=0, CodonFreq=2, getSE=0,
              RateAncestor=0, noisy=0))
    _check(_lnls(str(tmp_path / "mlc"))[0], -906.017440)


def test_stewart_wag_gamma4(tmp_path):
    _run("codeml", tmp_path, ["stewart.aa", "stewart.trees", "wag.dat"], "aaml.ctl",
         dict(seqtype=2, model=2, aaRatefile="wag.dat", fix_alpha=0, alpha=0.5, ncatG=4, getSE=0, RateAncestor=0, noisy=0, verbose=0,
              cleandata=1))
    _check(_lnls(str(tmp_path / "mlc"))[-1], -1028.924417)


def test_mhc_stock_control_file(tmp_path):
    """The flagship example exactly as shipped: method = 1, NSsites = 0 1 2 7 8, branch lengths fixed at the tree file's."""
    _run("codeml", tmp_path, ["bigmhc.phy", "bigmhc.trees"], "mhc_codeml.ctl", dict(noisy=0))
    got = _lnls(str(tmp_path / "mlc"))
    want = [-8225.154790, -7490.993363, -7231.154540, -7502.792534, -7238.014961]
    assert len(got) == len(want), got
    for g, w in zip(got, want):
        _check(g, w)


def test_mhc_m0_branch_lengths_by_minbranches(tmp_path):
    """method = 1 for real: 381 branch lengths of the 192-taxon tree estimated one at a time on HBM-resident conditional vectors
    (10 scaling nodes).  The tree file holds the M0 estimates, so the optimum is the README's M0 value."""
    out = _run("codeml", tmp_path, ["bigmhc.phy", "bigmhc.trees"], "mhc_codeml.ctl", dict(noisy=0, NSsites=0, fix_blength=1),
               env=dict(PB200_VERBOSE="1"))
    assert "minbranches stays on the reference" not in out
    got = _lnls(str(tmp_path / "mlc"))[-1]
    assert got >= -8225.154790 - REL * 8225.15 and got <= -8225.10, got
