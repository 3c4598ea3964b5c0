"""mcmctree's exact sequence likelihood (usedata = 1) through the C ABI: paml_b200/bin/mcmctree_b200 is the
reference's mcmctree -- control file, priors, proposals, accept / reject, all unmodified -- with lfun / lfundG answered by ONE
pb200_eval per locus (paml_b200/host/mcmctree_b200.c).  Every scenario runs it beside the unmodified reference binary
(oracle/_ref/mcmctree) with the same seed on the reference's own DatingSoftBound example (7 primates, 3 loci = codon positions,
ambiguity data): the MCMC sample file must agree column for column -- a likelihood that differed beyond rounding would change an
accept / reject decision within a few iterations and the chains would part.
CPU: the same wrapper translation unit linked against the test-only mock of the C ABI (host logic).  GPU: the real library."""
import os
import subprocess

import numpy as np
import pytest

from test_host_logic_cpu import mock_bin  # noqa: F401  (fixture: builds tests/mock/_build)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(REPO, "tests", "golden", "examples")
REF = os.path.join(REPO, "oracle", "_ref", "mcmctree")

CTL = """          seed = 12345
       seqfile = mtCDNApri123.txt
      treefile = mtCDNApri.trees
      mcmcfile = mcmc.txt
       outfile = out.txt
         ndata = {ndata}
       seqtype = 0
       usedata = 1
         clock = {clock}
       RootAge = '<1.0'
         model = {model}
         alpha = {alpha}
         ncatG = {ncatG}
     cleandata = {cleandata}
       BDparas = 1 1 0.1 multiplicative
   kappa_gamma = 6 2
   alpha_gamma = 1 1
   rgene_gamma = 2 20 1
  sigma2_gamma = 1 10 1
      finetune = 1: .1 .1 .1 .1 .1 .1
         print = 1
        burnin = {burnin}
      sampfreq = 2
       nsample = {nsample}
"""
# stock example (HKY85 + G5, independent rates, ambiguity data) | JC69, strict clock | K80 + G4, correlated rates | F84, one rate
SCENARIOS = {
    "hky_g5_clock2": dict(model=4, alpha=1, ncatG=5, clock=2, cleandata=0, ndata=3),
    "jc69_clock1": dict(model=0, alpha=0, ncatG=5, clock=1, cleandata=1, ndata=2),
    "k80_g4_clock3": dict(model=1, alpha=0.5, ncatG=4, clock=3, cleandata=0, ndata=2),
    "f84_clock2": dict(model=3, alpha=0, ncatG=5, clock=2, cleandata=1, ndata=3),
}


# the reference's TipDate example with the exact likelihood (examples/TipDate.HIV2/mcmctree.ExactlnL.ctl): 33 dated HIV-2 sequences,
# birth-death-sequential-sampling prior, strict clock, HKY85 + G5, ambiguity data -- the stock control file with a shorter chain
CTL_TIPDATE = """          seed = 1357
       seqfile = HIV2ge.txt
      treefile = HIV2ge.tre
       outfile = out
      mcmcfile = mcmc.txt
       seqtype = 0
       usedata = 1
         ndata = 1
         clock = {clock}
       TipDate = 1 100
       RootAge = B(0.5, 2.0, 0.01, 0.01)
         model = 4
         alpha = 0.5
         ncatG = 5
     cleandata = 0
       BDparas = 2 1 0 1.8 M
   kappa_gamma = 6 2
   alpha_gamma = 1 1
   rgene_gamma = 2 10
  sigma2_gamma = 1 1
      finetune = 1: 0.5 1.2 .08 .05 .05 .05
         print = 1
        burnin = {burnin}
      sampfreq = 2
       nsample = {nsample}
"""
SCENARIOS["tipdate_hiv2_clock1"] = dict(ctl=CTL_TIPDATE, files=("HIV2ge.txt", "HIV2ge.tre"), clock=1)
SCENARIOS["tipdate_hiv2_clock2"] = dict(ctl=CTL_TIPDATE, files=("HIV2ge.txt", "HIV2ge.tre"), clock=2)


def _table(path):
    rows = open(path).read().split("\n")
    hdr = rows[0].split("\t")
    data = np.array([[float(x) for x in r.split("\t")] for r in rows[1:] if r.strip()])
    return hdr, data


def _scenario(exe, tmp_path, name, burnin=100, nsample=150):
    out = {}
    for tag, prog in (("new", exe), ("ref", REF)):
        d = tmp_path / tag
        d.mkdir()
        sc = dict(SCENARIOS[name])
        ctl = sc.pop("ctl", CTL)
        for f in sc.pop("files", ("mtCDNApri123.txt", "mtCDNApri.trees")):
            if f == "HIV2ge.tre":
                # the shipped file holds five trees; this version of mcmctree accepts one: keep the first (the RAxML tree)
                txt = open(os.path.join(EX, f)).read()
                first = txt[txt.index("("):txt.index(";") + 1]
                open(d / f, "w").write("33 1\n\n" + first + "\n")
            else:
                os.symlink(os.path.join(EX, f), d / f)
        open(d / "mcmctree.ctl", "w").write(ctl.format(burnin=burnin, nsample=nsample, **sc))
        r = subprocess.run([prog, "mcmctree.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:]
        out[tag] = _table(d / "mcmc.txt")
    (h1, a), (h2, b) = out["new"], out["ref"]
    assert h1 == h2 and a.shape == b.shape and a.shape[0] == nsample + 1, (a.shape, b.shape)
    assert h1[-1] == "lnL"
    # lnL to the north-star tolerance; the sampled times / rates / kappa / alpha follow from identical accept / reject decisions
    np.testing.assert_allclose(a[:, -1], b[:, -1], rtol=1e-6, atol=0)
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_mcmctree_host_logic(mock_bin, tmp_path, name):  # noqa: F811
    if not os.path.exists(REF):
        pytest.skip("reference binary not built (make -C oracle ref)")
    big = name.startswith("tipdate")                    # 33 taxa x 1107 sites through the scalar oracle: keep the chain short
    _scenario(os.path.join(mock_bin, "mcmctree_mock"), tmp_path, name, burnin=10 if big else 40, nsample=16 if big else 60)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_mcmctree_gpu(tmp_path, name):
    exe = os.path.join(REPO, "paml_b200", "bin", "mcmctree_b200")
    for p in (exe, REF):
        if not os.path.exists(p):
            pytest.skip("prebuilt binary missing: " + p)
    _scenario(exe, tmp_path, name)
