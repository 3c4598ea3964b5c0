"""Multi-GPU paths on real hardware (skipped on a single-GPU box): several contexts in one process (the drop-in's
PB200_NGPU mode) and one process per GPU with the NCCL all-reduce inside pb200_eval (bench.py's mode)."""
import os
import re
import subprocess
import sys

import pytest
import torch

from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood
from paml_b200.shard import shard_bounds

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_two_contexts_one_process():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    p = simulate.codon_m0_problem(npatt=2500, seed=31)
    whole = TreeLikelihood(0); whole.set_problem(p); ref = whole.eval_problem(p); whole.close()
    ctxs, tot = [], 0.0
    for r in range(2):
        t = TreeLikelihood(r); h0, h1 = shard_bounds(p.npatt, 2, r); t.set_problem(p, h0, h1); ctxs.append(t)
    for t in ctxs:
        tot += t.eval_problem(p)
    for t in ctxs:
        t.close()
    assert abs(tot - ref) <= 1e-12 * abs(ref)


def test_torchrun_nccl_allreduce():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29671", os.path.join(REPO, "tests", "multi_rank_worker.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    m = re.search(r"MULTI_RANK_RESULT (\d+) .* (OK|MISMATCH)", r.stdout)
    assert r.returncode == 0 and m and m.group(2) == "OK", r.stdout[-3000:]


def test_dropin_two_gpus(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from paml_b200 import paml_io
    exe = os.path.join(REPO, "paml_b200", "bin", "codeml_b200")
    if not os.path.exists(exe):
        pytest.skip("prebuilt binary missing")
    p = simulate.codon_m0_problem(npatt=600, seed=41)
    vals = []
    for ng in (1, 2):
        d = tmp_path / ("g%d" % ng); d.mkdir()
        paml_io.write_pattern_files(p, str(d / "s"))
        paml_io.write_codeml_ctl(str(d / "run.ctl"), "s.phy", "s.tre", kappa=3.0, omega=0.5)
        r = subprocess.run([exe, "run.ctl"], cwd=str(d), env=dict(os.environ, PB200_NGPU=str(ng)), stdin=subprocess.DEVNULL,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:]
        vals.append(float(re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(d / "mlc").read())[-1]))
    assert abs(vals[0] - vals[1]) <= 1e-9 * abs(vals[0]), vals


def _fit_lnl(exe, d, env, outfile):
    base = dict(os.environ)
    lib = os.path.join(REPO, "oracle", "_ref", "libfixedseed.so")     # same random starting values in every run (oracle/fixed_seed.c)
    if os.path.exists(lib):
        base.update(PB200_TEST_SEED=os.environ.get("PB200_TEST_SEED", "20261017"))           # the drop-in seeds itself from it (SetSeed wrapper in the shim); no preload
    r = subprocess.run([exe, "run.ctl"], cwd=str(d), env=dict(base, **env), stdin=subprocess.DEVNULL,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    return float(re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(d / outfile).read())[-1])


def test_dropin_two_gpus_method1(tmp_path):
    """method = 1 with the patterns sharded over two GPUs: every shard keeps its own HBM-resident conditional vectors, the
    shim adds the shards' -lnL, d/dt and d2/dt2 of every Newton step (paml_b200/host/plfun_b200.inc, pbh_mb_branch)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from paml_b200 import paml_io
    exe = os.path.join(REPO, "paml_b200", "bin", "codeml_b200")
    if not os.path.exists(exe):
        pytest.skip("prebuilt binary missing")
    p = simulate.codon_m0_problem(ns=8, npatt=600, seed=71)
    vals = []
    for ng in (1, 2):
        d = tmp_path / ("g%d" % ng); d.mkdir()
        paml_io.write_pattern_files(p, str(d / "s"))
        paml_io.write_codeml_ctl(str(d / "run.ctl"), "s.phy", "s.tre", kappa=2.0, omega=0.4,
                                 extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, method=1))
        vals.append(_fit_lnl(exe, d, dict(PB200_NGPU=str(ng)), "mlc"))
    assert abs(vals[0] - vals[1]) <= 2e-6 * abs(vals[0]), vals


def test_dropin_two_gpus_genes(tmp_path):
    """Two partitions with their own kappa and pi (Mgene = 4): gene g is evaluated on GPU g % 2."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from paml_b200 import paml_io
    exe = os.path.join(REPO, "paml_b200", "bin", "baseml_b200")
    if not os.path.exists(exe):
        pytest.skip("prebuilt binary missing")
    p = simulate.nuc_gtr_gamma_problem(ns=6, npatt=200, seed=33, K=4)
    n = int(p.npatt)
    vals = []
    for ng in (1, 2):
        d = tmp_path / ("g%d" % ng); d.mkdir()
        paml_io.write_pattern_files(p, str(d / "s"), genes=[n - n // 3, n // 3])
        with open(d / "run.ctl", "w") as f:
            f.write("seqfile = s.phy\ntreefile = s.tre\noutfile = mlb\nnoisy = 0\nverbose = 0\nrunmode = 0\nMgene = 4\nclock = 0\n"
                    "Malpha = 1\ngetSE = 0\nRateAncestor = 0\nSmall_Diff = 7e-6\ncleandata = 1\nmethod = 0\nmodel = 4\nfix_kappa = 0\n"
                    "kappa = 2\nfix_alpha = 0\nalpha = 0.5\nncatG = 4\nnhomo = 0\nfix_blength = 1\n")
        vals.append(_fit_lnl(exe, d, dict(PB200_NGPU=str(ng)), "mlb"))
    assert abs(vals[0] - vals[1]) <= 2e-6 * abs(vals[0]), vals
