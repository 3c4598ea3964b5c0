"""`method = 1` end to end: the same codeml run (M0, branch lengths by minbranches, kappa and omega by ming2) with
  (1) the unmodified reference, (2) the drop-in with minbranches kept on the reference's CPU routines (PB200_MINB=0: only the
  com.plfun calls on the GPU -- what the drop-in did before), (3) the drop-in with minbranches on conditional vectors resident in HBM.
usage: python tools/method1_bench.py [ns] [npatt]"""
import os, re, subprocess, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from paml_b200 import paml_io, simulate

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 32
npatt = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
p = simulate.codon_m0_problem(ns=ns, npatt=npatt, seed=91)
rows = (("unmodified reference (1 core)", REPO + "/oracle/_ref/codeml", {}),
        ("drop-in, minbranches on the CPU (PB200_MINB=0)", REPO + "/paml_b200/bin/codeml_b200", {"PB200_MINB": "0"}),
        ("drop-in, minbranches on the GPU, FP64-vector kernels, host-driven Newton (round 1)", REPO + "/paml_b200/bin/codeml_b200", {"PB200_COND_DMMA": "0", "PB200_NEWTON": "host"}),
        ("drop-in, minbranches on the GPU, tensor-path kernels, host-driven Newton", REPO + "/paml_b200/bin/codeml_b200", {"PB200_NEWTON": "host"}),
        ("drop-in, minbranches on the GPU, tensor-path kernels, device-side Newton (default)", REPO + "/paml_b200/bin/codeml_b200", {}))
if len(sys.argv) > 3 and sys.argv[3] == "gpu":
    rows = rows[2:]
for tag, exe, env in rows:
    if not os.path.exists(exe):
        continue
    d = tempfile.mkdtemp(prefix="pb200_m1_")
    paml_io.write_pattern_files(p, d + "/s")
    paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                             extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1, method=1))
    t0 = time.time()
    r = subprocess.run([exe, "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, env=dict(os.environ, PB200_PROFILE="1", PB200_TEST_SEED="20261017",
                                           **(dict(env, LD_PRELOAD=REPO + "/oracle/_ref/libfixedseed.so") if "_ref/" in exe else env)))
    dt = time.time() - t0
    l = re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(d + "/mlc").read()) if r.returncode == 0 else []
    m = re.search(r"minbranches (\d+) calls, ([\d.]+) s in all, (\d+) branch iterations \(([\d.]+) us each\), (\d+) branch-function evaluations", r.stdout)
    prof = ("  | minbranches: %s s in %s calls, %s evaluations" % (m.group(2), m.group(1), m.group(5))) if m else ""
    print("%-86s %8.2f s wall  lnL %s   (%d taxa, %d codon patterns)%s" % (tag, dt, l[-1] if l else "FAILED rc=%d" % r.returncode, ns, int(p.npatt), prof))
