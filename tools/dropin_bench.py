"""Evaluations per second of the drop-in itself: codeml_b200 fitting kappa and omega (branch lengths fixed) on a
BASELINE-config-2-shaped data set, i.e. the whole com.plfun call as ming2 drives it -- SetParameters (host Q
construction + eigen-solution), gather, H2D, kernels, D2H.  Run once with the reference's host eigenQREV
(PB200_EIGEN=host) and once with the device Jacobi solver (default), and, if present, the unmodified reference.
usage: python tools/dropin_bench.py [npatt] [m0|sites]   (sites: BASELINE config 3 shape, 32 taxa, NSsites = 3 with 4 classes:
four eigen-solutions per evaluation, which the drop-in batches into one device launch)"""
import os, re, subprocess, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from paml_b200 import paml_io, simulate

npatt = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kind = sys.argv[2] if len(sys.argv) > 2 else "m0"
if kind == "m0":
    p = simulate.codon_m0_problem(npatt=npatt, seed=13147)
    EXTRA, INIT = dict(fix_kappa=0, fix_omega=0, fix_blength=2), "-1 2.0 0.4\n"
else:
    p = simulate.codon_sites_problem(ns=32, npatt=npatt, seed=13149)
    EXTRA, INIT = dict(fix_kappa=0, fix_omega=0, fix_blength=2, NSsites=3, ncatG=4), "-1 2.0  0.4 0.3 0.2  0.05 0.3 1.0 2.5\n"


def run(exe, env, fit):
    d = tempfile.mkdtemp(prefix="pb200_db_")
    paml_io.write_pattern_files(p, d + "/s")
    paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                             extra=EXTRA)
    if not fit:
        open(d + "/in.codeml", "w").write(INIT)
    t0 = time.time()
    r = subprocess.run([exe, "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, env=dict(os.environ, **env))
    dt = time.time() - t0
    m = re.search(r"(\d+) lfun", r.stdout)
    l = re.findall(r"lnL\(ntime:\s*\d+\s+np:\s*\d+\):\s*(-?[\d.]+)", open(d + "/mlc").read())
    prof = re.search(r"paml_b200 profile: (.*)", r.stdout)
    return dt, int(m.group(1)) if m else 0, float(l[-1]) if l else float("nan"), prof.group(1) if prof else ""


for tag, exe, env in (("drop-in, device eigen", REPO + "/paml_b200/bin/codeml_b200", {}),
                      ("drop-in, host eigen (reference eigenQREV)", REPO + "/paml_b200/bin/codeml_b200", {"PB200_EIGEN": "host"}),
                      ("unmodified reference", REPO + "/oracle/_ref/codeml", {})):
    if not os.path.exists(exe):
        continue
    if "reference" == tag.split()[-1] and npatt > 20000:
        print("%-45s skipped (npatt > 20000 takes minutes on one core)" % tag)
        continue
    env = dict(env, PB200_PROFILE="1")
    t1, n1, l1, _ = run(exe, env, False)
    t2, n2, l2, prof = run(exe, env, True)
    per = (t2 - t1) / max(1, n2 - n1)
    print("%-45s fit: %.2f s wall, %d lfun, lnL %.6f; wall difference to a 1-evaluation run: %.3f ms per evaluation"
          % (tag, t2, n2, l2, per * 1e3))
    if prof:
        print("    in-process: " + prof)
