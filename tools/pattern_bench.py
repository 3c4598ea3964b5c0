"""Pattern compression at scale: pb200_compress_patterns (hash + sort) next to the unmodified reference's PatternWeight
(sorted insertion, O(npatt^2) moves), both on the host.  The reference is timed as a whole baseml start-up on the same
sequence file (read + compress; it stops at the missing tree file), ours as the library call plus the file parse.
usage: python tools/pattern_bench.py [ns] [ls]"""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paml_b200 import patterns

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ls = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
rng = np.random.default_rng(1)
anc = rng.integers(0, 4, ls)
z = np.empty((ns, ls), np.uint8)
for j in range(ns):
    mut = rng.random(ls) < 0.25
    z[j] = np.where(mut, rng.integers(0, 4, ls), anc)
z = np.frombuffer(b"TCAG", np.uint8)[z]
t0 = time.time(); got = patterns.pattern_weight(z); t1 = time.time()
print("ours: ns %d ls %d -> npatt %d in %.3f s" % (ns, ls, got["npatt"], t1 - t0))
ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "baseml")
if os.path.exists(ref):
    d = tempfile.mkdtemp(prefix="pb200_pw_")
    with open(d + "/s.phy", "w") as f:
        f.write("%d %d\n" % (ns, ls))
        for j in range(ns):
            f.write("s%d  %s\n" % (j + 1, z[j].tobytes().decode()))
    open(d + "/s.tre", "w").write("%d 1\n(%s);\n" % (ns, ",".join("s%d:0.1" % (j + 1) for j in range(ns))))
    open(d + "/run.ctl", "w").write("seqfile = s.phy\ntreefile = s.tre\noutfile = mlb\nnoisy = 3\nverbose = 0\nrunmode = 0\nmodel = 4\n"
                                    "fix_kappa = 1\nkappa = 2\nfix_alpha = 1\nalpha = 0\nncatG = 1\ncleandata = 1\n")
    # the reference prints its progress through PatternWeight when noisy: stop it at the first line after the step
    t0 = time.time()
    pr = subprocess.Popen([ref, "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    tread = None
    for ln in pr.stdout:
        if "Counting site patterns" in ln and tread is None:
            tread = time.time()
        if "Counting frequencies" in ln:
            break
    t1 = time.time()
    pr.kill()
    print("reference: read %.2f s, PatternWeight %.2f s" % (tread - t0, t1 - tread))
