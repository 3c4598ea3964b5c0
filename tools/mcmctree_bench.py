"""mcmctree with the exact likelihood (usedata = 1) on a locus that is large enough to matter: `ns` species, `nsites` sites simulated
under HKY85 + G5 on a random clock tree with two fossil calibrations.  Runs paml_b200/bin/mcmctree_b200 and (unless --no-ref) the
unmodified reference (oracle/_ref/mcmctree) with the same seed and control file, reports the wall time per MCMC iteration of each
and checks that the two sample files agree.
usage: python tools/mcmctree_bench.py [ns] [nsites] [nsample] [--no-ref] [--no-gpu]"""
import os, subprocess, sys, tempfile, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

args = [a for a in sys.argv[1:] if not a.startswith("--")]
ns = int(args[0]) if len(args) > 0 else 40
nsites = int(args[1]) if len(args) > 1 else 50000
nsample = int(args[2]) if len(args) > 2 else 20
rng = np.random.default_rng(7)

# random rooted binary tree with node ages (root age 1): repeatedly join two lineages
nodes = [dict(name="s%d" % (i + 1), age=0.0, kids=None) for i in range(ns)]
live = list(range(ns))
ages = np.sort(rng.uniform(0.05, 1.0, ns - 1)); ages[-1] = 1.0
for a in ages:
    i, j = rng.choice(len(live), 2, replace=False)
    x, y = live[i], live[j]
    nodes.append(dict(name=None, age=float(a), kids=(x, y)))
    live = [l for k, l in enumerate(live) if k not in (i, j)] + [len(nodes) - 1]
root = live[0]

def newick(v, calib):
    nd = nodes[v]
    if nd["kids"] is None:
        return nd["name"]
    s = "(" + newick(nd["kids"][0], calib) + ", " + newick(nd["kids"][1], calib) + ")"
    if v in calib:
        s += " '>%.3f<%.3f'" % (0.8 * nd["age"], 1.2 * nd["age"])
    return s

internal = [v for v in range(ns, len(nodes))]
calib = {root, internal[len(internal) // 2]}
kappa, alpha, rate = 4.0, 0.6, 0.4
pi = np.array([0.3, 0.25, 0.2, 0.25])
Q = np.ones((4, 4)); Q[0, 1] = Q[1, 0] = Q[2, 3] = Q[3, 2] = kappa
Q = Q * pi[None, :]; np.fill_diagonal(Q, 0); np.fill_diagonal(Q, -Q.sum(1)); Q /= -(pi * np.diag(Q)).sum()
w, V = np.linalg.eig(Q); Vi = np.linalg.inv(V)
site_rate = rng.gamma(alpha, 1 / alpha, nsites)
seq = {root: rng.choice(4, nsites, p=pi)}
def evolve(v):
    nd = nodes[v]
    if nd["kids"] is None:
        return
    for c in nd["kids"]:
        t = (nd["age"] - nodes[c]["age"]) * rate
        out = np.empty(nsites, np.int64)
        # a few rate bins keep the simulation cheap
        bins = np.minimum((site_rate * 4).astype(int), 15)
        for b in np.unique(bins):
            sel = bins == b
            P = np.real((V * np.exp(w * t * (b + 0.5) / 4)[None, :]) @ Vi); P = np.clip(P, 0, None); P /= P.sum(1, keepdims=True)
            cum = P.cumsum(1)
            u = rng.random(sel.sum())
            out[sel] = (u[:, None] > cum[seq[v][sel]]).sum(1)
        seq[c] = np.minimum(out, 3)
        evolve(c)
sys.setrecursionlimit(10000)
evolve(root)
d = tempfile.mkdtemp(prefix="pb200_mt_")
with open(d + "/seq.txt", "w") as f:
    f.write("%d %d\n" % (ns, nsites))
    for i in range(ns):
        f.write("%-10s %s\n" % (nodes[i]["name"], "".join("TCAG"[k] for k in seq[i])))
with open(d + "/tree.txt", "w") as f:
    f.write("%d 1\n%s;\n" % (ns, newick(root, calib)))
ctl = """seed = 4321\nseqfile = seq.txt\ntreefile = tree.txt\nmcmcfile = mcmc.txt\noutfile = out.txt\nndata = 1\nseqtype = 0\nusedata = 1\nclock = 2\n
RootAge = '<2.0'\nmodel = 4\nalpha = 0.5\nncatG = 5\ncleandata = 0\nBDparas = 1 1 0.1 multiplicative\nkappa_gamma = 6 2\nalpha_gamma = 1 1\nrgene_gamma = 2 5 1\n
sigma2_gamma = 1 10 1\nfinetune = 1: .1 .1 .1 .1 .1 .1\nprint = 1\nburnin = %d\nsampfreq = 1\nnsample = %d\n""" % (max(2, nsample // 4), nsample)
res = {}
for tag, exe in (("mcmctree_b200 (GPU)", REPO + "/paml_b200/bin/mcmctree_b200"), ("reference mcmctree (1 core)", REPO + "/oracle/_ref/mcmctree")):
    if ("reference" in tag and "--no-ref" in sys.argv) or ("GPU" in tag and "--no-gpu" in sys.argv):
        continue
    if not os.path.exists(exe):
        print(tag, "missing"); continue
    dd = d + "/" + ("gpu" if "GPU" in tag else "ref"); os.mkdir(dd)
    for fn in ("seq.txt", "tree.txt"):
        os.symlink(d + "/" + fn, dd + "/" + fn)
    open(dd + "/mcmctree.ctl", "w").write(ctl)
    t0 = time.time()
    r = subprocess.run([exe, "mcmctree.ctl"], cwd=dd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.time() - t0
    assert r.returncode == 0, r.stdout[-2000:]
    npatt = [l for l in r.stdout.split("\n") if "patterns" in l.lower()][:1]
    rows = [l for l in open(dd + "/mcmc.txt").read().split("\n")[1:] if l.strip()]
    res[tag] = np.array([[float(x) for x in l.split("\t")] for l in rows])
    iters = max(2, nsample // 4) + nsample
    print("%-30s %d species, %d sites %s: %.2f s wall for %d iterations = %.1f ms per iteration" % (tag, ns, nsites, npatt, dt, iters, dt / iters * 1e3))
if len(res) == 2:
    a, b = res.values()
    print("sample files agree: max relative difference %.2e over %d x %d values" % (np.max(np.abs(a - b) / np.maximum(1e-12, np.abs(b))), *a.shape))
