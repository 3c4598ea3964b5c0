"""Latency of one complete pb200_eval call on SMALL codon alignments (the usual size of a codeml data set), default kernel against
the P-stationary kernel (PB200_KERNEL=ps).  usage: python tools/small_latency.py"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from paml_b200 import simulate
    from paml_b200.likelihood import TreeLikelihood
    tl = TreeLikelihood(0)
    for ns, npatt, K in ((8, 130, 1), (8, 1000, 1), (16, 300, 1), (16, 300, 3), (40, 300, 1), (40, 3000, 1), (8, 10000, 1), (16, 20000, 1)):
        p = simulate.codon_m0_problem(ns=ns, npatt=npatt, seed=5) if K == 1 else simulate.codon_sites_problem(ns=ns, npatt=npatt, seed=3)
        tl.set_problem(p)
        for _ in range(5):
            v = tl.eval_problem(p)
        n = 500
        t0 = time.perf_counter()
        for _ in range(n):
            tl.eval_problem(p)
        dt = (time.perf_counter() - t0) / n
        r = tl.bench_resident(20, 0)
        print("  %2d taxa x %5d patterns x %d classes: %7.1f us per call, prune kernel %6.1f us   %s" % (ns, npatt, p.nclass, dt * 1e6, r["prune"].mean() * 1e3, tl.last_kernel().split(" grid")[0]))
    tl.close()
else:
    for k in ("warp", "ps"):
        print("PB200_KERNEL=%s" % k); sys.stdout.flush()
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, PB200_KERNEL=k))
