"""Pruning-kernel time of the two codon kernels on MID-SIZE shards (what one GPU holds when BASELINE configs 2 / 3 are sharded over
4-8 GPUs): the warp-independent kernel has a tail there (items / resident warp slots between 1 and 3).  usage: python tools/mid_size.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from paml_b200 import simulate
    from paml_b200.likelihood import TreeLikelihood
    tl = TreeLikelihood(0)
    for ns, npatt, K in ((32, 6250, 4), (32, 12500, 4), (32, 25000, 4), (32, 6250, 1), (32, 12500, 1), (8, 12500, 1), (8, 25000, 1), (8, 50000, 1), (16, 10000, 3)):
        if K == 1:
            p = simulate.codon_m0_problem(ns=ns, npatt=npatt, seed=5)
        else:
            p = simulate.codon_sites_problem(ns=ns, npatt=npatt, seed=3, freqs=tuple([1.0 / K] * K), omegas=tuple([0.1 + 0.6 * i for i in range(K)]))
        tl.set_problem(p)
        for _ in range(3):
            tl.eval_problem(p)
        r = tl.bench_resident(20, 256 << 20)
        print("  %2d taxa x %5d patterns x %d classes (%5d tiles): prune %7.1f us   %s" % (ns, npatt, p.nclass, p.nclass * ((npatt + 7) // 8), r["prune"].mean() * 1e3, tl.last_kernel().split(" grid")[0]))
    tl.close()
else:
    for k in ("warp", "ps"):
        print("PB200_KERNEL=%s" % k); sys.stdout.flush()
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, PB200_KERNEL=k))
