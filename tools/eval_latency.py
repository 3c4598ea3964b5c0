"""Latency of one complete pb200_eval call (host buffers in, -lnL out) on small problems, where an evaluation is launch-bound:
with the evaluation replayed as a CUDA graph (default) and launch by launch (PB200_EVAL_GRAPH=0).
usage: python tools/eval_latency.py"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from paml_b200 import simulate
    from paml_b200.likelihood import TreeLikelihood
    tl = TreeLikelihood(0)
    for name, p in (("codon M0, 8 taxa x 300 patterns", simulate.codon_m0_problem(npatt=300, seed=5)),
                    ("codon 4 classes, 16 taxa x 700 patterns", simulate.codon_sites_problem(ns=16, npatt=700, seed=3)),
                    ("nucleotide GTR+G5, 40 taxa x 20000 patterns", simulate.nuc_gtr_gamma_problem(ns=40, npatt=20000, seed=4, K=5)),
                    ("amino acid +G4, 24 taxa x 3000 patterns", simulate.aa_gamma_problem(ns=24, npatt=3000, seed=4))):
        tl.set_problem(p)
        for _ in range(5):
            v = tl.eval_problem(p)
        t0 = time.perf_counter()
        n = 2000
        for _ in range(n):
            v2 = tl.eval_problem(p)
        dt = (time.perf_counter() - t0) / n
        assert v2 == v
        print("  %-48s %7.1f us per evaluation (through the Python wrapper)   -lnL %.6f" % (name, dt * 1e6, v))
    tl.close()
else:
    for g in ("1", "0"):
        print("PB200_EVAL_GRAPH=%s" % g)
        sys.stdout.flush()
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, PB200_EVAL_GRAPH=g))
