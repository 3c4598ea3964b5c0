"""Quick kernel-time probe for tuning: prints per-kernel device times of BASELINE config 2 (or others)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
npatt = int(sys.argv[2]) if len(sys.argv) > 2 else None
if which == "cfg2":
    p = simulate.codon_m0_problem(npatt=npatt or 100_000, seed=13147)
elif which == "cfg3":
    p = simulate.codon_sites_problem(ns=32, npatt=npatt or 50_000, seed=13149, freqs=(.4, .3, .2, .1), omegas=(.05, .3, 1, 2.5))
elif which == "cfg4":
    p = simulate.aa_gamma_problem(ns=64, npatt=npatt or 200_000)
elif which == "cfg5":
    p = simulate.nuc_gtr_gamma_problem(ns=128, npatt=npatt or 1_000_000)
tl = TreeLikelihood(0)
tl.set_problem(p)
for _ in range(3):
    v = tl.eval_problem(p)
r = tl.bench_resident(10, 256 << 20)
upd = p.updates_per_eval()
print("%s env[%s] npatt %d K %d: step %.4f ms  pt %.4f  prune %.4f  mix %.4f  -> %.3e upd/s (prune only %.3e)  -lnL %.6f" % (
    which, " ".join("%s=%s" % (k, v_) for k, v_ in os.environ.items() if k.startswith("PB200_")), p.npatt, p.nclass,
    r["step"].mean(), r["pt"].mean(), r["prune"].mean(), r["mix"].mean(), upd / (r["step"].mean() * 1e-3),
    upd / (r["prune"].mean() * 1e-3), v))
print("   ", tl.last_kernel())
