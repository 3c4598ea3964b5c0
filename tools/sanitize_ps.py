"""Small run of the P-stationary kernel (PB200_KERNEL=ps) and of the warm-started eigen-solver for compute-sanitizer
(memcheck / racecheck / synccheck): codon problems with classes, scaling nodes, ambiguity, ragged sizes."""
import os, sys
os.environ["PB200_KERNEL"] = "ps"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood
tl = TreeLikelihood(0)
probs = [simulate.codon_m0_problem(npatt=333, seed=5), simulate.codon_sites_problem(ns=12, npatt=150, seed=3),
         simulate.add_ambiguity(simulate.codon_m0_problem(ns=10, npatt=90, seed=6), 0.05), simulate.codon_m0_problem(ns=40, npatt=70, seed=9)]
for p in probs:
    tl.set_problem(p)
    v = tl.eval_problem(p); tl.lnf()
    print(p.n, p.ns, p.npatt, p.nclass, v, tl.last_kernel())
rng = np.random.default_rng(1)
n = 61
S = rng.random((n, n)); S = S + S.T
pi = rng.random(n) + 0.05; pi /= pi.sum()
for k in range(3):
    Q = S * (1 + 1e-6 * k) * pi[None, :]; np.fill_diagonal(Q, 0); np.fill_diagonal(Q, -Q.sum(1))
    print("eigen sweeps", tl.eigen_qrev(Q, pi)[3])
tl.close()
