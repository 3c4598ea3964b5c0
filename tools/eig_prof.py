"""One cold solve of the device eigen-solver (61 x 61) for ncu: PB200_EIGEN_WARM=0.  usage: ncu ... python tools/eig_prof.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paml_b200.likelihood import TreeLikelihood
os.environ["PB200_EIGEN_WARM"] = "0"
rng = np.random.default_rng(11); n = 61
S = rng.random((n, n)); S = S + S.T
pi = rng.random(n) + 0.05; pi /= pi.sum()
Q = S * pi[None, :]; np.fill_diagonal(Q, 0); np.fill_diagonal(Q, -Q.sum(1))
tl = TreeLikelihood(0)
for _ in range(3):
    print(tl.eigen_qrev(Q, pi)[3])
tl.close()
