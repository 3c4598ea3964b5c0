"""Small end-to-end run for compute-sanitizer: every kernel family once (codon, amino acid, nucleotide, scaling,
ambiguity, classes) plus the TMA variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood
tl = TreeLikelihood(0)
probs = [simulate.codon_m0_problem(npatt=300, seed=5), simulate.codon_sites_problem(ns=20, npatt=200, seed=3),
         simulate.add_ambiguity(simulate.codon_m0_problem(ns=10, npatt=200, seed=6), 0.05),
         simulate.aa_gamma_problem(ns=12, npatt=300, seed=4), simulate.nuc_gtr_gamma_problem(ns=30, npatt=700, seed=4, K=3),
         simulate.add_ambiguity(simulate.nuc_gtr_gamma_problem(ns=10, npatt=300, seed=6, K=2), 0.05)]
for p in probs:
    tl.set_problem(p)
    v = tl.eval_problem(p); tl.lnf(); tl.fhK()
    print(p.n, p.ns, p.npatt, p.nclass, v)
tl.close()
