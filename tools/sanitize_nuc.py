import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood, plan_tree
p = simulate.nuc_gtr_gamma_problem(ns=30, npatt=700, seed=4, K=3)
prog, nslots = plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
print("nslots", nslots, "npatt", p.npatt, "nnode", p.nnode)
tl = TreeLikelihood(0); tl.set_problem(p); print(tl.eval_problem(p)); tl.close()
