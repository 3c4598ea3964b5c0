"""Workload for profiling the kernels under `method = 1` (csrc/cond.cuh) and the eigen-solver: BASELINE-config-3 shape
(32 taxa, 61 states, K site classes), a full update of every internal node's resident vectors, then Newton-step evaluations
of one branch.  Prints host-clock times of the (synchronous) C-ABI calls; run under ncu for the kernel metrics.
usage: python tools/cond_profile.py [npatt] [K]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood

npatt = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
freqs = tuple([1.0 / K] * K)
omegas = tuple(np.linspace(0.1, 2.5, K))
p = simulate.codon_sites_problem(ns=32, npatt=npatt, seed=13149, freqs=freqs, omegas=omegas)
tl = TreeLikelihood(0)
tl.set_problem(p)
sons = {v: [int(c) for c in p.sons[v, :p.nson[v]]] for v in range(p.nnode)}
order = []
def rec(v):
    for c in sons[v]:
        if c >= p.ns:
            rec(c)
    order.append((v, sons[v]))
rec(p.root)
nint = p.nnode - p.ns
for rep in range(3):
    t0 = time.perf_counter()
    tl.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, order, None)
    dt = time.perf_counter() - t0
flops = sum(2.0 * p.n * p.n * sum(1 for c in sn if c >= p.ns) for _, sn in order) * p.npatt * K
bytes_ = sum((1 + sum(1 for c in sn if c >= p.ns)) * 8.0 * p.n for _, sn in order) * p.npatt * K
print("cond_update, all %d internal nodes x %d classes x %d patterns: %.3f ms  (%.2f TFLOP/s of contraction, %.0f GB/s of vectors)"
      % (nint, K, p.npatt, dt * 1e3, flops / dt / 1e12, bytes_ / dt / 1e9))
a = int(p.sons[p.root, 0]); b = int(p.root)
if a < p.ns:
    a = [c for c in sons[p.root] if c >= p.ns][0]
# vectors for the branch root -> a as minbranches sees them: b's side = the root without a
own = dict(sons); own[b] = [c for c in sons[b] if c != a]
tl.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, [(b, own[b])], None)
t = float(p.branch_t[:, a].max()); mult = p.branch_t[:, a] / t
for want in (True, False):
    for rep in range(3):
        t0 = time.perf_counter()
        out = tl.branch_lnl(p.freqK, a, b, t, mult, p.branch_set[:, a], p.pi[:p.n], 1, want)
        dt = time.perf_counter() - t0
    f = (3 if want else 1) * 2.0 * p.n * p.n * p.npatt * K
    print("branch_lnl (want_dd=%d): %.3f ms  (%.2f TFLOP/s)  -lnL %.6f" % (want, dt * 1e3, f / dt / 1e12, out[0]))
# the eigen-solver on the K class matrices
n = p.n
Qs = []
for k in range(p.nset):
    s = p.set_data[k]; R, U, V = s[:n], s[n:n + n * n].reshape(n, n), s[n + n * n:].reshape(n, n)
    Qs.append((U * R[None, :]) @ V)
for rep in range(3):
    t0 = time.perf_counter()
    tl.eigen_qrev(np.array(Qs), np.tile(p.pi[:n], (len(Qs), 1)))
    dt = time.perf_counter() - t0
print("eigen_qrev, %d matrices: %.3f ms" % (len(Qs), dt * 1e3))
tl.close()
