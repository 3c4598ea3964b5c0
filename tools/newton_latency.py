"""Per-branch latency of the device-side iteration (pb200_branch_newton) against a host-driven loop over pb200_branch_lnl, on the
branch fixtures (small data: the launch / round-trip overheads dominate)."""
import glob, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from paml_b200.fixture import load_fixture
from paml_b200.likelihood import TreeLikelihood
import test_cond_gpu as T

for path in sorted(glob.glob(os.path.join(os.path.dirname(T.__file__), "golden", "branch", "*.pbfx"))):
    p = load_fixture(path); r = dict(p.branch_rec); a, b = r["a"], r["b"]
    t_ = TreeLikelihood(0); t_.set_problem(p)
    sons = T._sons(p); own = dict(sons); own[b] = [c for c in sons[b] if c != a]
    t_.cond_update(p.set_kind, p.nr, p.set_data, p.branch_set, p.branch_t, T._postorder(p, a, sons) + T._postorder(p, b, own), p.scale_flags if p.nscale else None)
    mult = p.branch_t[:, a] / r["branch_len"]; sem = 1 if r["siteclass"] else 0
    start = 2.0 * r["t"] + 0.02; r["t"] = start
    l0 = t_.branch_lnl(p.freqK, a, b, start, mult, p.branch_set[:, a], p.pi[:p.n], sem, False)[0]
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(50): want = T._host_branch_iteration(t_, p, r, mult, sem, l0, False, 1e-6)
        th = (time.perf_counter() - t0) / 50
        t0 = time.perf_counter()
        for _ in range(50): got = t_.branch_newton(p.freqK, a, b, start, l0, mult, p.branch_set[:, a], p.pi[:p.n], sem, False, 1e-6)
        td = (time.perf_counter() - t0) / 50
        t0 = time.perf_counter()
        for _ in range(50): t_.branch_lnl(p.freqK, a, b, start, mult, p.branch_set[:, a], p.pi[:p.n], sem, True)
        t1 = (time.perf_counter() - t0) / 50
    print("%-24s n=%2d npatt=%4d K=%d: %2d evaluations; host-driven loop (Python over the C ABI) %.1f us, device-side %.1f us, one pb200_branch_lnl call %.1f us"
          % (os.path.basename(path)[:-5], p.n, p.npatt, p.nclass, got[4], th * 1e6, td * 1e6, t1 * 1e6))
    t_.close()
