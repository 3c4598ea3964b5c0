"""Time the phases of the pruning kernel in isolation with hand-made programs (codon shape, 1e5 patterns)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood, _ptr, check

p = simulate.codon_m0_problem(npatt=100_000, seed=13147)
tl = TreeLikelihood(0)
tl.set_problem(p)
tl.eval_problem(p)
base = tl.bench_resident(5, 0)["prune"].mean()
ns = p.ns
internal = [v for v in range(p.nnode) if p.nson[v] > 0 and v != p.root]

def prog_of(ops, tips):
    mats = [k[2] for (children, _) in ops for k in children if k[0] != 2]
    out = [len(ops), 2, len(tips), 6 + len(tips) + len(mats), len(mats), 6 + len(tips)] + list(tips) + mats
    for (children, out_slot) in ops:
        ntip = sum(1 for k in children if k[0] == 0)
        out += [len(children), out_slot, -1, ns, 0, ntip]
        for k in children: out += list(k)
    return np.array(out, np.int32)

def run(name, ops, tips, nint, ntipc):
    prog = prog_of(ops, tips)
    check(tl._L.pb200_debug_set_program(tl._ctx, _ptr(prog, C.c_int), C.c_int(len(prog)), C.c_int(2), C.c_int(0)))
    tl.bench_resident(2, 0)
    ms = tl.bench_resident(5, 0)["prune"].mean()
    flops = 1e5 * nint * 2 * 61 * 61
    print("%-46s %.4f ms  (%d internal, %d tip children per pattern)  DMMA-equivalent %.1f TF" % (name, ms, nint, ntipc, flops / ms / 1e9))

print("real cfg2 program: %.4f ms" % base)
NI = 20
run("internal only, one P (L1 resident)", [([(1, 0, internal[0])] * NI, 1), ([(1, 1, internal[0])], -1)], [], NI + 1, 0)
run("internal only, 5 P's cycling", [([(1, 0, internal[i % len(internal)]) for i in range(NI)], 1), ([(1, 1, internal[0])], -1)], [], NI + 1, 0)
tips = [i % ns for i in range(16)]
ops = [([(0, tips[2 * i], tips[2 * i]), (0, tips[2 * i + 1], tips[2 * i + 1])], 0) for i in range(8)]
run("tip pairs only (8 ops x 2 tips) + 1 internal", ops + [([(1, 0, internal[0])], -1)], tips, 1, 16)
ops = [([(0, tips[i], tips[i]), (1, 0, internal[i % len(internal)])], 0) for i in range(8)]
run("caterpillar: 8 x (1 tip + 1 internal)", ops + [([(1, 0, internal[0])], -1)], tips[:8], 9, 8)
