"""Round-2 tuning sweep (one process, env knobs re-read by the library at every launch): device-resident pruning-kernel time
of the BASELINE shapes under kernel variants; every variant's -lnL must equal the default's.  Usage: sweep_r2.py [nuc|codon|aa|all]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood

what = sys.argv[1] if len(sys.argv) > 1 else "all"
FLUSH = 256 << 20


def flops(p):
    Bi = sum(1 for v in range(p.nnode) if v != p.root and p.nson[v] > 0)
    return p.nclass * p.npatt * (Bi * 2 * p.n * p.n + (p.nnode - 1) * p.n + 2 * p.n)


def sweep(name, p, variants, peak):
    tl = TreeLikelihood(0)
    tl.set_problem(p)
    ref = None
    for tag, env in variants:
        for k in list(os.environ):
            if k.startswith("PB200_") and k not in ("PB200_LIB",):
                del os.environ[k]
        os.environ.update(env)
        try:
            tl.set_problem(p)                      # a new plan/packing where the variant needs one
            v = tl.eval_problem(p)
            tl.bench_resident(3, FLUSH)
            r = tl.bench_resident(10, FLUSH)
            ms = float(np.median(r["prune"]))
            ok = "" if ref is None or abs(v - ref) <= 1e-12 * abs(ref) else "  LNL MISMATCH %.12g vs %.12g" % (v, ref)
            if ref is None:
                ref = v
            tf = flops(p) / ms / 1e9
            print("%-10s %-44s prune %.4f ms  %.2f TF  frac %.3f  step %.4f ms  [%s]%s" % (name, tag, ms, tf, tf / peak, float(np.median(r["step"])), tl.last_kernel(), ok), flush=True)
        except Exception as exc:
            print("%-10s %-44s FAILED: %s" % (name, tag, exc), flush=True)
    tl.close()


t0 = TreeLikelihood(0)
dmma, dfma = t0.measure_fp64_peak()
t0.close()
print("peaks: DMMA %.2f TF, DFMA %.2f TF" % (dmma, dfma), flush=True)

if what in ("nuc", "all"):
    p = simulate.nuc_gtr_gamma_problem()
    V = [("v2 default (P in smem)", dict(PB200_NUC_PCONST="0")),
         ("PCONST ppt2 256x2", dict()),
         ("PCONST ppt2 256x2, 1 class/launch", dict(PB200_NUC_CLS_PER_LAUNCH="1")),
         ("PCONST ppt2 128x4", dict(PB200_NUC_PPT="21")),
         ("PCONST ppt2 384x2", dict(PB200_NUC_PPT="23")),
         ("PCONST ppt1 512x2", dict(PB200_NUC_PPT="13")),
         ("PCONST ppt1 256x3", dict(PB200_NUC_PPT="1")),
         ("PCONST ppt3 128x3", dict(PB200_NUC_PPT="3")),
         ("PCONST ppt4 128x2", dict(PB200_NUC_PPT="4")),
         ("PCONST ppt4 256x1", dict(PB200_NUC_PPT="41"))]
    sweep("cfg5", p, V, dfma)
if what == "nuc2":
    p = simulate.nuc_gtr_gamma_problem()
    V = [("PCONST ppt4 128x2", dict(PB200_NUC_PPT="4")),
         ("v2 ppt4 128x2 (P in smem)", dict(PB200_NUC_PPT="4", PB200_NUC_PCONST="0")),
         ("PCONST ppt4 128x3", dict(PB200_NUC_PPT="43")),
         ("PCONST ppt4 64x4", dict(PB200_NUC_PPT="44")),
         ("PCONST ppt4 64x6", dict(PB200_NUC_PPT="46")),
         ("PCONST ppt6 128x2", dict(PB200_NUC_PPT="6")),
         ("PCONST ppt8 64x2", dict(PB200_NUC_PPT="8")),
         ("PCONST ppt8 64x4", dict(PB200_NUC_PPT="84")),
         ("PCONST ppt3 256x2", dict(PB200_NUC_PPT="32")),
         ("PCONST ppt4 128x2, 1 class/launch", dict(PB200_NUC_PPT="4", PB200_NUC_CLS_PER_LAUNCH="1")),
         ("v2 ppt8 64x4 (P in smem)", dict(PB200_NUC_PPT="84", PB200_NUC_PCONST="0")),
         ("v2 ppt6 128x2 (P in smem)", dict(PB200_NUC_PPT="6", PB200_NUC_PCONST="0"))]
    sweep("cfg5", p, V, dfma)
if what == "aa2":
    p = simulate.aa_gamma_problem(npatt=200_000, wag=True)
    V = [("default (MT=2)", dict()), ("MT=3", dict(PB200_AA_MT="3")), ("MT=4", dict(PB200_AA_MT="4")),
         ("MT=4, 8-warp CTAs", dict(PB200_AA_MT="4", PB200_WARPS="8")), ("MT=3, 8-warp CTAs", dict(PB200_AA_MT="3", PB200_WARPS="8")),
         ("MT=4, smem 227", dict(PB200_AA_MT="4", PB200_SMEM_KB="227", PB200_WARPS="12"))]
    sweep("cfg4", p, V, dmma)
if what == "codon3":
    p = simulate.codon_m0_problem(npatt=100_000, seed=13147)
    V = [("default", dict())] + [("stagger CTAs %d ns" % ns, dict(PB200_STAGGER_CTAS="1", PB200_STAGGER_NS=str(ns))) for ns in (500, 1000, 2000, 4000, 8000, 16000)]
    V += [("16-warp CTA stagger %d ns" % ns, dict(PB200_WARPS="16", PB200_STAGGER_NS=str(ns))) for ns in (1000, 4000)]
    sweep("cfg2", p, V, dmma)
    p = simulate.codon_sites_problem(npatt=50_000, seed=13149, freqs=(0.4, 0.3, 0.2, 0.1), omegas=(0.05, 0.3, 1.0, 2.5))
    sweep("cfg3_K4", p, V[:5], dmma)
if what == "codon2":
    p = simulate.codon_m0_problem(npatt=100_000, seed=13147)
    V = [("default", dict()), ("regs for 20 warps/SM", dict(PB200_MAXW="20")), ("regs for 24 warps/SM", dict(PB200_MAXW="24"))]
    sweep("cfg2", p, V, dmma)
    q = simulate.Problem(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    q.z = np.zeros_like(p.z)
    sweep("cfg2 z=0", q, V[:1], dmma)             # every tip gather hits the same row: tip phases without their L2 traffic
    q.z = (np.arange(p.npatt)[None, :] // 8 % 61 + np.zeros((p.ns, 1), int)).astype(np.uint8)
    sweep("cfg2 z=tile", q, V[:1], dmma)          # the 8 patterns of a warp share their codes: one row per gather instruction
    p = simulate.codon_sites_problem(npatt=50_000, seed=13149, freqs=(0.4, 0.3, 0.2, 0.1), omegas=(0.05, 0.3, 1.0, 2.5))
    sweep("cfg3_K4", p, V, dmma)
if what in ("codon", "all"):
    p = simulate.codon_m0_problem(npatt=100_000, seed=13147)
    V = [("default", dict()),
         ("static items", dict(PB200_STATIC_ITEMS="1")),
         ("8-warp CTAs", dict(PB200_WARPS="8")),
         ("16-warp CTAs", dict(PB200_WARPS="16")),
         ("16-warp CTAs, pipe lock 2", dict(PB200_WARPS="16", PB200_PIPE_LOCK="2")),
         ("MT=2 (16 patterns/warp, 255 regs)", dict(PB200_MT_SMALL="0")),
         ("smem budget 100 KB", dict(PB200_SMEM_KB="100")),
         ("carveout 25", dict(PB200_CARVEOUT="25")),
         ("carveout 50", dict(PB200_CARVEOUT="50")),
         ("tma variant", dict(PB200_KERNEL="tma"))]
    sweep("cfg2", p, V, dmma)
    p = simulate.codon_sites_problem(npatt=50_000, seed=13149, freqs=(0.4, 0.3, 0.2, 0.1), omegas=(0.05, 0.3, 1.0, 2.5))
    sweep("cfg3_K4", p, V[:6], dmma)
if what in ("aa", "all"):
    p = simulate.aa_gamma_problem(npatt=200_000, wag=True)
    V = [("default", dict()), ("MT small", dict(PB200_MT_SMALL="1")), ("8-warp CTAs", dict(PB200_WARPS="8")),
         ("16-warp CTAs", dict(PB200_WARPS="16")), ("static items", dict(PB200_STATIC_ITEMS="1"))]
    sweep("cfg4", p, V, dmma)
