#!/usr/bin/env python
"""bench.py -- throughput of the tree-likelihood hot path on B200 (BASELINE.json metric).

A "step" is ONE likelihood evaluation (one call of PAML's com.plfun): batched P(t) for every branch, pruning
of every site pattern through the whole tree, root reduction, weighted deterministic sum and -- with N > 1
ranks -- one all-reduce of the scalar (fused into the reduction kernel over NVLink peer memory; NCCL fallback).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (GPU)
    python bench.py --impl reference [...]                          the reference's own CPU path (oracle/_ref)

Headline (`value`, `e2e`, `roofline`): BASELINE config 2, codeml M0, 61 codon states, 8 taxa, 100 000 synthetic site
patterns PER GPU (weak scaling: patterns shard naturally).
`configs`: the other BASELINE shapes, each at its stated TOTAL size sharded over the N ranks (strong scaling):
cfg3_K1..K4 (codeml NSsites M0/M1a/M2a/M3, 32 taxa, 50 000 codon patterns, K site classes), cfg4 (aaml WAG+G4, 64 taxa,
200 000 patterns), cfg5 (baseml GTR+G8, 128 taxa, 1 000 000 patterns) -- updates/s, evaluations/s, the kernel and its
recomputed roofline fraction, the end-to-end figure, the shard-sum check of the all-reduced -lnL and, at N = 1, parity
against the oracle on a 3 000-pattern subsample plus the reference's CPU time on a bounded sample.

Keys of the JSON line (one line, rank 0):
  value / ms_per_step   device-resident throughput: data, tree and parameters already in HBM; per-step device
                        time from CUDA events on the launching stream, L2 flushed before every step, max over ranks.
                        The K timed steps are enqueued straight after W further untimed ones in the same stream, behind a
                        barrier, so that the ranks (which synchronise inside every step's all-reduce) enter step 1 together
  steady                (N > 1) the same loop over 200 steps: median and max step time over ranks
  e2e                   the same metric through the reference-facing C-ABI call pb200_eval with HOST buffers:
                        every step packs the parameters, copies them host->device from pinned memory, runs the
                        kernels, reads -lnL back and synchronises; timed with the host clock around
                        the call (it contains host work), barrier + device sync on both sides, max over ranks
  roofline              dominant kernel (prune_kernel): algorithmic FP64 flops per launch / its device time;
                        peak = FP64 DMMA rate measured live (MEASURED_PEAKS.json has no FP64 figure)
  lnl_check             all-reduced -lnL against the sum over ranks of the shards' own values (relative difference)
  cpu_baseline          the unmodified reference (oracle/_ref/codeml_hook = codeml + lfun timer) on 1 host core, full workload
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

NPATT = 100_000          # per GPU (headline)
SEED = 13147
METRIC = "site_pattern_node_updates_per_sec"
UNIT = "updates/s"
WORKLOAD = "codeml M0 (seqtype=1, 61-state codon, F3x4), 8 taxa unrooted, %d synthetic site patterns per GPU" % NPATT
FLUSH = 256 << 20        # > 126 MB L2
STEADY_STEPS = 200


def n_branches(p):
    Bi = sum(1 for v in range(p.nnode) if v != p.root and p.nson[v] > 0)
    return p.nnode - 1, Bi


def algorithmic_flops(p):
    """flops/eval = K*N*(Bi*2n^2 + B*n + 2n)."""
    n, N, K = p.n, p.npatt, p.nclass
    B, Bi = n_branches(p)
    return K * N * (Bi * 2 * n * n + B * n + 2 * n)


def algorithmic_bytes(p):
    """SURVEY 8d (materialised scheme): bytes/eval = K*N*((I+Bi)*8n) + N*(ns+8)."""
    n, N, K = p.n, p.npatt, p.nclass
    B, Bi = n_branches(p)
    I = p.nnode - p.ns
    return K * N * ((I + Bi) * 8 * n) + N * (p.ns + 8)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (~1 ms period, in-process thread that
    sleeps between samples so it neither holds the GIL nor hammers the driver while kernels are being launched)."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.t, self.h, self.skip = gpu_index, [], False, None, None, 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # CUDA_VISIBLE_DEVICES remaps ordinals; NVML uses physical indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].strip().isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        pw, it = 0.0, 0
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = get_reasons(self.h)
                if it % 8 == 0:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, pw, rs))
                it += 1
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        # the first NVML queries pay one-off driver costs (tens of ms): wait for the loop's steady period BEFORE the
        # barrier that opens the timed region, and count only the samples taken from here on
        t0 = time.time()
        while len(self.rows) < 3 and time.time() - t0 < 1.0:
            time.sleep(0.001)
        self.skip = len(self.rows)

    def stop(self):
        if self.nv is None or self.t is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["NVML unavailable"])
        self.stop_flag = True
        self.t.join(timeout=1)
        nv = self.nv
        try:
            smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            smax = None
        if len(self.rows) > self.skip:
            self.rows = self.rows[self.skip:]
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=smax, reasons=["no samples"])
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown"}
        reasons = set()
        for _, _, rs in self.rows:
            for bit, nm in names.items():
                if rs & bit:
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median([r[0] for r in self.rows])), sm_max_mhz=smax, reasons=sorted(reasons),
                    power_w_max=float(max(r[1] for r in self.rows)), samples=len(self.rows))


def kernel_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of a pruning kernel, per launch, from the committed `ncu --set full`
    capture of the same workload (profiles/*_traffic.json, newest round first); None if absent."""
    for rnd in ("r2", "r1"):
        try:
            j = json.load(open(os.path.join(REPO, "profiles", "%s_%s_traffic.json" % (rnd, name))))
            return j["dram_bytes_read"] + j["dram_bytes_write"]
        except Exception:
            continue
    return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def sub_problem(p, h0, h1):
    from paml_b200.fixture import Problem
    q = Problem(**{f: getattr(p, f) for f in p.__dataclass_fields__})
    q.z = np.ascontiguousarray(p.z[:, h0:h1]); q.fpatt = np.ascontiguousarray(p.fpatt[h0:h1]); q.npatt = q.z.shape[1]
    return q


# ---------------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------------------
def make_problem(cfg, npatt, seed):
    from paml_b200 import simulate
    if cfg == "cfg2":
        return simulate.codon_m0_problem(npatt=npatt, seed=seed)
    if cfg == "cfg3_K1":
        return simulate.codon_m0_problem(ns=32, npatt=npatt, seed=seed, kappa=2.0, omega=0.4, bl=(0.02, 0.2), sim_factor=1.6)
    if cfg == "cfg3_K2":
        return simulate.codon_sites_problem(npatt=npatt, seed=seed, freqs=(0.7, 0.3), omegas=(0.1, 1.0))
    if cfg == "cfg3_K3":
        return simulate.codon_sites_problem(npatt=npatt, seed=seed, freqs=(0.6, 0.3, 0.1), omegas=(0.1, 1.0, 2.5))
    if cfg == "cfg3_K4":
        return simulate.codon_sites_problem(npatt=npatt, seed=seed, freqs=(0.4, 0.3, 0.2, 0.1), omegas=(0.05, 0.3, 1.0, 2.5))
    if cfg == "cfg4":
        return simulate.aa_gamma_problem(npatt=npatt, seed=seed, wag=True)
    if cfg == "cfg5":
        return simulate.nuc_gtr_gamma_problem(npatt=npatt, seed=seed)
    raise ValueError(cfg)


CONFIGS = {
    #            total patterns, base seed, description, reference-arm sample (patterns)
    "cfg3_K1": (50_000, 13149, "codeml NSsites=0 (M0), 32 taxa, 61-state codon, 50 000 patterns, 1 class", 5000),
    "cfg3_K2": (50_000, 13149, "codeml NSsites=1 (M1a), 32 taxa, 61-state codon, 50 000 patterns, 2 site classes", 4000),
    "cfg3_K3": (50_000, 13149, "codeml NSsites=2 (M2a), 32 taxa, 61-state codon, 50 000 patterns, 3 site classes", 3000),
    "cfg3_K4": (50_000, 13149, "codeml NSsites=3 ncatG=4 (M3), 32 taxa, 61-state codon, 50 000 patterns, 4 site classes", 2500),
    "cfg4": (200_000, 13151, "codeml aaml (seqtype=2) WAG + discrete gamma (4 categories, alpha 0.5), 64 taxa, 200 000 patterns", 20000),
    "cfg5": (1_000_000, 1234567, "baseml REV/GTR + discrete gamma (8 categories, alpha 0.5), 128 taxa, 1 000 000 patterns", 50000),
}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference timed by oracle/_ref/{codeml,baseml}_hook
# ---------------------------------------------------------------------------------------------------------
def write_reference_run(cfg, p, d, sample):
    """Input files + control file (+ in.codeml / in.baseml fixing every parameter: one evaluation) for the reference."""
    from paml_b200 import paml_io
    n_used = paml_io.write_pattern_files(p, os.path.join(d, "s"), sample)
    ctl = os.path.join(d, "run.ctl")
    hook = "codeml_hook"
    if cfg in ("cfg2", "cfg3_K1"):
        kappa, omega = (float(p.x[0]), float(p.x[1])) if getattr(p, "x", None) is not None else (2.0, 0.4)
        paml_io.write_codeml_ctl(ctl, "s.phy", "s.tre", kappa=kappa, omega=omega)
    elif cfg == "cfg3_K2":
        paml_io.write_codeml_ctl(ctl, "s.phy", "s.tre", NSsites=1, kappa=2.0, omega=0.4, extra=dict(fix_kappa=0, fix_omega=0, ncatG=2))
        open(os.path.join(d, "in.codeml"), "w").write("-1  2.0  0.7  0.1\n")
    elif cfg == "cfg3_K3":
        paml_io.write_codeml_ctl(ctl, "s.phy", "s.tre", NSsites=2, kappa=2.0, omega=0.4, extra=dict(fix_kappa=0, fix_omega=0, ncatG=3))
        open(os.path.join(d, "in.codeml"), "w").write("-1  2.0  0.6 0.3  0.1 2.5\n")
    elif cfg == "cfg3_K4":
        paml_io.write_codeml_ctl(ctl, "s.phy", "s.tre", NSsites=3, kappa=2.0, omega=0.4, extra=dict(fix_kappa=0, fix_omega=0, ncatG=4))
        open(os.path.join(d, "in.codeml"), "w").write("-1  2.0  0.4 0.3 0.2  0.05 0.3 1.0 2.5\n")
    elif cfg == "cfg4":
        wag = os.path.join(REPO, "oracle", "_ref", "dat", "wag.dat")
        if not os.path.exists(wag):
            return None, 0
        shutil.copy(wag, os.path.join(d, "wag.dat"))
        paml_io.write_codeml_ctl(ctl, "s.phy", "s.tre", seqtype=2, model=2, extra=dict(aaRatefile="wag.dat", fix_alpha=1, alpha=0.5, ncatG=4))
    elif cfg == "cfg5":
        paml_io.write_baseml_ctl(ctl, "s.phy", "s.tre")
        open(os.path.join(d, "in.baseml"), "w").write("-1  0.9 0.3 0.2 0.4 0.25\n")
        hook = "baseml_hook"
    else:
        raise ValueError(cfg)
    return os.path.join(REPO, "oracle", "_ref", hook), n_used


def time_reference(cfg, p, sample, reps, nproc=1, warm=1):
    """Run the unmodified reference (+ lfun timer) on the first `sample` patterns of problem p in `nproc` concurrent
    single-threaded processes.  Returns dict or None if the binaries are absent."""
    work = tempfile.mkdtemp(prefix="pb200_ref_")
    procs = []
    for k in range(nproc):
        d = os.path.join(work, "p%d" % k)
        os.makedirs(d)
        hook, n_used = write_reference_run(cfg, p, d, sample)
        if hook is None or not os.path.exists(hook):
            shutil.rmtree(work, ignore_errors=True)
            return None
        env = dict(os.environ, PBX_TIME=str(reps), PBX_WARM=str(warm), PBX_TIME_OUT=os.path.join(d, "t.json"))
        procs.append((d, n_used, subprocess.Popen([hook, "run.ctl"], cwd=d, env=env, stdin=subprocess.DEVNULL,
                                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    t0 = time.time()
    out = []
    for d, n_used, pr in procs:
        pr.wait()
        try:
            out.append((n_used, json.load(open(os.path.join(d, "t.json")))))
        except Exception:
            shutil.rmtree(work, ignore_errors=True)
            return None
    wall = time.time() - t0
    shutil.rmtree(work, ignore_errors=True)
    I, K = p.nnode - p.ns, p.nclass
    per_proc = [n_used * I * K / j["sec_per_eval_mean"] for n_used, j in out]
    return dict(updates_per_s=float(sum(per_proc)), per_process=float(np.mean(per_proc)), nproc=nproc, sample=out[0][0],
                sec_per_eval=float(np.mean([j["sec_per_eval_mean"] for _, j in out])), reps=reps, wall=wall,
                neg_lnL=out[0][1]["neg_lnL"])


def time_oracle_port(p, sample):
    from oracle import oracle
    q = sub_problem(p, 0, sample)
    oracle.lnl(q)
    t0 = time.time(); r = oracle.lnl(q); dt = time.time() - t0
    return dict(updates_per_s=sample * (p.nnode - p.ns) * p.nclass / dt, sample=sample, sec_per_eval=dt, nproc=1, reps=1, neg_lnL=r["neg_lnL"])


def time_dropin(p):
    """BASELINE metric (ii), complete com.plfun calls per second: the drop-in binary (the reference's own driver, ming2 and
    SetParameters incl. the host Q construction / eigen-solution, with lfun on the GPU) fitting kappa and omega on the
    workload; its in-process profile (PB200_PROFILE) gives the wall time per call and where it goes."""
    import re
    from paml_b200 import paml_io
    exe = os.path.join(REPO, "paml_b200", "bin", "codeml_b200")
    if not os.path.exists(exe):
        return None
    d = tempfile.mkdtemp(prefix="pb200_dropin_")
    paml_io.write_pattern_files(p, d + "/s")
    paml_io.write_codeml_ctl(d + "/run.ctl", "s.phy", "s.tre", kappa=2.0, omega=0.4,
                             extra=dict(fix_kappa=0, fix_omega=0, fix_blength=1))
    try:
        r = subprocess.run([exe, "run.ctl"], cwd=d, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600, env=dict(os.environ, PB200_PROFILE="1"))
    except subprocess.TimeoutExpired:
        return None
    finally:
        pass
    m = re.search(r"then (\d+) evaluations, ([\d.]+) ms each.*SetParameters ([\d.]+) ms, gather ([\d.]+) ms, GPU \(H2D \+ kernels \+ D2H\) "
                  r"([\d.]+) ms, per-pattern outputs ([\d.]+) ms", r.stdout)
    shutil.rmtree(d, ignore_errors=True)
    if r.returncode != 0 or not m:
        return None
    n, each, sp, ga, gpu, outp = int(m.group(1)), *[float(m.group(i)) for i in range(2, 7)]
    memo = re.search(r"eigen-solutions remembered (\d+), solved (\d+)", r.stdout)
    return dict(what="codeml_b200 (reference driver + ming2, lfun on the GPU) fitting branch lengths, kappa and omega on the workload",
                evaluations=n, ms_per_plfun_call=sp + ga + gpu, evals_per_sec=1e3 / (sp + ga + gpu),
                ms=dict(SetParameters_host=sp, gather=ga, gpu_h2d_kernels_d2h=gpu, per_pattern_outputs_final_calls=outp),
                eigen_memo=dict(remembered=int(memo.group(1)), solved=int(memo.group(2))) if memo else None)


def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the box's host cores: the unmodified codeml (single-threaded, it
    has no threads) as one process per physical core, EVERY process evaluating the complete 100 000-pattern workload
    (the config of our arm at N = 1); a step = one lfun evaluation per process; aggregate over the processes."""
    if rank != 0:
        return
    from paml_b200 import simulate
    p = simulate.codon_m0_problem(npatt=args.npatt, seed=SEED)
    ncores = int(os.environ.get("PB200_REF_PROCS", max(1, (os.cpu_count() or 2) // 2)))
    r = time_reference("cfg2", p, p.npatt, reps=args.steps, nproc=ncores, warm=max(1, args.warmup))
    kind = "reference"
    if r is None:
        r = time_oracle_port(p, min(p.npatt, 20000)); kind = "port"; ncores = 1
    line = dict(impl="reference", metric=METRIC, value=r["updates_per_s"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=r["sec_per_eval"] * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, note="reference CPU arm: each step = one lfun evaluation of ALL %d patterns of the "
                            "workload per process; %d independent single-threaded reference processes (the reference has no "
                            "threads), one per physical core, each on its own replica of the workload" % (r["sample"], ncores)),
                cpu_baseline=dict(value=r["updates_per_s"], unit=UNIT, cores=ncores, kind=kind,
                                  sample="all %d patterns x %d evaluations per process" % (r["sample"], args.steps),
                                  single_core_value=r["per_process"] if "per_process" in r else r["updates_per_s"]),
                e2e=dict(value=r["updates_per_s"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                evals_per_sec=ncores / r["sec_per_eval"], neg_lnL=r.get("neg_lnL"))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class Ranks:
    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.world > 1 else None)


def measure(R, tl, solo, p, steps, warmup, sampler=None, steady=0):
    """Time one workload on this rank's shard p: shard-sum check, device-resident loop, end-to-end loop."""
    # this shard's own -lnL from a context without communicator: the all-reduced value must be the sum of these
    solo.set_problem(p)
    part = solo.eval_problem(p)
    tl.set_problem(p)
    for _ in range(warmup):                      # W evaluations through the public call (also makes the parameters resident)
        v = tl.eval_problem(p)
    total = R.sum(part)
    check = abs(v - total) / abs(total)
    tl.bench_resident(2, FLUSH)
    launches0 = tl.launch_count()
    if sampler is not None:
        sampler.start()
    R.barrier()
    res = tl.bench_resident(warmup + steps, FLUSH)          # W more untimed steps, then the K timed ones, one enqueue
    R.barrier()
    nl = tl.launch_count() - launches0
    launches = int(round((nl - (warmup + steps)) * steps / float(warmup + steps)))   # minus the L2-flush launches, timed steps only
    sl = slice(warmup, warmup + steps)
    dev_ms_total = R.max(res["step"][sl].sum())
    out = dict(neg_lnL=v, lnl_check=check, dev_ms_total=dev_ms_total, launches=launches, kernel=tl.last_kernel(),
               prune_ms=float(res["prune"][sl].mean()), pt_ms=float(res["pt"][sl].mean()), mix_ms=float(res["mix"][sl].mean()))
    if steady and R.world > 1:
        R.barrier()
        rs = tl.bench_resident(warmup + steady, FLUSH)["step"][warmup:]
        R.barrier()
        out["steady"] = dict(steps=steady, ms_per_step_median=R.max(float(np.median(rs))), ms_per_step_max=R.max(float(rs.max())),
                             ms_per_step_mean=R.max(float(rs.mean())))
    R.barrier()
    e2e_times, v2 = tl.bench_e2e(p, steps, FLUSH)            # C loop around pb200_eval: flush, sync, [timed] eval with host buffers
    R.barrier()
    out["e2e_total"] = R.max(sum(e2e_times))
    out["e2e_neg_lnL"] = v2
    out["h2d"] = int(8 * (p.nclass + 64 + p.set_data.size + p.branch_t.size) + 4 * p.branch_set.size)
    return out


def config_block(R, tl, solo, cfg, rank, world, steps, warmup, peaks, fp64, want_cpu):
    total_patt, seed, desc, ref_sample = CONFIGS[cfg]
    per = total_patt // world
    t0 = time.time()
    p = make_problem(cfg, per, seed + 101 * rank)
    gen_s = time.time() - t0
    m = measure(R, tl, solo, p, steps, warmup)
    upd_total = R.sum(p.updates_per_eval())
    npatt_total = R.sum(p.npatt)
    if rank != 0:
        return None
    dmma, dfma = fp64
    flops = algorithmic_flops(p)
    ach = flops / (m["prune_ms"] * 1e-3) / 1e12
    tensor = p.n > 4
    peak = dmma if tensor else dfma
    hbm = algorithmic_bytes(p) / (m["prune_ms"] * 1e-3) / 1e9
    blk = dict(workload=desc, scaling="strong" if world > 1 else "single GPU", patterns_total=int(npatt_total), patterns_per_gpu=int(p.npatt),
               taxa=p.ns, states=p.n, classes=p.nclass, internal_nodes=p.ninternal, model=getattr(p, "model_name", None),
               steps=steps, warmup=warmup,
               updates_per_sec=upd_total * steps / (m["dev_ms_total"] * 1e-3), evals_per_sec=steps / (m["dev_ms_total"] * 1e-3),
               ms_per_step=m["dev_ms_total"] / steps,
               e2e=dict(value=upd_total * steps / m["e2e_total"], unit=UNIT, evals_per_sec=steps / m["e2e_total"],
                        ms_per_step=m["e2e_total"] / steps * 1e3, h2d_bytes_per_step=m["h2d"], d2h_bytes_per_step=16),
               kernel=m["kernel"], kernel_ms=dict(pt=m["pt_ms"], prune=m["prune_ms"], mix_allreduce=m["mix_ms"]),
               roofline=dict(bound="tensor" if tensor else "fp64-vector", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak,
                             flops_per_launch=flops, peak_source="FP64 %s rate measured live (pb200_measure_fp64_peak)" % ("DMMA" if tensor else "DFMA"),
                             traffic=kernel_traffic("prune4v2" if p.n == 4 else ("prune_aa" if p.n <= 20 else "prune")),
                             hbm_equivalent=dict(achieved=hbm, peak=peaks.get("hbm_gbs"), unit="GB/s",
                                                 frac=(hbm / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None)),
               neg_lnL=m["neg_lnL"], lnl_check=dict(all_reduced_vs_sum_of_shards_rel=m["lnl_check"], ok=bool(m["lnl_check"] <= 1e-12)),
               data_generation_s=gen_s)
    if want_cpu:
        blk.update(cpu_and_parity(solo, cfg, p, ref_sample))
    return blk


def cpu_and_parity(solo, cfg, p, ref_sample):
    """cpu_baseline leg (rank 0, N = 1): the reference's CPU time on a bounded sample, and the GPU result on a 3 000-pattern
    subsample checked against the oracle (and against the reference binary's own -lnL where its parameters are ours)."""
    out = {}
    nsub = min(3000, p.npatt)
    q = sub_problem(p, 0, nsub)
    solo.set_problem(q)
    g = solo.eval_problem(q)
    glnf = solo.lnf()
    try:
        from oracle import oracle
        r = oracle.lnl(q)
        out["parity"] = dict(patterns=nsub, checker="oracle/liboracle.so (C restatement pinned bit-exact on the reference's fixtures)",
                             neg_lnL_gpu=g, neg_lnL_oracle=r["neg_lnL"], rel=abs(g - r["neg_lnL"]) / abs(r["neg_lnL"]),
                             max_abs_dlnf=float(np.abs(glnf - r["lnf"]).max()), ok=bool(abs(g - r["neg_lnL"]) <= 1e-9 * abs(r["neg_lnL"])))
    except Exception as exc:                      # checker missing on this box: say so, never pretend
        out["parity"] = dict(unavailable=str(exc))
    ref = time_reference(cfg, p, ref_sample, reps=3, nproc=1)
    if ref is not None:
        out["cpu_baseline"] = dict(value=ref["updates_per_s"], unit=UNIT, cores=1, kind="reference",
                                   sample="first %d of the %d patterns, %d evaluations (unmodified %s, gcc -O3)"
                                          % (ref["sample"], p.npatt, ref["reps"], "baseml" if cfg == "cfg5" else "codeml"),
                                   sec_per_eval=ref["sec_per_eval"])
        if cfg == "cfg4":
            # WAG with its own frequencies, alpha fixed, branch lengths from the tree file (8 decimals): the reference's
            # parameters are ours, so its -lnL on the sample is directly comparable
            s = sub_problem(p, 0, ref["sample"])
            solo.set_problem(s)
            gs = solo.eval_problem(s)
            out["parity"]["reference_binary"] = dict(patterns=ref["sample"], neg_lnL_gpu=gs, neg_lnL_reference=ref["neg_lnL"],
                                                     rel=abs(gs - ref["neg_lnL"]) / abs(ref["neg_lnL"]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npatt", type=int, default=NPATT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="all", help="'all', 'none' or a comma list of %s" % ",".join(CONFIGS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from paml_b200 import simulate
    from paml_b200.likelihood import TreeLikelihood, plan_tree

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R = Ranks(torch, dist, world)

    tl = TreeLikelihood(local)         # the timed context: communicator + fused all-reduce when N > 1
    solo = TreeLikelihood(local)       # no communicator: shard-only values (lnL check) and the parity subsample
    if world > 1:
        uid = [TreeLikelihood.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        tl.comm_init(world, rank, uid[0])
        allreduce = "nccl"
        if os.environ.get("PB200_ALLREDUCE", "p2p") == "p2p":
            # fused all-reduce over NVLink peer memory inside the reduction kernel; NCCL stays as the fallback
            handles = [None] * world
            dist.all_gather_object(handles, tl.p2p_local_handle())
            ok = torch.tensor([1], device="cuda")
            try:
                tl.p2p_init(world, rank, handles)
            except Exception as exc:          # e.g. IPC not permitted in this container
                ok[0] = 0
                sys.stderr.write("rank %d: p2p all-reduce unavailable (%s); using NCCL\n" % (rank, exc))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                allreduce = "p2p (fused in mix_kernel, NVLink peer stores)"
            else:                              # some rank could not map its peers: everybody stays on NCCL
                try:
                    tl.p2p_enable(False)
                except Exception:
                    pass
    else:
        allreduce = "none"

    fp64 = tl.measure_fp64_peak()
    peaks = measured_peaks()

    # ---- headline: config 2, weak scaling (every rank simulates its own 100k-pattern shard: different seed = different sites)
    p = simulate.codon_m0_problem(npatt=args.npatt, seed=SEED + 101 * rank)
    sampler = ClockSampler(local)
    m = measure(R, tl, solo, p, args.steps, args.warmup, sampler=sampler, steady=STEADY_STEPS)
    clocks = sampler.stop()
    upd_total = R.sum(p.updates_per_eval())
    value = upd_total * args.steps / (m["dev_ms_total"] * 1e-3)
    e2e_value = upd_total * args.steps / m["e2e_total"]

    line = None
    if rank == 0:
        prog, nslots = plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
        flops = algorithmic_flops(p)
        ach_tf = flops / (m["prune_ms"] * 1e-3) / 1e12
        hbm_gbs = algorithmic_bytes(p) / (m["prune_ms"] * 1e-3) / 1e9
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=m["dev_ms_total"] / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload=WORKLOAD, patterns_per_gpu=int(p.npatt), taxa=p.ns, states=p.n, classes=p.nclass,
                                internal_nodes=p.ninternal, l2="flushed before every timed step (256 MiB write)",
                                parallelism="site-pattern shards, one process per GPU, one all-reduce of the scalar -lnL per step (%s)" % allreduce,
                                on_chip_slots=nslots),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=m["h2d"], d2h_bytes_per_step=16,
                             ms_per_step=m["e2e_total"] / args.steps * 1e3,
                             timer="host clock (clock_gettime inside the library) around each pb200_eval call: pack + H2D from pinned + kernels + all-reduce + D2H + sync; L2 flushed and stream synchronised before each call"),
                    gpu_launches=int(m["launches"]),
                    evals_per_sec=dict(device_resident=args.steps / (m["dev_ms_total"] * 1e-3), e2e=args.steps / m["e2e_total"]),
                    kernel_ms=dict(pt=m["pt_ms"], prune=m["prune_ms"], mix_allreduce=m["mix_ms"]),
                    neg_lnL=m["neg_lnL"],
                    lnl_check=dict(all_reduced_vs_sum_of_shards_rel=m["lnl_check"], ok=bool(m["lnl_check"] <= 1e-12)),
                    roofline=dict(bound="tensor", achieved=ach_tf, peak=fp64[0], unit="TFLOP/s", frac=ach_tf / fp64[0],
                                  traffic=kernel_traffic("prune"), kernel=m["kernel"],
                                  peak_source="FP64 DMMA (mma.sync.m8n8k4.f64) rate measured live by pb200_measure_fp64_peak; "
                                              "MEASURED_PEAKS.json has no FP64 figure (tcgen05 has no FP64 kind)",
                                  flops_per_launch=flops, fp64_dfma_peak=fp64[1],
                                  hbm_equivalent=dict(achieved=hbm_gbs, peak=peaks.get("hbm_gbs"), unit="GB/s",
                                                      frac=(hbm_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                                      note="algorithmic bytes of the materialised scheme (SURVEY 8d); the fused "
                                                           "kernel keeps conditional vectors on chip, so real DRAM traffic is far lower")))
        if "steady" in m:
            st = m["steady"]
            st["value_at_median"] = upd_total / (st["ms_per_step_median"] * 1e-3)
            line["steady"] = st

    # ---- the other BASELINE shapes, sharded over the ranks (strong scaling) ----
    want = [] if args.configs == "none" else (list(CONFIGS) if args.configs == "all" else [c for c in args.configs.split(",") if c in CONFIGS])
    blocks = {}
    want_cpu = (world == 1 and not args.no_cpu_baseline)
    for cfg in want:
        try:
            b = config_block(R, tl, solo, cfg, rank, world, min(args.steps, 10), 3, peaks, fp64, want_cpu)
        except Exception as exc:                 # a failing extra block must not take the headline with it -- but say so loudly
            sys.stderr.write("bench.py: config %s FAILED on rank %d: %r\n" % (cfg, rank, exc))
            if world > 1:
                raise                            # ranks must stay in step: no partial continuation across ranks
            b = dict(failed=repr(exc))
        if rank == 0:
            blocks[cfg] = b

    if rank == 0 and world == 1:
        # the P-stationary, warp-specialised kernel (csrc/prune_ps.cuh; default for small alignments only) on the headline workload
        try:
            os.environ["PB200_KERNEL"] = "ps"
            alt = TreeLikelihood(local)
            alt.set_problem(p)                       # PB200_KERNEL is read when the tree is set
            os.environ.pop("PB200_KERNEL")
            va = alt.eval_problem(p)
            alt.bench_resident(3, FLUSH)
            ra = alt.bench_resident(args.steps, FLUSH)
            ach = algorithmic_flops(p) / (float(ra["prune"].mean()) * 1e-3) / 1e12
            line["variants"] = dict(prune_ps_kernel=dict(kernel=alt.last_kernel(), prune_ms=float(ra["prune"].mean()), ms_per_step=float(ra["step"].mean()),
                                                         roofline_frac=ach / fp64[0], neg_lnL_rel_diff_to_default=abs(va - m["neg_lnL"]) / abs(m["neg_lnL"]),
                                                         note="PB200_KERNEL=ps; the default only for small alignments, where it cuts latency (see DESIGN.md section 4)"))
            alt.close()
        except Exception as exc:
            os.environ.pop("PB200_KERNEL", None)
            line["variants"] = dict(prune_ps_kernel=dict(failed=repr(exc)))
    if rank == 0:
        line["configs"] = blocks
        if want_cpu:
            r = time_reference("cfg2", p, p.npatt, reps=3, nproc=1)
            kind = "reference"
            if r is None:
                r = time_oracle_port(p, 4000); kind = "port"
            line["cpu_baseline"] = dict(value=r["updates_per_s"], unit=UNIT, cores=1, kind=kind,
                                        sample="all %d of the %d patterns, %d lfun evaluations (unmodified codeml, gcc -O3)"
                                               % (r["sample"], p.npatt, r.get("reps", 1)) if kind == "reference" else
                                               "first %d patterns, 1 evaluation of the C restatement" % r["sample"],
                                        sec_per_eval=r["sec_per_eval"], neg_lnL=r.get("neg_lnL"))
            line.update(cpu_and_parity_headline(solo, p))
            d = time_dropin(p)
            if d:
                line["evals_per_sec"]["dropin_plfun"] = d["evals_per_sec"]
                line["dropin"] = d
        print(json.dumps(line))
    tl.close()
    solo.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_and_parity_headline(solo, p):
    nsub = min(3000, p.npatt)
    q = sub_problem(p, 0, nsub)
    solo.set_problem(q)
    g = solo.eval_problem(q)
    try:
        from oracle import oracle
        r = oracle.lnl(q)
        return dict(parity=dict(patterns=nsub, checker="oracle/liboracle.so", neg_lnL_gpu=g, neg_lnL_oracle=r["neg_lnL"],
                                rel=abs(g - r["neg_lnL"]) / abs(r["neg_lnL"]), ok=bool(abs(g - r["neg_lnL"]) <= 1e-9 * abs(r["neg_lnL"]))))
    except Exception as exc:
        return dict(parity=dict(unavailable=str(exc)))


if __name__ == "__main__":
    main()
