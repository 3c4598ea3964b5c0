#!/usr/bin/env python
"""bench.py -- throughput of the tree-likelihood hot path on B200 (BASELINE.json metric).

A "step" is ONE likelihood evaluation (one call of PAML's com.plfun): batched P(t) for every branch, pruning
of every site pattern through the whole tree, root reduction, weighted deterministic sum and -- with N > 1
ranks -- one NCCL all-reduce of the scalar.  Workload at every N: BASELINE config 2, codeml M0, 61 codon
states, 8 taxa, 100 000 synthetic site patterns PER GPU (weak scaling: patterns shard naturally).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (GPU)
    python bench.py --impl reference [...]                          the reference's own CPU path (oracle/_ref)

Keys of the JSON line (one line, rank 0):
  value / ms_per_step   device-resident throughput: data, tree and parameters already in HBM; per-step device
                        time from CUDA events on the launching stream, L2 flushed before every step, max over ranks
  e2e                   the same metric through the reference-facing C-ABI call pb200_eval with HOST buffers:
                        every step packs the parameters, copies them host->device from pinned memory, runs the
                        kernels, reads -lnL back (8 bytes) and synchronises; timed with the host clock around
                        the call (it contains host work), barrier + device sync on both sides, max over ranks
  roofline              dominant kernel (prune_kernel): algorithmic FP64 flops per launch / its device time;
                        peak = FP64 DMMA rate measured live (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline          the unmodified reference (oracle/_ref/codeml_hook = codeml + lfun timer) on 1 host core
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

NPATT = 100_000          # per GPU
SEED = 13147
METRIC = "site_pattern_node_updates_per_sec"
UNIT = "updates/s"
WORKLOAD = "codeml M0 (seqtype=1, 61-state codon, F3x4), 8 taxa unrooted, %d synthetic site patterns per GPU" % NPATT


def algorithmic_flops(p, plan_internal_children, plan_tip_children):
    """flops/eval = K*N*(Bi*2n^2 + B*n + 2n)."""
    n, N, K = p.n, p.npatt, p.nclass
    Bi = sum(1 for v in range(p.nnode) if v != p.root and p.nson[v] > 0)
    B = p.nnode - 1
    return K * N * (Bi * 2 * n * n + B * n + 2 * n)


def algorithmic_bytes(p):
    """SURVEY 8d (materialised scheme): bytes/eval = K*N*((I+Bi)*8n) + N*(ns+8)."""
    n, N, K = p.n, p.npatt, p.nclass
    Bi = sum(1 for v in range(p.nnode) if v != p.root and p.nson[v] > 0)
    I = p.nnode - p.ns
    return K * N * ((I + Bi) * 8 * n) + N * (p.ns + 8)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (1 ms period, in-process thread);
    falls back to `nvidia-smi -lms` (B200_PROFILING.md recipe) if NVML is unavailable."""

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

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        # the first NVML queries pay one-off driver costs (tens of ms): let the loop reach its steady ~0.1 ms period before
        # the timed region opens, and count only the samples taken from here on
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


def prune_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the pruning kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/r1_prune_traffic.json); None if absent."""
    try:
        j = json.load(open(os.path.join(REPO, "profiles", "r1_prune_traffic.json")))
        return j["dram_bytes_read"] + j["dram_bytes_write"]
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference timed by oracle/_ref/codeml_hook
# ---------------------------------------------------------------------------------------------------------
def time_reference(p, sample, reps, nproc=1, warm=1):
    """Run oracle/_ref/codeml_hook (unmodified codeml + lfun timer) on `sample` patterns of problem p in `nproc`
    concurrent processes (each its own shard of the sample size).  Returns dict or None if binaries are absent."""
    from paml_b200 import paml_io
    hook = os.path.join(REPO, "oracle", "_ref", "codeml_hook")
    if not os.path.exists(hook):
        return None
    work = tempfile.mkdtemp(prefix="pb200_ref_")
    procs = []
    for k in range(nproc):
        d = os.path.join(work, "p%d" % k)
        os.makedirs(d)
        q = p
        if nproc > 1:   # a different shard per process
            lo = (k * sample) % max(1, p.npatt - sample)
            from paml_b200.fixture import Problem
            q = Problem(**{f: getattr(p, f) for f in p.__dataclass_fields__})
            q.z = p.z[:, lo:lo + sample]; q.fpatt = p.fpatt[lo:lo + sample]; q.npatt = q.z.shape[1]
        n_used = paml_io.write_pattern_files(q, os.path.join(d, "s"), sample)
        paml_io.write_codeml_ctl(os.path.join(d, "run.ctl"), "s.phy", "s.tre", kappa=float(p.x[0]), omega=float(p.x[1]))
        env = dict(os.environ, PBX_TIME=str(reps), PBX_TIME_OUT=os.path.join(d, "t.json"))
        procs.append((d, n_used, subprocess.Popen([hook, "run.ctl"], cwd=d, env=env, stdin=subprocess.DEVNULL,
                                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    t0 = time.time()
    out = []
    for d, n_used, pr in procs:
        pr.wait()
        try:
            out.append((n_used, json.load(open(os.path.join(d, "t.json")))))
        except Exception:
            return None
    wall = time.time() - t0
    I = p.nnode - p.ns
    per_proc = [n_used * I / j["sec_per_eval_mean"] for n_used, j in out]
    return dict(updates_per_s=float(sum(per_proc)), per_process=float(np.mean(per_proc)), nproc=nproc, sample=out[0][0],
                sec_per_eval=float(np.mean([j["sec_per_eval_mean"] for _, j in out])), reps=reps, wall=wall,
                neg_lnL=out[0][1]["neg_lnL"])


def time_oracle_port(p, sample):
    from oracle import oracle
    from paml_b200.fixture import Problem
    q = Problem(**{f: getattr(p, f) for f in p.__dataclass_fields__})
    q.z = p.z[:, :sample].copy(); q.fpatt = p.fpatt[:sample].copy(); q.npatt = sample
    oracle.lnl(q)
    t0 = time.time(); oracle.lnl(q); dt = time.time() - t0
    return dict(updates_per_s=sample * (p.nnode - p.ns) / dt, sample=sample, sec_per_eval=dt, nproc=1)


def time_dropin(p):
    """BASELINE metric (ii), complete com.plfun calls per second: the drop-in binary (the reference's own driver, ming2 and
    SetParameters incl. the host Q construction / eigen-solution, with lfun on the GPU) fitting kappa and omega on the
    workload; its in-process profile (PB200_PROFILE) gives the wall time per call and where it goes."""
    import re
    import tempfile
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
    m = re.search(r"then (\d+) evaluations, ([\d.]+) ms each.*SetParameters ([\d.]+) ms, gather ([\d.]+) ms, GPU \(H2D \+ kernels \+ D2H\) "
                  r"([\d.]+) ms, per-pattern outputs ([\d.]+) ms", r.stdout)
    if r.returncode != 0 or not m:
        return None
    n, each, sp, ga, gpu, outp = int(m.group(1)), *[float(m.group(i)) for i in range(2, 7)]
    memo = re.search(r"eigen-solutions remembered (\d+), solved (\d+)", r.stdout)
    return dict(what="codeml_b200 (reference driver + ming2, lfun on the GPU) fitting branch lengths, kappa and omega on the workload",
                evaluations=n, ms_per_plfun_call=sp + ga + gpu, evals_per_sec=1e3 / (sp + ga + gpu),
                ms=dict(SetParameters_host=sp, gather=ga, gpu_h2d_kernels_d2h=gpu, per_pattern_outputs_final_calls=outp),
                eigen_memo=dict(remembered=int(memo.group(1)), solved=int(memo.group(2))) if memo else None)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from paml_b200 import simulate
    sample = 20_000
    p = simulate.codon_m0_problem(npatt=sample, seed=SEED, sim_factor=2.5)
    ncores = int(os.environ.get("PB200_REF_PROCS", max(1, (os.cpu_count() or 2) // 2)))
    r = time_reference(p, sample, reps=args.steps, nproc=ncores)
    kind = "reference"
    if r is None:
        r = time_oracle_port(p, sample); kind = "port"; ncores = 1
    single = time_reference(p, sample, reps=max(2, min(args.steps, 5)), nproc=1) if kind == "reference" else r
    line = dict(impl="reference", metric=METRIC, value=r["updates_per_s"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=r["sec_per_eval"] * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, note="reference CPU arm: each step = one lfun evaluation of a %d-pattern sample "
                            "of the workload per process; %d independent single-threaded reference processes (the reference "
                            "has no threads), one per physical core, each on its own pattern shard" % (r["sample"], ncores)),
                cpu_baseline=dict(value=r["updates_per_s"], unit=UNIT, cores=ncores, kind=kind,
                                  sample="%d patterns x %d evaluations per process" % (r["sample"], args.steps),
                                  single_core_value=single["updates_per_s"]),
                e2e=dict(value=r["updates_per_s"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                evals_per_sec=ncores / r["sec_per_eval"])
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npatt", type=int, default=NPATT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    # every rank simulates its own 100k-pattern shard of the alignment (different seed = different sites)
    p = simulate.codon_m0_problem(npatt=args.npatt, seed=SEED + 101 * rank)
    tl = TreeLikelihood(local)
    tl.set_problem(p)
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    FLUSH = 256 << 20            # > 126 MB L2
    # warm-up: W evaluations through the public call (also makes the parameters resident)
    for _ in range(args.warmup):
        neg_lnL = tl.eval_problem(p)
    tl.bench_resident(args.warmup, FLUSH)

    fp64_dmma, fp64_dfma = tl.measure_fp64_peak()

    # ---- device-resident timed region ----
    sampler = ClockSampler(local)
    launches0 = tl.launch_count()
    barrier()
    sampler.start()
    res = tl.bench_resident(args.steps, FLUSH)
    barrier()
    launches = tl.launch_count() - launches0 - args.steps      # minus the L2-flush launches (outside the timed events)
    dev_ms_total = max_over_ranks(res["step"].sum())
    prune_ms = float(res["prune"].mean()); pt_ms = float(res["pt"].mean()); mix_ms = float(res["mix"].mean())

    # ---- end-to-end timed region: public C-ABI call with host buffers ----
    barrier()
    e2e_times, v = tl.bench_e2e(p, args.steps, FLUSH)     # C loop around pb200_eval: flush, sync, [timed] eval with host buffers
    barrier()
    clocks = sampler.stop()
    e2e_total = max_over_ranks(sum(e2e_times))

    upd_local = p.updates_per_eval()
    upd_total = sum_over_ranks(upd_local)
    value = upd_total * args.steps / (dev_ms_total * 1e-3)
    e2e_value = upd_total * args.steps / e2e_total

    h2d = int(8 * (p.nclass + 64 + p.set_data.size + p.branch_t.size) + 4 * p.branch_set.size)

    if rank == 0:
        prog, nslots = plan_tree(p.nnode, p.ns, p.root, p.nson, p.sons, p.scale_flags)
        flops = algorithmic_flops(p, 0, 0)
        peaks = measured_peaks()
        ach_tf = flops / (prune_ms * 1e-3) / 1e12
        hbm_gbs = algorithmic_bytes(p) / (prune_ms * 1e-3) / 1e9
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=dev_ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload=WORKLOAD, patterns_per_gpu=int(p.npatt), taxa=p.ns, states=p.n, classes=p.nclass,
                                internal_nodes=p.ninternal, l2="flushed before every timed step (256 MiB write)",
                                parallelism="site-pattern shards, one process per GPU, one all-reduce of the scalar -lnL per step (%s)" % allreduce,
                                on_chip_slots=nslots),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
                             ms_per_step=e2e_total / args.steps * 1e3,
                             timer="host clock (clock_gettime inside the library) around each pb200_eval call: pack + H2D from pinned + kernels + all-reduce + D2H + sync; L2 flushed and stream synchronised before each call"),
                    gpu_launches=int(launches),
                    evals_per_sec=dict(device_resident=args.steps / (dev_ms_total * 1e-3), e2e=args.steps / e2e_total),
                    kernel_ms=dict(pt=pt_ms, prune=prune_ms, mix_allreduce=mix_ms),
                    neg_lnL=v,
                    roofline=dict(bound="tensor", achieved=ach_tf, peak=fp64_dmma, unit="TFLOP/s", frac=ach_tf / fp64_dmma,
                                  traffic=prune_traffic(), kernel="prune_kernel<8,16,1,16,true> (4-warp CTAs)",
                                  peak_source="FP64 DMMA (mma.sync.m8n8k4.f64) rate measured live by pb200_measure_fp64_peak; "
                                              "MEASURED_PEAKS.json has no FP64 figure (tcgen05 has no FP64 kind)",
                                  flops_per_launch=flops, fp64_dfma_peak=fp64_dfma,
                                  hbm_equivalent=dict(achieved=hbm_gbs, peak=peaks.get("hbm_gbs"), unit="GB/s",
                                                      frac=(hbm_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                                      note="algorithmic bytes of the materialised scheme (SURVEY 8d); the fused "
                                                           "kernel keeps conditional vectors on chip, so real DRAM traffic is far lower")))
        if world == 1 and not args.no_cpu_baseline:
            sample = 20_000
            r = time_reference(p, sample, reps=3, nproc=1)
            kind = "reference"
            if r is None:
                r = time_oracle_port(p, 4000); kind = "port"
            line["cpu_baseline"] = dict(value=r["updates_per_s"], unit=UNIT, cores=1, kind=kind,
                                        sample="%d of the %d patterns, %d lfun evaluations (unmodified codeml, gcc -O3)"
                                               % (r["sample"], p.npatt, r.get("reps", 1)),
                                        sec_per_eval=r["sec_per_eval"])
            d = time_dropin(p)
            if d:
                line["evals_per_sec"]["dropin_plfun"] = d["evals_per_sec"]
                line["dropin"] = d
        print(json.dumps(line))
    tl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
