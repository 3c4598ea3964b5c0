"""A/B of the two all-reduce paths inside one torchrun session (same box, same data): NCCL call vs the all-reduce
fused into mix_kernel over NVLink peer memory.  Prints per-mode mean step time (max over ranks)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = simulate.codon_m0_problem(npatt=100_000, seed=13147 + 101 * rank)
tl = TreeLikelihood(local); tl.set_problem(p)
uid = [TreeLikelihood.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0); tl.comm_init(world, rank, uid[0])
handles = [None] * world; dist.all_gather_object(handles, tl.p2p_local_handle()); tl.p2p_init(world, rank, handles)
def run(mode, steps=40):
    tl.p2p_enable(mode == "p2p")
    tl.eval_problem(p); tl.bench_resident(3, 256 << 20)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    r = tl.bench_resident(steps, 256 << 20)
    sec, _ = tl.bench_e2e(p, steps, 256 << 20)
    t = torch.tensor([r["step"].mean(), r["mix"].mean(), sec.mean() * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
for mode in ["nccl", "p2p", "nccl", "p2p", "nccl", "p2p"]:
    res = run(mode)
    if rank == 0:
        print("%-5s N=%d  step %.4f ms  mix+allreduce %.4f ms  e2e %.4f ms" % (mode, world, *res))
tl.close(); dist.barrier(); dist.destroy_process_group()
