"""Worker for tests/test_multi_gpu.py: launched with torchrun, one rank per GPU.  Every rank uploads its shard,
joins the NCCL communicator through the C ABI and evaluates; the all-reduced -lnL must be the whole-alignment value."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paml_b200 import simulate
from paml_b200.likelihood import TreeLikelihood
from paml_b200.shard import shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = simulate.codon_m0_problem(npatt=3000, seed=77)
tl = TreeLikelihood(local)
h0, h1 = shard_bounds(p.npatt, world, rank)
tl.set_problem(p, h0, h1)
part = tl.eval_problem(p)                       # before the communicator exists: this rank's shard only
uid = [TreeLikelihood.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
tl.comm_init(world, rank, uid[0])
total = tl.eval_problem(p)                      # now ends with one ncclAllReduce of the scalar
# the same reduction fused into the mixture kernel over NVLink peer memory
handles = [None] * world
dist.all_gather_object(handles, tl.p2p_local_handle())
tl.p2p_init(world, rank, handles)
totals_p2p = [tl.eval_problem(p) for _ in range(5)]
parts = [None] * world
dist.all_gather_object(parts, part)
allp2p = [None] * world
dist.all_gather_object(allp2p, totals_p2p[0])
assert len(set(allp2p)) == 1, "ranks disagree on the fused all-reduce result"
if rank == 0:
    whole = TreeLikelihood(local)
    whole.set_problem(p)
    ref = whole.eval_problem(p)
    ok = abs(total - ref) <= 1e-12 * abs(ref) and abs(sum(parts) - ref) <= 1e-12 * abs(ref)
    ok = ok and all(abs(t - ref) <= 1e-12 * abs(ref) for t in totals_p2p) and len(set(totals_p2p)) == 1
    whole.close()
# a peer that never arrives: the fused all-reduce must give up and the call must fail loudly (no hang inside the kernel).
# New hand-shake (clears the exchange buffers, restarts the evaluation numbers) with a 300 ms limit; only rank 0 evaluates.
os.environ["PB200_P2P_TIMEOUT_MS"] = "300"
dist.barrier()
handles = [None] * world
dist.all_gather_object(handles, tl.p2p_local_handle())
tl.p2p_init(world, rank, handles)
again = tl.eval_problem(p)                      # everybody: still the right sum after the re-initialisation
dist.barrier()
timed_out = True
if rank == 0:
    from paml_b200.likelihood import PB200Error
    try:
        tl.eval_problem(p)
        timed_out = False
    except PB200Error as exc:
        timed_out = "did not arrive" in str(exc)
    ok = ok and timed_out and abs(again - ref) <= 1e-12 * abs(ref)
    print("MULTI_RANK_RESULT", world, total, ref, sum(parts), "OK" if ok else "MISMATCH", "timeout-detected" if timed_out else "NO-TIMEOUT")
dist.barrier()
tl.close()
dist.barrier()
dist.destroy_process_group()
