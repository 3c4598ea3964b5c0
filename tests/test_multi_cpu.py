"""The N > 1 path on CPU (gloo, world_size 2): shard rule, per-rank partial -lnL, one SUM all-reduce of the scalar,
max-over-ranks timing reduction -- the host-side logic bench.py and the drop-in use around the GPU library.
The per-rank evaluator here is the CPU oracle (this is a test); on the GPU box the same flow runs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN
from paml_b200.fixture import load_fixture, Problem
from paml_b200.shard import shard_bounds, all_bounds


def test_shard_bounds_cover_and_align():
    for npatt in (1, 7, 8, 9, 100, 1000, 100_000, 12_345):
        for world in (1, 2, 3, 4, 8):
            b = all_bounds(npatt, world)
            assert b[0][0] == 0 and b[-1][1] == npatt
            for (a0, a1), (b0, b1) in zip(b, b[1:]):
                assert a1 == b0 and a0 <= a1
            for (h0, h1) in b[:-1]:
                assert h1 % 8 == 0 or h1 == npatt or npatt < 8 * world
            sizes = [h1 - h0 for h0, h1 in b]
            if npatt >= 8 * world:
                assert max(sizes) - min(sizes) <= 8 + npatt % 8


def _worker(rank, world, port, path, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    p = load_fixture(path)
    h0, h1 = shard_bounds(p.npatt, world, rank)
    q = Problem(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    q.z = p.z[:, h0:h1].copy(); q.fpatt = p.fpatt[h0:h1].copy(); q.npatt = h1 - h0
    part = oracle.lnl(q)["neg_lnL"] if q.npatt else 0.0
    t = torch.tensor([part], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                 # the ONE collective of an evaluation
    tm = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)                # timing: max over ranks
    if rank == 0:
        np.save(out, np.array([t.item(), tm.item(), p.ref_neg_lnL]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["mhc_m0", "brown_rev_g8", "stewart_wag_g4"])
def test_two_rank_gloo_sum_matches_reference(tmp_path, name):
    path = [g for g in GOLDEN if g.endswith("/" + name + ".pbfx")][0]
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, path, out), nprocs=2, join=True)
    total, tmax, ref = np.load(out)
    assert abs(total - ref) <= 1e-11 * abs(ref)
    assert abs(tmax - 0.002) < 1e-12
