"""Site-pattern sharding across GPUs: contiguous ranges of (nearly) equal pattern COUNT per rank,
interior boundaries on multiples of 8 patterns (a DMMA m-tile), same rule as the C host shim
(paml_b200/host/plfun_b200.inc).  Pattern weights do not affect cost, so counts -- not weights -- are balanced."""


def shard_bounds(npatt, world, rank):
    """[h0, h1) of `rank` among `world` shards of npatt patterns."""
    if world <= 1 or npatt < 8 * world:
        return (0, npatt) if rank == 0 else (npatt, npatt)
    cut = lambda g: npatt if g >= world else (npatt * g // world) // 8 * 8
    return cut(rank), cut(rank + 1)


def all_bounds(npatt, world):
    return [shard_bounds(npatt, world, r) for r in range(world)]
