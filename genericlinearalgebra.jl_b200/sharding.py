"""Host-side partition of the two paths that shard across GPUs (one process per GPU).

batched QR : independent matrices -> contiguous index ranges, NO collective.
TSQR       : contiguous row blocks -> local R (n x n) per rank -> ONE exchange of the R factors
             (all-gather) -> every rank reduces the stack of R factors to the same final R.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """(start, count) of rank's contiguous share of `total` units; the first `total % world` ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def tsqr_R_sharded(local_R, all_gather, reduce_stack):
    """The N>1 TSQR step as data flow: `local_R` (n x n) of this rank -> `all_gather(local_R)` returns the
    (world, n, n) stack -> `reduce_stack(stack)` returns the final R.  On GPUs the three callables are
    gla_dtsqr_local_dev / ncclAllGather / gla_dtsqr_combine_dev (fused in gla_dtsqr_allreduce_dev)."""
    return reduce_stack(all_gather(local_R))
