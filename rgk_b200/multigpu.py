"""Sharding of the hot path across ranks (one process per GPU).  The scene is replicated; there is no data-path
collective -- only the per-round sum of the partial framebuffers.

round sharding (weak scaling): rank g of N renders round step*N + g; its tiles use seedcount_base = round * ntasks,
    exactly the seeds RenderDriver::RenderFrame would have used for that round (src/render_driver.cpp:160,222).
tile sharding (strong scaling): the centre-sorted task list is dealt round-robin; tile i keeps seed index i, every
    pixel is owned by exactly one rank, so the summed image does not depend on N.
"""


def round_for_rank(step, rank, world):
    return step * world + rank


def seedcount_base(round_index, ntasks):
    return round_index * ntasks


def tiles_for_rank(ntasks, rank, world):
    return list(range(rank, ntasks, world))


def reduce_framebuffer(fb, cnt, dst=0):
    """One sum-reduce of (rgb_sum, count) to `dst` per round; tensors may live on cuda (nccl) or cpu (gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(fb, dst=dst, op=dist.ReduceOp.SUM)
        dist.reduce(cnt, dst=dst, op=dist.ReduceOp.SUM)
    return fb, cnt
