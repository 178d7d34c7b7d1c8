"""Multi-GPU partitioning of the render: one process per GPU (torchrun), scene replicated, image
tiles dealt to the ranks, one collective at the end.

The image partitions into the Hilbert-ordered bucket list of Renderer::render (src/renderer.rs:138-180);
the buckets are dealt to the ranks by size so that every rank gets about the same number of samples
(psy_render_image_shard computes the same assignment on the device side).  Ranks
render into zero-initialised full-size images, so the gather to rank 0 is a single SUM reduce of
disjoint tiles (NCCL on GPUs; gloo in the CPU tests).
"""
import ctypes as C
import math

from . import abi


def bucket_dim(max_samples_per_bucket, spp):
    """renderer.rs:138-147"""
    target = max_samples_per_bucket / spp
    return 1 if math.sqrt(target) < 1.0 else int(math.sqrt(target))


def bucket_list(width, height, bucket_w, start=(0, 0)):
    """All (x, y, w, h) jobs in the reference's Hilbert order."""
    h = abi.load_host()
    n = h.psyh_bucket_list(width, height, start[0], start[1], bucket_w, bucket_w, None, 0)
    buf = (C.c_uint32 * (4 * max(n, 1)))()
    h.psyh_bucket_list(width, height, start[0], start[1], bucket_w, bucket_w, buf, n)
    return [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]


def buckets_for_rank(buckets, rank, world):
    """The buckets of one rank: largest first, each to the rank with the fewest pixels so far (ties: lowest rank),
    returned in Hilbert order — the assignment psy_render_image_shard computes on the device side."""
    order = sorted(range(len(buckets)), key=lambda j: -(buckets[j][2] * buckets[j][3]))  # stable, like std::stable_sort
    load = [0] * world
    mine = []
    for j in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += buckets[j][2] * buckets[j][3]
        if r == rank:
            mine.append(j)
    return [buckets[j] for j in sorted(mine)]


def gather_image(image, dst=0, group=None):
    """Sums the ranks' disjoint tiles onto `dst` (torch tensor, in place)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(image, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return image
