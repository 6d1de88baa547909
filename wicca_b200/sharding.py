"""Multi-GPU sharding of the icon path: images are independent units, so image ``i`` belongs to
rank ``i % world`` (one process per GPU) and nothing crosses between GPUs on the transform path.
The only exchange is a host-side gather of the (small) icons, in input order - the reference's
per-image loop (``classifying_tools.py:312-321``) has no cross-image state either.
"""
from __future__ import annotations

from typing import Callable, Sequence


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the items rank ``rank`` owns (round-robin keeps ragged sizes balanced)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def merge_in_order(n_items: int, shards: Sequence[Sequence]) -> list:
    """Inverse of :func:`shard_indices`: ``shards[r][j]`` is the result of item ``r + j*world``."""
    world = len(shards)
    out = [None] * n_items
    for r, part in enumerate(shards):
        idx = shard_indices(n_items, r, world)
        if len(part) != len(idx):
            raise ValueError(f"rank {r} returned {len(part)} results for {len(idx)} items")
        for i, v in zip(idx, part):
            out[i] = v
    return out


def sharded_small_copies(get_image: Callable[[int], "np.ndarray"], n_images: int, depths: Sequence[int],
                         transform: Callable, group=None, gather: bool = True):
    """Run ``transform(images, depths) -> list[list[icon]]`` on this rank's shard of the images and
    gather every rank's icons on the host (``torch.distributed.all_gather_object``; works with the
    ``nccl`` and ``gloo`` backends).  Returns the full, input-ordered list when ``gather`` else only
    the local shard.  ``transform`` is normally ``HaarCoder().get_small_copies_batch`` bound to the
    rank's own device."""
    import torch.distributed as dist  # noqa: PLC0415

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = shard_indices(n_images, rank, world)
    local = transform([get_image(i) for i in mine], list(depths)) if mine else []
    if not gather or world == 1:
        return local if not gather else merge_in_order(n_images, [local])
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    return merge_in_order(n_images, parts)
