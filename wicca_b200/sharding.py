"""Multi-GPU sharding of the icon path: images are independent units, so image ``i`` belongs to
rank ``i % world`` (one process per GPU) and nothing crosses between GPUs on the transform path.
The only exchange is a host-side gather of the (small) icons, in input order - the reference's
per-image loop (``classifying_tools.py:312-321``) has no cross-image state either.

Two gathers are offered:

* :class:`IconArena` (one node - the 8 x B200 box): a POSIX shared-memory segment laid out for every
  icon of the whole job.  Each rank page-locks the segment and its GPU writes the icons of its own
  images straight into their final place (``get_small_copies_batch(out=...)``), so the "gather" is
  the D2H copy that had to happen anyway plus one barrier: no pickling, no second copy.
* ``all_gather_object`` (any topology, ``nccl`` or ``gloo``): the portable fallback.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


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


def icon_shape(h: int, w: int, c: int, depth: int) -> tuple[int, int, int]:
    """``(ceil(H/2^d), ceil(W/2^d), C)`` - the shape ``get_small_copy`` returns (``wavelet_coder.py:58-67``)."""
    if depth <= 0:
        return (h, w, c)
    return (-(-h // (1 << depth)), -(-w // (1 << depth)), c)


class IconArena:
    """Shared-memory gather target for the icons of ``len(image_shapes)`` images at ``depths``.

    Every rank of ``group`` constructs it collectively with the same arguments; rank 0 creates the
    segment and broadcasts its name.  ``views(i)`` are the uint8 arrays of image ``i`` (one per depth),
    valid in every rank after :meth:`barrier`.  ``pin=True`` page-locks the segment for the local GPU
    (``wicca_host_register``) so icons are DMA'd into it without a bounce buffer."""

    ALIGN = 256

    def __init__(self, image_shapes: Sequence[tuple[int, int, int]], depths: Sequence[int], group=None, pin: bool = True):
        from multiprocessing import resource_tracker, shared_memory  # noqa: PLC0415

        self.depths = [int(d) for d in depths]
        self.group = group
        self._shapes, self._offsets = [], []
        off = 0
        for (h, w, c) in image_shapes:
            row_s, row_o = [], []
            for d in self.depths:
                sh = icon_shape(int(h), int(w), int(c), d)
                row_s.append(sh)
                row_o.append(off)
                off += -(-int(np.prod(sh)) // self.ALIGN) * self.ALIGN
            self._shapes.append(row_s)
            self._offsets.append(row_o)
        self.nbytes = max(off, self.ALIGN)
        dist = self._dist()
        self.rank = dist.get_rank(group) if dist else 0
        self.world = dist.get_world_size(group) if dist else 1
        self._owner = self.rank == 0
        if self._owner:
            self._shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
        name = [self._shm.name if self._owner else None]
        if dist:
            dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if not self._owner:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:        # only the creator may unlink; keep this process's tracker from doing it at exit
                resource_tracker.unregister(self._shm._name, "shared_memory")  # noqa: SLF001
            except Exception:  # noqa: BLE001
                pass
        self._buf = np.frombuffer(self._shm.buf, dtype=np.uint8, count=self.nbytes)
        self._pinned = False
        if pin:
            from . import _capi  # noqa: PLC0415
            lib = _capi.load()
            _capi.check(lib.wicca_host_register(self._buf.ctypes.data, self.nbytes), "wicca_host_register")
            self._pinned = True

    @staticmethod
    def _dist():
        try:
            import torch.distributed as dist  # noqa: PLC0415
        except Exception:  # noqa: BLE001
            return None
        return dist if dist.is_available() and dist.is_initialized() else None

    def views(self, image: int) -> list[np.ndarray]:
        return [self._buf[o:o + int(np.prod(sh))].reshape(sh) for o, sh in zip(self._offsets[image], self._shapes[image])]

    def barrier(self) -> None:
        dist = self._dist()
        if dist:
            dist.barrier(group=self.group)

    def close(self) -> None:
        if getattr(self, "_shm", None) is None:
            return
        if self._pinned:
            from . import _capi  # noqa: PLC0415
            _capi.load().wicca_host_unregister(self._buf.ctypes.data)
            self._pinned = False
        self.barrier()                      # nobody reads after this point
        self._buf = None
        try:
            self._shm.close()
        except BufferError:
            pass
        if self._owner:
            self._shm.unlink()
        self._shm = None


def sharded_small_copies(get_image: Callable[[int], "np.ndarray"], n_images: int, depths: Sequence[int],
                         transform: Callable, group=None, gather: bool = True, arena: IconArena | None = None):
    """Run ``transform(images, depths) -> list[list[icon]]`` on this rank's shard of the images and
    gather every rank's icons on the host.  Returns the full, input-ordered list when ``gather`` else only
    the local shard.  ``transform`` is normally ``HaarCoder().get_small_copies_batch`` bound to the
    rank's own device.  With ``arena`` (an :class:`IconArena` built for these images) ``transform`` is called as
    ``transform(images, depths, out=...)`` and the gather is a barrier; otherwise it is
    ``torch.distributed.all_gather_object`` (works with the ``nccl`` and ``gloo`` backends)."""
    import torch.distributed as dist  # noqa: PLC0415

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = shard_indices(n_images, rank, world)
    if arena is not None:
        if mine:
            transform([get_image(i) for i in mine], list(depths), out=[arena.views(i) for i in mine])
        if not gather:
            return [arena.views(i) for i in mine]
        arena.barrier()
        return [arena.views(i) for i in range(n_images)]
    local = transform([get_image(i) for i in mine], list(depths)) if mine else []
    if not gather or world == 1:
        return local if not gather else merge_in_order(n_images, [local])
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    return merge_in_order(n_images, parts)
