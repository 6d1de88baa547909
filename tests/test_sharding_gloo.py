"""N>1 path on CPU: world_size-2 gloo processes shard the images, transform their shard (the
oracle stands in for the GPU here - this tests the host-side sharding/gather logic, not the
kernels) and gather the icons in input order."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from wicca_b200.sharding import merge_in_order, shard_indices, sharded_small_copies


def test_shard_indices_partition():
    for n in (0, 1, 7, 130):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert merge_in_order(n, [[i * 10 for i in p] for p in parts]) == [i * 10 for i in range(n)]
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
    with pytest.raises(ValueError):
        merge_in_order(5, [[0, 2], [1]])


def _image(i):
    return np.random.default_rng(500 + i).integers(0, 256, (40 + 3 * i, 70 + 5 * i, 3), dtype=np.uint8)


def _oracle_transform(images, depths):
    from oracle import haar_oracle as ho
    return [[ho.haar_icon_blocksum(im, d) for d in depths] for im in images]


def _worker(rank, world, port, n_images, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = sharded_small_copies(_image, n_images, [1, 3], _oracle_transform)
        ok = len(out) == n_images
        for i, row in enumerate(out):
            exp = _oracle_transform([_image(i)], [1, 3])[0]
            ok = ok and all(np.array_equal(a, b) for a, b in zip(row, exp))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_in_order():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 7, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def _oracle_transform_out(images, depths, out=None):
    res = _oracle_transform(images, depths)
    if out is None:
        return res
    for row, dst in zip(res, out):
        for a, d in zip(row, dst):
            assert d.shape == a.shape and d.dtype == np.uint8
            d[...] = a
    return out


def _arena_worker(rank, world, port, n_images, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wicca_b200.sharding import IconArena
        depths = [1, 3]
        arena = IconArena([_image(i).shape for i in range(n_images)], depths, pin=False)   # pin needs a GPU
        out = sharded_small_copies(_image, n_images, depths, _oracle_transform_out, arena=arena)
        ok = len(out) == n_images
        for i, row in enumerate(out):
            exp = _oracle_transform([_image(i)], depths)[0]
            ok = ok and all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(row, exp))
        ret[rank] = ok
        del out, row
        arena.close()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_shared_memory_arena():
    """The one-node gather: every rank writes its own images' icons into a shared segment; after the
    barrier every rank sees all of them, in input order, bit for bit."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_arena_worker, args=(2, port, 9, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_icon_arena_single_process_layout():
    from wicca_b200.sharding import IconArena, icon_shape
    shapes = [(37, 53, 3), (64, 64, 3), (1, 1, 3)]
    arena = IconArena(shapes, [0, 2, 6], pin=False)
    seen = []
    for i, (h, w, c) in enumerate(shapes):
        vs = arena.views(i)
        assert [v.shape for v in vs] == [icon_shape(h, w, c, d) for d in (0, 2, 6)]
        for v in vs:
            v[...] = len(seen) + 1
            seen.append(v)
    for k, v in enumerate(seen):                      # no view overlaps another
        assert (v == k + 1).all()
    del vs, v, seen
    arena.close()
