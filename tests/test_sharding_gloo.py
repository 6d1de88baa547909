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
