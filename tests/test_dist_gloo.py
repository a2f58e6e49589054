"""The N>1 path on CPU: two gloo ranks shard a height map into strips, exchange the
halo row with the ring helper the GPU path uses, and the assembled result equals the
whole-image oracle.  Also the round-robin graph sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from kanter_core_b200 import dist as kdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hgt = np.random.default_rng(3).random((h, w), dtype=np.float32)  # every rank can rebuild the input
        y0, y1 = kdist.strip_rows(h, rank, world)
        strip = hgt[y0:y1]
        halo = kdist.ring_halo_rows(torch.from_numpy(strip[-1].copy())).numpy()
        assert np.array_equal(halo, hgt[(y0 - 1) % h]), "rank %d got the wrong halo row" % rank
        planes = oracle.height_to_normal_strip(strip, h, halo)
        np.save(os.path.join(out_dir, "strip_%d.npy" % rank), np.stack(planes))
        t = kdist.max_over_ranks(float(rank + 1))
        assert t == float(world)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 64), (2, 37), (3, 50)])
def test_strip_sharding_with_halo_exchange_matches_whole_image(tmp_path, world, h):
    w = 48
    mp.spawn(_worker, args=(world, _free_port(), h, w, str(tmp_path)), nprocs=world, join=True)
    hgt = np.random.default_rng(3).random((h, w), dtype=np.float32)
    want = np.stack(oracle.height_to_normal(hgt))
    got = np.concatenate([np.load(tmp_path / ("strip_%d.npy" % r)) for r in range(world)], axis=1)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_strip_rows_cover_the_image_exactly():
    for h in (1, 7, 64, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            rows = [kdist.strip_rows(h, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1


def test_shard_units_round_robin():
    for world in (1, 2, 4, 8):
        got = sorted(sum((kdist.shard_units(64, r, world) for r in range(world)), []))
        assert got == list(range(64))
        assert all(len(kdist.shard_units(64, r, world)) == 64 // world for r in range(world))


def test_oracle_strip_equals_whole_image():
    hgt = np.random.default_rng(5).random((40, 33), dtype=np.float32)
    want = oracle.height_to_normal(hgt)
    for (y0, y1) in [(0, 13), (13, 27), (27, 40)]:
        got = oracle.height_to_normal_strip(hgt[y0:y1], 40, hgt[(y0 - 1) % 40])
        for c in range(3):
            assert np.array_equal(got[c], want[c][y0:y1])
