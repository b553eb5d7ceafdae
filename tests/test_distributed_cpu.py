"""world_size-2 gloo tests of the multi-GPU plumbing (temgymcore_b200/distributed.py) on CPU:
shard arithmetic, the coefficient-table broadcast and the row all-gather.  The per-rank
compute is injected (rows come from the CPU oracle) because there is no GPU here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from temgymcore_b200 import distributed as D


def test_shard_range_partitions():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_row_shards_are_tile_aligned_and_cover():
    for H in (1, 31, 32, 100, 1024, 2048, 2050):
        for world in (1, 2, 4, 8):
            sh = D.row_shards(H, world)
            assert sum(n for _, n in sh) == H
            pos = 0
            for r0, n in sh:
                assert r0 == pos or n == 0
                if n:
                    assert r0 % D.ROW_ALIGN == 0
                pos += n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, nb, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import temgym_oracle as O
        from tests import models as M
        g, model = M.aperture_diffraction_case(nb, (H, W))
        grid = model[-1]

        def table_fn():
            # rank 0 owns the true table; other ranks start with garbage that the broadcast
            # must overwrite
            t = torch.arange(nb * 12, dtype=torch.float64).reshape(nb, 12)
            return (t if rank == 0 else torch.full((nb, 12), -1.0, dtype=torch.float64)), nb, "cpu"

        seen = {}

        def rows_fn(poly, n, row0, nrows):
            seen["poly"] = poly.clone()
            full = torch.from_numpy(O.make_gaussian_image(g, model))
            return full[row0:row0 + nrows].contiguous()

        img = D.make_gaussian_image_sharded(g, model, rows_fn=rows_fn, table_fn=table_fn)
        ref = torch.from_numpy(O.make_gaussian_image(g, model))
        ok_table = bool((seen["poly"] == torch.arange(nb * 12, dtype=torch.float64).reshape(nb, 12)).all())
        ret[rank] = (ok_table, bool(torch.equal(img, ref)), tuple(img.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H,W", [(96, 40), (70, 33)])
def test_sharded_image_plumbing_gloo(H, W):
    world, nb = 2, 24
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, W, nb, ret), nprocs=world, join=True)
    for r in range(world):
        ok_table, ok_img, shape = ret[r]
        assert ok_table, "broadcast did not deliver rank 0's table"
        assert ok_img and shape == (H, W)


class _SharedMemPeerImage:
    """CPU stand-in for distributed.PeerImage with the same interface: one full image per rank in shared host
    memory (the peers' images 'mapped' into every process), field_sum writes this rank's row block into ALL of
    them, barrier is a gloo barrier.  The compute is the CPU oracle."""

    def __init__(self, images, rank, world, rows_of):
        self.images, self.rank, self.world = images, rank, world
        self.H, self.W = images[0].shape
        self.image = images[rank]
        self.rows_of = rows_of
        self.calls = []

    def field_sum(self, poly, nb, grid, *, cull_bits=None, method="auto"):
        r0, nr = D.row_shards(self.H, self.world)[self.rank]
        block = self.rows_of(r0, nr)
        for img in self.images:                       # the kernels' peer stores
            img[r0:r0 + nr] = block
        self.calls.append((r0, nr, method))
        return r0, nr

    def barrier(self):
        dist.barrier()

    def check_grid(self, grid, out_dtype=None):
        assert tuple(int(v) for v in grid.shape) == (self.H, self.W)

    def step(self, poly, nb, grid, *, cull_bits=None, method="auto"):
        self.field_sum(poly, nb, grid, cull_bits=cull_bits, method=method)
        self.barrier()
        return self.image


def _peer_worker(rank, world, port, images, nb, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import temgym_oracle as O
        from tests import models as M
        H, W = images[0].shape
        g, model = M.aperture_diffraction_case(nb, (H, W))
        ref = torch.from_numpy(O.make_gaussian_image(g, model))

        def table_fn():                               # every rank builds the table itself: nothing is broadcast
            return torch.full((nb, 12), float(rank), dtype=torch.float64), nb, "cpu"

        pi = _SharedMemPeerImage(images, rank, world, lambda r0, nr: ref[r0:r0 + nr])
        out = D.make_gaussian_image_sharded(g, model, method="sfu", table_fn=table_fn, peer_image=pi)
        dist.barrier()
        ret[rank] = (bool(torch.equal(out, ref)), out.data_ptr() == images[rank].data_ptr(), pi.calls)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H,W", [(96, 40), (70, 33)])
def test_fused_peer_image_plumbing_gloo(H, W):
    """make_gaussian_image_sharded(peer_image=...): no broadcast, every rank's row block lands in every rank's
    image, the barrier closes the step and each rank returns its own (complete) image."""
    world, nb = 2, 24
    images = [torch.zeros((H, W), dtype=torch.complex128).share_memory_() for _ in range(world)]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_peer_worker, args=(world, port, images, nb, ret), nprocs=world, join=True)
    shards = D.row_shards(H, world)
    for r in range(world):
        ok_img, own, calls = ret[r]
        assert ok_img and own
        assert calls == [(shards[r][0], shards[r][1], "sfu")]
