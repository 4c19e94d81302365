"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path (batch sharding, the single gradient
all-reduce and the 1/world scale handed to the fused optimizer)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from resdepth_b200.lib.distributed import (BucketedAllReduce, allreduce_gradients, allreduce_mean_of_meter,
                                            broadcast_state, loader_is_sharded, owns_batch, replicas_identical,
                                            shard_batch, shard_bounds, state_checksum, sum_partial_rasters, world)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        assert world() == (rank, world_size)
        g = torch.Generator().manual_seed(0)
        full = {'input': torch.randn(5, 3, 8, 8, generator=g), 'target': torch.randn(5, 1, 8, 8, generator=g),
                'dsm_std': torch.full((5,), 3.5), 'tile_size': 8}
        mine = shard_batch(full, rank, world_size)
        lo, hi = shard_bounds(5, rank, world_size)
        assert mine['input'].shape[0] == hi - lo and mine['tile_size'] == 8
        assert torch.equal(mine['input'], full['input'][lo:hi])
        # per-rank "gradient arena": the sum over ranks must equal the gradient of the whole batch
        grads = torch.stack([full['input'][i].sum() * torch.arange(16.) for i in range(lo, hi)]).sum(0)
        scale = allreduce_gradients(grads)
        expect = sum(full['input'][i].sum() for i in range(5)) * torch.arange(16.)
        assert abs(scale - 1.0 / world_size) < 1e-12
        assert torch.allclose(grads, expect, rtol=1e-5, atol=1e-5)
        # tiled inference: batches round-robin over ranks, partial float64 rasters summed by one all-reduce
        weights = torch.arange(1, 8, dtype=torch.float64)             # 7 "batches", each adds its weight to the raster
        raster = torch.zeros(4, 6, dtype=torch.float64)
        for bi in range(7):
            if owns_batch(bi, rank, world_size):
                raster[bi % 4, bi % 6] += weights[bi]
        raster = sum_partial_rasters(raster)
        # replica synchronisation at start-up: different per-rank values -> rank 0's everywhere
        arena = torch.full((10,), float(rank + 1))
        counters = torch.tensor([rank], dtype=torch.int64)
        same_before, _ = replicas_identical(arena)
        broadcast_state([arena, None, counters])
        same_after, cs = replicas_identical(arena)
        assert not same_before and same_after and torch.equal(arena, torch.ones(10)) and int(counters) == 0
        # the step's collective issued slice by slice (CPU tensors: the same sequence, synchronously)
        flat = torch.arange(12.) * (rank + 1)
        red = BucketedAllReduce()
        for lo, hi in ((8, 12), (3, 8), (0, 3)):              # backward order: decoder slice first
            red.launch(flat[lo:hi])
        bscale = red.finish()
        assert bscale == 0.5 and torch.equal(flat, torch.arange(12.) * 3)
        # validation metric: every rank ends with the global mean
        tot, cnt = allreduce_mean_of_meter(2.0 * (rank + 1), rank + 1)
        assert cnt == 3 and abs(tot - 6.0) < 1e-12
        torch.save({'grads': grads, 'scale': scale, 'raster': raster, 'checksum': cs},
                   os.path.join(out_dir, f'rank{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / 'rank0.pt')
    b = torch.load(tmp_path / 'rank1.pt')
    assert torch.equal(a['grads'], b['grads']) and a['scale'] == b['scale'] == 0.5
    expect = torch.zeros(4, 6, dtype=torch.float64)
    for bi in range(7):
        expect[bi % 4, bi % 6] += bi + 1
    assert torch.equal(a['raster'], expect) and torch.equal(b['raster'], expect)


def test_owns_batch_partitions_the_loader():
    for w in (1, 2, 3, 8):
        owners = [[r for r in range(w) if owns_batch(bi, r, w)] for bi in range(20)]
        assert all(len(o) == 1 for o in owners)
    with pytest.raises(ValueError):
        owns_batch(0, 3, 2)
    assert torch.equal(sum_partial_rasters(torch.ones(2, 2, dtype=torch.float64)), torch.ones(2, 2, dtype=torch.float64))


def test_shard_bounds_cover_the_batch():
    for n in (1, 5, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_single_process_is_a_no_op():
    g = torch.ones(8)
    assert allreduce_gradients(g) == 1.0 and torch.equal(g, torch.ones(8))


def test_loader_shard_detection_and_checksum():
    from torch.utils.data import DataLoader, TensorDataset
    from torch.utils.data.distributed import DistributedSampler
    ds = TensorDataset(torch.arange(8.))
    assert not loader_is_sharded(DataLoader(ds, batch_size=2))
    assert not loader_is_sharded([1, 2, 3])
    assert loader_is_sharded(DataLoader(ds, batch_size=2, sampler=DistributedSampler(ds, num_replicas=2, rank=0)))
    a = torch.arange(20.)
    b = a.clone()
    assert state_checksum(a) == state_checksum(b)
    b[3], b[4] = a[4], a[3]                                  # a permutation changes it (position-weighted)
    assert state_checksum(a) != state_checksum(b)
    assert replicas_identical(a) == (True, state_checksum(a))
    assert allreduce_mean_of_meter(3.0, 2) == (3.0, 2)
    red = BucketedAllReduce()
    red.launch(a)
    assert red.finish() == 1.0
