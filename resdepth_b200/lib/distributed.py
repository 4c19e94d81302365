"""Data-parallel plumbing of the hot path (one process per GPU, ``torch.distributed``).

ResDepth itself is single-device (reference lib/Trainer.py:34).  Tiles are independent, so the training step
shards over ranks by batch: every rank runs forward/backward on its own tiles with a full parameter replica
(broadcast from rank 0 when the trainer starts) and the flat gradient arena is summed over ranks -- one all-reduce
(NCCL over NVLink on GPUs, gloo in the CPU tests) issued in three slices as the backward pass completes them, so
that the collective overlaps the rest of the backward pass -- before the fused Adam step, which folds the
1/world_size scale into the update.  BatchNorm statistics stay per rank (as in plain DDP).  The validation metric
is averaged over ranks so that every replica takes the same scheduler / checkpoint decisions.  Nothing else
crosses ranks.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_tiles: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n_tiles`` over ranks: rank r gets [lo, hi); the first ``n % world`` ranks
    get one extra tile."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank/world_size {rank}/{world_size}')
    base, extra = divmod(n_tiles, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world_size: int) -> Dict[str, torch.Tensor]:
    """The slice of a DataLoader batch dict (reference lib/DsmOrthoDataset.py:281-291) owned by ``rank``."""
    n = batch['input'].shape[0]
    lo, hi = shard_bounds(n, rank, world_size)
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def owns_batch(batch_index: int, rank: int, world_size: int) -> bool:
    """Tiled inference shards the DataLoader's batches round-robin over ranks (tiles are independent; the partial
    rasters are summed once at the end, ``sum_partial_rasters``)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank/world_size {rank}/{world_size}')
    return batch_index % world_size == rank


def sum_partial_rasters(raster: torch.Tensor) -> torch.Tensor:
    """Blended float64 raster of one rank -> the complete raster on every rank (one all-reduce; no-op on one rank).
    Linear blending is a sum of weighted tiles, so the partial rasters of disjoint tile subsets simply add."""
    _, world_size = world()
    if world_size > 1:
        dist.all_reduce(raster, op=dist.ReduceOp.SUM)
    return raster


def loader_is_sharded(loader) -> bool:
    """True when the DataLoader already hands every rank its own tiles (a ``DistributedSampler`` or any sampler /
    batch sampler that knows ``num_replicas``); the trainer then takes the batches as they come."""
    for attr in ('sampler', 'batch_sampler'):
        smp = getattr(loader, attr, None)
        inner = getattr(smp, 'sampler', None)
        for cand in (smp, inner):
            if cand is not None and getattr(cand, 'num_replicas', 1) > 1:
                return True
    return False


def broadcast_state(tensors, src: int = 0) -> None:
    """Makes every rank start from rank ``src``'s values (parameter arena, BatchNorm buffers, counters): replicas
    built from different seeds or different checkpoints would otherwise drift apart silently.  No-op on one rank."""
    _, world_size = world()
    if world_size == 1:
        return
    for t in tensors:
        if t is not None and t.numel() > 0:
            dist.broadcast(t, src=src)


def allreduce_mean_of_meter(total: float, count: int, device=None) -> Tuple[float, int]:
    """(sum, count) of a per-rank average meter -> the global (sum, count): the validation metric that drives
    ReduceLROnPlateau, the best-model decision and the checkpoint must be the same number on every rank."""
    _, world_size = world()
    if world_size == 1:
        return total, count
    dev = device if (device is not None and dist.get_backend() == 'nccl') else 'cpu'
    t = torch.tensor([total, float(count)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0].item()), int(round(float(t[1].item())))


def state_checksum(*tensors) -> float:
    """Order-dependent float64 checksum of flat tensors (sum of x * (1 + (i mod 7))): equal on replicas that hold the
    same bits, different as soon as one element moves."""
    acc = 0.0
    for t in tensors:
        if t is None or t.numel() == 0:
            continue
        f = t.detach().double().flatten()
        w = (torch.arange(f.numel(), device=f.device) % 7 + 1).double()
        acc += float((f * w).sum().item())
    return acc


def replicas_identical(*tensors, device=None) -> Tuple[bool, float]:
    """(all ranks hold the same checksum, this rank's checksum); (True, checksum) on one rank."""
    cs = state_checksum(*tensors)
    _, world_size = world()
    if world_size == 1:
        return True, cs
    dev = device if (device is not None and dist.get_backend() == 'nccl') else 'cpu'
    lo = torch.tensor([cs], dtype=torch.float64, device=dev)
    hi = lo.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(lo.item() == hi.item()), cs


class BucketedAllReduce:
    """The gradient all-reduce of one step, issued slice by slice as the backward pass completes them.

    ``launch(slice)`` is called right after a backward stage has been enqueued on the compute stream: the slice's
    all-reduce is enqueued on a communication stream that waits for that point only, so it overlaps the stages that
    follow.  ``finish()`` makes the compute stream wait for every slice and returns the 1/world_size scale for the
    optimizer.  CPU tensors (gloo tests) run the same sequence synchronously."""

    def __init__(self):
        self._works = []
        self._comm_stream = None

    def launch(self, flat_slice: torch.Tensor) -> None:
        _, world_size = world()
        if world_size == 1:
            return
        if not flat_slice.is_cuda:
            dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM)
            return
        dev = flat_slice.device
        if self._comm_stream is None or self._comm_stream.device != dev:
            self._comm_stream = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        self._comm_stream.wait_stream(main)
        with torch.cuda.stream(self._comm_stream):
            self._works.append(dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, async_op=True))

    def finish(self) -> float:
        _, world_size = world()
        for w in self._works:
            w.wait()                      # orders the current (compute) stream after the collective
        if self._works and self._comm_stream is not None:
            torch.cuda.current_stream(self._comm_stream.device).wait_stream(self._comm_stream)
        self._works = []
        return 1.0 / world_size


def allreduce_gradients(flat_grads: torch.Tensor) -> float:
    """Sums the flat gradient arena over all ranks in place (the ONE collective of the path) and returns the
    scale the optimizer must apply to turn the sum into the data-parallel mean."""
    rank, world_size = world()
    if world_size == 1:
        return 1.0
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / world_size
