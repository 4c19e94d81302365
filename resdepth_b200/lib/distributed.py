"""Data-parallel plumbing of the hot path (one process per GPU, ``torch.distributed``).

ResDepth itself is single-device (reference lib/Trainer.py:34).  Tiles are independent, so the training step
shards over ranks by batch: every rank runs forward/backward on its own tiles with a full parameter replica
and the flat gradient arena is summed with ONE all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests)
before the fused Adam step, which folds the 1/world_size scale into the update.  BatchNorm statistics stay
per rank (as in plain DDP).  Nothing else crosses ranks.
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


def allreduce_gradients(flat_grads: torch.Tensor) -> float:
    """Sums the flat gradient arena over all ranks in place (the ONE collective of the path) and returns the
    scale the optimizer must apply to turn the sum into the data-parallel mean."""
    rank, world_size = world()
    if world_size == 1:
        return 1.0
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / world_size
