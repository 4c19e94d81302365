"""CPU oracle of the training-tile producer (TEST INFRASTRUCTURE ONLY).

Restates ``DsmOrthoDataset.__getitem__`` of the reference (lib/DsmOrthoDataset.py:161-291) for the training
sampling strategy -- crop at (y, x), per-tile masked mean-centring and division by sigma
(lib/DsmOrthoDataset.py:191-210, lib/data_normalization.py:6-26), ortho-image gather / optional permutation /
normalisation (:213-255), loss mask (:433-470), rot90 / flipud / fliplr augmentation
(lib/torch_transforms.py:15-157) -- with the random decisions passed in explicitly (the reference draws them from
``random`` / ``np.random``).  Pinned by tests/golden/tiles.npz, produced by the unmodified reference method
(oracle/make_golden_tiles.py).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def loss_mask(dsm_target: np.ndarray, nodata: float) -> np.ndarray:
    # _get_dsm_loss_mask with patch_valid_pixels=None (train): lib/DsmOrthoDataset.py:433-470
    return np.logical_and(dsm_target != 0, dsm_target != np.float32(nodata))


def augment(stack: np.ndarray, k: int, vflip: bool, hflip: bool) -> np.ndarray:
    """Rotate(k) -> RandomVerticalFlip -> RandomHorizontalFlip on a [C, T, T] stack (lib/torch_transforms.py)."""
    out = np.stack([np.rot90(c, k) for c in stack])
    if vflip:
        out = np.stack([np.flipud(c) for c in out])
    if hflip:
        out = np.stack([np.fliplr(c) for c in out])
    return np.ascontiguousarray(out)


def make_tile(dsm_input: np.ndarray, dsm_target: np.ndarray, orthos: Optional[np.ndarray], y: int, x: int, tile: int,
              views: Sequence[int], nodata: float, dsm_std: float, ortho_std: float, input_channels: str = 'geom-stereo',
              dsm_mean: Optional[float] = None, ortho_mean: Optional[float] = None, k: int = 0, vflip: bool = False,
              hflip: bool = False, do_augment: bool = True):
    """Returns (input [C,T,T] f32, target [1,T,T] f32, loss_mask [1,T,T] bool, dsm_mean f32)."""
    nod = np.float32(nodata)
    d_in = dsm_input[y:y + tile, x:x + tile]
    d_gt = dsm_target[y:y + tile, x:x + tile]
    mask = loss_mask(d_gt, nodata)[None]
    if dsm_mean is None:                                           # lib/DsmOrthoDataset.py:193-195
        mean = np.ma.mean(np.ma.masked_where(d_in == nod, d_in))
    else:
        mean = dsm_mean
    mean32, std32 = np.float32(mean), np.float32(dsm_std)
    d_in_n = ((d_in - mean32) / std32)[None].astype(np.float32)   # ToTensor + Normalize: (x - mean) / std in fp32
    d_gt_n = ((d_gt - mean32) / std32)[None].astype(np.float32)
    if input_channels != 'geom':
        o = orthos[y:y + tile, x:x + tile, list(views)].transpose((2, 0, 1)).copy()
        omean = np.float32(o.mean() if ortho_mean is None else ortho_mean)
        o = ((o - omean) / np.float32(ortho_std)).astype(np.float32)
        inputs = o if input_channels == 'stereo' else np.concatenate([d_in_n, o], axis=0)
    else:
        inputs = d_in_n
    if do_augment:
        stack = augment(np.concatenate([mask.astype(np.float32), d_gt_n, inputs], axis=0), k, vflip, hflip)
        mask, d_gt_n, inputs = stack[0:1] != 0, stack[1:2], stack[2:]
    return inputs, d_gt_n, mask, mean32
