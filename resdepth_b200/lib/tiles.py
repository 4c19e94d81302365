"""On-device training-tile producer: the hot part of the reference's ``DsmOrthoDataset.__getitem__``
(lib/DsmOrthoDataset.py:161-291, sampling_strategy='train') for whole batches.

The reference keeps the rasters fully in host RAM (lib/DsmOrthoDataset.py:293-314) and builds every tile with
numpy slicing plus per-channel ``np.rot90`` / ``np.flip`` loops in DataLoader worker processes; at thousands of
tiles per second that producer, not the network, becomes the bottleneck.  ``DeviceTileProducer`` uploads the
rasters once and produces a batch dict with the same keys and values (``input``, ``target``, ``loss_mask``,
``dsm_mean``, ``dsm_std``, ``patch_offset_x/y``) directly in HBM with one CUDA call (``rd_make_tiles``).
Random decisions are drawn on the host like the reference's sampler does (tile positions, image pair, optional
permutation of the pair, ``k`` quarter turns, vertical / horizontal flip) or passed in explicitly.
"""
from __future__ import annotations

import math
import random
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _native

INPUT_CHANNELS = ('geom', 'geom-mono', 'geom-multiview', 'geom-stereo', 'stereo')


class DeviceTileProducer:
    def __init__(self, dsm_input, dsm_target, orthos, nodata: float, tile_size: int, input_channels: str = 'geom-stereo',
                 image_pairs: Optional[Sequence[Sequence[int]]] = None, dsm_mean: Optional[float] = None,
                 dsm_std: float = 1.0, ortho_mean: Optional[float] = None, ortho_std: float = 1.0, augment: bool = True,
                 permute_images_within_pair: bool = False, device='cuda'):
        if input_channels not in INPUT_CHANNELS:
            raise ValueError(f"Unknown input channel configuration: '{input_channels}'. Choose among {INPUT_CHANNELS}.")
        if not torch.cuda.is_available():
            raise RuntimeError('resdepth_b200: DeviceTileProducer needs a CUDA device (no CPU fallback)')
        self.device = torch.device(device)
        self.input_channels, self.tile_size = input_channels, int(tile_size)
        self.nodata, self.dsm_mean, self.dsm_std = float(nodata), dsm_mean, float(dsm_std)
        self.ortho_mean, self.ortho_std = ortho_mean, float(ortho_std)
        self.augment, self.permute = augment, permute_images_within_pair
        self.dsm_input = torch.as_tensor(np.ascontiguousarray(dsm_input, dtype=np.float32)).to(self.device)
        self.dsm_target = torch.as_tensor(np.ascontiguousarray(dsm_target, dtype=np.float32)).to(self.device)
        self.rows, self.cols = self.dsm_input.shape
        self.include_dsm = input_channels != 'stereo'
        if input_channels != 'geom':
            if orthos is None or not image_pairs:
                raise ValueError("ortho images and image_pairs are required unless input_channels == 'geom'")
            # reference layout is [rows, cols, views] (np.dstack); the kernels read planar [views, rows, cols]
            self.orthos = torch.as_tensor(np.ascontiguousarray(orthos, dtype=np.float32)).to(self.device) \
                .permute(2, 0, 1).contiguous()
            self.image_pairs = [list(p) for p in image_pairs]
            self.n_ortho = len(self.image_pairs[0])
            if any(len(p) != self.n_ortho for p in self.image_pairs):
                raise ValueError('all image pairs must have the same number of views')
        else:
            self.orthos, self.image_pairs, self.n_ortho = None, None, 0
        self.n_channels = self.n_ortho + (1 if self.include_dsm else 0)

    @classmethod
    def from_device(cls, dsm_input, dsm_target, orthos_planar, nodata, tile_size, input_channels='geom-stereo',
                    image_pairs=None, dsm_mean=None, dsm_std=1.0, ortho_mean=None, ortho_std=1.0, augment=True,
                    permute_images_within_pair=False):
        """Adopt rasters that already live on the device (orthos planar: [views, rows, cols])."""
        self = object.__new__(cls)
        self.device = dsm_input.device
        self.input_channels, self.tile_size = input_channels, int(tile_size)
        self.nodata, self.dsm_mean, self.dsm_std = float(nodata), dsm_mean, float(dsm_std)
        self.ortho_mean, self.ortho_std = ortho_mean, float(ortho_std)
        self.augment, self.permute = augment, permute_images_within_pair
        self.dsm_input, self.dsm_target = dsm_input.float().contiguous(), dsm_target.float().contiguous()
        self.rows, self.cols = self.dsm_input.shape
        self.include_dsm = input_channels != 'stereo'
        self.orthos = orthos_planar.float().contiguous() if input_channels != 'geom' else None
        self.image_pairs = [list(p) for p in image_pairs] if self.orthos is not None else None
        self.n_ortho = len(self.image_pairs[0]) if self.orthos is not None else 0
        self.n_channels = self.n_ortho + (1 if self.include_dsm else 0)
        return self

    def draw(self, n: int):
        """Random decisions of one batch, drawn like the reference: uniformly sampled tile origins, one image pair per
        tile, optional permutation inside the pair, Rotate() / RandomVerticalFlip() / RandomHorizontalFlip()."""
        T = self.tile_size
        pos = [(random.randint(0, self.rows - T), random.randint(0, self.cols - T)) for _ in range(n)]
        views = []
        for _ in range(n):
            v = list(self.image_pairs[random.randrange(len(self.image_pairs))]) if self.n_ortho else []
            if self.permute:
                random.shuffle(v)
            views.append(v)
        if self.augment:
            aug = [(random.randint(0, 3), int(random.random() < 0.5), int(random.random() < 0.5)) for _ in range(n)]
        else:
            aug = [(0, 0, 0)] * n
        return pos, views, aug

    def make_batch(self, positions, views=None, aug=None) -> dict:
        """positions: n x (y, x); views: n x n_ortho view indices (already permuted); aug: n x (k, vflip, hflip)."""
        n, T, dev = len(positions), self.tile_size, self.device
        pos_t = torch.tensor(np.asarray(positions, dtype=np.int32).reshape(n, 2), device=dev)
        pa = np.asarray(positions).reshape(n, 2)
        if np.any(pa < 0) or np.any(pa[:, 0] + T > self.rows) or np.any(pa[:, 1] + T > self.cols):
            raise ValueError('tile position outside the raster')
        aug_t = torch.tensor(np.asarray(aug if aug is not None else [(0, 0, 0)] * n, dtype=np.int32).reshape(n, 3), device=dev)
        if self.n_ortho:
            va = np.asarray(views, dtype=np.int32).reshape(n, self.n_ortho)
            if np.any(va < 0) or np.any(va >= self.orthos.shape[0]):
                raise ValueError('ortho view index outside the raster stack')
            views_t = torch.tensor(va, device=dev)
        else:
            views_t = None
        inp = torch.empty((n, self.n_channels, T, T), device=dev)
        tgt = torch.empty((n, 1, T, T), device=dev)
        mask = torch.empty((n, 1, T, T), device=dev, dtype=torch.uint8)
        mean = torch.empty(n, device=dev)
        scratch = torch.empty(2 * n, device=dev)
        with torch.cuda.device(dev):
            _native.make_tiles(self.dsm_input.data_ptr(), self.dsm_target.data_ptr(),
                               self.orthos.data_ptr() if self.orthos is not None else None, self.rows, self.cols,
                               self.orthos.shape[0] if self.orthos is not None else 0, pos_t.data_ptr(),
                               views_t.data_ptr() if views_t is not None else None, aug_t.data_ptr(), n, T, self.n_ortho,
                               int(self.include_dsm), self.nodata, self.dsm_std, self.ortho_std,
                               math.nan if self.dsm_mean is None else float(self.dsm_mean),
                               math.nan if self.ortho_mean is None else float(self.ortho_mean), inp.data_ptr(),
                               tgt.data_ptr(), mask.data_ptr(), mean.data_ptr(), scratch.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
        ys = torch.tensor([p[0] for p in positions])
        xs = torch.tensor([p[1] for p in positions])
        return {'input': inp, 'target': tgt, 'loss_mask': mask.view(torch.bool), 'dsm_mean': mean,
                'dsm_std': torch.full((n,), self.dsm_std, device=dev), 'patch_offset_x': xs, 'patch_offset_y': ys,
                'nodata': torch.full((n,), self.nodata)}

    def sample_batch(self, n: int) -> dict:
        pos, views, aug = self.draw(n)
        return self.make_batch(pos, views, aug)
