"""Tiled inference with linear blending: mirror of ``predict_linear_blend`` (reference lib/evaluation.py:460-567).

Same signature and return value (a float64 ``numpy`` raster with the extent of the input DSM).  The reference
runs one forward per DataLoader batch, copies the prediction to the host, de-normalises it in numpy and
accumulates ``tile * weights`` tile by tile in Python; here the forward (``rd_forward``), the de-normalisation
(``denormalize_numpy``, lib/data_normalization.py:41-53), the per-tile ramp weights (``_get_blend_weights``,
lib/evaluation.py:516-567) and the accumulation all stay on the device (``rd_blend_accumulate``), and the raster
crosses PCIe once at the end.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from .UNet import UNet


def _as_int_array(v):
    return v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def blend_tiles_into(raster: torch.Tensor, tiles: torch.Tensor, mean: torch.Tensor, std: torch.Tensor,
                     geom: torch.Tensor, tile_size: int, stride: int):
    """raster [rows, cols] float64 (device) += blend of ``tiles`` [n,1,T,T]; geom int32 [n,6] = (y, x, uly, ulx,
    lry, lrx) per tile (offsets in the raster + the non-overlapping box of lib/rasterutils.py:100-191)."""
    if raster.dtype != torch.float64 or not raster.is_cuda or not raster.is_contiguous():
        raise ValueError('resdepth_b200: raster must be a contiguous float64 CUDA tensor')
    rows, cols = raster.shape
    tiles = tiles.contiguous()
    with torch.cuda.device(raster.device):
        _native.blend_accumulate(tiles.data_ptr(), mean.data_ptr(), std.data_ptr(), geom.data_ptr(), tiles.shape[0],
                                 tile_size, stride, raster.data_ptr(), rows, cols,
                                 torch.cuda.current_stream().cuda_stream)


def predict_linear_blend(dataloader, model):
    if not torch.cuda.is_available():
        raise RuntimeError('resdepth_b200: predict_linear_blend needs a CUDA device (no CPU fallback)')
    if not isinstance(model, UNet):
        raise TypeError('resdepth_b200: predict_linear_blend drives resdepth_b200.lib.UNet.UNet models only')
    device = torch.device('cuda', torch.cuda.current_device())
    model.eval()
    model.to(device)

    dataset = dataloader.dataset
    cols = dataset.dsm_input_gdal.RasterXSize
    rows = dataset.dsm_input_gdal.RasterYSize
    tile_size, stride = dataset.tile_size, dataset.stride
    raster = torch.zeros((rows, cols), dtype=torch.float64, device=device)

    with torch.no_grad():
        for batch in dataloader:
            x = batch['input']
            if not x.is_cuda and not x.is_pinned():
                x = x.pin_memory()
            x = x.to(device, dtype=torch.float32, non_blocking=True)
            n = x.shape[0]
            geom = np.stack([_as_int_array(batch[k]).reshape(n) for k in (
                'patch_offset_y', 'patch_offset_x', 'patch_valid_pixels_uly', 'patch_valid_pixels_ulx',
                'patch_valid_pixels_lry', 'patch_valid_pixels_lrx')], axis=1).astype(np.int32)
            geom_d = torch.from_numpy(np.ascontiguousarray(geom)).to(device, non_blocking=True)
            mean = torch.flatten(batch['dsm_mean']).to(device, dtype=torch.float32, non_blocking=True)
            std = torch.flatten(batch['dsm_std']).to(device, dtype=torch.float32, non_blocking=True)
            y_pred = model(x)
            blend_tiles_into(raster, y_pred, mean, std, geom_d, tile_size, stride)
    return raster.cpu().numpy()
