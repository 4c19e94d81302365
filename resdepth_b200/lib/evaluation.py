"""Tiled inference with linear blending: mirror of ``predict_linear_blend`` (reference lib/evaluation.py:460-567).

Same signature and return value (a float64 ``numpy`` raster with the extent of the input DSM).  The reference
runs one forward per DataLoader batch, copies the prediction to the host, de-normalises it in numpy and
accumulates ``tile * weights`` tile by tile in Python; here the forward (``rd_forward``), the de-normalisation
(``denormalize_numpy``, lib/data_normalization.py:41-53), the per-tile ramp weights (``_get_blend_weights``,
lib/evaluation.py:516-567) and the accumulation all stay on the device (``rd_blend_accumulate``), the input tiles of
batch i+1 are copied while batch i computes, and the raster crosses PCIe once at the end.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from .distributed import owns_batch, sum_partial_rasters, world
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


_PINNED = {}                 # (device index) -> two pinned float64 staging buffers, kept for the life of the process


def _raster_to_host(raster: torch.Tensor, chunk_bytes: int = 8 << 20) -> np.ndarray:
    """Device float64 raster -> fresh numpy array, through two small PINNED staging buffers: chunk i+1 crosses PCIe
    while the host copies chunk i into the result.  Measured on the B200 box (4096 x 4096): ``raster.cpu()`` into
    pageable memory ran at 2.1 GB/s (63 ms, more than the 53 ms of all 31 forward passes of that raster), a pinned
    destination at 56 GB/s -- but pinning 134 MB costs 50-70 ms itself, hence the small reusable buffers (first call 83 ms
    with two allocations, later calls 32 ms)."""
    rows, cols = raster.shape
    out = np.empty((rows, cols), dtype=np.float64)
    rpc = max(1, min(rows, chunk_bytes // max(cols * 8, 1)))
    key = (raster.device.index, rpc * cols)
    bufs = _PINNED.get(key)
    if bufs is None:
        _PINNED.clear()                                   # one raster width at a time: do not accumulate pinned memory
        both = torch.empty(2 * rpc * cols, dtype=torch.float64, pin_memory=True)      # ONE allocation: pinning has a
        bufs = _PINNED[key] = [both[:rpc * cols], both[rpc * cols:]]                  # fixed cost of ~25 ms on this box
    events = [torch.cuda.Event(), torch.cuda.Event()]
    flat = raster.reshape(-1)
    pending = None                                        # (slot, first row, number of rows) whose D2H is in flight
    for i, r0 in enumerate(range(0, rows, rpc)):
        n = min(rpc, rows - r0)
        slot = i & 1
        bufs[slot][:n * cols].copy_(flat[r0 * cols:(r0 + n) * cols], non_blocking=True)
        events[slot].record()
        if pending is not None:
            ps, pr, pn = pending
            events[ps].synchronize()
            out[pr:pr + pn] = bufs[ps][:pn * cols].numpy().reshape(pn, cols)
        pending = (slot, r0, n)
    if pending is not None:
        ps, pr, pn = pending
        events[ps].synchronize()
        out[pr:pr + pn] = bufs[ps][:pn * cols].numpy().reshape(pn, cols)
    return out


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

    rank, world_size = world()          # one process per GPU: every rank blends its share of the batches
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)

    def stage(batch):
        """H2D copies of one batch on the copy stream (one batch ahead of the forward pass that consumes it)."""
        x = batch['input']
        n = x.shape[0]
        geom = np.stack([_as_int_array(batch[k]).reshape(n) for k in (
            'patch_offset_y', 'patch_offset_x', 'patch_valid_pixels_uly', 'patch_valid_pixels_ulx',
            'patch_valid_pixels_lry', 'patch_valid_pixels_lrx')], axis=1).astype(np.int32)
        if x.is_cuda:
            copy.wait_stream(main)
        with torch.cuda.stream(copy):
            if not x.is_cuda and not x.is_pinned():
                x = x.pin_memory()
            xd = x.to(device, dtype=torch.float32, non_blocking=True)
            geom_d = torch.from_numpy(np.ascontiguousarray(geom)).to(device, non_blocking=True)
            mean = torch.flatten(batch['dsm_mean']).to(device, dtype=torch.float32, non_blocking=True)
            std = torch.flatten(batch['dsm_std']).to(device, dtype=torch.float32, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        for t in (xd, geom_d, mean, std):
            t.record_stream(main)
        return xd, geom_d, mean, std, ready

    mine = (batch for bi, batch in enumerate(dataloader) if owns_batch(bi, rank, world_size))
    with torch.no_grad(), model.constant_weights(device):     # nothing in the loop touches the weights: pack once
        nxt = next(mine, None)
        nxt = stage(nxt) if nxt is not None else None
        while nxt is not None:
            xd, geom_d, mean, std, ready = nxt
            following = next(mine, None)
            nxt = stage(following) if following is not None else None     # in flight while this batch computes
            main.wait_event(ready)
            y_pred = model(xd)
            blend_tiles_into(raster, y_pred, mean, std, geom_d, tile_size, stride)
    return _raster_to_host(sum_partial_rasters(raster))


# -------------------------------------------------------------------------------------------------
# Residual statistics on the device (reference lib/evaluation.py:11-131)
# -------------------------------------------------------------------------------------------------
class _Stats(dict):
    """Attribute dictionary with the keys of the reference's EasyDict result."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class DeviceResiduals:
    """Residual errors ``raster - raster_gt`` held on the device: ``data`` float64, ``valid`` bool (True = unmasked).
    Stands in for the ``np.ma`` array the reference's ``compute_residuals`` returns; ``to_masked_array()`` gives
    exactly that array, ``compressed()`` its valid values, slicing returns a view-like ``DeviceResiduals``."""

    def __init__(self, data: torch.Tensor, valid: torch.Tensor):
        self.data, self.valid = data, valid

    @property
    def shape(self):
        return tuple(self.data.shape)

    def __getitem__(self, idx):
        return DeviceResiduals(self.data[idx], self.valid[idx])

    def to_masked_array(self):
        return np.ma.masked_array(self.data.cpu().numpy(), mask=~self.valid.cpu().numpy())

    def compressed(self):
        return self.data[self.valid].cpu().numpy()


def _device_array(a, device, allow=(torch.float32, torch.float64)):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in allow:
        t = t.to(torch.float64)
    return t.to(device).contiguous()


def compute_residuals(raster, raster_gt, nodata, mask_gt=None) -> DeviceResiduals:
    """Same arguments and masking rules as the reference (numpy arrays or CUDA tensors); the result stays on the
    device.  A positive error means the predicted height is larger than the reference value."""
    if not torch.cuda.is_available():
        raise RuntimeError('resdepth_b200: compute_residuals needs a CUDA device (no CPU fallback)')
    device = raster.device if isinstance(raster, torch.Tensor) and raster.is_cuda else torch.device('cuda', torch.cuda.current_device())
    r = _device_array(raster, device)
    g = _device_array(raster_gt, device)
    if r.shape != g.shape:
        raise ValueError('raster and raster_gt must have the same shape')
    m = None
    if mask_gt is not None:
        m = mask_gt if isinstance(mask_gt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mask_gt))
        m = (m != 0).to(device).contiguous().view(torch.uint8)
    res = torch.empty(r.shape, dtype=torch.float64, device=device)
    valid = torch.empty(r.shape, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _native.residuals(r.data_ptr(), int(r.dtype == torch.float64), g.data_ptr(), int(g.dtype == torch.float64),
                          m.data_ptr() if m is not None else None, r.numel(), float(nodata), res.data_ptr(),
                          valid.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return DeviceResiduals(res, valid.view(torch.bool))


def get_statistics(residuals_masked, residual_threshold=None):
    """Evaluation metrics of the reference's ``get_statistics`` (same keys, same definitions -- including NMAD around
    the median of the absolute residuals) computed by ``rd_residual_stats``.  Accepts a ``DeviceResiduals`` or a numpy
    masked array."""
    if not torch.cuda.is_available():
        raise RuntimeError('resdepth_b200: get_statistics needs a CUDA device (no CPU fallback)')
    if isinstance(residuals_masked, DeviceResiduals):
        data, valid = residuals_masked.data.contiguous(), residuals_masked.valid.contiguous()
    else:
        device = torch.device('cuda', torch.cuda.current_device())
        ma = np.ma.asarray(residuals_masked)
        data = torch.from_numpy(np.ascontiguousarray(np.ma.filled(ma.astype(np.float64), 0.0))).to(device)
        valid = torch.from_numpy(np.ascontiguousarray(~np.ma.getmaskarray(ma))).to(device)
    if data.dtype != torch.float64:
        data = data.double()
    with torch.cuda.device(data.device):
        o = _native.residual_stats(data.data_ptr(), valid.view(torch.uint8).data_ptr(), data.numel(),
                                   float(residual_threshold) if residual_threshold else 0.0,
                                   torch.cuda.current_stream().cuda_stream)
    stats = _Stats(truncation=True if residual_threshold else False, count_total=o[0], diff_max=o[1], diff_min=o[2],
                   MAE=o[3], RMSE=o[4], absolute_median=o[5], median=o[6], NMAD=o[7])
    if residual_threshold:
        stats['truncated'] = _Stats(count_total=o[8], threshold=residual_threshold, MAE=o[9], RMSE=o[10],
                                    absolute_median=o[11], median=o[12], NMAD=o[13])
    return stats
