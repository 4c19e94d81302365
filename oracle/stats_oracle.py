"""CPU oracle of the evaluation-side reductions (TEST INFRASTRUCTURE ONLY).

Restates, with plain numpy masked arrays,
  * ``compute_residuals`` / ``truncate_residuals`` / ``get_statistics`` of the reference (lib/evaluation.py:11-131), and
  * the per-tile part and the robust average of ``compute_local_dsm_std_per_centered_patch`` (lib/utils.py:111-158).
Pinned by tests/golden/stats.npz, produced by the unmodified reference functions (oracle/make_golden_stats.py).
"""
from __future__ import annotations

import numpy as np

STAT_KEYS = ('count_total', 'diff_max', 'diff_min', 'MAE', 'RMSE', 'absolute_median', 'median', 'NMAD')
TRUNC_KEYS = ('count_total', 'MAE', 'RMSE', 'absolute_median', 'median', 'NMAD')


def compute_residuals(raster, raster_gt, nodata, mask_gt=None):
    # lib/evaluation.py:22-37
    if mask_gt is not None:
        mask = np.ma.mask_or(raster_gt == nodata, ~mask_gt)
        gt = np.ma.masked_array(raster_gt, mask=mask)
    else:
        gt = np.ma.masked_where(raster_gt == nodata, raster_gt)
    return np.ma.masked_where(raster == nodata, raster) - gt


def _block(res):
    # lib/evaluation.py:98-119 (the NMAD centre is the median of the ABSOLUTE residuals, as in the reference)
    a = np.ma.abs(res)
    absmed = np.ma.median(a)
    return {'count_total': float(np.ma.count(res)), 'diff_max': float(res.max()), 'diff_min': float(res.min()),
            'MAE': float(np.ma.mean(a)), 'RMSE': float(np.ma.sqrt(np.ma.mean(a ** 2))), 'absolute_median': float(absmed),
            'median': float(np.ma.median(res)), 'NMAD': float(1.4826 * np.ma.median(np.ma.abs(res - absmed)))}


def get_statistics(residuals_masked, residual_threshold=None):
    stats = _block(residuals_masked)
    stats['truncation'] = bool(residual_threshold)
    if residual_threshold:
        t = _block(np.ma.masked_outside(residuals_masked, -residual_threshold, residual_threshold))   # :40-48
        stats['truncated'] = {k: t[k] for k in TRUNC_KEYS}
        stats['truncated']['threshold'] = residual_threshold
    return stats


def tile_stds(dsm, positions, tile, nodata):
    # lib/utils.py:130-151 with batch_size 1: masked mean-centring, sqrt(sum (x - mean)^2 / (count - 1)), wide floats
    out = np.zeros(len(positions), dtype=float)
    for i, (y, x) in enumerate(positions):
        t = dsm[y:y + tile, x:x + tile].astype(np.longdouble)
        t = np.ma.masked_where(t == nodata, t)
        out[i] = np.sqrt(((t - t.mean()) ** 2).sum() / (t.count() - 1))
    return out


def robust_std(stds):
    # lib/utils.py:153-157
    p95, p5 = np.percentile(stds, 95), np.percentile(stds, 5)
    return stds[np.logical_and(stds >= p5, stds <= p95)].mean().item()
