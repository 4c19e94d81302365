#!/usr/bin/env python
"""Generate tests/golden/stats.npz with the UNMODIFIED reference functions (container only):
``lib.evaluation.compute_residuals`` + ``get_statistics`` on synthetic rasters, and
``lib.utils.compute_local_dsm_std_per_centered_patch`` driven by a list "dataloader" of one-tile batches."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402
from oracle.stats_oracle import STAT_KEYS, TRUNC_KEYS  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'stats.npz')
NODATA = -9999.0
THRESHOLD = 2.5


def rasters(case: int):
    rng = np.random.default_rng(100 + case)
    rows, cols = [(37, 53), (64, 64), (90, 41)][case]
    gt = (400 + 6 * np.sin(np.arange(rows)[:, None] / 5.0) + rng.standard_normal((rows, cols))).astype(np.float32)
    pred = gt.astype(np.float64) + rng.standard_normal((rows, cols)) * [1.0, 2.0, 0.5][case] + 0.2
    pred[rng.random((rows, cols)) < 0.01] += 9.0                       # outliers beyond the truncation threshold
    gt[rng.random((rows, cols)) < 0.04] = NODATA
    pred[rng.random((rows, cols)) < 0.02] = NODATA
    mask_gt = rng.random((rows, cols)) > 0.1 if case != 1 else None
    if case == 2:
        pred = pred.astype(np.float32)                                 # float32 - float32 residuals (initial DSM case)
    return pred, gt, mask_gt


def main():
    import importlib
    ref_shims.import_reference()
    ev = importlib.import_module('lib.evaluation')
    ut = importlib.import_module('lib.utils')
    out = {}
    for case in range(3):
        pred, gt, mask_gt = rasters(case)
        res = ev.compute_residuals(pred, gt, NODATA, mask_gt)
        st = ev.get_statistics(res, THRESHOLD)
        out[f'c{case}_pred'] = pred
        out[f'c{case}_gt'] = gt
        if mask_gt is not None:
            out[f'c{case}_mask'] = mask_gt
        out[f'c{case}_res'] = np.ma.filled(res.astype(np.float64), 0.0)
        out[f'c{case}_valid'] = ~np.ma.getmaskarray(res)
        out[f'c{case}_stats'] = np.array([float(st[k]) for k in STAT_KEYS])
        out[f'c{case}_trunc'] = np.array([float(st.truncated[k]) for k in TRUNC_KEYS])
        st0 = ev.get_statistics(res, None)
        assert st0.truncation is False
    # sigma_DSM estimation: a list of batches is a valid "dataloader" (len + iteration)
    rng = np.random.default_rng(7)
    dsm = (500 + 10 * np.sin(np.arange(128)[:, None] / 11.0) + 4 * rng.standard_normal((128, 160))).astype(np.float32)
    dsm[rng.random(dsm.shape) < 0.03] = NODATA
    T = 32
    pos = [(int(rng.integers(0, 128 - T + 1)), int(rng.integers(0, 160 - T + 1))) for _ in range(40)]
    batches = []
    for (y, x) in pos:
        tile = dsm[y:y + T, x:x + T]
        batches.append({'input': torch.from_numpy(tile.copy())[None, None], 'nodata': torch.tensor([NODATA])})
    std = ut.compute_local_dsm_std_per_centered_patch(batches, 'raster_in')
    out.update(std_dsm=dsm, std_pos=np.array(pos, dtype=np.int32), std_tile=np.array(T), std_value=np.array(std))
    out.update(nodata=np.array(NODATA), threshold=np.array(THRESHOLD))
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: out[k].shape for k in out if k.endswith('stats')}, 'robust std', std)


if __name__ == '__main__':
    main()
