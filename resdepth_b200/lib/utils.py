"""Device-side counterpart of one pre-training pass of the reference's ``lib/utils.py``."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native


def compute_local_dsm_std_per_centered_patch(producer, raster_identifier='raster_in', positions=None):
    """Single robust scale factor across the DSM training tiles (reference lib/utils.py:111-158).

    The reference walks a batch-size-1 DataLoader over all training tiles and does the arithmetic in float128 on the
    host; here the tiles are crops of the raster that ``producer`` (a ``DeviceTileProducer``) already holds in HBM and
    the per-tile standard deviations come from one kernel launch (``rd_tile_stds``, float64, two passes).  The
    5th / 95th percentile trimming and the final average are the reference's numpy lines on the ``n`` results.

    :param producer:           resdepth_b200.lib.tiles.DeviceTileProducer
    :param raster_identifier:  'raster_in' (initial DSM) or anything else for the ground-truth DSM, as in the reference
    :param positions:          sequence of (y, x) tile origins (the training patches); required
    :return:                   float, standard deviation of the zero-centred DSM training tiles
    """
    if positions is None or len(positions) == 0:
        raise ValueError('positions (tile origins of the training patches) are required')
    dsm = producer.dsm_input if raster_identifier == 'raster_in' else producer.dsm_target
    T = producer.tile_size
    pa = np.asarray(positions, dtype=np.int32).reshape(-1, 2)
    if np.any(pa < 0) or np.any(pa[:, 0] + T > producer.rows) or np.any(pa[:, 1] + T > producer.cols):
        raise ValueError('tile position outside the raster')
    pos = torch.from_numpy(np.ascontiguousarray(pa)).to(dsm.device)
    stds = torch.empty(len(pa), dtype=torch.float64, device=dsm.device)
    with torch.cuda.device(dsm.device):
        _native.tile_stds(dsm.data_ptr(), producer.rows, producer.cols, pos.data_ptr(), len(pa), T, float(producer.nodata),
                          stds.data_ptr(), torch.cuda.current_stream().cuda_stream)
    stds = stds.cpu().numpy()
    perc95 = np.percentile(stds, 95)
    perc5 = np.percentile(stds, 5)
    return stds[np.logical_and(stds >= perc5, stds <= perc95)].mean().item()
