#!/usr/bin/env python
"""Generate tests/golden/tiles.npz by calling the UNMODIFIED reference ``DsmOrthoDataset.__getitem__``
(container only).  The dataset object is created without running ``__init__`` (which needs GDAL rasters on disk);
its attributes are set to synthetic in-memory rasters, exactly the state ``_load_data`` / ``_determine_patches``
would leave (lib/DsmOrthoDataset.py:293-431).  ``random`` / ``np.random`` are seeded per sample so that the
oracle can be given the same augmentation decisions."""
from __future__ import annotations

import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'tiles.npz')


def synthetic_rasters(rows=96, cols=120, n_views=4, seed=3):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    dsm_gt = (400 + 8 * np.sin(yy / 9.0) + 5 * np.cos(xx / 7.0) + rng.standard_normal((rows, cols))).astype(np.float32)
    dsm_in = (dsm_gt + 1.5 * rng.standard_normal((rows, cols))).astype(np.float32)
    nodata = np.float32(-9999.0)
    dsm_gt[rng.random((rows, cols)) < 0.03] = nodata
    dsm_in[rng.random((rows, cols)) < 0.02] = nodata
    dsm_gt[5:8, 10:14] = 0.0                       # exact zeros are masked out by the reference's mask1
    orthos = (120 + 40 * rng.standard_normal((rows, cols, n_views))).astype(np.float32)
    return dsm_in, dsm_gt, orthos, nodata


def main():
    import importlib
    ref_shims.import_reference()
    ds_mod = importlib.import_module('lib.DsmOrthoDataset')
    dsm_in, dsm_gt, orthos, nodata = synthetic_rasters()
    T = 32
    rng = np.random.default_rng(11)
    out = {'dsm_in': dsm_in, 'dsm_gt': dsm_gt, 'orthos': orthos, 'nodata': nodata, 'tile': np.int64(T)}
    cases = []
    for ci, (channels, pairs, permute, dmean, omean) in enumerate([
            ('geom-stereo', [[0, 1], [2, 3], [1, 3]], False, None, None),
            ('geom-stereo', [[0, 1], [2, 3]], True, None, 118.5),
            ('geom-mono', [[0], [2]], False, 401.25, None),
            ('geom', None, False, None, None),
            ('stereo', [[3, 0]], False, None, None)]):
        ds = object.__new__(ds_mod.DsmOrthoDataset)
        ds.input_channels, ds.tile_size, ds.sampling_strategy = channels, T, 'train'
        ds.augment, ds.transform_dsm, ds.transform_orthos = True, True, True
        ds.dsm_mean, ds.dsm_std, ds.ortho_mean, ds.ortho_std = dmean, 3.5, omean, 41.0
        ds.permute_images_within_pair = permute
        ds.raster_gt = 'in-memory'
        ds.dsm_input, ds.dsm_target, ds.nodata = dsm_in, dsm_gt, np.array(nodata)
        n = 12
        pos = [(int(rng.integers(0, dsm_in.shape[0] - T + 1)), int(rng.integers(0, dsm_in.shape[1] - T + 1))) for _ in range(n)]
        ds.patch_position = pos
        if pairs is not None:
            ds.orthos, ds.image_pairs = orthos, pairs
            ds.image_pair_indices = rng.integers(0, len(pairs), n)
        for i in range(n):
            seed = 1000 * ci + i
            random.seed(seed)
            np.random.seed(seed)
            item = ds[i]
            # replay the decisions the reference drew (np.random: permutation; random: k, vflip, hflip)
            np.random.seed(seed)
            views = list(pairs[ds.image_pair_indices[i]]) if pairs is not None else []
            if permute:
                perm = np.arange(len(views))
                np.random.shuffle(perm)
                views = [views[p] for p in perm]
            random.seed(seed)
            k = random.randint(0, 3)
            vflip = random.random() < 0.5
            hflip = random.random() < 0.5
            key = f'c{ci}_s{i}'
            out[key + '_input'] = item['input'].numpy()
            out[key + '_target'] = item['target'].numpy()
            out[key + '_mask'] = item['loss_mask'].numpy()
            out[key + '_mean'] = np.float32(item['dsm_mean'])
            out[key + '_meta'] = np.array([pos[i][0], pos[i][1], k, int(vflip), int(hflip)] + views, dtype=np.int64)
        cases.append([channels, str(dmean), str(omean), str(n)])
    out['cases'] = np.array(cases)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
