"""CPU: the tile-producer oracle (oracle/tile_oracle.py) against vectors produced by the unmodified reference
``DsmOrthoDataset.__getitem__`` (tests/golden/tiles.npz, generator oracle/make_golden_tiles.py)."""
import os

import numpy as np
import pytest

from oracle import tile_oracle as TO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tiles.npz')


def _cases():
    g = np.load(GOLD)
    out = []
    for ci, (channels, dmean, omean, n) in enumerate(g['cases']):
        for i in range(int(n)):
            out.append((ci, i, str(channels), None if dmean == 'None' else float(dmean),
                        None if omean == 'None' else float(omean)))
    return out


@pytest.mark.parametrize('ci,i,channels,dmean,omean', _cases())
def test_tile_oracle_matches_reference(ci, i, channels, dmean, omean):
    g = np.load(GOLD)
    key = f'c{ci}_s{i}'
    meta = g[key + '_meta']
    y, x, k, vflip, hflip = [int(v) for v in meta[:5]]
    views = [int(v) for v in meta[5:]]
    inp, tgt, mask, mean = TO.make_tile(g['dsm_in'], g['dsm_gt'], g['orthos'], y, x, int(g['tile']), views,
                                        float(g['nodata']), 3.5, 41.0, channels, dmean, omean, k, bool(vflip), bool(hflip))
    np.testing.assert_array_equal(mask, g[key + '_mask'])
    np.testing.assert_allclose(mean, g[key + '_mean'], rtol=1e-7)
    np.testing.assert_allclose(inp, g[key + '_input'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(tgt, g[key + '_target'], rtol=0, atol=2e-3)   # nodata pixels are ~ -3000: fp32 ulp 2.4e-4
    valid = g[key + '_mask']
    np.testing.assert_allclose(tgt[valid], g[key + '_target'][valid], rtol=0, atol=1e-6)
