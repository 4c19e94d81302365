"""Evaluation-side reductions: residual statistics and the sigma_DSM estimation pass.
CPU: the oracle against goldens produced by the unmodified reference functions.  GPU: the CUDA path against both."""
import os

import numpy as np
import pytest
import torch

from oracle import stats_oracle as SO

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'stats.npz'))
NODATA, THR = float(G['nodata']), float(G['threshold'])


def _case(c):
    mask = G[f'c{c}_mask'] if f'c{c}_mask' in G.files else None
    return G[f'c{c}_pred'], G[f'c{c}_gt'], mask


@pytest.mark.parametrize('c', [0, 1, 2])
def test_oracle_matches_reference_goldens(c):
    pred, gt, mask = _case(c)
    res = SO.compute_residuals(pred, gt, NODATA, mask)
    np.testing.assert_array_equal(~np.ma.getmaskarray(res), G[f'c{c}_valid'])
    np.testing.assert_allclose(np.ma.filled(res.astype(np.float64), 0.0), G[f'c{c}_res'], rtol=0, atol=0)
    st = SO.get_statistics(res, THR)
    np.testing.assert_allclose([st[k] for k in SO.STAT_KEYS], G[f'c{c}_stats'], rtol=1e-12)
    np.testing.assert_allclose([st['truncated'][k] for k in SO.TRUNC_KEYS], G[f'c{c}_trunc'], rtol=1e-12)
    assert SO.get_statistics(res, None)['truncation'] is False


def test_oracle_tile_std_matches_reference_golden():
    stds = SO.tile_stds(G['std_dsm'], [tuple(p) for p in G['std_pos']], int(G['std_tile']), NODATA)
    assert abs(SO.robust_std(stds) - float(G['std_value'])) <= 1e-9 * float(G['std_value'])


@pytest.mark.gpu
@pytest.mark.parametrize('c', [0, 1, 2])
def test_cuda_residual_statistics(c):
    from resdepth_b200.lib.evaluation import compute_residuals, get_statistics
    pred, gt, mask = _case(c)
    res = compute_residuals(pred, gt, NODATA, mask)
    np.testing.assert_array_equal(res.valid.cpu().numpy(), G[f'c{c}_valid'])
    np.testing.assert_array_equal(res.data.cpu().numpy(), G[f'c{c}_res'])                   # element-wise: bit-exact
    ma = res.to_masked_array()
    assert ma.count() == int(G[f'c{c}_stats'][0])
    st = get_statistics(res, THR)
    # float32 - float32 residuals (case 2): the reference accumulates in float32, the kernel in float64
    rtol = 1e-11 if c != 2 else 2e-6
    got = np.array([st[k] for k in SO.STAT_KEYS])
    np.testing.assert_allclose(got, G[f'c{c}_stats'], rtol=rtol)
    got_t = np.array([st.truncated[k] for k in SO.TRUNC_KEYS])
    np.testing.assert_allclose(got_t, G[f'c{c}_trunc'], rtol=rtol)
    # order statistics are exact selections: identical to the oracle's float64 sort
    o = SO.get_statistics(SO.compute_residuals(pred, gt, NODATA, mask).astype(np.float64), THR)
    for k in ('absolute_median', 'median', 'NMAD', 'diff_max', 'diff_min', 'count_total'):
        assert st[k] == o[k], k
    for k in ('absolute_median', 'median', 'NMAD', 'count_total'):
        assert st.truncated[k] == o['truncated'][k], k
    # numpy masked array in, no threshold
    st2 = get_statistics(ma, None)
    assert st2.truncation is False and 'truncated' not in st2 and st2.median == st.median


@pytest.mark.gpu
def test_cuda_residual_statistics_large_and_degenerate():
    from resdepth_b200.lib.evaluation import DeviceResiduals, get_statistics
    g = torch.Generator().manual_seed(5)
    n = 3_000_001                                           # odd count, more elements than one pass of the grid
    r = torch.randn(n, generator=g, dtype=torch.float64) * 3 + 0.25
    valid = torch.rand(n, generator=g) > 0.2
    st = get_statistics(DeviceResiduals(r.cuda(), valid.cuda()), 4.0)
    ma = np.ma.masked_array(r.numpy(), mask=~valid.numpy())
    o = SO.get_statistics(ma, 4.0)
    for k in ('count_total', 'diff_max', 'diff_min', 'absolute_median', 'median', 'NMAD'):
        assert st[k] == o[k], k
    for k in ('count_total', 'absolute_median', 'median', 'NMAD'):
        assert st.truncated[k] == o['truncated'][k], k
    np.testing.assert_allclose([st.MAE, st.RMSE, st.truncated.MAE, st.truncated.RMSE],
                               [o['MAE'], o['RMSE'], o['truncated']['MAE'], o['truncated']['RMSE']], rtol=1e-11)
    # nothing valid: count 0, the rest NaN
    st0 = get_statistics(DeviceResiduals(r[:10].cuda(), torch.zeros(10, dtype=torch.bool).cuda()), None)
    assert st0.count_total == 0 and np.isnan(st0.median) and np.isnan(st0.MAE)
    # two valid values: even count -> mean of the two
    st2 = get_statistics(DeviceResiduals(torch.tensor([1.0, -3.0, 7.0], dtype=torch.float64).cuda(),
                                         torch.tensor([True, True, False]).cuda()), None)
    assert st2.median == -1.0 and st2.absolute_median == 2.0 and st2.MAE == 2.0


@pytest.mark.gpu
def test_cuda_tile_std_matches_reference_golden():
    from resdepth_b200.lib.tiles import DeviceTileProducer
    from resdepth_b200.lib.utils import compute_local_dsm_std_per_centered_patch
    dsm = G['std_dsm']
    prod = DeviceTileProducer(dsm, dsm, None, NODATA, int(G['std_tile']), 'geom')
    pos = [tuple(int(v) for v in p) for p in G['std_pos']]
    std = compute_local_dsm_std_per_centered_patch(prod, 'raster_in', pos)
    assert abs(std - float(G['std_value'])) <= 1e-9 * float(G['std_value'])
    with pytest.raises(ValueError):
        compute_local_dsm_std_per_centered_patch(prod, 'raster_in', [(1000, 0)])
