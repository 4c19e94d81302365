"""GPU parity tests (run with ``-m gpu`` on the B200 box): the CUDA path, called through the C ABI, against
(a) the committed golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded inputs.

Tolerances.  fp32 math mode (CUDA-core FMA): y within 1e-4 absolute of the reference output (fp32 summation
order only).  TF32 math mode (tcgen05 kind::tf32): the bar of BASELINE.json -- relative L2 error of the height
residual r = y - x0 at most 1e-3, identical argmax|r| per tile, height MAE within 1e-3 m at sigma = 3.5 m.
"""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
from tests.cases import CASES, NATIVE_CASES, batch_of, load_golden, spec_of

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _model(kwargs):
    from resdepth_b200.lib.UNet import UNet
    torch.manual_seed(0)
    return UNet(**kwargs)


def _cuda_batch(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def _train_step(model, batch, opt=None):
    """Forward(train) + fused loss + backward through the C ABI; returns (y, loss, grads by name)."""
    from resdepth_b200 import _native
    b = _cuda_batch(batch)
    model.train()
    with torch.no_grad():
        y = model._forward_native(b['input'], _native.FWD_TRAIN)
        h = model.native_handle(torch.device(DEV))
        loss = torch.empty(1, device=DEV)
        dy = torch.empty_like(y)
        h.loss(y.data_ptr(), b['target'].data_ptr(), b['loss_mask'].view(torch.uint8).data_ptr(),
               b['dsm_mean'].data_ptr(), b['dsm_std'].data_ptr(), loss.data_ptr(), dy.data_ptr(),
               y.shape[0], y.shape[2], torch.cuda.current_stream().cuda_stream)
        grads = model._backward_native(b['input'], dy, detach_copy=True)
    named = {n: g for (n, _), g in zip(model.named_parameters(), grads)}
    if opt is not None:
        for p, g in zip(model.parameters(), grads):
            p.grad = g
        opt.step()
    return y, float(loss.item()), named, dy


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(params=['fp32', 'tf32'])
def math_mode(request, monkeypatch):
    monkeypatch.setenv('RESDEPTH_MATH', request.param)
    return request.param


@pytest.mark.parametrize('name', NATIVE_CASES)
def test_train_step_and_eval_against_reference_golden(name, math_mode):
    from resdepth_b200.lib.optim import Adam
    kwargs, B, T = CASES[name]
    g = load_golden(name)
    model = _model(kwargs).to(DEV)
    batch = batch_of(name)
    opt = Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    y, loss, grads, _ = _train_step(model, batch, opt)
    y_ref = torch.from_numpy(g['y_train'])
    x0 = batch['input'][:, :1]
    y_cpu = y.cpu()
    if math_mode == 'fp32':
        np.testing.assert_allclose(y_cpu.numpy(), g['y_train'], rtol=0, atol=1e-4)
        assert abs(loss - float(g['loss_train'])) < 1e-5 * abs(float(g['loss_train']))
        gtol = 5e-3
    else:
        if kwargs.get('outer_skip', True):
            rel, same, mae = O.residual_metrics(y_cpu, y_ref, x0)
            assert rel <= 1e-3, rel
            assert same
            assert mae <= 1e-3, mae
        else:
            assert _rel(y_cpu, y_ref) <= 1e-3
        assert abs(loss - float(g['loss_train'])) < 2e-3 * abs(float(g['loss_train']))
        gtol = 6e-2          # per-tensor norms of tiny bias gradients move by a few % under TF32 sign flips
    pkeys = [str(k) for k in g['param_keys']]
    gn = np.array([float(grads[k].double().norm()) for k in pkeys])
    np.testing.assert_allclose(gn, g['grad_norm'], rtol=gtol, atol=1e-5)
    for k in g.files:
        if k.startswith('grad::'):
            assert _rel(grads[k[6:]].cpu(), torch.from_numpy(g[k])) <= (2e-3 if math_mode == 'fp32' else 3e-2), k
    # running statistics and the step counter moved exactly once
    sd = model.state_dict()
    for k in sd:
        if k.endswith('num_batches_tracked'):
            assert int(sd[k]) == 1
    # the first Adam step moves each weight by ~lr*sign(grad): per-tensor sums are stable only up to sign flips of
    # noise-level gradients (tolerance grows with tensor size; the kernel itself is checked element-wise below)
    keys = [str(s) for s in g['keys']]
    post = np.array([float(sd[k].double().sum()) for k in keys])
    tol = np.array([2e-4 * max(2.0, (0.004 if math_mode == 'fp32' else 0.1) * sd[k].numel()) for k in keys])
    assert np.all(np.abs(post - g['post_sum']) <= tol), float(np.abs(post - g['post_sum']).max())


@pytest.mark.parametrize('name', NATIVE_CASES)
def test_forward_backward_against_oracle_same_weights(name, math_mode):
    """Oracle and CUDA path start from identical weights (no optimizer step in between): train-mode forward,
    loss, every gradient, updated running statistics, then eval-mode forward."""
    kwargs, B, T = CASES[name]
    spec = spec_of(kwargs)
    model = _model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys:
        sd[k].requires_grad_(True)
    batch = batch_of(name)
    loss_ref, grads_ref, y_ref = O.train_step(sd, pkeys, batch, spec, None)

    model = model.to(DEV)
    y, loss, grads, dy = _train_step(model, batch, None)
    tight = math_mode == 'fp32'
    assert _rel(y.cpu(), y_ref) <= (2e-6 if tight else 1e-3)
    assert abs(loss - loss_ref) <= (1e-5 if tight else 2e-3) * abs(loss_ref)
    # Per-parameter gradients of this network are differences of large cancelling sums (every conv is followed
    # by a BatchNorm): PyTorch's own fp32 result is only within ~5e-3 of its fp64 result on the small ones
    # (measured with the oracle on kat2), so the per-tensor bound is 1e-2 (fp32) / 1e-1 (TF32 activations flip
    # ReLU / pooling decisions); the concatenated gradient is held much tighter.
    for k in pkeys:
        ref = grads_ref[k]
        if float(ref.norm()) < 1e-7:
            assert float(grads[k].norm()) < 1e-5
            continue
        assert _rel(grads[k].cpu(), ref) <= (1e-2 if tight else 1e-1), (k, _rel(grads[k].cpu(), ref))
    flat = torch.cat([grads[k].cpu().flatten() for k in pkeys])
    flat_ref = torch.cat([grads_ref[k].flatten() for k in pkeys])
    # TF32 forward + bf16 backward: the two BASELINE-sized cases sit at ~1e-3; the tiny constructor variants
    # (2 tiles of 32x32, 32 filters) move between 2e-3 and 1.2e-2 when a forward value changes in its last fp32
    # bit (a LeakyReLU / pooling decision flips), so they get 2e-2.
    flat_tol = 1e-3 if tight else (5e-3 if name in ('kat1', 'kat2') else 2e-2)
    assert _rel(flat, flat_ref) <= flat_tol, _rel(flat, flat_ref)
    msd = model.state_dict()
    for k, v in sd.items():
        if 'running_' in k:
            np.testing.assert_allclose(msd[k].cpu().numpy(), v.detach().numpy(), rtol=1e-4 if tight else 2e-3,
                                       atol=1e-5 if tight else 1e-3)
    # eval-mode forward with the updated running statistics
    with torch.no_grad():
        y_eval_ref = O.unet_forward(sd, batch['input'], spec, training=False)
        model.eval()
        y_eval = model(batch['input'].to(DEV)).cpu()
    if tight:
        assert _rel(y_eval, y_eval_ref) <= 5e-6
    elif kwargs.get('outer_skip', True):
        rel, same, mae = O.residual_metrics(y_eval, y_eval_ref, batch['input'][:, :1])
        assert rel <= 1e-3 and same and mae <= 1e-3, (rel, same, mae)


def test_autograd_path_fills_param_grad_like_the_fused_path(math_mode):
    kwargs, B, T = CASES['var_base']
    batch = _cuda_batch(batch_of('var_base'))
    m1 = _model(kwargs).to(DEV)
    m2 = copy.deepcopy(m1)
    _, _, fused, _ = _train_step(m1, {k: v.cpu() for k, v in batch.items()}, None)
    m2.train()
    y = m2(batch['input'])
    assert y.requires_grad
    s = batch['dsm_std'].view(-1, 1, 1, 1)
    msk = batch['loss_mask'].float()
    loss = (msk * s * (y - batch['target']).abs()).sum() / msk.sum()
    loss.backward()
    for n, p in m2.named_parameters():
        assert p.grad is not None, n
        assert _rel(p.grad, fused[n]) <= 1e-5, n
    # a second forward invalidates the first graph
    y1 = m2(batch['input'])
    _ = m2(batch['input'])
    with pytest.raises(RuntimeError):
        y1.sum().backward()


def test_loss_kernel_matches_reference_formula():
    from resdepth_b200.lib.UNet import UNet
    kwargs, B, T = CASES['var_base']
    model = _model(kwargs).to(DEV)
    h = model.native_handle(torch.device(DEV))
    h.reserve(B, T, False)
    b = O.synthetic_batch(5, 1, 64, seed=11)
    b['dsm_mean'] = torch.tensor([400., 512.5, 13., 1050.25, 0.])
    b['dsm_std'] = torch.tensor([3.5, 1.25, 7., 0.5, 2.])
    yp = (b['input'] + 0.2 * torch.randn(5, 1, 64, 64, generator=torch.Generator().manual_seed(2))).requires_grad_(True)
    ref = O.denormalized_l1(yp, b['target'], b['loss_mask'], b['dsm_mean'], b['dsm_std'])
    ref.backward()
    d = _cuda_batch(b)
    loss = torch.empty(1, device=DEV)
    dy = torch.empty(5, 1, 64, 64, device=DEV)
    ypd = yp.detach().to(DEV)
    h.loss(ypd.data_ptr(), d['target'].data_ptr(), d['loss_mask'].view(torch.uint8).data_ptr(), d['dsm_mean'].data_ptr(),
           d['dsm_std'].data_ptr(), loss.data_ptr(), dy.data_ptr(), 5, 64, torch.cuda.current_stream().cuda_stream)
    assert abs(float(loss) - float(ref)) <= 2e-6 * float(ref)
    np.testing.assert_allclose(dy.cpu().numpy(), yp.grad.numpy(), rtol=1e-5, atol=1e-12)


def test_adam_and_sgd_kernels_match_torch_cpu():
    from resdepth_b200.lib.optim import SGD, Adam
    gen = torch.Generator().manual_seed(4)
    shapes = [(64, 3, 3, 3), (64,), (1,), (128, 64, 3, 3), (7,)]
    ref = [torch.randn(s, generator=gen).requires_grad_(True) for s in shapes]
    for cls_ref, cls_ours in ((torch.optim.Adam, Adam), (torch.optim.SGD, SGD)):
        ours = [p.detach().clone().to(DEV).requires_grad_(True) for p in ref]
        o_ref = cls_ref(ref, lr=2e-4, weight_decay=1e-5)
        o_ours = cls_ours(ours, lr=2e-4, weight_decay=1e-5)
        for _ in range(3):
            for p, q in zip(ref, ours):
                gr = torch.randn(p.shape, generator=gen)
                p.grad = gr.clone()
                q.grad = gr.to(DEV)
            o_ref.step()
            o_ours.step()
        for p, q in zip(ref, ours):
            np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().numpy(), rtol=2e-6, atol=1e-7)
    # state_dict round trip keeps the PyTorch format
    sd = o_ours.state_dict()
    assert set(sd.keys()) == {'state', 'param_groups'}


@pytest.mark.parametrize('name', ['blend_a', 'blend_b', 'blend_c'])
def test_blend_kernel_matches_reference_golden(name, golden_dir):
    from resdepth_b200.lib.evaluation import blend_tiles_into
    g = np.load(os.path.join(golden_dir, 'blend.npz'))
    rows, cols, tile, stride = [int(v) for v in g[name + '_geom']]
    pos, box = g[name + '_pos'], g[name + '_box']
    geom = np.concatenate([pos, box], axis=1).astype(np.int32)
    raster = torch.zeros(rows, cols, dtype=torch.float64, device=DEV)
    blend_tiles_into(raster, torch.from_numpy(g[name + '_tiles']).to(DEV), torch.from_numpy(g[name + '_mean']).to(DEV),
                     torch.from_numpy(g[name + '_std']).to(DEV), torch.from_numpy(geom).to(DEV), tile, stride)
    np.testing.assert_allclose(raster.cpu().numpy(), g[name + '_raster'], rtol=0, atol=1e-9)


def test_input_validation_and_error_paths():
    kwargs, B, T = CASES['var_base']
    model = _model(kwargs).to(DEV)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 32, 32))                      # CPU tensor: no fallback
    with pytest.raises(ValueError):
        model(torch.zeros(1, 2, 32, 32, device=DEV))          # wrong channel count
    with pytest.raises(RuntimeError):
        model.eval()
        with torch.no_grad():
            model(torch.zeros(1, 3, 30, 30, device=DEV))      # tile not a multiple of 2^depth
    with torch.no_grad():
        y = model(torch.zeros(0 + 1, 3, 8, 8, device=DEV))    # smallest legal tile for depth 2 ... 2^depth = 4
    assert y.shape == (1, 1, 8, 8)


def test_eval_tiles_are_independent_at_full_size(math_mode):
    """BASELINE config 2 (3-ch 256x256, depth 5, batch 32, forward only): in eval mode every tile is independent,
    so one batch-32 call equals four batch-8 calls; two tiles are also checked against the CPU oracle."""
    kwargs = dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True)
    model = _model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).eval()
    x = O.synthetic_batch(32, 3, 256)['input']
    with torch.no_grad():
        xd = x.to(DEV)
        y_all = model(xd)
        y_parts = torch.cat([model(xd[i:i + 8]) for i in range(0, 32, 8)])
        # the tile plan of the deep layers depends on the batch (128- vs 256-column tiles): summation order only
        assert _rel(y_parts, y_all) <= (1e-6 if math_mode == 'fp32' else 1e-5)
        y_ref = O.unet_forward(sd, x[:2], spec_of(kwargs), training=False)
    rel, same, mae = O.residual_metrics(y_all[:2].cpu(), y_ref, x[:2, :1])
    assert rel <= (1e-5 if math_mode == 'fp32' else 1e-3) and same and mae <= 1e-3, (rel, same, mae)


def test_train_batch_replication_property_at_full_size(math_mode):
    """BASELINE config 3 (batch 64 train step): a batch made of 8 copies of 8 tiles has the same BatchNorm
    statistics, per-tile outputs, loss and parameter gradients as the 8 tiles alone."""
    kwargs = dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True)
    small = O.synthetic_batch(8, 3, 256)
    big = {k: torch.cat([v] * 8) for k, v in small.items()}
    m1 = _model(kwargs).to(DEV)
    m2 = copy.deepcopy(m1)
    y1, l1, g1, _ = _train_step(m1, small)
    y2, l2, g2, _ = _train_step(m2, big)
    assert _rel(y2[:8], y1) <= (1e-5 if math_mode == 'fp32' else 1e-3)
    assert _rel(y2[56:], y2[:8]) <= 1e-6
    assert abs(l1 - l2) <= 1e-4 * abs(l1)
    for k in g1:
        if float(g1[k].norm()) > 1e-6:
            assert _rel(g2[k], g1[k]) <= (1e-2 if math_mode == 'fp32' else 1e-1), k
    f1 = torch.cat([g1[k].flatten() for k in g1])
    f2 = torch.cat([g2[k].flatten() for k in g1])
    assert _rel(f2, f1) <= (1e-3 if math_mode == 'fp32' else 1e-2), _rel(f2, f1)


def test_trainer_three_steps_follow_the_oracle(tmp_path, math_mode):
    """resdepth_b200.lib.Trainer.inference_one_epoch (train) against the oracle running the same three steps."""
    from types import SimpleNamespace

    from resdepth_b200.lib.Trainer import Trainer
    kwargs, B, T = CASES['kat1']
    spec = spec_of(kwargs)
    model = _model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys:
        sd[k].requires_grad_(True)
    batches = [O.synthetic_batch(B, 1, T, seed=100 + i) for i in range(3)]
    opt_ref = torch.optim.Adam([sd[k] for k in pkeys], lr=2e-4, weight_decay=1e-5)
    ref_losses = [O.train_step(sd, pkeys, b, spec, opt_ref)[0] for b in batches]

    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=10)
    args = SimpleNamespace(trainloader=batches, valloader=batches[:1], model=model, optimizer=opt, scheduler=sched,
                           criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=1, evaluate_rate=1, save_model_rate=1,
                           freq_average_train_loss=1, save_dir=str(tmp_path), log_file=str(tmp_path / 'training.log'),
                           checkpoint_dir=str(tmp_path / 'ckpt'), tboard_log_dir=None, pretrained_path=None)
    tr = Trainer(args)
    assert tr.optimizer.__class__.__name__ == 'Adam'
    losses = []
    for b in batches:
        losses.append(tr.inference_one_batch(b, 'train')['MAE_metric'])
        tr.optimizer.step()
        for p in tr.model.parameters():
            p.grad = None
    tol = 1e-4 if math_mode == 'fp32' else 3e-3
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= tol * abs(b), (losses, ref_losses)
    val = tr.inference_one_batch(batches[0], 'val')['MAE_metric']
    with torch.no_grad():
        y = O.unet_forward(sd, batches[0]['input'], spec, training=False)
        val_ref = float(O.denormalized_l1(y, batches[0]['target'], batches[0]['loss_mask'], batches[0]['dsm_mean'],
                                          batches[0]['dsm_std']))
    assert abs(val - val_ref) <= (2e-4 if math_mode == 'fp32' else 5e-3) * abs(val_ref)
    # full loop incl. checkpoint files
    tr.train()
    assert os.path.isfile(tr.path_model_last)
    ck = torch.load(tr.path_model_last, weights_only=False)
    assert set(ck) >= {'epoch', 'model_state_dict', 'optimizer_state_dict', 'loss_train', 'loss_val'}
    assert list(ck['model_state_dict'].keys()) == list(model.state_dict().keys())


def test_predict_linear_blend_matches_oracle(math_mode):
    from types import SimpleNamespace

    from resdepth_b200.lib.evaluation import predict_linear_blend
    kwargs = dict(n_input_channels=2, start_kernel=32, depth=2, bias_conv_layer=True)
    spec = spec_of(kwargs)
    model = _model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    rows, cols, tile, stride = 80, 100, 32, 16
    pos, box = O.regular_grid((0, cols - 1), (0, rows - 1), tile, stride)
    gen = torch.Generator().manual_seed(9)
    n = len(pos)
    x = torch.randn(n, 2, tile, tile, generator=gen)
    mean = 400 + torch.randn(n, generator=gen)
    std = torch.full((n,), 3.5)
    with torch.no_grad():
        y_ref = O.unet_forward(sd, x, spec, training=False).numpy()
    ref = O.linear_blend(y_ref, mean.numpy(), std.numpy(), pos, box, rows, cols, tile, stride)
    batches = []
    for i in range(0, n, 5):
        sl = slice(i, min(i + 5, n))
        batches.append({'input': x[sl], 'dsm_mean': mean[sl], 'dsm_std': std[sl],
                        'patch_offset_y': torch.tensor([p[0] for p in pos[sl]]),
                        'patch_offset_x': torch.tensor([p[1] for p in pos[sl]]),
                        'patch_valid_pixels_uly': torch.tensor([b[0] for b in box[sl]]),
                        'patch_valid_pixels_ulx': torch.tensor([b[1] for b in box[sl]]),
                        'patch_valid_pixels_lry': torch.tensor([b[2] for b in box[sl]]),
                        'patch_valid_pixels_lrx': torch.tensor([b[3] for b in box[sl]])})

    class Loader(list):
        pass
    loader = Loader(batches)
    loader.dataset = SimpleNamespace(dsm_input_gdal=SimpleNamespace(RasterXSize=cols, RasterYSize=rows),
                                     tile_size=tile, stride=stride)
    out = predict_linear_blend(loader, model)
    assert out.dtype == np.float64 and out.shape == (rows, cols)
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-4 if math_mode == 'fp32' else 5e-3)


def test_evaluation_pipeline_blend_residuals_statistics(math_mode):
    """test.py's evaluation chain on the device -- predict_linear_blend -> compute_residuals -> get_statistics --
    against the oracle chain (UNet oracle -> blending oracle -> numpy masked statistics)."""
    from types import SimpleNamespace

    from oracle import stats_oracle as SO
    from resdepth_b200.lib.evaluation import compute_residuals, get_statistics, predict_linear_blend
    kwargs = dict(n_input_channels=1, start_kernel=32, depth=2, bias_conv_layer=True)
    spec = spec_of(kwargs)
    model = _model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    rows, cols, tile, stride = 64, 96, 32, 16
    pos, box = O.regular_grid((0, cols - 1), (0, rows - 1), tile, stride)
    gen = torch.Generator().manual_seed(21)
    n = len(pos)
    x = torch.randn(n, 1, tile, tile, generator=gen)
    mean = 400 + torch.randn(n, generator=gen)
    std = torch.full((n,), 2.0)
    with torch.no_grad():
        y_ref = O.unet_forward(sd, x, spec, training=False).numpy()
    ref_raster = O.linear_blend(y_ref, mean.numpy(), std.numpy(), pos, box, rows, cols, tile, stride)
    batches = []
    for i in range(0, n, 4):
        sl = slice(i, min(i + 4, n))
        batches.append({'input': x[sl], 'dsm_mean': mean[sl], 'dsm_std': std[sl],
                        'patch_offset_y': torch.tensor([p[0] for p in pos[sl]]),
                        'patch_offset_x': torch.tensor([p[1] for p in pos[sl]]),
                        'patch_valid_pixels_uly': torch.tensor([b[0] for b in box[sl]]),
                        'patch_valid_pixels_ulx': torch.tensor([b[1] for b in box[sl]]),
                        'patch_valid_pixels_lry': torch.tensor([b[2] for b in box[sl]]),
                        'patch_valid_pixels_lrx': torch.tensor([b[3] for b in box[sl]])})

    class Loader(list):
        pass
    loader = Loader(batches)
    loader.dataset = SimpleNamespace(dsm_input_gdal=SimpleNamespace(RasterXSize=cols, RasterYSize=rows),
                                     tile_size=tile, stride=stride)
    raster = predict_linear_blend(loader, model)
    rng = np.random.default_rng(4)
    nodata = -9999.0
    gt = (ref_raster + rng.standard_normal((rows, cols))).astype(np.float32)
    gt[rng.random((rows, cols)) < 0.05] = nodata
    mask_gt = rng.random((rows, cols)) > 0.1
    st = get_statistics(compute_residuals(raster, gt, nodata, mask_gt), 2.0)
    o = SO.get_statistics(SO.compute_residuals(ref_raster, gt, nodata, mask_gt), 2.0)
    tol = 2e-4 if math_mode == 'fp32' else 5e-3                      # metres: the blended prediction's own tolerance
    assert st.count_total == o['count_total']
    for k in ('MAE', 'RMSE', 'absolute_median', 'median', 'NMAD', 'diff_max', 'diff_min'):
        assert abs(st[k] - o[k]) <= tol, (k, st[k], o[k])
    for k in ('MAE', 'RMSE', 'absolute_median', 'median', 'NMAD'):
        assert abs(st.truncated[k] - o['truncated'][k]) <= 4 * tol, (k, st.truncated[k], o['truncated'][k])


EXTRA_SHAPES = [
    # kwargs, B, T  -- shapes the golden table does not cover: non-power-of-two tiles (masked partial GEMM tiles),
    # batch 1, wider inputs, channel counts that are multiples of 32 but not powers of two
    (dict(n_input_channels=2, start_kernel=32, depth=3, bias_conv_layer=True), 1, 96),
    (dict(n_input_channels=4, start_kernel=96, depth=2, bias_conv_layer=False), 3, 40),
    (dict(n_input_channels=6, start_kernel=64, max_filter_depth=128, depth=4, bias_conv_layer=True), 2, 48),
    (dict(n_input_channels=1, start_kernel=32, depth=1, bias_conv_layer=True), 5, 16),
]


@pytest.mark.parametrize('kwargs,B,T', EXTRA_SHAPES)
def test_extra_shapes_against_oracle(kwargs, B, T, math_mode):
    spec = spec_of(kwargs)
    model = _model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys:
        sd[k].requires_grad_(True)
    batch = O.synthetic_batch(B, kwargs['n_input_channels'], T, seed=77)
    loss_ref, grads_ref, y_ref = O.train_step(sd, pkeys, batch, spec, None)
    model = model.to(DEV)
    y, loss, grads, _ = _train_step(model, batch, None)
    tight = math_mode == 'fp32'
    rel, same, mae = O.residual_metrics(y.cpu(), y_ref, batch['input'][:, :1])
    assert rel <= (1e-5 if tight else 1e-3) and same and mae <= 1e-3, (rel, same, mae)
    assert abs(loss - loss_ref) <= (1e-5 if tight else 2e-3) * abs(loss_ref)
    flat = torch.cat([grads[k].cpu().flatten() for k in pkeys])
    flat_ref = torch.cat([grads_ref[k].flatten() for k in pkeys])
    assert _rel(flat, flat_ref) <= (1e-3 if tight else 2e-2), _rel(flat, flat_ref)     # tiny nets: see above


_ORACLE_CACHE = {}


def _oracle_step(kwargs, B, T, seed=1234):
    """One oracle train step (forward, loss, every gradient) on the synthetic batch; cached per configuration so
    that the two math modes share the CPU work."""
    key = (tuple(sorted(kwargs.items())), B, T, seed)
    if key not in _ORACLE_CACHE:
        torch.set_num_threads(os.cpu_count() or 1)
        model = _model(kwargs)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        pkeys = [k for k, _ in model.named_parameters()]
        for k in pkeys:
            sd[k].requires_grad_(True)
        batch = O.synthetic_batch(B, kwargs['n_input_channels'], T, seed=seed)
        loss_ref, grads_ref, y_ref = O.train_step(sd, pkeys, batch, spec_of(kwargs), None)
        _ORACLE_CACHE[key] = (batch, pkeys, loss_ref, {k: v.clone() for k, v in grads_ref.items()}, y_ref.clone())
    return _ORACLE_CACHE[key]


def _check_step_against_oracle(kwargs, B, T, math_mode, flat_tol_tf32):
    batch, pkeys, loss_ref, grads_ref, y_ref = _oracle_step(kwargs, B, T)
    model = _model(kwargs).to(DEV)
    y, loss, grads, _ = _train_step(model, batch, None)
    tight = math_mode == 'fp32'
    rel, same, mae = O.residual_metrics(y.cpu(), y_ref, batch['input'][:, :1])
    assert rel <= (1e-5 if tight else 1e-3) and same and mae <= 1e-3, (rel, same, mae)
    assert abs(loss - loss_ref) <= (1e-5 if tight else 2e-3) * abs(loss_ref), (loss, loss_ref)
    flat = torch.cat([grads[k].cpu().flatten() for k in pkeys])
    flat_ref = torch.cat([grads_ref[k].flatten() for k in pkeys])
    err = _rel(flat, flat_ref)
    assert err <= (1e-3 if tight else flat_tol_tf32), err
    for k in pkeys:                                            # every tensor, not only the dominant ones
        if float(grads_ref[k].norm()) > 1e-7:
            assert _rel(grads[k].cpu(), grads_ref[k]) <= (1e-2 if tight else 1e-1), (k, _rel(grads[k].cpu(), grads_ref[k]))
    return err


def test_depth6_512_tiles_against_oracle(math_mode):
    """BASELINE configs[4] shape (3-ch 512x512 tiles, U-Net depth 6; width cap of lib/UNet.py:152-155) on two tiles:
    train-mode forward, loss and EVERY gradient against the oracle (flat gradient <= 5e-3 like kat2)."""
    kwargs = dict(n_input_channels=3, start_kernel=64, depth=6, bias_conv_layer=True)
    _check_step_against_oracle(kwargs, 2, 512, math_mode, 5e-3)


def test_full_batch64_train_step_against_oracle(math_mode):
    """BASELINE configs[2] at its real size -- 64 tiles of 3x256x256, depth 5 -- against the CPU oracle running the
    same 64-tile step (forward output, loss, flat gradient and every per-tensor gradient)."""
    kwargs = dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True)
    _check_step_against_oracle(kwargs, 64, 256, math_mode, 5e-3)


def _trainer_for(model, batches, tmp_path, opt=None, sched=None, pretrained=None, n_epochs=1):
    from types import SimpleNamespace

    from resdepth_b200.lib.Trainer import Trainer
    opt = opt or torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    args = SimpleNamespace(trainloader=batches, valloader=batches[:1], model=model, optimizer=opt, scheduler=sched,
                           criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=n_epochs, evaluate_rate=1,
                           save_model_rate=1, freq_average_train_loss=1000, save_dir=str(tmp_path), log_file=None,
                           checkpoint_dir=str(tmp_path / 'ckpt'), tboard_log_dir=None, pretrained_path=pretrained)
    return Trainer(args)


def test_loss_trajectory_200_steps_bf16_backward_follows_fp32(tmp_path, monkeypatch):
    """Training behaviour, not one step: 200 Adam steps from the same seed on the same batches with (a) the exact
    CUDA-core fp32 path, (b) the default tcgen05 path (TF32 forward, bf16-operand backward) and (c) TF32 forward +
    TF32 backward, at the reference's learning rate (lib/config.py:100).  The loss must fall to less than half, and the
    mean loss of the last 50 steps must agree within 2 %.  Resolution of the test: the loss is still falling at step 200
    (0.26 -> 0.22 over the window), so a trajectory that is a few steps ahead or behind moves the window mean by ~1 %; a
    CPU study with the oracle moved it by 0.1 % / 0.5 % when every initial weight was perturbed by 0.1 % / 0.3 %, and
    across kernel revisions of round 2 (different summation orders only) the bf16-backward run landed between 0.3 % and
    1.4 % of the fp32 run."""
    kwargs = dict(n_input_channels=3, start_kernel=64, depth=3, bias_conv_layer=True)
    batches = [O.synthetic_batch(8, 3, 64, seed=700 + i) for i in range(10)]
    curves = {}
    for label, math, bwd in (('fp32', 'fp32', 'auto'), ('default', 'tf32', 'auto'), ('tf32bwd', 'tf32', 'tf32')):
        monkeypatch.setenv('RESDEPTH_MATH', math)
        model = _model(kwargs)
        model.backward_math = bwd
        tr = _trainer_for(model, batches, tmp_path / label)
        assert tr.model.native_handle().bwd_mode_name() == {'fp32': 'fp32', 'default': 'bf16', 'tf32bwd': 'tf32'}[label]
        losses = []
        for step in range(200):
            loss = tr._launch_batch(batches[step % len(batches)], 'train')
            tr.optimizer.step()
            losses.append(loss.clone())
        curves[label] = torch.cat(losses).cpu().double()
    ref = curves['fp32']
    assert float(ref[-50:].mean()) < 0.5 * float(ref[:20].mean()), 'the reference trajectory does not train'
    for label in ('default', 'tf32bwd'):
        c = curves[label]
        assert abs(float(c[-50:].mean()) - float(ref[-50:].mean())) <= 2e-2 * float(ref[-50:].mean()), \
            (label, float(c[-50:].mean()), float(ref[-50:].mean()))
        assert abs(float(c[:20].mean()) - float(ref[:20].mean())) <= 5e-3 * float(ref[:20].mean())


def test_graph_replay_is_bitwise_the_eager_step(tmp_path, monkeypatch):
    """Steady-state steps replay CUDA graphs of the same launches: losses and parameters after 6 steps are bit-equal
    to RESDEPTH_GRAPHS=0, through inference_one_epoch (static staging sets) and through device-resident batches."""
    kwargs, B, T = CASES['kat1']
    batches = [O.synthetic_batch(B, 1, T, seed=40 + i) for i in range(3)] * 2
    results = {}
    for graphs in ('0', '1'):
        monkeypatch.setenv('RESDEPTH_GRAPHS', graphs)
        model = _model(kwargs)
        tr = _trainer_for(model, batches, tmp_path / graphs)
        assert tr.use_graphs == (graphs == '1')
        tr.inference_one_epoch(0, 'train')
        dev_batch = _cuda_batch(batches[0])
        losses = []
        for _ in range(4):
            losses.append(float(tr.inference_one_batch(dev_batch, 'train')['MAE_metric']))
            tr.optimizer.step()
        val = tr.inference_one_batch(batches[1], 'val')['MAE_metric']
        results[graphs] = (losses, val, tr.model._rt['arena'].clone(), tr.model._rt['bufs'].clone(),
                           int(tr.model.state_dict()['bottleneck.1.num_batches_tracked']), len(tr._graphs))
    assert results['1'][5] >= 2 and results['0'][5] == 0          # graphs were actually captured and replayed
    assert results['0'][0] == results['1'][0] and results['0'][1] == results['1'][1]
    assert torch.equal(results['0'][2], results['1'][2]) and torch.equal(results['0'][3], results['1'][3])
    assert results['0'][4] == results['1'][4] == 10


def test_workspace_layouts_are_kept_across_shape_changes():
    """Alternating batch shapes (train batch / smaller validation batch / partial last batch) switches between
    cached workspace layouts: results do not change and the first layout stays alive (no re-carving)."""
    kwargs, B, T = CASES['kat1']
    model = _model(kwargs).to(DEV)
    h = model.native_handle(torch.device(DEV))
    big, small = O.synthetic_batch(4, 1, 64, seed=1), O.synthetic_batch(2, 1, 64, seed=2)
    y1, l1, g1, _ = _train_step(model, big)
    first = h.workspace_id()
    y2, l2, g2, _ = _train_step(model, small)
    assert h.workspace_id() != first and h.workspace_alive(first)
    model.eval()
    with torch.no_grad():
        model(small['input'][:1].to(DEV))
        model(big['input'].to(DEV))
    # identical weights and running statistics as before the first step? no: train steps moved the running stats only;
    # the parameters are untouched (no optimizer), so the same batch reproduces the same gradients bit for bit
    y3, l3, g3, _ = _train_step(model, big)
    assert h.workspace_id() == first
    assert torch.equal(y1, y3) and l1 == l3
    assert all(torch.equal(g1[k], g3[k]) for k in g1)
    for _ in range(6):                                            # more shapes than cached layouts: the oldest is freed
        _train_step(model, O.synthetic_batch(1 + _, 1, 32, seed=3))
    assert not h.workspace_alive(first)
    y4, l4, g4, _ = _train_step(model, big)
    assert torch.equal(y1, y4)


def test_constant_weights_block_reuses_packs_and_sees_new_weights_afterwards(math_mode):
    """model.constant_weights() (rd_freeze_params): eval forwards inside the block skip the weight pack / BatchNorm
    finalize launches and return bit-identical outputs, on every workspace layout the block touches; a weight change
    between two blocks is seen by the second one; a training forward inside a block invalidates that layout's packs."""
    from resdepth_b200 import _native
    kwargs, B, T = CASES['kat1']
    model = _model(kwargs).to(DEV)
    a = O.synthetic_batch(4, 1, 64, seed=1)['input'].to(DEV)
    b = O.synthetic_batch(2, 1, 64, seed=2)['input'].to(DEV)
    _train_step(model, O.synthetic_batch(4, 1, 64, seed=3))          # non-trivial running statistics
    model.eval()
    lib = _native.lib()
    with torch.no_grad():
        ya, yb = model(a), model(b)
        lib.rd_launch_count(1)
        model(a)
        plain = lib.rd_launch_count(1)
        with model.constant_weights():
            assert torch.equal(model(a), ya)                          # first call of the block packs
            lib.rd_launch_count(1)
            y2 = model(a)
            frozen = lib.rd_launch_count(1)
            assert torch.equal(y2, ya)
            assert frozen < plain                                     # no pack, no BatchNorm finalize launch
            assert torch.equal(model(b), yb) and torch.equal(model(b), yb)        # another layout packs once, too
            assert torch.equal(model(a), ya)
            model.train()
            model._forward_native(a, _native.FWD_TRAIN)               # overwrites the layout's BatchNorm vectors
            model.eval()
            y_after_train = model(a)
        assert torch.equal(y_after_train, model(a))                   # the block re-packed after the training forward
        for p in model.parameters():
            p.mul_(1.01)
        y_new = model(a)
        assert not torch.equal(y_new, y_after_train)
        with model.constant_weights():
            assert torch.equal(model(a), y_new) and torch.equal(model(a), y_new)


def test_staged_backward_equals_the_single_call():
    """rd_backward_stage 0,1,2 (the data-parallel schedule) writes bit-identical gradients to rd_backward, the stage
    ranges tile the gradient arena, and stages out of order are refused."""
    from resdepth_b200 import _native
    kwargs, B, T = CASES['kat2']
    batch = _cuda_batch(batch_of('kat2'))
    model = _model(kwargs).to(DEV).train()
    _, _, g_ref, dy = _train_step(model, {k: v.cpu() for k, v in batch.items()})
    h = model.native_handle(torch.device(DEV))
    spans = [h.grad_stage_range(s) for s in range(3)]
    assert spans[2][0] == 0 and spans[2][0] + spans[2][1] == spans[1][0] and spans[1][0] + spans[1][1] == spans[0][0]
    assert spans[0][0] + spans[0][1] == h.param_arena_size()
    with torch.no_grad():
        model._forward_native(batch['input'], _native.FWD_TRAIN)
        model._rt['grads'].fill_(float('nan'))
        done = []
        grads = model._backward_native(batch['input'], dy, detach_copy=True, on_stage_done=lambda sl: done.append(sl.numel()))
    assert done == [spans[0][1], spans[1][1], spans[2][1]]
    for (n, _), g in zip(model.named_parameters(), grads):
        assert torch.equal(g, g_ref[n]), n
    stream = torch.cuda.current_stream().cuda_stream
    with torch.no_grad():
        model._forward_native(batch['input'], _native.FWD_TRAIN)
        with pytest.raises(RuntimeError):
            h.backward_stage(batch['input'].data_ptr(), dy.data_ptr(), 1, stream)


def test_checkpoint_written_by_the_reference_classes_resumes(tmp_path, golden_dir, monkeypatch):
    """tests/golden/ref_checkpoint.pth was written by the UNMODIFIED reference Trainer._save_checkpoint
    (oracle/make_golden_checkpoint.py).  Loading it through our Trainer(pretrained_path=...) must continue exactly
    where the reference continues: eval output, next train loss, state after the next Adam step, scheduler state."""
    monkeypatch.setenv('RESDEPTH_MATH', 'fp32')
    g = np.load(os.path.join(golden_dir, 'ref_checkpoint.npz'))
    kwargs = dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True)
    B, T = int(g['B']), int(g['T'])
    torch.manual_seed(123)                                        # different initial weights: everything must come from the file
    from resdepth_b200.lib.UNet import UNet
    model = UNet(**kwargs)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=3, gamma=0.5)
    batches = [O.synthetic_batch(B, 3, T, seed=502)]
    tr = _trainer_for(model, batches, tmp_path, opt=opt, sched=sched,
                      pretrained=os.path.join(golden_dir, 'ref_checkpoint.pth'), n_epochs=3)
    assert tr.start_epoch == int(g['epoch']) + 1 and tr.n_epochs == 3 + tr.start_epoch
    assert tr.best_loss == float(g['loss_val']) and tr.index_best_loss == int(g['epoch'])
    assert abs(tr._get_lr() - float(g['lr'])) < 1e-15 and tr.scheduler.last_epoch == 4
    ev = tr.inference_one_batch(batches[0], 'val')['MAE_metric']
    assert abs(ev - float(g['loss_eval'])) <= 1e-5 * float(g['loss_eval'])
    tr.model.eval()
    with torch.no_grad():
        y = tr.model(batches[0]['input'].to(DEV)).cpu().numpy()
    np.testing.assert_allclose(y, g['y_eval'], rtol=0, atol=2e-5)
    nxt = tr.inference_one_batch(O.synthetic_batch(B, 3, T, seed=503), 'train')['MAE_metric']
    assert abs(nxt - float(g['loss_next'])) <= 1e-5 * float(g['loss_next'])
    tr.optimizer.step()
    sd = tr.model.state_dict()
    keys = [str(k) for k in g['keys']]
    assert keys == list(sd.keys())
    post = np.array([float(sd[k].double().sum()) for k in keys])
    # third Adam step (moments come from the file): per-tensor sums are stable up to sign flips of noise-level gradients
    tol = np.array([2e-4 * max(2.0, 0.004 * sd[k].numel()) for k in keys])
    assert np.all(np.abs(post - g['post_sum']) <= tol), float(np.abs(post - g['post_sum']).max())


def test_trainer_with_sgd_and_checkpoint_resume(tmp_path):
    from types import SimpleNamespace

    from resdepth_b200.lib.Trainer import Trainer
    kwargs, B, T = CASES['var_base']
    batches = [O.synthetic_batch(B, 3, T, seed=300 + i) for i in range(2)]

    def make(pretrained=None, n_epochs=1):
        model = _model(kwargs)
        opt = torch.optim.SGD(model.parameters(), lr=1e-3, weight_decay=1e-5)
        args = SimpleNamespace(trainloader=batches, valloader=batches[:1], model=model, optimizer=opt, scheduler=None,
                               criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=n_epochs, evaluate_rate=1,
                               save_model_rate=1, freq_average_train_loss=1, save_dir=str(tmp_path),
                               log_file=str(tmp_path / 'training.log'), checkpoint_dir=str(tmp_path / 'ckpt'),
                               tboard_log_dir=None, pretrained_path=pretrained)
        return Trainer(args)
    tr = make()
    assert tr.optimizer.__class__.__name__ == 'SGD'
    tr.train()
    ck = torch.load(tr.path_model_last, weights_only=False)
    tr2 = make(pretrained=tr.path_model_last)
    assert tr2.start_epoch == 1 and tr2.n_epochs == 2
    for k, v in tr2.model.state_dict().items():
        assert torch.equal(v.cpu(), ck['model_state_dict'][k].cpu()), k
    v0 = tr.inference_one_batch(batches[0], 'val')['MAE_metric']
    v1 = tr2.inference_one_batch(batches[0], 'val')['MAE_metric']
    assert abs(v0 - v1) <= 1e-6 * abs(v0)


def test_device_tile_producer_matches_reference_golden(golden_dir):
    """rd_make_tiles against the vectors of the unmodified reference DsmOrthoDataset.__getitem__ (all five channel
    configurations, every rotation / flip combination that was drawn, user-specified and per-tile means)."""
    from resdepth_b200.lib.tiles import DeviceTileProducer
    g = np.load(os.path.join(golden_dir, 'tiles.npz'))
    pairs = {0: [[0, 1], [2, 3], [1, 3]], 1: [[0, 1], [2, 3]], 2: [[0], [2]], 3: None, 4: [[3, 0]]}
    for ci, (channels, dmean, omean, n) in enumerate(g['cases']):
        prod = DeviceTileProducer(g['dsm_in'], g['dsm_gt'], g['orthos'] if pairs[ci] else None, float(g['nodata']),
                                  int(g['tile']), str(channels), pairs[ci],
                                  None if dmean == 'None' else float(dmean), 3.5,
                                  None if omean == 'None' else float(omean), 41.0)
        metas = [g[f'c{ci}_s{i}_meta'] for i in range(int(n))]
        batch = prod.make_batch([(int(m[0]), int(m[1])) for m in metas], [[int(v) for v in m[5:]] for m in metas],
                                [(int(m[2]), int(m[3]), int(m[4])) for m in metas])
        for i in range(int(n)):
            key = f'c{ci}_s{i}'
            np.testing.assert_array_equal(batch['loss_mask'][i].cpu().numpy(), g[key + '_mask'])
            np.testing.assert_allclose(float(batch['dsm_mean'][i]), float(g[key + '_mean']), rtol=1e-6)
            np.testing.assert_allclose(batch['input'][i].cpu().numpy(), g[key + '_input'], rtol=0, atol=2e-5)
            valid = g[key + '_mask']
            np.testing.assert_allclose(batch['target'][i].cpu().numpy()[valid], g[key + '_target'][valid], rtol=0, atol=2e-5)


def test_device_tile_producer_feeds_the_trainer(tmp_path):
    """Batches produced on the device go straight into Trainer.inference_one_batch (no host round trip)."""
    from types import SimpleNamespace

    from oracle import tile_oracle as TO
    from resdepth_b200.lib.tiles import DeviceTileProducer
    from resdepth_b200.lib.Trainer import Trainer
    rng = np.random.default_rng(5)
    rows, cols, T = 160, 200, 64
    gt = (400 + 3 * rng.standard_normal((rows, cols))).astype(np.float32)
    din = (gt + rng.standard_normal((rows, cols))).astype(np.float32)
    orth = (100 + 30 * rng.standard_normal((rows, cols, 2))).astype(np.float32)
    prod = DeviceTileProducer(din, gt, orth, -9999.0, T, 'geom-stereo', [[0, 1]], None, 3.5, None, 30.0)
    pos, views, aug = prod.draw(6)
    batch = prod.make_batch(pos, views, aug)
    for i in range(6):
        inp, tgt, mask, mean = TO.make_tile(din, gt, orth, pos[i][0], pos[i][1], T, views[i], -9999.0, 3.5, 30.0,
                                            'geom-stereo', None, None, aug[i][0], bool(aug[i][1]), bool(aug[i][2]))
        np.testing.assert_allclose(batch['input'][i].cpu().numpy(), inp, rtol=0, atol=2e-5)
        np.testing.assert_allclose(batch['target'][i].cpu().numpy(), tgt, rtol=0, atol=2e-5)
        np.testing.assert_array_equal(batch['loss_mask'][i].cpu().numpy(), mask)
    model = _model(dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True))
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    args = SimpleNamespace(trainloader=[batch], valloader=[batch], model=model, optimizer=opt, scheduler=None,
                           criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=1, evaluate_rate=1, save_model_rate=1,
                           freq_average_train_loss=1, save_dir=str(tmp_path), log_file=None,
                           checkpoint_dir=str(tmp_path / 'ckpt'), tboard_log_dir=None, pretrained_path=None)
    tr = Trainer(args)
    l0 = tr.inference_one_batch(batch, 'train')['MAE_metric']
    assert np.isfinite(l0) and l0 > 0


def test_eval_mode_autograd_gradients_against_oracle(math_mode):
    """model.eval() with autograd recording (RD_FWD_EVAL_SAVE): BatchNorm uses the running statistics as constants, so its
    backward has no mean / projection terms -- also through the first block's Gram-matrix weight gradient (c1 = c2 = 0)."""
    kwargs, B, T = CASES['kat1']
    spec = spec_of(kwargs)
    model = _model(kwargs)
    batch = batch_of('kat1')
    # non-trivial running statistics first: one oracle train-mode forward updates them
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        O.unet_forward(sd, batch['input'], spec, training=True)
    model.load_state_dict(sd)
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys:
        sd[k].requires_grad_(True)
    y_ref = O.unet_forward(sd, batch['input'], spec, training=False)
    (y_ref - batch['target']).square().mean().backward()
    model = model.to(DEV).eval()
    y = model(batch['input'].to(DEV))
    assert y.requires_grad
    (y - batch['target'].to(DEV)).square().mean().backward()
    tight = math_mode == 'fp32'
    assert _rel(y.detach().cpu(), y_ref.detach()) <= (5e-6 if tight else 1e-3)
    flat = torch.cat([p.grad.cpu().flatten() for _, p in model.named_parameters()])
    flat_ref = torch.cat([sd[k].grad.flatten() for k in pkeys])
    assert _rel(flat, flat_ref) <= (1e-3 if tight else 1e-2), _rel(flat, flat_ref)
    g0, g0_ref = dict(model.named_parameters())['encoder.0.0.0.weight'].grad.cpu(), sd['encoder.0.0.0.weight'].grad
    assert _rel(g0, g0_ref) <= (5e-3 if tight else 5e-2), _rel(g0, g0_ref)
