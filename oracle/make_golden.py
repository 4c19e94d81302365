#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (container only).

Usage (build container, /root/reference mounted):  python -m oracle.make_golden

What is recorded per case (all produced by the reference's own classes, fp32 CPU):
  * ``lib.UNet.UNet`` built under ``torch.manual_seed(0)``: per-key sum / abs-sum of the
    initial state_dict (pins seed-for-seed initialisation and the key set);
  * ``lib.Trainer.Trainer.inference_one_batch(batch, 'train')`` (unmodified; the object is
    created without running ``__init__`` so no TensorBoard/log files are written): train-mode
    forward output, loss value, per-parameter gradient L2 norms (float64);
  * ``torch.optim.Adam(lr=2e-4, weight_decay=1e-5).step()`` as ``lib/utils.py:329-331``
    builds it: per-key sum of the post-step state_dict (float64);
  * ``inference_one_batch(batch, 'val')`` + eval-mode forward after the step;
  * ``lib.evaluation._get_blend_weights`` / ``lib.rasterutils.create_regular_grid`` and the
    accumulation loop of ``predict_linear_blend`` on a synthetic raster.
The synthetic inputs follow SURVEY.md 8c (``oracle.unet_oracle.synthetic_batch``).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shims  # noqa: E402
from oracle.unet_oracle import synthetic_batch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = {
    # name: (ctor kwargs, B, T)
    'kat1': (dict(n_input_channels=1, start_kernel=64, depth=3, bias_conv_layer=True), 4, 64),
    'kat2': (dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True), 2, 256),
    'var_base': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True), 2, 32),
    'var_lrelu': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True,
                       act_fn_encoder='lrelu', act_fn_decoder='lrelu', act_fn_bottleneck='lrelu'), 2, 32),
    'var_prelu': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True,
                       act_fn_encoder='prelu', act_fn_decoder='prelu', act_fn_bottleneck='prelu'), 2, 32),
    'var_nobn': (dict(n_input_channels=2, start_kernel=32, depth=2, bias_conv_layer=True, do_BN=False), 2, 32),
    'var_outerbn': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=False,
                         outer_skip_BN=True), 2, 32),
    'var_noouter': (dict(n_input_channels=1, start_kernel=32, depth=2, bias_conv_layer=True,
                         outer_skip=False), 2, 32),
    'var_bilinear': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True,
                          up_mode='bilinear'), 2, 32),
    'var_cap': (dict(n_input_channels=3, start_kernel=32, max_filter_depth=64, depth=3,
                     bias_conv_layer=True), 3, 32),
}


def _bare_trainer(ref, model):
    tr = object.__new__(ref.Trainer.Trainer)        # skip __init__ (TensorBoard / log files / loaders)
    tr.model = model
    tr.device = torch.device('cpu')
    tr.criterion = torch.nn.L1Loss(reduction='mean')   # lib/utils.py:284-285
    return tr


def run_case(ref, name, kwargs, B, T):
    torch.manual_seed(0)
    model = ref.UNet.UNet(**kwargs)
    out = {}
    sd0 = model.state_dict()
    out['keys'] = np.array(list(sd0.keys()))
    out['init_sum'] = np.array([float(v.double().sum()) for v in sd0.values()])
    out['init_abs'] = np.array([float(v.double().abs().sum()) for v in sd0.values()])
    out['shapes'] = np.array([str(tuple(v.shape)) for v in sd0.values()])
    pkeys = [k for k, _ in model.named_parameters()]
    out['param_keys'] = np.array(pkeys)

    batch = synthetic_batch(B, kwargs['n_input_channels'], T)
    out['x_sum'] = np.float64(batch['input'].double().sum())
    out['x_abs'] = np.float64(batch['input'].double().abs().sum())
    tr = _bare_trainer(ref, model)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)   # lib/utils.py:329-331

    # train-mode forward output (hook: no second forward, running stats updated exactly once)
    captured = {}
    h = model.register_forward_hook(lambda m, i, o: captured.__setitem__('y', o.detach().clone()))
    stats = tr.inference_one_batch(batch, 'train')
    h.remove()
    out['y_train'] = captured['y'].numpy()
    out['loss_train'] = np.float64(stats['MAE_metric'])
    out['grad_norm'] = np.array([float(p.grad.double().norm()) for _, p in model.named_parameters()])
    out['grad_sum'] = np.array([float(p.grad.double().sum()) for _, p in model.named_parameters()])
    # a handful of full gradients for element-wise checks
    named = dict(model.named_parameters())
    for k in ('last_layer.weight', 'encoder.0.0.0.weight', f'decoder.{kwargs["depth"] - 1}.weight',
              f'decoder.{kwargs["depth"] - 1}.0.weight', f'decoder.{kwargs["depth"] - 1}.1.weight'):
        if k in named:
            out['grad::' + k] = named[k].grad.numpy().copy()
    opt.step()
    sd1 = model.state_dict()
    out['post_sum'] = np.array([float(v.double().sum()) for v in sd1.values()])
    out['post_abs'] = np.array([float(v.double().abs().sum()) for v in sd1.values()])

    h = model.register_forward_hook(lambda m, i, o: captured.__setitem__('y', o.detach().clone()))
    stats = tr.inference_one_batch(batch, 'val')
    h.remove()
    out['y_eval'] = captured['y'].numpy()
    out['loss_eval'] = np.float64(stats['MAE_metric'])
    np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **out)
    print(f'{name}: loss_train={out["loss_train"]:.8f} loss_eval={out["loss_eval"]:.8f} '
          f'params={sum(p.numel() for p in model.parameters())}')


def run_blend(ref):
    rng = np.random.default_rng(7)
    out = {}
    for name, (rows, cols, tile, stride) in {'blend_a': (80, 100, 32, 16), 'blend_b': (64, 64, 32, 16),
                                              'blend_c': (37, 90, 32, 16)}.items():
        if rows < tile:
            continue
        area = {'x_extent': [(0, cols - 1)], 'y_extent': [(0, rows - 1)]}
        pos, box = ref.rasterutils.create_regular_grid(area, tile_size=tile, stride=stride)
        n = len(pos)
        tiles = rng.standard_normal((n, 1, tile, tile)).astype(np.float32)
        mean = (400 + rng.standard_normal(n)).astype(np.float32)
        std = np.full(n, 3.5, np.float32)
        raster = np.zeros((rows, cols))
        den = ref.data_normalization.denormalize_numpy(torch.from_numpy(tiles), torch.from_numpy(mean),
                                                        torch.from_numpy(std))
        for i in range(n):
            uly, ulx, lry, lrx = box[i]
            w = ref.evaluation._get_blend_weights(tile, stride, ulx, uly, lrx, lry)
            y, x = pos[i]
            raster[y:y + tile, x:x + tile] += den[i, 0] * w
        out[name + '_geom'] = np.array([rows, cols, tile, stride])
        out[name + '_pos'] = np.array(pos)
        out[name + '_box'] = np.array(box)
        out[name + '_tiles'] = tiles
        out[name + '_mean'] = mean
        out[name + '_std'] = std
        out[name + '_raster'] = raster
    np.savez_compressed(os.path.join(OUT, 'blend.npz'), **out)
    print('blend: ok')


def main():
    torch.set_num_threads(8)
    ref = ref_shims.import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, (kw, B, T) in CASES.items():
        run_case(ref, name, kw, B, T)
    run_blend(ref)


if __name__ == '__main__':
    main()
