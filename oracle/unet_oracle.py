"""CPU oracle: functional restatement of the ResDepth hot path (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it follows (paths relative to the
upstream repo root).  The arithmetic of the path lives in PyTorch (third-party,
pinned ``torch==1.9.0`` upstream, ``requirements.txt:5``); the oracle therefore
states the path with ``torch.nn.functional`` primitives plus an explicit
BatchNorm formula, runs on the CPU in fp32 (or fp64 for a higher-precision
truth) and is pinned by ``tests/golden/*.npz`` which were produced by the
unmodified reference classes (see ``oracle/make_golden.py``).

The product (``resdepth_b200``) never imports this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default, lib/UNet.py:45,66,86
BN_MOMENTUM = 0.1    # nn.BatchNorm2d default
LRELU_SLOPE = 0.01   # nn.LeakyReLU default, lib/UNet.py:30


@dataclass
class NetSpec:
    """Constructor arguments of the reference network (lib/UNet.py:105-107)."""
    n_input_channels: int = 1
    start_kernel: int = 64
    max_filter_depth: int = 512
    depth: int = 8
    act_fn_encoder: str = 'relu'
    act_fn_decoder: str = 'relu'
    act_fn_bottleneck: str = 'relu'
    up_mode: str = 'transpose'
    do_BN: bool = True
    bias_conv_layer: bool = False
    outer_skip: bool = True
    outer_skip_BN: bool = False

    def widths(self) -> List[int]:
        # lib/UNet.py:152-155 -- 64*2^i, clipped at max_filter_depth
        w = [self.start_kernel * (2 ** i) for i in range(self.depth)]
        return [self.max_filter_depth if c > self.max_filter_depth else c for c in w]


def _act(h: torch.Tensor, kind: str, prelu_weight: Optional[torch.Tensor]) -> torch.Tensor:
    # lib/UNet.py:27-33
    if kind == 'relu':
        return F.relu(h, inplace=True)                 # nn.ReLU(inplace=True)
    if kind == 'lrelu':
        return F.leaky_relu(h, LRELU_SLOPE, inplace=True)
    if kind == 'prelu':
        return F.prelu(h, prelu_weight)
    raise ValueError(kind)


EXPLICIT_BN = False   # True: evaluate the BatchNorm formula term by term (cross-check of the ATen call)


def _bn(z: torch.Tensor, prefix: str, sd: Dict[str, torch.Tensor], training: bool,
        update_running: bool) -> torch.Tensor:
    """nn.BatchNorm2d forward (lib/UNet.py:45,66,86,193).

    Default: the same ATen operator nn.BatchNorm2d.forward dispatches to (``F.batch_norm`` with momentum 0.1,
    eps 1e-5), so the CPU baseline timed from this oracle executes the reference's own operator sequence.
    ``EXPLICIT_BN``: the formula written out -- train: batch mean / biased variance normalise, running stats
    get the unbiased variance with momentum 0.1; eval: running stats.
    """
    g, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if not EXPLICIT_BN:
        if training and not update_running:
            rm, rv = rm.clone(), rv.clone()
        out = F.batch_norm(z, rm, rv, g, b, training, BN_MOMENTUM, BN_EPS)
        if training and update_running:
            sd[prefix + '.num_batches_tracked'] += 1
        return out
    if training:
        n = z.numel() // z.shape[1]
        mu = z.mean(dim=(0, 2, 3))
        var = ((z - mu[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        if update_running:
            with torch.no_grad():
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mu.detach())
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * (n / max(n - 1, 1)))
                sd[prefix + '.num_batches_tracked'] += 1
    else:
        mu, var = rm, rv
    xhat = (z - mu[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    return xhat * g[None, :, None, None] + b[None, :, None, None]


def _conv_block(h, conv_prefix, bn_prefix, act_prefix, act_kind, sd, spec, training, update_running):
    # conv_block / bottleneck / inner Sequential of conv_up_block: lib/UNet.py:36-52,64-67,78-93
    if spec.do_BN:
        z = F.conv2d(h, sd[conv_prefix + '.weight'], None, stride=1, padding=1)
        z = _bn(z, bn_prefix, sd, training, update_running)
    else:
        z = F.conv2d(h, sd[conv_prefix + '.weight'], sd[conv_prefix + '.bias'], stride=1, padding=1)
    return _act(z, act_kind, sd.get(act_prefix + '.weight'))


def _upconv(h, prefix, sd, spec):
    # upconv(): lib/UNet.py:17-24
    if spec.up_mode == 'transpose':
        return F.conv_transpose2d(h, sd[prefix + '.weight'], sd[prefix + '.bias'], stride=2)
    h = F.interpolate(h, scale_factor=2, mode='bilinear')
    return F.conv2d(h, sd[prefix + '.1.weight'], sd[prefix + '.1.bias'])


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, spec: NetSpec, training: bool,
                 update_running: bool = True) -> torch.Tensor:
    """UNet.forward, lib/UNet.py:196-246; ``sd`` uses the reference's state_dict keys."""
    d = spec.depth
    bn = 1 if spec.do_BN else None
    act_i = 2 if spec.do_BN else 1
    skips = []
    h = x
    for i in range(d):                                     # lib/UNet.py:201-207
        p = f'encoder.{i}.0'
        a = _conv_block(h, p + '.0', p + '.1', f'{p}.{act_i}', spec.act_fn_encoder, sd, spec,
                        training, update_running)
        skips.append(a)
        h = F.max_pool2d(a, 2, 2)
    h = _conv_block(h, 'bottleneck.0', 'bottleneck.1', f'bottleneck.{act_i}', spec.act_fn_bottleneck,
                    sd, spec, training, update_running)    # lib/UNet.py:210
    for j in range(d):                                     # lib/UNet.py:213-224
        if j < d - 1:
            u = _upconv(h, f'decoder.{j}.0', sd, spec) + skips[-1 - j]
            p = f'decoder.{j}.1'
            h = _conv_block(u, p + '.0', p + '.1', f'{p}.{act_i}', spec.act_fn_decoder, sd, spec,
                            training, update_running)
        else:
            h = _upconv(h, f'decoder.{j}', sd, spec) + skips[-1 - j]
    y = F.conv2d(h, sd['last_layer.weight'], sd.get('last_layer.bias'), padding=1)   # lib/UNet.py:227
    if spec.outer_skip:                                    # lib/UNet.py:229-244
        x0 = x[:, 0:1]
        if spec.outer_skip_BN:
            x0 = _bn(x0, 'layer_outer_skip.0', sd, training, update_running)
        y = x0 + y
    return y


def denormalized_l1(y_pred, y, loss_mask, mean, std):
    """Trainer._compute_denormalized_loss, lib/Trainer.py:87-100 with
    denormalize_torch, lib/data_normalization.py:29-38 and L1Loss(mean), lib/utils.py:284-285.
    ``mean``/``std`` are per-sample vectors [B]."""
    m = mean.to(y.dtype).view(-1, 1, 1, 1)
    s = std.to(y.dtype).view(-1, 1, 1, 1)
    yp = y_pred * s + m
    yt = y * s + m
    keep = (loss_mask != 0)
    yp = torch.where(keep, yp, torch.zeros_like(yp))
    yt = torch.where(keep, yt, torch.zeros_like(yt))
    loss = (yp - yt).abs().mean()
    return loss * loss_mask.numel() / loss_mask.sum()


def masked_l1_closed_form(y_pred, y, loss_mask, std):
    """Algebraically equal form: sum m*sigma*|yhat-y| / sum m."""
    s = std.to(y.dtype).view(-1, 1, 1, 1)
    m = loss_mask.to(y.dtype)
    return (m * s * (y_pred - y).abs()).sum() / m.sum()


class RefAdam:
    """torch.optim.Adam exactly as the reference builds it (lib/utils.py:329-331):
    betas (0.9, 0.999), eps 1e-8, coupled L2 weight decay on every parameter."""

    def __init__(self, params: Sequence[torch.Tensor], lr=2e-4, weight_decay=1e-5):
        self.opt = torch.optim.Adam(list(params), lr=lr, weight_decay=weight_decay)

    def step(self):
        self.opt.step()


def adam_reference_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """Single-tensor statement of the Adam update; returns new (p, m, v)."""
    g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def train_step(sd: Dict[str, torch.Tensor], param_keys: Sequence[str], batch: dict, spec: NetSpec,
               optimizer: Optional[torch.optim.Optimizer]) -> Tuple[float, Dict[str, torch.Tensor], torch.Tensor]:
    """One Trainer.inference_one_batch('train') + optimizer.step() (lib/Trainer.py:159-199,218).

    ``sd`` tensors named in ``param_keys`` must be leaf tensors with requires_grad.
    Returns (loss value, {key: grad}, y_pred)."""
    for k in param_keys:
        sd[k].grad = None
    y_pred = unet_forward(sd, batch['input'], spec, training=True)
    loss = denormalized_l1(y_pred, batch['target'], batch['loss_mask'],
                           torch.flatten(batch['dsm_mean']), torch.flatten(batch['dsm_std']))
    loss.backward()
    grads = {k: sd[k].grad.detach().clone() for k in param_keys if sd[k].grad is not None}
    if optimizer is not None:
        optimizer.step()
    return float(loss.item()), grads, y_pred.detach()


# ----------------------------------------------------------------------------------------------
# Tiled inference with linear blending
# ----------------------------------------------------------------------------------------------

def regular_grid(x_extent: Tuple[int, int], y_extent: Tuple[int, int], tile: int, stride: int):
    """create_regular_grid for one rectangular region, lib/rasterutils.py:100-191.
    Returns tile origins [(uly, ulx)] and no-overlap boxes [(uly, ulx, lry, lrx)]."""
    pos, box = [], []
    uly = lry = y_extent[0]
    b_uly, b_lry = 0, stride - 1
    while lry < y_extent[1]:
        ulx = lrx = x_extent[0]
        b_ulx, b_lrx = 0, stride - 1
        lry = uly + tile - 1
        if lry >= y_extent[1]:
            b_uly += lry - y_extent[1]
            lry = y_extent[1]
            uly = y_extent[1] - tile + 1
            b_lry = tile - 1
        while lrx < x_extent[1]:
            lrx = ulx + tile - 1
            if lrx >= x_extent[1]:
                b_ulx += lrx - x_extent[1]
                lrx = x_extent[1]
                ulx = x_extent[1] - tile + 1
                b_lrx = tile - 1
            pos.append((int(uly), int(ulx)))
            box.append((int(b_uly), int(b_ulx), int(b_lry), int(b_lrx)))
            ulx += stride
            b_ulx = tile - stride
        uly += stride
        b_uly = tile - stride
    return pos, box


def blend_weights(tile: int, stride: int, ulx: int, uly: int, lrx: int, lry: int) -> np.ndarray:
    """_get_blend_weights, lib/evaluation.py:516-567 (float64 weights)."""
    w = np.ones((tile, tile))
    overlap = tile - stride
    ramp = np.linspace(0, 1, overlap, endpoint=True)
    if ulx > 0:
        w[:, ulx - overlap:ulx] *= ramp[None, :]
        w[:, 0:ulx - overlap] = 0
    if lrx < tile - 1:
        w[:, lrx + 1:] *= ramp[::-1][None, :]
    if uly > 0:
        w[uly - overlap:uly, :] *= ramp[:, None]
        w[0:uly - overlap, :] = 0
    if lry < tile - 1:
        w[lry + 1:, :] *= ramp[::-1][:, None]
    return w


def linear_blend(tiles_pred: np.ndarray, mean: np.ndarray, std: np.ndarray, pos, box,
                 rows: int, cols: int, tile: int, stride: int) -> np.ndarray:
    """Accumulation loop of predict_linear_blend, lib/evaluation.py:478-513, with
    denormalize_numpy (lib/data_normalization.py:41-53).  tiles_pred: [N,1,T,T] fp32."""
    out = np.zeros((rows, cols))
    for i, ((y, x), (uly, ulx, lry, lrx)) in enumerate(zip(pos, box)):
        # python-float (double) scalars times an fp32 array stay fp32 (numpy weak scalars),
        # as in denormalize_numpy where mean_i/std_i come from .tolist()
        den = tiles_pred[i, 0] * float(std[i]) + float(mean[i])
        out[y:y + tile, x:x + tile] += den * blend_weights(tile, stride, ulx, uly, lrx, lry)
    return out


# ----------------------------------------------------------------------------------------------
# Synthetic inputs shared by tests, goldens and the bench (SURVEY.md 8c/8d recipe)
# ----------------------------------------------------------------------------------------------

def synthetic_batch(B: int, C: int, T: int, seed: int = 1234, dsm_mean: float = 400.0, dsm_std: float = 3.5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, T, T, generator=g)
    tgt = x[:, :1] + 0.1 * torch.randn(B, 1, T, T, generator=g)
    mask = torch.rand(B, 1, T, T, generator=g) > 0.05
    return {
        'input': x, 'target': tgt, 'loss_mask': mask,
        'dsm_mean': torch.full((B,), dsm_mean), 'dsm_std': torch.full((B,), dsm_std),
    }


def residual_metrics(y_new: torch.Tensor, y_ref: torch.Tensor, x0: torch.Tensor, std: float = 3.5):
    """The parity bar of BASELINE.json: relative L2 error of the height residual r = y - x0,
    per-tile argmax|r| equality and height MAE in metres."""
    r_new = (y_new - x0).double().flatten(1)
    r_ref = (y_ref - x0).double().flatten(1)
    rel = float((r_new - r_ref).norm() / r_ref.norm())
    same_argmax = bool((r_new.abs().argmax(1) == r_ref.abs().argmax(1)).all())
    mae_m = float((std * (y_new.double() - y_ref.double())).abs().mean())
    return rel, same_argmax, mae_m
