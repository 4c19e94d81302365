"""Same-box library baseline (SURVEY.md 2b / 8d "secondary bar"): the ResDepth network written as an ordinary
``torch.nn`` module tree and trained with stock ``torch.optim.Adam``, so that on a B200 every convolution runs a
cuDNN engine and everything else an ATen kernel.  It never touches ``resdepth_b200``'s kernels and is not part of
the product: ``bench.py`` times it (``--impl cudnn`` and the ``cudnn_baseline`` key) next to the hand-written path.

Architecture restated from the reference's description of the model (lib/UNet.py:104-246): per encoder level
Conv3x3(no bias) - BatchNorm - ReLU - MaxPool2, a bottleneck block, per decoder level ConvTranspose2d(k2, s2) +
additive skip followed by a conv block, a final ConvTranspose2d + skip, Conv3x3(64 -> 1, bias) and the outer residual
on input channel 0.  The loss is the reference's de-normalised masked L1 (lib/Trainer.py:87-100) in vectorised form.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _block(c_in: int, c_out: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(c_in, c_out, 3, padding=1, bias=False), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True))


class TorchUNet(nn.Module):
    def __init__(self, n_input_channels=3, start_kernel=64, depth=5, max_filter_depth=512):
        super().__init__()
        widths = [min(start_kernel << i, max_filter_depth) for i in range(depth)]
        self.down = nn.ModuleList(_block(ci, co) for ci, co in zip([n_input_channels] + widths[:-1], widths))
        self.mid = _block(widths[-1], widths[-1])
        rev = widths[::-1]
        self.up = nn.ModuleList(nn.ConvTranspose2d(c, c, 2, stride=2) for c in rev)
        self.up_conv = nn.ModuleList(_block(ci, co) for ci, co in zip(rev[:-1], rev[1:]))
        self.head = nn.Conv2d(start_kernel, 1, 3, padding=1, bias=True)

    def forward(self, x):
        skips, h = [], x
        for blk in self.down:
            h = blk(h)
            skips.append(h)
            h = F.max_pool2d(h, 2)
        h = self.mid(h)
        for j, up in enumerate(self.up):
            h = up(h) + skips[-1 - j]
            if j < len(self.up_conv):
                h = self.up_conv[j](h)
        return self.head(h) + x[:, :1]


def masked_l1(y_pred, target, mask, mean, std):
    m = mean.view(-1, 1, 1, 1)
    s = std.view(-1, 1, 1, 1)
    keep = mask != 0
    yp = torch.where(keep, y_pred.float() * s + m, 0.0)
    yt = torch.where(keep, target * s + m, 0.0)
    return (yp - yt).abs().mean() * mask.numel() / mask.sum()


def time_train_steps(batches, steps, warmup, variant, model_kw, lr=2e-4, weight_decay=1e-5):
    """CUDA-event time (ms) of ``steps`` train steps (forward, loss, backward, Adam) after ``warmup``.
    variant: 'fp32' (NCHW, cuDNN TF32 convolutions as PyTorch enables them by default) or 'bf16_channels_last'
    (channels_last weights / inputs under bf16 autocast)."""
    dev = batches[0]['input'].device
    torch.manual_seed(0)
    model = TorchUNet(**model_kw).to(dev).train()
    cl = variant == 'bf16_channels_last'
    if cl:
        model = model.to(memory_format=torch.channels_last)
        batches = [dict(b, input=b['input'].contiguous(memory_format=torch.channels_last)) for b in batches]
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)

    def step(i):
        b = batches[i % len(batches)]
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=cl):
            y = model(b['input'])
        loss = masked_l1(y, b['target'], b['loss_mask'], b['dsm_mean'], b['dsm_std'])
        loss.backward()
        opt.step()
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(warmup + i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), float(loss.item())


def time_inference(x, steps, warmup, variant, model_kw):
    dev = x.device
    torch.manual_seed(0)
    model = TorchUNet(**model_kw).to(dev).eval()
    cl = variant == 'bf16_channels_last'
    if cl:
        model = model.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=cl):
        for _ in range(warmup):
            model(x)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model(x)
        e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)
