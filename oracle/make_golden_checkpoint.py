#!/usr/bin/env python
"""Generate tests/golden/ref_checkpoint.{pth,npz} with the UNMODIFIED reference classes (container only).

    python -m oracle.make_golden_checkpoint

The ``.pth`` is written by the reference's own ``Trainer._save_checkpoint`` (lib/Trainer.py:145-157) after two
train steps of the reference ``UNet`` + ``torch.optim.Adam`` + ``StepLR`` on the CPU; the ``.npz`` records what the
reference itself computes when it continues from that state (eval-mode output on a fresh batch, the loss of the next
train step, per-key sums of the state after the next Adam step).  ``tests/test_gpu_parity.py`` loads the file
through ``resdepth_b200.lib.Trainer._load_pretrain`` and must reproduce those numbers.
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
KWARGS = dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True)
B, T = 2, 16


def main():
    torch.set_num_threads(4)
    ref = ref_shims.import_reference()
    torch.manual_seed(0)
    model = ref.UNet.UNet(**KWARGS)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)           # lib/utils.py:329-331
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=3, gamma=0.5)
    tr = object.__new__(ref.Trainer.Trainer)            # skip __init__ (TensorBoard / log files / loaders)
    tr.model, tr.optimizer, tr.scheduler = model, opt, sched
    tr.device = torch.device('cpu')
    tr.criterion = torch.nn.L1Loss(reduction='mean')
    losses = []
    for seed in (500, 501):
        losses.append(tr.inference_one_batch(synthetic_batch(B, 3, T, seed=seed), 'train')['MAE_metric'])
        opt.step()
        for p in model.parameters():
            p.grad = None
    for _ in range(4):
        sched.step()                                    # lr has decayed once (step_size 3)
    tr._save_checkpoint(4, losses[-1], 0.987, os.path.join(OUT, 'ref_checkpoint.pth'))

    out = {'kwargs_depth': KWARGS['depth'], 'B': B, 'T': T, 'epoch': 4, 'loss_val': 0.987, 'lr': opt.param_groups[0]['lr']}
    captured = {}
    hook = model.register_forward_hook(lambda m, i, o: captured.__setitem__('y', o.detach().clone()))
    out['loss_eval'] = tr.inference_one_batch(synthetic_batch(B, 3, T, seed=502), 'val')['MAE_metric']
    out['y_eval'] = captured['y'].numpy()
    out['loss_next'] = tr.inference_one_batch(synthetic_batch(B, 3, T, seed=503), 'train')['MAE_metric']
    hook.remove()
    opt.step()
    sd = model.state_dict()
    out['keys'] = np.array(list(sd.keys()))
    out['post_sum'] = np.array([float(v.double().sum()) for v in sd.values()])
    out['post_abs'] = np.array([float(v.double().abs().sum()) for v in sd.values()])
    np.savez_compressed(os.path.join(OUT, 'ref_checkpoint.npz'), **out)
    print('ref_checkpoint: losses', losses, 'eval', out['loss_eval'], 'next', out['loss_next'], 'lr', out['lr'],
          os.path.getsize(os.path.join(OUT, 'ref_checkpoint.pth')), 'bytes')


if __name__ == '__main__':
    main()
