"""A/B timing of the train step on ONE box (box-to-box clocks differ by several %): alternates configurations given as
NAME=ENV1:VAL1,ENV2:VAL2 ... (environment knobs read at handle creation / launch) plus the overlap switch
(`overlap:0`), several rounds each, device-resident batches, eager launches, CUDA events.

    python profiles/ab_step.py base= nooverlap=overlap:0 bn3=RESDEPTH_BN_CTAS:3
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['RESDEPTH_GRAPHS'] = '0'
import bench  # noqa: E402

K, W, ROUNDS = 20, 5, 3
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
configs = []
for arg in sys.argv[1:]:
    name, _, spec = arg.partition('=')
    kv = dict(item.split(':') for item in spec.split(',') if item)
    configs.append((name, kv))
arms = {}
for name, kv in configs:
    for k, v in kv.items():
        if k != 'overlap':
            os.environ[k] = v
    arm = bench.Arm('cfg3', 64, dev, 0, n_sets=2)
    if 'overlap' in kv:
        arm.handle.set_overlap(kv['overlap'] != '0')
    for i in range(W):
        arm.device_step(i)
    arms[name] = arm
    for k in kv:
        os.environ.pop(k, None)
res = {n: [] for n in arms}
for r in range(ROUNDS):
    for name, arm in arms.items():
        for i in range(2):
            arm.device_step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            arm.device_step(i)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / K)
for name, v in res.items():
    print(f'{name:12s} ms/step: ' + ' '.join(f'{x:.3f}' for x in v) + f'   best {min(v):.3f}  -> {64 / min(v) * 1e3:.0f} tiles/s')
