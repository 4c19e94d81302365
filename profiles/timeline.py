"""Two-stream timeline of one train step (batch 64, cfg3) from the library's own event brackets (RESDEPTH_TIMELINE=1):
every kernel category with its start / end in ms relative to the step's first bracket, weight gradients on the side
stream.  Timing events perturb the step slightly; use it to see what overlaps what, not for absolute numbers.

    RESDEPTH_TIMELINE=1 python profiles/timeline.py 2> gpurun_out/timeline.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['RESDEPTH_GRAPHS'] = '0'
import bench  # noqa: E402

dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
arm = bench.Arm('cfg3', 64, dev, 0, n_sets=2)
for i in range(6):
    arm.device_step(i)
torch.cuda.synchronize()
arm.handle.profile_enable(True)
arm.device_step(0)
torch.cuda.synchronize()
sys.stderr.write('STEP BEGIN\n')
arm.handle.profile_read()
sys.stderr.write('STEP END\n')
arm.handle.profile_enable(False)
