"""Profiling driver: N train steps of the benchmark workload (3-ch 256x256, depth 5, batch 64) with no timing
code around them.  Used under ncu (see profiles/README.md); prints nothing that is a benchmark value."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resdepth_b200 import _native  # noqa: E402
from resdepth_b200.lib.optim import Adam  # noqa: E402
from resdepth_b200.lib.UNet import UNet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = sys.argv[3] if len(sys.argv) > 3 else 'train'
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = UNet(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True).to(dev)
opt = Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
g = torch.Generator().manual_seed(1234)
x = torch.randn(B, 3, 256, 256, generator=g)
tgt = (x[:, :1] + 0.1 * torch.randn(B, 1, 256, 256, generator=g)).to(dev)
mask = (torch.rand(B, 1, 256, 256, generator=g) > 0.05).to(dev).view(torch.uint8)
x = x.to(dev)
mean = torch.full((B,), 400.0, device=dev)
std = torch.full((B,), 3.5, device=dev)
loss = torch.empty(1, device=dev)
s = torch.cuda.current_stream().cuda_stream
if mode == 'eval':
    model.eval()
for i in range(steps):
    with torch.no_grad():
        if mode == 'eval':
            y = model(x)
            continue
        y = model._forward_native(x, _native.FWD_TRAIN)
        h = model.native_handle(dev)
        dy = torch.empty_like(y)
        h.loss(y.data_ptr(), tgt.data_ptr(), mask.data_ptr(), mean.data_ptr(), std.data_ptr(), loss.data_ptr(),
               dy.data_ptr(), B, 256, s)
        grads = model._backward_native(x, dy, detach_copy=False)
        for p, gr in zip(model.parameters(), grads):
            p.grad = gr
        opt.step()
torch.cuda.synchronize()
print('done', float(loss) if mode != 'eval' else float(y.sum()))
