"""Where the wall clock of predict_linear_blend goes (4096 x 4096 raster, 961 tiles, batches of 32 from pinned host
memory): per-phase CUDA-event / host timings of the same loop, run on the GPU box.  Not a benchmark value."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resdepth_b200.lib.evaluation import _raster_to_host, blend_tiles_into  # noqa: E402
from resdepth_b200.lib.UNet import UNet  # noqa: E402

dev = torch.device('cuda:0')
R, tile, stride, bs = 4096, 256, 128, 32
starts = list(range(0, R - tile + 1, stride))
pos = [(y, x) for y in starts for x in starts]
g = torch.Generator().manual_seed(5)
pool = [torch.randn(bs, 3, tile, tile, generator=g).pin_memory() for _ in range(4)]
torch.manual_seed(0)
net = UNet(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True).to(dev).eval()
raster = torch.zeros((R, R), dtype=torch.float64, device=dev)
nb = (len(pos) + bs - 1) // bs
geoms = []
for i in range(0, len(pos), bs):
    p = pos[i:i + bs]
    geoms.append(torch.tensor([[y, x, 0 if y == 0 else tile - stride, 0 if x == 0 else tile - stride,
                                tile - 1 if y == starts[-1] else stride - 1, tile - 1 if x == starts[-1] else stride - 1]
                               for (y, x) in p], dtype=torch.int32).to(dev))
mean = torch.full((bs,), 400.0, device=dev)
std = torch.full((bs,), 3.5, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


with torch.no_grad(), net.constant_weights(dev):
    for rep in range(2):
        t_h2d = t_fwd = t_blend = 0.0
        raster.zero_()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for b in range(nb):
            n = geoms[b].shape[0]
            e = [ev() for _ in range(4)]
            e[0].record()
            x = pool[b % 4][:n].to(dev, non_blocking=True)
            e[1].record()
            y = net(x)
            e[2].record()
            blend_tiles_into(raster, y, mean[:n], std[:n], geoms[b], tile, stride)
            e[3].record()
            torch.cuda.synchronize()
            t_h2d += e[0].elapsed_time(e[1]); t_fwd += e[1].elapsed_time(e[2]); t_blend += e[2].elapsed_time(e[3])
        w1 = time.perf_counter()
        out = raster.cpu().numpy()
        w2 = time.perf_counter()
        pin = torch.empty((R, R), dtype=torch.float64, pin_memory=True)
        w3 = time.perf_counter()
        pin.copy_(raster, non_blocking=True); torch.cuda.synchronize()
        w4 = time.perf_counter()
        w5 = time.perf_counter()
        out2 = _raster_to_host(raster)
        w6 = time.perf_counter()
        assert np.array_equal(out, out2)
        print(f'rep {rep}: serialised loop {1e3 * (w1 - w0):.1f} ms (h2d {t_h2d:.1f}, forward {t_fwd:.1f}, blend {t_blend:.1f}); '
              f'raster.cpu() {1e3 * (w2 - w1):.1f} ms; pinned alloc {1e3 * (w3 - w2):.1f} ms; pinned D2H {1e3 * (w4 - w3):.1f} ms; chunked pinned D2H + host copy {1e3 * (w6 - w5):.1f} ms')
