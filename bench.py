#!/usr/bin/env python
"""Benchmark of the ResDepth hot path on B200 (contract: see the round brief).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (N=1): BASELINE.json configs[2] -- ResDepth-stereo training step: 3-ch 256x256 tiles, U-Net depth 5,
batch 64 per GPU, Adam + masked L1; one "step" = forward + loss + backward (+ gradient all-reduce) + Adam.
Metric: 256x256 3-ch DSM tiles/sec per train step, whole job (all ranks).  Weak scaling: 64 tiles per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = '256x256 3-ch DSM tiles/sec per train step'
UNIT = 'tiles/s'
MODEL_KW = dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True)
TILE = 256
N_INPUT_SETS = 4          # distinct resident batches cycled through (4 x 71 MB of inputs > 126 MB L2)
TRAIN_GFLOP_PER_TILE = 59.165   # fwd + dgrad + wgrad, cfg A


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'],
                    bf16_tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']), source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback')


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port of lib/UNet.py + lib/Trainer.py step + torch.optim.Adam)
# --------------------------------------------------------------------------------------------------------
def cpu_train_steps(batch_tiles: int, steps: int, warmup: int, threads: int):
    """Times `steps` CPU train steps (after `warmup`) on `batch_tiles` tiles of the benchmark workload.
    Returns (tiles/s, seconds per step).  Uses oracle/ only as the reported baseline."""
    import torch
    from oracle import unet_oracle as O
    from resdepth_b200.lib.UNet import UNet           # parameter container only (ordinary nn modules on the CPU)
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = UNet(**MODEL_KW)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in pkeys], lr=2e-4, weight_decay=1e-5)
    spec = O.NetSpec(**MODEL_KW)
    batch = O.synthetic_batch(batch_tiles, MODEL_KW['n_input_channels'], TILE)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(sd, pkeys, batch, spec, opt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch_tiles * len(times) / total, total / len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample_tiles = 8
    steps, warmup = args.steps, max(1, min(args.warmup, 2))
    # bound the run to a few minutes: a CPU step on 8 tiles takes 1-3 s
    steps = max(1, min(steps, 20))
    t0 = time.perf_counter()
    tps, sec = cpu_train_steps(sample_tiles, steps, warmup, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': tps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'ResDepth-stereo training step: 3-ch 256x256, depth 5, batch 64/GPU, Adam+L1 '
                               '(BASELINE configs[2]; configs[3] at 8 GPUs)',
                   'tile': TILE, 'sample_tiles_per_step': sample_tiles, 'device': 'host CPU',
                   'note': 'each step is a bounded sample of the 64-tile batch (same network, tile size, loss, optimizer)'},
        'cpu_baseline': {'value': tps, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{steps} train steps of {sample_tiles} tiles (of the 64-tile batch) after '
                                   f'{warmup} warm-up, oracle port of lib/UNet.py + lib/Trainer.py step + '
                                   'torch.optim.Adam on all host threads'},
        'e2e': {'value': tps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace

    from resdepth_b200 import _native
    from resdepth_b200.lib.Trainer import Trainer
    from resdepth_b200.lib.UNet import UNet

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the native arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, T, C = args.batch, TILE, MODEL_KW['n_input_channels']

    torch.manual_seed(0)
    model = UNet(**MODEL_KW)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)       # lib/utils.py:329-331
    # synthetic inputs (SURVEY 8d recipe), generated on the host so e2e can start from pinned host memory
    g = torch.Generator().manual_seed(1234 + rank)
    host = []
    for _ in range(N_INPUT_SETS):
        x = torch.randn(B, C, T, T, generator=g)
        host.append({'input': x.pin_memory(), 'target': (x[:, :1] + 0.1 * torch.randn(B, 1, T, T, generator=g)).pin_memory(),
                     'loss_mask': (torch.rand(B, 1, T, T, generator=g) > 0.05).pin_memory(),
                     'dsm_mean': torch.full((B,), 400.0).pin_memory(), 'dsm_std': torch.full((B,), 3.5).pin_memory()})
    args_tr = SimpleNamespace(trainloader=[host[0]], valloader=[host[0]], model=model, optimizer=opt, scheduler=None,
                              criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=1, evaluate_rate=1,
                              save_model_rate=1, freq_average_train_loss=20, save_dir='', log_file=None,
                              checkpoint_dir='', tboard_log_dir=None, pretrained_path=None)
    import logging
    logging.getLogger('train_logger').addHandler(logging.NullHandler())
    logging.getLogger('train_logger').propagate = False
    tr = Trainer.__new__(Trainer)
    _init_quiet(tr, args_tr, dev)
    model.train()
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    handle = model.native_handle(dev)

    def device_step(i):
        b = resident[i % N_INPUT_SETS]
        loss = tr.device_step(b['input'], b['target'], b['loss_mask'], b['dsm_mean'], b['dsm_std'], True)
        tr.optimizer.step()
        return loss

    def e2e_step(i):
        stats = tr.inference_one_batch(host[i % N_INPUT_SETS], 'train')
        tr.optimizer.step()
        return stats['MAE_metric']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        if profile:
            handle.profile_enable(True)
        _native.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(steps):
            last = fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _native.launch_count()
        prof = handle.profile_read() if profile else None
        if profile:
            handle.profile_enable(False)
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof, clocks, last

    K, W = args.steps, max(args.warmup, 3)
    # secondary figure (BASELINE configs[1]): eval-mode forward only, 32 tiles per call, device-resident inputs
    infer_x = resident[0]['input'][:32].contiguous()

    def infer_step(i):
        model.eval()
        with torch.no_grad():
            return model(infer_x)
    infer_ms, _, _, _, _ = timed(infer_step, K, W)
    model.train()
    # tier-next figure: the on-device tile producer that feeds the step (DsmOrthoDataset.__getitem__
    # for 64 tiles per call: geom-stereo, 4 views on a 4096x4096 raster, random positions / pairs / rot90 / flips)
    from resdepth_b200.lib.tiles import DeviceTileProducer
    g = torch.Generator(device=dev).manual_seed(11)
    R = 4096
    prod = DeviceTileProducer.from_device(
        400.0 + 3.0 * torch.randn(R, R, device=dev, generator=g), 400.0 + 3.0 * torch.randn(R, R, device=dev, generator=g),
        100.0 + 30.0 * torch.randn(4, R, R, device=dev, generator=g), -9999.0, T, 'geom-stereo', [[0, 1], [2, 3], [1, 3]],
        None, 3.5, None, 40.0, permute_images_within_pair=True)

    def producer_step(i):
        return prod.sample_batch(B)['input']
    prod_ms, _, _, _, _ = timed(producer_step, K, W)
    del prod
    # headline pass: no profiling events, weight gradients overlapped on the side stream
    ms, launches, _, clocks, last_loss = timed(device_step, K, W)
    loss_value = float(last_loss.item())
    # per-kernel pass (roofline, breakdown): CUDA events around every kernel category; the side stream is switched off
    # so that each category is timed alone (concurrent kernels would inflate one another's brackets)
    handle.set_overlap(False)
    prof_ms, _, prof, _, _ = timed(device_step, K, 2, profile=True)
    handle.set_overlap(True)
    # end to end through the public API the reference's train.py drives: Trainer.inference_one_epoch over a loader of
    # K pinned HOST batches (per step: H2D copies of that step's inputs -- enqueued one batch ahead so they overlap
    # the previous step's compute --, forward, loss, backward, optimizer.step, D2H of the loss, consumed by the host
    # one iteration later so the GPU never waits for the host).  The un-pipelined
    # per-call figure (inference_one_batch on a host batch: copy, then compute) is kept beside it.
    def e2e_epoch(i):
        tr.loader['train'] = [host[j % N_INPUT_SETS] for j in range(K)]
        return tr.inference_one_epoch(0, 'train')['MAE_metric'].avg
    e2e_call_ms, _, _, _, _ = timed(e2e_step, K, 2)
    e2e_ms, _, _, e2e_clocks, _ = timed(e2e_epoch, 1, 1)
    tiles = B * world * K
    value = tiles / (ms * 1e-3)
    e2e_value = tiles / (e2e_ms * 1e-3)
    hb = host[0]
    h2d = sum(hb[k].numel() * hb[k].element_size() for k in ('input', 'target', 'loss_mask', 'dsm_mean', 'dsm_std'))

    if rank == 0:
        pk = peaks()
        adam_ms = None
        cats = {k: v for k, v in prof.items() if v['calls'] > 0}
        total_ms = sum(v['ms'] for v in cats.values())
        dom_name, dom = max(cats.items(), key=lambda kv: kv[1]['ms'])
        gemm_like = dom['flops'] > 0 and dom['flops'] / max(dom['bytes'], 1.0) > 50.0
        if gemm_like:
            achieved = dom['flops'] / (dom['ms'] * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                    'frac': achieved / pk['bf16_tflops_sustained'], 'traffic': None,
                    'peak_note': f"bf16 cuBLAS sustained ({pk['source']}); this kernel computes in "
                                 f"{handle.math_mode_name().split(' ')[0]}: tf32 tensor ceiling is half of it"}
        else:
            achieved = dom['bytes'] / (dom['ms'] * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                    'frac': achieved / pk['hbm_gbs'], 'traffic': None, 'peak_note': f"copy bandwidth ({pk['source']})"}
        roof.update(kernel=dom_name, launches_per_step=dom['launches'] / K, avg_launch_ms=dom['ms'] / max(dom['launches'], 1),
                    share_of_step=dom['ms'] / max(total_ms, 1e-9),
                    algorithmic_per_step={'gflop': dom['flops'] / K / 1e9, 'mbytes': dom['bytes'] / K / 1e6})
        traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.isfile(traffic_path):
            with open(traffic_path) as fh:
                roof['traffic'] = json.load(fh).get(dom_name)
        breakdown = {k: round(v['ms'] / K, 4) for k, v in sorted(cats.items(), key=lambda kv: -kv[1]['ms'])}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'tf32' if 'tf32' in handle.math_mode_name() else 'f32', 'data': 'synthetic',
            'config': {'workload': 'ResDepth-stereo training step: 3-ch 256x256, depth 5, batch 64/GPU, Adam+L1 '
                                   '(BASELINE configs[2]; configs[3] at 8 GPUs)',
                       'tiles_per_gpu': B, 'global_batch': B * world, 'tile': T, 'parallelism': f'dp{world}',
                       'l2': f'{N_INPUT_SETS} resident input sets cycled (inputs {N_INPUT_SETS * h2d / 1e6:.0f} MB and '
                             '~11 GB of activations per step exceed the 126 MB L2); no explicit flush'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': e2e_ms / K,
                    'api': 'resdepth_b200.lib.Trainer.inference_one_epoch over K pinned host batches (per step: H2D of '
                           'the batch -- enqueued one batch ahead --, forward, loss, backward, optimizer.step, 4-byte D2H '
                           'of the loss into pinned memory, read by the host one iteration later)',
                    'unpipelined_per_call': {'value': tiles / (e2e_call_ms * 1e-3), 'unit': UNIT,
                                             'api': 'Trainer.inference_one_batch(host batch) + optimizer.step'},
                    'clocks': e2e_clocks},
            'gpu_launches': launches,
            'roofline': roof,
            'step_tflops': value * TRAIN_GFLOP_PER_TILE / 1e3,
            'kernel_ms_per_step': breakdown,
            'kernel_ms_total_per_step': total_ms / K,
            'kernel_ms_note': f'separate pass of {K} steps with per-category CUDA events and the weight-gradient side '
                              f'stream off ({prof_ms / K:.3f} ms per step); the headline pass has neither',
            'loss': loss_value,
            'inference': {'workload': 'BASELINE configs[1]: eval-mode forward, 3-ch 256x256, depth 5, batch 32/GPU',
                          'value': 32 * world * K / (infer_ms * 1e-3), 'unit': UNIT, 'ms_per_call': infer_ms / K},
            'tile_producer': {'workload': 'rd_make_tiles: 64 geom-stereo 256x256 training tiles per call from a 4096x4096 '
                                          'raster with 4 views (host-drawn positions / pairs / rot90 / flips included)',
                              'value': B * world * K / (prod_ms * 1e-3), 'unit': UNIT, 'ms_per_call': prod_ms / K},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tps, sec = cpu_train_steps(8, 3, 1, threads)
            line['cpu_baseline'] = {'value': tps, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                    'sample': '3 train steps of 8 tiles after 1 warm-up (oracle port of lib/UNet.py + '
                                              'lib/Trainer.py step + torch.optim.Adam, all host threads); '
                                              f'{sec:.2f} s per 8-tile step'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _init_quiet(tr, args, dev):
    """Trainer.__init__ without log files / TensorBoard / the look-ahead batch fetch (bench only)."""
    import math

    import torch

    from resdepth_b200.lib.optim import fuse_optimizer
    from resdepth_b200.lib.Trainer import _NullWriter, _setup_logger
    tr.config = args
    tr.distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
    tr.rank = torch.distributed.get_rank() if tr.distributed else 0
    tr.world_size = torch.distributed.get_world_size() if tr.distributed else 1
    tr.writer = _NullWriter()
    tr.logger = _setup_logger('train_logger', None, to_console=False)
    tr.device = dev
    tr.model = args.model.to(dev)
    tr.optimizer = fuse_optimizer(args.optimizer)
    tr.scheduler = None
    tr.criterion = args.criterion
    tr.loader = {'train': args.trainloader, 'val': args.valloader}
    tr.best_loss = math.inf
    tr.freq_average_train_loss = 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', choices=['native', 'reference'], default='native')
    ap.add_argument('--batch', type=int, default=64, help='tiles per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_native(args)


if __name__ == '__main__':
    sys.exit(main())
