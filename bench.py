#!/usr/bin/env python
"""Benchmark of the ResDepth hot path on B200 (contract: see the round brief).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores
    python bench.py --impl cudnn --steps K ...               # the same network through stock PyTorch / cuDNN on the GPU
    python bench.py --config cfg5 --gpus N ...               # BASELINE configs[4] as the headline workload

Workload (default): BASELINE.json configs[2] -- ResDepth-stereo training step: 3-ch 256x256 tiles, U-Net depth 5,
batch 64 per GPU, Adam + masked L1; one "step" = forward + loss + backward (+ gradient all-reduce) + Adam.
Metric: 256x256 3-ch DSM tiles/sec per train step, whole job (all ranks).  Weak scaling: 64 tiles per GPU
(configs[3] = 8 GPUs, global batch 512).
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
N_INPUT_SETS = 4          # distinct resident batches cycled through (4 x 71 MB of inputs > 126 MB L2)

# name -> (constructor arguments, tile, default tiles per GPU, train GFLOP per tile, description)
CONFIGS = {
    'cfg3': (dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True), 256, 64, 59.165,
             'ResDepth-stereo training step: 3-ch 256x256, depth 5, batch 64/GPU, Adam+L1 '
             '(BASELINE configs[2]; configs[3] at 8 GPUs)'),
    'cfg5': (dict(n_input_channels=3, start_kernel=64, depth=6, bias_conv_layer=True), 512, 32, 241.59,
             'ResDepth-stereo_generalized training step: 3-ch 512x512, depth 6, batch 32/GPU, Adam+L1 '
             '(BASELINE configs[4]: global batch 256 at 8 GPUs)'),
    'cfg1': (dict(n_input_channels=1, start_kernel=64, depth=3, bias_conv_layer=True), 64, 4, 2.364,
             'ResDepth-0 training step: 1-ch 64x64, depth 3, batch 4 (BASELINE configs[0]; launch-latency-bound)'),
}


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


def make_host_batches(B, C, T, seed, n_sets, pin=True):
    """Synthetic batches of the SURVEY 8d recipe, generated on the host (e2e starts from pinned host memory)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_sets):
        x = torch.randn(B, C, T, T, generator=g)
        b = {'input': x, 'target': x[:, :1] + 0.1 * torch.randn(B, 1, T, T, generator=g),
             'loss_mask': torch.rand(B, 1, T, T, generator=g) > 0.05,
             'dsm_mean': torch.full((B,), 400.0), 'dsm_std': torch.full((B,), 3.5)}
        out.append({k: (v.pin_memory() if pin else v) for k, v in b.items()})
    return out


# --------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port of lib/UNet.py + lib/Trainer.py step + torch.optim.Adam)
# --------------------------------------------------------------------------------------------------------
def cpu_train_steps(cfg: str, batch_tiles: int, steps: int, warmup: int, threads: int, budget_s: float = 1e9):
    """Times up to `steps` CPU train steps (after `warmup`) on `batch_tiles` tiles of the benchmark workload; stops
    early once `budget_s` seconds of timed steps have run.  Returns (tiles/s, seconds per step, steps timed).
    Uses oracle/ only as the reported baseline."""
    import torch
    from oracle import unet_oracle as O
    from resdepth_b200.lib.UNet import UNet           # parameter container only (ordinary nn modules on the CPU)
    kw, tile = CONFIGS[cfg][0], CONFIGS[cfg][1]
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = UNet(**kw)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in pkeys], lr=2e-4, weight_decay=1e-5)
    spec = O.NetSpec(**kw)
    batch = O.synthetic_batch(batch_tiles, kw['n_input_channels'], tile)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(sd, pkeys, batch, spec, opt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
            if sum(times) > budget_s:
                break
    total = sum(times)
    return batch_tiles * len(times) / total, total / len(times), len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return 0
    kw, tile, default_b, _, workload = CONFIGS[args.config]
    B = args.batch or default_b
    threads = os.cpu_count() or 1
    # full-size steps (the same 64-tile batch the GPU arm runs), the requested step / warm-up counts; the timed part is
    # cut once it has used its share of "a few minutes" (a 64-tile CPU step takes ~4 s on the box's 16 cores)
    steps, warmup = args.steps, args.warmup
    t0 = time.perf_counter()
    tps, sec, done = cpu_train_steps(args.config, B, steps, warmup, threads, budget_s=150.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': tps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done,
        'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload, 'tiles_per_gpu': B, 'global_batch': B, 'tile': tile, 'device': 'host CPU',
                   'parallelism': f'{threads} host threads',
                   'note': 'full-size steps of the GPU arm\'s batch on the host cores'
                           + ('' if done == steps else f'; stopped after {done} of {steps} steps (150 s budget)')},
        'cpu_baseline': {'value': tps, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{done} train steps of {B} tiles after {warmup} warm-up, oracle port of '
                                   'lib/UNet.py + lib/Trainer.py step + torch.optim.Adam on all host threads'},
        'e2e': {'value': tps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'wall_s': time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------
# Library baseline on the same GPU: stock PyTorch modules (cuDNN convolutions) + torch.optim.Adam
# --------------------------------------------------------------------------------------------------------
def cudnn_numbers(cfg, B, steps, warmup, dev, resident=None, with_inference=True):
    """tiles/s of baseline/torch_unet.py (never touches resdepth_b200's kernels) for both library variants."""
    import torch
    from baseline import torch_unet as TU
    kw, tile = CONFIGS[cfg][0], CONFIGS[cfg][1]
    mkw = dict(n_input_channels=kw['n_input_channels'], start_kernel=kw['start_kernel'], depth=kw['depth'])
    if resident is None:
        resident = [{k: v.to(dev) for k, v in hb.items()}
                    for hb in make_host_batches(B, kw['n_input_channels'], tile, 1234, N_INPUT_SETS, pin=False)]
    out = {'torch': torch.__version__, 'cudnn': torch.backends.cudnn.version(),
           'allow_tf32_conv': bool(torch.backends.cudnn.allow_tf32), 'tiles_per_step': B, 'steps': steps,
           'warmup': warmup, 'api': 'baseline/torch_unet.py: nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d / '
                                    'ConvTranspose2d modules, autograd, torch.optim.Adam (stock eager PyTorch)'}
    for variant in ('fp32', 'bf16_channels_last'):
        ms, loss = TU.time_train_steps(resident, steps, warmup, variant, mkw)
        out[f'train_{variant}'] = {'value': B * steps / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / steps, 'loss': loss}
        torch.cuda.empty_cache()
    if with_inference:
        x = resident[0]['input'][:32].contiguous()
        for variant in ('fp32', 'bf16_channels_last'):
            ms = TU.time_inference(x, steps, warmup, variant, mkw)
            out[f'inference_{variant}'] = {'value': x.shape[0] * steps / (ms * 1e-3), 'unit': UNIT, 'ms_per_call': ms / steps}
        torch.cuda.empty_cache()
    return out


def run_cudnn(args):
    import torch
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl cudnn needs a CUDA device')
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    kw, tile, default_b, _, workload = CONFIGS[args.config]
    B = args.batch or default_b
    sampler = ClockSampler(0)
    sampler.start()
    nums = cudnn_numbers(args.config, B, args.steps, max(args.warmup, 3), dev)
    clocks = sampler.stop()
    best = max(nums['train_fp32']['value'], nums['train_bf16_channels_last']['value'])
    line = {'impl': 'cudnn', 'metric': METRIC, 'value': best, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': B / best * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'tf32 (fp32 variant) / bf16 autocast (channels_last variant)',
            'data': 'synthetic', 'config': {'workload': workload, 'tiles_per_gpu': B, 'tile': tile,
                                            'note': 'value = the faster of the two library variants'},
            'clocks': clocks, 'cudnn_baseline': nums}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def roofline_of(prof, K, pk, math_name, traffic_key_prefix=''):
    """Roofline block of the dominant kernel category of a per-category profile (rd_profile_*)."""
    cats = {k: v for k, v in prof.items() if v['calls'] > 0}
    total_ms = sum(v['ms'] for v in cats.values())
    dom_name, dom = max(cats.items(), key=lambda kv: kv[1]['ms'])
    gemm_like = dom['flops'] > 0 and dom['flops'] / max(dom['bytes'], 1.0) > 50.0
    if gemm_like:
        achieved = dom['flops'] / (dom['ms'] * 1e-3) / 1e12
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['bf16_tflops_sustained'], 'traffic': None,
                'peak_note': f"bf16 cuBLAS sustained ({pk['source']}); this kernel computes in "
                             f"{math_name.split(' ')[0]}: the tf32 tensor ceiling is half of it"}
        if 'tf32' in math_name and dom_name.endswith('_fwd'):
            roof['frac_of_tf32_ceiling'] = achieved / (0.5 * pk['bf16_tflops_sustained'])
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
            roof['traffic'] = json.load(fh).get(traffic_key_prefix + dom_name)
    breakdown = {k: round(v['ms'] / K, 4) for k, v in sorted(cats.items(), key=lambda kv: -kv[1]['ms'])}
    # every memory-bound category against the HBM roofline, every contraction against the tensor ceiling
    fracs = {}
    for k, v in cats.items():
        if v['ms'] <= 0:
            continue
        if v['flops'] > 0 and v['flops'] / max(v['bytes'], 1.0) > 50.0:
            fracs[k] = {'tflops': round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1),
                        'frac_of_bf16_sustained': round(v['flops'] / (v['ms'] * 1e-3) / 1e12 / pk['bf16_tflops_sustained'], 3)}
        elif v['bytes'] > 0:
            fracs[k] = {'gbs': round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 0),
                        'frac_of_hbm': round(v['bytes'] / (v['ms'] * 1e-3) / 1e9 / pk['hbm_gbs'], 3)}
    return roof, breakdown, total_ms, fracs


class Arm:
    """One model + Trainer on this rank's GPU with resident and host copies of its synthetic batches."""

    def __init__(self, cfg, B, dev, rank, backward_math='auto', n_sets=N_INPUT_SETS, pin=True):
        import logging
        from types import SimpleNamespace

        import torch

        from resdepth_b200.lib.Trainer import Trainer
        from resdepth_b200.lib.UNet import UNet
        kw, tile = CONFIGS[cfg][0], CONFIGS[cfg][1]
        self.cfg, self.B, self.T, self.dev = cfg, B, tile, dev
        torch.manual_seed(0)
        self.model = UNet(**kw)
        self.model.backward_math = backward_math
        opt = torch.optim.Adam(self.model.parameters(), lr=2e-4, weight_decay=1e-5)       # lib/utils.py:329-331
        # every rank draws its own tiles (seed 1234 + rank): the loaders below are already per-rank shards, so the
        # Trainer's batch partition is switched off (it is exercised by tests/dist_gpu_check.py)
        self.host = make_host_batches(B, kw['n_input_channels'], tile, 1234 + rank, n_sets, pin=pin)
        args_tr = SimpleNamespace(trainloader=[self.host[0]], valloader=[self.host[0]], model=self.model, optimizer=opt,
                                  scheduler=None, criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=1,
                                  evaluate_rate=1, save_model_rate=1, freq_average_train_loss=20, save_dir='',
                                  log_file=None, checkpoint_dir='', tboard_log_dir=None, pretrained_path=None)
        logging.getLogger('train_logger').addHandler(logging.NullHandler())
        logging.getLogger('train_logger').propagate = False
        self.tr = Trainer.__new__(Trainer)
        _init_quiet(self.tr, args_tr, dev)
        self.tr.shard_batches = {'train': False, 'val': False}
        self.model.train()
        self.resident = [{k: v.to(dev) for k, v in hb.items()} for hb in self.host]
        self.handle = self.model.native_handle(dev)

    def device_step(self, i):
        b = self.resident[i % len(self.resident)]
        loss = self.tr.device_step(b['input'], b['target'], b['loss_mask'], b['dsm_mean'], b['dsm_std'], True)
        self.tr.optimizer.step()
        return loss

    def e2e_step(self, i):
        stats = self.tr.inference_one_batch(self.host[i % len(self.host)], 'train')
        self.tr.optimizer.step()
        return stats['MAE_metric']

    def e2e_epoch(self, K):
        self.tr.loader['train'] = [self.host[j % len(self.host)] for j in range(K)]
        return self.tr.inference_one_epoch(0, 'train')['MAE_metric'].avg

    def h2d_bytes(self):
        hb = self.host[0]
        return sum(hb[k].numel() * hb[k].element_size() for k in ('input', 'target', 'loss_mask', 'dsm_mean', 'dsm_std'))

    def close(self):
        import torch
        self.tr._graphs.clear()
        self.model._rt['handle'].close()
        self.model._rt.clear()
        del self.resident, self.host
        torch.cuda.empty_cache()


def run_native(args):
    import torch
    import torch.distributed as dist

    from resdepth_b200 import _native
    from resdepth_b200.lib.distributed import replicas_identical

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
    kw, T, default_b, gflop_per_tile, workload = CONFIGS[args.config]
    B, C = args.batch or default_b, kw['n_input_channels']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, arm=None, profile=False, prime=0):
        # `prime` untimed steps before the W warm-up steps: every resident input set must have been seen twice for its
        # CUDA graph to exist (capture happens at the second sighting), otherwise captures land in the timed region
        for i in range(prime):
            fn(i)
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        if profile:
            arm.handle.profile_enable(True)
        _native.launch_count(reset=True)
        if arm is not None:
            arm.tr.replayed_kernels = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(steps):
            last = fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _native.launch_count() + (arm.tr.replayed_kernels if arm is not None else 0)
        prof = arm.handle.profile_read() if profile else None
        if profile:
            arm.handle.profile_enable(False)
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof, clocks, last

    K, W = args.steps, max(args.warmup, 3)
    main = Arm(args.config, B, dev, rank)
    model, tr, handle = main.model, main.tr, main.handle
    math_name, bwd_name = handle.math_mode_name(), handle.bwd_mode_name()
    extras = {}

    # ---- secondary figure (BASELINE configs[1]): eval-mode forward only, 32 tiles per call, device-resident inputs
    if args.config == 'cfg3':
        infer_x = main.resident[0]['input'][:32].contiguous()

        def infer_step(i):
            model.eval()
            with torch.no_grad():
                return model(infer_x)
        infer_repack_ms, _, _, _, _ = timed(infer_step, K, W)        # a bare model(x): weights re-packed every call
        with model.constant_weights(dev):                           # what predict_linear_blend's tile loop runs
            infer_ms, _, _, _, _ = timed(infer_step, K, W)
            _, _, infer_prof, _, _ = timed(infer_step, K, 2, arm=main, profile=True)
        model.train()
        if rank == 0:
            iroof, ibreak, itotal, _ = roofline_of(infer_prof, K, peaks(), math_name, 'inference:')
            extras['inference'] = {
                'workload': 'BASELINE configs[1]: eval-mode forward, 3-ch 256x256, depth 5, batch 32/GPU',
                'value': 32 * world * K / (infer_ms * 1e-3), 'unit': UNIT, 'ms_per_call': infer_ms / K,
                'note': 'inside model.constant_weights() (rd_freeze_params), as in predict_linear_blend; a bare '
                        'model(x) re-packs the weights on every call',
                'ms_per_call_repacking': infer_repack_ms / K,
                'gflop_per_tile': 19.797, 'tflops': 32 * K / (infer_ms * 1e-3) * 19.797 / 1e3,
                'roofline': iroof, 'kernel_ms_per_call': ibreak, 'kernel_ms_total_per_call': itotal / K}
        # tier-next figure: the on-device tile producer that feeds the step
        from resdepth_b200.lib.tiles import DeviceTileProducer
        g = torch.Generator(device=dev).manual_seed(11)
        R = 4096
        prod = DeviceTileProducer.from_device(
            400.0 + 3.0 * torch.randn(R, R, device=dev, generator=g), 400.0 + 3.0 * torch.randn(R, R, device=dev, generator=g),
            100.0 + 30.0 * torch.randn(4, R, R, device=dev, generator=g), -9999.0, T, 'geom-stereo', [[0, 1], [2, 3], [1, 3]],
            None, 3.5, None, 40.0, permute_images_within_pair=True)
        prod_ms, _, _, _, _ = timed(lambda i: prod.sample_batch(B)['input'], K, W)
        del prod
        extras['tile_producer'] = {'workload': 'rd_make_tiles: 64 geom-stereo 256x256 training tiles per call from a '
                                               '4096x4096 raster with 4 views (host-drawn positions / pairs / rot90 / flips)',
                                   'value': B * world * K / (prod_ms * 1e-3), 'unit': UNIT, 'ms_per_call': prod_ms / K}

    # ---- headline pass: the product path (CUDA-graph replay of forward + loss + backward, weight gradients on the
    # side stream, all-reduce slices overlapped), no profiling events
    ms, launches, _, clocks, last_loss = timed(main.device_step, K, W, arm=main, prime=2 * N_INPUT_SETS)
    loss_value = float(last_loss.item())
    identical, checksum = replicas_identical(model._rt['arena'], device=dev)
    # ---- per-kernel pass (roofline, breakdown): eager launches with CUDA events around every kernel category; the side
    # stream is switched off so that each category is timed alone
    tr.use_graphs = False
    handle.set_overlap(False)
    prof_ms, _, prof, _, _ = timed(main.device_step, K, 2, arm=main, profile=True)
    handle.set_overlap(True)
    eager_ms, eager_launches, _, _, _ = timed(main.device_step, K, 2, arm=main)
    tr.use_graphs = True
    # ---- end to end through the public API the reference's train.py drives: Trainer.inference_one_epoch over a loader
    # of K pinned HOST batches (per step: H2D of that step's inputs into a static staging set -- enqueued one batch
    # ahead so it overlaps the previous step's compute --, graph replay, optimizer.step, 4-byte D2H of the loss, read
    # by the host one iteration later).  The un-pipelined per-call figure (inference_one_batch) is kept beside it.
    e2e_call_ms, _, _, _, _ = timed(main.e2e_step, K, 3, arm=main, prime=4)
    e2e_ms, _, _, e2e_clocks, _ = timed(lambda i: main.e2e_epoch(K), 1, 1, arm=main)
    tiles = B * world * K
    value = tiles / (ms * 1e-3)
    e2e_value = tiles / (e2e_ms * 1e-3)
    h2d = main.h2d_bytes()

    # ---- the other backward precision, the other configs, the library baseline (one GPU each, rank-local)
    def guarded(name, fn):
        """Secondary blocks must not cost the headline line: on one GPU a failure is recorded under the block's key.
        With several ranks every rank must take the same path through the collectives, so errors propagate."""
        if world > 1:
            return fn()
        try:
            return fn()
        except Exception as exc:                      # noqa: BLE001
            extras[name] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()

    def tf32_block():
        alt = Arm('cfg3', B, dev, rank, backward_math='tf32')
        alt_ms, _, _, _, _ = timed(alt.device_step, K, W, arm=alt, prime=2 * N_INPUT_SETS)
        extras['value_tf32_bwd'] = {'value': tiles / (alt_ms * 1e-3), 'unit': UNIT, 'ms_per_step': alt_ms / K,
                                    'note': "model.backward_math = 'tf32': TF32 operands in every backward GEMM "
                                            "(what cuDNN's autograd would use), same step otherwise"}
        alt.close()

    def other_config(other):
        okw, oT, oB, ogf, odesc = CONFIGS[other]
        arm = Arm(other, oB, dev, rank, n_sets=2)
        oms, ol, _, _, _ = timed(arm.device_step, K, W, arm=arm, prime=4)
        arm.tr.use_graphs = False
        oms_eager, _, _, _, _ = timed(arm.device_step, K, 2, arm=arm)
        arm.tr.use_graphs = True
        oe2e, _, _, _, _ = timed(lambda i: arm.e2e_epoch(K), 1, 1, arm=arm)
        extras[other] = {'workload': odesc, 'value': oB * world * K / (oms * 1e-3), 'unit': f'{oT}x{oT} tiles/s',
                         'ms_per_step': oms / K, 'step_tflops': oB * K / (oms * 1e-3) * ogf / 1e3,
                         'eager_launch_value': oB * world * K / (oms_eager * 1e-3),
                         'e2e': oB * world * K / (oe2e * 1e-3), 'gpu_launches': ol, 'tiles_per_gpu': oB}
        arm.close()

    def cudnn_block():
        nums = cudnn_numbers(args.config, B, max(5, K // 2), 3, dev)
        best = max(nums['train_fp32']['value'], nums['train_bf16_channels_last']['value'])
        extras['cudnn_baseline'] = nums
        extras['vs_cudnn'] = {'train_vs_fp32_tf32': value / nums['train_fp32']['value'],
                              'train_vs_bf16_channels_last': value / nums['train_bf16_channels_last']['value'],
                              'train_vs_best': value / best}
        if 'inference' in extras:
            ib = max(nums['inference_fp32']['value'], nums['inference_bf16_channels_last']['value'])
            extras['vs_cudnn']['inference_vs_best'] = extras['inference']['value'] / ib

    def blend_block():
        """test.py's inference loop (predict_linear_blend): a 4096 x 4096 raster in 256 x 256 tiles with stride 128 (961
        tiles in batches of 32 from pinned host memory), forward + de-normalise + ramp-weighted accumulation on the
        device, the float64 raster read back once."""
        from types import SimpleNamespace

        from resdepth_b200.lib.evaluation import predict_linear_blend
        from resdepth_b200.lib.UNet import UNet
        R, tile, stride, bs = 4096, 256, 128, 32
        starts = list(range(0, R - tile + 1, stride))
        pos = [(y, x) for y in starts for x in starts]
        box = []
        for (y, x) in pos:                                   # non-overlapping boxes of a regular grid (rasterutils)
            box.append((0 if y == 0 else tile - stride, 0 if x == 0 else tile - stride,
                        tile - 1 if y == starts[-1] else stride - 1, tile - 1 if x == starts[-1] else stride - 1))
        g = torch.Generator().manual_seed(5)
        pool = [torch.randn(bs, 3, tile, tile, generator=g).pin_memory() for _ in range(4)]
        batches = []
        for i in range(0, len(pos), bs):
            sl = slice(i, min(i + bs, len(pos)))
            n = sl.stop - sl.start
            batches.append({'input': pool[(i // bs) % 4][:n], 'dsm_mean': torch.full((n,), 400.0), 'dsm_std': torch.full((n,), 3.5),
                            'patch_offset_y': torch.tensor([p[0] for p in pos[sl]]), 'patch_offset_x': torch.tensor([p[1] for p in pos[sl]]),
                            'patch_valid_pixels_uly': torch.tensor([b[0] for b in box[sl]]),
                            'patch_valid_pixels_ulx': torch.tensor([b[1] for b in box[sl]]),
                            'patch_valid_pixels_lry': torch.tensor([b[2] for b in box[sl]]),
                            'patch_valid_pixels_lrx': torch.tensor([b[3] for b in box[sl]])})

        class Loader(list):
            pass
        loader = Loader(batches)
        loader.dataset = SimpleNamespace(dsm_input_gdal=SimpleNamespace(RasterXSize=R, RasterYSize=R), tile_size=tile, stride=stride)
        torch.manual_seed(0)
        net = UNet(**CONFIGS['cfg3'][0]).to(dev)
        predict_linear_blend(loader, net)                    # warm-up (workspace, plans)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = predict_linear_blend(loader, net)
        dt = time.perf_counter() - t0
        extras['predict_linear_blend'] = {
            'workload': f'{R}x{R} raster, {len(pos)} tiles of {tile}x{tile} (stride {stride}), batches of {bs} from pinned host '
                        'memory; wall clock incl. the H2D copies and the D2H of the float64 raster',
            'value': len(pos) / dt, 'unit': UNIT, 'seconds': dt, 'raster_mean': float(out.mean())}
        net._rt['handle'].close()
        net._rt.clear()

    if args.config == 'cfg3' and not args.quick:
        main.close()
        if world == 1:
            guarded('predict_linear_blend', blend_block)
        guarded('value_tf32_bwd', tf32_block)
        for other in ('cfg5', 'cfg1'):
            guarded(other, lambda o=other: other_config(o))
    if rank == 0 and world == 1 and not args.no_cudnn and not args.quick:
        guarded('cudnn_baseline', cudnn_block)

    if rank == 0:
        pk = peaks()
        roof, breakdown, total_ms, fracs = roofline_of(prof, K, pk, math_name)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': ('f32' if 'tf32' not in math_name else
                      f'tf32 fwd / {bwd_name} bwd (tcgen05 operands; fp32 accumulation, fp32 parameters and activations)'),
            'data': 'synthetic',
            'config': {'workload': workload, 'tiles_per_gpu': B, 'global_batch': B * world, 'tile': T,
                       'parallelism': f'dp{world}',
                       'sharding': 'rank r draws its own tiles (seed 1234 + r): per-rank loaders, Trainer batch partition off',
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
            'graph_priming_steps': 2 * N_INPUT_SETS,
            'host_launches_per_step': 'CUDA-graph replays + Adam (+ all-reduce slices); see eager_launch',
            'eager_launch': {'value': tiles / (eager_ms * 1e-3), 'unit': UNIT, 'kernel_launches': eager_launches,
                             'note': 'RESDEPTH_GRAPHS=0: every kernel launched from the host'},
            'replicas': {'identical_after_steps': identical, 'param_checksum': checksum},
            'roofline': roof,
            'step_tflops': value * gflop_per_tile / 1e3,
            'kernel_ms_per_step': breakdown,
            'kernel_roofline_fracs': fracs,
            'kernel_ms_total_per_step': total_ms / K,
            'kernel_ms_note': f'separate pass of {K} steps with per-category CUDA events, eager launches and the '
                              f'weight-gradient side stream off ({prof_ms / K:.3f} ms per step); the headline pass has none',
            'loss': loss_value,
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tps, sec, done = cpu_train_steps(args.config, B, 3, 1, threads, budget_s=25.0)
            line['cpu_baseline'] = {'value': tps, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                    'sample': f'{done} train steps of {B} tiles after 1 warm-up (oracle port of lib/UNet.py + '
                                              'lib/Trainer.py step + torch.optim.Adam, all host threads); '
                                              f'{sec:.2f} s per {B}-tile step'}
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
    tr._init_runtime()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', choices=['native', 'reference', 'cudnn'], default='native')
    ap.add_argument('--config', choices=list(CONFIGS), default='cfg3')
    ap.add_argument('--batch', type=int, default=0, help='tiles per GPU per step (default: the config\'s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cudnn', action='store_true')
    ap.add_argument('--quick', action='store_true', help='headline workload only (no extra configs / baselines)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.impl == 'cudnn':
        return run_cudnn(args)
    return run_native(args)


if __name__ == '__main__':
    sys.exit(main())
