"""Drop-in mirror of the reference ``lib/Trainer.py`` (``Trainer(args)``, reference lib/Trainer.py:13-318).

Same constructor contract (``args`` carries ``trainloader, valloader, model, optimizer, scheduler, criterion,
n_epochs, evaluate_rate, save_model_rate, freq_average_train_loss, save_dir, log_file, checkpoint_dir,
tboard_log_dir, pretrained_path`` -- lib/utils.py:395-439), same public methods (``train``,
``inference_one_epoch``, ``inference_one_batch`` -> ``{'MAE_metric': float}``, ``stats_dict``, ``stats_meter``),
same checkpoint dictionaries and file names, same TensorBoard scalar names and log lines.

What changes is the step itself: forward, masked de-normalised L1 loss, backward and the optimizer update run
as hand-written CUDA kernels through the C ABI (``rd_forward``, ``rd_loss``, ``rd_backward``,
``rd_adam_step``); the per-sample Python loop of ``denormalize_torch`` (lib/data_normalization.py:29-38) and
the ~2B+8 small kernels of ``_compute_denormalized_loss`` (lib/Trainer.py:87-100) collapse into one launch
sequence, and host-to-device copies use pinned, non-blocking transfers.

Data parallel (``torch.distributed`` initialised, one process per GPU): the parameter / BatchNorm arenas are
broadcast from rank 0 at construction, every loader batch is partitioned over the ranks (rank r takes tiles
[r*B/N, (r+1)*B/N)) unless the loader's sampler already shards, the gradient arena is all-reduced
in three slices that overlap the backward pass, and the validation metric is averaged over ranks before it
drives the scheduler / best-model decisions.

Steady-state steps replay CUDA graphs (one launch for forward + loss + backward; three with the all-reduce
slices in between) captured per batch shape once the same device buffers have been seen twice; host batches are
copied into two alternating static staging sets so that the pointers repeat.  ``RESDEPTH_GRAPHS=0`` switches
the capture off.
"""
from __future__ import annotations

import logging
import math
import os
import time

import torch

from .. import _native
from .AverageMeter import AverageMeter
from .distributed import (BucketedAllReduce, allreduce_mean_of_meter, broadcast_state, loader_is_sharded,
                          shard_batch)
from .optim import fuse_optimizer
from .UNet import UNet


class _LevelFormatter(logging.Formatter):
    """INFO records print the bare message, WARNING/ERROR get a level prefix (behaviour of the reference's
    lib/formatter.py + lib/utils.py:640-670)."""

    def format(self, record):
        msg = record.getMessage()
        if record.levelno in (logging.WARNING, logging.ERROR):
            return f'{record.levelname}: {msg}'
        if record.levelno == logging.INFO:
            return msg
        return f'{self.formatTime(record, "%Y-%m-%d %H:%M:%S")} - {record.name} - {record.levelname} - {msg}'


def _setup_logger(name, log_file=None, to_console=True):
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    have = {type(h) for h in logger.handlers}
    if to_console and logging.StreamHandler not in have:
        h = logging.StreamHandler()
        h.setFormatter(_LevelFormatter())
        logger.addHandler(h)
    if log_file and not any(isinstance(h, logging.FileHandler) and
                            os.path.abspath(h.baseFilename) == os.path.abspath(log_file) for h in logger.handlers):
        h = logging.FileHandler(log_file, mode='a')
        h.setFormatter(_LevelFormatter())
        logger.addHandler(h)
    return logger


class _NullWriter:
    def add_scalar(self, *a, **k): pass
    def add_hparams(self, *a, **k): pass
    def close(self): pass


class _StagedBatch(dict):
    """A batch whose tensors already live on the device; ``ready`` = CUDA event recorded after its H2D copies."""
    ready = None


_STEP_KEYS = ('input', 'target', 'loss_mask', 'dsm_mean', 'dsm_std')
_MAX_GRAPHS = 16


class _GraphedStep:
    """CUDA graphs of one step on fixed device buffers: ``segments`` are replayed in order, ``between[i]`` (a flat
    gradient slice or None) is handed to the all-reduce after segment i."""

    def __init__(self):
        self.segments = []
        self.between = []
        self.loss = None
        self.keep = None            # tensors captured by address (inputs, y, dy): kept alive with the graphs


class Trainer(object):
    def __init__(self, args):
        self.config = args

        self.save_dir = args.save_dir
        self.checkpoint_dir = args.checkpoint_dir
        self.tboard_log_dir = args.tboard_log_dir
        self.pretrained_path = args.pretrained_path
        self.log_file = args.log_file

        for d in (self.save_dir, self.checkpoint_dir):
            if d:
                os.makedirs(d, exist_ok=True)
        self.path_model_best = os.path.join(self.checkpoint_dir, 'Model_best.pth')
        self.path_model_last = os.path.join(self.checkpoint_dir, 'Model_last.pth')

        # data-parallel context: one process per GPU (torchrun); rank 0 owns logging and checkpoints
        self.distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank() if self.distributed else 0
        self.world_size = torch.distributed.get_world_size() if self.distributed else 1

        if self.rank == 0 and self.tboard_log_dir:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=self.tboard_log_dir)
        else:
            self.writer = _NullWriter()
        self.logger = _setup_logger('train_logger', log_file=self.log_file if self.rank == 0 else None,
                                    to_console=self.rank == 0)

        self.start_epoch = 0
        self.n_epochs = args.n_epochs
        if not torch.cuda.is_available():
            raise RuntimeError('resdepth_b200: Trainer needs a CUDA device (B200); there is no CPU fallback')
        local = int(os.environ.get('LOCAL_RANK', 0)) if self.distributed else 0
        self.device = torch.device('cuda', local)      # reference: cuda:0 (lib/Trainer.py:34)

        self.model = args.model
        if not isinstance(self.model, UNet):
            raise TypeError('resdepth_b200: Trainer drives resdepth_b200.lib.UNet.UNet models only, got '
                            f'{type(self.model).__name__}')
        self.optimizer = fuse_optimizer(args.optimizer)
        self.scheduler = args.scheduler
        self.criterion = args.criterion
        if self.criterion is not None and not (isinstance(self.criterion, torch.nn.L1Loss)
                                               and self.criterion.reduction == 'mean'):
            raise NotImplementedError("resdepth_b200: the fused loss implements L1Loss(reduction='mean') "
                                      '(the only criterion of the reference, lib/utils.py:284-285)')

        self.evaluate_rate = args.evaluate_rate
        self.save_model_rate = args.save_model_rate
        self.freq_average_train_loss = args.freq_average_train_loss

        self.best_loss = math.inf
        self.index_best_loss = math.inf

        if self.pretrained_path is not None:
            self._load_pretrain(self.pretrained_path)
        else:
            self.logger.info('\nStart training from scratch.\n')
            self.model = self.model.to(self.device)

        self.loader = {'train': args.trainloader, 'val': args.valloader}
        self._init_runtime()

        batch = next(iter(self.loader['train']))
        x, _, _ = self._extract_inputs_outputs_loss_masks(batch)
        self.batch_size = x.shape[0]

        self.hparams = {'batch_size': self.batch_size, 'lr_initial': self._get_lr(),
                        'optimizer': self.optimizer.__class__.__name__}
        if self.scheduler is not None:
            name = self.scheduler.__class__.__name__
            self.hparams['scheduler'] = name
            if name == 'ReduceLROnPlateau':
                self.hparams['patience'] = self.scheduler.patience
                self.hparams['step_size'] = -1
            elif name == 'StepLR':
                self.hparams['patience'] = -1
                self.hparams['step_size'] = self.scheduler.step_size
        else:
            self.hparams.update(scheduler='None', patience=-1, step_size=-1)

    def _init_runtime(self):
        """Device-side state behind the step API: replica synchronisation, the batch-partition policy, staging
        buffers and the CUDA-graph cache (also used by bench.py, which skips the logging part of ``__init__``)."""
        self._copy_stream = None
        self._static = {}            # (set index, shapes) -> static staging tensors of host batches
        self._static_done = {}       # set index -> event: the step that last read this staging set has finished
        self._stage_count = 0
        self.replayed_kernels = 0    # library kernels launched through graph replays (bench: gpu_launches)
        self._graphs = {}            # pointer key -> _GraphedStep
        self._graph_seen = {}
        self.use_graphs = os.environ.get('RESDEPTH_GRAPHS', '1') != '0'
        self._reducer = BucketedAllReduce() if self.distributed else None
        # every rank partitions the loader's batches unless the sampler already hands it its own tiles
        self.shard_batches = {ph: self.distributed and ld is not None and not loader_is_sharded(ld)
                              for ph, ld in self.loader.items()}
        rt = self.model._runtime(self.device)
        if self.distributed:
            broadcast_state([rt['arena'], rt['bufs'], rt['nbt']])
        h = rt['handle']
        self.logger.info(f'resdepth_b200: forward {h.math_mode_name()}; backward GEMM operands {h.bwd_mode_name()}'
                         + (f'; data parallel over {self.world_size} ranks' if self.distributed else ''))

    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def _extract_inputs_outputs_loss_masks(batch):
        return batch['input'], batch['target'], batch['loss_mask']

    def _get_lr(self, group=0):
        return self.optimizer.param_groups[group]['lr']

    def _to_device(self, t, dtype=None):
        if not t.is_cuda and not t.is_pinned():
            t = t.pin_memory()
        return t.to(self.device, dtype=dtype, non_blocking=True)

    def _my_shard(self, batch, phase):
        """The tiles of a loader batch this rank owns; the whole batch on one rank or when the
        loader's sampler already partitions the data."""
        if self.shard_batches.get(phase) and not isinstance(batch, _StagedBatch):
            return shard_batch(batch, self.rank, self.world_size)
        return batch

    def _static_set(self, batch):
        """Two alternating sets of device staging tensors per batch shape: host batches always land at the same
        addresses, which is what lets the step replay a CUDA graph."""
        k = self._stage_count & 1
        self._stage_count += 1
        sig = (k,) + tuple((tuple(batch[n].shape), batch[n].dtype) for n in _STEP_KEYS)
        st = self._static.get(sig)
        if st is None:
            if len(self._static) >= 8:
                self._static.clear()
            want = {'input': torch.float32, 'target': torch.float32, 'loss_mask': batch['loss_mask'].dtype,
                    'dsm_mean': torch.float32, 'dsm_std': torch.float32}
            st = {n: torch.empty(torch.flatten(batch[n]).shape if n.startswith('dsm_') else batch[n].shape,
                                 dtype=want[n], device=self.device) for n in _STEP_KEYS}
            self._static[sig] = st
        return k, st

    def _stage_batch(self, batch, phase='train'):
        """Enqueue the H2D copies of one host batch on the copy stream and return the device-resident batch
        (``_StagedBatch``) with the event that marks the copies complete."""
        batch = self._my_shard(batch, phase)
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        staged = _StagedBatch(batch)
        on_device = any(batch[n].is_cuda for n in _STEP_KEYS)
        if on_device:
            # device-resident batches (e.g. DeviceTileProducer) may still be being written on the compute stream
            self._copy_stream.wait_stream(main)
        use_static = self.use_graphs and not on_device
        if use_static:
            k, st = self._static_set(batch)
            if k in self._static_done:                       # the step that read this set two batches ago
                self._copy_stream.wait_event(self._static_done[k])
            staged.static_set = k
        with torch.cuda.stream(self._copy_stream):
            for n in _STEP_KEYS:
                src = torch.flatten(batch[n]) if n.startswith('dsm_') else batch[n]
                if use_static:
                    # pinned batches (DataLoader(pin_memory=True), the reference's setting) copy asynchronously; pageable
                    # ones go through the driver's staged copy -- pinning 70 MB per batch here would cost more than it saves
                    st[n].copy_(src, non_blocking=True)
                    staged[n] = st[n]
                else:
                    staged[n] = self._to_device(src, None if n == 'loss_mask' else torch.float32)
            staged.ready = torch.cuda.Event()
            staged.ready.record(self._copy_stream)
        if not use_static:
            for n in _STEP_KEYS:
                staged[n].record_stream(main)
        return staged

    def _prefetched(self, loader, phase='train'):
        """Iterate ``loader`` one batch ahead: the H2D copies of batch i+1 are enqueued (copy stream) before the
        step of batch i is launched, so they overlap its compute.  Same batches, same order as the plain loop of
        the reference (lib/Trainer.py:212-213)."""
        it = iter(loader)
        try:
            nxt = self._stage_batch(next(it), phase)
        except StopIteration:
            return
        while nxt is not None:
            cur = nxt
            try:
                nxt = self._stage_batch(next(it), phase)
            except StopIteration:
                nxt = None
            yield cur

    def _compute_denormalized_loss(self, y_pred, y, loss_mask, mean, std, want_grad=False):
        """Fused masked, de-normalised L1 (lib/Trainer.py:87-100).  Returns (loss tensor [1], dy or None)."""
        B, _, T, _ = y_pred.shape
        handle = self.model.native_handle(self.device)
        loss = torch.empty(1, device=self.device, dtype=torch.float32)
        dy = torch.empty_like(y_pred) if want_grad else None
        mask_u8 = self._mask_u8(loss_mask)
        with torch.cuda.device(self.device):
            handle.loss(y_pred.data_ptr(), y.contiguous().data_ptr(), mask_u8.data_ptr(),
                        mean.data_ptr(), std.data_ptr(), loss.data_ptr(), dy.data_ptr() if want_grad else None,
                        B, T, torch.cuda.current_stream().cuda_stream)
        return loss, dy

    def _load_pretrain(self, resume):
        if not os.path.isfile(resume):
            raise ValueError(f"No checkpoint found at '{resume}.\n'")
        # tensors + plain scalars only (what both implementations write); full unpickling of an untrusted file is
        # opt-in: RESDEPTH_TRUSTED_CHECKPOINT=1
        trusted = os.environ.get('RESDEPTH_TRUSTED_CHECKPOINT', '0') == '1'
        checkpoint = torch.load(resume, map_location='cpu', weights_only=not trusted)
        self.model.load_state_dict(checkpoint['model_state_dict'])
        self.model = self.model.to(self.device)            # before the optimizer state, as the reference does
        self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        if 'scheduler_state_dict' in checkpoint and self.scheduler is not None:
            self.scheduler.load_state_dict(checkpoint['scheduler_state_dict'])
        self.start_epoch = checkpoint['epoch'] + 1
        self.n_epochs += self.start_epoch
        self.best_loss = checkpoint['loss_val']
        self.index_best_loss = checkpoint['epoch']
        self.logger.info('\n\nRestoring the pretrained model from epoch {}.'.format(self.start_epoch))
        self.logger.info(f'Successfully load pretrained model from {resume}!\n')
        self.logger.info(f'Current best loss {self.best_loss}\n')

    def _save_checkpoint(self, epoch, loss_train, loss_val, filepath):
        if self.rank != 0:
            return
        # clone(): parameters / optimizer moments are views of flat arenas; saving clones keeps each tensor's file
        # footprint its own.  The optimizer's live state dictionaries are copied, never rebound.
        opt_sd = self.optimizer.state_dict()
        opt_sd = {'state': {pid: {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v)
                                  for k, v in st.items()} for pid, st in opt_sd.get('state', {}).items()},
                  'param_groups': opt_sd['param_groups']}
        state = {
            'epoch': epoch,
            'model_state_dict': {k: v.detach().clone() for k, v in self.model.state_dict().items()},
            'optimizer_state_dict': opt_sd,
            'loss_train': loss_train,
            'loss_val': loss_val,
        }
        if self.scheduler is not None:
            state['scheduler_state_dict'] = self.scheduler.state_dict()
        torch.save(state, filepath)

    # ---------------------------------------------------------------------------------------------
    def inference_one_batch(self, batch, phase):
        return {'MAE_metric': float(self._launch_batch(batch, phase).item())}

    def _launch_batch(self, batch, phase):
        """``inference_one_batch`` without the host synchronisation: enqueues the step and returns the loss as a
        device tensor [1]."""
        assert phase in ['train', 'val']
        train = phase == 'train'
        self.model.train() if train else self.model.eval()

        if isinstance(batch, _StagedBatch):                 # copies already in flight (``_prefetched``)
            main = torch.cuda.current_stream(self.device)
            main.wait_event(batch.ready)
            loss = self.device_step(batch['input'], batch['target'], batch['loss_mask'], batch['dsm_mean'],
                                    batch['dsm_std'], train)
            k = getattr(batch, 'static_set', None)
            if k is not None:                               # the staging set may be overwritten after this point
                ev = self._static_done.get(k)
                if ev is None:
                    ev = self._static_done[k] = torch.cuda.Event()
                ev.record(main)
            return loss

        batch = self._my_shard(batch, phase)
        if all(batch[n].is_cuda for n in _STEP_KEYS):
            # device-resident batch (DeviceTileProducer, a pre-staged loader): nothing to copy
            return self.device_step(batch['input'].float(), batch['target'].float(), batch['loss_mask'],
                                    torch.flatten(batch['dsm_mean']).float(), torch.flatten(batch['dsm_std']).float(),
                                    train)
        if self.use_graphs and not any(batch[n].is_cuda for n in _STEP_KEYS):
            # one code path for host batches: stage (static buffers), then the graphed step
            return self._launch_batch(self._stage_batch(batch, phase), phase)
        x, y, loss_mask = self._extract_inputs_outputs_loss_masks(batch)
        # the input tiles are needed first: copy them on the compute stream; target / mask / normalisation constants
        # are only needed by the loss, so their copies ride a side stream and overlap the forward pass
        x = self._to_device(x, torch.float32)
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            y = self._to_device(y, torch.float32)
            loss_mask = self._to_device(loss_mask)
            mean = self._to_device(torch.flatten(batch['dsm_mean']), torch.float32)
            std = self._to_device(torch.flatten(batch['dsm_std']), torch.float32)
            copied = torch.cuda.Event()
            copied.record(self._copy_stream)
        for t in (y, loss_mask, mean, std):
            t.record_stream(main)

        return self.device_step(x, y, loss_mask, mean, std, train, wait_before_loss=copied)

    # -- the step on device-resident tensors ---------------------------------------------------------------
    def _mask_u8(self, loss_mask):
        if loss_mask.dtype == torch.bool:
            return loss_mask.contiguous().view(torch.uint8)            # zero-copy reinterpretation
        if loss_mask.dtype == torch.uint8:
            return loss_mask.contiguous()
        return (loss_mask != 0).view(torch.uint8)

    def _enqueue_forward_loss(self, x, y, loss_mask, mean, std, train):
        y_pred = self.model._forward_native(x, _native.FWD_TRAIN if train else _native.FWD_EVAL)
        return y_pred, self._compute_denormalized_loss(y_pred, y, loss_mask, mean, std, want_grad=train)

    def _set_grads(self):
        """``param.grad`` = views of the persistent flat gradient arena (what ``loss.backward()`` leaves behind in the
        reference).  The views are built once per arena; per step this is one attribute assignment per parameter."""
        rt = self.model._rt
        flat = rt['grads']
        cache = rt.get('grad_views')
        if cache is None or cache[0] != flat.data_ptr():
            named = dict(self.model.named_parameters())
            cache = rt['grad_views'] = (flat.data_ptr(), [(named[name], flat[off:off + numel].view(named[name].shape))
                                                          for name, numel, off in rt['pinfos']])
        for p, g in cache[1]:
            p.grad = g

    def device_step(self, x, y, loss_mask, mean, std, train, wait_before_loss=None):
        """The step on device-resident tensors: forward, fused loss (+ gradient seed), backward, gradient
        all-reduce; leaves ``param.grad`` set (views of the flat gradient arena) and returns the loss as a
        device tensor [1] without synchronising.  ``inference_one_batch`` = H2D copies + this + ``loss.item()``."""
        with torch.no_grad():
            if self.use_graphs and wait_before_loss is None:
                loss = self._graphed_step(x, y, loss_mask, mean, std, train)
                if loss is not None:
                    return loss
            y_pred = self.model._forward_native(x, _native.FWD_TRAIN if train else _native.FWD_EVAL)
            if wait_before_loss is not None:
                torch.cuda.current_stream(self.device).wait_event(wait_before_loss)
            loss, dy = self._compute_denormalized_loss(y_pred, y, loss_mask, mean, std, want_grad=train)
            if train:
                # the ONE collective of the path: the flat gradient arena summed over ranks (NCCL / NVLink), issued
                # slice by slice as the backward stages complete so that it overlaps the rest of the backward pass
                red = self._reducer
                self.model._backward_native(x, dy, detach_copy=False,
                                            on_stage_done=red.launch if red is not None else None)
                self.optimizer.grad_scale = red.finish() if red is not None else 1.0
                self._set_grads()
        return loss

    def _graphed_step(self, x, y, loss_mask, mean, std, train):
        """Replays (or, the second time the same device buffers are seen, captures) the CUDA graphs of a step on
        these buffers; returns None when the step should run eagerly instead."""
        tensors = (x, y, loss_mask, mean, std)
        handle = self.model._rt['handle']
        if handle.profiling or handle.frozen:             # timing events cannot be captured; a graph must not depend
            return None                                   # on packs that outlive a constant_weights() block
        if not all(t.is_cuda and t.is_contiguous() for t in tensors) or x.dtype != torch.float32 \
                or y.dtype != torch.float32 or mean.dtype != torch.float32 or std.dtype != torch.float32:
            return None
        key = tuple(t.data_ptr() for t in tensors) + (tuple(x.shape), bool(train), self.model._rt.get('math'))
        g = self._graphs.get(key)
        if g is None:
            seen = self._graph_seen.get(key, 0) + 1
            if len(self._graph_seen) > 4096:
                self._graph_seen.clear()
            self._graph_seen[key] = seen
            if seen < 2:
                return None                      # first sight: eager (also reserves the workspace for this shape)
            try:
                g = self._capture_step(x, y, loss_mask, mean, std, train)
            except Exception as exc:                      # capture is an optimisation: fall back to eager launches
                self.use_graphs = False
                self._graphs.clear()
                self.logger.warning(f'resdepth_b200: CUDA-graph capture failed ({exc}); continuing with eager launches')
                torch.cuda.synchronize(self.device)
                return None
            if len(self._graphs) >= _MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        if not self.model._rt['handle'].workspace_alive(g.ws_id):
            # the workspace layout this graph was captured on has been freed: its addresses are stale
            self._graphs.pop(key, None)
            self._graph_seen.pop(key, None)
            return None
        red = self._reducer
        self.replayed_kernels += g.n_kernels
        for seg, grad_slice in zip(g.segments, g.between):
            seg.replay()
            if grad_slice is not None and red is not None:
                red.launch(grad_slice)
        self.model._rt['token'] += 1                 # a replay overwrites the saved activations like a forward call
        if train:
            self.optimizer.grad_scale = red.finish() if red is not None else 1.0
            self._set_grads()
        return g.loss

    def _capture_step(self, x, y, loss_mask, mean, std, train):
        rt = self.model._rt
        handle = rt['handle']
        B, _, T, _ = x.shape
        handle.reserve(B, T, train)
        g = _GraphedStep()
        g.keep = (x, y, loss_mask, mean, std)
        kernels_before = _native.launch_count()
        stream = torch.cuda.Stream(self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        pool = None
        staged = self.distributed and train

        def segment(fn):
            nonlocal pool
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, pool=pool, stream=stream):
                fn()
            if pool is None:
                pool = cg.pool()
            g.segments.append(cg)

        state = {}

        def first():
            state['y_pred'], (state['loss'], state['dy']) = self._enqueue_forward_loss(x, y, loss_mask, mean, std, train)
            if train:
                s = torch.cuda.current_stream(self.device).cuda_stream
                if staged:
                    handle.backward_stage(x.data_ptr(), state['dy'].data_ptr(), 0, s)
                else:
                    handle.backward(x.data_ptr(), state['dy'].data_ptr(), s)

        segment(first)
        if staged:
            off, n = handle.grad_stage_range(0)
            g.between.append(rt['grads'][off:off + n] if n > 0 else None)
            for stage in (1, 2):
                segment(lambda st=stage: handle.backward_stage(x.data_ptr(), state['dy'].data_ptr(), st,
                                                               torch.cuda.current_stream(self.device).cuda_stream))
                off, n = handle.grad_stage_range(stage)
                g.between.append(rt['grads'][off:off + n] if n > 0 else None)
        else:
            g.between.append(None)
        torch.cuda.current_stream(self.device).wait_stream(stream)
        g.loss = state['loss']
        g.keep = g.keep + (state['y_pred'], state['dy'])
        g.ws_id = handle.workspace_id()
        g.n_kernels = _native.launch_count() - kernels_before
        return g

    def inference_one_epoch(self, epoch, phase):
        assert phase in ['train', 'val']
        stats_meter = self.stats_meter()
        num_iter = len(self.loader[phase])

        for param in self.model.parameters():
            param.grad = None

        # Same loop as the reference (lib/Trainer.py:212-238), software-pipelined: batch i+1 is copied to the device
        # while batch i computes (``_prefetched``), and the loss of batch i is read back (4 bytes, pinned memory) only
        # after batch i+1 has been enqueued, so the GPU never waits for the host between steps.  Every batch's
        # MAE_metric reaches the meters / the log exactly as in the reference, one iteration later.
        def consume(pending):
            host, ev, c_iter = pending
            ev.synchronize()
            stats_meter['MAE_metric'].update(float(host.item()))
            if phase == 'train' and (c_iter + 1) % self.freq_average_train_loss == 0:
                curr_iter = num_iter * epoch + (c_iter + 1)
                message = f'{phase}:\tEpoch: {epoch} [{c_iter + 1}/{num_iter}]\t'
                for key, value in stats_meter.items():
                    self.writer.add_scalar(f"train/{key}", value.avg, curr_iter)
                    message += f'{key}: {value.avg:.6f}\t'
                    stats_meter[key].reset()
                self.logger.info(message)
                self.writer.add_scalar("train/learning_rate", self._get_lr(), curr_iter)

        pending = None
        for c_iter, batch in enumerate(self._prefetched(self.loader[phase], phase)):
            loss = self._launch_batch(batch, phase)

            if phase == 'train':
                self.optimizer.step()
                for param in self.model.parameters():
                    param.grad = None

            # pinned landing slots for the 4-byte loss read-back: two alternate (the previous one is consumed only after
            # this step has been enqueued); allocated once -- a pinned allocation per step costs a cudaHostAlloc each
            if getattr(self, '_loss_slots', None) is None:
                self._loss_slots = torch.empty(2, dtype=torch.float32, pin_memory=True)
            host = self._loss_slots[c_iter & 1:(c_iter & 1) + 1]
            host.copy_(loss, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            if pending is not None:
                consume(pending)
            pending = (host, ev, c_iter)
        if pending is not None:
            consume(pending)

        return stats_meter

    @staticmethod
    def stats_dict():
        return {'MAE_metric': 0.}

    def stats_meter(self):
        return {key: AverageMeter() for key in self.stats_dict()}

    # -- the epoch loop of lib/Trainer.py:255-318, split into its decisions ----------------------------------
    def _validate(self, epoch, train_avg):
        """One validation pass plus everything the reference hangs on its result (lib/Trainer.py:271-300): scalar
        logging, best-model checkpoint + hparams, scheduler step.  Under data parallelism the metric is the mean
        over all ranks' tiles, so every replica takes the same decisions."""
        if self.distributed:
            # BatchNorm running statistics are per-rank during training (every rank sees its own tiles, as in plain
            # DDP); before they are USED -- validation, checkpoints -- they become the mean over ranks on every replica
            bufs = self.model._rt['bufs']
            torch.distributed.all_reduce(bufs, op=torch.distributed.ReduceOp.SUM)
            bufs.mul_(1.0 / self.world_size)
        meters = self.inference_one_epoch(epoch, 'val')
        if self.distributed:
            for meter in meters.values():
                total, count = allreduce_mean_of_meter(meter.sum, meter.count, self.device)
                if count > 0:
                    meter.sum, meter.count, meter.avg = total, count, total / count
        line = f"\nval:\tEpoch: {epoch}\t\t"
        for key, meter in meters.items():
            self.writer.add_scalar(f"val/{key}", meter.avg, epoch)
            line += f'{key}: {meter.avg:.6f}\t'
        self.logger.info(line + '\n')
        self.writer.add_scalar("val/learning_rate", self._get_lr(), epoch)

        val_avg = meters['MAE_metric'].avg
        if val_avg < self.best_loss:
            self.best_loss, self.index_best_loss = val_avg, epoch
            self._save_checkpoint(epoch, train_avg, val_avg, self.path_model_best)
            self.writer.add_hparams(hparam_dict=self.hparams, metric_dict={'hparam/MAE_metric': val_avg},
                                    run_name=self.tboard_log_dir)
        if self.scheduler is not None:
            plateau = self.scheduler.__class__.__name__ == 'ReduceLROnPlateau'
            self.scheduler.step(val_avg) if plateau else self.scheduler.step()
        return meters

    def train(self):
        self.logger.info('Start training...\n')
        t_start = time.time()
        epoch = self.start_epoch
        train_meters = val_meters = self.stats_meter()

        for epoch in range(self.start_epoch, self.n_epochs):
            title = f'Epoch {epoch}/{self.n_epochs - 1}'
            self.logger.info('\n{}\n{}\n'.format(title, '-' * len(title)))

            train_meters = self.inference_one_epoch(epoch, 'train')
            if (epoch + 1) % self.evaluate_rate == 0:
                val_meters = self._validate(epoch, train_meters['MAE_metric'].avg)
            if (epoch + 1) % self.save_model_rate == 0 and epoch > self.evaluate_rate:
                self._save_checkpoint(epoch, train_meters['MAE_metric'].avg, val_meters['MAE_metric'].avg,
                                      os.path.join(self.checkpoint_dir, f'Model_after_{epoch + 1}_epochs.pth'))

        elapsed = time.strftime('%H:%M:%S', time.gmtime(time.time() - t_start))
        self.logger.info(f'\n\nTraining finished!\nTraining time: {elapsed}')
        self.logger.info(f'\nBest model at epoch: {self.index_best_loss}')
        self.logger.info('Validation loss of the best model: {:.6f}'.format(self.best_loss))
        self.writer.close()
        self._save_checkpoint(epoch, train_meters['MAE_metric'].avg, val_meters['MAE_metric'].avg, self.path_model_last)
