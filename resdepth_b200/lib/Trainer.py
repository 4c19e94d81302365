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
sequence, and host-to-device copies use pinned, non-blocking transfers.  With ``torch.distributed``
initialised (one process per GPU) the gradient arena is summed with a single NCCL all-reduce before the
optimizer step.
"""
from __future__ import annotations

import logging
import math
import os
import time

import torch

from .. import _native
from .AverageMeter import AverageMeter
from .distributed import allreduce_gradients
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

        self._loss_buf = None

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

    def _stage_batch(self, batch):
        """Enqueue the H2D copies of one host batch on the copy stream and return the device-resident batch
        (``_StagedBatch``) with the event that marks the copies complete."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        staged = _StagedBatch(batch)
        with torch.cuda.stream(self._copy_stream):
            staged['input'] = self._to_device(batch['input'], torch.float32)
            staged['target'] = self._to_device(batch['target'], torch.float32)
            staged['loss_mask'] = self._to_device(batch['loss_mask'])
            staged['dsm_mean'] = self._to_device(torch.flatten(batch['dsm_mean']), torch.float32)
            staged['dsm_std'] = self._to_device(torch.flatten(batch['dsm_std']), torch.float32)
            staged.ready = torch.cuda.Event()
            staged.ready.record(self._copy_stream)
        for k in ('input', 'target', 'loss_mask', 'dsm_mean', 'dsm_std'):
            staged[k].record_stream(main)
        return staged

    def _prefetched(self, loader):
        """Iterate ``loader`` one batch ahead: the H2D copies of batch i+1 are enqueued (copy stream) before the
        step of batch i is launched, so they overlap its compute.  Same batches, same order as the plain loop of
        the reference (lib/Trainer.py:212-213)."""
        it = iter(loader)
        try:
            nxt = self._stage_batch(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur = nxt
            try:
                nxt = self._stage_batch(next(it))
            except StopIteration:
                nxt = None
            yield cur

    def _compute_denormalized_loss(self, y_pred, y, loss_mask, mean, std, want_grad=False):
        """Fused masked, de-normalised L1 (lib/Trainer.py:87-100).  Returns (loss tensor [1], dy or None)."""
        B, _, T, _ = y_pred.shape
        handle = self.model.native_handle(self.device)
        loss = torch.empty(1, device=self.device, dtype=torch.float32)
        dy = torch.empty_like(y_pred) if want_grad else None
        if loss_mask.dtype == torch.bool:
            mask_u8 = loss_mask.contiguous().view(torch.uint8)        # zero-copy reinterpretation
        else:
            mask_u8 = (loss_mask != 0).view(torch.uint8) if loss_mask.dtype != torch.uint8 else loss_mask
        with torch.cuda.device(self.device):
            handle.loss(y_pred.data_ptr(), y.contiguous().data_ptr(), mask_u8.contiguous().data_ptr(),
                        mean.data_ptr(), std.data_ptr(), loss.data_ptr(), dy.data_ptr() if want_grad else None,
                        B, T, torch.cuda.current_stream().cuda_stream)
        return loss, dy

    def _load_pretrain(self, resume):
        if not os.path.isfile(resume):
            raise ValueError(f"No checkpoint found at '{resume}.\n'")
        checkpoint = torch.load(resume, map_location='cpu', weights_only=False)
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
        # clone(): parameters are views of one arena; saving clones keeps each tensor's file footprint its own
        state = {
            'epoch': epoch,
            'model_state_dict': {k: v.detach().clone() for k, v in self.model.state_dict().items()},
            'optimizer_state_dict': self.optimizer.state_dict(),
            'loss_train': loss_train,
            'loss_val': loss_val,
        }
        opt_state = state['optimizer_state_dict'].get('state', {})
        for st in opt_state.values():
            for k, v in list(st.items()):
                if isinstance(v, torch.Tensor):
                    st[k] = v.detach().clone()
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
            torch.cuda.current_stream(self.device).wait_event(batch.ready)
            return self.device_step(batch['input'], batch['target'], batch['loss_mask'], batch['dsm_mean'],
                                    batch['dsm_std'], train)

        x, y, loss_mask = self._extract_inputs_outputs_loss_masks(batch)
        # the input tiles are needed first: copy them on the compute stream; target / mask / normalisation constants
        # are only needed by the loss, so their copies ride a side stream and overlap the forward pass
        x = self._to_device(x, torch.float32)
        main = torch.cuda.current_stream(self.device)
        if getattr(self, '_copy_stream', None) is None:
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

    def device_step(self, x, y, loss_mask, mean, std, train, wait_before_loss=None):
        """The step on device-resident tensors: forward, fused loss (+ gradient seed), backward, gradient
        all-reduce; leaves ``param.grad`` set (views of the flat gradient arena) and returns the loss as a
        device tensor [1] without synchronising.  ``inference_one_batch`` = H2D copies + this + ``loss.item()``."""
        with torch.no_grad():
            y_pred = self.model._forward_native(x, _native.FWD_TRAIN if train else _native.FWD_EVAL)
            if wait_before_loss is not None:
                torch.cuda.current_stream(self.device).wait_event(wait_before_loss)
            loss, dy = self._compute_denormalized_loss(y_pred, y, loss_mask, mean, std, want_grad=train)
            if train:
                grads = self.model._backward_native(x, dy, detach_copy=False)
                # the ONE collective of the path: sum the flat gradient arena over ranks (NCCL / NVLink)
                self.optimizer.grad_scale = allreduce_gradients(self.model._rt['grads'])
                for p, g in zip(self.model.parameters(), grads):
                    p.grad = g
        return loss

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
        for c_iter, batch in enumerate(self._prefetched(self.loader[phase])):
            loss = self._launch_batch(batch, phase)

            if phase == 'train':
                self.optimizer.step()
                for param in self.model.parameters():
                    param.grad = None

            host = torch.empty(1, dtype=torch.float32, pin_memory=True)
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

    def train(self):
        self.logger.info('Start training...\n')
        start_time = time.time()
        epoch = self.start_epoch
        train_stats_meter = val_stats_meter = self.stats_meter()

        for epoch in range(self.start_epoch, self.n_epochs):
            print_msg = f'Epoch {epoch}/{self.n_epochs - 1}'
            self.logger.info('\n{}\n{}\n'.format(print_msg, '-' * len(print_msg)))

            train_stats_meter = self.inference_one_epoch(epoch, 'train')

            if (epoch + 1) % self.evaluate_rate == 0:
                val_stats_meter = self.inference_one_epoch(epoch, 'val')

                message = f"\nval:\tEpoch: {epoch}\t\t"
                for key, value in val_stats_meter.items():
                    self.writer.add_scalar(f"val/{key}", value.avg, epoch)
                    message += f'{key}: {value.avg:.6f}\t'
                self.logger.info(message + '\n')
                self.writer.add_scalar("val/learning_rate", self._get_lr(), epoch)

                if val_stats_meter['MAE_metric'].avg < self.best_loss:
                    self.best_loss = val_stats_meter['MAE_metric'].avg
                    self.index_best_loss = epoch
                    self._save_checkpoint(epoch, train_stats_meter['MAE_metric'].avg,
                                          val_stats_meter['MAE_metric'].avg, self.path_model_best)
                    self.writer.add_hparams(hparam_dict=self.hparams,
                                            metric_dict={'hparam/MAE_metric': val_stats_meter['MAE_metric'].avg},
                                            run_name=self.tboard_log_dir)

                if self.scheduler is not None:
                    if self.scheduler.__class__.__name__ == 'ReduceLROnPlateau':
                        self.scheduler.step(val_stats_meter['MAE_metric'].avg)
                    else:
                        self.scheduler.step()

            if (epoch + 1) % self.save_model_rate == 0 and epoch > self.evaluate_rate:
                name = 'Model_after_' + str(epoch + 1) + '_epochs.pth'
                self._save_checkpoint(epoch, train_stats_meter['MAE_metric'].avg, val_stats_meter['MAE_metric'].avg,
                                      os.path.join(self.checkpoint_dir, name))

        time_text = time.strftime('%H:%M:%S', time.gmtime(time.time() - start_time))
        self.logger.info(f'\n\nTraining finished!\nTraining time: {time_text}')
        self.logger.info(f'\nBest model at epoch: {self.index_best_loss}')
        self.logger.info('Validation loss of the best model: {:.6f}'.format(self.best_loss))
        self.writer.close()

        self._save_checkpoint(epoch, train_stats_meter['MAE_metric'].avg, val_stats_meter['MAE_metric'].avg,
                              self.path_model_last)
