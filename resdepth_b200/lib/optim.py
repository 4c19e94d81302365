"""Fused optimizers for the hot path: ``optimizer.step()`` of lib/Trainer.py:218.

The reference builds ``torch.optim.Adam(model.parameters(), lr, weight_decay)`` or
``torch.optim.SGD(model.parameters(), lr, weight_decay)`` (lib/utils.py:329-334): default betas/eps, no amsgrad,
no momentum, L2 decay coupled into the gradient.  ``Adam`` / ``SGD`` below subclass the PyTorch classes, so
``param_groups``, ``state_dict()`` / ``load_state_dict()`` (checkpoint format of lib/Trainer.py:145-157),
``__class__.__name__`` (hparams, lib/Trainer.py:70) and LR schedulers behave identically; only ``step()`` is
replaced: one CUDA kernel (``rd_adam_step`` / ``rd_sgd_step``) over the flat parameter / gradient arenas that
``resdepth_b200.lib.UNet`` maintains, or one launch per tensor when the parameters are not arena views.
``fuse_optimizer`` converts an already-built PyTorch optimizer in place (what ``Trainer`` does with
``args.optimizer``)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import _native


def _align4(n: int) -> int:
    return (n + 3) & ~3


def _flat_span(tensors: List[torch.Tensor]) -> Optional[Tuple[int, int]]:
    """If ``tensors`` (float32, contiguous) tile one storage back-to-back with the arena's 4-element alignment,
    returns (data_ptr of the first, total float count incl. padding); else None."""
    if not tensors:
        return None
    base = tensors[0].untyped_storage().data_ptr()
    for t in tensors:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.untyped_storage().data_ptr() != base:
            return None
    order = sorted(tensors, key=lambda t: t.storage_offset())
    first = order[0].storage_offset()
    if (order[0].data_ptr() & 15) != 0:
        return None
    expect = first
    for t in order:
        if t.storage_offset() != expect:
            return None
        expect = _align4(expect + t.numel() - first) + first
    storage_floats = tensors[0].untyped_storage().nbytes() // 4
    if expect > storage_floats:           # the padding of the last tensor is outside the storage
        expect = order[-1].storage_offset() + order[-1].numel()
    return order[0].data_ptr(), expect - first


class _FlatState:
    """Flat float32 state arenas (exp_avg, exp_avg_sq) laid out like the parameters of one param group."""

    def __init__(self):
        self.key = None
        self.arenas = {}

    def ensure(self, params: List[torch.Tensor], names: Tuple[str, ...], state: dict) -> bool:
        """Makes state[p][name] views into flat arenas with the parameters' relative offsets.  Returns True when
        the parameters themselves are one flat span (so a single kernel launch covers everything)."""
        span = _flat_span(params)
        if span is None:
            return False
        base_ptr, total = span
        key = (base_ptr, total, params[0].device)
        if self.key != key:
            self.key = key
            self.arenas = {n: torch.zeros(total, dtype=torch.float32, device=params[0].device) for n in names}
        for p in params:
            off = (p.data_ptr() - base_ptr) // 4
            st = state[p]
            for n in names:
                arena = self.arenas[n]
                cur = st.get(n)
                want_ptr = arena.data_ptr() + 4 * off
                if cur is None or cur.data_ptr() != want_ptr or cur.device != arena.device:
                    view = arena[off:off + p.numel()].view(p.shape)
                    if cur is not None:
                        view.copy_(cur)          # e.g. state restored by load_state_dict
                    st[n] = view
        return True


def _stream_for(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class Adam(torch.optim.Adam):
    """``torch.optim.Adam`` with a fused CUDA ``step`` (coupled L2 decay, bias correction; no amsgrad)."""

    grad_scale = 1.0      # multiplies every gradient inside the kernel (1/world_size after a sum all-reduce)

    def _check_group(self, group):
        if group.get('amsgrad') or group.get('maximize') or group.get('capturable') or group.get('differentiable'):
            raise NotImplementedError('resdepth_b200 Adam: amsgrad / maximize / capturable / differentiable are not '
                                      'supported by the fused CUDA step')
        if isinstance(group['lr'], torch.Tensor):
            raise NotImplementedError('resdepth_b200 Adam: tensor learning rates are not supported')

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not hasattr(self, '_flat'):
            self._flat = {}
        for gi, group in enumerate(self.param_groups):
            self._check_group(group)
            params = [p for p in group['params'] if p.grad is not None]
            if not params:
                continue
            beta1, beta2 = group['betas']
            for p in params:
                if not p.is_cuda:
                    raise RuntimeError('resdepth_b200 Adam: parameters must live on a CUDA device (no CPU fallback)')
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)
                    st['exp_avg'] = None
                    st['exp_avg_sq'] = None
            steps = {int(self.state[p]['step']) for p in params}
            all_have_grad = len(params) == len(group['params'])
            flat = self._flat.setdefault(gi, _FlatState())
            fused = (all_have_grad and len(steps) == 1
                     and flat.ensure(params, ('exp_avg', 'exp_avg_sq'), self.state))
            gspan = _flat_span([p.grad for p in params]) if fused else None
            if fused and gspan is not None:
                pspan = _flat_span(params)
                same_layout = gspan[1] == pspan[1] and all(
                    p.grad.data_ptr() - gspan[0] == p.data_ptr() - pspan[0] for p in params)
            else:
                same_layout = False
            for p in params:
                self.state[p]['step'] += 1
            if fused and same_layout:
                t = steps.pop() + 1
                with torch.cuda.device(params[0].device):
                    _native.adam_step(pspan[0], gspan[0], flat.arenas['exp_avg'].data_ptr(),
                                      flat.arenas['exp_avg_sq'].data_ptr(), pspan[1], float(group['lr']), beta1, beta2,
                                      group['eps'], group['weight_decay'], t, float(self.grad_scale),
                                      _stream_for(params[0]))
                continue
            # per-tensor launches of the same kernel
            for p in params:
                st = self.state[p]
                for n in ('exp_avg', 'exp_avg_sq'):
                    if st.get(n) is None:
                        st[n] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if not p.is_contiguous() or p.dtype != torch.float32:
                    raise RuntimeError('resdepth_b200 Adam: parameters must be contiguous float32')
                with torch.cuda.device(p.device):
                    _native.adam_step(p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(),
                                      st['exp_avg_sq'].data_ptr(), p.numel(), float(group['lr']), beta1, beta2,
                                      group['eps'], group['weight_decay'], int(st['step']), float(self.grad_scale),
                                      _stream_for(p))
        return loss


class SGD(torch.optim.SGD):
    """``torch.optim.SGD`` as the reference builds it (no momentum, coupled L2 decay) with a fused CUDA step."""

    grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group.get('momentum', 0) != 0 or group.get('nesterov') or group.get('maximize') \
                    or group.get('dampening', 0) != 0:
                raise NotImplementedError('resdepth_b200 SGD: momentum / nesterov / maximize are not supported by the '
                                          'fused CUDA step')
            params = [p for p in group['params'] if p.grad is not None]
            if not params:
                continue
            if any(not p.is_cuda for p in params):
                raise RuntimeError('resdepth_b200 SGD: parameters must live on a CUDA device (no CPU fallback)')
            pspan = _flat_span(params) if len(params) == len(group['params']) else None
            gspan = _flat_span([p.grad for p in params]) if pspan is not None else None
            if pspan is not None and gspan is not None and gspan[1] == pspan[1] and all(
                    p.grad.data_ptr() - gspan[0] == p.data_ptr() - pspan[0] for p in params):
                with torch.cuda.device(params[0].device):
                    _native.sgd_step(pspan[0], gspan[0], pspan[1], float(group['lr']), group['weight_decay'],
                                     float(self.grad_scale), _stream_for(params[0]))
                continue
            for p in params:
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                with torch.cuda.device(p.device):
                    _native.sgd_step(p.data_ptr(), g.data_ptr(), p.numel(), float(group['lr']), group['weight_decay'],
                                     float(self.grad_scale), _stream_for(p))
        return loss


def fuse_optimizer(optimizer: torch.optim.Optimizer) -> torch.optim.Optimizer:
    """Switches a PyTorch ``Adam`` / ``SGD`` instance to the fused CUDA ``step`` in place (same object, same
    ``param_groups`` and state, same class name).  Other optimizer types are rejected loudly."""
    if isinstance(optimizer, (Adam, SGD)):
        return optimizer
    target = {torch.optim.Adam: Adam, torch.optim.SGD: SGD}.get(type(optimizer))
    if target is not None:
        # an LR scheduler built earlier has patched ``optimizer.step`` on the INSTANCE with a wrapper around the
        # original class's step (torch.optim.lr_scheduler: patch_track_step_called); re-point it at the fused step
        patched = optimizer.__dict__.pop('step', None)
        optimizer.__class__ = target
        if patched is not None:
            def step(*args, **kwargs):
                optimizer._opt_called = True        # what the scheduler's wrapper records
                return target.step(optimizer, *args, **kwargs)
            step._wrapped_by_lr_sched = True
            optimizer.step = step
        return optimizer
    raise NotImplementedError(f'resdepth_b200: no fused CUDA step for optimizer {type(optimizer).__name__} '
                              "(the reference's get_optimizer builds Adam or SGD, lib/utils.py:329-334)")
