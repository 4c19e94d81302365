"""Drop-in mirror of the reference ``lib/UNet.py`` (``UNet(nn.Module)``, reference lib/UNet.py:104-246).

Same constructor arguments, same sub-module tree (hence the same ``state_dict`` keys, shapes and -- seed for
seed -- the same initial values, including the RNG draws of the unused up-sampling branch the reference
builds at lib/UNet.py:19-22), same ``ValueError`` behaviour.  The sub-modules are *parameter containers* only:
``forward`` does not call them.  It hands the flat parameter arena to the CUDA library
(``include/resdepth_b200.h``: ``rd_forward`` / ``rd_backward``) which runs the whole encoder-decoder as
hand-written sm_100a kernels.  There is no PyTorch/CPU fallback: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _native

_ACTIVATIONS = ('relu', 'lrelu', 'prelu')
_UP_MODES = ('transpose', 'bilinear')


def _make_activation(kind: str) -> nn.Module:
    # lib/UNet.py:27-33
    if kind == 'relu':
        return nn.ReLU(inplace=True)
    if kind == 'lrelu':
        return nn.LeakyReLU(inplace=True)
    return nn.PReLU()


def _make_conv_block(c_in: int, c_out: int, activation: str, do_bn: bool) -> nn.Sequential:
    # conv_block / bottleneck / inner block of conv_up_block: lib/UNet.py:36-52,64-75,78-93
    conv = nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1, bias=not do_bn)
    if do_bn:
        return nn.Sequential(conv, nn.BatchNorm2d(c_out), _make_activation(activation))
    return nn.Sequential(conv, _make_activation(activation))


def _make_upconv(channels_in: int, channels_out: int, mode: str) -> nn.Module:
    # lib/UNet.py:17-24 builds BOTH variants (1x1 conv first, then the transposed conv) and returns one of
    # them; constructing both in that order keeps the global RNG stream identical to the reference's.
    pointwise = nn.Conv2d(channels_in, channels_out, kernel_size=1, stride=1)
    transposed = nn.ConvTranspose2d(channels_in, channels_out, kernel_size=2, stride=2)
    if mode == 'transpose':
        return transposed
    return nn.Sequential(nn.Upsample(mode='bilinear', scale_factor=2), pointwise)


class SkipConnection(nn.Module):
    """Additive skip (lib/UNet.py:96-101); kept in the module tree for ``print(model)`` parity."""

    def forward(self, x_skip, x_up):
        return x_skip + x_up


class _NativeUNetFn(torch.autograd.Function):
    """Connects the CUDA forward/backward to autograd so ``loss.backward()`` (lib/Trainer.py:179) fills
    ``param.grad`` of every parameter."""

    @staticmethod
    def forward(ctx, x, model, *params):
        mode = _native.FWD_TRAIN if model.training else _native.FWD_EVAL_SAVE
        y = model._forward_native(x, mode)
        ctx.model = model
        ctx.token = model._rt['token']
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        model = ctx.model
        (x,) = ctx.saved_tensors
        if model._rt.get('token') != ctx.token:
            raise RuntimeError('resdepth_b200: the activations of this forward pass were overwritten by a later '
                               'forward call of the same model (one in-flight step per model)')
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('resdepth_b200: gradients with respect to the input tiles are not computed')
        grads = model._backward_native(x, dy, detach_copy=True)
        return (None, None) + tuple(grads)


class UNet(nn.Module):
    def __init__(self, n_input_channels=1, start_kernel=64, max_filter_depth=512, depth=8,
                 act_fn_encoder='relu', act_fn_decoder='relu', act_fn_bottleneck='relu', up_mode='transpose',
                 do_BN=True, bias_conv_layer=False, outer_skip=True, outer_skip_BN=False):
        super().__init__()
        for choice in (act_fn_encoder, act_fn_decoder, act_fn_bottleneck):
            if choice not in _ACTIVATIONS:
                raise ValueError(f"'{choice}' is not a valid activation function. "
                                 f"Choose among {list(_ACTIVATIONS)}.\n")
        if up_mode not in _UP_MODES:
            raise ValueError(f"'{up_mode}' is not a valid mode for upsampling. Choose among {list(_UP_MODES)} "
                             "to specify 'up_mode'.\n")

        self.n_input_channels = n_input_channels
        self.start_kernel = start_kernel
        self.depth = depth
        self.act_fn_encoder = act_fn_encoder
        self.act_fn_decoder = act_fn_decoder
        self.act_fn_bottleneck = act_fn_bottleneck
        self.up_mode = up_mode
        self.max_filter_depth = max_filter_depth
        self.do_BN = do_BN
        self.bias_conv_layer = bias_conv_layer
        self.do_outer_skip = outer_skip
        self.do_outer_skip_BN = outer_skip_BN
        widths = [min(start_kernel * 2 ** i, max_filter_depth) for i in range(depth)]   # lib/UNet.py:152-155
        self.filter_depths = widths
        self.filter_depths_up = widths[::-1]

        # parameter containers, registered in the reference's order (lib/UNet.py:157-194)
        self.encoder = nn.ModuleList()
        for c_in, c_out in zip([n_input_channels] + widths[:-1], widths):
            self.encoder.append(nn.Sequential(_make_conv_block(c_in, c_out, act_fn_encoder, do_BN),
                                              nn.MaxPool2d(kernel_size=2, stride=2)))
        self.bottleneck = _make_conv_block(widths[-1], widths[-1], act_fn_bottleneck, do_BN)
        self.decoder = nn.ModuleList()
        ups = self.filter_depths_up
        for c_in, c_out in zip(ups[:-1], ups[1:]):
            self.decoder.append(nn.Sequential(_make_upconv(c_in, c_in, up_mode),
                                              _make_conv_block(c_in, c_out, act_fn_decoder, do_BN)))
        self.decoder.append(_make_upconv(ups[-1], ups[-1], up_mode))
        self.last_layer = nn.Conv2d(start_kernel, 1, kernel_size=3, stride=1, padding=1, bias=bias_conv_layer)
        self.skipconnect = SkipConnection()
        if outer_skip:
            self.layer_outer_skip = nn.ModuleList()
            if outer_skip_BN:
                self.layer_outer_skip.append(nn.BatchNorm2d(1))
            self.layer_outer_skip.append(SkipConnection())

        object.__setattr__(self, '_rt', {})          # runtime state (native handle, arenas); never pickled/copied

    # -- pickling / deepcopy: drop the native runtime state, it is rebuilt lazily -----------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop('_rt', None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        object.__setattr__(self, '_rt', {})

    # -- native plumbing ------------------------------------------------------------------------------
    @staticmethod
    def _math_mode() -> int:
        m = os.environ.get('RESDEPTH_MATH', 'tf32').lower()
        if m not in ('tf32', 'fp32'):
            raise ValueError(f"RESDEPTH_MATH must be 'tf32' or 'fp32', got '{m}'")
        return _native.MATH_TF32 if m == 'tf32' else _native.MATH_FP32

    # Operand type of the backward GEMMs on the tcgen05 path.  The reference's autograd runs them in fp32 (TF32 under
    # cuDNN); this implementation defaults to bf16 operands with fp32 accumulation ('auto' = bf16 unless the
    # environment says RESDEPTH_BWD=tf32).  Set ``model.backward_math = 'tf32'`` before the first forward call for
    # TF32 operands; gradient error of either mode against the fp32 oracle: DESIGN.md section 5.
    backward_math = 'auto'

    def _bwd_mode(self) -> int:
        if self.backward_math not in _native.BWD_IDS:
            raise ValueError(f"backward_math must be one of {list(_native.BWD_IDS)}, got '{self.backward_math}'")
        mode = _native.BWD_IDS[self.backward_math]
        if mode == 0 and os.environ.get('RESDEPTH_BWD', '').lower() == 'tf32':
            mode = _native.BWD_IDS['tf32']
        return mode

    def _config(self) -> _native.RdConfig:
        return _native.RdConfig(
            n_input_channels=self.n_input_channels, start_kernel=self.start_kernel,
            max_filter_depth=self.max_filter_depth, depth=self.depth,
            act_encoder=_native.ACT_IDS[self.act_fn_encoder], act_decoder=_native.ACT_IDS[self.act_fn_decoder],
            act_bottleneck=_native.ACT_IDS[self.act_fn_bottleneck], do_bn=int(bool(self.do_BN)),
            bias_conv_layer=int(bool(self.bias_conv_layer)), outer_skip=int(bool(self.do_outer_skip)),
            outer_skip_bn=int(bool(self.do_outer_skip_BN)), math_mode=self._math_mode(),
            up_mode=_native.UP_IDS[self.up_mode], bwd_mode=self._bwd_mode())

    def _runtime(self, device: torch.device) -> dict:
        """Creates (once per device) the native handle and moves parameters/buffers into flat arenas."""
        rt = self._rt
        math_key = (self._math_mode(), self._bwd_mode())
        if rt.get('device') != device or rt.get('math') != math_key:
            if 'handle' in rt:
                rt['handle'].close()
            rt.clear()
            handle = _native.Handle(self._config(), device.index if device.index is not None else 0)
            named = dict(self.named_parameters())
            infos = handle.param_infos()
            if [n for n, _, _ in infos] != list(named.keys()):
                raise RuntimeError('resdepth_b200: native parameter plan does not match the module tree:\n'
                                   f'{[n for n, _, _ in infos]}\nvs\n{list(named.keys())}')
            for name, numel, _ in infos:
                if named[name].numel() != numel:
                    raise RuntimeError(f'resdepth_b200: parameter {name} has {named[name].numel()} elements, '
                                       f'native plan expects {numel}')
            rt.update(device=device, math=math_key, handle=handle, pinfos=infos,
                      binfos=handle.buffer_infos(), token=0,
                      arena=torch.zeros(max(handle.param_arena_size(), 4), device=device),
                      grads=torch.zeros(max(handle.param_arena_size(), 4), device=device),
                      bufs=torch.zeros(max(handle.buffer_arena_size(), 4), device=device),
                      nbt=None, bound=None)
        self._adopt_tensors(rt)
        return rt

    def _adopt_tensors(self, rt: dict):
        """Makes every parameter / BatchNorm buffer a view into the flat arenas (copying current values in when
        it is not one already, e.g. after ``.to(device)``); ``load_state_dict`` copies in place and keeps them."""
        arena, bufs, device = rt['arena'], rt['bufs'], rt['device']
        named = dict(self.named_parameters())
        base = arena.data_ptr()
        for name, numel, off in rt['pinfos']:
            p = named[name]
            if p.dtype != torch.float32:
                raise TypeError(f'resdepth_b200: parameter {name} is {p.dtype}; the CUDA path is float32 only')
            if p.device != device or p.data_ptr() != base + 4 * off or not p.is_contiguous():
                view = arena[off:off + numel].view(p.shape)
                view.copy_(p.data)
                p.data = view
                if rt['handle'].frozen:                       # constant_weights(): the packed copies are stale now
                    rt['handle'].freeze_params(True)
        named_b = dict(self.named_buffers())
        bbase = bufs.data_ptr()
        for name, numel, off in rt['binfos']:
            b = named_b[name]
            if b.device != device or b.data_ptr() != bbase + 4 * off:
                view = bufs[off:off + numel].view(b.shape)
                view.copy_(b.data)
                mod_name, _, leaf = name.rpartition('.')
                self.get_submodule(mod_name)._buffers[leaf] = view
        if self.do_BN:
            # num_batches_tracked counters: one int64 arena so a training step bumps them with one add_
            counters = [(n, b) for n, b in self.named_buffers() if n.endswith('num_batches_tracked')]
            nbt = rt['nbt']
            if nbt is None or nbt.numel() != len(counters):
                nbt = rt['nbt'] = torch.zeros(len(counters), dtype=torch.int64, device=device)
            for i, (name, b) in enumerate(counters):
                if b.device != device or b.data_ptr() != nbt.data_ptr() + 8 * i:
                    nbt[i] = b.to(device)
                    mod_name, _, leaf = name.rpartition('.')
                    self.get_submodule(mod_name)._buffers[leaf] = nbt[i]
        key = (arena.data_ptr(), rt['grads'].data_ptr(), bufs.data_ptr())
        if rt['bound'] != key:
            rt['handle'].bind(*key)
            rt['bound'] = key

    def _check_input(self, x: torch.Tensor):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError('resdepth_b200: expected a [B, C, T, T] tensor')
        if not x.is_cuda:
            raise RuntimeError('resdepth_b200: the UNet runs on CUDA tensors only (there is no CPU fallback); '
                               'move the model and its inputs to a B200 device')
        if x.dtype != torch.float32:
            raise TypeError(f'resdepth_b200: input must be float32, got {x.dtype}')
        B, C, H, W = x.shape
        if C != self.n_input_channels:
            raise ValueError(f'resdepth_b200: input has {C} channels, the model expects {self.n_input_channels}')
        if H != W:
            raise ValueError(f'resdepth_b200: tiles must be square, got {H}x{W}')

    def _forward_native(self, x: torch.Tensor, mode: int) -> torch.Tensor:
        self._check_input(x)
        x = x.contiguous()
        rt = self._runtime(x.device)
        B, _, T, _ = x.shape
        y = torch.empty((B, 1, T, T), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            rt['handle'].forward(x.data_ptr(), y.data_ptr(), B, T, mode, stream)
            if mode == _native.FWD_TRAIN and rt['nbt'] is not None:
                rt['nbt'].add_(1)
        rt['token'] += 1
        return y

    def _backward_native(self, x: torch.Tensor, dy: torch.Tensor, detach_copy: bool,
                         on_stage_done=None) -> List[torch.Tensor]:
        """Runs rd_backward; returns per-parameter gradient views (into the persistent gradient arena, or into a
        fresh copy of it when ``detach_copy``).  With ``on_stage_done`` the pass runs in its three stages
        (rd_backward_stage) and ``on_stage_done(flat_gradient_slice)`` is called after each one has been enqueued:
        the hook of the data-parallel trainer, which starts that slice's all-reduce while the next stage computes."""
        rt = self._rt
        dy = dy.contiguous()
        x = x.contiguous()
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            if on_stage_done is None:
                rt['handle'].backward(x.data_ptr(), dy.data_ptr(), stream)
            else:
                for stage in range(3):
                    rt['handle'].backward_stage(x.data_ptr(), dy.data_ptr(), stage, stream)
                    off, n = rt['handle'].grad_stage_range(stage)
                    if n > 0:
                        on_stage_done(rt['grads'][off:off + n])
        flat = rt['grads'].clone() if detach_copy else rt['grads']
        named = dict(self.named_parameters())
        return [flat[off:off + numel].view(named[name].shape) for name, numel, off in rt['pinfos']]

    def native_handle(self, device: Optional[torch.device] = None):
        """The ``rd_handle`` wrapper of this model on ``device`` (default: the parameters' device)."""
        if device is None:
            device = next(self.parameters()).device
        return self._runtime(torch.device(device))['handle']

    @contextlib.contextmanager
    def constant_weights(self, device: Optional[torch.device] = None):
        """Inference with weights that do not change inside the block (the tile loop of ``test.py``): eval-mode
        forwards re-use the packed GEMM copies of the weights and the BatchNorm scale / shift vectors instead of
        rebuilding them on every call (``rd_freeze_params``).  The caller promises not to modify parameters or
        buffers inside the block; ``load_state_dict`` / optimizer steps belong outside it."""
        handle = self.native_handle(device)
        handle.freeze_params(True)
        try:
            yield self
        finally:
            if handle._h:                                   # the handle may have been replaced (math-mode change)
                handle.freeze_params(False)

    # -- public API -------------------------------------------------------------------------------
    def forward(self, x):
        self._check_input(x)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            self._runtime(x.device)                       # parameters become arena views before autograd sees them
            return _NativeUNetFn.apply(x, self, *self.parameters())
        return self._forward_native(x, _native.FWD_TRAIN if self.training else _native.FWD_EVAL)
