"""ctypes binding of ``libresdepth_b200.so`` (the C ABI in ``include/resdepth_b200.h``).

There is no fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_LIB = None
_LOCK = threading.Lock()

ACT_IDS = {'relu': 0, 'lrelu': 1, 'prelu': 2}          # RD_ACT_*
MATH_FP32, MATH_TF32 = 0, 1                             # RD_MATH_*
FWD_EVAL, FWD_TRAIN, FWD_EVAL_SAVE = 0, 1, 2            # RD_FWD_*
UP_IDS = {'transpose': 0, 'bilinear': 1}                # RD_UP_*
ABI_VERSION = 4
BWD_IDS = {'auto': 0, 'tf32': 1, 'bf16': 2}                  # RD_BWD_*

# every symbol include/resdepth_b200.h declares
EXPORTED_SYMBOLS = (
    'rd_abi_version', 'rd_last_error', 'rd_create', 'rd_destroy', 'rd_num_params', 'rd_param_info',
    'rd_param_arena_size', 'rd_num_buffers', 'rd_buffer_info', 'rd_buffer_arena_size', 'rd_bind', 'rd_reserve',
    'rd_workspace_bytes', 'rd_forward', 'rd_loss', 'rd_backward', 'rd_adam_step', 'rd_sgd_step',
    'rd_blend_accumulate', 'rd_launch_count', 'rd_math_mode_name', 'rd_profile_enable', 'rd_profile_collect',
    'rd_profile_read', 'rd_debug_rows', 'rd_debug_reduce', 'rd_make_tiles', 'rd_residuals', 'rd_residual_stats',
    'rd_tile_stds', 'rd_set_overlap', 'rd_freeze_params', 'rd_backward_stage', 'rd_grad_stage_range', 'rd_bwd_mode_name',
    'rd_workspace_id', 'rd_workspace_alive',
)
PROF_NUM = 18                                           # RD_PROF_NUM


class RdConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'n_input_channels', 'start_kernel', 'max_filter_depth', 'depth', 'act_encoder', 'act_decoder',
        'act_bottleneck', 'do_bn', 'bias_conv_layer', 'outer_skip', 'outer_skip_bn', 'math_mode', 'up_mode',
        'bwd_mode')]


def library_path() -> str:
    return os.environ.get('RESDEPTH_B200_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '_lib',
                                                                'libresdepth_b200.so')


def _declare(lib):
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    lib.rd_abi_version.restype = i32
    lib.rd_abi_version.argtypes = []
    lib.rd_last_error.restype = C.c_char_p
    lib.rd_last_error.argtypes = []
    lib.rd_create.restype = i32
    lib.rd_create.argtypes = [C.POINTER(RdConfig), i32, C.POINTER(vp)]
    lib.rd_destroy.restype = i32
    lib.rd_destroy.argtypes = [vp]
    lib.rd_num_params.restype = i32
    lib.rd_num_params.argtypes = [vp]
    lib.rd_param_info.restype = i32
    lib.rd_param_info.argtypes = [vp, i32, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]
    lib.rd_param_arena_size.restype = i64
    lib.rd_param_arena_size.argtypes = [vp]
    lib.rd_num_buffers.restype = i32
    lib.rd_num_buffers.argtypes = [vp]
    lib.rd_buffer_info.restype = i32
    lib.rd_buffer_info.argtypes = [vp, i32, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]
    lib.rd_buffer_arena_size.restype = i64
    lib.rd_buffer_arena_size.argtypes = [vp]
    lib.rd_bind.restype = i32
    lib.rd_bind.argtypes = [vp, vp, vp, vp]
    lib.rd_reserve.restype = i32
    lib.rd_reserve.argtypes = [vp, i32, i32, i32]
    lib.rd_workspace_bytes.restype = i64
    lib.rd_workspace_bytes.argtypes = [vp]
    lib.rd_workspace_id.restype = i64
    lib.rd_workspace_id.argtypes = [vp]
    lib.rd_workspace_alive.restype = i32
    lib.rd_workspace_alive.argtypes = [vp, i64]
    lib.rd_forward.restype = i32
    lib.rd_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.rd_loss.restype = i32
    lib.rd_loss.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.rd_backward.restype = i32
    lib.rd_backward.argtypes = [vp, vp, vp, vp]
    lib.rd_backward_stage.restype = i32
    lib.rd_backward_stage.argtypes = [vp, vp, vp, i32, vp]
    lib.rd_grad_stage_range.restype = i32
    lib.rd_grad_stage_range.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.rd_bwd_mode_name.restype = C.c_char_p
    lib.rd_bwd_mode_name.argtypes = [vp]
    lib.rd_adam_step.restype = i32
    lib.rd_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, f32, vp]
    lib.rd_sgd_step.restype = i32
    lib.rd_sgd_step.argtypes = [vp, vp, i64, f32, f32, f32, vp]
    lib.rd_blend_accumulate.restype = i32
    lib.rd_blend_accumulate.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, i32, i32, vp]
    lib.rd_profile_enable.restype = i32
    lib.rd_profile_enable.argtypes = [vp, i32]
    lib.rd_profile_collect.restype = i32
    lib.rd_profile_collect.argtypes = [vp]
    lib.rd_profile_read.restype = i32
    lib.rd_profile_read.argtypes = [vp, i32, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64)]
    lib.rd_debug_rows.restype = i32
    lib.rd_debug_rows.argtypes = [i32, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]
    lib.rd_debug_reduce.restype = i32
    lib.rd_debug_reduce.argtypes = [i32, i32, vp, i32, i32, i32, i32, vp, i32, vp, vp, i64, vp]
    lib.rd_make_tiles.restype = i32
    lib.rd_make_tiles.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, f32, f32, f32, f32, f32,
                                  vp, vp, vp, vp, vp, vp]
    lib.rd_residuals.restype = i32
    lib.rd_residuals.argtypes = [vp, i32, vp, i32, vp, i64, f64, vp, vp, vp]
    lib.rd_residual_stats.restype = i32
    lib.rd_residual_stats.argtypes = [vp, vp, i64, f64, vp, vp]
    lib.rd_tile_stds.restype = i32
    lib.rd_tile_stds.argtypes = [vp, i32, i32, vp, i32, i32, f32, vp, vp]
    lib.rd_set_overlap.restype = i32
    lib.rd_set_overlap.argtypes = [vp, i32]
    lib.rd_freeze_params.restype = i32
    lib.rd_freeze_params.argtypes = [vp, i32]
    lib.rd_launch_count.restype = i64
    lib.rd_launch_count.argtypes = [i32]
    lib.rd_math_mode_name.restype = C.c_char_p
    lib.rd_math_mode_name.argtypes = [vp]


def lib():
    """Loads the shared library once; raises if it is absent (there is no CPU/PyTorch fallback)."""
    global _LIB
    if _LIB is None:
        with _LOCK:
            if _LIB is None:
                path = library_path()
                if not os.path.isfile(path):
                    raise RuntimeError(
                        f'resdepth_b200: CUDA library not built ({path} missing). Run `python -m resdepth_b200._build` '
                        '(needs nvcc); there is no fallback path.')
                handle = C.CDLL(path, mode=C.RTLD_GLOBAL)
                _declare(handle)
                v = handle.rd_abi_version()
                if v != ABI_VERSION:
                    raise RuntimeError(f'resdepth_b200: ABI version mismatch (library {v}, binding {ABI_VERSION})')
                _LIB = handle
    return _LIB


def check(status: int, what: str = ''):
    if status != 0:
        msg = lib().rd_last_error()
        raise RuntimeError(f'resdepth_b200 {what} failed: {msg.decode() if msg else "unknown error"}')


def launch_count(reset: bool = False) -> int:
    return int(lib().rd_launch_count(1 if reset else 0))


class Handle:
    """Owns one ``rd_handle`` (a layer plan + workspace on one device)."""

    def __init__(self, cfg: RdConfig, device_index: int):
        self._h = C.c_void_p()
        self._lib = lib()
        check(self._lib.rd_create(C.byref(cfg), int(device_index), C.byref(self._h)), 'rd_create')
        self.device_index = int(device_index)
        self.profiling = False
        self.frozen = False

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self._lib.rd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _infos(self, count_fn, info_fn):
        out = []
        name = C.create_string_buffer(64)
        numel, off = C.c_int64(), C.c_int64()
        for i in range(count_fn(self._h)):
            check(info_fn(self._h, i, name, C.byref(numel), C.byref(off)), 'info')
            out.append((name.value.decode(), int(numel.value), int(off.value)))
        return out

    def param_infos(self):
        return self._infos(self._lib.rd_num_params, self._lib.rd_param_info)

    def buffer_infos(self):
        return self._infos(self._lib.rd_num_buffers, self._lib.rd_buffer_info)

    def param_arena_size(self) -> int:
        return int(self._lib.rd_param_arena_size(self._h))

    def buffer_arena_size(self) -> int:
        return int(self._lib.rd_buffer_arena_size(self._h))

    def bind(self, params_ptr: int, grads_ptr: int, buffers_ptr: int):
        check(self._lib.rd_bind(self._h, params_ptr, grads_ptr, buffers_ptr), 'rd_bind')

    def reserve(self, batch: int, tile: int, with_backward: bool):
        check(self._lib.rd_reserve(self._h, batch, tile, 1 if with_backward else 0), 'rd_reserve')

    def workspace_bytes(self) -> int:
        return int(self._lib.rd_workspace_bytes(self._h))

    def workspace_id(self) -> int:
        return int(self._lib.rd_workspace_id(self._h))

    def workspace_alive(self, ws_id: int) -> bool:
        return bool(self._lib.rd_workspace_alive(self._h, int(ws_id)))

    def forward(self, x_ptr, y_ptr, batch, tile, mode, stream):
        check(self._lib.rd_forward(self._h, x_ptr, y_ptr, batch, tile, mode, stream), 'rd_forward')

    def loss(self, y_pred, target, mask, mean, std, loss_out, dy_out, batch, tile, stream):
        check(self._lib.rd_loss(self._h, y_pred, target, mask, mean, std, loss_out, dy_out, batch, tile, stream),
              'rd_loss')

    def backward(self, x_ptr, dy_ptr, stream):
        check(self._lib.rd_backward(self._h, x_ptr, dy_ptr, stream), 'rd_backward')

    def backward_stage(self, x_ptr, dy_ptr, stage, stream):
        check(self._lib.rd_backward_stage(self._h, x_ptr, dy_ptr, stage, stream), 'rd_backward_stage')

    def grad_stage_range(self, stage: int):
        """(offset, numel) in floats of the gradient-arena slice that backward stage ``stage`` completes."""
        off, n = C.c_int64(), C.c_int64()
        check(self._lib.rd_grad_stage_range(self._h, stage, C.byref(off), C.byref(n)), 'rd_grad_stage_range')
        return int(off.value), int(n.value)

    def bwd_mode_name(self) -> str:
        return self._lib.rd_bwd_mode_name(self._h).decode()

    def set_overlap(self, on: bool):
        """Side-stream weight gradients in rd_backward on (default) / off (everything on the caller's stream)."""
        check(self._lib.rd_set_overlap(self._h, 1 if on else 0), 'rd_set_overlap')

    def freeze_params(self, on: bool):
        """Constant-weight inference: while on, eval-mode forwards re-use the packed weights (rd_freeze_params)."""
        check(self._lib.rd_freeze_params(self._h, 1 if on else 0), 'rd_freeze_params')
        self.frozen = bool(on)

    def profile_enable(self, on: bool):
        check(self._lib.rd_profile_enable(self._h, 1 if on else 0), 'rd_profile_enable')
        self.profiling = bool(on)

    def profile_read(self):
        """Collects pending events; returns {category: dict(ms, flops, bytes, launches, calls)}."""
        check(self._lib.rd_profile_collect(self._h), 'rd_profile_collect')
        out = {}
        name = C.create_string_buffer(64)
        ms, fl, by = C.c_double(), C.c_double(), C.c_double()
        la, ca = C.c_int64(), C.c_int64()
        for i in range(PROF_NUM):
            check(self._lib.rd_profile_read(self._h, i, name, C.byref(ms), C.byref(fl), C.byref(by), C.byref(la),
                                            C.byref(ca)), 'rd_profile_read')
            out[name.value.decode()] = dict(ms=ms.value, flops=fl.value, bytes=by.value, launches=int(la.value),
                                            calls=int(ca.value))
        return out

    def math_mode_name(self) -> str:
        return self._lib.rd_math_mode_name(self._h).decode()


def debug_rows(engine, kind, src, batch, h, w, c, w_kn, w_nk, n, out, stream):
    check(lib().rd_debug_rows(engine, kind, src, batch, h, w, c, w_kn, w_nk, n, out, stream), 'rd_debug_rows')


def debug_reduce(engine, kind, src, batch, h, w, c, g, n, out, scratch, scratch_floats, stream):
    check(lib().rd_debug_reduce(engine, kind, src, batch, h, w, c, g, n, out, scratch, scratch_floats, stream),
          'rd_debug_reduce')


def residuals(*args):
    check(lib().rd_residuals(*args), 'rd_residuals')


def residual_stats(res_ptr, valid_ptr, n, threshold, stream):
    out = (C.c_double * 16)()
    check(lib().rd_residual_stats(res_ptr, valid_ptr, n, float(threshold), out, stream), 'rd_residual_stats')
    return list(out)


def tile_stds(*args):
    check(lib().rd_tile_stds(*args), 'rd_tile_stds')


def make_tiles(*args):
    check(lib().rd_make_tiles(*args), 'rd_make_tiles')


def adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, stream):
    check(lib().rd_adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, stream),
          'rd_adam_step')


def sgd_step(p, g, n, lr, weight_decay, grad_scale, stream):
    check(lib().rd_sgd_step(p, g, n, lr, weight_decay, grad_scale, stream), 'rd_sgd_step')


def blend_accumulate(tiles, mean, std, geom, n, tile, stride, raster, rows, cols, stream):
    check(lib().rd_blend_accumulate(tiles, mean, std, geom, n, tile, stride, raster, rows, cols, stream),
          'rd_blend_accumulate')
