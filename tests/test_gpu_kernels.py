"""GPU kernel-level tests through the C-ABI test hooks: each GEMM-shaped kernel (CUDA-core fp32 and tcgen05
TF32) against a float64 statement of the same contraction computed with plain tensor indexing on the CPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _gather_cpu(src, kind, B, H, W, C):
    """A[p][(tap, c)] as float64; src NHWC."""
    s = src.double()
    if kind == 0:
        pad = torch.zeros(B, H + 2, W + 2, C, dtype=torch.float64)
        pad[:, 1:-1, 1:-1] = s
        taps = [pad[:, r:r + H, q:q + W] for r in range(3) for q in range(3)]
    elif kind == 1:
        taps = [s]
    else:
        taps = [s[:, a::2, b::2] for a in range(2) for b in range(2)]
    return torch.stack(taps, dim=3).reshape(B * H * W, len(taps) * C)


def _tf32(t):
    return (t.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if False else \
        ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


ROWS_CASES = [  # kind, B, H, W, C, N
    (0, 2, 16, 16, 64, 128), (0, 4, 8, 8, 256, 256), (0, 3, 4, 4, 64, 64), (0, 2, 32, 32, 32, 32),
    (0, 1, 24, 24, 64, 64), (1, 2, 8, 8, 128, 512), (1, 3, 4, 4, 64, 256), (2, 2, 8, 8, 64, 64),
    (2, 3, 4, 4, 128, 128), (2, 1, 16, 16, 32, 32), (0, 2, 64, 64, 64, 512),
]


@pytest.mark.parametrize('engine', [0, 1])
@pytest.mark.parametrize('kind,B,H,W,C,N', ROWS_CASES)
def test_rows_kernels(engine, kind, B, H, W, C, N):
    from resdepth_b200 import _native
    g = torch.Generator().manual_seed(kind * 100 + C + N + H)
    ups = 2 if kind == 2 else 1
    src = _tf32(torch.randn(B, ups * H, ups * W, C, generator=g))
    ntaps = {0: 9, 1: 1, 2: 4}[kind]
    w_kn = _tf32(torch.randn(ntaps * C, N, generator=g) / (ntaps * C) ** 0.5)
    ref = _gather_cpu(src, kind, B, H, W, C) @ w_kn.double()
    d_src, d_kn, d_nk = src.to(DEV), w_kn.to(DEV), w_kn.t().contiguous().to(DEV)
    out = torch.full((B * H * W, N), float('nan'), device=DEV)
    _native.debug_rows(engine, kind, d_src.data_ptr(), B, H, W, C, d_kn.data_ptr(), d_nk.data_ptr(), N,
                       out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = float((out.cpu().double() - ref).abs().max())
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), err     # operands are exact TF32 values: only fp32 accumulation differs


REDUCE_CASES = [  # kind, B, H, W, C, N
    (0, 4, 8, 8, 64, 128), (0, 2, 16, 16, 128, 64), (0, 3, 4, 4, 64, 64), (0, 2, 32, 32, 32, 32),
    (0, 1, 24, 24, 64, 64), (2, 2, 8, 8, 64, 64), (2, 3, 4, 4, 128, 128), (2, 1, 16, 16, 32, 32),
    (0, 2, 64, 64, 64, 256), (0, 4, 8, 8, 256, 512),
]


@pytest.mark.parametrize('engine', [0, 1])
@pytest.mark.parametrize('kind,B,H,W,C,N', REDUCE_CASES)
def test_reduce_kernels(engine, kind, B, H, W, C, N):
    from resdepth_b200 import _native
    g = torch.Generator().manual_seed(kind * 100 + C + N + H + 7)
    ups = 2 if kind == 2 else 1
    src = _tf32(torch.randn(B, ups * H, ups * W, C, generator=g))
    G = _tf32(torch.randn(B * H * W, N, generator=g))
    ntaps = {0: 9, 2: 4}[kind]
    ref = _gather_cpu(src, kind, B, H, W, C).t() @ G.double()
    d_src, d_G = src.to(DEV), G.to(DEV)
    out = torch.full((ntaps * C, N), float('nan'), device=DEV)
    scratch = torch.zeros(8 << 20, device=DEV)
    _native.debug_reduce(engine, kind, d_src.data_ptr(), B, H, W, C, d_G.data_ptr(), N, out.data_ptr(),
                         scratch.data_ptr(), scratch.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = float((out.cpu().double() - ref).abs().max())
    assert err <= 1e-4 * max(1.0, float(ref.abs().max())), err


BF16_ROWS = [(0, 2, 16, 16, 64, 128), (0, 4, 8, 8, 256, 256), (0, 2, 32, 32, 64, 64), (0, 1, 24, 24, 128, 64),
             (2, 2, 8, 8, 64, 64), (2, 3, 4, 4, 128, 128), (0, 2, 64, 64, 64, 512)]


@pytest.mark.parametrize('kind,B,H,W,C,N', BF16_ROWS)
def test_rows_kernel_bf16(kind, B, H, W, C, N):
    """tcgen05 kind::f16 path of the backward GEMMs: operands are bf16 tensors, accumulation fp32."""
    from resdepth_b200 import _native
    g = torch.Generator().manual_seed(kind * 100 + C + N + H + 3)
    ups = 2 if kind == 2 else 1
    src = torch.randn(B, ups * H, ups * W, C, generator=g).bfloat16()
    ntaps = {0: 9, 1: 1, 2: 4}[kind]
    w_kn = (torch.randn(ntaps * C, N, generator=g) / (ntaps * C) ** 0.5).bfloat16()
    ref = _gather_cpu(src.float(), kind, B, H, W, C) @ w_kn.double()
    d_src, d_nk = src.to(DEV), w_kn.t().contiguous().to(DEV)
    out = torch.full((B * H * W, N), float('nan'), device=DEV)
    _native.debug_rows(2, kind, d_src.data_ptr(), B, H, W, C, None, d_nk.data_ptr(), N, out.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = float((out.cpu().double() - ref).abs().max())
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), err


BF16_REDUCE = [(0, 4, 8, 8, 64, 128), (0, 2, 16, 16, 128, 64), (0, 2, 32, 32, 64, 64), (2, 2, 8, 8, 64, 64),
               (2, 3, 4, 4, 128, 128), (0, 2, 64, 64, 64, 256), (0, 4, 8, 8, 256, 512), (1, 2, 32, 32, 64, 64)]


@pytest.mark.parametrize('kind,B,H,W,C,N', BF16_REDUCE)
def test_reduce_kernel_bf16(kind, B, H, W, C, N):
    from resdepth_b200 import _native
    g = torch.Generator().manual_seed(kind * 100 + C + N + H + 11)
    ups = 2 if kind == 2 else 1
    src = torch.randn(B, ups * H, ups * W, C, generator=g).bfloat16()
    G = torch.randn(B * H * W, N, generator=g).bfloat16()
    ntaps = {0: 9, 1: 1, 2: 4}[kind]
    ref = _gather_cpu(src.float(), kind, B, H, W, C).t() @ G.double()
    d_src, d_G = src.to(DEV), G.to(DEV)
    out = torch.full((ntaps * C, N), float('nan'), device=DEV)
    scratch = torch.zeros(8 << 20, device=DEV)
    _native.debug_reduce(2, kind, d_src.data_ptr(), B, H, W, C, d_G.data_ptr(), N, out.data_ptr(),
                         scratch.data_ptr(), scratch.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = float((out.cpu().double() - ref).abs().max())
    assert err <= 1e-4 * max(1.0, float(ref.abs().max())), err
