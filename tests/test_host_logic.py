"""CPU tests: the C-ABI library loads and exports every symbol of include/resdepth_b200.h, the native layer plan
mirrors the module tree of every constructor variant, host-side error behaviour, optimizer arena detection."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from resdepth_b200 import _native
from tests.cases import CASES, NATIVE_CASES, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'resdepth_b200.h')).read()
    declared = set(re.findall(r'\b(rd_[a-z0-9_]+)\s*\(', header))
    declared -= {'rd_handle', 'rd_config'}
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_native.library_path())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert _native.lib().rd_abi_version() == _native.ABI_VERSION


@pytest.mark.parametrize('name', NATIVE_CASES)
def test_native_plan_matches_module_tree(name):
    from resdepth_b200.lib.UNet import UNet
    kwargs, _, _ = CASES[name]
    torch.manual_seed(0)
    model = UNet(**kwargs)
    h = _native.Handle(model._config(), 0)           # rd_create builds the plan only; no device work
    named = dict(model.named_parameters())
    infos = h.param_infos()
    assert [n for n, _, _ in infos] == list(named.keys())
    assert [n for n, _, _ in infos] == [str(k) for k in load_golden(name)['param_keys']]
    end = 0
    for n, numel, off in infos:
        assert numel == named[n].numel()
        assert off % 4 == 0 and off >= end
        end = off + numel
    assert h.param_arena_size() >= end
    bufs = dict(model.named_buffers())
    for n, numel, off in h.buffer_infos():
        assert bufs[n].numel() == numel and off % 4 == 0
    assert {n for n, _, _ in h.buffer_infos()} == {k for k in bufs if not k.endswith('num_batches_tracked')}
    h.close()


def test_constructor_errors_match_reference_behaviour():
    from resdepth_b200.lib.UNet import UNet
    with pytest.raises(ValueError):
        UNet(act_fn_encoder='gelu')
    with pytest.raises(ValueError):
        UNet(up_mode='nearest')
    m = UNet(n_input_channels=3, start_kernel=32, depth=2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))                  # CPU tensors are refused: no fallback path


def test_rd_create_rejects_unsupported_plans():
    cfg = _native.RdConfig(n_input_channels=3, start_kernel=48, max_filter_depth=512, depth=3, do_bn=1, outer_skip=1)
    with pytest.raises(RuntimeError, match='start_kernel'):
        _native.Handle(cfg, 0)
    cfg = _native.RdConfig(n_input_channels=3, start_kernel=64, max_filter_depth=512, depth=3, do_bn=1, outer_skip=1,
                           up_mode=7)
    with pytest.raises(RuntimeError, match='up_mode'):
        _native.Handle(cfg, 0)
    cfg = _native.RdConfig(n_input_channels=9, start_kernel=64, max_filter_depth=512, depth=3, do_bn=1)
    with pytest.raises(RuntimeError, match='n_input_channels'):
        _native.Handle(cfg, 0)


def test_flat_span_detection():
    from resdepth_b200.lib.optim import _flat_span
    arena = torch.zeros(64)
    a, b, c = arena[0:10].view(2, 5), arena[12:13], arena[16:48].view(4, 8)
    ptr, n = _flat_span([c, a, b])
    assert ptr == arena.data_ptr() and n == 48
    assert _flat_span([a, c]) is None                  # hole where b should be
    assert _flat_span([a, torch.zeros(3)]) is None     # different storages
    assert _flat_span([arena[1:5]]) is None            # misaligned start


def test_fuse_optimizer_keeps_class_name_state_and_scheduler():
    from resdepth_b200.lib.optim import Adam, SGD, fuse_optimizer
    p = [torch.nn.Parameter(torch.zeros(4))]
    opt = torch.optim.Adam(p, lr=2e-4, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    fused = fuse_optimizer(opt)
    assert fused is opt and isinstance(opt, Adam) and opt.__class__.__name__ == 'Adam'
    assert isinstance(fuse_optimizer(torch.optim.SGD(p, lr=0.1)), SGD)
    with pytest.raises(NotImplementedError):
        fuse_optimizer(torch.optim.RMSprop(p))
    p[0].grad = torch.ones(4)
    with pytest.raises(RuntimeError):                 # CPU parameters: the fused step refuses, no fallback
        opt.step()
    sched.step()
    assert abs(opt.param_groups[0]['lr'] - 1e-4) < 1e-12


def test_average_meter_and_denormalize_helpers():
    from resdepth_b200.lib.AverageMeter import AverageMeter
    from resdepth_b200.lib.data_normalization import denormalize_numpy, denormalize_torch
    m = AverageMeter()
    m.update(2.0)
    m.update(4.0, n=3)
    assert m.avg == pytest.approx(3.5) and m.count == 4 and m.val == 4.0
    x = torch.arange(8.).view(2, 1, 2, 2)
    mean, std = torch.tensor([10., 20.]), torch.tensor([2., 3.])
    out = denormalize_torch(x, mean, std)
    assert torch.equal(out[1], x[1] * 3 + 20)
    np.testing.assert_array_equal(denormalize_numpy(x, mean, std), out.numpy())
    assert torch.equal(denormalize_torch(x, 1.0, 2.0), x * 2 + 1)


def test_epoch_prefetcher_keeps_order_and_stages_one_batch_ahead():
    """Trainer._prefetched yields the loader's batches in order and has batch i+1 staged (its H2D copies enqueued)
    before batch i is handed to the step -- checked with a recording stand-in for the CUDA staging."""
    from resdepth_b200.lib.Trainer import Trainer
    tr = Trainer.__new__(Trainer)
    events = []

    def fake_stage(batch, phase="train"):
        events.append(('stage', batch))
        return {'staged': batch}
    tr._stage_batch = fake_stage
    seen = []
    for b in tr._prefetched(iter([10, 11, 12, 13])):
        events.append(('step', b['staged']))
        seen.append(b['staged'])
    assert seen == [10, 11, 12, 13]
    assert events == [('stage', 10), ('stage', 11), ('step', 10), ('stage', 12), ('step', 11), ('stage', 13),
                      ('step', 12), ('step', 13)]
    assert list(tr._prefetched(iter([]))) == []
    assert [b['staged'] for b in tr._prefetched([7])] == [7]


def test_rd_config_layout_matches_header_and_integration_snippet():
    """The ctypes struct, the C header and the binding snippet in INTEGRATION.md list the same int32 fields in the
    same order (a short struct handed to rd_create reads garbage: VERDICT r1 'docs bug with teeth')."""
    import ctypes
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, 'include', 'resdepth_b200.h')).read()
    body = header[header.index('typedef struct rd_config {'):header.index('} rd_config;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in re.findall(r'int32_t\s+([^;]+);', body):
        fields += [n.strip() for n in decl.split(',')]
    from resdepth_b200 import _native
    assert [n for n, _ in _native.RdConfig._fields_] == fields
    assert ctypes.sizeof(_native.RdConfig) == 4 * len(fields) == 56
    doc = open(os.path.join(root, 'INTEGRATION.md')).read()
    snippet = doc[doc.index('class RdConfig(C.Structure)'):doc.index('lib.rd_create.argtypes')]
    assert re.findall(r"'(\w+)'", snippet) == fields
    # every symbol the header declares is bound, and vice versa
    declared = set(re.findall(r'\b(rd_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)


def test_backward_stage_ranges_tile_the_gradient_arena_for_every_depth():
    """rd_grad_stage_range (the data-parallel slices): stage 2 | stage 1 | stage 0 are contiguous, start at 0 and end at
    the arena size, stage 1 starts at encoder level min(3, depth-1) and stage 0 at the first up-conv -- for every depth,
    without a GPU (the layer plan is host-side)."""
    for depth in (1, 2, 3, 5, 6):
        cfg = _native.RdConfig(n_input_channels=3, start_kernel=64, max_filter_depth=512, depth=depth, do_bn=1,
                               bias_conv_layer=1, outer_skip=1, math_mode=1)
        h = _native.Handle(cfg, 0)
        spans = [h.grad_stage_range(s) for s in range(3)]
        assert spans[2][0] == 0
        assert spans[2][0] + spans[2][1] == spans[1][0]
        assert spans[1][0] + spans[1][1] == spans[0][0]
        assert spans[0][0] + spans[0][1] == h.param_arena_size()
        offsets = {n: off for n, _, off in h.param_infos()}
        assert spans[1][0] == offsets[f'encoder.{min(3, depth - 1)}.0.0.weight']
        assert spans[0][0] == offsets['decoder.0.0.weight' if depth > 1 else 'decoder.0.weight']
        assert all(n >= 0 for _, n in spans) and (depth == 1) == (spans[2][1] == 0)
        assert h.workspace_id() == 0 and not h.workspace_alive(1)      # no layout before the first forward call
        h.close()


def test_backward_math_knob_reaches_the_native_config():
    from resdepth_b200.lib.UNet import UNet
    m = UNet(n_input_channels=3, start_kernel=32, depth=2)
    assert m._config().bwd_mode == _native.BWD_IDS['auto']
    m.backward_math = 'tf32'
    assert m._config().bwd_mode == _native.BWD_IDS['tf32']
    m.backward_math = 'fp8'
    with pytest.raises(ValueError, match='backward_math'):
        m._config()
    cfg = _native.RdConfig(n_input_channels=3, start_kernel=64, max_filter_depth=512, depth=3, do_bn=1, bwd_mode=7)
    with pytest.raises(RuntimeError, match='bwd_mode'):
        _native.Handle(cfg, 0)
    for mode, name in ((0, 'bf16'), (1, 'tf32'), (2, 'bf16')):
        h = _native.Handle(_native.RdConfig(n_input_channels=3, start_kernel=64, max_filter_depth=512, depth=3, do_bn=1,
                                            math_mode=1, bwd_mode=mode), 0)
        assert h.bwd_mode_name() == name
        h.close()
    h = _native.Handle(_native.RdConfig(n_input_channels=3, start_kernel=64, max_filter_depth=512, depth=3, do_bn=1,
                                        math_mode=0), 0)
    assert h.bwd_mode_name() == 'fp32'
    h.close()


def test_trainer_shards_host_batches_before_staging():
    """Trainer._my_shard: rank r of a data-parallel job takes tiles [r*B/N, (r+1)*B/N) of every loader batch unless the
    loader already shards; staged (device) batches pass through untouched."""
    from resdepth_b200.lib.Trainer import Trainer, _StagedBatch
    tr = Trainer.__new__(Trainer)
    tr.rank, tr.world_size = 1, 4
    tr.shard_batches = {'train': True, 'val': False}
    batch = {'input': torch.arange(8.).view(8, 1, 1, 1), 'dsm_std': torch.arange(8.), 'tile_size': 256}
    mine = tr._my_shard(batch, 'train')
    assert mine['input'].flatten().tolist() == [2.0, 3.0] and mine['dsm_std'].tolist() == [2.0, 3.0]
    assert mine['tile_size'] == 256
    assert tr._my_shard(batch, 'val') is batch
    staged = _StagedBatch(batch)
    assert tr._my_shard(staged, 'train') is staged


def test_first_block_weight_gradient_identity_used_by_the_cuda_path():
    """DESIGN.md 4.7, checked against PyTorch autograd in float64 on the CPU: for conv(bias-free) -> BatchNorm(train) ->
    ReLU, the weight gradient equals  cs * ( X^T gY - c1 * cx - c2 * (G W - cx * mean) )  with X the im2col expansion,
    G = X^T X, cx = X^T 1, gY the gradient at the BatchNorm output times the ReLU mask, c1 = mean(gY),
    c2 = mean(gY * xhat) / sigma, cs = gamma / sigma -- no second pass over z."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    B, Cin, Co, H = 3, 3, 8, 10
    x = torch.randn(B, Cin, H, H, generator=g, dtype=torch.float64)
    w = (0.3 * torch.randn(Co, Cin, 3, 3, generator=g, dtype=torch.float64)).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(Co, generator=g, dtype=torch.float64))
    beta = 0.1 * torch.randn(Co, generator=g, dtype=torch.float64)
    z = F.conv2d(x, w, padding=1)
    y = F.batch_norm(z, None, None, gamma, beta, True, 0.1, 1e-5)
    a = F.relu(y)
    up = torch.randn(a.shape, generator=g, dtype=torch.float64)            # gradient arriving at the activation
    (a * up).sum().backward()
    # the pieces the CUDA path has: X (im2col), gY, batch statistics
    X = F.unfold(x, 3, padding=1).transpose(1, 2).reshape(-1, Cin * 9)      # [pixels][27], k = ci*9 + r*3 + s
    n = X.shape[0]
    zc = z.detach().permute(0, 2, 3, 1).reshape(n, Co)
    mean, var = zc.mean(0), zc.var(0, unbiased=False)
    sigma = torch.sqrt(var + 1e-5)
    gY = (up * (y.detach() > 0)).permute(0, 2, 3, 1).reshape(n, Co)
    xhat = (zc - mean) / sigma
    cs, c1, c2 = gamma / sigma, gY.mean(0), (gY * xhat).mean(0) / sigma
    G, cx, A = X.T @ X, X.sum(0), X.T @ gY                                   # [27][27], [27], [27][Co]
    W2 = w.detach().reshape(Co, Cin * 9)
    dW = cs[:, None] * (A.T - c1[:, None] * cx[None, :] - c2[:, None] * (W2 @ G - mean[:, None] * cx[None, :]))
    assert torch.allclose(dW.reshape(w.shape), w.grad, rtol=1e-9, atol=1e-10)
