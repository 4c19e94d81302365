"""CPU: the oracle (oracle/unet_oracle.py) reproduces the vectors the UNMODIFIED reference produced
(tests/golden/*.npz, generator oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
from tests.cases import CASES, batch_of, load_golden, spec_of, torch_reference_module


def _state(kwargs):
    m = torch_reference_module(kwargs)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    pkeys = [k for k, _ in m.named_parameters()]
    return sd, pkeys


@pytest.mark.parametrize('name', list(CASES))
def test_init_matches_reference_seed_for_seed(name):
    kwargs, _, _ = CASES[name]
    g = load_golden(name)
    sd, pkeys = _state(kwargs)
    assert list(sd.keys()) == [str(k) for k in g['keys']]
    assert pkeys == [str(k) for k in g['param_keys']]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g['shapes']]
    got = np.array([float(v.double().sum()) for v in sd.values()])
    np.testing.assert_allclose(got, g['init_sum'], rtol=0, atol=1e-9)
    got = np.array([float(v.double().abs().sum()) for v in sd.values()])
    np.testing.assert_allclose(got, g['init_abs'], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_train_step_and_eval_match_reference(name):
    kwargs, B, T = CASES[name]
    g = load_golden(name)
    spec = spec_of(kwargs)
    sd, pkeys = _state(kwargs)
    batch = batch_of(name)
    assert abs(float(batch['input'].double().sum()) - float(g['x_sum'])) < 1e-9
    for k in pkeys:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in pkeys], lr=2e-4, weight_decay=1e-5)
    loss, grads, y = O.train_step(sd, pkeys, batch, spec, opt)
    np.testing.assert_allclose(y.numpy(), g['y_train'], rtol=0, atol=2e-5)
    assert abs(loss - float(g['loss_train'])) < 2e-6 * max(1.0, abs(float(g['loss_train'])))
    gn = np.array([float(grads[k].double().norm()) for k in pkeys])
    np.testing.assert_allclose(gn, g['grad_norm'], rtol=2e-3, atol=1e-6)
    for k in g.files:
        if k.startswith('grad::'):
            ref = g[k]
            np.testing.assert_allclose(grads[k[6:]].numpy(), ref, rtol=0, atol=2e-4 * max(1e-3, np.abs(ref).max()))
    # the first Adam step moves every weight by ~lr*sign(grad): sums are only stable up to sign flips of
    # noise-level gradients, so the tolerance scales with the tensor size (tight check: test_adam_* below)
    keys = [str(s) for s in g['keys']]
    post = np.array([float(sd[k].detach().double().sum()) for k in keys])
    tol = np.array([2e-4 * max(2.0, 0.002 * sd[k].numel()) for k in keys])
    assert np.all(np.abs(post - g['post_sum']) <= tol), np.abs(post - g['post_sum']).max()
    with torch.no_grad():
        y_eval = O.unet_forward(sd, batch['input'], spec, training=False)
        loss_eval = O.denormalized_l1(y_eval, batch['target'], batch['loss_mask'], batch['dsm_mean'], batch['dsm_std'])
    np.testing.assert_allclose(y_eval.numpy(), g['y_eval'], rtol=0, atol=5e-5)
    assert abs(float(loss_eval) - float(g['loss_eval'])) < 5e-6 * max(1.0, abs(float(g['loss_eval'])))


def test_explicit_batchnorm_formula_equals_aten_call():
    kwargs, B, T = CASES['var_base']
    spec = spec_of(kwargs)
    outs = []
    for explicit in (False, True):
        O.EXPLICIT_BN = explicit
        try:
            sd, _ = _state(kwargs)
            with torch.no_grad():
                y = O.unet_forward(sd, batch_of('var_base')['input'], spec, training=True)
            outs.append((y, sd['encoder.1.0.1.running_var'].clone(), int(sd['bottleneck.1.num_batches_tracked'])))
        finally:
            O.EXPLICIT_BN = False
    np.testing.assert_allclose(outs[0][0].numpy(), outs[1][0].numpy(), rtol=0, atol=5e-6)
    np.testing.assert_allclose(outs[0][1].numpy(), outs[1][1].numpy(), rtol=1e-5)
    assert outs[0][2] == outs[1][2] == 1


def test_closed_form_loss_equals_reference_form():
    b = O.synthetic_batch(3, 1, 16)
    y = b['input'] + 0.3 * torch.randn(3, 1, 16, 16, generator=torch.Generator().manual_seed(5))
    a = O.denormalized_l1(y, b['target'], b['loss_mask'], b['dsm_mean'], b['dsm_std'])
    c = O.masked_l1_closed_form(y, b['target'], b['loss_mask'], b['dsm_std'])
    assert abs(float(a) - float(c)) < 1e-5 * float(a)


def test_adam_single_tensor_statement_matches_torch():
    g0 = torch.Generator().manual_seed(3)
    p = torch.randn(1000, generator=g0)
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=2e-4, weight_decay=1e-5)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(1000, generator=g0)
        p_ref.grad = grad.clone()
        opt.step()
        p, m, v = O.adam_reference_step(p, grad, m, v, step, 2e-4, wd=1e-5)
        np.testing.assert_allclose(p.numpy(), p_ref.detach().numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('name', ['blend_a', 'blend_b', 'blend_c'])
def test_blend_oracle_matches_reference(name):
    g = np.load(__import__('os').path.join(__import__('tests.cases', fromlist=['GOLDEN']).GOLDEN, 'blend.npz'))
    rows, cols, tile, stride = [int(v) for v in g[name + '_geom']]
    pos, box = O.regular_grid((0, cols - 1), (0, rows - 1), tile, stride)
    assert [tuple(p) for p in g[name + '_pos']] == pos
    assert [tuple(b) for b in g[name + '_box']] == box
    out = O.linear_blend(g[name + '_tiles'], g[name + '_mean'], g[name + '_std'], pos, box, rows, cols, tile, stride)
    np.testing.assert_allclose(out, g[name + '_raster'], rtol=0, atol=1e-9)
