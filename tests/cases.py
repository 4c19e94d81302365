"""Golden cases shared by the CPU and GPU test modules (same table as oracle/make_golden.py)."""
import os

import numpy as np
import torch

from oracle.unet_oracle import NetSpec, synthetic_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = {
    # name: (ctor kwargs, B, T)
    'kat1': (dict(n_input_channels=1, start_kernel=64, depth=3, bias_conv_layer=True), 4, 64),
    'kat2': (dict(n_input_channels=3, start_kernel=64, depth=5, bias_conv_layer=True), 2, 256),
    'var_base': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True), 2, 32),
    'var_lrelu': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True,
                       act_fn_encoder='lrelu', act_fn_decoder='lrelu', act_fn_bottleneck='lrelu'), 2, 32),
    'var_prelu': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True,
                       act_fn_encoder='prelu', act_fn_decoder='prelu', act_fn_bottleneck='prelu'), 2, 32),
    'var_nobn': (dict(n_input_channels=2, start_kernel=32, depth=2, bias_conv_layer=True, do_BN=False), 2, 32),
    'var_outerbn': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=False,
                         outer_skip_BN=True), 2, 32),
    'var_noouter': (dict(n_input_channels=1, start_kernel=32, depth=2, bias_conv_layer=True,
                         outer_skip=False), 2, 32),
    'var_bilinear': (dict(n_input_channels=3, start_kernel=32, depth=2, bias_conv_layer=True,
                          up_mode='bilinear'), 2, 32),
    'var_cap': (dict(n_input_channels=3, start_kernel=32, max_filter_depth=64, depth=3,
                     bias_conv_layer=True), 3, 32),
}
# constructor variants the CUDA path implements
NATIVE_CASES = list(CASES)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f'{name}.npz'), allow_pickle=False)


def spec_of(kwargs) -> NetSpec:
    return NetSpec(**kwargs)


def torch_reference_module(kwargs):
    """A plain-PyTorch module tree with the reference's registration order, built by OUR mirror class (its
    sub-modules are ordinary nn modules); used on the CPU only to obtain seed-for-seed initial weights."""
    from resdepth_b200.lib.UNet import UNet
    torch.manual_seed(0)
    return UNet(**kwargs)


def batch_of(name):
    kwargs, B, T = CASES[name]
    return synthetic_batch(B, kwargs['n_input_channels'], T)
