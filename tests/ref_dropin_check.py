#!/usr/bin/env python
"""Container-only check (needs /root/reference): the three shims of INTEGRATION.md installed as ``lib.UNet`` /
``lib.Trainer`` and the reference's OWN construction path driven on top of them -- ``lib.utils.get_model``
(lib/utils.py:295-316), ``get_optimizer`` (:319-341), ``get_scheduler``, ``get_loss`` and ``get_trainer`` (:380-441).

Run as a script (its own interpreter: the reference's top-level package is called ``lib``); prints one JSON line.
Without a CUDA device the construction is followed up to the first statement of our Trainer that needs the GPU.
"""
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import ref_shims  # noqa: E402
from oracle.unet_oracle import synthetic_batch  # noqa: E402


def main():
    ref_shims._install_stub_modules()
    sys.path.insert(0, ref_shims.REFERENCE_ROOT)
    import importlib
    our_trainer = importlib.import_module('resdepth_b200.lib.Trainer')   # the package re-exports the classes under
    our_unet = importlib.import_module('resdepth_b200.lib.UNet')         # the same names as these sub-modules

    seen = {}

    class RecordingTrainer(our_trainer.Trainer):          # same class, remembers what the reference handed over
        def __init__(self, args):
            seen['args'] = args
            super().__init__(args)

    # the shims a maintainer writes: the whole of lib/UNet.py and lib/Trainer.py
    shim_unet = types.ModuleType('lib.UNet')
    shim_unet.UNet, shim_unet.SkipConnection = our_unet.UNet, our_unet.SkipConnection
    shim_trainer = types.ModuleType('lib.Trainer')
    shim_trainer.Trainer = RecordingTrainer
    sys.modules['lib.UNet'] = shim_unet
    sys.modules['lib.Trainer'] = shim_trainer
    import lib.utils as utils                             # the UNMODIFIED reference module
    from easydict import EasyDict as edict
    assert utils.UNet is our_unet.UNet and utils.Trainer is RecordingTrainer

    out_dir = tempfile.mkdtemp(prefix='rd_dropin_')
    cfg = edict({
        'model': {'name': 'UNet', 'input_channels': 'geom-stereo', 'start_kernel': 32, 'depth': 2,
                  'act_fn_encoder': 'relu', 'act_fn_decoder': 'relu', 'act_fn_bottleneck': 'relu',
                  'up_mode': 'transpose', 'do_BN': True, 'outer_skip': True, 'outer_skip_BN': False,
                  'bias_conv_layer': True},
        'optimizer': {'name': 'Adam', 'learning_rate': 2e-4, 'weight_decay': 1e-5},
        'scheduler': {'enabled': True, 'name': 'StepLR', 'settings': {'gamma': 0.5, 'step_size': 3}},
        'training_settings': {'loss': 'L1', 'n_epochs': 2},
        'general': {'evaluate_rate': 1, 'save_model_rate': 1},
        'output': {'output_directory': out_dir, 'checkpoint_dir': os.path.join(out_dir, 'checkpoints'),
                   'tboard_log_dir': os.path.join(out_dir, 'tb')},
    })
    result = {}
    torch.manual_seed(0)
    model, args_model = utils.get_model(cfg)
    assert type(model) is our_unet.UNet
    result['model'] = type(model).__module__ + '.' + type(model).__name__
    result['n_input_channels'] = model.n_input_channels
    assert model.n_input_channels == 3 and model.depth == 2 and model.start_kernel == 32 and model.bias_conv_layer
    # same state_dict as the reference class built from the same settings and seed
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_unet_file', os.path.join(ref_shims.REFERENCE_ROOT, 'lib', 'UNet.py'))
    ref_unet = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_unet)
    torch.manual_seed(0)
    ref_model = ref_unet.UNet(**args_model.settings)
    sd_ref, sd = ref_model.state_dict(), model.state_dict()
    assert list(sd_ref) == list(sd)
    assert all(torch.equal(sd_ref[k], sd[k]) for k in sd)
    result['state_dict_equal'] = True

    optimizer = utils.get_optimizer(cfg, model)
    assert type(optimizer) is torch.optim.Adam
    try:
        scheduler = utils.get_scheduler(cfg, optimizer)
    except TypeError as e:                                 # torch >= 2.7 dropped StepLR(verbose=...): SURVEY.md section 7
        result['get_scheduler'] = f'reference incompatible with this torch: {e}'
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=3, gamma=0.5)
    criterion = utils.get_loss(cfg)
    batches = [synthetic_batch(4, 3, 32, seed=900 + i) for i in range(3)]
    try:
        trainer = utils.get_trainer(cfg, batches, batches[:1], model, optimizer, scheduler, criterion)
    except RuntimeError as e:
        if torch.cuda.is_available():
            raise
        assert 'needs a CUDA device' in str(e), e
        result['trainer'] = 'reached resdepth_b200 Trainer.__init__ (no CUDA device here)'
        trainer = None
    args = seen['args']
    want = ['trainloader', 'valloader', 'model', 'optimizer', 'scheduler', 'criterion', 'n_epochs', 'evaluate_rate',
            'save_model_rate', 'freq_average_train_loss', 'save_dir', 'log_file', 'checkpoint_dir', 'tboard_log_dir',
            'pretrained_path']
    missing = [k for k in want if k not in args]
    assert not missing, missing
    assert args.model is model and args.optimizer is optimizer and args.pretrained_path is None
    result['trainer_args'] = sorted(args.keys())
    if trainer is not None:                                # on a GPU box with the reference mounted: run it
        trainer.train()
        result['trained'] = os.path.isfile(trainer.path_model_last)
        assert result['trained'] and trainer.optimizer.__class__.__name__ == 'Adam'
    print(json.dumps(result))


if __name__ == '__main__':
    main()
