"""Container-only tests (``reference`` marker): need the upstream repository mounted at /root/reference."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shims.reference_available(), reason='reference not mounted')]


def test_reference_construction_path_runs_on_the_shims():
    """lib/utils.py get_model / get_optimizer / get_loss / get_trainer of the unmodified reference, with
    resdepth_b200's classes installed as lib.UNet / lib.Trainer."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_dropin_check.py')], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out['model'] == 'resdepth_b200.lib.UNet.UNet' and out['state_dict_equal']
    assert 'pretrained_path' in out['trainer_args']


def test_golden_checkpoint_is_what_the_generator_writes(tmp_path):
    """tests/golden/ref_checkpoint.pth really is the reference's checkpoint dictionary (keys, scalar types) and loads
    with weights_only=True -- the mode _load_pretrain uses."""
    import torch
    ck = torch.load(os.path.join(ROOT, 'tests', 'golden', 'ref_checkpoint.pth'), map_location='cpu', weights_only=True)
    assert set(ck) == {'epoch', 'model_state_dict', 'optimizer_state_dict', 'loss_train', 'loss_val', 'scheduler_state_dict'}
    assert ck['epoch'] == 4 and abs(ck['loss_val'] - 0.987) < 1e-12
    assert set(ck['optimizer_state_dict']) == {'state', 'param_groups'}
