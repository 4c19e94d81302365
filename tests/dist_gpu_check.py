#!/usr/bin/env python
"""Multi-GPU check of the data-parallel product path (run by hand on a GPU box, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py

Every rank builds the model from a DIFFERENT seed and hands the SAME global batches to
``resdepth_b200.lib.Trainer`` (fixed seed, as the reference's train.py does, train.py:86-87).  Checks:
  * the constructor broadcast makes the replicas identical (rank 0's weights);
  * the Trainer partitions every loader batch (rank r trains tiles [r*B/N, (r+1)*B/N));
  * after K steps through inference_one_epoch (graph replays, all-reduce slices overlapping the backward pass) the
    parameter arenas are bit-identical on all ranks AND bit-identical to a single-process emulation on rank 0 that
    runs the N shards one after the other, sums their gradient arenas and applies Adam with grad_scale 1/N;
  * the validation metric is the same number on every rank, and validation leaves the BatchNorm running statistics
    (per rank during training, as in DDP) averaged and identical on all replicas.
Prints one JSON line on rank 0.
"""
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle.unet_oracle import synthetic_batch  # noqa: E402


def make_trainer(model, batches, out_dir, single_process=False):
    """The real constructor (replica broadcast, shard policy, logging on rank 0 only).  single_process=True builds the
    rank-0-only emulation trainer: torch.distributed is hidden from it so that it issues no collectives."""
    from resdepth_b200.lib.Trainer import Trainer
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-5)
    args = SimpleNamespace(trainloader=batches, valloader=batches[:2], model=model, optimizer=opt, scheduler=None,
                           criterion=torch.nn.L1Loss(reduction='mean'), n_epochs=1, evaluate_rate=1, save_model_rate=1,
                           freq_average_train_loss=1000, save_dir=out_dir, log_file=None,
                           checkpoint_dir=os.path.join(out_dir, 'ckpt'), tboard_log_dir=None, pretrained_path=None)
    if not single_process:
        return Trainer(args)
    real = torch.distributed.is_initialized
    torch.distributed.is_initialized = lambda: False
    try:
        return Trainer(args)
    finally:
        torch.distributed.is_initialized = real


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from resdepth_b200.lib.distributed import replicas_identical, shard_batch
    from resdepth_b200.lib.UNet import UNet
    kwargs = dict(n_input_channels=3, start_kernel=64, depth=4, bias_conv_layer=True)
    B, T, K = 4 * world, 128, 6
    batches = [synthetic_batch(B, 3, T, seed=60 + i) for i in range(3)] * (K // 3)

    torch.manual_seed(1000 + rank)                      # replicas start DIFFERENT on purpose
    model = UNet(**kwargs)
    import tempfile
    out_dir = tempfile.mkdtemp(prefix=f'rd_dist_r{rank}_')
    tr = make_trainer(model, batches, out_dir)
    assert tr.distributed and tr.shard_batches == {'train': True, 'val': True}
    same0, _ = replicas_identical(model._rt['arena'], model._rt['bufs'], device=dev)
    start = model._rt['arena'].clone()
    meters = tr.inference_one_epoch(0, 'train')
    same1, cs = replicas_identical(model._rt['arena'], device=dev)        # parameters; BN statistics are per rank
    n_graphs = len(tr._graphs)
    val = tr._validate(0, meters['MAE_metric'].avg)['MAE_metric'].avg
    same2, _ = replicas_identical(model._rt['arena'], model._rt['bufs'], device=dev)   # statistics averaged by _validate
    v = torch.tensor([val], dtype=torch.float64, device=dev)
    vmin, vmax = v.clone(), v.clone()
    dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)

    result = {'world': world, 'identical_after_broadcast': same0, 'identical_after_steps': same1, 'checksum': cs,
              'graphs_captured': n_graphs, 'buffers_identical_after_validation': same2, 'val_same_on_all_ranks': bool(vmin.item() == vmax.item()), 'val': val}
    if rank == 0:
        # single-process emulation of the same K data-parallel steps, eager launches
        os.environ['RESDEPTH_GRAPHS'] = '0'
        torch.manual_seed(1000)
        emu = UNet(**kwargs)
        etr = make_trainer(emu, batches, out_dir + '_emu', single_process=True)
        assert not etr.distributed and etr._reducer is None and etr.shard_batches == {'train': False, 'val': False}
        assert torch.equal(emu._rt['arena'], start)
        bufs_r0 = None
        for b in batches:
            total = torch.zeros_like(emu._rt['grads'])
            keep = emu._rt['bufs'].clone()
            for r in range(world):
                emu._rt['bufs'].copy_(keep)                     # BatchNorm statistics are per rank: follow rank 0's
                loss = etr._launch_batch(shard_batch(b, r, world), 'train')
                total += emu._rt['grads']
                if r == 0:
                    bufs_r0 = emu._rt['bufs'].clone()
            emu._rt['bufs'].copy_(bufs_r0)
            emu._rt['grads'].copy_(total)
            etr.optimizer.grad_scale = 1.0 / world
            etr.optimizer.step()
        result['bitwise_equal_to_emulation'] = bool(torch.equal(emu._rt['arena'], model._rt['arena']))
        result['max_abs_diff_to_emulation'] = float((emu._rt['arena'] - model._rt['arena']).abs().max())
        ok = same0 and same1 and same2 and result['val_same_on_all_ranks'] and result['max_abs_diff_to_emulation'] <= 1e-6
        result['ok'] = bool(ok)
        print(json.dumps(result), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if (rank != 0 or result.get('ok')) else 1


if __name__ == '__main__':
    sys.exit(main())
