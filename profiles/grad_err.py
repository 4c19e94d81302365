import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.cases import CASES, NATIVE_CASES, spec_of, batch_of
from tests import test_gpu_parity as T
from oracle import unet_oracle as O
for name in NATIVE_CASES:
    kwargs, B, Tt = CASES[name]
    spec = spec_of(kwargs)
    model = T._model(kwargs)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pkeys = [k for k, _ in model.named_parameters()]
    for k in pkeys: sd[k].requires_grad_(True)
    batch = batch_of(name)
    loss_ref, grads_ref, y_ref = O.train_step(sd, pkeys, batch, spec, None)
    model = model.to(T.DEV)
    y, loss, grads, dy = T._train_step(model, batch, None)
    flat = torch.cat([grads[k].cpu().flatten() for k in pkeys]); flat_ref = torch.cat([grads_ref[k].flatten() for k in pkeys])
    worst = max((T._rel(grads[k].cpu(), grads_ref[k]), k) for k in pkeys if float(grads_ref[k].norm()) > 1e-7)
    print(f'{name:14s} y_rel={T._rel(y.cpu(), y_ref):.2e} flat_grad_rel={T._rel(flat, flat_ref):.3e} worst={worst[0]:.2e} {worst[1]}')
