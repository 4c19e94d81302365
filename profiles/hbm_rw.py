import torch, time
x = torch.empty(1<<28, dtype=torch.float32, device='cuda')   # 1 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
gb = x.numel()*4/1e9
print('fill (write only)  GB/s', gb/t(lambda: x.fill_(1.0))*1e3)
print('sum (read only)    GB/s', gb/t(lambda: x.sum())*1e3)
print('copy (read+write)  GB/s', 2*gb/t(lambda: y.copy_(x))*1e3)
