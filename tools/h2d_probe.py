import torch, time, sys, os
sys.path.insert(0, '/root/repo')
from mmearth_train_b200 import synthetic as fo
dev = torch.device('cuda', 0)
host = [{k: v.pin_memory() for k, v in fo.synthetic_batch(256, 56, None, seed=i).items()} for i in range(4)]
nbytes = sum(v.numel() * v.element_size() for v in host[0].values())
bufs = {k: torch.empty_like(v, device=dev) for k, v in host[0].items()}
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20):
        for k, v in host[i % 4].items():
            bufs[k].copy_(v, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D {nbytes/1e6:.1f} MB x20: {dt/20*1e3:.2f} ms per batch, {nbytes*20/dt/1e9:.1f} GB/s")
print({k: (tuple(v.shape), str(v.dtype)) for k, v in host[0].items()})
