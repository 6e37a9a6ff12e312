"""Host-side cost of one step through the module API (enqueue time without synchronising), to see how far the CPU is ahead
of the GPU:  python tools/host_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmearth_train_b200 as mp
from bench import CONFIGS
from mmearth_train_b200.optim import FlatAdamW
from mmearth_train_b200 import synthetic as fo

cfg = CONFIGS["cfg2"]
args = fo.make_args(cfg["out_modalities"], cfg["loss_aggr"])
model = mp.convnextv2_atto(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True, patch_size=8, img_size=56,
                           args=args, loss_fn=mp.UncertaintyWeightingStrategy(12)).cuda()
opt = FlatAdamW(model)
batch = {k: v.cuda() for k, v in fo.synthetic_batch(256, 56, None, seed=1).items()}
def step():
    loss = model(batch, mask_ratio=0.6)[0]
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
for _ in range(5):
    step()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    tf = tb = 0.0
    for _ in range(20):
        a = time.perf_counter()
        loss = model(batch, mask_ratio=0.6)[0]
        b = time.perf_counter()
        loss.backward()
        c = time.perf_counter()
        opt.step(); opt.zero_grad(set_to_none=True)
        tf += b - a; tb += c - b
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0) / 20:.2f} ms/step (forward {1e3 * tf / 20:.2f}, backward {1e3 * tb / 20:.2f}), "
          f"with final sync {1e3 * (t2 - t0) / 20:.2f} ms/step")
