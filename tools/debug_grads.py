import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import golden_util as gu
from tests.test_parity_gpu import build_native
case = sys.argv[1] if len(sys.argv) > 1 else "atto_p8_all_unc"
backend = int(sys.argv[2]) if len(sys.argv) > 2 else 1
z, meta, orc, batch, noise = gu.inputs(case)
model = build_native(meta["cfg"], orc, backend)
model.noise_override = noise
loss = model({k: v.cuda() for k, v in batch.items()})[0]
loss.backward()
o_loss = orc(batch, mask_ratio=0.6, noise=noise)[0]
og = gu.oracle_grads(orc, o_loss)
named = dict(model.named_parameters())
for n, g in og.items():
    if g is None: continue
    e = gu.rel_err(named[n].grad, g)
    if e > 1e-3: print(f"{n:60s} {e:.3e}  norm {float(g.norm()):.3e}")
print("loss", float(loss), float(o_loss))
