"""Runs bg_conv_pool4_dgrad on one layer a few times (for ncu): python tools/one_pool4_dgrad.py Rp C [batch]
Rp = pooled resolution (the output map is 2Rp x 2Rp), C = channels of both maps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import bg_native as bgn  # noqa: E402

Rp, C = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 32
DEV = "cuda"
w = torch.randn(C, C, 3, 3, device=DEV)
wt = torch.empty(16, C, C, dtype=torch.bfloat16, device=DEV)
bgn.call("bg_pack_weight_tconv4", w, wt, C, C, 0.05)
gpool = torch.randn(n, Rp, Rp, C, device=DEV).to(torch.bfloat16)
gate = torch.randn(n, 2 * Rp, 2 * Rp, C, device=DEV).to(torch.bfloat16)
gx = torch.empty(n, 2 * Rp, 2 * Rp, C, dtype=torch.bfloat16, device=DEV)
db = torch.empty(C, device=DEV)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 2:
        s.record()
    bgn.call("bg_conv_pool4_dgrad", gpool, wt, gx, n, Rp, Rp, C, C, gate, 0.2, db)
e.record()
torch.cuda.synchronize()
print(f"pool4_dgrad pooled {Rp} C {C} batch {n}: {s.elapsed_time(e) / 4 * 1e3:.1f} us per call")
