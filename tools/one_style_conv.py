"""Runs bg_style_modulate + bg_conv_style_fprop on one generator layer a few times (for ncu):
python tools/one_style_conv.py R Cin Cout upsample(0|1) [batch]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import bg_native as bgn  # noqa: E402

R, ci, co, up = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 64
DEV = "cuda"
rin = R // 2 if up else R
x = torch.randn(n, rin, rin, ci, device=DEV).to(torch.bfloat16)
w = torch.randn(co, ci, 3, 3, device=DEV)
bias = torch.randn(co, device=DEV) * 0.1
stats_prev = torch.stack([torch.randn(n, ci, device=DEV) * rin * rin * 0.1, torch.rand(n, ci, device=DEV) * rin * rin + rin * rin], dim=2).contiguous()
style = torch.cat([1 + 0.1 * torch.randn(n, ci, device=DEV), 0.1 * torch.randn(n, ci, device=DEV)], dim=1).contiguous()
wmod = torch.empty(n, 9, co, ci, dtype=torch.bfloat16, device=DEV)
btab = torch.empty(n, 9, co, device=DEV)
bgn.call("bg_style_modulate", w, bias, stats_prev, style, wmod, btab, n, ci, co, rin * rin, math.sqrt(2 / (9 * ci)), 1e-8)
noise = torch.randn(n, 1, R, R, device=DEV)
nw = torch.randn(co, device=DEV) * 0.1
out = torch.empty(n, R, R, co, dtype=torch.bfloat16, device=DEV)
stats = torch.empty(n, co, 2, device=DEV)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 2:
        s.record()
    bgn.call("bg_conv_style_fprop", x, wmod, btab, out, n, R, R, ci, co, up, noise, nw, 0.2, stats)
e.record()
torch.cuda.synchronize()
by = 2.0 * n * (ci * rin * rin + co * R * R)
t = s.elapsed_time(e) / 4
print(f"style conv {R}x{R} {ci}->{co} up={up} batch {n}: {t * 1e3:.1f} us per call, {by / t / 1e6:.0f} GB/s algorithmic")
