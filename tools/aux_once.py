"""The HBM-bound helpers of the 256x256 stage, each launched alone at its train256 shape (batch 32): event time and
algorithmic GB/s per call, and a fixed launch sequence for ncu (`ncu --set full -k regex:<kernel> ... python tools/aux_once.py`).

    python tools/aux_once.py [R] [C] [batch]
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import bg_native as bgn  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 32
DEV = "cuda"
HW, P = R * R, n * R * R
bf = lambda *s: torch.randn(*s, device=DEV).to(torch.bfloat16)  # noqa: E731
a, g, out_map = bf(n, R, R, C), bf(n, R, R, C), torch.empty(n, R, R, C, dtype=torch.bfloat16, device=DEV)
planes = torch.randn(n, 3, R, R, device=DEV)
out_planes = torch.empty(n, 3, R, R, device=DEV)
w_to, w_from = torch.randn(3, C, 1, 1, device=DEV), torch.randn(C, 3, 1, 1, device=DEV)
b3, bc = torch.randn(3, device=DEV), torch.randn(C, device=DEV)
stats = torch.empty(n, C, 2, device=DEV)
bgn.call("bg_in_stats", a, stats, n, HW, C)
style = torch.cat([1 + 0.1 * torch.randn(n, C, device=DEV), 0.1 * torch.randn(n, C, device=DEV)], dim=1).contiguous()
bs = torch.zeros(n, C, 2, device=DEV)
ws = torch.zeros(2, C, device=DEV)
wsum4 = torch.zeros(4, C, device=DEV)
noise = torch.randn(n, 1, R, R, device=DEV)
g_hi = bf(n, R, R, 2 * C)                      # hi-res gradient of the layer below (R/2 -> R, 2C channels)
g_lo = torch.empty(n, R // 2, R // 2, 2 * C, dtype=torch.bfloat16, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

MAP = 2.0 * P * C
PL = 4.0 * 3 * P
CALLS = [
    ("bg_nhwc_to_planes3 (fromRGB dgrad)", MAP + PL,
     lambda: bgn.call("bg_nhwc_to_planes3", g, w_from, None, out_planes, P, HW, C, 3, 1, 0.8)),
    ("bg_to_rgb_adain", MAP + PL,
     lambda: bgn.call("bg_to_rgb_adain", a, stats, style, w_to, b3, out_planes, n, HW, C, 0.25, 1e-8)),
    ("bg_planes3_to_nhwc (fromRGB fwd)", MAP + PL,
     lambda: bgn.call("bg_planes3_to_nhwc", planes, w_from, bc, None, out_map, P, HW, C, 3, 1, 0.8, 1, 0.2)),
    ("bg_planes3_to_nhwc (gated)", 2 * MAP + PL,
     lambda: bgn.call("bg_planes3_to_nhwc", planes, w_from, None, a, out_map, P, HW, C, 3, 1, 0.8, 0, 0.2)),
    ("bg_channel_wsum (3 planes)", MAP + PL,
     lambda: bgn.call("bg_channel_wsum", g, planes, wsum4, P, C, HW, 3 * HW, HW, 3)),
    ("bg_adain_bwd_reduce", 2 * MAP,
     lambda: bgn.call("bg_adain_bwd_reduce", g, a, stats, bs, n, HW, C, 1e-8)),
    ("bg_adain_bwd_apply (+bias/noise sums)", 3 * MAP + 4.0 * P,
     lambda: bgn.call("bg_adain_bwd_apply", g, a, stats, style, bs, out_map, n, HW, C, 1e-8, 0.2, 1, noise, ws)),
    ("bg_adain_apply", 2 * MAP,
     lambda: bgn.call("bg_adain_apply", a, stats, style, out_map, n, HW, C, 1e-8)),
    ("bg_upsample2x_bwd (2C channels)", 2.0 * P * 2 * C * 1.25,
     lambda: bgn.call("bg_upsample2x_bwd", g_hi, g_lo, n, R // 2, R // 2, 2 * C)),
]
print(f"{R}x{R}, C {C}, batch {n}: one call after an L2 flush, median of 5")
for name, nbytes, fn in CALLS:
    ts = []
    for _ in range(6):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts[1:])[2]
    print(f"  {name:40s} {t * 1e3:7.1f} us   {nbytes / t / 1e6:7.0f} GB/s")
