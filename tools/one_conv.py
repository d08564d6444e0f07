"""Runs one conv layer a few times (for ncu): python tools/one_conv.py R Cin Cout [batch] [mode: fprop|wgrad]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import bg_native as bgn  # noqa: E402

R, ci, co = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 32
mode = sys.argv[5] if len(sys.argv) > 5 else "fprop"
DEV = "cuda"
x = torch.randn(n, R, R, ci, device=DEV).to(torch.bfloat16)
g = torch.randn(n, R, R, co, device=DEV).to(torch.bfloat16)
w = torch.randn(co, ci, 3, 3, device=DEV)
wf = torch.empty(9, co, ci, dtype=torch.bfloat16, device=DEV)
wd = torch.empty(9, ci, co, dtype=torch.bfloat16, device=DEV)
bgn.call("bg_pack_weight", w, wf, wd, co, ci, ci, 3, math.sqrt(2 / (9 * ci)))
out = torch.empty(n, R, R, co, dtype=torch.bfloat16, device=DEV)
dwp = torch.empty(9, co, ci, device=DEV)
bias = torch.zeros(co, device=DEV)
for _ in range(4):
    if mode == "fprop":
        bgn.call("bg_conv_fprop", x, wf, out, n, R, R, ci, co, 3, bias, None, None, None, 1, 0.2)
    else:
        bgn.call("bg_conv_wgrad", x, g, dwp, n, R, R, ci, co, 0)
torch.cuda.synchronize()
print("ok")
