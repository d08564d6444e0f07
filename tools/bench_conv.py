"""Micro-benchmark of the tcgen05 conv kernels on the model's layer shapes (CUDA events, L2 flushed by
rotating over buffers larger than L2).  Usage: python tools/bench_conv.py [batch] [max_res]"""
import os
import sys
import math

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import bg_native as bgn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
MAXR = int(sys.argv[2]) if len(sys.argv) > 2 else 256
DEV = "cuda"

# (R, Cin, Cout) distinct 3x3 layers of G (gan.py:159-166) and D (gan.py:320-327)
LAYERS = [(4, 512, 512), (8, 512, 512), (16, 512, 512), (32, 512, 256), (32, 256, 256), (32, 256, 512), (32, 512, 512),
          (64, 256, 128), (64, 128, 128), (64, 128, 256), (64, 256, 256), (128, 128, 64), (128, 64, 64), (128, 64, 128),
          (128, 128, 128), (256, 64, 32), (256, 32, 32), (256, 32, 64), (256, 64, 64), (512, 32, 16), (512, 16, 16),
          (512, 16, 32), (512, 32, 32)]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


print(f"batch {B}")
print(f"{'R':>4} {'Cin':>4} {'Cout':>4} | {'fprop ms':>9} {'TF/s':>7} {'GB/s':>7} | {'dgrad ms':>9} {'TF/s':>7} | {'wgrad ms':>9} {'TF/s':>7} | {'fp+stats ms':>10} {'TF/s':>7}")
for (R, ci, co) in LAYERS:
    if R > MAXR:
        continue
    n = B
    x = torch.randn(n, R, R, ci, device=DEV).to(torch.bfloat16)
    g = torch.randn(n, R, R, co, device=DEV).to(torch.bfloat16)
    w = torch.randn(co, ci, 3, 3, device=DEV)
    wf = torch.empty(9, co, ci, dtype=torch.bfloat16, device=DEV)
    wd = torch.empty(9, ci, co, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight", w, wf, wd, co, ci, ci, 3, math.sqrt(2 / (9 * ci)))
    out = torch.empty(n, R, R, co, dtype=torch.bfloat16, device=DEV)
    gx = torch.empty(n, R, R, ci, dtype=torch.bfloat16, device=DEV)
    dwp = torch.empty(9, co, ci, device=DEV)
    bias = torch.zeros(co, device=DEV)
    flops = 2.0 * n * R * R * 9 * ci * co
    t_f = timeit(lambda: bgn.call("bg_conv_fprop", x, wf, out, n, R, R, ci, co, 3, bias, None, None, None, 1, 0.2))
    t_d = timeit(lambda: bgn.call("bg_conv_fprop", g, wd, gx, n, R, R, co, ci, 3, None, None, None, None, 0, 0.2))
    t_w = timeit(lambda: bgn.call("bg_conv_wgrad", x, g, dwp, n, R, R, ci, co, 0))
    stats = torch.empty(n, co, 2, device=DEV)
    t_t = timeit(lambda: bgn.call("bg_conv_fprop_stats", x, wf, out, n, R, R, ci, co, 3, bias, None, None, None, 1, 0.2, stats, 1))
    byts = 2.0 * n * R * R * (ci + co)
    print(f"{R:>4} {ci:>4} {co:>4} | {t_f:9.3f} {flops / t_f / 1e9:7.1f} {byts / t_f / 1e6:7.0f} | {t_d:9.3f} {flops / t_d / 1e9:7.1f} | "
          f"{t_w:9.3f} {flops / t_w / 1e9:7.1f} | {t_t:10.3f} {flops / t_t / 1e9:7.1f}", flush=True)
