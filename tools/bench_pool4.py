"""conv3x3+AvgPool2d: epilogue-pooled 3x3 kernel vs the 4x4 stride-2 formulation (python tools/bench_pool4.py [batch])"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import bg_native as bgn  # noqa: E402

DEV = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32


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


for (R, ci, co) in [(256, 64, 64), (128, 128, 128), (64, 256, 256), (32, 512, 512)]:
    x = torch.randn(n, R, R, ci, device=DEV).to(torch.bfloat16)
    w = torch.randn(co, ci, 3, 3, device=DEV)
    wf = torch.empty(9, co, ci, dtype=torch.bfloat16, device=DEV)
    wd = torch.empty(9, ci, co, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight", w, wf, wd, co, ci, ci, 3, 0.05)
    w16 = torch.empty(16, co, ci, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight_pool4", w, w16, co, ci, 0.05)
    out = torch.empty(n, R // 2, R // 2, co, dtype=torch.bfloat16, device=DEV)
    bias = torch.zeros(co, device=DEV)
    t1 = timeit(lambda: bgn.call("bg_conv_pool_fprop", x, wf, out, n, R, R, ci, co, bias, None, 1, 0.2))
    t2 = timeit(lambda: bgn.call("bg_conv_pool4_fprop", x, w16, out, n, R, R, ci, co, bias, None, 1, 0.2))
    # input gradient: pool adjoint materialised at full resolution + 3x3 dgrad, vs the transposed 4x4 stride-2 form
    gpool = torch.randn(n, R // 2, R // 2, co, device=DEV).to(torch.bfloat16)
    y2 = torch.randn(n, R // 2, R // 2, co, device=DEV).to(torch.bfloat16)
    gate = torch.randn(n, R, R, ci, device=DEV).to(torch.bfloat16)
    gu = torch.empty(n, R, R, co, dtype=torch.bfloat16, device=DEV)
    gx = torch.empty(n, R, R, ci, dtype=torch.bfloat16, device=DEV)
    wt = torch.empty(16, ci, co, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight_tconv4", w, wt, co, ci, 0.05)

    def old_dgrad():
        bgn.call("bg_pool_act_bwd", gpool, y2, gu, n, R // 2, R // 2, co, 0.2, None)
        bgn.call("bg_conv_fprop", gu, wd, gx, n, R, R, co, ci, 3, None, None, None, gate, 0, 0.2)

    def new_dgrad():
        bgn.call("bg_act_gate", gpool, y2, gpool, gpool.numel(), 0.2)
        bgn.call("bg_conv_pool4_dgrad", gpool, wt, gx, n, R // 2, R // 2, co, ci, gate, 0.2, None)

    t3, t4 = timeit(old_dgrad), timeit(new_dgrad)
    print(f"{R:4d} {ci:4d} {co:4d}  fprop: conv+pool epilogue {t1:.3f} ms, 4x4 stride-2 {t2:.3f} ms | "
          f"dgrad: pool_bwd + 3x3 {t3:.3f} ms, gate + transposed 4x4s2 {t4:.3f} ms")
