"""Two of the streaming reductions at their largest train256 shapes, for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:"in_reduce|channel_wsum" -o gpurun_out/elem python tools/one_elem.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import torch  # noqa: E402
import bg_native as bgn  # noqa: E402

dev = torch.device("cuda", 0)
N, HW, C = 32, 65536, 32
a = torch.randn(N, HW, C, device=dev).to(torch.bfloat16)
g = torch.randn(N, HW, C, device=dev).to(torch.bfloat16)
stats = torch.empty(N, C, 2, device=dev)
bsums = torch.empty(N, C, 2, device=dev)
img = torch.randn(N, 3, HW, device=dev)
out = torch.empty(4, C, device=dev)
bgn.call("bg_in_stats", a, stats, N, HW, C)
for _ in range(3):
    bgn.call("bg_adain_bwd_reduce", g, a, stats, bsums, N, HW, C, 1e-5)
    bgn.call("bg_channel_wsum", g, img, out, N * HW, C, HW, 3 * HW, HW, 3)
torch.cuda.synchronize()
t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t[0].record()
bgn.call("bg_adain_bwd_reduce", g, a, stats, bsums, N, HW, C, 1e-5)
t[1].record()
bgn.call("bg_channel_wsum", g, img, out, N * HW, C, HW, 3 * HW, HW, 3)
t[2].record()
torch.cuda.synchronize()
print("adain_bwd_reduce %.1f us, channel_wsum %.1f us" % (t[0].elapsed_time(t[1]) * 1e3, t[1].elapsed_time(t[2]) * 1e3))
