"""Device time of every C-ABI call of one 512x512 sampling batch (Generator.forward under no_grad), grouped by entry point and
shape, with the algorithmic bytes / flops of the conv calls: python tools/sample_call_times.py [batch]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts byo-gan_b200 on sys.path)
import torch  # noqa: E402
import bg_native as bgn  # noqa: E402
import gan  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
torch.manual_seed(0)
g = gan.Generator().to(dev)
z = torch.randn(batch, 512, device=dev).clamp_(-0.75, 0.75)
with torch.no_grad():
    for _ in range(3):
        g(z, steps=8)
    torch.cuda.synchronize()
    bgn.start_timing()
    g(z, steps=8)
    rec = bgn.stop_timing()
agg = collections.OrderedDict()
for name, a, t in rec:
    e = agg.setdefault((name, a[:6]), [0, 0.0, a])
    e[0] += 1
    e[1] += t
tot = sum(v[1] for v in agg.values())
print(f"sampling 512x512 batch {batch}: {len(rec)} calls, {tot:.3f} ms inside calls")
for (name, a), (c, t, full) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = ""
    if name in bench.FPROP_CALLS:
        fl, ex, by = bench.conv_call_work(name, full)
        extra = f"  {by * c / (t / 1e3) / 1e9:7.0f} GB/s {fl * c / (t / 1e3) / 1e12:6.0f} TF/s"
    print(f"  {c:2d} x {t / c * 1e3:8.1f} us  {name:24s} {a}{extra}")
