"""Kernel timeline of G+D iterations in flight (CUPTI through torch.profiler, NOT ncu: kernels run back to back with
warm caches): device-busy time, idle gaps between kernels, and per-kernel hot durations.

    python tools/timeline_step.py [workload] [batch] [iters] > gpurun_out/timeline.txt
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts byo-gan_b200 on sys.path)
import torch  # noqa: E402
import dist as bdist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "train256"
steps, alpha, batch, _, _ = bench.WORKLOADS[workload]
if len(sys.argv) > 2:
    batch = int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
tr = bench.make_trainer(steps, alpha, batch, dev, style_mixing=workload in bench.STYLE_MIXING_DEFAULT)
R = 4 * 2 ** (steps - 1)
real = torch.rand(batch, 3, R, R, device=dev) * 2 - 1
z = torch.randn(2, batch, 512, device=dev).clamp_(-0.75, 0.75)
for _ in range(12):
    tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
    torch.cuda.synchronize()

ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() >= 0:
        ev.append((e.time_range.start, e.time_range.end, e.name))
ev.sort()
t0, t1 = ev[0][0], max(e[1] for e in ev)
busy, gaps, cur_end = 0.0, [], ev[0][0]
for s, e, _ in ev:
    if s > cur_end:
        gaps.append(s - cur_end)
        busy += e - s
        cur_end = e
    elif e > cur_end:
        busy += e - cur_end
        cur_end = e
span = t1 - t0
# the long gaps, with the activities either side (where the device waited for the host)
cur_end, prev_name = ev[0][0], ""
for s, e, nme in ev:
    if s - cur_end >= 20:
        print(f"gap {s - cur_end:7.1f} us at +{(s - t0) / 1e3:8.3f} ms   after {prev_name[:70]!r}   before {nme[:70]!r}")
    if e > cur_end:
        cur_end, prev_name = e, nme
print(f"{workload} batch {batch}: {iters} iterations, {len(ev)} device activities, span {span / iters / 1e3:.3f} ms/iter, "
      f"busy {busy / iters / 1e3:.3f} ms/iter, idle {(span - busy) / iters / 1e3:.3f} ms/iter in {len(gaps) // iters} gaps/iter")
hist = collections.Counter()
for g in gaps:
    hist[min(int(g), 20)] += 1
print("gap histogram (us: count/iter, total us/iter):",
      {k: (round(v / iters, 1), round(sum(g for g in gaps if min(int(g), 20) == k) / iters, 1)) for k, v in sorted(hist.items())})
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ev:
    a = agg[n.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]]
    a[0] += 1
    a[1] += e - s
print(f"{'kernel':60s} {'n/iter':>7s} {'us/iter':>9s} {'share':>6s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{n:60s} {c / iters:7.1f} {t / iters:9.1f} {100 * t / busy:5.1f}%")
