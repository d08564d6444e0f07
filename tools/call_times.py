"""Device time of every C-ABI call of one G+D iteration, grouped by (entry point, integer arguments): CUDA events
around each call (so a call's time includes its memset/zero kernel and the gap in front of it).

    python tools/call_times.py [workload] [batch] > gpurun_out/call_times.txt
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts byo-gan_b200 on sys.path)
import torch  # noqa: E402
import dist as bdist  # noqa: E402
import engine  # noqa: E402
import gan  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "train256"
steps, alpha, batch, _, _ = bench.WORKLOADS[workload]
if len(sys.argv) > 2:
    batch = int(sys.argv[2])
dev = torch.device("cuda", 0)
tr = bench.make_trainer(steps, alpha, batch, dev, style_mixing=workload in bench.STYLE_MIXING_DEFAULT)
R = 4 * 2 ** (steps - 1)
real = torch.rand(batch, 3, R, R, device=dev) * 2 - 1
z = torch.randn(2, batch, 512, device=dev).clamp_(-0.75, 0.75)
for _ in range(6):
    tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
torch.cuda.synchronize()

records = []
inner = engine.call


def timed_call(name, *args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = inner(name, *args)
    b.record()
    key = (name, tuple(x for x in args if isinstance(x, int) and not isinstance(x, bool)))
    records.append((key, a, b))
    return r


for mod in (engine, gan):
    if getattr(mod, "call", None) is inner:
        mod.call = timed_call
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
t1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, a, b in records:
    e = agg.setdefault(key, [0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b) * 1e3
total = sum(v[1] for v in agg.values())
print(f"{workload} batch {batch}: {len(records)} calls, {total / 1e3:.3f} ms inside calls, iteration {t0.elapsed_time(t1):.3f} ms (event overhead included)")
byname = collections.defaultdict(float)
for (name, _), (c, t) in agg.items():
    byname[name] += t
for name, t in sorted(byname.items(), key=lambda kv: -kv[1]):
    print(f"== {name}: {t:.1f} us")
    for (n2, ints), (c, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if n2 == name:
            print(f"     {c:3d} x {tt / c:8.1f} us   {ints}")
