"""One profiled G+D iteration (after 3 warm-ups) for ncu:  ncu --profile-from-start off ... python tools/profile_step.py [workload] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (puts byo-gan_b200 on sys.path)
import torch  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "train256"
steps, alpha, batch, _, _ = bench.WORKLOADS[workload]
if len(sys.argv) > 2:
    batch = int(sys.argv[2])
dev = torch.device("cuda", 0)
tr = bench.make_trainer(steps, alpha, batch, dev, style_mixing=workload in bench.STYLE_MIXING_DEFAULT)
R = 4 * 2 ** (steps - 1)
real = torch.rand(batch, 3, R, R, device=dev) * 2 - 1
z = torch.randn(2, batch, 512, device=dev).clamp_(-0.75, 0.75)
for _ in range(3):
    tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one iteration", workload, "batch", batch)
