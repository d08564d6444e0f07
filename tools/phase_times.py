"""Device time between the phase boundaries of the training iteration (critic forward / R1 backward / all-reduce join / Adam,
generator forward / backward / join / Adam), averaged over a few iterations; run alone or under torchrun to see what the
data-parallel run adds and where:  [torchrun --nproc-per-node 2 ...] python tools/phase_times.py [workload]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import dist as bdist  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "train256"
rank, world, local = bdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
steps, alpha, batch, _, _ = bench.WORKLOADS[workload]
tr = bench.make_trainer(steps, alpha, batch, dev, style_mixing=workload in bench.STYLE_MIXING_DEFAULT)
R = 4 * 2 ** (steps - 1)
real = torch.rand(batch, 3, R, R, device=dev) * 2 - 1
z = torch.randn(2, batch, 512, device=dev).clamp_(-0.75, 0.75)
for _ in range(5):
    tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
torch.cuda.synchronize()
agg = collections.OrderedDict()
N = 8
for _ in range(N):
    tr._marks = []
    tr.iteration(real.clone(), z[0].clone(), z[1].clone(), read_losses=False)
    torch.cuda.synchronize()
    marks = tr._marks
    for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
        agg[n1] = agg.get(n1, 0.0) + e0.elapsed_time(e1)
    agg["total"] = agg.get("total", 0.0) + marks[0][1].elapsed_time(marks[-1][1])
if rank == 0:
    print(f"{workload} world {world}: device ms per phase (ends at the named mark), mean of {N} iterations")
    for k, v in agg.items():
        print(f"  {k:22s} {v / N:8.3f}")
if world > 1:
    torch.distributed.destroy_process_group()
