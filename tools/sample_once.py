"""One profiled 512x512 sampling forward (after warm-ups) for ncu: python tools/sample_once.py [batch] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
import torch  # noqa: E402
import gan  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
g = gan.Generator().to(dev).eval()
with torch.no_grad():
    for n, p in g.named_parameters():
        if n.endswith("bias") or n.endswith("inject_noise.weights"):
            p.add_(0.05 * torch.randn_like(p))
z = torch.randn(batch, 512, device=dev).clamp_(-0.75, 0.75)
with torch.no_grad():
    for _ in range(2):
        g(z, steps=steps)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    img = g(z, steps=steps)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", tuple(img.shape))
