"""One small G+D iteration per stage (for compute-sanitizer): python tools/tiny_iteration.py [max_steps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "byo-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import parity_util as U  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

max_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g, c = U.build_models(2)
for steps in range(1, max_steps + 1):
    for alpha in (None, 0.4):
        args = (O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps), O.make_images(batch, steps, 30 + steps),
                O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps))
        r = U.cuda_iteration(g, c, *args, steps, alpha, 10.0)
        r2 = U.cuda_iteration(g, c, *args, steps, alpha, 10.0, loss="wgan", epsilon=O.make_epsilon(batch, steps))
        torch.cuda.synchronize()
        print(f"steps {steps} alpha {alpha}: c_loss {r['c_loss'].item():.4f} g_loss {r['g_loss'].item():.4f} wgan {r2['c_loss'].item():.4f}", flush=True)
print("done")
