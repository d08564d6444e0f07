"""Run-to-run spread of the gradient parity errors of one golden train-iteration case (atomics make the CUDA path
non-deterministic in the last bits; the piecewise-linear net amplifies that through LeakyReLU gate flips)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "byo-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import parity_util as U  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 5
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
case = json.load(open(os.path.join(ROOT, "tests", "golden", "train_iteration.json")))[idx]
U.no_tf32()
steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
args = (O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps), O.make_images(batch, steps, 30 + steps),
        O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps))
o = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, lam, device="cuda")
O.QUANT[0] = True
q = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, lam, device="cuda")
O.QUANT[0] = False
worst = {}
for rep in range(reps):
    g, c = U.build_models(2)
    r = U.cuda_iteration(g, c, *args, steps, alpha, lam)
    for kind in ("d_grads", "g_grads"):
        for k, ref in o[kind].items():
            if ref is None or ref.norm().item() == 0:
                continue
            e = U.rel(r[kind][k], ref)
            w = worst.setdefault((kind, k), [])
            w.append(e)
rows = sorted(worst.items(), key=lambda kv: -max(kv[1]))[:12]
print(f"case s{steps} b{batch} a{alpha}: worst tensors over {reps} runs (min / max rel-L2; bf16-emulated oracle's rel-L2)")
for (kind, k), es in rows:
    print(f"  {kind:8s} {k:48s} {min(es):.3f} / {max(es):.3f}   emu {U.rel(q[kind][k], o[kind][k]):.3f}")
