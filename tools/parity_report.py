"""Prints a per-tensor parity table (CUDA path vs oracle fp32 on GPU) for one G+D iteration.
Usage: python tools/parity_report.py steps batch [alpha]"""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "byo-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402
import parity_util as U  # noqa: E402

steps, batch = int(sys.argv[1]), int(sys.argv[2])
alpha = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "None" else None
lam = 10.0
U.no_tf32()
g, c = U.build_models(2)
args = (O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps), O.make_images(batch, steps, 30 + steps),
        O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps))
o = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, lam, device="cuda")
O.QUANT[0] = True
oq = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, lam, device="cuda")
O.QUANT[0] = False
try:
    r = U.cuda_iteration(g, c, *args, steps, alpha, lam)
except Exception:
    traceback.print_exc()
    sys.exit(1)
print(f"steps={steps} B={batch} alpha={alpha}")
print(f"c_loss {r['c_loss'].item():.6f} vs fp32 {o['c_loss'].item():.6f} / bf16-emu {oq['c_loss'].item():.6f}   g_loss {r['g_loss'].item():.6f} vs {o['g_loss'].item():.6f} / {oq['g_loss'].item():.6f}")
print(f"fake rel {U.rel(r['fake_d'], o['fake_d']):.3e}  pred_fake {r['pred_fake'].flatten()[:4].tolist()} vs {o['pred_fake'].flatten()[:4].tolist()}")
print(f"pred_real {r['pred_real'].flatten()[:4].tolist()} vs {o['pred_real'].flatten()[:4].tolist()}")
print(f"z_grad rel {U.rel(r['z_grad'], o['z_grad']):.3e} cos {U.cos(r['z_grad'], o['z_grad']):.5f}")
for kind in ("d_grads", "g_grads"):
    for k, ref in o[kind].items():
        got = r[kind][k]
        if ref is None or got is None:
            if (ref is None) != (got is None):
                print(f"{kind:8s} {k:50s} NONE-MISMATCH got={got is not None} ref={ref is not None}")
            continue
        print(f"{kind:8s} {k:50s} rel {U.rel(got, ref):.3e} cos {U.cos(got, ref):.5f} |ref| {ref.norm().item():.3e}"
              f"  || vs bf16-emu: rel {U.rel(got, oq[kind][k]):.3e} cos {U.cos(got, oq[kind][k]):.5f}")
