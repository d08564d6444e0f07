"""Per-tensor parity table over every golden train-iteration case (deterministic mode): rel-L2 and cosine of each
parameter gradient vs the fp32 oracle, the same for the oracle's bf16-storage emulation, their ratio and the norm ratio —
the data the per-tensor gate in tests/parity_util.py is calibrated on.
Usage: python tools/parity_table.py [out.txt]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "byo-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import parity_util as U  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
cases = [(c["steps"], c["batch"], c["alpha"], "r1") for c in json.load(open(os.path.join(ROOT, "tests/golden/train_iteration.json")))]
cases += [(c["steps"], c["batch"], c["alpha"], "r1_penalty") for c in json.load(open(os.path.join(ROOT, "tests/golden/r1_penalty.json")))]
cases += [(c["steps"], c["batch"], c["alpha"], "wgan") for c in json.load(open(os.path.join(ROOT, "tests/golden/wgan_gp.json")))]
cases += [(7, 32, None, "r1"), (8, 16, None, "r1")]
U.no_tf32()
worst_ratio = []
for steps, batch, alpha, loss in cases:
    args = (O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps), O.make_images(batch, steps, 30 + steps),
            O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps))
    kw = dict(loss=loss)
    if loss == "wgan":
        kw["epsilon"] = O.make_epsilon(batch, 40 + steps)
    g, c = U.build_models(2)
    with U.deterministic():
        r = U.cuda_iteration(g, c, *args, steps, alpha, 10.0, **kw)
    del g, c
    torch.cuda.empty_cache()
    o = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, 10.0, device="cuda", **kw)
    O.QUANT[0] = True
    q = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, 10.0, device="cuda", **kw)
    O.QUANT[0] = False
    print(f"== steps={steps} B={batch} alpha={alpha} loss={loss}: c_loss {r['c_loss'].item():.6f} / oracle {o['c_loss'].item():.6f} / emu "
          f"{q['c_loss'].item():.6f}   g_loss {r['g_loss'].item():.6f} / {o['g_loss'].item():.6f} / {q['g_loss'].item():.6f}   "
          f"image rel {U.rel(r['fake_d'], o['fake_d']):.3e} (emu {U.rel(q['fake_d'], o['fake_d']):.3e})", file=out)
    for kind in ("d_grads", "g_grads"):
        for k, ref in o[kind].items():
            got = r[kind][k]
            if ref is None or ref.norm().item() == 0:
                continue
            e, ee = U.rel(got, ref), U.rel(q[kind][k], ref)
            ratio = got.double().norm().item() / ref.double().norm().item()
            gate = U.grad_gate(k, got, ref, q[kind][k])
            worst_ratio.append((e / (ee + 1e-12), e, ee, f"s{steps}b{batch}a{alpha}/{loss}", kind, k, ref.numel()))
            print(f"{kind:8s} {k:48s} n={ref.numel():8d} rel {e:.4f} emu {ee:.4f} x{e / (ee + 1e-12):5.2f} cos {U.cos(got, ref):.5f} "
                  f"(emu {U.cos(q[kind][k], ref):.5f}) |g|/|ref| {ratio:.3f} {'FAIL ' + gate if gate else ''}", file=out)
    out.flush()
print("== largest rel / emulated-rel ratios", file=out)
for row in sorted(worst_ratio, reverse=True)[:40]:
    print("  x%5.2f rel %.4f emu %.4f  %s %s %s n=%d" % row, file=out)
