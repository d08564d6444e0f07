"""Per-layer forward error of the generator's pre-AdaIN activations vs the fp32 oracle, fused vs unfused path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "byo-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import parity_util as U  # noqa: E402
import engine  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

steps, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 6, 4
U.no_tf32()
dev = "cuda"
P = {k: v.to(dev) for k, v in O.make_state("gen", 2).items()}
z = O.make_latents(batch, 10 + steps).to(dev)
noise = [n.to(dev) for n in O.make_noise(batch, steps, 10 + steps)]
# oracle activations
w_lat = O.mapping(P, z)
ref = []
x = None
for k in range(steps):
    if k > 0:
        x = O.bilinear_up2(x)
    for j in (1, 2):
        pre = f"gen_blocks.{k}.conv_{j}"
        if k == 0 and j == 1:
            out = P[f"{pre}.conv"].repeat(batch, 1, 1, 1)
        else:
            out = O.eq_conv2d(x, P[f"{pre}.conv.weight"], P[f"{pre}.conv.bias"], padding=1)
        out = O.lrelu(out + P[f"{pre}.inject_noise.weights"] * noise[k])
        ref.append(out)
        style = O.eq_linear(w_lat, P[f"{pre}.adain.style.weight"], P[f"{pre}.adain.style.bias"])
        c = out.shape[1]
        x = style[:, :c, None, None] * O.instance_norm(out) + style[:, c:, None, None]
img_ref = O.to_rgb(P, steps - 1, x)

g, _ = U.build_models(2)
for label, fn in (("fused", engine.style_conv_fusable), ("unfused", lambda R, c: False)):
    engine.style_conv_fusable = fn
    with torch.no_grad():
        img, tape = engine.generator_forward(g, g._packs, z, noise, steps, None, keep_tape=True)
    errs = [U.rel(L["a"].float().permute(0, 3, 1, 2), r) for L, r in zip(tape["layers"], ref)]
    print(label, "a rel-L2 per layer:", " ".join(f"{e:.4f}" for e in errs), " img:", f"{U.rel(img, img_ref):.4f}")
