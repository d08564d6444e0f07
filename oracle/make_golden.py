"""Generates tests/golden/*.json by running the UNMODIFIED reference (`/root/reference/gan.py`).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Inputs are the deterministic recipes of oracle/gan_oracle.py (make_state / make_latents / make_noise /
make_images), loaded into the reference's own nn.Modules with load_state_dict(strict=True).  Outputs are
stored as compact fingerprints (shape, sum, norm, abs-max, 64 strided samples) so the fixtures stay small.
The training-iteration cases restate train.py:135-217 around the real modules because train.py/helper.py
cannot be imported here (matplotlib missing, helper.py:42 hard-codes .cuda()).
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import gan as ref_gan  # the reference  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

FORWARD_CASES = [
    # steps, batch, alpha
    (1, 2, None), (2, 2, None), (3, 2, None), (4, 2, None), (5, 2, None), (6, 2, None), (7, 2, None), (8, 1, None),
    (1, 4, 0.5), (2, 4, 0.0), (2, 4, 0.3), (3, 4, 1.7), (4, 4, 0.3), (5, 4, 0.5), (6, 2, 1.0),
]
TRAIN_CASES = [
    # steps, batch, alpha, lambda
    (1, 16, None, 10.0),   # BASELINE config 1
    (2, 8, 0.4, 10.0),
    (3, 8, None, 10.0),
    (4, 4, 0.5, 10.0),
    (5, 4, 0.5, 10.0),     # BASELINE config 2 shape at reduced batch
    (6, 4, None, 10.0),
    (7, 4, None, 10.0),    # BASELINE config 3 shape at reduced batch
    (8, 4, None, 10.0),    # BASELINE config 4 shape (512x512) at reduced batch
    (8, 4, 0.5, 10.0),     # ... with the fade-in branches
]
# critic step with ONLY the R1 penalty term (gan.py:398-404): purely second-order parameter gradients
PENALTY_CASES = [(2, 8, 0.4, 10.0), (4, 4, None, 10.0), (5, 4, 0.5, 10.0), (6, 4, None, 10.0)]
# WGAN-GP (train.py:177-185 -> gan.py:357-391), dead code in the reference: run here with its two defects repaired
WGAN_CASES = [(1, 8, None, 10.0), (2, 8, 0.4, 10.0), (4, 4, None, 10.0), (6, 4, 0.5, 10.0)]
MBSTD_BATCHES = [4, 8, 16, 32, 6, 12, 8]   # the tail reproduces the group_size mutation 6 -> 6 -> 8... (gan.py:277-278)


def load_ref(seed):
    gen = ref_gan.Generator()
    gen.load_state_dict(O.make_state("gen", seed), strict=True)
    critic = ref_gan.Critic()
    critic.load_state_dict(O.make_state("critic", seed), strict=True)
    return gen, critic


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    gen, critic = load_ref(seed=1)

    # ---- key/shape contract of the checkpoint layout (train.py:247-259)
    layout = {
        "gen": {k: list(v.shape) for k, v in ref_gan.Generator().state_dict().items()},
        "critic": {k: list(v.shape) for k, v in ref_gan.Critic().state_dict().items()},
    }
    json.dump(layout, open(os.path.join(OUT, "state_layout.json"), "w"), indent=0)

    # ---- forward passes
    fwd = []
    for steps, batch, alpha in FORWARD_CASES:
        t0 = time.time()
        z = O.make_latents(batch, seed=steps)
        noise = O.make_noise(batch, steps, seed=steps)
        img = O.make_images(batch, steps, seed=steps)
        with torch.no_grad():
            fake = gen(z, noise=noise, steps=steps, alpha=alpha)
            c = ref_gan.Critic()
            c.load_state_dict(critic.state_dict())
            pred_real = c(img, steps, alpha)
            pred_fake = c(fake, steps, alpha)
        fwd.append({"steps": steps, "batch": batch, "alpha": alpha, "fake": O.fingerprint(fake),
                    "pred_real": O.fingerprint(pred_real), "pred_fake": O.fingerprint(pred_fake)})
        print(f"forward steps={steps} B={batch} alpha={alpha}: {time.time() - t0:.1f}s", flush=True)
    json.dump(fwd, open(os.path.join(OUT, "forward.json"), "w"))

    # ---- minibatch stddev incl. the stateful group_size
    mb = ref_gan.MiniBatchStdDev()
    rows = []
    for i, b in enumerate(MBSTD_BATCHES):
        g = torch.Generator().manual_seed(100 + i)
        x = torch.randn(b, 512, 4, 4, generator=g)
        y = mb(x)
        rows.append({"batch": b, "seed": 100 + i, "group_size_after": mb.group_size, "plane": y[:, 512, 0, 0].tolist()})
    json.dump(rows, open(os.path.join(OUT, "mbstd.json"), "w"))

    # ---- bilinear impulse response (gan.py:112) and instance-norm (gan.py:59) known answers
    imp = torch.zeros(1, 1, 4, 4)
    imp[0, 0, 1, 2] = 1.0
    up = torch.nn.Upsample(scale_factor=2, mode="bilinear")(imp)
    g = torch.Generator().manual_seed(5)
    xin = torch.randn(2, 3, 4, 4, generator=g) * 3 + 1
    json.dump({"impulse_up": up[0, 0].tolist(), "in_x": xin.tolist(),
               "in_y": torch.nn.InstanceNorm2d(3, eps=1e-8)(xin).tolist()},
              open(os.path.join(OUT, "layers.json"), "w"))

    # ---- full G+D iterations (train.py:135-217 restated around the real modules, no optimizer step)
    train = []
    for steps, batch, alpha, lam in TRAIN_CASES:
        t0 = time.time()
        gen, critic = load_ref(seed=2)
        z_d, z_g = O.make_latents(batch, seed=10 + steps), O.make_latents(batch, seed=20 + steps)
        n_d, n_g = O.make_noise(batch, steps, seed=10 + steps), O.make_noise(batch, steps, seed=20 + steps)
        real = O.make_images(batch, steps, seed=30 + steps)
        # critic step
        for p in critic.parameters():
            p.requires_grad = True
        for p in gen.parameters():
            p.requires_grad = False
        z = z_d.clone().requires_grad_()
        fake = gen(z, noise=n_d, steps=steps, alpha=alpha)
        real_im = real.clone().requires_grad_()
        pf = critic(fake.detach(), steps, alpha)
        pr = critic(real_im, steps, alpha)
        critic.zero_grad()
        c_loss = critic.get_r1_loss(pf, pr, real_im, fake, steps, alpha, lam)
        d_grads = {k: O.fingerprint(p.grad) for k, p in critic.named_parameters()}
        # generator step
        for p in critic.parameters():
            p.requires_grad = False
        for p in gen.parameters():
            p.requires_grad = True
        z2 = z_g.clone().requires_grad_()
        fake2 = gen(z2, noise=n_g, steps=steps, alpha=alpha)
        pred = critic(fake2, steps, alpha)
        g_loss = gen.get_r1_loss(pred)
        gen.zero_grad()
        g_loss.backward()
        g_grads = {k: O.fingerprint(p.grad) for k, p in gen.named_parameters()}
        train.append({"steps": steps, "batch": batch, "alpha": alpha, "lambda": lam,
                      "c_loss": c_loss.item(), "g_loss": g_loss.item(),
                      "pred_fake": O.fingerprint(pf), "pred_real": O.fingerprint(pr),
                      "fake_d": O.fingerprint(fake), "z_grad": O.fingerprint(z2.grad),
                      "d_grads": d_grads, "g_grads": g_grads})
        print(f"train steps={steps} B={batch} alpha={alpha}: c_loss={c_loss.item():.6f} g_loss={g_loss.item():.6f} "
              f"{time.time() - t0:.1f}s", flush=True)
    json.dump(train, open(os.path.join(OUT, "train_iteration.json"), "w"))

    # ---- R1 penalty alone: gan.py:398-404 on the real modules, then backward of just that term
    pen = []
    for steps, batch, alpha, lam in PENALTY_CASES:
        gen, critic = load_ref(seed=2)
        real_im = O.make_images(batch, steps, seed=30 + steps).requires_grad_()
        pr = critic(real_im, steps, alpha)
        critic.zero_grad()
        grad_real = torch.autograd.grad(outputs=pr.sum(), inputs=real_im, create_graph=True)[0]
        penalty = lam / 2 * (grad_real.view(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()
        penalty.backward()
        pen.append({"steps": steps, "batch": batch, "alpha": alpha, "lambda": lam, "penalty": penalty.item(),
                    "grad_real": O.fingerprint(grad_real),
                    "d_grads": {k: O.fingerprint(p.grad) for k, p in critic.named_parameters()}})
        print(f"penalty steps={steps} B={batch} alpha={alpha}: {penalty.item():.6f}", flush=True)
    json.dump(pen, open(os.path.join(OUT, "r1_penalty.json"), "w"))

    # ---- WGAN-GP: the body of Critic.get_wgan_loss (gan.py:357-391) executed on the real modules with the two
    # defects that keep it from running repaired (self.device -> real_im.device; the undefined fake_im passed in) and
    # epsilon supplied instead of torch.rand; generator step with Generator.get_wgan_loss (gan.py:224-225)
    wg = []
    for steps, batch, alpha, lam in WGAN_CASES:
        gen, critic = load_ref(seed=2)
        z_d, z_g = O.make_latents(batch, seed=10 + steps), O.make_latents(batch, seed=20 + steps)
        n_d, n_g = O.make_noise(batch, steps, seed=10 + steps), O.make_noise(batch, steps, seed=20 + steps)
        real = O.make_images(batch, steps, seed=30 + steps)
        eps = O.make_epsilon(batch, seed=40 + steps)
        for p in gen.parameters():
            p.requires_grad = False
        fake = gen(z_d, noise=n_d, steps=steps, alpha=alpha)
        real_im = real.clone().requires_grad_()
        pf = critic(fake.detach(), steps, alpha)
        pr = critic(real_im, steps, alpha)
        critic.zero_grad()
        mixed_images = real_im * eps + (1 - eps) * fake                                    # gan.py:372
        mixed_image_scores = critic.forward(mixed_images, steps=steps, alpha=alpha)        # gan.py:373
        gradient = torch.autograd.grad(inputs=mixed_images, outputs=mixed_image_scores,
                                       grad_outputs=torch.ones_like(mixed_image_scores),
                                       create_graph=True, retain_graph=True)[0]            # gan.py:375-381
        gp = ((gradient.view(gradient.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()        # gan.py:385
        wgan_loss = -pr.mean() + pf.mean() + (lam * gp)                                    # gan.py:387
        wgan_loss.backward()                                                               # gan.py:389
        d_grads = {k: O.fingerprint(p.grad) for k, p in critic.named_parameters()}
        for p in critic.parameters():
            p.requires_grad = False
        for p in gen.parameters():
            p.requires_grad = True
        fake2 = gen(z_g, noise=n_g, steps=steps, alpha=alpha)
        g_loss = gen.get_wgan_loss(critic(fake2, steps, alpha))
        gen.zero_grad()
        g_loss.backward()
        wg.append({"steps": steps, "batch": batch, "alpha": alpha, "lambda": lam, "c_loss": wgan_loss.item(),
                   "g_loss": g_loss.item(), "gp": gp.item(), "d_grads": d_grads,
                   "g_grads": {k: O.fingerprint(p.grad) for k, p in gen.named_parameters()}})
        print(f"wgan steps={steps} B={batch} alpha={alpha}: c_loss={wgan_loss.item():.6f} gp={gp.item():.6f}", flush=True)
    json.dump(wg, open(os.path.join(OUT, "wgan_gp.json"), "w"))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
