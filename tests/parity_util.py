"""Shared helpers for the GPU parity tests: build the CUDA modules and the oracle from the same seeded
state, run one G+D iteration through the public API exactly as train.py:135-217 drives it."""
import torch

import gan
from oracle import gan_oracle as O

# Stated tolerances.  The CUDA path keeps feature maps, their gradients and the conv weights in bf16 (fp32
# accumulation, fp32 statistics); the reference is fp32 throughout.  Forward quantities agree to ~1e-2.
# Gradients are noisier than the 2^-9 storage error suggests because the network is piecewise linear: a
# rounding-induced flip of a LeakyReLU gate (slope 1 <-> 0.2) changes that element's gradient by 80 %, and
# the G step chains ~50 such layers (G fwd, D fwd, D bwd, G bwd).  The yardstick for "this is storage noise,
# not wrong arithmetic" is the oracle itself run with bf16-rounded storage (gan_oracle.QUANT): the CUDA path
# must be no further from the fp32 reference than ~1.5x that emulation, tensor population by tensor population.
TOL_IMG = 4e-2        # rel-L2 of generated images
TOL_PRED = 5e-2       # critic scores: |diff| <= TOL_PRED * (rms(pred) + 1)
TOL_LOSS = 3e-2       # relative, losses
# Per-tensor hard cap.  The CUDA path is not bit-reproducible (fp32 atomics in the fused reductions and split-K
# sums), and the tiny heavily-cancelling gradients (a noise-weight gradient is 16..512 numbers, each a signed sum over
# every pixel) move a lot between runs: tools/flaky_probe.py measured 0.23..0.41 (per-layer kernels) and 0.39..0.60
# (fused style-conv forward, whose roundings differ from the backward's linearisation at four more layers) for
# gen_blocks.5.conv_2.inject_noise.weights at 128x128, batch 4, where the deterministic bf16 emulation of the reference
# sits at 0.21 and every other tensor stays within ~1.3x of its emulated error.  The cap catches wrong arithmetic
# (errors >= 1), the median criterion below is the accuracy statement.
TOL_GRAD_REL = 0.85   # per-tensor rel-L2 of gradients vs the fp32 reference (hard cap)
TOL_GRAD_COS = 0.9    # per-tensor cosine of gradients vs the fp32 reference (hard cap)
TOL_VS_EMU = 1.5      # median rel-L2 over a network's tensors <= TOL_VS_EMU * same statistic of the bf16 emulation + 0.01


def no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def build_models(seed, device="cuda"):
    g, c = gan.Generator(), gan.Critic()
    g.load_state_dict(O.make_state("gen", seed))
    c.load_state_dict(O.make_state("critic", seed))
    return g.to(device), c.to(device)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def cuda_iteration(g, c, z_d, z_g, real, n_d, n_g, steps, alpha, lam):
    """train.py:135-217 with the optimizer steps removed (so the D and G gradients refer to the same weights)."""
    dev = "cuda"
    out = {}
    for p in c.parameters():
        p.requires_grad = True
    for p in g.parameters():
        p.requires_grad = False
    z = z_d.to(dev).requires_grad_()
    fake = g(z, noise=[n.to(dev) for n in n_d], steps=steps, alpha=alpha)
    real_im = real.to(dev).requires_grad_()
    pf = c(fake.detach(), steps, alpha)
    pr = c(real_im, steps, alpha)
    c.zero_grad()
    c_loss = c.get_r1_loss(pf, pr, real_im, fake, steps, alpha, lam)
    out.update(c_loss=c_loss.detach(), fake_d=fake.detach(), pred_fake=pf.detach(), pred_real=pr.detach(),
               d_grads={k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in c.named_parameters()},
               real_grad=c.last_real_image_grad)
    for p in c.parameters():
        p.requires_grad = False
    for p in g.parameters():
        p.requires_grad = True
    z2 = z_g.to(dev).requires_grad_()
    fake2 = g(z2, noise=[n.to(dev) for n in n_g], steps=steps, alpha=alpha)
    pred = c(fake2, steps, alpha)
    g_loss = g.get_r1_loss(pred)
    g.zero_grad()
    g_loss.backward()
    out.update(g_loss=g_loss.detach(), pred_g=pred.detach(), z_grad=z2.grad.detach(),
               g_grads={k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in g.named_parameters()})
    return out
