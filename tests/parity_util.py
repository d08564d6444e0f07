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
# Gradient gate, PER TENSOR (grad_gate below).  The yardstick of every criterion is the error that the reference's OWN
# bf16-storage emulation (gan_oracle.QUANT) makes on that very tensor: profiles/r2_parity_table.txt lists, for every
# parameter gradient of every golden case, the CUDA path's and the emulation's rel-L2 / cosine side by side — the CUDA
# path sits at 0.6-1.6x the emulated error on all but a handful of tiny tensors (worst: a 64-number noise-weight
# gradient at 2.4x).  A tensor with a lost or mis-scaled term (a missing R1 second-order contribution, a dropped
# minibatch-stddev curvature term, a wrong constant) is off by O(its share) and lands far outside these bands:
#   * rel-L2 vs the fp32 reference  <= K_EMU x emulated rel-L2 + EPS_EMU   (K_EMU_SMALL for tensors of < 4096 numbers:
#     per-channel sums over every pixel — biases, noise weights, toRGB — cancel heavily);
#   * magnitude: | |g|/|ref| - 1 |  <= 0.2 + 1.5 x emulated rel-L2  (a constant factor has cosine 1.0);
#   * direction: 1 - cosine         <= 2.5 x (1 - emulated cosine) + 0.02.
# The tests run the library in chain-deterministic mode (bg_set_deterministic): images, scores and activation
# gradients are then bit-reproducible, so the gate sees ONE fixed realisation of the bf16 rounding noise.
K_EMU = 2.0
K_EMU_SMALL = 3.0
EPS_EMU = 0.02
TOL_VS_EMU = 1.5      # median rel-L2 over a network's tensors <= TOL_VS_EMU * same statistic of the bf16 emulation + 0.01


def grad_gate(name, got, ref, emu):
    """None if the gradient tensor passes the per-tensor gate, else a description of the failure."""
    if ref.norm().item() == 0.0:
        return None if got.abs().max().item() < 1e-6 else f"{name}: reference gradient is 0, got {got.abs().max().item():.2e}"
    e, e_emu, cs, cs_emu = rel(got, ref), rel(emu, ref), cos(got, ref), cos(emu, ref)
    k = K_EMU_SMALL if ref.numel() < 4096 else K_EMU
    ratio = got.double().norm().item() / ref.double().norm().item()
    why = []
    if e > k * e_emu + EPS_EMU:
        why.append(f"rel-L2 {e:.4f} > {k} x emulation {e_emu:.4f} + {EPS_EMU}")
    if abs(ratio - 1.0) > 0.2 + 1.5 * e_emu:
        why.append(f"norm ratio {ratio:.3f} (emulated rel-L2 {e_emu:.3f})")
    if ref.numel() >= 2 and 1.0 - cs > 2.5 * (1.0 - cs_emu) + 0.02:
        why.append(f"cosine {cs:.4f} (emulation {cs_emu:.4f})")
    return f"{name} [{ref.numel()}]: " + ", ".join(why) if why else None


class deterministic:
    """with U.deterministic(): ... — chain-deterministic reductions for the duration (include/bg_b200.h)."""

    def __enter__(self):
        import bg_native

        self.prev = bg_native.set_deterministic(True)

    def __exit__(self, *exc):
        import bg_native

        bg_native.set_deterministic(self.prev)


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


def cuda_iteration(g, c, z_d, z_g, real, n_d, n_g, steps, alpha, lam, loss="r1", epsilon=None):
    """train.py:135-217 with the optimizer steps removed (so the D and G gradients refer to the same weights).
    loss: "r1" (use_r1=True), "r1_penalty" (the R1 penalty term alone) or "wgan" (train.py:177-185,213)."""
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
    if loss == "wgan":
        c_loss = c.get_wgan_loss(pf, pr, real_im, steps, alpha, lam, epsilon=epsilon.to(dev))
    else:
        c._r1_penalty_only = loss == "r1_penalty"
        try:
            c_loss = c.get_r1_loss(pf, pr, real_im, fake, steps, alpha, lam)
        finally:
            c._r1_penalty_only = False
    out.update(c_loss=c_loss.detach(), fake_d=fake.detach(), pred_fake=pf.detach(), pred_real=pr.detach(),
               d_grads={k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in c.named_parameters()},
               real_grad=c.last_mixed_image_grad if loss == "wgan" else c.last_real_image_grad)
    for p in c.parameters():
        p.requires_grad = False
    for p in g.parameters():
        p.requires_grad = True
    z2 = z_g.to(dev).requires_grad_()
    fake2 = g(z2, noise=[n.to(dev) for n in n_g], steps=steps, alpha=alpha)
    pred = c(fake2, steps, alpha)
    g_loss = g.get_wgan_loss(pred) if loss == "wgan" else g.get_r1_loss(pred)
    g.zero_grad()
    g_loss.backward()
    out.update(g_loss=g_loss.detach(), pred_g=pred.detach(), z_grad=z2.grad.detach(),
               g_grads={k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in g.named_parameters()})
    return out
