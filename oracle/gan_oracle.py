"""CPU/torch-fp32 oracle for the BYO-GAN StyleGAN hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a reference-layout ``state_dict``, the arithmetic of the
reference's ``gan.py`` (``/root/reference/gan.py``; every function cites the lines it follows) and of the
G+D iteration in ``train.py:135-217``.  It exists to CHECK the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  The product (``byo-gan_b200/gan.py``) never does and has no CPU fallback.

Pinning: the reference ships no tests, golden vectors or checkpoints, so parity is pinned
against outputs of the reference itself: ``oracle/make_golden.py`` imports the unmodified reference
``gan.py`` in the build container, runs it on the deterministic inputs produced by :func:`make_state`,
:func:`make_latents` ... below, and commits compact fingerprints under ``tests/golden/``.
``tests/test_oracle_golden.py`` holds this restatement to those fixtures.

Everything is fp32 and NCHW like the reference.  Gradients come from torch autograd over these functions,
which is exactly how the reference obtains them (``gan.py:398-410``, ``train.py:216``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# Channel plans (gan.py:159-166 generator blocks, gan.py:320-327 critic blocks, gan.py:307-314 fromRGB).
GEN_CHANNELS = [(512, 512), (512, 512), (512, 512), (512, 256), (256, 128), (128, 64), (64, 32), (32, 16)]
CRITIC_CHANNELS = [(16, 32), (32, 64), (64, 128), (128, 256), (256, 512), (512, 512), (512, 512), (512, 512)]
Z_DIM = 512
NUM_BLOCKS = 8
LRELU_SLOPE = 0.2
IN_EPS = 1e-8      # nn.InstanceNorm2d(eps=1e-8), gan.py:59
MBSTD_EPS = 1e-8   # gan.py:287


# --------------------------------------------------------------------------------------------------------
# deterministic inputs (shared by the golden generator, the tests, smoke() and bench.py)
# --------------------------------------------------------------------------------------------------------
def generator_param_shapes() -> Dict[str, tuple]:
    """state_dict keys/shapes of the reference Generator (probed from gan.py:151-181; 111 tensors)."""
    shapes: Dict[str, tuple] = {}
    for i in range(8):  # MappingLayers, gan.py:130-145
        shapes[f"to_w_noise.0.layers.{i}.0.weight"] = (512, 512)
        shapes[f"to_w_noise.0.layers.{i}.0.bias"] = (512,)
    for k, (cin, cout) in enumerate(GEN_CHANNELS):
        for j, ci in ((1, cin), (2, cout)):
            p = f"gen_blocks.{k}.conv_{j}"
            if k == 0 and j == 1:
                shapes[f"{p}.conv"] = (1, cin, 4, 4)  # learned constant, gan.py:81
            else:
                shapes[f"{p}.conv.weight"] = (cout, ci, 3, 3)
                shapes[f"{p}.conv.bias"] = (cout,)
            shapes[f"{p}.inject_noise.weights"] = (1, cout, 1, 1)
            shapes[f"{p}.adain.style.weight"] = (2 * cout, 512)
            shapes[f"{p}.adain.style.bias"] = (2 * cout,)
    for k, (_, cout) in enumerate(GEN_CHANNELS):
        shapes[f"to_rgbs.{k}.weight"] = (3, cout, 1, 1)
        shapes[f"to_rgbs.{k}.bias"] = (3,)
    return shapes


def critic_param_shapes() -> Dict[str, tuple]:
    """state_dict keys/shapes of the reference Critic (gan.py:301-329, 231-262; 52 tensors)."""
    shapes: Dict[str, tuple] = {}
    for k, (cin, _) in enumerate(CRITIC_CHANNELS):
        shapes[f"from_rgbs.{k}.0.weight"] = (cin, 3, 1, 1)
        shapes[f"from_rgbs.{k}.0.bias"] = (cin,)
    for k, (cin, cout) in enumerate(CRITIC_CHANNELS):
        p = f"conv_blocks.{k}"
        if k < 7:
            shapes[f"{p}.conv_1.0.weight"] = (cout, cin, 3, 3)
            shapes[f"{p}.conv_1.0.bias"] = (cout,)
            shapes[f"{p}.conv_2.0.weight"] = (cout, cout, 3, 3)
            shapes[f"{p}.conv_2.0.bias"] = (cout,)
        else:
            shapes[f"{p}.conv_1.1.weight"] = (cout, cin + 1, 3, 3)
            shapes[f"{p}.conv_1.1.bias"] = (cout,)
            shapes[f"{p}.conv_2.0.weight"] = (cout, cout, 4, 4)
            shapes[f"{p}.conv_2.0.bias"] = (cout,)
            shapes[f"{p}.conv_2.3.weight"] = (cout, cout)
            shapes[f"{p}.conv_2.3.bias"] = (cout,)
            shapes[f"{p}.conv_2.5.weight"] = (1, cout)
            shapes[f"{p}.conv_2.5.bias"] = (1,)
    return shapes


def make_state(kind: str, seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic parameters in the reference's state_dict layout (CPU fp32).

    Weights ~ N(0,1) as the reference initialises them (gan.py:10,23,81).  Unlike the reference's init,
    biases and noise-injection weights are N(0, 0.1^2) instead of 0 (gan.py:11,24,44) and the AdaIN style
    bias is [1..|0..] + N(0, 0.1^2) (gan.py:62-63), so every additive path is live in the parity tests
    (SURVEY.md §8d).  Values depend only on (kind, seed, key order) and the torch CPU generator.
    """
    shapes = generator_param_shapes() if kind == "gen" else critic_param_shapes()
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + (17 if kind == "gen" else 29))
    state = {}
    for key, shape in shapes.items():
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if key.endswith("bias") or key.endswith("inject_noise.weights"):
            t = t * 0.1
            if key.endswith("adain.style.bias"):
                t[: shape[0] // 2] += 1.0
        state[key] = t
    return state


def make_latents(batch: int, seed: int, trunc: float = 0.75) -> torch.Tensor:
    """Truncated-normal latents like helper.get_truncated_noise (helper.py:36-45), by rejection from the
    seeded torch CPU generator instead of scipy so the values are reproducible on any box."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    z = torch.randn(batch, Z_DIM, generator=g)
    bad = z.abs() > trunc
    while bad.any():
        z = torch.where(bad, torch.randn(batch, Z_DIM, generator=g), z)
        bad = z.abs() > trunc
    return z


def make_noise(batch: int, steps: int, seed: int) -> List[torch.Tensor]:
    """Explicit per-block noise maps (B,1,4*2^i,4*2^i), the `noise=` argument of Generator.forward
    (gan.py:183,193-197)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(104729 * seed + 11)
    return [torch.randn(batch, 1, 4 * 2 ** i, 4 * 2 ** i, generator=g) for i in range(steps)]


def make_images(batch: int, steps: int, seed: int) -> torch.Tensor:
    """Synthetic "real" images U(-1,1), (B,3,R,R) (the range transforms.Normalize(.5,.5) yields, train.py:47)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(15485863 * seed + 5)
    r = 4 * 2 ** (steps - 1)
    return torch.rand(batch, 3, r, r, generator=g) * 2 - 1


# --------------------------------------------------------------------------------------------------------
# optional bf16 emulation: with QUANT[0] = True the oracle rounds conv weights and every feature map that
# the CUDA path stores in bf16 (forward value AND its gradient) to bf16.  This is the yardstick that separates
# "bf16 storage noise" from "wrong arithmetic" in the GPU parity tests; the fp32 mode is the reference.
# --------------------------------------------------------------------------------------------------------
QUANT = [False]


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def q(x):
    return _RoundBf16.apply(x) if QUANT[0] else x


def qw(w):
    """weights are rounded once (value only): the master copy and its gradient stay fp32"""
    return (w.to(torch.bfloat16).to(w.dtype) - w).detach() + w if QUANT[0] else w


# --------------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------------
def eq_coef(weight: torch.Tensor) -> float:
    """sqrt(2 / fan_in), fan_in = in_features * kernel area (gan.py:13-14, 26-27)."""
    fan_in = weight.shape[1] * (weight[0][0].numel())
    return math.sqrt(2.0 / fan_in)


def eq_linear(x, weight, bias):
    """EqualizedLinear.forward, gan.py:16-17: the runtime scale multiplies the weight only."""
    return F.linear(x, weight * eq_coef(weight), bias)


def eq_conv2d(x, weight, bias, padding=0):
    """EqualizedConv2d.forward, gan.py:29-38 (stride 1, no dilation/groups anywhere in the model)."""
    return F.conv2d(x, qw(weight * eq_coef(weight)) if weight.shape[-1] == 3 else weight * eq_coef(weight), bias,
                    padding=padding)


def lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


def bilinear_up2(x):
    """nn.Upsample(scale_factor=2, mode='bilinear') / F.interpolate(..., 'bilinear'), align_corners=False
    (gan.py:112, 213-217): separable taps (.25,.75)/(.75,.25) with the source index clamped at the edges."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def instance_norm(x):
    """nn.InstanceNorm2d(C, eps=1e-8): no affine, no running stats, biased variance (gan.py:59)."""
    # (x - mean_hw) / sqrt(biased_var_hw + eps); evaluated with the same ATen primitive the reference's
    # nn.InstanceNorm2d dispatches to, so fp32 rounding matches it (the 4x4 planes are ill-conditioned).
    return F.instance_norm(x, eps=IN_EPS)


def style_conv(P, prefix, x, w_lat, noise, batch, initial=False):
    """StyleConvBlock.forward, gan.py:89-98: conv|const -> + noise_w*noise -> LeakyReLU -> AdaIN."""
    if initial:
        out = P[f"{prefix}.conv"].repeat(batch, 1, 1, 1)            # gan.py:92
    else:
        out = eq_conv2d(x, P[f"{prefix}.conv.weight"], P[f"{prefix}.conv.bias"], padding=1)  # gan.py:94
    out = out + P[f"{prefix}.inject_noise.weights"] * noise        # gan.py:52
    out = q(lrelu(out))                                            # gan.py:97
    style = eq_linear(w_lat, P[f"{prefix}.adain.style.weight"], P[f"{prefix}.adain.style.bias"])  # gan.py:66
    c = out.shape[1]
    gamma = style[:, :c, None, None]
    beta = style[:, c:, None, None]
    return q(gamma * instance_norm(out) + beta)                    # gan.py:69


def synthesis_block(P, k, x, w_lat, noise, batch):
    """StyleGanBlock.forward, gan.py:118-127; both convs receive the same noise map."""
    if k > 0:
        x = q(bilinear_up2(x))
    out = style_conv(P, f"gen_blocks.{k}.conv_1", x, w_lat, noise, batch, initial=(k == 0))
    return style_conv(P, f"gen_blocks.{k}.conv_2", out, w_lat, noise, batch)


def mapping(P, z):
    """MappingLayers.forward, gan.py:130-148: 8 x (EqualizedLinear 512->512, LeakyReLU 0.2); no pixel norm."""
    x = z
    for i in range(8):
        x = lrelu(eq_linear(x, P[f"to_w_noise.0.layers.{i}.0.weight"], P[f"to_w_noise.0.layers.{i}.0.bias"]))
    return x


def to_rgb(P, k, x):
    return eq_conv2d(x, P[f"to_rgbs.{k}.weight"], P[f"to_rgbs.{k}.bias"])


def generator_forward(P, z, noise: Optional[Sequence[torch.Tensor]] = None, steps: int = 1,
                      alpha: Optional[float] = None, w_override: Optional[Sequence[torch.Tensor]] = None):
    """Generator.forward, gan.py:183-222.

    ``w_override`` (one latent-w per block) is the oracle of the opt-in style-mixing extension: it drives
    the reference's own sub-blocks with a per-block style, exactly as SURVEY.md §7(9) defines it.  With the
    default ``None`` every block sees ``mapping(z)`` as in the reference.
    """
    batch = len(z)
    w_lat = mapping(P, z)
    if noise is None:
        noise = [torch.randn(batch, 1, 4 * 2 ** i, 4 * 2 ** i, device=z.device) for i in range(steps)]
    out = noise[0]
    for k in range(NUM_BLOCKS):
        previous = out
        wk = w_lat if w_override is None else w_override[k]
        out = synthesis_block(P, k, out, wk, noise[k], batch)
        if k + 1 >= steps:
            if alpha is not None and k > 0:
                a = min(1.0, max(0.0, alpha))                                  # gan.py:211
                small = bilinear_up2(to_rgb(P, k - 1, previous))              # gan.py:213-217
                return torch.lerp(small, to_rgb(P, k, out), a)                # gan.py:220
            return to_rgb(P, k, out)                                           # gan.py:222
    return None


def minibatch_stddev(x, group_size: int = 4):
    """MiniBatchStdDev.forward, gan.py:273-298.  Returns (output, group_size actually used).

    Quirks kept: the mean subtracted is over the WHOLE batch (gan.py:282), groups are strided
    (sample n sits in slot n mod (B/G)), and group_size falls back to B when B % group_size != 0
    (the module then keeps that value, gan.py:277-278 — the caller threads the returned value back in).
    """
    b, c, h, w = x.shape
    if b % group_size != 0:
        group_size = b
    m = b // group_size
    grouped = x.reshape(group_size, m, c, h, w)
    dev = grouped - x.mean(0, keepdim=True)                # broadcast of the whole-batch mean
    var = (dev ** 2).mean(0)                               # (m, c, h, w)
    s = torch.sqrt(var + MBSTD_EPS).mean(dim=(1, 2, 3))    # (m,)
    plane = s.repeat(group_size)                           # sample n -> s[n mod m]
    plane = plane.view(b, 1, 1, 1).expand(b, 1, h, w)
    return torch.cat([x, plane], dim=1), group_size


def from_rgb(P, k, img):
    """gen_from_rgbs, gan.py:351-355: 1x1 EqualizedConv2d + LeakyReLU."""
    return q(lrelu(eq_conv2d(img, P[f"from_rgbs.{k}.0.weight"], P[f"from_rgbs.{k}.0.bias"])))


def critic_block(P, k, x, group_size: int = 4):
    """CriticBlock.forward, gan.py:264-265 with the layer lists of gan.py:237-262."""
    p = f"conv_blocks.{k}"
    if k < 7:
        x = q(lrelu(eq_conv2d(x, P[f"{p}.conv_1.0.weight"], P[f"{p}.conv_1.0.bias"], padding=1)))
        x = q(eq_conv2d(x, P[f"{p}.conv_2.0.weight"], P[f"{p}.conv_2.0.bias"], padding=1))
        return q(lrelu(F.avg_pool2d(x, 2))), group_size    # pool BEFORE the activation, gan.py:258-262
    x, group_size = minibatch_stddev(x, group_size)
    x = q(lrelu(eq_conv2d(q(x), P[f"{p}.conv_1.1.weight"], P[f"{p}.conv_1.1.bias"], padding=1)))
    x = lrelu(eq_conv2d(x, P[f"{p}.conv_2.0.weight"], P[f"{p}.conv_2.0.bias"]))   # 4x4 valid -> (B,512,1,1)
    x = x.flatten(1)
    x = lrelu(eq_linear(x, P[f"{p}.conv_2.3.weight"], P[f"{p}.conv_2.3.bias"]))
    return eq_linear(x, P[f"{p}.conv_2.5.weight"], P[f"{p}.conv_2.5.bias"]), group_size


def critic_forward(P, images, steps: int = 1, alpha: Optional[float] = None, group_size: int = 4,
                   return_group_size: bool = False):
    """Critic.forward, gan.py:331-349."""
    start = NUM_BLOCKS - steps
    out = None
    for j, k in enumerate(range(start, NUM_BLOCKS)):
        if j == 0:
            out = from_rgb(P, start, images)
        out, group_size = critic_block(P, k, out, group_size)
        if j == 0 and steps > 1 and alpha is not None:
            a = min(1.0, max(0.0, alpha))
            down = from_rgb(P, start + 1, F.avg_pool2d(images, 2))     # gan.py:345
            out = q(torch.lerp(down, out, a))                           # gan.py:347
    return (out, group_size) if return_group_size else out


# --------------------------------------------------------------------------------------------------------
# losses and the G+D iteration
# --------------------------------------------------------------------------------------------------------
def generator_r1_loss(pred_fake):
    """Generator.get_r1_loss, gan.py:227-228."""
    return F.softplus(-pred_fake).mean()


def critic_r1_loss(pred_fake, pred_real, real_im, c_lambda=1.0):
    """Critic.get_r1_loss, gan.py:393-412, WITHOUT the internal .backward() (the caller does it):
    softplus(-D(real)).mean() + softplus(D(fake)).mean() + lambda/2 * mean_n ||d sum(D(real)) / d real_n||^2."""
    real_term = F.softplus(-pred_real).mean()
    (grad_real,) = torch.autograd.grad(outputs=pred_real.sum(), inputs=real_im, create_graph=True)
    penalty = (grad_real.reshape(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()
    fake_term = F.softplus(pred_fake).mean()
    return real_term + fake_term + c_lambda / 2 * penalty


def critic_r1_penalty(pred_real, real_im, c_lambda=1.0):
    """Only the gradient-penalty term of Critic.get_r1_loss (gan.py:398-404): lambda/2 * mean_n ||d sum D(real) / d real_n||^2.
    Its parameter gradient is PURELY second order, so checking it alone isolates the double-backward of every layer."""
    (grad_real,) = torch.autograd.grad(outputs=pred_real.sum(), inputs=real_im, create_graph=True)
    return c_lambda / 2 * (grad_real.reshape(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()


def make_epsilon(batch: int, seed: int) -> torch.Tensor:
    """The per-sample interpolation factor of WGAN-GP, U(0,1) of shape (B,1,1,1) (gan.py:367-369), from a seeded
    CPU generator so the reference restatement, the oracle and the CUDA path mix the same images."""
    g = torch.Generator(device="cpu")
    g.manual_seed(32452843 * seed + 7)
    return torch.rand(batch, 1, 1, 1, generator=g)


def generator_wgan_loss(pred_fake):
    """Generator.get_wgan_loss, gan.py:224-225."""
    return -pred_fake.mean()


def critic_wgan_gp_loss(D, pred_fake, pred_real, real_im, fake_im, epsilon, steps, alpha, c_lambda=1.0, group_size=4):
    """Critic.get_wgan_loss as gan.py:357-391 INTENDS it, WITHOUT the internal .backward().  The reference's body cannot
    run: it reads `self.device` (nn.Module has none, gan.py:368) and an undefined `fake_im` (gan.py:372; train.py:178-185
    does not pass it).  Repaired here and in oracle/make_golden.py the same way: device taken from real_im, fake_im
    passed in, epsilon supplied (the reference draws torch.rand).  Everything else follows the lines:
      mixed = real * eps + (1 - eps) * fake                      gan.py:372
      g = d sum D(mixed) / d mixed  (create_graph)               gan.py:373-381
      gp = mean_n (||g_n||_2 - 1)^2                              gan.py:385
      loss = -mean D(real) + mean D(fake) + lambda * gp          gan.py:387"""
    mixed = real_im * epsilon + (1 - epsilon) * fake_im
    scores = critic_forward(D, mixed, steps, alpha, group_size)
    (gradient,) = torch.autograd.grad(inputs=mixed, outputs=scores, grad_outputs=torch.ones_like(scores),
                                      create_graph=True, retain_graph=True)
    gp = ((gradient.reshape(gradient.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()
    return -pred_real.mean() + pred_fake.mean() + c_lambda * gp


def _as_params(state, requires_grad, device=None, dtype=torch.float32):
    return {k: v.detach().to(device=device, dtype=dtype).clone().requires_grad_(requires_grad)
            for k, v in state.items()}


def train_iteration(gen_state, critic_state, z_d, z_g, real, noise_d, noise_g, steps, alpha, c_lambda=10.0,
                    device=None, dtype=torch.float32, loss="r1", epsilon=None):
    """One reference iteration (train.py:135-217) without the optimizer steps: the critic step
    (train.py:135-191: G frozen, fake detached, R1 loss, backward into D) then the generator step
    (train.py:193-217: D frozen, non-saturating loss, backward into G).  Both steps see the SAME weights
    (no Adam update in between) so gradients are comparable tensor by tensor.

    loss: "r1" (use_r1=True, the reference's working path), "r1_penalty" (critic step = the penalty term alone: second-order
    gradients in isolation) or "wgan" (train.py:177-185,213 with the repaired gan.py:357-391; needs `epsilon` (B,1,1,1)).

    Returns dict(c_loss, g_loss, fake_d, pred_fake, pred_real, d_grads{key: tensor|None}, g_grads{...}).
    """
    out = {}
    # ---- critic step
    G = _as_params(gen_state, False, device, dtype)
    D = _as_params(critic_state, True, device, dtype)
    z = z_d.detach().to(device=device, dtype=dtype).requires_grad_()          # helper.py:44
    fake = generator_forward(G, z, [n.to(device=device, dtype=dtype) for n in noise_d], steps, alpha)
    real_im = real.detach().to(device=device, dtype=dtype).requires_grad_()   # train.py:150-158
    pred_fake = critic_forward(D, fake.detach(), steps, alpha)
    pred_real = critic_forward(D, real_im, steps, alpha)
    if loss == "r1":
        c_loss = critic_r1_loss(pred_fake, pred_real, real_im, c_lambda)
    elif loss == "r1_penalty":
        c_loss = critic_r1_penalty(pred_real, real_im, c_lambda)
    else:
        c_loss = critic_wgan_gp_loss(D, pred_fake, pred_real, real_im, fake.detach(),
                                     epsilon.detach().to(device=device, dtype=dtype), steps, alpha, c_lambda)
    c_loss.backward()                                                          # gan.py:410 / gan.py:389
    out.update(c_loss=c_loss.detach(), fake_d=fake.detach(), pred_fake=pred_fake.detach(),
               pred_real=pred_real.detach(),
               d_grads={k: (v.grad.detach() if v.grad is not None else None) for k, v in D.items()})
    # ---- generator step
    G = _as_params(gen_state, True, device, dtype)
    D = _as_params(critic_state, False, device, dtype)
    z = z_g.detach().to(device=device, dtype=dtype).requires_grad_()
    fake = generator_forward(G, z, [n.to(device=device, dtype=dtype) for n in noise_g], steps, alpha)
    pred = critic_forward(D, fake, steps, alpha)
    g_loss = generator_wgan_loss(pred) if loss == "wgan" else generator_r1_loss(pred)   # train.py:207-213
    g_loss.backward()                                                          # train.py:216
    out.update(g_loss=g_loss.detach(), fake_g=fake.detach(), pred_g=pred.detach(),
               g_grads={k: (v.grad.detach() if v.grad is not None else None) for k, v in G.items()},
               z_grad=z.grad.detach())
    return out


# --------------------------------------------------------------------------------------------------------
# compact fingerprints (what the golden fixtures store instead of full tensors)
# --------------------------------------------------------------------------------------------------------
def fingerprint(t: Optional[torch.Tensor], samples: int = 64) -> Optional[dict]:
    """Size-independent summary of a tensor: shape, sum, L2 norm, abs-max and `samples` evenly strided
    values.  Two tensors with equal fingerprints to 1e-5 are, for parity purposes, the same tensor."""
    if t is None:
        return None
    f = t.detach().to("cpu", torch.float64).flatten()
    n = f.numel()
    idx = torch.linspace(0, n - 1, min(samples, n)).round().long()
    return {"shape": list(t.shape), "sum": f.sum().item(), "norm": f.norm().item(),
            "absmax": f.abs().max().item() if n else 0.0, "samples": f[idx].to(torch.float32).tolist()}
