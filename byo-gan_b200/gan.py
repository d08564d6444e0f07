"""Drop-in replacement for the reference's ``gan.py`` (BYO-GAN), running on hand-written sm_100a kernels.

Import it as ``gan`` (put ``byo-gan_b200/`` first on ``sys.path``): ``main.py`` / ``train.py`` /
``generate_samples.py`` / ``interpolate.py`` of the reference then run unchanged, because this module keeps

* the constructors ``Generator()`` / ``Critic()`` (no arguments; reference gan.py:151-181, 301-329),
* the attribute tree and therefore the ``state_dict`` keys/shapes of a reference checkpoint
  (``to_w_noise.0.layers.i.0.weight`` ... ``conv_blocks.7.conv_2.5.bias``; 111 + 52 fp32 tensors),
* ``Generator.forward(z_noise, noise=None, steps=1, alpha=None)`` (gan.py:183-222),
  ``Critic.forward(images, steps=1, alpha=None)`` (gan.py:331-349),
* the loss methods ``get_r1_loss`` / ``get_wgan_loss`` with the reference's signatures; like the reference,
  ``Critic.get_r1_loss`` runs its own backward and fills ``.grad`` (gan.py:393-412).

Everything between those calls is different: feature maps live in NHWC bf16, every convolution is a
tcgen05 implicit GEMM, and forward, backward and the R1 double-backward are scheduled by ``engine.py`` over
the C ABI of ``libbg_b200.so`` (``include/bg_b200.h``).  The sub-modules below only HOLD parameters (so
optimizers, ``requires_grad`` toggling, ``zero_grad`` and checkpoints work as in the reference); they have
no forward of their own.  There is no CPU path: calling the model with CPU tensors raises.

Opt-in extension (not in the reference, default off): style mixing — ``Generator.forward(..., z2=, crossover=)``
feeds blocks ``>= crossover`` with the mapped second latent.
"""
from math import sqrt

import torch
from torch import nn

import os

import engine
import lazy
from engine import PackCache

# Opt-in host-overhead shims for an unmodified train.py (lazy.py); both default off
DEFER_LOSS_ITEMS = os.environ.get("BG_DEFER_ITEMS", "0") == "1"        # loss.item() returns a float-like that syncs on use
LAZY_NO_GRAD_FORWARD = os.environ.get("BG_LAZY_PREVIEW", "0") == "1"   # no_grad forwards run when their result is first used


# --------------------------------------------------------------------------------------------------------
# parameter holders (same construction order and initialisation as the reference, so a given
# torch.manual_seed yields the same initial weights)
# --------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} only holds parameters in the B200 build; call Generator/Critic.forward")


class EqualizedLinear(nn.Linear):
    """N(0,1) weight, zero bias, runtime scale sqrt(2/fan_in) applied inside the kernels (gan.py:7-17)."""

    def __init__(self, in_features, out_features):
        super().__init__(in_features, out_features)
        with torch.no_grad():
            self.weight.normal_()
            self.bias.zero_()
        self.equalized_coefficient = sqrt(2 / in_features)

    forward = _Holder.forward


class EqualizedConv2d(nn.Conv2d):
    """N(0,1) weight, zero bias, runtime scale sqrt(2/(Cin*k*k)) applied at weight-pack time (gan.py:20-38)."""

    def __init__(self, in_chan, out_chan, kernel_size, padding=0):
        super().__init__(in_chan, out_chan, kernel_size=kernel_size, padding=padding)
        with torch.no_grad():
            self.weight.normal_()
            self.bias.zero_()
        self.equalized_coefficient = sqrt(2 / (in_chan * kernel_size * kernel_size))

    forward = _Holder.forward


class InjectSecondaryNoise(_Holder):
    def __init__(self, channels):
        super().__init__()
        self.weights = nn.Parameter(torch.zeros((1, channels, 1, 1)))      # gan.py:44


class AdaINBlock(_Holder):
    def __init__(self, in_channel, style_dim=512):
        super().__init__()
        self.style = EqualizedLinear(style_dim, in_channel * 2)            # gan.py:60
        with torch.no_grad():                                              # gan.py:62-63
            self.style.bias[:in_channel] = 1
            self.style.bias[in_channel:] = 0


class StyleConvBlock(_Holder):
    def __init__(self, in_chan, out_chan, is_initial=False):
        super().__init__()
        self.is_initial = is_initial
        if is_initial:
            self.conv = nn.Parameter(torch.randn(1, in_chan, 4, 4))        # gan.py:81
        else:
            self.conv = EqualizedConv2d(in_chan, out_chan, kernel_size=3, padding=1)
        self.inject_noise = InjectSecondaryNoise(out_chan)
        self.adain = AdaINBlock(out_chan)


class StyleGanBlock(_Holder):
    def __init__(self, in_chan, out_chan, is_initial=False, does_upsample=True):
        super().__init__()
        if is_initial and does_upsample:
            raise ValueError("You cannot use the Starting Constant and Upsample.")   # gan.py:105-106
        self.is_initial = is_initial
        self.does_upsample = does_upsample
        self.conv_1 = StyleConvBlock(in_chan, out_chan, is_initial=is_initial)
        self.conv_2 = StyleConvBlock(out_chan, out_chan)


class _Slot(_Holder):
    """Parameter-free placeholder keeping nn.Sequential indices equal to the reference's
    (LeakyReLU / AvgPool2d / Flatten positions, gan.py:145,238-262,353-355)."""


class MappingLayers(_Holder):
    def __init__(self, channels=512):
        super().__init__()
        self.layers = nn.Sequential(*[nn.Sequential(EqualizedLinear(channels, channels), _Slot()) for _ in range(8)])


class MiniBatchStdDev(_Holder):
    def __init__(self, group_size=4):
        super().__init__()
        self.group_size = group_size           # mutated like the reference when B % group_size != 0 (gan.py:277-278)


class CriticBlock(_Holder):
    def __init__(self, in_chan, out_chan, is_final_layer=False):
        super().__init__()
        self.is_final_layer = is_final_layer
        if is_final_layer:
            self.conv_1 = nn.Sequential(MiniBatchStdDev(), EqualizedConv2d(in_chan + 1, out_chan, 3, padding=1), _Slot())
            self.conv_2 = nn.Sequential(EqualizedConv2d(out_chan, out_chan, 4), _Slot(), _Slot(),
                                        EqualizedLinear(out_chan, out_chan), _Slot(), EqualizedLinear(out_chan, 1))
        else:
            self.conv_1 = nn.Sequential(EqualizedConv2d(in_chan, out_chan, 3, padding=1), _Slot())
            self.conv_2 = nn.Sequential(EqualizedConv2d(out_chan, out_chan, 3, padding=1), _Slot(), _Slot())


# --------------------------------------------------------------------------------------------------------
# autograd glue: one node per network
# --------------------------------------------------------------------------------------------------------
def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the B200 build has no CPU path")


_REPLICA_MSG = (
    "this {} is an nn.DataParallel replica: the B200 build runs one process per GPU (torchrun + dist.py, see "
    "INTEGRATION.md); single-process multi-device nn.DataParallel (train.py:71,79 with several visible GPUs) is not "
    "supported — launch with CUDA_VISIBLE_DEVICES set to one device per process, where the unchanged wrapper takes "
    "torch's single-device path")


def _reject_replica(module):
    # torch.nn.parallel.replicate() marks its shallow copies with _is_replica; their parameters are plain tensors, so
    # the parameter walk below would fail with an opaque AttributeError long before get_r1_loss could explain
    if getattr(module, "_is_replica", False):
        raise RuntimeError(_REPLICA_MSG.format(type(module).__name__))


def _install_pack_invalidation(module):
    """load_state_dict copies into the parameters in place (version counters move, so the packed bf16 copies refresh
    by themselves), but a checkpoint load is also the natural point to drop every cached pack and transpose."""
    module.register_load_state_dict_post_hook(lambda mod, incompatible: mod._packs.invalidate())


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gen, steps, alpha, crossover, n_noise, z, z2, *rest):
        noise, params = rest[:n_noise], rest[n_noise:]
        track = any(ctx.needs_input_grad)   # all False under torch.no_grad(): sampling keeps no tape
        img, tape = engine.generator_forward(gen, gen._packs, z, noise, steps, alpha, z2=z2, crossover=crossover,
                                             keep_tape=track)
        ctx.gen, ctx.tape, ctx.params, ctx.n_noise = gen, tape, params, n_noise
        ctx.set_materialize_grads(False)
        return img

    @staticmethod
    def backward(ctx, g_img):
        head = (None,) * 5
        n_noise, params = ctx.n_noise, ctx.params
        if g_img is None:
            return head + (None, None) + (None,) * (n_noise + len(params))
        need = {id(p): ctx.needs_input_grad[7 + n_noise + i] for i, p in enumerate(params)}
        hook = getattr(ctx.gen, "_grad_ready_hook", None)
        if hook is None:
            grads, dz, dz2 = engine.generator_backward(ctx.gen, ctx.gen._packs, ctx.tape, g_img, need,
                                                       need_z=ctx.needs_input_grad[5], need_z2=ctx.needs_input_grad[6])
            return head + (dz, dz2) + (None,) * n_noise + tuple(grads.get(id(p)) for p in params)
        # overlapped data-parallel path (dist.GradSync.ready as the hook): every parameter gradient goes into .grad the
        # moment its kernels are queued and is handed to the hook, so its all-reduce runs under the rest of the backward;
        # autograd gets None for the parameters (nothing left to accumulate)
        by_id = {id(p): p for p in params}

        def emit(key, g):
            p = by_id[key]
            g = g.reshape(p.shape)
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
            hook(p)

        _, dz, dz2 = engine.generator_backward(ctx.gen, ctx.gen._packs, ctx.tape, g_img, need,
                                               need_z=ctx.needs_input_grad[5], need_z2=ctx.needs_input_grad[6], emit=emit)
        return head + (dz, dz2) + (None,) * (n_noise + len(params))


class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, critic, steps, alpha, images, *params):
        pred, tape = engine.critic_forward(critic, critic._packs, images, steps, alpha)
        ctx.critic, ctx.tape, ctx.params = critic, tape, params
        critic._last_tape = tape
        ctx.set_materialize_grads(False)
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        params = ctx.params
        if g_pred is None:
            return (None,) * (4 + len(params))
        need = {id(p): ctx.needs_input_grad[4 + i] for i, p in enumerate(params)}
        grads, g_img = engine.critic_backward(ctx.critic, ctx.critic._packs, ctx.tape, g_pred, need,
                                              need_img=ctx.needs_input_grad[3])
        return (None, None, None, g_img) + tuple(grads.get(id(p)) for p in params)


class _LogisticLossFn(torch.autograd.Function):
    """mean softplus(sign * pred) with its gradient from the same kernel (gan.py:228)."""

    @staticmethod
    def forward(ctx, pred, sign):
        p = pred.detach().float().contiguous()
        n = p.numel()
        loss = torch.empty(1, device=p.device)
        seed = torch.empty_like(p)
        engine.call("bg_logistic_loss", p, n, float(sign), loss, seed, 1.0)
        ctx.save_for_backward(seed)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (seed,) = ctx.saved_tensors
        return seed * g, None


# --------------------------------------------------------------------------------------------------------
# public modules
# --------------------------------------------------------------------------------------------------------
class Generator(nn.Module):
    def __init__(self):
        super().__init__()
        self.to_w_noise = nn.Sequential(MappingLayers())
        self.gen_blocks = nn.ModuleList(
            [StyleGanBlock(512, 512, is_initial=True, does_upsample=False)]
            + [StyleGanBlock(ci, co) for ci, co in engine.GEN_CHANNELS[1:]])
        self.to_rgbs = nn.ModuleList([EqualizedConv2d(co, 3, kernel_size=1) for _, co in engine.GEN_CHANNELS])
        self._packs = PackCache()
        _install_pack_invalidation(self)

    def invalidate_packs(self):
        """Drop the cached bf16 weight packs.  They refresh automatically when a parameter's version counter or storage
        changes (optimizer steps, load_state_dict, copy_/mul_ on the parameter); writes through `.data`
        (p.data.mul_(...), EMA / weight-clipping code) do NOT bump the counter and must be followed by this call."""
        self._packs.invalidate()

    def forward(self, z_noise, noise=None, steps=1, alpha=None, z2=None, crossover=None):
        _reject_replica(self)
        _require_cuda(z_noise, "z_noise")
        steps = int(steps)
        if steps > len(self.gen_blocks):
            return None                                   # the reference falls off its loop (gan.py:201-222)
        batch = len(z_noise)
        if noise is None:                                 # gan.py:189-197: torch global RNG, on z's device
            noise = [torch.randn(batch, 1, 4 * 2 ** i, 4 * 2 ** i, device=z_noise.device) for i in range(steps)]
        noise = tuple(noise[:steps])
        fade = alpha is not None and steps > 1
        params = engine.generator_params(self, steps, fade)
        if LAZY_NO_GRAD_FORWARD and not torch.is_grad_enabled():
            # the noise above is already drawn (same RNG consumption as the reference); the kernels run on first use
            z_now, z2_now = z_noise.detach().clone(), None if z2 is None else z2.detach().clone()

            def thunk():
                with torch.no_grad():
                    return _GeneratorFn.apply(self, steps, alpha, crossover, len(noise), z_now, z2_now, *noise, *params)

            r = 4 * 2 ** (steps - 1)
            return lazy.LazyImages(thunk, (batch, 3, r, r), z_noise.device)
        return _GeneratorFn.apply(self, steps, alpha, crossover, len(noise), z_noise, z2, *noise, *params)

    def get_wgan_loss(self, crit_fake_pred):
        loss = -crit_fake_pred.mean()                     # gan.py:224-225
        return lazy.defer_item(loss) if DEFER_LOSS_ITEMS else loss

    def get_r1_loss(self, crit_fake_pred):
        loss = _LogisticLossFn.apply(crit_fake_pred, -1.0)   # softplus(-pred).mean(), gan.py:227-228
        return lazy.defer_item(loss) if DEFER_LOSS_ITEMS else loss


class Critic(nn.Module):
    def __init__(self):
        super().__init__()
        self.from_rgbs = nn.ModuleList([self.gen_from_rgbs(ci) for ci, _ in engine.CRITIC_CHANNELS])
        self.conv_blocks = nn.ModuleList(
            [CriticBlock(ci, co) for ci, co in engine.CRITIC_CHANNELS[:7]] + [CriticBlock(512, 512, is_final_layer=True)])
        self._packs = PackCache()
        self._last_tape = None
        _install_pack_invalidation(self)

    invalidate_packs = Generator.invalidate_packs

    def gen_from_rgbs(self, out_chan, image_chan=3):
        return nn.Sequential(EqualizedConv2d(image_chan, out_chan, kernel_size=1), _Slot())

    def forward(self, images, steps=1, alpha=None):
        _reject_replica(self)
        _require_cuda(images, "images")
        steps = int(steps)
        fade = alpha is not None and steps > 1
        params = engine.critic_params(self, steps, fade)
        pred = _CriticFn.apply(self, steps, alpha, images, *params)
        pred._bg_tape = self._last_tape                   # lets get_r1_loss schedule the double-backward by hand
        self._last_tape = None
        return pred

    def _emit_into_grad(self):
        hook = getattr(self, "_grad_ready_hook", None)

        def emit(p, g):
            # called while the backward is still running, as each parameter's total gradient becomes final
            g = g.reshape(p.shape)
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
            if hook is not None:
                hook(p)                                   # e.g. dist.GradSync.ready: bucketed all-reduce, overlapped

        return emit

    @staticmethod
    def _tapes(*preds):
        tapes = [getattr(p, "_bg_tape", None) for p in preds]
        if any(t is None for t in tapes):
            raise RuntimeError(
                "the critic losses need the predictions returned by this Critic's forward (they carry the saved "
                "activations).  Under multi-device nn.DataParallel the gather drops them: launch one process "
                "per GPU instead (see INTEGRATION.md).")
        return tapes

    def get_wgan_loss(self, crit_fake_pred, crit_real_pred, real_im, steps, alpha, c_lambda=1, fake_im=None,
                      epsilon=None):
        """WGAN-GP critic loss as gan.py:357-391 intends it; runs its own backward like the reference (gan.py:389).

        The reference's body cannot execute: it reads `self.device` (gan.py:368; nn.Module has none) and an undefined
        `fake_im` (gan.py:372; train.py:178-185 does not pass one).  The signature train.py uses is kept; the fake images
        are the ones `crit_fake_pred` was computed from (saved with its activations), or `fake_im=` when given;
        `epsilon=` (B,1,1,1) overrides the torch.rand draw of gan.py:367-369 (parity tests)."""
        tape_f, tape_r = self._tapes(crit_fake_pred, crit_real_pred)
        _require_cuda(real_im, "real_im")
        fake = tape_f["img"] if fake_im is None else fake_im.detach().float()
        if epsilon is None:
            epsilon = torch.rand(real_im.shape[0], 1, 1, 1, device=real_im.device)
        mixed = torch.lerp(fake, real_im.detach().float(), epsilon.to(real_im.device))      # gan.py:372
        _, tape_m = engine.critic_forward(self, self._packs, mixed, int(steps), alpha)      # gan.py:373
        loss, _, g_m = engine.critic_wgan_gp_step(self, self._packs, tape_f, crit_fake_pred, tape_r, crit_real_pred,
                                                  tape_m, c_lambda, emit=self._emit_into_grad())
        self.last_mixed_image_grad = g_m                  # d sum(D(mixed)) / d mixed, what autograd.grad returned
        return lazy.defer_item(loss) if DEFER_LOSS_ITEMS else loss

    def get_r1_loss(self, crit_fake_pred, crit_real_pred, real_im, fake_im, steps, alpha, c_lambda=1):
        """gan.py:393-412.  Runs the backward itself (like the reference's r1_loss.backward()) and accumulates
        into .grad of every critic parameter that requires grad; returns the loss value."""
        tape_f, tape_r = self._tapes(crit_fake_pred, crit_real_pred)
        loss, _, g_x = engine.critic_r1_step(self, self._packs, tape_f, crit_fake_pred, tape_r, crit_real_pred,
                                             c_lambda, emit=self._emit_into_grad(),
                                             first_order=not getattr(self, "_r1_penalty_only", False))
        self.last_real_image_grad = g_x                   # d sum(D(real)) / d real, what autograd.grad returned
        return lazy.defer_item(loss) if DEFER_LOSS_ITEMS else loss
