"""Host-side schedule of the sm_100a kernels for the StyleGAN hot path.

This module owns WHAT runs in which order for
  * Generator.forward / its backward            (reference gan.py:183-222 + autograd),
  * Critic.forward / its first-order backward    (reference gan.py:331-349 + autograd),
  * the R1 critic loss with its double-backward  (reference gan.py:393-412),
and nothing else: all arithmetic on feature maps happens inside libbg_b200.so (include/bg_b200.h).  torch is
used for device memory (the caching allocator), the current stream, and a handful of reshapes/concats on
(B, C)-sized tensors.  There is no CPU path and no eager fallback: every op goes through bg_native.call(),
which raises if the library is missing or a kernel reports an error.

Layout: feature maps NHWC bf16; images at the module boundary NCHW fp32; the mapping network, style
vectors, instance-norm statistics, critic head and every parameter gradient are fp32.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

import bg_native as bgn

SLOPE = 0.2          # nn.LeakyReLU(0.2) everywhere (gan.py:86,145,241,...)
IN_EPS = 1e-8        # nn.InstanceNorm2d(eps=1e-8), gan.py:59
MBSTD_EPS = 1e-8     # gan.py:287
GEN_CHANNELS = [(512, 512), (512, 512), (512, 512), (512, 256), (256, 128), (128, 64), (64, 32), (32, 16)]
CRITIC_CHANNELS = [(16, 32), (32, 64), (64, 128), (128, 256), (256, 512), (512, 512), (512, 512), (512, 512)]
MBSTD_CPAD = 576     # 512 features + 1 stddev plane, padded to a multiple of 64 for the tensor-core conv

call = bgn.call


def _bf16(*shape, device):
    return torch.empty(shape, dtype=torch.bfloat16, device=device)


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


class ZeroSlab:
    """One zero-filled fp32 slab per pass that every accumulation target of the pass (fused statistics, split-K weight-
    gradient scratch, channel sums) is carved from: ONE memset instead of one zero-fill launch per accumulator (~120 per
    iteration).  While active the slab is registered with the library (bg_set_prezeroed_range), which then skips its
    own zero fill for targets inside it.  Buffers that do not fit fall back to ordinary allocations (zeroed by the call)."""

    current: "Optional[ZeroSlab]" = None

    def __init__(self, device, nfloats: int):
        self.device = device
        self.buf = torch.zeros(max(int(nfloats), 4), dtype=torch.float32, device=device)
        self.off = 0
        self.prev = None

    def __enter__(self):
        self.prev = ZeroSlab.current
        ZeroSlab.current = self
        bgn.lib().bg_set_prezeroed_range(self.buf.data_ptr(), self.buf.numel() * 4)
        return self

    def __exit__(self, *exc):
        ZeroSlab.current = self.prev
        if self.prev is not None:
            bgn.lib().bg_set_prezeroed_range(self.prev.buf.data_ptr(), self.prev.buf.numel() * 4)
        else:
            bgn.lib().bg_set_prezeroed_range(None, 0)

    def take(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        n_pad = (n + 3) & ~3                       # keep every buffer 16-byte aligned (vector reductions)
        if self.off + n_pad > self.buf.numel():
            return None
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += n_pad
        return v


def _acc(*shape, device):
    """fp32 accumulation target: a carved piece of the active ZeroSlab (already zero), else a plain allocation that the
    library zeroes itself."""
    slab = ZeroSlab.current
    if slab is not None and slab.device == device:
        v = slab.take(shape)
        if v is not None:
            return v
    return torch.empty(shape, dtype=torch.float32, device=device)


def coef_of(weight: torch.Tensor) -> float:
    """Equalized-lr runtime scale sqrt(2 / fan_in) (gan.py:13-14, 26-27)."""
    fan_in = weight.shape[1] * (weight[0][0].numel() if weight.dim() > 2 else 1)
    return math.sqrt(2.0 / fan_in)


# ------------------------------------------------------------------------------------------------------
# packed-weight cache: bf16 [tap][Cout][Cin] (fprop) and [tap'][Cin][Cout] (dgrad) copies of the fp32
# master weights with the equalized coefficient folded in; fp32 transposes of the linear weights.
# A pack is stale when the parameter's version counter or storage changed (load_state_dict, copy_/mul_ on the
# parameter, for-each optimizers) OR an optimizer has stepped it since: the multi-tensor `fused=True` optimizers
# update the masters WITHOUT moving the version counter (torch 2.11: Adam(fused=True).step() leaves p._version
# unchanged), so every torch optimizer step is observed through a global step post-hook that stamps the parameters of
# that optimizer.  Writes through `.data` are invisible to both and need PackCache.invalidate().
# ------------------------------------------------------------------------------------------------------
_OPT_EPOCH: Dict[int, int] = {}


def _note_optimizer_step(optimizer, args, kwargs):
    for group in optimizer.param_groups:
        for p in group["params"]:
            _OPT_EPOCH[id(p)] = _OPT_EPOCH.get(id(p), 0) + 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook  # noqa: E402

_register_step_hook(_note_optimizer_step)


def _tag(w: torch.Tensor, *extra):
    return (w._version, w.data_ptr(), _OPT_EPOCH.get(id(w), 0)) + extra


class PackCache:
    def __init__(self):
        self._conv: Dict[int, tuple] = {}
        self._lin: Dict[int, tuple] = {}

    def conv(self, w: torch.Tensor, cin_pad: Optional[int] = None):
        key = id(w)
        tag = _tag(w, cin_pad)
        hit = self._conv.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1], hit[2]
        cout, cin, ks, _ = w.shape
        cp = cin_pad or cin
        wf = _bf16(ks * ks, cout, cp, device=w.device)
        wd = _bf16(ks * ks, cp, cout, device=w.device)
        call("bg_pack_weight", w.detach(), wf, wd, cout, cin, cp, ks, coef_of(w))
        self._conv[key] = (tag, wf, wd)
        return wf, wd

    def conv_pool4(self, w: torch.Tensor):
        """16-tap pack of conv3x3 followed by AvgPool2d(2) as one 4x4 stride-2 conv (bg_pack_weight_pool4)."""
        key = ("p4", id(w))
        tag = _tag(w)
        hit = self._conv.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        cout, cin = w.shape[0], w.shape[1]
        w16 = _bf16(16, cout, cin, device=w.device)
        call("bg_pack_weight_pool4", w.detach(), w16, cout, cin, coef_of(w))
        self._conv[key] = (tag, w16, None)
        return w16

    def conv_tconv4(self, w: torch.Tensor):
        """16-tile pack of the input gradient of conv3x3 -> AvgPool2d(2) (transposed 4x4 stride-2 conv)."""
        key = ("t4", id(w))
        tag = _tag(w)
        hit = self._conv.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        cout, cin = w.shape[0], w.shape[1]
        wt = _bf16(16, cin, cout, device=w.device)
        call("bg_pack_weight_tconv4", w.detach(), wt, cout, cin, coef_of(w))
        self._conv[key] = (tag, wt, None)
        return wt

    def invalidate(self):
        """Forget every pack: the next forward re-packs from the fp32 masters.  Needed after writes that bypass the
        version counter (p.data.mul_(...), p.data.copy_(...)); optimizer steps and load_state_dict do not need it."""
        self._conv.clear()
        self._lin.clear()

    def refresh_convs(self, items):
        """Re-pack every stale conv weight of `items` = [(weight, cin_pad or None)] in ONE grouped launch (after an
        optimizer step all of a network's packs are stale); the per-layer conv() lookups that follow then hit."""
        stale = []
        for w, cin_pad in items:
            hit = self._conv.get(id(w))
            if hit is None or hit[0] != _tag(w, cin_pad):
                stale.append((w, cin_pad))
        for i in range(0, len(stale), 32):
            grp = stale[i:i + 32]
            wfs, wds, meta = [], [], []
            for w, cin_pad in grp:
                cout, cin, ks, _ = w.shape
                cp = cin_pad or cin
                wfs.append(_bf16(ks * ks, cout, cp, device=w.device))
                wds.append(_bf16(ks * ks, cp, cout, device=w.device))
                meta.append((cout, cin, cp, ks, coef_of(w)))
            call("bg_pack_weight_grouped", [w.detach() for w, _ in grp], wfs, wds, [m[0] for m in meta],
                 [m[1] for m in meta], [m[2] for m in meta], [m[3] for m in meta], [m[4] for m in meta], len(grp))
            for (w, cin_pad), wf, wd in zip(grp, wfs, wds):
                self._conv[id(w)] = (_tag(w, cin_pad), wf, wd)


# ------------------------------------------------------------------------------------------------------
# thin op helpers
# ------------------------------------------------------------------------------------------------------
def linear_fwd(x, w, bias, act, coef=None):
    w2 = w.detach().reshape(w.shape[0], -1)
    m, k = x.shape
    n = w2.shape[0]
    y = _f32(m, n, device=x.device)
    call("bg_linear_fwd", x, w2, None if bias is None else bias.detach(), y, m, n, k,
         coef_of(w) if coef is None else coef, 1 if act else 0, SLOPE)
    return y


def linear_bwd_input(gy, w, packs: PackCache = None):
    """gx = coef * gy @ W, read straight from the (N, K) weight (no transposed copy)."""
    w2 = w.detach().reshape(w.shape[0], -1)
    n, k = w2.shape
    m = gy.shape[0]
    gx = _f32(m, k, device=gy.device)
    call("bg_linear_bwd_input", gy, w2, gx, m, n, k, coef_of(w))
    return gx


def linear_bwd_weight(gy, x, w, want_bias=True, into=None, into_b=None):
    """Returns (dW shaped like w, db) ; with `into` accumulates in place instead."""
    m, n = gy.shape
    k = x.shape[1]
    acc = into is not None
    dw = into if acc else _f32(*w.shape, device=gy.device)
    db = into_b if acc else (_f32(n, device=gy.device) if want_bias else None)
    call("bg_linear_bwd_weight", gy, x, dw, db, m, n, k, coef_of(w), 1 if acc else 0)
    return dw, db


def gate_f32(g, y):
    out = torch.empty_like(g)
    call("bg_act_gate_f32", g, y, out, g.numel(), SLOPE)
    return out


def conv3x3(x, wpack, cin, cout, bias=None, noise=None, noise_w=None, gate_src=None, act=False, stats=0):
    """3x3 conv with the fused epilogue.  stats=1 also returns the instance-norm sums (N,Cout,2) of the output,
    stats=2 its per-channel total (Cout) — both reduced inside the conv epilogue."""
    n, h, w_, _ = x.shape
    out = _bf16(n, h, w_, cout, device=x.device)
    if stats:
        st = _acc(n, cout, 2, device=x.device) if stats == 1 else _acc(cout, device=x.device)
        call("bg_conv_fprop_stats", x, wpack, out, n, h, w_, cin, cout, 3, bias, noise, noise_w, gate_src,
             1 if act else 0, SLOPE, st, stats)
        return out, st
    call("bg_conv_fprop", x, wpack, out, n, h, w_, cin, cout, 3, bias, noise, noise_w, gate_src, 1 if act else 0, SLOPE)
    return out


def conv3x3_pool(x, wpack, cin, cout, bias=None, gate_src=None, act=True, w16=None):
    """conv3x3 -> AvgPool2d(2) -> LeakyReLU (gan.py:258-262), or with act=False and gate_src the R1 tangent of it.
    H,W >= 32 (and a 16-tap pack w16): ONE 4x4 stride-2 convolution, 2.25x fewer MACs; H,W == 16: 3x3 conv with the
    pool in its epilogue; below that conv + pool kernels."""
    n, h, w_, _ = x.shape
    out = _bf16(n, h // 2, w_ // 2, cout, device=x.device)
    if w16 is not None and h >= 32 and w_ >= 32 and cin % 32 == 0:
        call("bg_conv_pool4_fprop", x, w16, out, n, h, w_, cin, cout, bias, gate_src, 1 if act else 0, SLOPE)
    elif h >= 16 and w_ >= 16:
        call("bg_conv_pool_fprop", x, wpack, out, n, h, w_, cin, cout, bias, gate_src, 1 if act else 0, SLOPE)
    else:
        u = conv3x3(x, wpack, cin, cout, bias=bias, act=False)
        call("bg_pool_act_fwd", u, gate_src, out, n, h // 2, w_ // 2, cout, SLOPE, 0 if act else 1)
    return out


def conv_wgrad(x, g, w_param, cin_pad=None, extra=None, into=None):
    """dW for a 3x3 conv parameter `w_param` from input x and output-gradient g (both NHWC bf16).
    `extra=(v, ghat)` adds the R1 second-order pair into the same accumulator (doubled-K contraction).
    `into`: an existing gradient of w_param's shape that the result is ADDED to by the unpack pass (returned)."""
    n, h, w_, cin_eff = x.shape
    cout = g.shape[3]
    dwp = _acc(9, cout, cin_eff, device=x.device)
    call("bg_conv_wgrad", x, g, dwp, n, h, w_, cin_eff, cout, 0)
    if extra is not None:
        call("bg_conv_wgrad", extra[0], extra[1], dwp, n, h, w_, cin_eff, cout, 1)
    dw = into if into is not None else _f32(*w_param.shape, device=x.device)
    call("bg_unpack_wgrad", dwp, dw, cout, w_param.shape[1], cin_eff, 3, coef_of(w_param), 0 if into is None else 1)
    return dw


def conv_pool_wgrad(x, gpool, w_param, extra=None, into=None):
    """dW of a 3x3 conv that is followed by AvgPool2d(2), from the gradient gpool at the POOLED map (NHWC bf16) and the
    conv input x: taken on the equivalent 4x4 stride-2 kernel (16 taps, a quarter of the pixels), folded back to 3x3.
    `extra=(v, ghat_pooled)` adds the R1 second-order pair into the same accumulator."""
    n, hp, wp, cout = gpool.shape
    cin = x.shape[3]
    dw16 = _acc(16, cout, cin, device=x.device)
    call("bg_conv_pool4_wgrad", x, gpool, dw16, n, hp, wp, cin, cout, 0)
    if extra is not None:
        call("bg_conv_pool4_wgrad", extra[0], extra[1], dw16, n, hp, wp, cin, cout, 1)
    dw = into if into is not None else _f32(*w_param.shape, device=x.device)
    call("bg_unpack_wgrad_pool4", dw16, dw, cout, cin, coef_of(w_param), 0 if into is None else 1)
    return dw


def channel_wsum(g, planes, nplanes, hw, img_stride, plane_stride):
    c = g.shape[-1]
    p = g.numel() // c
    out = _acc(1 + nplanes, c, device=g.device)
    call("bg_channel_wsum", g, planes, out, p, c, hw, img_stride, plane_stride, nplanes)
    return out


def axpby(a, b, ca, cb):
    out = torch.empty_like(a)
    call("bg_axpby", a, b, out, a.numel(), float(ca), float(cb))
    return out


def clamp_alpha(alpha):
    return min(1.0, max(0.0, float(alpha)))      # gan.py:211,344


# ======================================================================================================
# Generator
# ======================================================================================================
def generator_params(gen, steps: int, fade: bool) -> List[torch.nn.Parameter]:
    """Parameters Generator.forward touches at (steps, fade), in a fixed order.  Everything else keeps
    .grad = None exactly as in the reference (its autograd never reaches unused blocks)."""
    ps = []
    for i in range(8):
        lin = gen.to_w_noise[0].layers[i][0]
        ps += [lin.weight, lin.bias]
    for k in range(steps):
        blk = gen.gen_blocks[k]
        for sc in (blk.conv_1, blk.conv_2):
            if isinstance(sc.conv, torch.nn.Parameter):
                ps.append(sc.conv)
            else:
                ps += [sc.conv.weight, sc.conv.bias]
            ps += [sc.inject_noise.weights, sc.adain.style.weight, sc.adain.style.bias]
    if fade:
        ps += [gen.to_rgbs[steps - 2].weight, gen.to_rgbs[steps - 2].bias]
    ps += [gen.to_rgbs[steps - 1].weight, gen.to_rgbs[steps - 1].bias]
    return ps


def style_conv_fusable(R: int, cout: int) -> bool:
    """Fold the previous AdaIN into per-sample weights (bg_style_modulate + bg_conv_style_fprop) when the halo kernel
    runs the layer (R >= 16) and the per-sample packs (9*Cin*Cout per sample) are cheaper to write and read than the
    normalised map (R*R*Cin per sample) the unfused path materialises."""
    if os.environ.get("BG_NO_STYLE_FUSION"):         # A/B switch (tools/flaky_probe.py); the default path is fused
        return False
    return R >= 16 and 18 * cout <= R * R


def layer_output(layers, idx):
    """AdaIN output gamma * IN(a) + beta (gan.py:69) of layer idx as a materialised NHWC bf16 map (cached)."""
    L = layers[idx]
    if L["xo"] is None:
        a = L["a"]
        xo = torch.empty_like(a)
        call("bg_adain_apply", a, L["stats"], L["style"], xo, a.shape[0], L["R"] * L["R"], L["C"], IN_EPS)
        L["xo"] = xo
    return L["xo"]


def layer_input(layers, idx, L):
    """Materialised conv input of layer idx: AdaIN output of the layer before it, bilinearly upsampled for conv_1
    (gan.py:122-123).  The fused forward never builds it; the unfused layers and the weight-gradient pass do."""
    xo = layer_output(layers, idx - 1)
    if not L["up"]:
        return xo
    B, r, _, c = xo.shape
    xin = _bf16(B, 2 * r, 2 * r, c, device=xo.device)
    call("bg_upsample2x_fwd", xo, xin, B, r, r, c)
    return xin


def _wgrad_scratch_floats(params, need) -> int:
    """Upper bound of the split-K weight-gradient scratch of a backward pass: 16/9 of every wanted 3x3 conv weight (the
    pooled 4x4-stride-2 form has 16 taps; the padded 513 -> 576 layer stays below that factor)."""
    n = 0
    for p in params:
        if p.dim() == 4 and p.shape[-1] == 3 and need.get(id(p), False):
            n += p.numel() * 16 // 9 + 64
    return n


def generator_forward(gen, packs: PackCache, z, noise, steps, alpha, z2=None, crossover=None, keep_tape=True):
    """Generator.forward (gan.py:183-222).  Returns (image (B,3,R,R) fp32, tape)."""
    B = z.shape[0]
    with ZeroSlab(z.device, sum(2 * (2 * B * GEN_CHANNELS[k][1] + 16) for k in range(steps))):
        return _generator_forward(gen, packs, z, noise, steps, alpha, z2, crossover, keep_tape)


def _generator_forward(gen, packs: PackCache, z, noise, steps, alpha, z2=None, crossover=None, keep_tape=True):
    dev = z.device
    B = z.shape[0]
    fade = alpha is not None and steps > 1
    a_mix = clamp_alpha(alpha) if fade else None

    def mapping(zz):
        hs = [zz.detach().float().contiguous()]
        for i in range(8):                                   # MappingLayers, gan.py:130-148
            lin = gen.to_w_noise[0].layers[i][0]
            hs.append(linear_fwd(hs[-1], lin.weight, lin.bias, act=True))
        return hs

    packs.refresh_convs([(sc.conv.weight, None) for k in range(steps)
                         for sc in (gen.gen_blocks[k].conv_1, gen.gen_blocks[k].conv_2) if not (k == 0 and sc.is_initial)])
    maps_cat = None
    if z2 is not None:
        # style mixing: both latents go through the mapping network as ONE batch of 2B rows (8 launches instead of 16,
        # and one backward pass instead of two); maps[i] are row views of it
        maps_cat = mapping(torch.cat([z.detach().float(), z2.detach().float()]))
        maps = [[h[:B] for h in maps_cat], [h[B:] for h in maps_cat]]
    else:
        maps = [mapping(z)]
    layers = []
    # AdaIN style vectors [gamma | beta] = EqualizedLinear(512 -> 2C)(w) of every layer (gan.py:60,66): all layers fed
    # by the same latent go in ONE grouped launch
    which_of = [1 if (z2 is not None and crossover is not None and k >= crossover) else 0 for k in range(steps)]
    styles = {}
    for wsel in sorted(set(which_of)):
        fcs = [(k, j, sc.adain.style) for k in range(steps) if which_of[k] == wsel
               for j, sc in enumerate((gen.gen_blocks[k].conv_1, gen.gen_blocks[k].conv_2))]
        outs = [_f32(B, fc.weight.shape[0], device=dev) for _, _, fc in fcs]
        call("bg_linear_fwd_grouped", maps[wsel][-1], [fc.weight.detach() for _, _, fc in fcs],
             [fc.bias.detach() for _, _, fc in fcs], outs, [fc.weight.shape[0] for _, _, fc in fcs],
             [coef_of(fc.weight) for _, _, fc in fcs], len(fcs), B, 512, 0, SLOPE)
        for (k, j, _), o in zip(fcs, outs):
            styles[(k, j)] = o
    for k in range(steps):
        blk = gen.gen_blocks[k]
        cin, cout = GEN_CHANNELS[k]
        R = 4 << k
        which = which_of[k]
        nz = noise[k].detach().float().contiguous()
        for j, sc in enumerate((blk.conv_1, blk.conv_2)):
            style = styles[(k, j)]
            nw = sc.inject_noise.weights.detach().reshape(-1)
            L = dict(k=k, j=j, sc=sc, xin=None, R=R, C=cout, which=which, noise=nz, style=style, xo=None,
                     prev=len(layers) - 1, up=(j == 0 and k > 0))
            if k == 0 and j == 0:
                a = _bf16(B, 4, 4, cout, device=dev)                             # gan.py:92,96-97
                call("bg_const_noise_act", sc.conv.detach(), nz, nw, a, B, 16, cout, SLOPE)
                stats = _acc(B, cout, 2, device=dev)
                call("bg_in_stats", a, stats, B, R * R, cout)
                L["const"] = True
            else:
                ci = cin if j == 0 else cout
                P = layers[-1]
                if style_conv_fusable(R, cout):
                    # gan.py:122-123 + 94-98 in ONE kernel: the previous layer's AdaIN is folded into per-sample
                    # weights / bias rows, the upsample happens in the operand feed, IN sums come out of the epilogue
                    wmod = _bf16(B, 9, cout, ci, device=dev)
                    btab = _f32(B, 9, cout, device=dev)
                    call("bg_style_modulate", sc.conv.weight.detach(), sc.conv.bias.detach(), P["stats"], P["style"],
                         wmod, btab, B, ci, cout, P["R"] * P["R"], coef_of(sc.conv.weight), IN_EPS)
                    a = _bf16(B, R, R, cout, device=dev)
                    stats = _acc(B, cout, 2, device=dev)
                    call("bg_conv_style_fprop", P["a"], wmod, btab, a, B, R, R, ci, cout, 1 if L["up"] else 0, nz, nw,
                         SLOPE, stats)
                    del wmod, btab
                else:
                    xin = layer_input(layers, len(layers), L)
                    wf, _ = packs.conv(sc.conv.weight)
                    a, stats = conv3x3(xin, wf, ci, cout, bias=sc.conv.bias.detach(), noise=nz, noise_w=nw, act=True,
                                       stats=1)                                  # IN sums ride in the conv epilogue
                    if keep_tape:
                        L["xin"] = xin
                L["const"] = False
            L["a"], L["stats"] = a, stats
            layers.append(L)
    R = 4 << (steps - 1)
    C = GEN_CHANNELS[steps - 1][1]
    rgb = gen.to_rgbs[steps - 1]
    img = _f32(B, 3, R, R, device=dev)
    last = layers[-1]
    call("bg_to_rgb_adain", last["a"], last["stats"], last["style"], rgb.weight.detach(), rgb.bias.detach(), img, B, R * R,
         C, coef_of(rgb.weight), IN_EPS)                                         # gan.py:218,222 on AdaIN(a)
    if fade:
        Cp = GEN_CHANNELS[steps - 2][1]
        rgb_p = gen.to_rgbs[steps - 2]
        small = _f32(B, 3, R // 2, R // 2, device=dev)
        lp = layers[2 * (steps - 1) - 1]                                         # conv_2 of the previous block
        call("bg_to_rgb_adain", lp["a"], lp["stats"], lp["style"], rgb_p.weight.detach(), rgb_p.bias.detach(), small, B,
             R * R // 4, Cp, coef_of(rgb_p.weight), IN_EPS)
        out = _f32(B, 3, R, R, device=dev)
        call("bg_img_up2_lerp", small, img, out, B * 3, R // 2, R // 2, a_mix)   # gan.py:213-220
        img = out
    if not keep_tape:
        for L in layers:
            L["xo"] = None
    tape = None
    if keep_tape:
        tape = dict(maps=maps, maps_cat=maps_cat, layers=layers, steps=steps, fade=fade, a_mix=a_mix, B=B)
    return img, tape


def generator_backward(gen, packs: PackCache, tape, g_img, need: Dict[int, bool], need_z=True, need_z2=False,
                       emit=None):
    """See _generator_backward; runs it inside one pre-zeroed slab for all of the pass's accumulators."""
    steps, B = tape["steps"], tape["B"]
    small = sum(2 * (2 * B * GEN_CHANNELS[k][1] + 6 * GEN_CHANNELS[k][1] + 64) for k in range(steps)) + 4096
    with ZeroSlab(g_img.device, small + _wgrad_scratch_floats(generator_params(gen, steps, tape["fade"]), need)):
        return _generator_backward(gen, packs, tape, g_img, need, need_z, need_z2, emit)


def _generator_backward(gen, packs: PackCache, tape, g_img, need: Dict[int, bool], need_z=True, need_z2=False,
                        emit=None):
    """Backward of generator_forward.  `need[id(param)]` says which parameter gradients to produce.
    emit(id(param), grad): called the moment a parameter's TOTAL gradient has been queued (synthesis layers one by
    one, the mapping network at the end because both latents of a style-mixing pass accumulate into it).
    Returns (grads {id(param): tensor}, dz, dz2)."""
    dev = g_img.device
    steps, fade, B = tape["steps"], tape["fade"], tape["B"]
    layers, maps = tape["layers"], tape["maps"]
    grads: Dict[int, torch.Tensor] = {} if emit is None else _EmitDict(emit)
    g_img = g_img.detach().float().contiguous()
    R = 4 << (steps - 1)
    C = GEN_CHANNELS[steps - 1][1]
    g_w = [None] * len(maps)
    dstyles = []

    def want(p):
        return need.get(id(p), False)

    def rgb_backward(rgb, x_feat, g_planes, r, c):
        """toRGB 1x1 conv (gan.py:172-179): returns grad wrt the NHWC feature map; fills weight/bias grads."""
        if want(rgb.weight):
            ws = channel_wsum(x_feat, g_planes, 3, r * r, 3 * r * r, r * r)
            grads[id(rgb.weight)] = (ws[1:4] * coef_of(rgb.weight)).reshape(3, c, 1, 1)
        if want(rgb.bias):
            sums = _f32(3, device=dev)
            call("bg_plane_sums", g_planes, sums, B, r * r)
            grads[id(rgb.bias)] = sums
        gx = _bf16(B, r, r, c, device=dev)
        call("bg_planes3_to_nhwc", g_planes, rgb.weight.detach(), None, None, gx, B * r * r, r * r, c, 1, c,
             coef_of(rgb.weight), 0, SLOPE)
        return gx

    g_prev_extra = None
    if fade:
        a_mix = tape["a_mix"]
        g_large = _f32(B, 3, R, R, device=dev)
        call("bg_axpby_f32", g_img, None, g_large, g_img.numel(), a_mix, 0.0)
        g_small = _f32(B, 3, R // 2, R // 2, device=dev)
        call("bg_img_up2_bwd", g_img, g_small, B * 3, R // 2, R // 2, 1.0 - a_mix)
        ip = 2 * (steps - 1) - 1                                   # conv_2 of the previous block feeds to_rgbs[steps-2]
        g_prev_extra = rgb_backward(gen.to_rgbs[steps - 2], layer_output(layers, ip), g_small, R // 2,
                                    GEN_CHANNELS[steps - 2][1])
    else:
        g_large = g_img
    gx = rgb_backward(gen.to_rgbs[steps - 1], layer_output(layers, len(layers) - 1), g_large, R, C)
    layers[-1]["xo"] = None

    for idx in reversed(range(len(layers))):
        L = layers[idx]
        sc, k, j, r, c = L["sc"], L["k"], L["j"], L["R"], L["C"]
        a, stats, style = L["a"], L["stats"], L["style"]
        bs = _acc(B, c, 2, device=dev)
        gpre = torch.empty_like(a)
        is_const = L["const"]
        need_b = (not is_const) and want(sc.conv.bias)
        need_nw = want(sc.inject_noise.weights)
        # bias / noise-weight gradients (gan.py:30,52) are reduced while gpre is written
        ws = _acc(2, c, device=dev) if (need_b or need_nw) else None
        call("bg_adain_bwd_reduce", gx, a, stats, bs, B, r * r, c, IN_EPS)
        call("bg_adain_bwd_apply", gx, a, stats, style, bs, gpre, B, r * r, c, IN_EPS, SLOPE, 1,
             L["noise"] if ws is not None else None, ws)
        # style FC (gan.py:60,66): dL/dgamma = sum g*ahat, dL/dbeta = sum g; the FC gradients of all layers are taken in
        # grouped launches after the loop
        dstyles.append((L["which"], sc.adain.style, torch.cat([bs[..., 1], bs[..., 0]], dim=1).contiguous()))
        if need_b:
            grads[id(sc.conv.bias)] = ws[0]
        if need_nw:
            grads[id(sc.inject_noise.weights)] = ws[1].reshape(1, c, 1, 1)
        if is_const:
            if want(sc.conv):
                dc = _f32(1, c, 4, 4, device=dev)
                call("bg_const_bwd", gpre, dc, B, 16, c)
                grads[id(sc.conv)] = dc
            gx = None
            continue
        ci = sc.conv.weight.shape[1]
        if want(sc.conv.weight):
            # the fused forward never built this layer's input: the weight gradient is its only consumer, so it is
            # rebuilt here (AdaIN apply [+ upsample] of the previous layer) and dropped right after
            xin = L["xin"] if L["xin"] is not None else layer_input(layers, idx, L)
            grads[id(sc.conv.weight)] = conv_wgrad(xin, gpre, sc.conv.weight)
            del xin
        L["xin"] = None
        layers[idx - 1]["xo"] = None
        _, wd = packs.conv(sc.conv.weight)
        gxin = conv3x3(gpre, wd, c, ci)                       # data gradient: same kernel, flipped pack
        if j == 0:
            gfeat = _bf16(B, r // 2, r // 2, ci, device=dev)
            call("bg_upsample2x_bwd", gxin, gfeat, B, r // 2, r // 2, ci)
            if g_prev_extra is not None and k == steps - 1:
                gfeat = axpby(gfeat, g_prev_extra, 1.0, 1.0)
            gx = gfeat
        else:
            gx = gxin

    # style FCs of all layers (gan.py:60,66), grouped per latent: weight / bias gradients and the latent gradient
    for wsel in range(len(maps)):
        grp = [(fc, d) for (w_, fc, d) in dstyles if w_ == wsel]
        if not grp:
            continue
        fcs, gys = [fc for fc, _ in grp], [d for _, d in grp]
        ns, cf = [fc.weight.shape[0] for fc in fcs], [coef_of(fc.weight) for fc in fcs]
        wanted = [fc for fc in fcs if want(fc.weight) or want(fc.bias)]
        if wanted:
            sel = [i for i, fc in enumerate(fcs) if want(fc.weight) or want(fc.bias)]
            dws = [_f32(*fcs[i].weight.shape, device=dev) for i in sel]
            dbs = [_f32(ns[i], device=dev) for i in sel]
            call("bg_linear_bwd_weight_grouped", [maps[wsel][-1]] * len(sel), [gys[i] for i in sel], dws, dbs, [ns[i] for i in sel],
                 [cf[i] for i in sel], len(sel), B, 512)
            for i, dw, db in zip(sel, dws, dbs):
                if want(fcs[i].weight):
                    grads[id(fcs[i].weight)] = dw
                if want(fcs[i].bias):
                    grads[id(fcs[i].bias)] = db
        gw = _f32(B, 512, device=dev)
        call("bg_linear_bwd_input_grouped", gys, [fc.weight.detach() for fc in fcs], ns, cf, len(fcs), B, 512, gw)
        g_w[wsel] = gw

    # mapping network backward (gan.py:130-148)
    dzs = [None, None]
    mgrads: Dict[int, torch.Tensor] = {}          # both latents accumulate here; handed over (emitted) when complete
    passes = list(enumerate(maps))
    if tape.get("maps_cat") is not None:
        # both latents were mapped as one batch of 2B rows: one backward pass over it
        g_w = [torch.cat([g if g is not None else torch.zeros(B, 512, device=dev) for g in g_w])]
        passes = [(0, tape["maps_cat"])]
    for which, hs in passes:
        g = g_w[which]
        if g is None:
            g = torch.zeros(B, 512, device=dev)
        need_in = (need_z or need_z2) if tape.get("maps_cat") is not None else (need_z if which == 0 else need_z2)
        pending = []                                  # (layer, gated gradient, layer input) of a single pass
        for i in reversed(range(8)):
            lin = gen.to_w_noise[0].layers[i][0]
            gp = gate_f32(g, hs[i + 1])
            if want(lin.weight) or want(lin.bias):
                if len(passes) == 1:
                    pending.append((lin, gp, hs[i]))
                else:
                    acc = id(lin.weight) in mgrads
                    dw, db = linear_bwd_weight(gp, hs[i], lin.weight, into=mgrads.get(id(lin.weight)),
                                               into_b=mgrads.get(id(lin.bias)))
                    if not acc:
                        mgrads[id(lin.weight)] = dw
                        mgrads[id(lin.bias)] = db
            if i > 0 or need_in:
                g = linear_bwd_input(gp, lin.weight, packs)
        if pending:
            # the weight gradients of the whole mapping network in one launch (each layer is a 1 MB output from a
            # 64-row batch: eight separate launches are latency, not work)
            dws = [_f32(*lin.weight.shape, device=dev) for lin, _, _ in pending]
            dbs = [_f32(lin.weight.shape[0], device=dev) for lin, _, _ in pending]
            call("bg_linear_bwd_weight_grouped", [x for _, _, x in pending], [gp for _, gp, _ in pending], dws, dbs,
                 [lin.weight.shape[0] for lin, _, _ in pending], [coef_of(lin.weight) for lin, _, _ in pending],
                 len(pending), pending[0][1].shape[0], 512)
            for (lin, _, _), dw, db in zip(pending, dws, dbs):
                mgrads[id(lin.weight)] = dw
                mgrads[id(lin.bias)] = db
        dzs[which] = g if need_in else None
    if tape.get("maps_cat") is not None and dzs[0] is not None:
        both = dzs[0]
        dzs = [both[:B].contiguous() if need_z else None, both[B:].contiguous() if need_z2 else None]
    for key, val in mgrads.items():
        grads[key] = val
    return grads, dzs[0], dzs[1]


# ======================================================================================================
# Critic
# ======================================================================================================
def critic_params(critic, steps: int, fade: bool) -> List[torch.nn.Parameter]:
    start = 8 - steps
    ps = [critic.from_rgbs[start][0].weight, critic.from_rgbs[start][0].bias]
    if fade:
        ps += [critic.from_rgbs[start + 1][0].weight, critic.from_rgbs[start + 1][0].bias]
    for k in range(start, 8):
        blk = critic.conv_blocks[k]
        if k < 7:
            ps += [blk.conv_1[0].weight, blk.conv_1[0].bias, blk.conv_2[0].weight, blk.conv_2[0].bias]
        else:
            ps += [blk.conv_1[1].weight, blk.conv_1[1].bias, blk.conv_2[0].weight, blk.conv_2[0].bias,
                   blk.conv_2[3].weight, blk.conv_2[3].bias, blk.conv_2[5].weight, blk.conv_2[5].bias]
    return ps


def critic_forward(critic, packs: PackCache, images, steps, alpha):
    """Critic.forward (gan.py:331-349).  Returns (pred (B,1) fp32, tape)."""
    dev = images.device
    img = images.detach().float().contiguous()
    B, _, R, _ = img.shape
    start = 8 - steps
    fade = alpha is not None and steps > 1
    a_mix = clamp_alpha(alpha) if fade else None
    packs.refresh_convs([(cv.weight, None) for k in range(start, 7)
                         for cv in (critic.conv_blocks[k].conv_1[0], critic.conv_blocks[k].conv_2[0])]
                        + [(critic.conv_blocks[7].conv_1[1].weight, MBSTD_CPAD)])
    fr = critic.from_rgbs[start][0]
    c0 = CRITIC_CHANNELS[start][0]
    x0 = _bf16(B, R, R, c0, device=dev)
    call("bg_planes3_to_nhwc", img, fr.weight.detach(), fr.bias.detach(), None, x0, B * R * R, R * R, c0, 3, 1,
         coef_of(fr.weight), 1, SLOPE)                                                        # gan.py:338,351-355
    feat = x0
    blocks = []
    for idx, k in enumerate(range(start, 7)):
        blk = critic.conv_blocks[k]
        cin, cout = CRITIC_CHANNELS[k]
        c1, c2 = blk.conv_1[0], blk.conv_2[0]
        wf1, _ = packs.conv(c1.weight)
        wf2, _ = packs.conv(c2.weight)
        y1 = conv3x3(feat, wf1, cin, cout, bias=c1.bias.detach(), act=True)                   # gan.py:254-255
        y2 = conv3x3_pool(y1, wf2, cout, cout, bias=c2.bias.detach(), act=True,               # gan.py:259-261
                          w16=packs.conv_pool4(c2.weight) if R >= 32 else None)
        e = dict(k=k, x=feat, y1=y1, y2=y2, R=R, cin=cin, cout=cout, c1=c1, c2=c2)
        if idx == 0 and fade:
            imgp = _f32(B, 3, R // 2, R // 2, device=dev)
            call("bg_img_avgpool2", img, imgp, B * 3, R // 2, R // 2)                         # gan.py:345
            fr2 = critic.from_rgbs[start + 1][0]
            d = _bf16(B, R // 2, R // 2, cout, device=dev)
            call("bg_planes3_to_nhwc", imgp, fr2.weight.detach(), fr2.bias.detach(), None, d, B * R * R // 4,
                 R * R // 4, cout, 3, 1, coef_of(fr2.weight), 1, SLOPE)
            feat = axpby(d, y2, 1.0 - a_mix, a_mix)                                           # gan.py:347
            e.update(d=d, imgp=imgp, fr2=fr2)
        else:
            feat = y2
        blocks.append(e)
        R //= 2
    # ---- final block (gan.py:237-251)
    blk = critic.conv_blocks[7]
    mb = blk.conv_1[0]
    if B % mb.group_size != 0:                    # the reference mutates the module here (gan.py:277-278)
        mb.group_size = B
    G = mb.group_size
    x7 = feat
    plane = _f32(B // G, device=dev)
    xpad = _bf16(B, 4, 4, MBSTD_CPAD, device=dev)
    call("bg_mbstd_fwd", x7, None, plane, xpad, B, G, 16, 512, MBSTD_CPAD, MBSTD_EPS)
    c1 = blk.conv_1[1]
    wf, _ = packs.conv(c1.weight, cin_pad=MBSTD_CPAD)
    y = conv3x3(xpad, wf, MBSTD_CPAD, 512, bias=c1.bias.detach(), act=True)
    yf = _f32(B, 512 * 16, device=dev)
    call("bg_nhwc_to_nchw_f32", y, yf, B, 16, 512)
    c4, l3, l5 = blk.conv_2[0], blk.conv_2[3], blk.conv_2[5]
    h1 = linear_fwd(yf, c4.weight, c4.bias, act=True)          # 4x4 valid conv == linear over (c,h,w), gan.py:245
    h2 = linear_fwd(h1, l3.weight, l3.bias, act=True)          # gan.py:248-249
    pred = linear_fwd(h2, l5.weight, l5.bias, act=False)       # gan.py:250
    tape = dict(img=img, x0=x0, fr=fr, blocks=blocks, x7=x7, xpad=xpad, y=y, yf=yf, h1=h1, h2=h2, G=G, B=B,
                steps=steps, fade=fade, a_mix=a_mix, c1=c1, c4=c4, l3=l3, l5=l5, R=img.shape[2])
    return pred, tape


def critic_tangent(critic, packs: PackCache, tape, v_img):
    """Forward-mode pass of the critic along an image-space direction v (bias-free, LeakyReLU gates taken
    from the saved activations).  Returns the tangents at the INPUT of every weight layer; they pair with
    the gated ones-backprop (`ghat`) in the R1 weight gradient  dP/dW_l = wgrad(v_{l-1}, ghat_l)."""
    dev = v_img.device
    B, R = tape["B"], tape["R"]
    T = dict(img=v_img, blocks=[])
    fr = tape["fr"]
    x0 = tape["x0"]
    c0 = x0.shape[3]
    v = _bf16(B, R, R, c0, device=dev)
    call("bg_planes3_to_nhwc", v_img, fr.weight.detach(), None, x0, v, B * R * R, R * R, c0, 3, 1,
         coef_of(fr.weight), 0, SLOPE)
    for idx, e in enumerate(tape["blocks"]):
        cin, cout, r = e["cin"], e["cout"], e["R"]
        wf1, _ = packs.conv(e["c1"].weight)
        wf2, _ = packs.conv(e["c2"].weight)
        t = dict(x=v)
        v1 = conv3x3(v, wf1, cin, cout, gate_src=e["y1"])
        t["y1"] = v1
        v2 = conv3x3_pool(v1, wf2, cout, cout, gate_src=e["y2"], act=False,
                          w16=packs.conv_pool4(e["c2"].weight) if r >= 32 else None)
        if "d" in e:
            a_mix = tape["a_mix"]
            vp = _f32(B, 3, r // 2, r // 2, device=dev)
            call("bg_img_avgpool2", v_img, vp, B * 3, r // 2, r // 2)
            fr2 = e["fr2"]
            vd = _bf16(B, r // 2, r // 2, cout, device=dev)
            call("bg_planes3_to_nhwc", vp, fr2.weight.detach(), None, e["d"], vd, B * r * r // 4, r * r // 4, cout, 3,
                 1, coef_of(fr2.weight), 0, SLOPE)
            t["imgp"] = vp
            v = axpby(vd, v2, 1.0 - a_mix, a_mix)
        else:
            v = v2
        T["blocks"].append(t)
    T["x7"] = v
    G = tape["G"]
    sdot = _f32(B // G, device=dev)
    vpad = _bf16(B, 4, 4, MBSTD_CPAD, device=dev)
    call("bg_mbstd_fwd", tape["x7"], v, sdot, vpad, B, G, 16, 512, MBSTD_CPAD, MBSTD_EPS)
    T["xpad"] = vpad
    wf, _ = packs.conv(tape["c1"].weight, cin_pad=MBSTD_CPAD)
    vy = conv3x3(vpad, wf, MBSTD_CPAD, 512, gate_src=tape["y"])
    vyf = _f32(B, 512 * 16, device=dev)
    call("bg_nhwc_to_nchw_f32", vy, vyf, B, 16, 512)
    T["yf"] = vyf
    vh1 = gate_f32(linear_fwd(vyf, tape["c4"].weight, None, act=False), tape["h1"])
    T["h1"] = vh1
    vh2 = gate_f32(linear_fwd(vh1, tape["l3"].weight, None, act=False), tape["h2"])
    T["h2"] = vh2
    return T


class _EmitDict(dict):
    """Gradient dict that reports every entry the moment it is stored (its kernels are already queued on the stream):
    lets the caller finish and all-reduce a layer's gradient while the backward pass is still running."""

    def __init__(self, emit):
        super().__init__()
        self._emit = emit

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self._emit(key, value)


def critic_backward(critic, packs: PackCache, tape, g_pred, need: Dict[int, bool], need_img: bool,
                    keep: Optional[dict] = None, r1: Optional[tuple] = None, emit=None, acc: Optional[dict] = None):
    """See _critic_backward; runs it inside one pre-zeroed slab for all of the pass's accumulators."""
    params = critic_params(critic, tape["steps"], tape["fade"])
    with ZeroSlab(g_pred.device, 16384 + _wgrad_scratch_floats(params, need)):
        return _critic_backward(critic, packs, tape, g_pred, need, need_img, keep, r1, emit, acc)


def _critic_backward(critic, packs: PackCache, tape, g_pred, need: Dict[int, bool], need_img: bool,
                     keep: Optional[dict] = None, r1: Optional[tuple] = None, emit=None, acc: Optional[dict] = None):
    """Reverse pass of critic_forward seeded with g_pred (B,1).

    need[id(p)] -> produce that parameter gradient.  keep: dict that receives the gated gradient at every
    weight layer's output (the `ghat` chain of the R1 penalty when seeded with ones).  r1=(T, ghat): adds
    the second-order R1 terms — wgrad(v, ghat) into every weight gradient and the minibatch-stddev
    curvature into the activation gradient (SURVEY.md §7 hard part 1).
    Returns (grads {id(param): tensor}, g_img or None).
    """
    dev = g_pred.device
    B, G = tape["B"], tape["G"]
    grads: Dict[int, torch.Tensor] = {} if emit is None else _EmitDict(emit)
    T, H = r1 if r1 is not None else (None, None)
    acc = acc or {}            # {id(conv weight): gradient of another branch}: this pass's result is added INTO it by the
                               # weight-gradient unpack kernel (no separate add pass over the 9.4 MB tensors)

    def want(p):
        return need.get(id(p), False)

    def lin_grads(layer, gy, x, gy2=None, x2=None):
        if want(layer.weight) or want(layer.bias):
            dw, db = linear_bwd_weight(gy, x, layer.weight)
            if gy2 is not None:                   # R1: + ghat^T v   (no bias term: the tangent pass is bias-free)
                linear_bwd_weight(gy2, x2, layer.weight, into=dw, into_b=None)
            if want(layer.weight):
                grads[id(layer.weight)] = dw
            if want(layer.bias):
                grads[id(layer.bias)] = db

    g_pred = g_pred.detach().float().contiguous()
    c1, c4, l3, l5 = tape["c1"], tape["c4"], tape["l3"], tape["l5"]
    # ---- head: Linear(512,1) <- LReLU <- Linear(512,512) <- LReLU <- conv4x4 (gan.py:244-251)
    lin_grads(l5, g_pred, tape["h2"], H["pred"] if H else None, T["h2"] if T else None)
    gh2 = gate_f32(linear_bwd_input(g_pred, l5.weight, packs), tape["h2"])
    lin_grads(l3, gh2, tape["h1"], H["h2"] if H else None, T["h1"] if T else None)
    gh1 = gate_f32(linear_bwd_input(gh2, l3.weight, packs), tape["h1"])
    lin_grads(c4, gh1, tape["yf"], H["h1"] if H else None, T["yf"] if T else None)
    gyf = linear_bwd_input(gh1, c4.weight, packs)                                   # (B, 8192) NCHW order
    gy = _bf16(B, 4, 4, 512, device=dev)
    call("bg_nchw_f32_to_nhwc", gyf, tape["y"], gy, B, 16, 512, SLOPE)             # + LReLU gate of conv_1
    if keep is not None:
        keep.update(pred=g_pred, h2=gh2, h1=gh1, y=gy)
    # ---- conv_1 of the final block on the 513(->576)-channel input
    if want(c1.weight):
        grads[id(c1.weight)] = conv_wgrad(tape["xpad"], gy, c1.weight,
                                          extra=(T["xpad"], H["y"]) if T else None, into=acc.get(id(c1.weight)))
    if want(c1.bias):
        grads[id(c1.bias)] = channel_wsum(gy, None, 0, 16, 0, 0)[0].clone()
    _, wd = packs.conv(c1.weight, cin_pad=MBSTD_CPAD)
    gxpad = conv3x3(gy, wd, 512, MBSTD_CPAD)
    if keep is not None:
        keep["xpad"] = gxpad
    gs_ws = _f32(2 * (B // G), device=dev)
    gx = _bf16(B, 4, 4, 512, device=dev)
    call("bg_mbstd_bwd", tape["x7"], T["x7"] if T else None, gxpad, H["xpad"] if H else None, gs_ws, gx, B, G, 16,
         512, MBSTD_CPAD, MBSTD_EPS)

    g_img = None
    if need_img:
        g_img = torch.zeros_like(tape["img"])
    blocks = tape["blocks"]
    gx_gated = False          # the gradient entering the topmost block comes from the mbstd layer: not gated yet
    for idx in reversed(range(len(blocks))):
        e = blocks[idx]
        t = T["blocks"][idx] if T else None
        h = H["blocks"][idx] if H else None
        kk = dict() if keep is not None else None
        cin, cout, r = e["cin"], e["cout"], e["R"]
        c1b, c2b = e["c1"], e["c2"]
        if "d" in e:                                             # fade-in lerp (gan.py:342-347)
            a_mix = tape["a_mix"]
            fr2 = e["fr2"]
            gd = axpby(gx, None, 1.0 - a_mix, 0.0)
            gdp = torch.empty_like(gd)
            call("bg_act_gate", gd, e["d"], gdp, gd.numel(), SLOPE)
            if kk is not None:
                kk["d"] = gdp
            hw = (r // 2) * (r // 2)
            if want(fr2.weight) or want(fr2.bias):
                ws = channel_wsum(gdp, e["imgp"], 3, hw, 3 * hw, hw)
                dw = ws[1:4]
                if t is not None:
                    dw = dw + channel_wsum(h["d"], t["imgp"], 3, hw, 3 * hw, hw)[1:4]
                if want(fr2.weight):
                    grads[id(fr2.weight)] = (dw * coef_of(fr2.weight)).t().reshape(cout, 3, 1, 1).contiguous()
                if want(fr2.bias):
                    grads[id(fr2.bias)] = ws[0].clone()
            if need_img:
                gp_img = _f32(B, 3, r // 2, r // 2, device=dev)
                call("bg_nhwc_to_planes3", gdp, fr2.weight.detach(), None, gp_img, B * hw, hw, cout, 3, 1,
                     coef_of(fr2.weight))
                call("bg_img_avgpool2_bwd", gp_img, g_img, B * 3, r // 2, r // 2, 1.0, 1)
            gy2 = axpby(gx, None, a_mix, 0.0)
        else:
            gy2 = gx
        db1 = _acc(cout, device=dev) if want(c1b.bias) else None                               # bias grad of conv_1
        gu = None
        if r >= 32:
            # conv_2 -> pool is handled as ONE 4x4 stride-2 conv in all three directions: everything works from the
            # gated gradient at the POOLED map, the full-resolution pool adjoint is never materialised
            if gx_gated:
                gpool = gy2                                                                     # already gated above
            else:
                gpool = torch.empty_like(gy2)
                call("bg_act_gate", gy2, e["y2"], gpool, gy2.numel(), SLOPE)                    # LReLU adjoint
            if kk is not None:
                kk["gp"] = gpool
            if want(c2b.weight):
                grads[id(c2b.weight)] = conv_pool_wgrad(e["y1"], gpool, c2b.weight,
                                                        extra=(t["y1"], h["gp"]) if t else None, into=acc.get(id(c2b.weight)))
            if want(c2b.bias):
                grads[id(c2b.bias)] = channel_wsum(gpool, None, 0, (r // 2) * (r // 2), 0, 0)[0]
            g1 = _bf16(B, r, r, cout, device=dev)
            call("bg_conv_pool4_dgrad", gpool, packs.conv_tconv4(c2b.weight), g1, B, r // 2, r // 2, cout, cout, e["y1"],
                 SLOPE, db1)                                                                    # dgrad + LReLU gate
            del gpool
        else:
            gu = _bf16(B, r, r, cout, device=dev)
            db2 = _acc(cout, device=dev) if want(c2b.bias) else None                           # bias grad of conv_2
            call("bg_pool_act_bwd", gy2, e["y2"], gu, B, r // 2, r // 2, cout, SLOPE, db2)     # pool + LReLU adjoint
            if kk is not None:
                kk["u"] = gu
            if want(c2b.weight):
                grads[id(c2b.weight)] = conv_wgrad(e["y1"], gu, c2b.weight, extra=(t["y1"], h["u"]) if t else None,
                                                   into=acc.get(id(c2b.weight)))
            if db2 is not None:
                grads[id(c2b.bias)] = db2
            _, wd2 = packs.conv(c2b.weight)
            if db1 is not None:
                g1, db1 = conv3x3(gu, wd2, cout, cout, gate_src=e["y1"], stats=2)
            else:
                g1 = conv3x3(gu, wd2, cout, cout, gate_src=e["y1"])
        if db1 is not None:
            grads[id(c1b.bias)] = db1
        del gu
        if kk is not None:
            kk["y1"] = g1
        if want(c1b.weight):
            grads[id(c1b.weight)] = conv_wgrad(e["x"], g1, c1b.weight, extra=(t["x"], h["y1"]) if t else None,
                                               into=acc.get(id(c1b.weight)))
        _, wd1 = packs.conv(c1b.weight)
        first = idx == 0
        if first:
            gx = conv3x3(g1, wd1, cout, cin, gate_src=tape["x0"])                              # gate of fromRGB's LReLU
            gx_gated = False
        else:
            # this block's input IS the block below's pooled output y2: when that block takes the folded path (and there
            # is no fade-in lerp in between) its LeakyReLU adjoint is applied right here, in this dgrad's epilogue
            below = blocks[idx - 1]
            gx_gated = ("d" not in below) and below["R"] >= 32
            gx = conv3x3(g1, wd1, cout, cin, gate_src=e["x"] if gx_gated else None)
        del g1
        if keep is not None:
            keep.setdefault("blocks", [None] * len(blocks))[idx] = kk
    # ---- fromRGB (gan.py:351-355); when steps == 1 gx is the (ungated) gradient at x7 = x0
    fr = tape["fr"]
    x0 = tape["x0"]
    R = tape["R"]
    c0 = x0.shape[3]
    if not blocks:
        g0 = torch.empty_like(gx)
        call("bg_act_gate", gx, x0, g0, gx.numel(), SLOPE)
    else:
        g0 = gx
    if keep is not None:
        keep["x0"] = g0
    if want(fr.weight) or want(fr.bias):
        ws = channel_wsum(g0, tape["img"], 3, R * R, 3 * R * R, R * R)
        dw = ws[1:4]
        if T is not None:
            dw = dw + channel_wsum(H["x0"], T["img"], 3, R * R, 3 * R * R, R * R)[1:4]
        if want(fr.weight):
            grads[id(fr.weight)] = (dw * coef_of(fr.weight)).t().reshape(c0, 3, 1, 1).contiguous()
        if want(fr.bias):
            grads[id(fr.bias)] = ws[0].clone()
    if need_img:
        gi = _f32(B, 3, R, R, device=dev)
        call("bg_nhwc_to_planes3", g0, fr.weight.detach(), None, gi, B * R * R, R * R, c0, 3, 1, coef_of(fr.weight))
        call("bg_axpby_f32", g_img, gi, g_img, gi.numel(), 1.0, 1.0)
    return grads, g_img


def _penalised_critic_step(critic, packs: PackCache, first_order, pen_tape, pen_seed, make_v0, emit=None):
    """Shared schedule of the two gradient-penalty critic losses (R1: gan.py:393-412, WGAN-GP: gan.py:357-391).

    first_order: [(tape, seed (B,1))] — ordinary backward passes whose parameter gradients are summed in.
    pen_tape:    the forward whose input-gradient g_x = d sum D(x) / d x is penalised; pen_seed (B,1) is ITS first-order
                 seed (zeros when the loss has no first-order term on that forward).
    make_v0(g_x) -> v0: the image-space direction (d penalty / d g_x) the second-order pass is taken along.
    The double-backward autograd would run is scheduled by hand:
      1. `ghat` = gated backprop of ones down to the image -> g_x;  v0 = make_v0(g_x);
      2. tangent forward of v0 through the critic (bias-free, saved LeakyReLU gates);
      3. ONE combined backward seeded with pen_seed whose weight gradients contract over the doubled K
         [x ; v] . [ybar ; ghat]  and whose activation gradient picks up the minibatch-stddev curvature term.
    Returns (grads {id(param): tensor}, g_x)."""
    steps, fade = pen_tape["steps"], pen_tape["fade"]
    params = critic_params(critic, steps, fade)
    need = {id(p): bool(p.requires_grad) for p in params}
    firsts = [critic_backward(critic, packs, tape, seed, need, need_img=False)[0] for tape, seed in first_order]
    B = pen_tape["B"]
    dev = pen_tape["img"].device
    ones = torch.ones(B, 1, device=dev)
    ghat: dict = {}
    _, g_x = critic_backward(critic, packs, pen_tape, ones, {}, need_img=True, keep=ghat)    # gan.py:398-400 / 375-381
    v0 = make_v0(g_x)
    T = critic_tangent(critic, packs, pen_tape, v0)
    grads = {}
    by_id = {id(p): p for p in params}

    def finish(k, gr):
        """A penalised-branch gradient just became final: add the first-order parts and hand the total out."""
        if not need.get(k, False):
            return
        for gf in firsts:
            part = gf.get(k)
            if part is None:
                raise RuntimeError("internal: missing critic gradient")
            if part.data_ptr() != gr.data_ptr():         # conv weights were accumulated in place by the unpack kernel
                call("bg_axpby_f32", part, gr, part, part.numel(), 1.0, 1.0)
            gr = part
        grads[k] = gr
        if emit is not None:
            emit(by_id[k], gr)

    # the last pass runs head -> high resolution: the big low-resolution weights finish first, so their all-reduce
    # (emit) overlaps the expensive high-resolution layers still to come
    conv_acc = None
    if len(firsts) == 1:
        conv_acc = {k: v for k, v in firsts[0].items() if v.dim() == 4 and v.shape[-1] == 3}
    critic_backward(critic, packs, pen_tape, pen_seed, need, need_img=False, r1=(T, ghat), emit=finish, acc=conv_acc)
    for p in params:
        if need[id(p)] and id(p) not in grads:
            raise RuntimeError("internal: missing critic gradient")
    return grads, g_x


def critic_r1_step(critic, packs: PackCache, tape_fake, pred_fake, tape_real, pred_real, c_lambda, emit=None,
                   first_order=True):
    """Critic.get_r1_loss (gan.py:393-412): loss value + gradients of every active critic parameter.

    loss = mean softplus(-D(real)) + mean softplus(D(fake)) + lambda/2 * mean_n ||d sum D(real) / d real_n||^2
    fake branch: ordinary backward seeded with sigmoid(D(fake))/B; real branch: _penalised_critic_step with
    v0 = (lambda/B) g_x and the first-order seed -sigmoid(-D(real))/B.  first_order=False keeps ONLY the penalty term
    (loss and gradients): the purely second-order part, checked in isolation by the parity tests.
    Returns (loss tensor (), grads {id(param): tensor}, g_x).
    """
    dev = pred_real.device
    B = pred_real.shape[0]
    terms = _f32(3, device=dev) if first_order else torch.zeros(3, device=dev)
    seed_f = _f32(B, 1, device=dev)
    seed_r = _f32(B, 1, device=dev)
    if first_order:
        pf = pred_fake.detach().float().contiguous()
        pr = pred_real.detach().float().contiguous()
        call("bg_logistic_loss", pf, B, 1.0, terms[0:1], seed_f, 1.0)         # softplus(D(fake)).mean(), gan.py:406
        call("bg_logistic_loss", pr, B, -1.0, terms[1:2], seed_r, 1.0)        # softplus(-D(real)).mean(), gan.py:396
    else:
        seed_r.zero_()

    def make_v0(g_x):
        call("bg_sumsq", g_x, g_x.numel(), float(c_lambda) / 2.0 / B, terms[2:3])            # gan.py:401-404
        v0 = torch.empty_like(g_x)
        call("bg_axpby_f32", g_x, None, v0, g_x.numel(), float(c_lambda) / B, 0.0)
        return v0

    grads, g_x = _penalised_critic_step(critic, packs, [(tape_fake, seed_f)] if first_order else [], tape_real, seed_r,
                                        make_v0, emit=emit)
    return terms.sum(), grads, g_x


def critic_wgan_gp_step(critic, packs: PackCache, tape_fake, pred_fake, tape_real, pred_real, tape_mixed, c_lambda,
                        emit=None):
    """Critic.get_wgan_loss as gan.py:357-391 intends it (the reference's body cannot run, see gan.Critic.get_wgan_loss):
    loss = -mean D(real) + mean D(fake) + lambda * mean_n (||d sum D(mixed) / d mixed_n||_2 - 1)^2.
    fake / real branches: ordinary backward passes seeded with +1/B, -1/B; mixed branch: _penalised_critic_step with
    v0_n = (lambda/B) * 2 (r_n - 1) / r_n * g_n, r_n = ||g_n||  and no first-order seed.
    Returns (loss tensor (), grads {id(param): tensor}, g_mixed)."""
    dev = pred_real.device
    B = pred_real.shape[0]
    pen = _f32(1, device=dev)
    seed_f = torch.full((B, 1), 1.0 / B, device=dev)
    seed_r = torch.full((B, 1), -1.0 / B, device=dev)

    def make_v0(g_x):
        v0 = torch.empty_like(g_x)
        call("bg_gp_rows", g_x, B, g_x.numel() // B, float(c_lambda) / B, float(c_lambda) / B, pen, v0)   # gan.py:385
        return v0

    grads, g_x = _penalised_critic_step(critic, packs, [(tape_fake, seed_f), (tape_real, seed_r)], tape_mixed,
                                        torch.zeros(B, 1, device=dev), make_v0, emit=emit)
    loss = pred_fake.detach().float().mean() - pred_real.detach().float().mean() + pen[0]    # gan.py:387
    return loss, grads, g_x
