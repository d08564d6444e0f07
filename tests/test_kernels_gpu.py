"""Kernel-level parity (GPU): every C-ABI entry point against a plain torch fp32 restatement of the
reference op it replaces (gan.py line cited per test).  bf16 operands / fp32 accumulate, so the
tolerance is stated per test: outputs are compared after rounding the *inputs* to bf16 exactly as the
kernel sees them, leaving only accumulation-order and output-rounding differences.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import bg_native as bgn  # noqa: E402

DEV = "cuda"


def nhwc(x):  # (N,C,H,W) fp32 -> (N,H,W,C) bf16 contiguous
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):  # (N,H,W,C) bf16 -> (N,C,H,W) fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def relerr(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def pack(w, coef, cin_pad=None):
    co, ci, ks, _ = w.shape
    cin_pad = cin_pad or ci
    wf = torch.empty(ks * ks, co, cin_pad, dtype=torch.bfloat16, device=DEV)
    wd = torch.empty(ks * ks, cin_pad, co, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight", w, wf, wd, co, ci, cin_pad, ks, float(coef))
    return wf, wd


CONV_SHAPES = [
    # N, H, W, Cin, Cout
    (2, 16, 16, 64, 64),
    (3, 4, 4, 512, 512),
    (8, 4, 4, 64, 128),
    (5, 8, 8, 128, 256),
    (2, 32, 32, 256, 128),
    (1, 64, 64, 32, 32),
    (1, 64, 64, 16, 32),
    (2, 32, 32, 32, 16),
    (1, 128, 128, 16, 16),
    (2, 16, 16, 512, 256),
    (1, 32, 32, 64, 48),
    (3, 16, 16, 128, 128),
    (1, 32, 32, 128, 64),
    (1, 64, 64, 64, 32),
    (2, 32, 32, 16, 16),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_pack_weight(shape):
    _, _, _, ci, co = shape
    torch.manual_seed(1)
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    wf, wd = pack(w, coef)
    ref = (w * coef).to(torch.bfloat16)
    assert torch.equal(wf, ref.permute(2, 3, 0, 1).reshape(9, co, ci))
    assert torch.equal(wd, ref.flip(2, 3).permute(2, 3, 1, 0).reshape(9, ci, co))


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("entry", ["bg_conv_fprop", "bg_conv_fprop_tapwise"])
def test_conv3x3_fprop(shape, fused, entry):
    """EqualizedConv2d.forward (gan.py:29-38) [+ InjectSecondaryNoise gan.py:52 + LeakyReLU gan.py:86]."""
    n, h, w_, ci, co = shape
    torch.manual_seed(0)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    wf, _ = pack(w, coef)
    bias = torch.randn(co, device=DEV) * 0.1 if fused else None
    noise = torch.randn(n, 1, h, w_, device=DEV) if fused else None
    nw = torch.randn(co, device=DEV) * 0.1 if fused else None
    out = torch.empty(n, h, w_, co, dtype=torch.bfloat16, device=DEV)
    bgn.call(entry, x, wf, out, n, h, w_, ci, co, 3, bias, noise, nw, None, 1 if fused else 0, 0.2)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x), (w * coef).to(torch.bfloat16).float(), None, padding=1)
    if fused:
        ref = ref + bias.view(1, -1, 1, 1) + nw.view(1, -1, 1, 1) * noise
        ref = F.leaky_relu(ref, 0.2)
    err = relerr(nchw(out), ref)
    assert err < 6e-3, f"conv3x3 fprop {entry} {shape} fused={fused}: rel-L2 {err:.3e}"


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (1, 64, 64, 16, 32), (4, 8, 8, 256, 128), (2, 32, 32, 32, 16)])
def test_conv3x3_dgrad_via_fprop(shape):
    """autograd convolution_backward w.r.t. input == the forward kernel on the flipped/transposed pack."""
    n, h, w_, ci, co = shape
    torch.manual_seed(0)
    g = nhwc(torch.randn(n, co, h, w_, device=DEV))
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    _, wd = pack(w, coef)
    gx = torch.empty(n, h, w_, ci, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_conv_fprop", g, wd, gx, n, h, w_, co, ci, 3, None, None, None, None, 0, 0.2)
    torch.cuda.synchronize()
    wq = (w * coef).to(torch.bfloat16).float()
    ref = torch.nn.grad.conv2d_input((n, ci, h, w_), wq, nchw(g), padding=1)
    err = relerr(nchw(gx), ref)
    assert err < 6e-3, f"dgrad {shape}: rel-L2 {err:.3e}"


STATS_SHAPES = CONV_SHAPES + [(3, 64, 64, 64, 128), (5, 32, 32, 64, 64), (2, 64, 64, 32, 256)]


@pytest.mark.parametrize("shape", STATS_SHAPES)
@pytest.mark.parametrize("mode", [1, 2])
def test_conv3x3_fprop_fused_stats(shape, mode):
    """Epilogue-fused reductions: mode 1 = nn.InstanceNorm2d statistics of the produced activation (gan.py:59,69),
    mode 2 = per-channel sum over batch and pixels (the bias gradient when the pass is a dgrad).  The map itself
    must be bit-identical to the plain entry; the sums are compared with fp64 sums of that bf16 map."""
    n, h, w_, ci, co = shape
    torch.manual_seed(3)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    w = torch.randn(co, ci, 3, 3, device=DEV)
    wf, _ = pack(w, math.sqrt(2 / (ci * 9)))
    bias = torch.randn(co, device=DEV) * 0.1
    noise = torch.randn(n, 1, h, w_, device=DEV)
    nw = torch.randn(co, device=DEV) * 0.1
    ref = torch.empty(n, h, w_, co, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_conv_fprop", x, wf, ref, n, h, w_, ci, co, 3, bias, noise, nw, None, 1, 0.2)
    out = torch.empty_like(ref)
    stats = torch.full((n, co, 2) if mode == 1 else (co,), 7.0, device=DEV)     # the call must zero it
    bgn.call("bg_conv_fprop_stats", x, wf, out, n, h, w_, ci, co, 3, bias, noise, nw, None, 1, 0.2, stats, mode)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    o = out.double()
    if mode == 1:
        want = torch.stack([o.sum(dim=(1, 2)), (o * o).sum(dim=(1, 2))], dim=-1)
    else:
        want = o.sum(dim=(0, 1, 2))
    err = (stats.double() - want).abs().max().item()
    scale = want.abs().max().item() + 1.0
    assert err < 2e-5 * scale * math.sqrt(h * w_), (err, scale)


@pytest.mark.parametrize("shape", [s for s in CONV_SHAPES if s[4] != 48])
@pytest.mark.parametrize("entry", ["bg_conv_wgrad", "bg_conv_wgrad_tapwise"])
def test_conv3x3_wgrad(shape, entry):
    """autograd convolution_backward w.r.t. weight (and the weight half of the R1 double-backward)."""
    n, h, w_, ci, co = shape
    torch.manual_seed(0)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    g = nhwc(torch.randn(n, co, h, w_, device=DEV))
    dwp = torch.empty(9, co, ci, dtype=torch.float32, device=DEV)
    bgn.call(entry, x, g, dwp, n, h, w_, ci, co, 0)
    bgn.call(entry, x, g, dwp, n, h, w_, ci, co, 1)          # accumulate=1: the R1 doubled-K form adds into dWp
    dwp *= 0.5
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(nchw(x), (co, ci, 3, 3), nchw(g), padding=1)
    got = dwp.reshape(3, 3, co, ci).permute(2, 3, 0, 1)
    err = relerr(got, ref)
    assert err < 2e-3, f"wgrad {shape}: rel-L2 {err:.3e}"


def test_upsample_pool_adain_aux():
    """nn.Upsample bilinear x2 (gan.py:112), AvgPool2d+LeakyReLU (gan.py:260-261), AdaIN (gan.py:65-71)."""
    torch.manual_seed(0)
    n, c, h, w_ = 3, 32, 8, 8
    xf = torch.randn(n, c, h, w_, device=DEV)
    x = nhwc(xf)
    y = torch.empty(n, 2 * h, 2 * w_, c, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_upsample2x_fwd", x, y, n, h, w_, c)
    ref = F.interpolate(nchw(x), scale_factor=2, mode="bilinear")
    assert relerr(nchw(y), ref) < 4e-3
    # adjoint
    gy = nhwc(torch.randn(n, c, 2 * h, 2 * w_, device=DEV))
    gx = torch.empty(n, h, w_, c, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_upsample2x_bwd", gy, gx, n, h, w_, c)
    xr = nchw(x).requires_grad_()
    F.interpolate(xr, scale_factor=2, mode="bilinear").backward(nchw(gy))
    assert relerr(nchw(gx), xr.grad) < 4e-3
    for (n2, c2, h2, w2) in [(2, 8, 4, 4), (1, 64, 4, 16), (2, 16, 32, 8), (1, 512, 16, 16)]:   # strips of 4 rows per thread
        gy2 = nhwc(torch.randn(n2, c2, 2 * h2, 2 * w2, device=DEV))
        gx2 = torch.full((n2, h2, w2, c2), float("nan"), dtype=torch.bfloat16, device=DEV)
        bgn.call("bg_upsample2x_bwd", gy2, gx2, n2, h2, w2, c2)
        xr2 = torch.zeros(n2, c2, h2, w2, device=DEV, requires_grad=True)
        F.interpolate(xr2, scale_factor=2, mode="bilinear").backward(nchw(gy2))
        assert relerr(nchw(gx2), xr2.grad) < 4e-3, (n2, c2, h2, w2)
    # pool + lrelu
    u = nhwc(torch.randn(n, c, 2 * h, 2 * w_, device=DEV))
    yo = torch.empty(n, h, w_, c, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pool_act_fwd", u, None, yo, n, h, w_, c, 0.2, 0)
    ref = F.leaky_relu(F.avg_pool2d(nchw(u), 2), 0.2)
    assert relerr(nchw(yo), ref) < 4e-3
    # instance-norm statistics + AdaIN apply
    stats = torch.empty(n, c, 2, device=DEV)
    a = nhwc(torch.randn(n, c, h, w_, device=DEV) * 2 + 0.5)
    bgn.call("bg_in_stats", a, stats, n, h * w_, c)
    af = nchw(a)
    assert torch.allclose(stats[..., 0], af.sum((2, 3)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[..., 1], (af * af).sum((2, 3)), rtol=1e-4, atol=1e-3)
    style = torch.randn(n, 2 * c, device=DEV)
    xo = torch.empty_like(a)
    bgn.call("bg_adain_apply", a, stats, style, xo, n, h * w_, c, 1e-8)
    ref = style[:, :c, None, None] * F.instance_norm(af, eps=1e-8) + style[:, c:, None, None]
    assert relerr(nchw(xo), ref) < 5e-3


STYLE_SHAPES = [
    # N, H, W (output), Cin, Cout, upsample
    (2, 16, 16, 64, 64, 0), (2, 16, 16, 64, 64, 1), (3, 32, 32, 128, 64, 1), (3, 32, 32, 64, 64, 0),
    (2, 64, 64, 32, 16, 1), (2, 64, 64, 16, 16, 0), (5, 32, 32, 256, 128, 1), (2, 64, 64, 64, 32, 1),
    (2, 32, 32, 512, 256, 1), (1, 128, 128, 32, 32, 0), (4, 16, 16, 128, 256, 0),
]


@pytest.mark.parametrize("shape", STYLE_SHAPES)
def test_style_conv_fused(shape):
    """StyleGanBlock / StyleConvBlock forward (gan.py:89-98,118-127): AdaIN of the previous activation (gan.py:65-71)
    -> [bilinear x2] -> conv3x3 -> + noise -> LeakyReLU, against torch fp32 on the same bf16 activation, and the
    fused instance-norm sums of the result."""
    n, h, w_, ci, co, up = shape
    hi, wi = (h // 2, w_ // 2) if up else (h, w_)
    torch.manual_seed(11)
    a_prev = nhwc(F.leaky_relu(torch.randn(n, ci, hi, wi, device=DEV) * 1.3 + 0.2, 0.2))
    style = torch.cat([1 + 0.3 * torch.randn(n, ci, device=DEV), 0.3 * torch.randn(n, ci, device=DEV)], dim=1).contiguous()
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    bias = torch.randn(co, device=DEV) * 0.1
    noise = torch.randn(n, 1, h, w_, device=DEV)
    nw = torch.randn(co, device=DEV) * 0.1
    # reference in fp32 from the bf16 activation the kernels see
    af = nchw(a_prev)
    xo = style[:, :ci, None, None] * F.instance_norm(af, eps=1e-8) + style[:, ci:, None, None]
    if up:
        xo = F.interpolate(xo, scale_factor=2, mode="bilinear")
    ref = F.leaky_relu(F.conv2d(xo, w * coef, bias, padding=1) + nw.view(1, co, 1, 1) * noise, 0.2)
    # fused path
    stats_prev = torch.empty(n, ci, 2, device=DEV)
    bgn.call("bg_in_stats", a_prev, stats_prev, n, hi * wi, ci)
    wmod = torch.empty(n, 9, co, ci, dtype=torch.bfloat16, device=DEV)
    btab = torch.empty(n, 9, co, device=DEV)
    bgn.call("bg_style_modulate", w, bias, stats_prev, style, wmod, btab, n, ci, co, hi * wi, coef, 1e-8)
    out = torch.empty(n, h, w_, co, dtype=torch.bfloat16, device=DEV)
    stats = torch.full((n, co, 2), 5.0, device=DEV)
    bgn.call("bg_conv_style_fprop", a_prev, wmod, btab, out, n, h, w_, ci, co, up, noise, nw, 0.2, stats)
    torch.cuda.synchronize()
    err = relerr(nchw(out), ref)
    assert err < 1.2e-2, err
    # border pixels exercise the per-class bias table: check them separately (they are few, so a wrong class would
    # hide inside the global norm)
    o, r = nchw(out), ref
    for sl in [(slice(None), slice(None), 0), (slice(None), slice(None), -1), (slice(None), slice(None), slice(None), 0),
               (slice(None), slice(None), slice(None), -1)]:
        assert relerr(o[sl], r[sl]) < 1.5e-2, (sl, relerr(o[sl], r[sl]))
    od = out.double()
    want = torch.stack([od.sum(dim=(1, 2)), (od * od).sum(dim=(1, 2))], dim=-1)
    assert (stats.double() - want).abs().max().item() < 2e-5 * (want.abs().max().item() + 1.0) * math.sqrt(h * w_)


def test_to_rgb_adain():
    """to_rgbs[k](AdaIN(a)) (gan.py:172-179 after gan.py:65-71) without the normalised map."""
    torch.manual_seed(2)
    for n, c, h in [(3, 32, 16), (2, 16, 64), (2, 512, 4), (1, 128, 32)]:
        a = nhwc(F.leaky_relu(torch.randn(n, c, h, h, device=DEV) + 0.1, 0.2))
        style = torch.cat([1 + 0.3 * torch.randn(n, c, device=DEV), 0.3 * torch.randn(n, c, device=DEV)], dim=1).contiguous()
        wm = torch.randn(3, c, 1, 1, device=DEV)
        b = torch.randn(3, device=DEV)
        coef = math.sqrt(2 / c)
        stats = torch.empty(n, c, 2, device=DEV)
        bgn.call("bg_in_stats", a, stats, n, h * h, c)
        out = torch.empty(n, 3, h, h, device=DEV)
        bgn.call("bg_to_rgb_adain", a, stats, style, wm, b, out, n, h * h, c, coef, 1e-8)
        torch.cuda.synchronize()
        af = nchw(a)
        xo = style[:, :c, None, None] * F.instance_norm(af, eps=1e-8) + style[:, c:, None, None]
        ref = F.conv2d(xo, wm * coef, b)
        assert relerr(out, ref) < 2e-3, (n, c, h, relerr(out, ref))


@pytest.mark.parametrize("n,c,h", [(3, 16, 8), (2, 32, 64), (5, 64, 16), (1, 256, 8), (2, 512, 4), (3, 1024, 4)])
def test_rgb_1x1_maps(n, c, h):
    """toRGB forward / fromRGB input gradient (gan.py:172-179, 351-355): NHWC bf16 features -> 3 fp32 planes, and the
    reverse map (fromRGB forward with bias + LeakyReLU, toRGB input gradient)."""
    torch.manual_seed(n * c + h)
    x = nhwc(torch.randn(n, c, h, h, device=DEV))
    coef = math.sqrt(2 / c)
    w_to = torch.randn(3, c, 1, 1, device=DEV)
    b = torch.randn(3, device=DEV)
    out = torch.full((n, 3, h, h), float("nan"), device=DEV)
    bgn.call("bg_nhwc_to_planes3", x, w_to, b, out, n * h * h, h * h, c, 1, c, coef)
    xd = nchw(x).double()                                            # fp64 references: cuDNN's fp32 conv may use TF32
    assert relerr(out, F.conv2d(xd, w_to.double() * coef, b.double())) < 1e-5
    w_from = torch.randn(c, 3, 1, 1, device=DEV)
    c3 = math.sqrt(2 / 3)
    bgn.call("bg_nhwc_to_planes3", x, w_from, None, out, n * h * h, h * h, c, 3, 1, c3)
    assert relerr(out, F.conv_transpose2d(xd, w_from.double() * c3)) < 1e-5
    img = torch.randn(n, 3, h, h, device=DEV)
    bc = torch.randn(c, device=DEV)
    y = torch.empty(n, h, h, c, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_planes3_to_nhwc", img, w_from, bc, None, y, n * h * h, h * h, c, 3, 1, c3, 1, 0.2)
    assert relerr(nchw(y), F.leaky_relu(F.conv2d(img.double(), w_from.double() * c3, bc.double()), 0.2)) < 4e-3
    bgn.call("bg_planes3_to_nhwc", img, w_to, None, None, y, n * h * h, h * h, c, 1, c, coef, 0, 0.2)
    assert relerr(nchw(y), F.conv_transpose2d(img.double(), w_to.double() * coef)) < 4e-3


@pytest.mark.parametrize("dims", [(3, 16, 8, 8), (2, 64, 32, 32), (2, 512, 4, 4), (1, 128, 64, 64)])
def test_pool_act_bwd_and_adain_bwd_with_fused_sums(dims):
    """Adjoint of AvgPool2d(2)+LeakyReLU (gan.py:258-262) with the conv bias gradient reduced in the same pass, and the
    instance-norm/AdaIN backward (gan.py:55-71, 94-98) with the conv-bias and noise-weight gradients (gan.py:30,52)."""
    n, c, h, w_ = dims
    torch.manual_seed(5)
    # ---- pool + lrelu adjoint against autograd
    u = nchw(nhwc(torch.randn(n, c, 2 * h, 2 * w_, device=DEV))).requires_grad_()
    y = F.leaky_relu(F.avg_pool2d(u, 2), 0.2)
    gy = nchw(nhwc(torch.randn(n, c, h, w_, device=DEV)))
    y.backward(gy)
    gu = torch.empty(n, 2 * h, 2 * w_, c, dtype=torch.bfloat16, device=DEV)
    csum = torch.full((c,), 3.0, device=DEV)
    bgn.call("bg_pool_act_bwd", nhwc(gy), nhwc(y.detach()), gu, n, h, w_, c, 0.2, csum)
    gu2 = torch.empty_like(gu)
    bgn.call("bg_pool_act_bwd", nhwc(gy), nhwc(y.detach()), gu2, n, h, w_, c, 0.2, None)
    torch.cuda.synchronize()
    assert torch.equal(gu, gu2)
    assert relerr(nchw(gu), u.grad) < 6e-3
    want = u.grad.sum((0, 2, 3))
    assert (csum - want).abs().max().item() < 2e-3 * (want.abs().max().item() + 1.0) * math.sqrt(h * w_ * n)
    # ---- AdaIN backward against autograd of: a -> lrelu -> instance_norm -> gamma * . + beta
    pre = nchw(nhwc(torch.randn(n, c, h, w_, device=DEV) * 1.5 + 0.3)).requires_grad_()
    a = F.leaky_relu(pre, 0.2)
    a_q = nchw(nhwc(a.detach()))                       # the kernels see the bf16 activation
    a_in = a_q.clone().requires_grad_()
    style = torch.randn(n, 2 * c, device=DEV)
    xo = style[:, :c, None, None] * F.instance_norm(a_in, eps=1e-8) + style[:, c:, None, None]
    g = nchw(nhwc(torch.randn(n, c, h, w_, device=DEV)))
    xo.backward(g)
    gate = torch.where(a_q > 0, 1.0, 0.2)
    gpre_ref = a_in.grad * gate
    noise = torch.randn(n, 1, h, w_, device=DEV)
    stats = torch.empty(n, c, 2, device=DEV)
    bgn.call("bg_in_stats", nhwc(a_q), stats, n, h * w_, c)
    bs = torch.empty(n, c, 2, device=DEV)
    bgn.call("bg_adain_bwd_reduce", nhwc(g), nhwc(a_q), stats, bs, n, h * w_, c, 1e-8)
    gpre = torch.empty(n, h, w_, c, dtype=torch.bfloat16, device=DEV)
    ws = torch.full((2, c), -1.0, device=DEV)
    bgn.call("bg_adain_bwd_apply", nhwc(g), nhwc(a_q), stats, style, bs, gpre, n, h * w_, c, 1e-8, 0.2, 1, noise, ws)
    gpre2 = torch.empty_like(gpre)
    bgn.call("bg_adain_bwd_apply", nhwc(g), nhwc(a_q), stats, style, bs, gpre2, n, h * w_, c, 1e-8, 0.2, 1, None, None)
    torch.cuda.synchronize()
    assert torch.equal(gpre, gpre2)
    assert relerr(nchw(gpre), gpre_ref) < 1e-2
    tol = 5e-3 * math.sqrt(h * w_ * n)
    wb, wn = gpre_ref.sum((0, 2, 3)), (gpre_ref * noise).sum((0, 2, 3))
    assert (ws[0] - wb).abs().max().item() < tol * (gpre_ref.abs().max().item())
    assert (ws[1] - wn).abs().max().item() < tol * (gpre_ref.abs().max().item()) * 3


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (1, 64, 64, 32, 64), (3, 32, 32, 128, 128), (1, 32, 32, 256, 256),
                                   (2, 16, 16, 512, 512)])
@pytest.mark.parametrize("tangent", [False, True])
def test_conv3x3_pool_fused(shape, tangent):
    """CriticBlock.conv_2: conv3x3 -> AvgPool2d(2) -> LeakyReLU in one kernel (gan.py:258-262); with tangent=True the
    bias-free R1 tangent form  avg_pool(conv(v)) * gate(saved y2)."""
    n, h, w_, ci, co = shape
    torch.manual_seed(0)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    wf, _ = pack(w, coef)
    out = torch.empty(n, h // 2, w_ // 2, co, dtype=torch.bfloat16, device=DEV)
    conv = F.conv2d(nchw(x), (w * coef).to(torch.bfloat16).float(), None, padding=1)
    if tangent:
        y2 = nhwc(torch.randn(n, co, h // 2, w_ // 2, device=DEV))
        bgn.call("bg_conv_pool_fprop", x, wf, out, n, h, w_, ci, co, None, y2, 0, 0.2)
        ref = F.avg_pool2d(conv, 2) * torch.where(nchw(y2) > 0, 1.0, 0.2)
    else:
        bias = torch.randn(co, device=DEV) * 0.1
        bgn.call("bg_conv_pool_fprop", x, wf, out, n, h, w_, ci, co, bias, None, 1, 0.2)
        ref = F.leaky_relu(F.avg_pool2d(conv + bias.view(1, -1, 1, 1), 2), 0.2)
    torch.cuda.synchronize()
    err = relerr(nchw(out), ref)
    assert err < 6e-3, f"fused conv+pool {shape} tangent={tangent}: rel-L2 {err:.3e}"


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64), (1, 64, 64, 32, 64), (3, 32, 32, 128, 128), (1, 64, 64, 256, 256),
                                   (2, 32, 32, 512, 512), (2, 64, 64, 64, 32), (1, 128, 128, 64, 64)])
@pytest.mark.parametrize("tangent", [False, True])
def test_conv_pool_as_4x4_stride2(shape, tangent):
    """CriticBlock.conv_2 (gan.py:258-262): conv3x3 -> AvgPool2d(2) -> LeakyReLU computed as ONE 4x4 stride-2 conv
    (bg_pack_weight_pool4 + bg_conv_pool4_fprop), against torch's conv2d + avg_pool2d in fp32; also the pack itself."""
    n, h, w_, ci, co = shape
    torch.manual_seed(0)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    w16 = torch.empty(16, co, ci, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight_pool4", w, w16, co, ci, coef)
    w4 = torch.zeros(co, ci, 4, 4, device=DEV)
    for dy in range(2):
        for dx in range(2):
            w4[:, :, dy:dy + 3, dx:dx + 3] += 0.25 * coef * w
    # fp32 summation order may differ by an ulp before the single bf16 rounding: allow one bf16 ulp
    assert (w16.float() - w4.permute(2, 3, 0, 1).reshape(16, co, ci)).abs().max().item() <= 2 ** -8 * w4.abs().max().item()
    out = torch.empty(n, h // 2, w_ // 2, co, dtype=torch.bfloat16, device=DEV)
    conv = F.conv2d(nchw(x), w * coef, None, padding=1)
    if tangent:
        y2 = nhwc(torch.randn(n, co, h // 2, w_ // 2, device=DEV))
        bgn.call("bg_conv_pool4_fprop", x, w16, out, n, h, w_, ci, co, None, y2, 0, 0.2)
        ref = F.avg_pool2d(conv, 2) * torch.where(nchw(y2) > 0, 1.0, 0.2)
    else:
        bias = torch.randn(co, device=DEV) * 0.1
        bgn.call("bg_conv_pool4_fprop", x, w16, out, n, h, w_, ci, co, bias, None, 1, 0.2)
        ref = F.leaky_relu(F.avg_pool2d(conv + bias.view(1, -1, 1, 1), 2), 0.2)
    torch.cuda.synchronize()
    err = relerr(nchw(out), ref)
    assert err < 8e-3, f"conv+pool as 4x4s2 {shape} tangent={tangent}: rel-L2 {err:.3e}"


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64), (1, 64, 64, 32, 64), (3, 32, 32, 128, 128), (1, 64, 64, 256, 256),
                                   (2, 32, 32, 512, 512), (1, 128, 128, 64, 64), (2, 32, 32, 64, 32)])
@pytest.mark.parametrize("gated", [False, True])
def test_conv_pool_dgrad_as_transposed_4x4_stride2(shape, gated):
    """autograd's input gradient of conv3x3 -> AvgPool2d(2) (gan.py:258-260) as the transposed 4x4 stride-2 conv
    (bg_pack_weight_tconv4 + bg_conv_pool4_dgrad), optionally with the LeakyReLU gate of the layer below and that
    layer's bias gradient (sum of the gated gradient) from the epilogue."""
    n, h, w_, ci, co = shape                      # conv input (n, ci, h, w_), pooled output (n, co, h/2, w_/2)
    torch.manual_seed(0)
    y1 = nchw(nhwc(torch.randn(n, ci, h, w_, device=DEV))).requires_grad_()
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    gpool = nhwc(torch.randn(n, co, h // 2, w_ // 2, device=DEV))
    F.avg_pool2d(F.conv2d(y1, w * coef, None, padding=1), 2).backward(nchw(gpool))
    ref = y1.grad
    gate = None
    if gated:
        gate = nhwc(torch.randn(n, ci, h, w_, device=DEV))
        ref = ref * torch.where(nchw(gate) > 0, 1.0, 0.2)
    wt = torch.empty(16, ci, co, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_pack_weight_tconv4", w, wt, co, ci, coef)
    gx = torch.empty(n, h, w_, ci, dtype=torch.bfloat16, device=DEV)
    db = torch.full((ci,), 9.0, device=DEV) if gated else None
    bgn.call("bg_conv_pool4_dgrad", gpool, wt, gx, n, h // 2, w_ // 2, co, ci, gate, 0.2, db)
    torch.cuda.synchronize()
    err = relerr(nchw(gx), ref)
    assert err < 8e-3, f"conv+pool dgrad as transposed 4x4s2 {shape} gated={gated}: rel-L2 {err:.3e}"
    if gated:
        want = gx.double().sum(dim=(0, 1, 2))
        assert (db.double() - want).abs().max().item() < 2e-5 * (want.abs().max().item() + 1.0) * math.sqrt(h * w_)


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64), (1, 64, 64, 32, 64), (3, 32, 32, 128, 128), (1, 64, 64, 256, 256),
                                   (2, 32, 32, 512, 512), (1, 128, 128, 64, 64), (2, 32, 32, 64, 32)])
def test_conv_pool_wgrad_on_4x4_stride2_form(shape):
    """autograd's weight gradient of conv3x3 -> AvgPool2d(2) (gan.py:258-260) from the POOLED output gradient
    (bg_conv_pool4_wgrad + bg_unpack_wgrad_pool4), incl. the accumulate=1 form of the R1 doubled-K contraction."""
    n, h, w_, ci, co = shape
    torch.manual_seed(0)
    y1 = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    gpool = nhwc(torch.randn(n, co, h // 2, w_ // 2, device=DEV))
    coef = math.sqrt(2 / (ci * 9))
    wparam = torch.randn(co, ci, 3, 3, device=DEV, requires_grad=True)
    F.avg_pool2d(F.conv2d(nchw(y1), wparam * coef, None, padding=1), 2).backward(nchw(gpool))
    dw16 = torch.empty(16, co, ci, device=DEV)
    bgn.call("bg_conv_pool4_wgrad", y1, gpool, dw16, n, h // 2, w_ // 2, ci, co, 0)
    bgn.call("bg_conv_pool4_wgrad", y1, gpool, dw16, n, h // 2, w_ // 2, ci, co, 1)
    dw = torch.empty(co, ci, 3, 3, device=DEV)
    bgn.call("bg_unpack_wgrad_pool4", dw16, dw, co, ci, 0.5 * coef, 0)
    torch.cuda.synchronize()
    err = relerr(dw, wparam.grad)
    assert err < 6e-3, f"conv+pool wgrad {shape}: rel-L2 {err:.3e}"


@pytest.mark.parametrize("shape", [(32, 8, 8, 512, 512), (32, 4, 4, 512, 576), (32, 4, 4, 576, 512), (16, 4, 4, 512, 512),
                                   (6, 8, 8, 512, 512), (3, 4, 4, 512, 512), (5, 8, 8, 128, 256), (4, 8, 8, 256, 128)])
@pytest.mark.parametrize("variant", ["plain", "fused", "gated"])
def test_small_map_conv_split_k_cluster(shape, variant):
    """conv_splitk.cu: the K loop of a small-map layer split over the CTAs of a cluster, partials reduced through
    distributed shared memory.  Against F.conv2d (the reference op, gan.py:29-38), against the tap-wise kernel, and
    bit-reproducible from run to run (ordered reduction, no atomics: these are forward activations)."""
    n, h, w_, ci, co = shape
    torch.manual_seed(3)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    w = torch.randn(co, ci, 3, 3, device=DEV)
    coef = math.sqrt(2 / (ci * 9))
    wf, _ = pack(w, coef)
    fused, gated = variant == "fused", variant == "gated"
    bias = torch.randn(co, device=DEV) * 0.1 if fused else None
    noise = torch.randn(n, 1, h, w_, device=DEV) if fused else None
    nw = torch.randn(co, device=DEV) * 0.1 if fused else None
    gate = nhwc(torch.randn(n, co, h, w_, device=DEV)) if gated else None
    outs = []
    for entry in ("bg_conv_fprop", "bg_conv_fprop", "bg_conv_fprop_tapwise"):
        out = torch.empty(n, h, w_, co, dtype=torch.bfloat16, device=DEV)
        bgn.call(entry, x, wf, out, n, h, w_, ci, co, 3, bias, noise, nw, gate, 1 if fused else 0, 0.2)
        outs.append(out)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x), (w * coef).to(torch.bfloat16).float(), None, padding=1)
    if fused:
        ref = F.leaky_relu(ref + bias.view(1, -1, 1, 1) + nw.view(1, -1, 1, 1) * noise, 0.2)
    if gated:
        ref = ref * torch.where(nchw(gate) > 0, 1.0, 0.2)
    assert relerr(nchw(outs[0]), ref) < 6e-3, relerr(nchw(outs[0]), ref)
    assert torch.equal(outs[0], outs[1]), "split-K result differs between two runs"
    assert relerr(outs[0].float(), outs[2].float()) < 4e-3        # other summation order than the tap-wise kernel


def test_pack_weight_grouped_equals_single_packs():
    """bg_pack_weight_grouped (all stale packs of a network in one launch, tiled kernel) == bg_pack_weight layer by layer,
    including the padded 513 -> 576 input channels of the critic's minibatch-stddev layer and ragged tile edges."""
    torch.manual_seed(2)
    layers = [(512, 513, 576), (32, 16, 16), (16, 32, 32), (256, 512, 512), (48, 64, 64), (64, 96, 96), (128, 128, 128)]
    ws, wfs, wds, meta = [], [], [], []
    for co, ci, cp in layers:
        w = torch.randn(co, ci, 3, 3, device=DEV)
        ws.append(w)
        wfs.append(torch.full((9, co, cp), 7.0, dtype=torch.bfloat16, device=DEV))
        wds.append(torch.full((9, cp, co), 7.0, dtype=torch.bfloat16, device=DEV))
        meta.append((co, ci, cp, 3, math.sqrt(2 / (ci * 9))))
    bgn.call("bg_pack_weight_grouped", ws, wfs, wds, [m[0] for m in meta], [m[1] for m in meta], [m[2] for m in meta],
             [m[3] for m in meta], [m[4] for m in meta], len(layers))
    for w, wf, wd, (co, ci, cp, _, coef) in zip(ws, wfs, wds, meta):
        rf = torch.empty(9, co, cp, dtype=torch.bfloat16, device=DEV)
        rd = torch.empty(9, cp, co, dtype=torch.bfloat16, device=DEV)
        bgn.call("bg_pack_weight", w, rf, rd, co, ci, cp, 3, coef)
        assert torch.equal(wf, rf) and torch.equal(wd, rd), (co, ci, cp)


@pytest.mark.parametrize("shape", [(32, 8, 8, 512, 512), (32, 4, 4, 512, 512), (16, 8, 8, 256, 128), (6, 4, 4, 128, 384)])
def test_small_map_wgrad_single_tap_output_stationary(shape):
    """conv_wgrad.cu single-tap mode (one CTA per (tap, 128 co, 128 ci), whole pixel range, no split-K): overwrite and
    accumulate forms against autograd's conv2d_weight and the split-K decomposition; bit-reproducible (no atomics)."""
    n, h, w_, ci, co = shape
    torch.manual_seed(4)
    x = nhwc(torch.randn(n, ci, h, w_, device=DEV))
    g = nhwc(torch.randn(n, co, h, w_, device=DEV))
    ref = torch.nn.grad.conv2d_weight(nchw(x), (co, ci, 3, 3), nchw(g), padding=1)
    outs = []
    for _ in range(2):
        dwp = torch.full((9, co, ci), 3.0, dtype=torch.float32, device=DEV)          # must be overwritten, not added to
        bgn.call("bg_conv_wgrad", x, g, dwp, n, h, w_, ci, co, 0)
        outs.append(dwp.clone())
        bgn.call("bg_conv_wgrad", x, g, dwp, n, h, w_, ci, co, 1)
        assert relerr(dwp, 2 * outs[-1]) < 1e-6
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    got = outs[0].reshape(3, 3, co, ci).permute(2, 3, 0, 1)
    assert relerr(got, ref) < 2e-3, relerr(got, ref)
