"""Kernel-level parity (GPU) for the fp32 helper kernels: EqualizedLinear fwd/bwd (gan.py:16-17), learned
constant (gan.py:92), image-plane fade ops (gan.py:213-220, 345), MiniBatchStdDev forward / tangent / backward /
second-order (gan.py:273-298) against autograd over the oracle restatement, and the loss terms (gan.py:228,396-406)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import bg_native as bgn  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

DEV = "cuda"


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize("m,n,k", [(16, 512, 512), (5, 1024, 512), (32, 512, 8192), (7, 1, 512), (9, 512, 1), (33, 32, 512),
                                   (32, 8192, 512), (5, 4098, 512), (3, 6, 2048), (11, 13, 4100), (3, 70, 2048), (11, 66, 4100), (32, 512, 516)])
def test_linear_fwd_and_bwd(m, n, k):
    torch.manual_seed(0)
    x = torch.randn(m, k, device=DEV)
    w = torch.randn(n, k, device=DEV)
    b = torch.randn(n, device=DEV)
    coef = math.sqrt(2 / k)
    y = torch.empty(m, n, device=DEV)
    bgn.call("bg_linear_fwd", x, w, b, y, m, n, k, coef, 1, 0.2)
    ref = F.leaky_relu(F.linear(x, w * coef, b), 0.2)
    assert rel(y, ref) < 1e-5
    gy = torch.randn(m, n, device=DEV)
    wt = torch.empty(k, n, device=DEV)
    bgn.call("bg_transpose_f32", w, wt, n, k)
    assert torch.equal(wt, w.t().contiguous())
    gx = torch.empty(m, k, device=DEV)
    bgn.call("bg_linear_fwd", gy, wt, None, gx, m, k, n, coef, 0, 0.2)
    assert rel(gx, gy @ (w * coef)) < 1e-5
    if k % 2 == 0:
        gx2 = torch.full((m, k), float("nan"), device=DEV)
        bgn.call("bg_linear_bwd_input", gy, w, gx2, m, n, k, coef)
        assert rel(gx2, gy @ (w * coef)) < 1e-5
    if k % 4 == 0:
        dw = torch.empty(n, k, device=DEV)
        db = torch.empty(n, device=DEV)
        bgn.call("bg_linear_bwd_weight", gy, x, dw, db, m, n, k, coef, 0)
        assert rel(dw, coef * gy.t() @ x) < 1e-5 and rel(db, gy.sum(0)) < 1e-5
        bgn.call("bg_linear_bwd_weight", gy, x, dw, db, m, n, k, coef, 1)
        assert rel(dw, 2 * coef * gy.t() @ x) < 1e-5 and rel(db, 2 * gy.sum(0)) < 1e-5


def test_const_noise_act_and_bwd():
    torch.manual_seed(0)
    n, c = 5, 512
    cst = torch.randn(1, c, 4, 4, device=DEV)
    noise = torch.randn(n, 1, 4, 4, device=DEV)
    nw = torch.randn(c, device=DEV) * 0.1
    a = torch.empty(n, 4, 4, c, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_const_noise_act", cst, noise, nw, a, n, 16, c, 0.2)
    ref = F.leaky_relu(cst.repeat(n, 1, 1, 1) + nw.view(1, c, 1, 1) * noise, 0.2)
    assert rel(a.float().permute(0, 3, 1, 2), ref) < 4e-3
    g = torch.randn(n, 4, 4, c, device=DEV).to(torch.bfloat16)
    dc = torch.empty(1, c, 4, 4, device=DEV)
    bgn.call("bg_const_bwd", g, dc, n, 16, c)
    assert rel(dc, g.float().permute(0, 3, 1, 2).sum(0, keepdim=True)) < 1e-5


@pytest.mark.parametrize("r", [2, 4, 16])
def test_image_plane_fade_ops(r):
    torch.manual_seed(0)
    b, alpha = 3, 0.3
    small = torch.randn(b, 3, r, r, device=DEV, requires_grad=True)
    large = torch.randn(b, 3, 2 * r, 2 * r, device=DEV, requires_grad=True)
    out = torch.empty(b, 3, 2 * r, 2 * r, device=DEV)
    bgn.call("bg_img_up2_lerp", small.detach(), large.detach(), out, b * 3, r, r, alpha)
    ref = torch.lerp(F.interpolate(small, scale_factor=2, mode="bilinear"), large, alpha)
    assert rel(out, ref) < 1e-6
    g = torch.randn_like(ref)
    ref.backward(g)
    gs = torch.empty(b, 3, r, r, device=DEV)
    bgn.call("bg_img_up2_bwd", g, gs, b * 3, r, r, 1 - alpha)
    assert rel(gs, small.grad) < 1e-5
    img = torch.randn(b, 3, 2 * r, 2 * r, device=DEV, requires_grad=True)
    p = torch.empty(b, 3, r, r, device=DEV)
    bgn.call("bg_img_avgpool2", img.detach(), p, b * 3, r, r)
    refp = F.avg_pool2d(img, 2)
    assert rel(p, refp) < 1e-6
    gp = torch.randn_like(refp)
    refp.backward(gp)
    gi = torch.ones(b, 3, 2 * r, 2 * r, device=DEV)
    bgn.call("bg_img_avgpool2_bwd", gp, gi, b * 3, r, r, 1.0, 1)
    assert rel(gi - 1, img.grad) < 1e-5
    sums = torch.empty(3, device=DEV)
    bgn.call("bg_plane_sums", g, sums, b, 4 * r * r)
    assert rel(sums, g.sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("b,gsz", [(4, 4), (8, 4), (16, 4), (32, 4), (6, 6)])
def test_minibatch_stddev_all_orders(b, gsz):
    """forward value, tangent (JVP), VJP and the second-order term  d/dx [ ghat_s . J(x) v ]  vs autograd."""
    torch.manual_seed(b)
    C, HW, CP = 512, 16, 576
    xb = (torch.randn(b, HW, C, device=DEV) * 1.5).to(torch.bfloat16)
    vb = torch.randn(b, HW, C, device=DEV).to(torch.bfloat16)
    M = b // gsz

    def s_of(x_nhwc):                       # oracle on NCHW fp64
        x = x_nhwc.permute(0, 2, 1).reshape(b, C, 4, 4)
        y, _ = O.minibatch_stddev(x, gsz)
        return y[:M, C, 0, 0]

    x64 = xb.double().requires_grad_()
    s_ref = s_of(x64)
    plane = torch.empty(M, device=DEV)
    xpad = torch.empty(b, HW, CP, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_mbstd_fwd", xb, None, plane, xpad, b, gsz, HW, C, CP, 1e-8)
    assert rel(plane, s_ref) < 1e-5
    assert torch.equal(xpad[..., :C], xb) and xpad[..., C + 1:].abs().sum() == 0
    assert rel(xpad[:, 0, C].float(), s_ref.repeat(gsz).float()) < 4e-3
    # tangent
    v64 = vb.double()
    (sdot_ref,) = torch.autograd.functional.jvp(s_of, (xb.double(),), (v64,))[1:]
    sdot = torch.empty(M, device=DEV)
    vpad = torch.empty(b, HW, CP, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_mbstd_fwd", xb, vb, sdot, vpad, b, gsz, HW, C, CP, 1e-8)
    assert rel(sdot, sdot_ref) < 1e-4
    assert torch.equal(vpad[..., :C], vb)
    # VJP:  gpad carries the pass-through gradient (channels < C) and the plane gradient (channel C)
    gpad = torch.randn(b, HW, CP, device=DEV).to(torch.bfloat16)
    gs = gpad[..., C].double().sum(1).reshape(gsz, M).sum(0)              # per-slot plane gradient
    (vjp_ref,) = torch.autograd.grad(s_ref, x64, gs, create_graph=False, retain_graph=True)
    ws = torch.empty(2 * M, device=DEV)
    gx = torch.empty(b, HW, C, dtype=torch.bfloat16, device=DEV)
    bgn.call("bg_mbstd_bwd", xb, None, gpad, None, ws, gx, b, gsz, HW, C, CP, 1e-8)
    ref_total = gpad[..., :C].double() + vjp_ref
    assert rel(gx.double(), ref_total) < 5e-3
    zero_pad = torch.zeros_like(gpad)                                     # isolate the (small) stddev term
    zero_pad[..., C] = gpad[..., C]
    bgn.call("bg_mbstd_bwd", xb, None, zero_pad, None, ws, gx, b, gsz, HW, C, CP, 1e-8)
    assert rel(gx.double(), vjp_ref) < 5e-3
    # second order: q = d/dx [ sum_m gs2[m] * sdot_m(x; v) ]
    gpad2 = torch.randn(b, HW, CP, device=DEV).to(torch.bfloat16)
    gs2 = gpad2[..., C].double().sum(1).reshape(gsz, M).sum(0)
    x2 = xb.double().requires_grad_()
    sd = torch.autograd.functional.jvp(s_of, (x2,), (v64,), create_graph=True)[1]
    (q_ref,) = torch.autograd.grad((sd * gs2).sum(), x2)
    bgn.call("bg_mbstd_bwd", xb, vb, zero_pad, gpad2, ws, gx, b, gsz, HW, C, CP, 1e-8)
    assert rel(gx.double(), vjp_ref + q_ref) < 2e-2


def test_loss_terms():
    torch.manual_seed(0)
    p = torch.randn(37, 1, device=DEV) * 3
    p[0] = 25.0
    loss = torch.empty(1, device=DEV)
    seed = torch.empty(37, 1, device=DEV)
    for sign in (1.0, -1.0):
        pr = p.clone().requires_grad_()
        ref = F.softplus(sign * pr).mean()
        ref.backward()
        bgn.call("bg_logistic_loss", p, 37, sign, loss, seed, 1.0)
        assert abs(loss.item() - ref.item()) < 1e-6 * abs(ref.item()) + 1e-7
        assert rel(seed, pr.grad) < 1e-5
    x = torch.randn(100003, device=DEV)
    out = torch.empty(1, device=DEV)
    bgn.call("bg_sumsq", x, x.numel(), 0.5, out)
    assert abs(out.item() - 0.5 * (x.double() ** 2).sum().item()) < 1e-3 * out.item()


def test_flatten_boundary_roundtrip():
    torch.manual_seed(0)
    n = 6
    x = torch.randn(n, 4, 4, 512, device=DEV).to(torch.bfloat16)
    f = torch.empty(n, 8192, device=DEV)
    bgn.call("bg_nhwc_to_nchw_f32", x, f, n, 16, 512)
    assert torch.equal(f, x.float().permute(0, 3, 1, 2).reshape(n, -1))
    back = torch.empty_like(x)
    bgn.call("bg_nchw_f32_to_nhwc", f, x, back, n, 16, 512, 0.2)
    assert torch.equal(back, (x.float() * torch.where(x.float() > 0, 1.0, 0.2)).to(torch.bfloat16))


def test_grouped_style_linears():
    """The generator's AdaIN style FCs (gan.py:60,66) as grouped launches: forward, weight/bias gradient and the summed
    input gradient against per-layer torch fp32."""
    torch.manual_seed(4)
    M, K = 12, 512
    Ns = [1024, 512, 64, 32, 256]
    x = torch.randn(M, K, device=DEV)
    Ws = [torch.randn(n, K, device=DEV) for n in Ns]
    bs = [torch.randn(n, device=DEV) for n in Ns]
    coefs = [math.sqrt(2 / K) * (1 + 0.1 * i) for i in range(len(Ns))]
    ys = [torch.empty(M, n, device=DEV) for n in Ns]
    bgn.call("bg_linear_fwd_grouped", x, Ws, bs, ys, Ns, coefs, len(Ns), M, K, 0, 0.2)
    for W, b, y, c in zip(Ws, bs, ys, coefs):
        assert torch.allclose(y, x @ (W * c).t() + b, rtol=1e-4, atol=1e-4)
    gys = [torch.randn(M, n, device=DEV) for n in Ns]
    dWs = [torch.empty(n, K, device=DEV) for n in Ns]
    dbs = [torch.empty(n, device=DEV) for n in Ns]
    bgn.call("bg_linear_bwd_weight_grouped", [x] * len(Ns), gys, dWs, dbs, Ns, coefs, len(Ns), M, K)
    for gy, dW, db, c in zip(gys, dWs, dbs, coefs):
        assert torch.allclose(dW, c * gy.t() @ x, rtol=1e-4, atol=1e-4)
        assert torch.allclose(db, gy.sum(0), rtol=1e-4, atol=1e-4)
    # one input per layer (the mapping network's eight 512 x 512 layers in one launch)
    xs = [torch.randn(M, K, device=DEV) for _ in Ns]
    bgn.call("bg_linear_bwd_weight_grouped", xs, gys, dWs, dbs, Ns, coefs, len(Ns), M, K)
    for xg, gy, dW, db, c in zip(xs, gys, dWs, dbs, coefs):
        assert torch.allclose(dW, c * gy.t() @ xg, rtol=1e-4, atol=1e-4)
        assert torch.allclose(db, gy.sum(0), rtol=1e-4, atol=1e-4)
    want = sum(c * gy @ W for gy, W, c in zip(gys, Ws, coefs))
    outs = []
    for _ in range(3):                                  # the cluster / DSMEM reduction is ordered: bit-identical reruns
        gx = torch.full((M, K), float("nan"), device=DEV)
        bgn.call("bg_linear_bwd_input_grouped", gys, Ws, Ns, coefs, len(Ns), M, K, gx)
        outs.append(gx)
    assert torch.allclose(outs[0], want, rtol=1e-4, atol=2e-3)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("M,K,Ns", [(64, 512, [1024] * 4 + [512, 256, 128, 64] * 2 + [32, 32]), (32, 512, [40, 4, 1000]),
                                    (3, 68, [4, 8]), (70, 132, [516])])
def test_grouped_linear_input_gradient_shapes(M, K, Ns):
    """bg_linear_bwd_input_grouped on the training shape (14 style FCs of the 256x256 stage, both latents) and ragged ones:
    group boundaries inside a CTA's share of the n range, n shares shorter than a warp span, K not a multiple of 64."""
    torch.manual_seed(len(Ns) + M)
    Ws = [torch.randn(n, K, device=DEV) for n in Ns]
    gys = [torch.randn(M, n, device=DEV) for n in Ns]
    coefs = [0.05 * (1 + 0.1 * i) for i in range(len(Ns))]
    gx = torch.full((M, K), float("nan"), device=DEV)
    bgn.call("bg_linear_bwd_input_grouped", gys, Ws, Ns, coefs, len(Ns), M, K, gx)
    want = sum(c * gy.double() @ W.double() for gy, W, c in zip(gys, Ws, coefs))
    assert rel(gx, want) < 1e-5


def test_helper_truncated_noise_distribution():
    """Drop-in helper.get_truncated_noise (helper.py:36-45): same distribution as scipy truncnorm.rvs(-t, t)."""
    import helper
    from scipy.stats import truncnorm

    torch.manual_seed(0)
    for t in (0.75, 2.0):
        x = helper.get_truncated_noise(4096, 512, t)
        assert x.is_cuda and x.dtype == torch.float32 and x.requires_grad and x.shape == (4096, 512)
        v = x.detach()
        assert v.abs().max().item() <= t + 1e-6
        assert abs(v.mean().item()) < 3e-3
        assert abs(v.std().item() - truncnorm.std(-t, t)) < 3e-3
        # quartiles against the analytic inverse CDF
        for q in (0.1, 0.25, 0.5, 0.9):
            assert abs(torch.quantile(v.flatten()[:1000000], q).item() - truncnorm.ppf(q, -t, t)) < 5e-3
