"""Model-level parity on the GPU: byo-gan_b200/gan.py (CUDA kernels through the C ABI) against
 (a) the oracle restatement evaluated in fp32 on the same device with TF32 off, on the same seeded inputs, and
 (b) the golden fingerprints recorded from the unmodified reference (tests/golden/*.json).
Tolerances are the stated bf16-vs-fp32 ones in tests/parity_util.py."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import gan_oracle as O  # noqa: E402
import parity_util as U  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def fp_close(t, fp, tol, what):
    got = O.fingerprint(t)
    assert got["shape"] == fp["shape"], what
    a, b = torch.tensor(got["samples"], dtype=torch.float64), torch.tensor(fp["samples"], dtype=torch.float64)
    err = ((a - b).norm() / (b.norm() + 1e-30)).item()
    assert err < tol, f"{what}: sampled rel-L2 {err:.3e} vs reference fingerprint"
    assert abs(got["norm"] - fp["norm"]) < tol * fp["norm"] + 1e-6, f"{what}: norm {got['norm']} vs {fp['norm']}"


@pytest.mark.parametrize("case", gold("forward.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_forward_vs_reference_golden_and_oracle(case):
    U.no_tf32()
    steps, batch, alpha = case["steps"], case["batch"], case["alpha"]
    g, c = U.build_models(1)
    z = O.make_latents(batch, steps).cuda()
    noise = [n.cuda() for n in O.make_noise(batch, steps, steps)]
    real = O.make_images(batch, steps, steps).cuda()
    with torch.no_grad():
        fake = g(z, noise=noise, steps=steps, alpha=alpha)
        pr = c(real, steps, alpha)
        Gs = {k: v.cuda() for k, v in O.make_state("gen", 1).items()}
        Ds = {k: v.cuda() for k, v in O.make_state("critic", 1).items()}
        fake_o = O.generator_forward(Gs, z, noise, steps, alpha)
        pr_o = O.critic_forward(Ds, real, steps, alpha)
        pf = c(fake_o, steps, alpha)
        pf_o = O.critic_forward(Ds, fake_o, steps, alpha)
    assert fake.shape == fake_o.shape and fake.dtype == torch.float32
    assert U.rel(fake, fake_o) < U.TOL_IMG, f"image rel-L2 {U.rel(fake, fake_o):.3e}"
    fp_close(fake, case["fake"], U.TOL_IMG, "image vs golden")
    for got, ref, name in ((pr, pr_o, "D(real)"), (pf, pf_o, "D(fake)")):
        scale = ref.pow(2).mean().sqrt().item() + 1.0
        assert (got - ref).abs().max().item() < U.TOL_PRED * scale, f"{name}: {got.flatten()} vs {ref.flatten()}"
    ref_pr = torch.tensor(case["pred_real"]["samples"])[: pr.numel()]
    assert (pr.flatten().cpu()[: ref_pr.numel()] - ref_pr).abs().max() < U.TOL_PRED * (ref_pr.pow(2).mean().sqrt() + 1)


def _iteration_inputs(steps, batch):
    return (O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps), O.make_images(batch, steps, 30 + steps),
            O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps))


def _oracle_pair(args, steps, alpha, lam, **kw):
    """The oracle in fp32 (the reference) and in its bf16-storage emulation (the per-tensor noise yardstick)."""
    o = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, lam, device="cuda", **kw)
    O.QUANT[0] = True
    try:
        q = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *args, steps, alpha, lam, device="cuda", **kw)
    finally:
        O.QUANT[0] = False
    return o, q


def _check_gradients(r, o, q, kinds=("d_grads", "g_grads"), case=None, zero_for_none=False):
    """Per-tensor gate (parity_util.grad_gate) on every parameter gradient + None-ness + the network-median criterion.
    zero_for_none: a parameter the reference's autograd never reaches (.grad None) may carry an all-zero gradient here
    (the penalty-only critic step: the last bias cannot influence d D / d x, the hand-scheduled pass still visits it)."""
    bad = []
    for kind in kinds:
        errs, errs_emu = [], []
        for k, ref in o[kind].items():
            got = r[kind][k]
            if zero_for_none and ref is None and got is not None:
                assert got.abs().max().item() == 0.0, f"{kind}[{k}]: the reference has no gradient, got a non-zero one"
                continue
            assert (got is None) == (ref is None), f"{kind}[{k}]: None-ness differs from the reference"
            if case is not None:
                assert (case[kind][k] is None) == (ref is None), f"{kind}[{k}]: oracle None-ness differs from the golden"
            if ref is None:
                continue
            assert got.shape == ref.shape
            why = U.grad_gate(f"{kind}[{k}]", got, ref, q[kind][k])
            if why:
                bad.append(why)
            if ref.norm().item() > 0:
                errs.append(U.rel(got, ref))
                errs_emu.append(U.rel(q[kind][k], ref))
        med, med_emu = sorted(errs)[len(errs) // 2], sorted(errs_emu)[len(errs_emu) // 2]
        assert med <= U.TOL_VS_EMU * med_emu + 0.01, \
            f"{kind}: median rel-L2 {med:.4f} vs bf16-storage emulation of the reference {med_emu:.4f}"
    assert not bad, f"{len(bad)} gradient tensors outside the per-tensor gate:\n  " + "\n  ".join(bad[:12])


@pytest.mark.parametrize("case", gold("train_iteration.json"),
                         ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_train_iteration_vs_reference(case):
    """One G+D iteration (train.py:135-217) incl. the 512x512 stage (steps=8, BASELINE configs[3] shape): losses, every
    parameter gradient through the per-tensor gate, which gradients are None."""
    U.no_tf32()
    steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
    g, c = U.build_models(2)
    args = _iteration_inputs(steps, batch)
    with U.deterministic():
        r = U.cuda_iteration(g, c, *args, steps, alpha, lam)
    o, q = _oracle_pair(args, steps, alpha, lam)
    # losses vs the reference's recorded values and vs the oracle
    assert abs(r["c_loss"].item() - case["c_loss"]) < U.TOL_LOSS * abs(case["c_loss"]), (r["c_loss"].item(), case["c_loss"])
    assert abs(r["g_loss"].item() - case["g_loss"]) < U.TOL_LOSS * abs(case["g_loss"]), (r["g_loss"].item(), case["g_loss"])
    assert abs(r["c_loss"].item() - o["c_loss"].item()) < U.TOL_LOSS * abs(o["c_loss"].item())
    assert U.rel(r["fake_d"], o["fake_d"]) < U.TOL_IMG
    fp_close(r["fake_d"], case["fake_d"], U.TOL_IMG, "fake (D step) vs golden")
    _check_gradients(r, o, q, case=case)
    assert U.cos(r["z_grad"], o["z_grad"]) > 0.95


@pytest.mark.parametrize("case", gold("r1_penalty.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_r1_penalty_alone_second_order_terms(case):
    """Critic step with ONLY lambda/2 * mean ||d sum D(real) / d real||^2 (gan.py:398-404): its parameter gradient is
    purely second order (tangent pass x gated ones-backprop, minibatch-stddev curvature), so a wrong or missing
    double-backward term in ANY layer fails here at O(1) instead of hiding under the first-order gradient."""
    U.no_tf32()
    steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
    g, c = U.build_models(2)
    args = _iteration_inputs(steps, batch)
    with U.deterministic():
        r = U.cuda_iteration(g, c, *args, steps, alpha, lam, loss="r1_penalty")
    o, q = _oracle_pair(args, steps, alpha, lam, loss="r1_penalty")
    assert abs(r["c_loss"].item() - case["penalty"]) < U.TOL_LOSS * abs(case["penalty"]), (r["c_loss"].item(), case["penalty"])
    assert abs(o["c_loss"].item() - case["penalty"]) < 1e-3 * abs(case["penalty"])
    # the image gradient itself (what autograd.grad returns at gan.py:398-400): against the oracle, its emulation, the golden
    real = args[2].cuda()
    grads = []
    for quant in (False, True):
        O.QUANT[0] = quant
        try:
            x = real.clone().requires_grad_()
            Ds = {k: v.cuda() for k, v in O.make_state("critic", 2).items()}
            grads.append(torch.autograd.grad(O.critic_forward(Ds, x, steps, alpha).sum(), x)[0])
        finally:
            O.QUANT[0] = False
    why = U.grad_gate("d sum D(real) / d real", r["real_grad"], grads[0], grads[1])
    assert why is None, why
    fp_close(grads[0], case["grad_real"], 5e-3, "oracle image gradient (GPU fp32) vs golden (CPU fp32)")
    _check_gradients(r, o, q, kinds=("d_grads",), case=case, zero_for_none=True)


@pytest.mark.parametrize("case", gold("wgan_gp.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_wgan_gp_iteration_vs_repaired_reference(case):
    """WGAN-GP (train.py:177-185,213 -> gan.py:357-391, 224-225), SURVEY §8f-2.  The golden values come from the
    reference's own modules executing the body of get_wgan_loss with its two defects repaired (oracle/make_golden.py)."""
    U.no_tf32()
    steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
    g, c = U.build_models(2)
    args = _iteration_inputs(steps, batch)
    eps = O.make_epsilon(batch, 40 + steps)
    with U.deterministic():
        r = U.cuda_iteration(g, c, *args, steps, alpha, lam, loss="wgan", epsilon=eps)
    o, q = _oracle_pair(args, steps, alpha, lam, loss="wgan", epsilon=eps)
    for key in ("c_loss", "g_loss"):
        assert abs(r[key].item() - case[key]) < U.TOL_LOSS * (abs(case[key]) + 0.05), (key, r[key].item(), case[key])
        assert abs(o[key].item() - case[key]) < 1e-3 * (abs(case[key]) + 0.05)
    _check_gradients(r, o, q, case=case)


@pytest.mark.parametrize("steps,batch,alpha", [(5, 8, 0.5), (6, 4, None)])
def test_deterministic_mode_is_bit_reproducible(steps, batch, alpha):
    """bg_set_deterministic: two runs of the same iteration give bit-identical images, scores, image gradients and latent
    gradients (every reduction that feeds later layers is ordered); parameter gradients — leaf sums that keep their fp32
    atomics, some heavily cancelling — agree to 1e-3.  Without the switch the same comparison needs 0.3-0.6 (see the emission test below)."""
    U.no_tf32()
    args = _iteration_inputs(steps, batch)
    runs = []
    with U.deterministic():
        for _ in range(2):
            g, c = U.build_models(2)
            runs.append(U.cuda_iteration(g, c, *args, steps, alpha, 10.0))
    a, b = runs
    for key in ("fake_d", "pred_fake", "pred_real", "real_grad", "pred_g", "z_grad", "c_loss", "g_loss"):
        assert torch.equal(a[key], b[key]), f"{key} differs between two deterministic runs"
    for kind in ("d_grads", "g_grads"):
        for k, ga in a[kind].items():
            if ga is not None:
                assert U.rel(b[kind][k], ga) < 1e-3, (kind, k, U.rel(b[kind][k], ga))


@pytest.mark.parametrize("steps,batch", [(7, 32), (8, 16)], ids=["256x256-b32", "512x512-b16"])
def test_full_batch_gradients_vs_oracle(steps, batch):
    """BASELINE.json configs[2] / configs[3] at their FULL per-GPU batch: one whole G+D iteration (R1 double-backward
    included) against the oracle evaluated in fp32 on the same device, every parameter gradient through the per-tensor
    gate.  (The golden fixtures stop at batch 4 because the reference runs on the container's CPU.)"""
    U.no_tf32()
    g, c = U.build_models(2)
    args = _iteration_inputs(steps, batch)
    with U.deterministic():
        r = U.cuda_iteration(g, c, *args, steps, None, 10.0)
    del g, c
    torch.cuda.empty_cache()
    o, q = _oracle_pair(args, steps, None, 10.0)
    assert abs(r["c_loss"].item() - o["c_loss"].item()) < U.TOL_LOSS * abs(o["c_loss"].item())
    assert abs(r["g_loss"].item() - o["g_loss"].item()) < U.TOL_LOSS * abs(o["g_loss"].item())
    assert U.rel(r["fake_d"], o["fake_d"]) < U.TOL_IMG
    _check_gradients(r, o, q)


def test_sampling_512_batch_256_vs_oracle():
    """BASELINE.json configs[4]: generate_samples at 512x512, batch 256, explicit noise — the whole
    batch in ONE call of the CUDA path, checked against the oracle evaluated slice by slice (the fp32 reference needs
    ~1.2 GB of activations per sample at this size)."""
    U.no_tf32()
    steps, batch = 8, 256
    g, _ = U.build_models(3)
    z = O.make_latents(batch, 90).cuda()
    noise = [n.cuda() for n in O.make_noise(batch, steps, 91)]
    Gs = {k: v.cuda() for k, v in O.make_state("gen", 3).items()}
    with torch.no_grad():
        with U.deterministic():
            img = g(z, noise=noise, steps=steps, alpha=None)
        assert img.shape == (batch, 3, 512, 512) and img.dtype == torch.float32
        worst = 0.0
        for i in range(0, batch, 8):
            ref = O.generator_forward(Gs, z[i:i + 8], [n[i:i + 8] for n in noise], steps, None)
            worst = max(worst, U.rel(img[i:i + 8], ref))
            per = (img[i:i + 8] - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)
            assert per.max().item() < 1.5 * U.TOL_IMG, (i, per.tolist())          # no single sample off either
        assert worst < U.TOL_IMG, worst


@pytest.mark.parametrize("steps,batch", [(7, 32), (8, 16)], ids=["256x256-b32", "512x512-b16"])
def test_full_size_forward_and_batch_independence(steps, batch):
    """BASELINE.json configs[2] / configs[3] at their FULL per-GPU batch: (a) images and critic scores against the
    fp32 oracle on the same inputs; (b) size-independent property — every sample of the generator's output is a
    function of its own latent and noise only (gan.py:183-222 has no cross-sample op), so the batch-32 result must equal
    the results of its 4-sample slices (this exercises the per-sample weight packs, bias tables and statistics of the
    fused style convolutions at the sizes the benchmark runs); (c) the critic's minibatch-stddev DOES couple samples
    (gan.py:273-298): permuting whole stddev groups must permute the scores."""
    U.no_tf32()
    g, c = U.build_models(3)
    z = O.make_latents(batch, 40 + steps).cuda()
    noise = [n.cuda() for n in O.make_noise(batch, steps, 40 + steps)]
    with torch.no_grad():
        fake = g(z, noise=noise, steps=steps, alpha=None)
        Gs = {k: v.cuda() for k, v in O.make_state("gen", 3).items()}
        Ds = {k: v.cuda() for k, v in O.make_state("critic", 3).items()}
        fake_o = torch.cat([O.generator_forward(Gs, z[i:i + 4], [n[i:i + 4] for n in noise], steps, None)
                            for i in range(0, batch, 4)])
        assert U.rel(fake, fake_o) < U.TOL_IMG, f"image rel-L2 {U.rel(fake, fake_o):.3e}"
        for i in range(0, batch, 4):
            part = g(z[i:i + 4].contiguous(), noise=[n[i:i + 4].contiguous() for n in noise], steps=steps, alpha=None)
            # not bit-equal: the fp32 atomics of the fused statistics sum in a different order, and one flipped bf16
            # rounding early in the 14-layer chain decorrelates the rest at the 2^-8 level; a wrong sample's weights or
            # statistics would show up as an O(1) error
            assert U.rel(part, fake[i:i + 4]) < 1.5e-2, (i, U.rel(part, fake[i:i + 4]))
        pf = c(fake_o, steps, None)
        pf_o = O.critic_forward(Ds, fake_o, steps, None)
        scale = pf_o.pow(2).mean().sqrt().item() + 1.0
        assert (pf - pf_o).abs().max().item() < U.TOL_PRED * scale
        # stddev groups are strided: sample n belongs to slot n mod (B/4); rotating the batch by B/4 keeps every slot's
        # member set, so scores rotate with the samples
        m = batch // 4
        rolled = torch.roll(fake_o, shifts=m, dims=0).contiguous()
        pf_r = c(rolled, steps, None)
        assert (torch.roll(pf, shifts=m, dims=0) - pf_r).abs().max().item() < 2e-2 * scale


@pytest.mark.parametrize("steps,crossover", [(4, 2), (6, 3), (6, 0), (5, 5)])
def test_style_mixing_extension(steps, crossover):
    """Opt-in extension (not in the reference): blocks >= crossover are styled by a second latent.  Oracle =
    the reference's own sub-blocks driven block by block with a per-block w (gan_oracle.generator_forward(w_override=)).
    Checks the image and the gradients w.r.t. both latents and the parameters; crossover >= steps must equal the plain
    forward."""
    U.no_tf32()
    batch = 4
    g, _ = U.build_models(5)
    Gs = {k: v.cuda().requires_grad_() for k, v in O.make_state("gen", 5).items()}
    z1 = O.make_latents(batch, 61).cuda().requires_grad_()
    z2 = O.make_latents(batch, 62).cuda().requires_grad_()
    noise = [n.cuda() for n in O.make_noise(batch, steps, 63)]
    probe = torch.randn(batch, 3, 4 << (steps - 1), 4 << (steps - 1), device="cuda")
    img = g(z1, noise=noise, steps=steps, alpha=None, z2=z2, crossover=crossover)
    (img * probe).sum().backward()
    zo1, zo2 = z1.detach().clone().requires_grad_(), z2.detach().clone().requires_grad_()
    w1, w2 = O.mapping(Gs, zo1), O.mapping(Gs, zo2)
    img_o = O.generator_forward(Gs, zo1, noise, steps, None, w_override=[w1 if k < crossover else w2 for k in range(8)])
    (img_o * probe).sum().backward()
    assert U.rel(img, img_o) < U.TOL_IMG
    if crossover > 0:
        assert U.cos(z1.grad, zo1.grad) > 0.97, U.cos(z1.grad, zo1.grad)
    else:                                        # no block is styled by the first latent
        assert zo1.grad is None and (z1.grad is None or z1.grad.abs().max().item() == 0.0)
    if crossover < steps:
        assert U.cos(z2.grad, zo2.grad) > 0.97, U.cos(z2.grad, zo2.grad)
    else:
        assert z2.grad is None or z2.grad.abs().max().item() == 0.0
        with torch.no_grad():
            plain = g(z1.detach(), noise=noise, steps=steps, alpha=None)
        assert U.rel(img, plain) < 1.5e-2
    worst = 1.0
    for name, p in g.named_parameters():
        ref = Gs[name].grad
        if p.grad is None:
            assert ref is None or ref.abs().max().item() == 0.0, name
            continue
        if ref.norm().item() > 0:
            worst = min(worst, U.cos(p.grad, ref))
    assert worst > 0.93, worst


@pytest.mark.parametrize("steps,alpha,mix", [(5, None, False), (4, 0.4, False), (5, None, True)])
def test_generator_layerwise_gradient_emission(steps, alpha, mix):
    """The overlapped data-parallel path (Generator._grad_ready_hook = dist.GradSync.ready): every parameter gradient is
    put into .grad during the backward and reported exactly once, in an order that lets its all-reduce overlap the rest;
    the values match the ones autograd accumulates without the hook to run-to-run noise, also on a second backward that
    accumulates into existing .grad."""
    U.no_tf32()
    batch = 4
    g, _ = U.build_models(7)
    z1 = O.make_latents(batch, 71).cuda()
    z2 = O.make_latents(batch, 72).cuda()
    noise = [n.cuda() for n in O.make_noise(batch, steps, 73)]
    probe = torch.randn(batch, 3, 4 << (steps - 1), 4 << (steps - 1), device="cuda")
    kw = dict(z2=z2, crossover=2) if mix else {}

    def run(times):
        g.zero_grad()
        for _ in range(times):
            (g(z1, noise=noise, steps=steps, alpha=alpha, **kw) * probe).sum().backward()
        return {n: p.grad.clone() for n, p in g.named_parameters() if p.grad is not None}

    with U.deterministic():
        plain1, plain2 = run(1), run(2)
        seen = []
        g._grad_ready_hook = lambda p: seen.append(id(p))
        try:
            hooked1 = run(1)
            first = list(seen)
            hooked2 = run(2)
        finally:
            del g._grad_ready_hook
        control = run(1)
    assert set(hooked1) == set(plain1) and len(first) == len(plain1) == len(set(first)), "each gradient reported once"
    names = {id(p): n for n, p in g.named_parameters()}
    order = [names[i] for i in first]
    assert order[0].startswith("to_rgbs") and order[-1].startswith("to_w_noise"), order[:2] + order[-2:]
    # chain-deterministic mode: the activations and their gradients are bit-identical between the runs, the parameter
    # gradients (leaf sums with fp32 atomics) differ by order noise only; a wrong parameter mapping or a lost
    # accumulation would be an O(1) error.  (Without the switch this comparison needed 0.3-0.6: a last-bit difference in
    # an instance-norm sum flips bf16 roundings and LeakyReLU gates in the ~30 layers behind it.)

    def tol(t):
        return 2e-3

    for ref, got, what in ((plain1, control, "control"), (plain1, hooked1, "hooked"), (plain2, hooked2, "hooked x2")):
        for n in ref:
            assert U.rel(got[n], ref[n]) < tol(ref[n]), (what, n, U.rel(got[n], ref[n]))
    for n in plain1:
        assert U.rel(hooked2[n], 2 * hooked1[n]) < tol(plain1[n]), n
