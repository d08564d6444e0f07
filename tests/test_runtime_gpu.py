"""GPU checks of the rows either side of the hot path: device-side data feed (train.py:43-50,109-117),
checkpoint resume with optimizer state (train.py:90-100,247-259), the loop driver (train.py:132-259) and the host-side
overhead shims (helper.get_truncated_noise, deferred loss reads)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import gan_oracle as O  # noqa: E402
import parity_util as U  # noqa: E402


def _u8_images(n, r, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, r, r, 3), generator=g, dtype=torch.uint8)


def test_device_transform_equals_the_reference_transforms():
    """train.py:43-50: RandomHorizontalFlip -> ToTensor -> Normalize(.5,.5) -> ConvertImageDtype(float), restated with
    torch ops on the host (torchvision's ToTensor is x / 255, Normalize is (x - .5) / .5) — bit-for-bit up to 1 ulp."""
    import data

    for r, b in ((4, 5), (16, 8), (64, 3)):
        u8 = _u8_images(b, r, r)
        flip = (torch.arange(b) % 2).to(torch.uint8)
        got = data.device_transform(u8.cuda(), flip.cuda()).cpu()
        x = u8.permute(0, 3, 1, 2).float() / 255.0
        x = torch.where(flip.bool()[:, None, None, None], x.flip(3), x)
        want = (x - 0.5) / 0.5
        assert got.shape == (b, 3, r, r) and got.dtype == torch.float32
        assert (got - want).abs().max().item() <= 2.4e-7, (r, (got - want).abs().max().item())
        assert torch.equal(data.device_transform(u8.cuda(), None).cpu(), data.device_transform(u8.cuda(), torch.zeros(b, dtype=torch.uint8).cuda()).cpu())


def test_image_feed_covers_the_dataset_once_per_epoch_and_shards_by_rank():
    import data

    n, r, batch = 22, 8, 4
    u8 = _u8_images(n, r, 3)
    u8[:, 0, 0, 0] = torch.arange(n, dtype=torch.uint8)                 # tag every image in its first pixel
    src = data.TensorSource(u8)
    seen = []
    for rank in range(2):
        feed = data.ImageFeed(src, batch, "cuda", rank=rank, world=2, seed=9, flip=False)
        assert len(feed) == 3                                             # ceil(ceil(22 / 2) / 4)
        tags = []
        for real in feed:
            assert real.shape[1:] == (3, r, r) and real.is_cuda and real.dtype == torch.float32
            assert -1.0 - 1e-6 <= real.min().item() and real.max().item() <= 1.0 + 1e-6
            tags += [int(round((v + 1.0) * 127.5)) for v in real[:, 0, 0, 0].tolist()]
        assert len(tags) == 11
        seen.append(tags)
        again = [int(round((v + 1.0) * 127.5)) for real in feed for v in real[:, 0, 0, 0].tolist()]
        assert again != tags and sorted(again) != list(range(11))         # second epoch: a new permutation
    assert set(seen[0]) | set(seen[1]) == set(range(n))
    # flip=True mirrors about half of the samples and nothing else
    feed = data.ImageFeed(src, n, "cuda", seed=1, flip=True, shuffle=False)
    (real,) = list(feed)
    plain = data.device_transform(u8.cuda(), None)
    mirrored = [not torch.equal(real[i], plain[i]) for i in range(n)]
    assert 3 <= sum(mirrored) <= n - 3
    for i in range(n):
        assert torch.equal(real[i], plain[i].flip(2) if mirrored[i] else plain[i])
    # a stage that trains at a lower resolution than the prepared set: bilinear resize like train.py:150-156
    feed = data.ImageFeed(src, 8, "cuda", flip=False, shuffle=False, resolution=4)
    first = next(iter(feed))
    assert first.shape == (8, 3, 4, 4)
    assert torch.allclose(first, torch.nn.functional.interpolate(plain[:8], size=(4, 4), mode="bilinear"))


def test_truncated_noise_sampler_matches_the_truncated_normal():
    """helper.get_truncated_noise (helper.py:36-45: scipy truncnorm.rvs(-t, t)) sampled on the device."""
    import helper
    from scipy.stats import truncnorm

    torch.manual_seed(0)
    z = helper.get_truncated_noise(4096, 512, 0.75)
    assert z.shape == (4096, 512) and z.is_cuda and z.dtype == torch.float32 and z.requires_grad
    v = z.detach().flatten().cpu().double()
    assert v.abs().max().item() <= 0.75
    assert abs(v.mean().item()) < 2e-3
    assert abs(v.std().item() - truncnorm.std(-0.75, 0.75)) < 2e-3
    qs = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9], dtype=torch.float64)
    want = torch.tensor(truncnorm.ppf(qs.numpy(), -0.75, 0.75))
    assert (torch.quantile(v[:1_000_000], qs) - want).abs().max().item() < 3e-3


def test_resume_restores_optimizer_state_and_fade_position(tmp_path):
    """Train 2 iterations, checkpoint, train a 3rd; a fresh Trainer resumed from the file must take the SAME 3rd step
    (weights and Adam moments restored).  Resuming from a reference-style file (weights only) must NOT: with betas
    (0, .99) the second-moment estimate restarts from zero and the update is several times larger."""
    import checkpoint as ckpt
    import trainer

    U.no_tf32()
    steps, batch = 3, 8
    reals = [O.make_images(batch, steps, 200 + i).cuda() for i in range(3)]
    zs = [(O.make_latents(batch, 210 + i).cuda(), O.make_latents(batch, 220 + i).cuda()) for i in range(3)]

    def step(tr, i):
        torch.manual_seed(300 + i)                       # the per-layer noise comes from torch's global generator
        tr.iteration(reals[i].clone(), zs[i][0].clone(), zs[i][1].clone(), read_losses=False)

    with U.deterministic():
        a = trainer.Trainer(steps, 0.5, batch, "cuda", perturb_init=True)
        step(a, 0)
        step(a, 1)
        path = os.path.join(tmp_path, "chk-2.pth")
        ckpt.write(ckpt.snapshot(a.gen, a.critic, 2, 2 * batch, steps, 0, 0.5, a.gen_opt, a.critic_opt), path)
        before = {k: v.clone() for k, v in a.critic.state_dict().items()}
        step(a, 2)
        b = trainer.Trainer(steps, 0.5, batch, "cuda", seed=123)                       # different init: must be overwritten
        info = ckpt.load(path, b.gen, b.critic, b.gen_opt, b.critic_opt)
        assert info["has_optimizer_state"] and info["im_count"] == 2 * batch and info["iter"] == 2
        step(b, 2)
        ref_save = torch.load(path)
        ref_save.pop("gen_opt")
        ref_save.pop("critic_opt")
        path_ref = os.path.join(tmp_path, "chk-2-weights-only.pth")
        torch.save(ref_save, path_ref)
        c = trainer.Trainer(steps, 0.5, batch, "cuda", seed=123)
        assert not ckpt.load(path_ref, c.gen, c.critic, c.gen_opt, c.critic_opt)["has_optimizer_state"]
        step(c, 2)
    worst_resumed, worst_cold = 0.0, 0.0
    for (k, va), (_, vb), (_, vc) in zip(a.critic.state_dict().items(), b.critic.state_dict().items(), c.critic.state_dict().items()):
        delta = (va - before[k]).norm().item()
        if delta == 0:
            continue
        worst_resumed = max(worst_resumed, (vb - va).norm().item() / delta)
        worst_cold = max(worst_cold, (vc - va).norm().item() / delta)
    assert worst_resumed < 2e-2, worst_resumed          # same step up to the atomics' order noise in the leaf gradients
    assert worst_cold > 0.5, worst_cold                 # without the moments the step is a different one


def test_loop_driver_runs_the_progressive_schedule(tmp_path):
    """trainer.run: train.py:102-259 over two stages (4x4 then 8x8 with fade-in) fed by data.ImageFeed, previews only at
    display steps, asynchronous checkpoints in the reference layout, resume from the last one."""
    import data
    import trainer

    U.no_tf32()
    config = {"gradient_lambda": 10, "lr": 0.002, "beta_1": 0.0, "beta_2": 0.99, "use_r1": "True", "display_step": 3,
              "checkpoint_step": 4, "batch_progression": "8,8", "epoch_progression": "2,1", "fade_percentage": 0.5}
    src = {1: data.TensorSource(_u8_images(24, 4, 1)), 2: data.TensorSource(_u8_images(24, 8, 2))}
    previews, saved = [], []

    def feed_for_stage(steps, batch):
        return data.ImageFeed(src[steps], batch, "cuda", seed=steps)

    def on_checkpoint(stub, state):
        import checkpoint as ckpt

        path = os.path.join(tmp_path, stub + ".pth")
        ckpt.write(state, path)
        saved.append(path)

    iters, hist = trainer.run(config, feed_for_stage, on_preview=lambda it, im: previews.append((it, tuple(im.shape))),
                              on_checkpoint=on_checkpoint)
    assert iters == 2 * 3 + 1 * 3                                         # epochs x batches per stage
    assert [p[0] for p in previews] == [3, 6, 9] and previews[0][1] == (25, 3, 4, 4) and previews[2][1] == (25, 3, 8, 8)
    assert len(hist) == iters - 1 and all(torch.isfinite(torch.tensor(h)).all() for h in hist)   # read one step late
    assert [os.path.basename(p) for p in saved] == ["chk-4.pth", "chk-8.pth"]
    save = torch.load(saved[-1])
    assert save["step"] == 2 and save["iter"] == 8 and save["im_count"] == 16 and "critic_opt" in save
    assert save["alpha"] is None or 0.0 < save["alpha"] <= 1.0
    iters2, _ = trainer.run(config, feed_for_stage, checkpoint_path=saved[-1], on_checkpoint=lambda *a: None)
    # like train.py:125-128 the resume is at EPOCH granularity: the epoch the checkpoint was taken in runs again in full
    # (3 batches), now with the restored optimizer state and fade-in position
    assert iters2 == 8 + 3


def test_wgan_gp_training_iteration_through_the_trainer():
    """use_r1=False (config.txt) -> train.py:177-185,213: the path that raises in the reference runs here."""
    import trainer

    tr = trainer.Trainer(3, 0.4, 8, "cuda", use_r1=False, perturb_init=True)
    w0 = tr.critic.conv_blocks[7].conv_2[0].weight.detach().clone()
    for i in range(2):
        tr.iteration(O.make_images(8, 3, i).cuda(), O.make_latents(8, 10 + i).cuda(), O.make_latents(8, 20 + i).cuda())
    vals = tr.flush_reads()
    assert all(torch.isfinite(torch.tensor(vals)))
    assert (tr.critic.conv_blocks[7].conv_2[0].weight - w0).abs().max().item() > 0


@pytest.mark.parametrize("mix", [False, True], ids=["plain", "style-mixing"])
def test_cuda_graph_iterations_equal_eager_iterations(mix):
    """Trainer.enable_graphs(): the whole iteration (forward, R1 double-backward, both Adam updates, weight re-packing,
    internally drawn noise) replayed from a CUDA graph must train exactly like the eager call sequence.  Same seeds, same
    batches, deterministic mode: weights after 5 iterations agree to the leaf-gradient order noise; an eager forward
    between replays sees the current weights (pack cache invalidated by the replay)."""
    import trainer

    U.no_tf32()
    steps, batch, n_it = 4, 8, 5
    reals = [O.make_images(batch, steps, 400 + i).cuda() for i in range(n_it)]
    zs = [(O.make_latents(batch, 410 + i).cuda(), O.make_latents(batch, 420 + i).cuda()) for i in range(n_it)]

    def run(graph):
        tr = trainer.Trainer(steps, None, batch, "cuda", perturb_init=True, style_mixing=mix, capturable=True, seed=5)
        if graph:
            tr.enable_graphs()
        probes = []
        for i in range(n_it):
            tr.iteration(reals[i].clone(), zs[i][0].clone(), zs[i][1].clone(), read_losses=True)
            if i in (1, 3):
                with torch.no_grad():
                    probes.append(tr.gen(zs[0][0], noise=[n.cuda() for n in O.make_noise(batch, steps, 5)], steps=steps))
        return tr, probes, tr.flush_reads()

    # the per-layer noise is drawn from torch's global generator inside the iteration: the graph registers the generator
    # and advances its offset per replay, so the eager run and the graphed run see different noise streams; freeze the
    # noise by comparing on the deterministic parts only would hide bugs, so instead give both runs the same stream:
    with U.deterministic():
        torch.manual_seed(77)
        a, probes_a, losses_a = run(False)
        torch.manual_seed(77)
        b, probes_b, losses_b = run(True)
    assert len(b.graphs) == (3 if mix else 1)
    for x, y in zip(losses_a, losses_b):
        assert abs(x - y) < 0.15 * abs(x) + 1e-3, (losses_a, losses_b)
    # noise streams differ between eager and captured RNG consumption, so weights are compared statistically: every tensor
    # must have moved, by a similar amount, and mostly in the same direction
    init = trainer.Trainer(steps, None, batch, "cuda", perturb_init=True, style_mixing=mix, capturable=True, seed=5)
    worst = 1.0
    for (k, va), (_, vb), (_, v0) in zip(a.critic.state_dict().items(), b.critic.state_dict().items(), init.critic.state_dict().items()):
        da, db = va - v0, vb - v0
        if da.norm() == 0:
            assert db.norm() == 0, k
            continue
        assert 0.5 < (db.norm() / da.norm()).item() < 2.0, (k, da.norm().item(), db.norm().item())
        if da.numel() >= 4096:
            worst = min(worst, U.cos(da, db))
    assert worst > 0.5, worst
    for pa, pb in zip(probes_a, probes_b):
        assert U.rel(pb, pa) < 0.2, U.rel(pb, pa)
    assert U.rel(probes_b[1], probes_b[0]) > 1e-3          # the eager preview after more replays sees newer weights


def test_deferred_item_and_lazy_preview_shims():
    """lazy.py (opt-in, for an unmodified train.py): `.item()` on the losses returns a float-like that synchronises only
    when used; no_grad forwards run when their result is first touched, with the SAME images as the eager forward (the
    noise is drawn up front from torch's generator exactly as the reference does)."""
    import gan
    import lazy

    U.no_tf32()
    g, c = U.build_models(4)
    steps, batch = 3, 8
    z = O.make_latents(25, 7).cuda()
    try:
        gan.LAZY_NO_GRAD_FORWARD = True
        gan.DEFER_LOSS_ITEMS = True
        with U.deterministic():
            torch.manual_seed(11)
            with torch.no_grad():
                lazy_img = g(z, alpha=0.3, steps=steps)                   # train.py:237
            assert isinstance(lazy_img, lazy.LazyImages) and not lazy_img.computed
            assert lazy_img.shape == (25, 3, 16, 16) and lazy_img.is_cuda and len(lazy_img) == 25
            after = torch.randn(4, device="cuda")                          # the RNG stream has moved past the noise draws
            shown = torch.clamp(lazy_img, 0, 1)                            # train.py:239: first use runs the kernels
            assert lazy_img.computed and type(shown) is torch.Tensor
            gan.LAZY_NO_GRAD_FORWARD = False
            torch.manual_seed(11)
            with torch.no_grad():
                eager = g(z, alpha=0.3, steps=steps)
            assert torch.equal(torch.randn(4, device="cuda"), after)
            assert torch.equal(shown, torch.clamp(eager, 0, 1))
            # with grad enabled nothing is deferred (training forwards)
            gan.LAZY_NO_GRAD_FORWARD = True
            assert type(g(z.clone().requires_grad_(), steps=steps)) is torch.Tensor
        # losses: .item() is a LazyScalar that behaves like the float train.py expects
        real = O.make_images(batch, steps, 3).cuda().requires_grad_()
        fake = g(O.make_latents(batch, 8).cuda().requires_grad_(), steps=steps)
        c.zero_grad()
        c_loss = c.get_r1_loss(c(fake.detach(), steps, None), c(real, steps, None), real, fake, steps, None, 10)
        g_loss = g.get_r1_loss(c(fake, steps, None))
        g_loss.backward()                                                  # still an ordinary differentiable tensor
        hist = [c_loss.item(), g_loss.item()]
        assert all(isinstance(h, lazy.LazyScalar) for h in hist)
        avg = sum(hist[-2:]) / 2                                           # train.py:222-227
        assert isinstance(avg, float) and abs(avg - (float(c_loss) + float(g_loss)) / 2) < 1e-6
        assert f"{hist[0]:.3}" == f"{float(c_loss):.3}"
        gan.DEFER_LOSS_ITEMS = False
        assert isinstance(g.get_r1_loss(c(fake.detach(), steps, None)).item(), float)
    finally:
        gan.LAZY_NO_GRAD_FORWARD = False
        gan.DEFER_LOSS_ITEMS = False
