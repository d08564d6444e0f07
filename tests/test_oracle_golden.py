"""Pins the oracle restatement (oracle/gan_oracle.py) to fixtures produced by the UNMODIFIED reference
gan.py (tests/golden/*.json, written by oracle/make_golden.py).  CPU only, fp32: tolerance 2e-5 relative
(same torch build, same ops; only association order of a few sums differs)."""
import json
import os

import pytest
import torch

from oracle import gan_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def check_fp(t, fp, what, tol=TOL):
    if fp is None:
        assert t is None, f"{what}: reference grad is None but the oracle produced a tensor"
        return
    assert t is not None, f"{what}: reference has a tensor, oracle has None"
    got = O.fingerprint(t)
    assert got["shape"] == fp["shape"], what
    scale = max(fp["absmax"], 1e-6)
    assert abs(got["norm"] - fp["norm"]) <= tol * max(fp["norm"], 1e-6) + 1e-9, f"{what}: norm {got['norm']} vs {fp['norm']}"
    ds = max(abs(a - b) for a, b in zip(got["samples"], fp["samples"]))
    assert ds <= 10 * tol * scale, f"{what}: sample diff {ds} (scale {scale})"


def test_state_layout_matches_reference_checkpoint_keys():
    """state_dict keys/shapes (train.py:250-251 saves them) — 111 G tensors, 52 D tensors."""
    layout = load("state_layout.json")
    assert {k: list(v) for k, v in O.generator_param_shapes().items()} == layout["gen"]
    assert {k: list(v) for k, v in O.critic_param_shapes().items()} == layout["critic"]
    assert len(layout["gen"]) == 111 and len(layout["critic"]) == 52


def test_layer_known_answers():
    g = load("layers.json")
    imp = torch.zeros(1, 1, 4, 4)
    imp[0, 0, 1, 2] = 1.0
    assert torch.allclose(O.bilinear_up2(imp)[0, 0], torch.tensor(g["impulse_up"]), atol=1e-7)
    # separable taps .25/.75 (SURVEY §8c KAT 3)
    assert torch.allclose(O.bilinear_up2(imp)[0, 0, 1:5, 3:7],
                          torch.outer(torch.tensor([.25, .75, .75, .25]), torch.tensor([.25, .75, .75, .25])))
    x = torch.tensor(g["in_x"])
    assert torch.allclose(O.instance_norm(x), torch.tensor(g["in_y"]), atol=2e-6)


def test_minibatch_stddev_sequence_with_stateful_group_size():
    rows = load("mbstd.json")
    gs = 4
    for r in rows:
        gen = torch.Generator().manual_seed(r["seed"])
        x = torch.randn(r["batch"], 512, 4, 4, generator=gen)
        y, gs = O.minibatch_stddev(x, gs)
        assert gs == r["group_size_after"], r
        assert y.shape == (r["batch"], 513, 4, 4)
        assert torch.allclose(y[:, 512, 0, 0], torch.tensor(r["plane"]), rtol=1e-5, atol=1e-6)
        assert torch.equal(y[:, 512], y[:, 512, :1, :1].expand(-1, 4, 4))


@pytest.mark.parametrize("case", load("forward.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_forward_matches_reference(case):
    steps, batch, alpha = case["steps"], case["batch"], case["alpha"]
    G, D = O.make_state("gen", 1), O.make_state("critic", 1)
    with torch.no_grad():
        fake = O.generator_forward(G, O.make_latents(batch, steps), O.make_noise(batch, steps, steps), steps, alpha)
        pr = O.critic_forward(D, O.make_images(batch, steps, steps), steps, alpha)
        pf = O.critic_forward(D, fake, steps, alpha)
    check_fp(fake, case["fake"], "fake image")
    check_fp(pr, case["pred_real"], "D(real)", tol=1e-4)
    check_fp(pf, case["pred_fake"], "D(fake)", tol=1e-4)


@pytest.mark.parametrize("case", load("train_iteration.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_train_iteration_matches_reference(case):
    steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
    r = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2),
                          O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps),
                          O.make_images(batch, steps, 30 + steps),
                          O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps),
                          steps, alpha, lam)
    assert abs(r["c_loss"].item() - case["c_loss"]) <= 1e-4 * abs(case["c_loss"])
    assert abs(r["g_loss"].item() - case["g_loss"]) <= 1e-4 * abs(case["g_loss"])
    check_fp(r["fake_d"], case["fake_d"], "fake (D step)")
    check_fp(r["z_grad"], case["z_grad"], "dL/dz", tol=2e-4)
    assert set(r["d_grads"]) == set(case["d_grads"]) and set(r["g_grads"]) == set(case["g_grads"])
    for k, fp in case["d_grads"].items():
        check_fp(r["d_grads"][k], fp, f"D grad {k}", tol=2e-4)
    for k, fp in case["g_grads"].items():
        check_fp(r["g_grads"][k], fp, f"G grad {k}", tol=2e-4)


def _iteration_inputs(steps, batch):
    return (O.make_latents(batch, 10 + steps), O.make_latents(batch, 20 + steps), O.make_images(batch, steps, 30 + steps),
            O.make_noise(batch, steps, 10 + steps), O.make_noise(batch, steps, 20 + steps))


@pytest.mark.parametrize("case", load("r1_penalty.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_r1_penalty_alone_matches_reference(case):
    """The penalty term of gan.py:398-404 by itself: purely second-order parameter gradients."""
    steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
    r = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *_iteration_inputs(steps, batch), steps, alpha,
                          lam, loss="r1_penalty")
    assert abs(r["c_loss"].item() - case["penalty"]) <= 1e-4 * abs(case["penalty"])
    for k, fp in case["d_grads"].items():
        check_fp(r["d_grads"][k], fp, f"D grad {k}", tol=2e-4)


@pytest.mark.parametrize("case", load("wgan_gp.json"), ids=lambda c: f"s{c['steps']}-b{c['batch']}-a{c['alpha']}")
def test_wgan_gp_iteration_matches_repaired_reference(case):
    """gan.py:357-391 executed on the reference's modules with its two defects repaired (oracle/make_golden.py)."""
    steps, batch, alpha, lam = case["steps"], case["batch"], case["alpha"], case["lambda"]
    r = O.train_iteration(O.make_state("gen", 2), O.make_state("critic", 2), *_iteration_inputs(steps, batch), steps, alpha,
                          lam, loss="wgan", epsilon=O.make_epsilon(batch, 40 + steps))
    assert abs(r["c_loss"].item() - case["c_loss"]) <= 1e-4 * abs(case["c_loss"])
    assert abs(r["g_loss"].item() - case["g_loss"]) <= 1e-4 * (abs(case["g_loss"]) + 1e-2)
    for k, fp in case["d_grads"].items():
        check_fp(r["d_grads"][k], fp, f"D grad {k}", tol=2e-4)
    for k, fp in case["g_grads"].items():
        check_fp(r["g_grads"][k], fp, f"G grad {k}", tol=2e-4)
