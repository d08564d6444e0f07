"""CPU checks of the host logic around the hot path: checkpoint layout / resume state,
data-feed sharding, fade-in schedule.  No kernels run here."""
import os

import pytest
import torch

import checkpoint as ckpt
import data
import gan
import trainer
from oracle import gan_oracle as O


def test_shard_indices_partition_every_epoch_like_distributed_sampler():
    n, world = 37, 4
    for epoch in range(3):
        parts = [data.shard_indices(n, epoch, r, world, seed=5) for r in range(world)]
        assert len({len(p) for p in parts}) == 1 and len(parts[0]) == 10          # ceil(37 / 4), equal on every rank
        flat = [i for p in parts for i in p]
        assert set(flat) == set(range(n))                                           # everything is seen ...
        assert len(flat) - len(set(flat)) == world * 10 - n                         # ... the pad wraps around
        assert parts == [data.shard_indices(n, epoch, r, world, seed=5) for r in range(world)]   # deterministic
    assert data.shard_indices(n, 0, 0, world, seed=5) != data.shard_indices(n, 1, 0, world, seed=5)
    assert data.shard_indices(6, 0, 1, 2, shuffle=False) == [1, 3, 5]
    assert data.shard_indices(5, 0, 0, 1, shuffle=False) == [0, 1, 2, 3, 4]


def test_fade_alpha_follows_train_py():
    assert trainer.fade_alpha(0, 100.0) == 0.0                                     # train.py:141
    assert trainer.fade_alpha(50, 100.0) == 0.5
    assert trainer.fade_alpha(100, 100.0) == 1.0
    assert trainer.fade_alpha(101, 100.0) is None                                  # train.py:143-145
    assert trainer.fade_alpha(10, 0.0) is None


def _models_and_opts():
    g, c = gan.Generator(), gan.Critic()
    g.load_state_dict(O.make_state("gen", 11))
    c.load_state_dict(O.make_state("critic", 11))
    g_opt = torch.optim.Adam([{"params": g.to_w_noise.parameters(), "lr": 2e-5}, {"params": g.gen_blocks.parameters()},
                              {"params": g.to_rgbs.parameters()}], lr=0.002, betas=(0.0, 0.99))
    c_opt = torch.optim.Adam(c.parameters(), lr=0.002, betas=(0.0, 0.99))
    for opt in (g_opt, c_opt):                                  # give a few parameters optimizer state
        ps = [p for grp in opt.param_groups for p in grp["params"]][:5]
        for p in ps:
            p.grad = torch.randn_like(p)
        opt.step()
        opt.zero_grad()
    return g, c, g_opt, c_opt


def test_checkpoint_keeps_reference_layout_and_restores_optimizer_state(tmp_path):
    g, c, g_opt, c_opt = _models_and_opts()
    state = ckpt.snapshot(torch.nn.DataParallel(g), c, iters=7, im_count=224, step=3, epoch=1, alpha=0.35,
                          gen_opt=g_opt, critic_opt=c_opt)
    # train.py:247-259: exactly these keys (+ the two optional ones), module.-prefixed fp32 state_dicts
    assert set(state) == {"gen", "critic", "iter", "im_count", "step", "epoch", "alpha", "gen_opt", "critic_opt"}
    assert all(k.startswith("module.") for k in state["gen"]) and all(k.startswith("module.") for k in state["critic"])
    assert [k[len("module."):] for k in state["gen"]] == list(O.generator_param_shapes())
    path = os.path.join(tmp_path, "checkpoints", "chk-7.pth")
    saver = ckpt.AsyncSaver()
    saver.save(state, path)
    saver.wait()
    assert os.path.exists(path) and not os.path.exists(path + ".tmp")
    # what generate_samples.py:48-52 does with the file
    save = torch.load(path)
    wrapped = torch.nn.DataParallel(gan.Generator())
    wrapped.load_state_dict(save["gen"])
    assert (save["step"], save["alpha"], save["iter"], save["im_count"]) == (3, 0.35, 7, 224)
    # resume into fresh objects: weights, Adam moments and step counts
    g2, c2 = gan.Generator(), torch.nn.DataParallel(gan.Critic())
    g2_opt = torch.optim.Adam([{"params": g2.to_w_noise.parameters(), "lr": 2e-5}, {"params": g2.gen_blocks.parameters()},
                               {"params": g2.to_rgbs.parameters()}], lr=0.002, betas=(0.0, 0.99))
    c2_opt = torch.optim.Adam(c2.parameters(), lr=0.002, betas=(0.0, 0.99))
    info = ckpt.load(path, g2, c2, g2_opt, c2_opt)
    assert info == {"iter": 7, "im_count": 224, "step": 3, "epoch": 1, "alpha": 0.35, "has_optimizer_state": True}
    for (k, a), (_, b) in zip(g.state_dict().items(), g2.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(c.state_dict().items(), c2.module.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = c_opt.state_dict()["state"], c2_opt.state_dict()["state"]
    assert set(s1) == set(s2) and len(s1) == 5
    for i in s1:
        assert torch.equal(s1[i]["exp_avg_sq"], s2[i]["exp_avg_sq"]) and float(s1[i]["step"]) == float(s2[i]["step"])
    assert [grp["lr"] for grp in g2_opt.param_groups] == [2e-5, 0.002, 0.002]


def test_checkpoint_loads_reference_files_without_optimizer_state(tmp_path):
    """A file exactly as the reference writes it (train.py:247-259: no optimizer keys)."""
    path = os.path.join(tmp_path, "FINAL.pth")
    torch.save({"gen": {"module." + k: v for k, v in O.make_state("gen", 12).items()},
                "critic": {"module." + k: v for k, v in O.make_state("critic", 12).items()},
                "iter": 100, "im_count": 3200, "step": 2, "epoch": 4, "alpha": None}, path)
    g, c = gan.Generator(), gan.Critic()
    opt = torch.optim.Adam(c.parameters(), lr=0.002)
    info = ckpt.load(path, g, c, None, opt)
    assert info["has_optimizer_state"] is False and info["step"] == 2 and info["alpha"] is None
    assert torch.equal(g.state_dict()["to_rgbs.3.weight"], O.make_state("gen", 12)["to_rgbs.3.weight"])
    with pytest.raises(RuntimeError):                                   # a layout mismatch must not pass silently
        bad = torch.load(path)
        del bad["gen"]["module.to_rgbs.3.weight"]
        torch.save(bad, path)
        ckpt.load(path, gan.Generator(), None)


def test_async_saver_reports_writer_errors(tmp_path):
    saver = ckpt.AsyncSaver()
    target = os.path.join(tmp_path, "file")
    open(target, "w").close()
    saver.save({"x": 1}, os.path.join(target, "sub", "chk.pth"))      # parent is a file: the writer thread fails
    with pytest.raises(BaseException):
        saver.wait()
