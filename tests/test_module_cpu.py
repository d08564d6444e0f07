"""CPU checks of the drop-in boundary: attribute tree / state_dict layout of byo-gan_b200/gan.py against the
reference's (tests/golden/state_layout.json), checkpoint round trips, C-ABI symbol export, error behaviour."""
import ctypes
import json
import os
import re

import pytest
import torch

import gan  # byo-gan_b200/gan.py (conftest puts it first on sys.path)
from oracle import gan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_state_dict_layout_equals_reference():
    layout = json.load(open(os.path.join(GOLD, "state_layout.json")))
    g, c = gan.Generator(), gan.Critic()
    assert {k: list(v.shape) for k, v in g.state_dict().items()} == layout["gen"]
    assert {k: list(v.shape) for k, v in c.state_dict().items()} == layout["critic"]
    assert list(g.state_dict().keys()) == list(layout["gen"].keys())          # same ORDER as the reference
    assert list(c.state_dict().keys()) == list(layout["critic"].keys())
    assert all(v.dtype == torch.float32 for v in g.state_dict().values())


def test_reference_checkpoint_layout_loads_strict_with_dataparallel_prefix(tmp_path):
    """train.py:250-251 saves DataParallel state_dicts (keys prefixed 'module.'); generate_samples.py:48-52
    loads them back into a DataParallel-wrapped Generator."""
    g = torch.nn.DataParallel(gan.Generator())
    c = torch.nn.DataParallel(gan.Critic())
    save = {"gen": {"module." + k: v for k, v in O.make_state("gen", 3).items()},
            "critic": {"module." + k: v for k, v in O.make_state("critic", 3).items()},
            "iter": 7, "im_count": 70, "step": 3, "epoch": 1, "alpha": 0.25}
    path = tmp_path / "chk-7.pth"
    torch.save(save, path)
    back = torch.load(path)
    g.load_state_dict(back["gen"])
    c.load_state_dict(back["critic"])
    for k, v in g.state_dict().items():
        assert torch.equal(v, save["gen"][k])
    assert hasattr(g.module, "get_r1_loss") and hasattr(c.module, "get_r1_loss") and hasattr(c.module, "get_wgan_loss")


def test_optimizer_groups_and_requires_grad_toggle():
    """train.py:59-70 builds three Adam groups from these attributes; helper.py:48-50 toggles requires_grad."""
    g = gan.Generator()
    n = sum(p.numel() for grp in (g.to_w_noise, g.gen_blocks, g.to_rgbs) for p in grp.parameters())
    assert n == sum(p.numel() for p in g.parameters())
    assert len(list(g.parameters())) == 111 and len(list(gan.Critic().parameters())) == 52
    for p in g.parameters():
        p.requires_grad = False
    assert not any(p.requires_grad for p in g.parameters())


def test_initialisation_follows_reference_rules():
    torch.manual_seed(0)
    g, c = gan.Generator(), gan.Critic()
    sd = g.state_dict()
    assert sd["gen_blocks.3.conv_1.inject_noise.weights"].abs().sum() == 0                 # gan.py:44
    assert sd["gen_blocks.3.conv_1.conv.bias"].abs().sum() == 0                            # gan.py:24
    b = sd["gen_blocks.3.conv_1.adain.style.bias"]
    assert torch.equal(b[:256], torch.ones(256)) and torch.equal(b[256:], torch.zeros(256))  # gan.py:62-63
    w = sd["gen_blocks.3.conv_1.conv.weight"]
    assert abs(w.std().item() - 1.0) < 0.01 and abs(w.mean().item()) < 0.01                  # gan.py:23
    assert c.conv_blocks[7].conv_1[0].group_size == 4                                        # gan.py:269
    assert g.gen_blocks[3].conv_1.conv.equalized_coefficient == pytest.approx((2 / (512 * 9)) ** 0.5)


def test_no_cpu_fallback_and_reference_error_behaviour():
    g, c = gan.Generator(), gan.Critic()
    with pytest.raises(RuntimeError, match="CUDA"):
        g(torch.zeros(2, 512))
    with pytest.raises(RuntimeError, match="CUDA"):
        c(torch.zeros(2, 3, 4, 4))
    with pytest.raises(RuntimeError, match="forward"):                 # WGAN-GP is implemented: same contract
        c.get_wgan_loss(torch.zeros(2, 1), torch.zeros(2, 1), torch.zeros(2, 3, 4, 4), 1, None)
    with pytest.raises(ValueError):                                    # gan.py:105-106
        gan.StyleGanBlock(4, 4, is_initial=True, does_upsample=True)
    with pytest.raises(RuntimeError, match="forward"):
        c.get_r1_loss(torch.zeros(2, 1), torch.zeros(2, 1), None, None, 1, None)


def test_dataparallel_replicas_are_rejected_with_an_explanation():
    """Unmodified train.py on a multi-GPU box would wrap the modules in a multi-device nn.DataParallel (train.py:71,79),
    whose replicas hold plain tensors instead of Parameters; forward must say 'one process per GPU' instead of failing
    with an opaque AttributeError deep inside the parameter walk."""
    for m, x in ((gan.Generator(), torch.zeros(2, 512)), (gan.Critic(), torch.zeros(2, 3, 4, 4))):
        m._is_replica = True                       # what torch.nn.parallel.replicate() sets on its copies
        with pytest.raises(RuntimeError, match="one process per GPU"):
            m(x)


def test_load_state_dict_drops_cached_weight_packs():
    g = gan.Generator()
    g._packs._conv["sentinel"] = ("tag", None, None)
    g._packs._lin["sentinel"] = ("tag", None)
    g.load_state_dict(O.make_state("gen", 4))
    assert not g._packs._conv and not g._packs._lin
    g._packs._conv["sentinel"] = ("tag", None, None)
    g.invalidate_packs()
    assert not g._packs._conv
    c = gan.Critic()
    c._packs._conv["sentinel"] = ("tag", None, None)
    torch.nn.DataParallel(c).load_state_dict({"module." + k: v for k, v in O.make_state("critic", 4).items()})
    assert not c._packs._conv


def test_pack_staleness_sees_fused_optimizer_steps():
    """torch's multi-tensor optimizers (fused=True) update parameters WITHOUT moving their version counter (observed on
    torch 2.11, CPU and CUDA): the pack cache must notice the step through its optimizer post-hook, or the convolutions
    would keep using the pre-step bf16 weights while the fp32 masters train."""
    import engine

    for fused in (False, True):
        p = torch.nn.Parameter(torch.randn(8, 8))
        opt = torch.optim.Adam([p], lr=0.1, fused=fused)
        p.grad = torch.randn(8, 8)
        before = engine._tag(p)
        opt.step()
        assert engine._tag(p) != before, f"optimizer step (fused={fused}) not visible to the pack cache"
    q = torch.nn.Parameter(torch.randn(4))
    before = engine._tag(q)
    with torch.no_grad():
        q.mul_(2.0)
    assert engine._tag(q) != before


REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "gan.py")), reason="the reference checkout only exists in the build container")
def test_checkpoints_cross_load_with_the_real_reference_classes(tmp_path):
    """SURVEY §8c KAT 10 with the REAL reference modules: a checkpoint written by train.py's save call (train.py:247-259)
    from reference modules loads strictly into the drop-in, and one written from the drop-in loads strictly into the
    reference's Generator / Critic (generate_samples.py:48-52), with and without the DataParallel 'module.' prefix —
    same keys, same order, same shapes, same values."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_gan_for_kat10", os.path.join(REF, "gan.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(3)
    rg, rc = ref.Generator(), ref.Critic()
    torch.manual_seed(3)
    ng, nc = gan.Generator(), gan.Critic()
    # same construction order and init rules: identical initial weights from the same seed
    for (ka, va), (kb, vb) in zip(rg.state_dict().items(), ng.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    for (ka, va), (kb, vb) in zip(rc.state_dict().items(), nc.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    # reference -> drop-in (DataParallel-wrapped on both sides, as train.py saves and generate_samples.py loads)
    for p in list(rg.parameters()) + list(rc.parameters()):
        p.data.normal_()
    path = tmp_path / "chk-ref.pth"
    torch.save({"gen": torch.nn.DataParallel(rg).state_dict(), "critic": torch.nn.DataParallel(rc).state_dict(),
                "iter": 11, "im_count": 352, "step": 2, "epoch": 0, "alpha": 0.4}, path)
    save = torch.load(path)
    wg, wc = torch.nn.DataParallel(ng), torch.nn.DataParallel(nc)
    wg.load_state_dict(save["gen"], strict=True)
    wc.load_state_dict(save["critic"], strict=True)
    for k, v in rg.state_dict().items():
        assert torch.equal(ng.state_dict()[k], v), k
    # drop-in -> reference
    for p in list(ng.parameters()) + list(nc.parameters()):
        p.data.mul_(0.5)
    path2 = tmp_path / "chk-new.pth"
    torch.save({"gen": wg.state_dict(), "critic": wc.state_dict(), "iter": 12, "im_count": 384, "step": 2, "epoch": 0,
                "alpha": None}, path2)
    back = torch.load(path2)
    torch.nn.DataParallel(rg).load_state_dict(back["gen"], strict=True)
    torch.nn.DataParallel(rc).load_state_dict(back["critic"], strict=True)
    for k, v in nc.state_dict().items():
        assert torch.equal(rc.state_dict()[k], v), k
    # the three optimizer groups train.py:59-70 builds see the same parameter sets in both implementations
    for attr in ("to_w_noise", "gen_blocks", "to_rgbs"):
        assert [tuple(p.shape) for p in getattr(rg, attr).parameters()] == [tuple(p.shape) for p in getattr(ng, attr).parameters()]


def test_c_abi_library_exports_every_declared_symbol():
    """include/bg_b200.h <-> libbg_b200.so <-> bg_native.SIGNATURES agree (no compute calls: no GPU here)."""
    import bg_native

    header = open(os.path.join(ROOT, "include", "bg_b200.h")).read()
    declared = set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(os.path.join(ROOT, "byo-gan_b200", "libbg_b200.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in bg_b200.h but not exported"
    assert declared - {"bg_last_error", "bg_abi_version"} == set(bg_native.SIGNATURES) | set(bg_native.HOST_FUNCS)
    lib.bg_abi_version.restype = ctypes.c_int
    assert lib.bg_abi_version() >= 2
    # the deterministic-reduction switch is plain host state: it can be exercised without a GPU
    assert bg_native.set_deterministic(True) is False and bg_native.set_deterministic(False) is True
    # argument counts in the binding match the header declarations
    for name, args in list(bg_native.SIGNATURES.items()) + list(bg_native.HOST_FUNCS.items()):
        m = re.search(r"int\s+" + name + r"\s*\(([^;]*?)\)\s*;", header, re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]) == len(args), name
