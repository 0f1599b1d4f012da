"""CPU: the oracle (oracle/) against the golden vectors produced by executing the reference's
own function bodies (tests/golden/make_golden.py).  Integer / index outputs bit-exact;
floating point within the tolerance written at each assert."""
import numpy as np
import pytest
import torch

import oracle
from oracle import assign as oassign
from oracle import boxes as oboxes
from oracle import crop as ocrop
from oracle import emd as oemd
from oracle import head as ohead
from oracle import hooks as ohooks
from tests._golden import CROP_CASES, load, procedural_image


def test_expand_bbox_bit_exact():
    g = load("boxes")
    for tag, coef, ratio in (("c05_r1", 0.5, 1), ("c11_r1", 1.1, 1), ("c05_r12", 0.5, 1.2)):
        got = np.array([oboxes.expand_bbox(b, coef, ratio) for b in g["boxes"]], dtype=np.int64)
        assert np.array_equal(got, g["expanded_" + tag])


def test_largest_face_bit_exact():
    g = load("boxes")
    got = [oboxes.largest_face_index(g["multi"][i, :g["counts"][i]], 512) for i in range(g["multi"].shape[0])]
    assert np.array_equal(np.array(got), g["picked"])


@pytest.mark.parametrize("idx", range(len(CROP_CASES)))
def test_crop_face_and_grad(idx):
    g = load("crop")
    seed, H, W, box, o = CROP_CASES[idx]
    img = torch.tensor(procedural_image(seed, 3, H, W), requires_grad=True)
    chip = ocrop.crop_face(img, box, [o, o], -1)
    assert np.array_equal(chip.detach().numpy(), g[f"chip_{idx}"])          # same torch ops -> identical
    direct = ocrop.direct_sampler(img.detach(), box, (o, o), -1.0)
    assert (direct - chip.detach().double()).abs().max() < 2e-5              # independent sampler, fp32 rounding
    up = torch.tensor(procedural_image(seed + 100, 3, o, o))
    (chip * up).sum().backward()
    grad = img.grad.numpy()
    if H <= 128:
        np.testing.assert_allclose(grad, g[f"grad_{idx}"], rtol=1e-6, atol=1e-6)
    else:
        np.testing.assert_allclose(grad[:, 192:256, 128:192], g[f"gradwin_{idx}"], rtol=1e-6, atol=1e-6)


def test_resize_small():
    g = load("crop")
    imgs = torch.tensor(procedural_image(31, 3, 512, 512)[None])
    assert np.array_equal(ocrop.resize_small(imgs, 224).numpy()[0], g["small"])


@pytest.mark.parametrize("tag,fn,kh", [("e1", ohead.get_face_gender, 80), ("e3", ohead.get_face_gender_race, 6),
                                       ("e4", ohead.get_face_gender_race_age, 8)])
def test_heads(tag, fn, kh):
    g = load("heads")
    sel = torch.tensor(g["selector"])
    chips = torch.zeros(sel.shape[0], 1)
    outs = fn(lambda x: torch.tensor(g[f"{tag}_logits_in"]), chips, selector=sel, fill_value=-1)
    for k, o in enumerate(outs):
        ref = g[f"{tag}_sel_{k}"]
        assert o.numpy().dtype == ref.dtype
        assert np.array_equal(o.numpy(), ref)
    outs = fn(lambda x: torch.tensor(g[f"{tag}_logits_in_full"]), chips, selector=None, fill_value=-1)
    for k, o in enumerate(outs):
        assert np.array_equal(o.numpy(), g[f"{tag}_nosel_{k}"])
    outs = fn(lambda x: 1 / 0, chips, selector=torch.zeros(sel.shape[0], dtype=torch.bool), fill_value=-1)
    for k, o in enumerate(outs):
        assert np.array_equal(o.numpy(), g[f"{tag}_empty_{k}"])


def test_assign_e1():
    g = load("assign_e1")
    for c in range(int(g["n_cases"])):
        p = torch.tensor(g[f"probs_{c}"])
        t, u = oassign.generate_dynamic_targets(p, target_ratio=float(g[f"ratio_{c}"]), w_uncertainty=True)
        assert np.array_equal(t.numpy(), g[f"targets_{c}"])
        assert np.array_equal(u.numpy(), g[f"unc_{c}"])


@pytest.mark.parametrize("tag", ["e3", "e4"])
@pytest.mark.parametrize("literal", [True, False])
@pytest.mark.parametrize("solver", ["c", "lsa"])
def test_assign_mc(tag, literal, solver):
    g = load("assign_" + tag)
    n_attr = 2 if tag == "e3" else 3
    fn = oassign.generate_dynamic_targets_gender_race if tag == "e3" else oassign.generate_dynamic_targets_gender_race_age
    emd = {"c": oemd.emd_c, "lsa": oemd.emd_lsa}[solver]
    for c in range(int(g["n_cases"])):
        probs = [torch.tensor(g[f"probs{k}_{c}"]) for k in range(n_attr)]
        world = int(g[f"world_{c}"])
        S = int(g[f"S_{c}"])
        if f"rand0_r0_{c}" in g:
            wr = [tuple(torch.tensor(g[f"rand{k}_r{r}_{c}"]) for k in range(n_attr)) for r in range(world)]
        else:
            wr = None
        outs = fn(*probs, w_uncertainty=True, num_samples_per_device=S, world_rand=wr, emd=emd, literal=literal)
        assert len(outs) == 2 * n_attr
        for k, o in enumerate(outs):
            ref = g[f"out{k}_{c}"]
            assert o.numpy().dtype == ref.dtype
            assert np.array_equal(o.numpy(), ref), (tag, c, k)


@pytest.mark.parametrize("tag", ["e1", "e3", "e4"])
def test_hooks_and_weights(tag):
    g = load("hooks")
    A = {"e1": 1, "e3": 2, "e4": 3}[tag]
    x = torch.tensor(g["images"], requires_grad=True)
    targets = [torch.tensor(g[f"{tag}_targets{a}"]) for a in range(A)]
    preds = [torch.tensor(g[f"{tag}_preds{a}"]) for a in range(A)]
    y = ohooks.apply_grad_hook_face(x, torch.tensor(g["box"]), torch.tensor(g["box_ori"]), targets, preds,
                                    list(g[f"{tag}_factors2"]), e1_rule=(tag == "e1"))
    assert np.array_equal(y.detach().numpy(), g[f"{tag}_forward"])
    (y * torch.tensor(g["upstream"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g[f"{tag}_grad"], rtol=1e-6, atol=0)
    face_ind = torch.tensor(~(g["box"] == -1).all(axis=1))
    w = ohooks.gen_dynamic_weights(face_ind, targets, preds, list(g[f"{tag}_factors1"]), torch.float32, e1_rule=(tag == "e1"))
    assert np.array_equal(w.numpy(), g[f"{tag}_weights"])


def test_oracle_header_says_test_infrastructure():
    assert "TEST INFRASTRUCTURE" in oracle.__doc__


# ----------------------------------------------------------------------------- 16-bit probability dtypes (half.npz)
HALF = ["f16", "bf16"]


@pytest.mark.parametrize("dn", HALF)
@pytest.mark.parametrize("tag,fn", [("e1", ohead.get_face_gender), ("e3", ohead.get_face_gender_race), ("e4", ohead.get_face_gender_race_age)])
def test_heads_half(dn, tag, fn):
    """fp16 is the reference's own classifier dtype (E1:933), bf16 the BASELINE headline dtype."""
    from tests._golden import half_equal, half_tensor
    g = load("half")
    sel = torch.tensor(g["heads_selector"])
    logits = half_tensor(g[f"heads_{dn}_{tag}_logits"], dn)
    outs = fn(lambda x: logits.clone(), torch.zeros(sel.shape[0], 1), selector=sel, fill_value=-1)
    for k, o in enumerate(outs):
        ref = g[f"heads_{dn}_{tag}_out{k}"]
        assert (np.array_equal(o.numpy(), ref) if o.dtype == torch.int64 else half_equal(o, ref)), (dn, tag, k)


@pytest.mark.parametrize("dn", HALF)
def test_assign_e1_half(dn):
    from tests._golden import half_equal, half_tensor
    g = load("half")
    for c in range(int(g[f"e1_{dn}_n_cases"])):
        p = half_tensor(g[f"e1_{dn}_probs_{c}"], dn)
        t, u = oassign.generate_dynamic_targets(p, target_ratio=float(g[f"e1_{dn}_ratio_{c}"]), w_uncertainty=True)
        assert np.array_equal(t.numpy(), g[f"e1_{dn}_targets_{c}"]) and half_equal(u, g[f"e1_{dn}_unc_{c}"])
        t_thr, _ = oassign.threshold_and_slice(t, u, 0.2, t.shape[0], 0)
        assert np.array_equal(t_thr.numpy(), g[f"e1_{dn}_thr_{c}"])


@pytest.mark.parametrize("dn", HALF)
@pytest.mark.parametrize("tag", ["e3", "e4"])
@pytest.mark.parametrize("literal", [True, False])
def test_assign_mc_half(dn, tag, literal):
    """Targets, uncertainties (bit patterns) and thresholded targets of the Monte-Carlo assignment on 16-bit
    probabilities, simulated world sizes up to 8 (fp16) / 4 (bf16)."""
    from tests._golden import half_equal, half_tensor
    g = load("half")
    n_attr = 2 if tag == "e3" else 3
    fn = oassign.generate_dynamic_targets_gender_race if tag == "e3" else oassign.generate_dynamic_targets_gender_race_age
    key = f"{tag}_{dn}"
    for c in range(int(g[f"{key}_n_cases"])):
        probs = [half_tensor(g[f"{key}_probs{k}_{c}"], dn) for k in range(n_attr)]
        world, S = int(g[f"{key}_world_{c}"]), int(g[f"{key}_S_{c}"])
        wr = [tuple(half_tensor(g[f"{key}_rand{k}_r{r}_{c}"], dn) for k in range(n_attr)) for r in range(world)]
        outs = fn(*probs, w_uncertainty=True, num_samples_per_device=S, world_rand=wr, literal=literal)
        for a in range(n_attr):
            assert np.array_equal(outs[2 * a].numpy(), g[f"{key}_out{2 * a}_{c}"]), (key, c, a)
            assert half_equal(outs[2 * a + 1], g[f"{key}_out{2 * a + 1}_{c}"]), (key, c, a)
            t_thr, _ = oassign.threshold_and_slice(outs[2 * a], outs[2 * a + 1], 0.2, outs[0].shape[0], 0)
            assert np.array_equal(t_thr.numpy(), g[f"{key}_thr{a}_{c}"]), (key, c, a)
