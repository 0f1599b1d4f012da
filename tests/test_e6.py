"""exp-6 (race only): get_face_race, get_evaluate_metrics(probs_race_all) and the one-attribute race hook / weights with
positional factors, against golden vectors from the reference's own function bodies (tests/golden/make_golden_e6.py).
CPU: the oracle; GPU: the CUDA path through the public call surface."""
import numpy as np
import pytest
import torch

from tests._golden import half_equal, half_tensor, load

DEV = "cuda"


def _dec(arr, dn):
    return half_tensor(arr, dn) if arr.dtype == np.uint16 else torch.tensor(arr)


def _same(t, arr):
    return half_equal(t, arr) if arr.dtype == np.uint16 else np.array_equal(t.detach().cpu().numpy(), arr)


@pytest.mark.parametrize("dn", ["f32", "f16"])
def test_oracle_get_face_race(dn):
    from oracle import head as ohead
    g = load("e6")
    sel = torch.tensor(g["race_head_selector"])
    n = sel.shape[0]
    lg, full = _dec(g[f"race_head_{dn}_logits"], dn), _dec(g[f"race_head_{dn}_logits_full"], dn)
    chips = torch.zeros(n, 1, dtype=lg.dtype)
    for tag, outs in (("sel", ohead.get_face_race(lambda x: lg, chips, selector=sel)),
                      ("nosel", ohead.get_face_race(lambda x: full, chips, selector=None)),
                      ("empty", ohead.get_face_race(lambda x: 1 / 0, chips, selector=torch.zeros(n, dtype=torch.bool)))):
        for k in range(3):
            assert _same(outs[k], g[f"race_head_{dn}_{tag}{k}"]), (dn, tag, k)


@pytest.mark.parametrize("dn", ["f32", "bf16", "f16"])
def test_oracle_race_metrics(dn):
    from oracle import nextrows
    g = load("e6")
    got = nextrows.evaluate_metrics_race(_dec(g[f"race_metrics_{dn}_pr"], dn))
    np.testing.assert_allclose(np.array(got), g[f"race_metrics_{dn}_out"], rtol=0, atol=2e-7)


def test_oracle_race_hook_and_weights():
    from oracle import hooks as ohooks
    g, h = load("e6"), load("hooks")
    x = torch.tensor(h["images"], requires_grad=True)
    t, p = [torch.tensor(g["race_hook_targets"])], [torch.tensor(g["race_hook_preds"])]
    y = ohooks.apply_grad_hook_face(x, torch.tensor(h["box"]), torch.tensor(h["box_ori"]), t, p, [0.15], e1_rule=True)
    (y * torch.tensor(h["upstream"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g["race_hook_grad"], rtol=1e-6, atol=0)
    face = torch.tensor(~(h["box"] == -1).all(axis=1))
    w = ohooks.gen_dynamic_weights(face, t, p, [0.35], torch.float32, e1_rule=True)
    assert np.array_equal(w.numpy(), g["race_hook_weights"])


def test_positional_factor_parsing():
    """Host logic of apply_grad_hook_face / gen_dynamic_weights: factors may arrive positionally (E1:1584, E3:1751, E4:1823)."""
    import fairguide
    f = fairguide.api._split_factors
    names = fairguide.api._FACTOR_NAMES
    t = torch.zeros(2)
    args, A, v = f((t, t, t, 0.15), {}, names, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})
    assert A == 1 and v == [0.15] and len(args) == 3
    args, A, v = f((t,) * 6 + (0.5,), {"factor_race": 0.7}, names, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})
    assert A == 2 and v == [0.5, 0.7]
    args, A, v = f((t,) * 9 + (0.5, 0.6, 0.7), {}, names, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})
    assert A == 3 and v == [0.5, 0.6, 0.7]
    args, A, v = f((t,) * 9, {}, names, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})
    assert A == 3 and v == [0.2, 0.3, 0.3]
    with pytest.raises(TypeError):
        f((t,) * 6 + (0.5,), {"factor_gender": 0.7}, names, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})
    with pytest.raises(TypeError):
        f((t,) * 6, {"factor": 0.7}, names, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})


def test_expand_bbox_host_matches_golden():
    """api.expand_bbox is host arithmetic (like the reference's); it must round every golden box like the reference."""
    import fairguide
    g = load("boxes")
    for tag, coef, ratio in (("c05_r1", 0.5, 1), ("c11_r1", 1.1, 1), ("c05_r12", 0.5, 1.2)):
        got = np.array([fairguide.expand_bbox(b, coef, ratio) for b in g["boxes"]], dtype=np.int64)
        assert np.array_equal(got, g["expanded_" + tag])


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dn", ["f32", "f16"])
def test_gpu_get_face_race(dn):
    import fairguide as fg
    g = load("e6")
    sel = torch.tensor(g["race_head_selector"], device=DEV)
    n = sel.shape[0]
    lg, full = _dec(g[f"race_head_{dn}_logits"], dn), _dec(g[f"race_head_{dn}_logits_full"], dn)
    chips = torch.zeros(n, 1, dtype=lg.dtype, device=DEV)
    api = fg.bind(race_classifier=lambda x: lg.float().to(DEV))
    api_full = fg.bind(race_classifier=lambda x: full.float().to(DEV))
    for tag, outs in (("sel", api.get_face_race(chips, sel)), ("nosel", api_full.get_face_race(chips)),
                      ("empty", api.get_face_race(chips, torch.zeros(n, dtype=torch.bool, device=DEV)))):
        for k in range(3):
            ref = g[f"race_head_{dn}_{tag}{k}"]
            if k == 1 and tag != "empty":          # probabilities: device expf vs the host's
                r = _dec(ref, dn).float()
                assert tuple(outs[k].shape) == tuple(r.shape) and outs[k].dtype == lg.dtype
                np.testing.assert_allclose(outs[k].float().cpu().numpy(), r.numpy(), rtol=1e-5 if dn == "f32" else 1e-3, atol=1e-6)
            else:
                assert _same(outs[k], ref), (dn, tag, k)


@pytest.mark.gpu
@pytest.mark.parametrize("dn", ["f32", "bf16", "f16"])
def test_gpu_race_metrics(dn):
    import fairguide as fg
    g = load("e6")
    got = fg.get_evaluate_metrics(_dec(g[f"race_metrics_{dn}_pr"], dn).to(DEV))
    assert len(got) == 6
    np.testing.assert_allclose(np.array(got), g[f"race_metrics_{dn}_out"], rtol=0, atol=2e-7)


@pytest.mark.gpu
def test_gpu_race_hook_and_weights_positional_factor():
    import fairguide as fg
    g, h = load("e6"), load("hooks")
    x = torch.tensor(h["images"], device=DEV, requires_grad=True)
    t, p = torch.tensor(g["race_hook_targets"], device=DEV), torch.tensor(g["race_hook_preds"], device=DEV)
    q = torch.full((x.shape[0], 4), 0.25, device=DEV)
    y = fg.apply_grad_hook_face(x, torch.tensor(h["box"], device=DEV), torch.tensor(h["box_ori"], device=DEV), t, p, q, 0.15)
    (y * torch.tensor(h["upstream"], device=DEV)).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["race_hook_grad"], rtol=1e-6, atol=0)
    face = torch.tensor(~(h["box"] == -1).all(axis=1), device=DEV)
    w = fg.gen_dynamic_weights(face, t, p, q, 0.35)
    assert np.array_equal(w.cpu().numpy(), g["race_hook_weights"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gpu_head_attributes_backward_kernel(dtype):
    """fg_head_attributes_bwd (one launch) against autograd through torch's slice / softmax on the same values, with a
    selector, gradients through BOTH the probabilities and the sliced logits, and attributes without any gradient."""
    import fairguide as fg
    torch.manual_seed(4)
    n = 37
    sel = torch.rand(n) > 0.3
    m = int(sel.sum())
    lg = torch.randn(m, 8, requires_grad=True)
    gp = [torch.randn(n, w) for w in (2, 4, 2)]
    gl = [torch.randn(n, w) for w in (2, 4, 2)]
    ref = 0
    for a, (c, w) in enumerate(((0, 2), (2, 4), (6, 2))):
        if a == 2:
            continue                                              # no gradient through the age outputs
        sl = lg[:, c:c + w].to(dtype).float() if dtype != torch.float32 else lg[:, c:c + w]
        ref = ref + (torch.softmax(sl, -1) * gp[a][sel].to(dtype).float()).sum() + (sl * gl[a][sel].to(dtype).float()).sum()
    ref.backward()
    lgd = lg.detach().to(DEV).requires_grad_(True)
    outs = fg.api._heads("gender_race_age", lambda x: lgd, torch.zeros(n, 1, device=DEV, dtype=dtype), sel.to(DEV), -1)
    loss = 0
    for a in range(2):
        loss = loss + (outs[3 * a + 1].float() * gp[a].to(DEV).to(dtype).float()).sum() + (outs[3 * a + 2].float() * gl[a].to(DEV).to(dtype).float()).sum()
    loss.backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    np.testing.assert_allclose(lgd.grad.cpu().numpy(), lg.grad.numpy(), rtol=tol, atol=tol)
    assert float(lgd.grad[:, 6:].abs().max()) == 0.0
