"""GPU (-m gpu): the kernels and the whole step AT THE BASELINE SHAPES (512x512 images, 224x224 crops; C5 = 1024 bf16
images, K = 16), compared DIRECTLY with the oracle -- not with the repo's own generic kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def fg():
    import fairguide
    assert fairguide._lib.lib().fg_abi_version() == 1
    return fairguide


def _c5_like_inputs(n, dtype, seed):
    from fairguide import pipeline
    cfg = pipeline.GuidanceConfig(kind="gender_race_age")
    b = pipeline.synth_batch(n, cfg, dtype, "cpu", seed=seed, host=True)
    from oracle import boxes as oboxes
    ind, boxes = oboxes.select_and_expand(b["cand_boxes"].numpy(), b["counts"].numpy(), 512, 0.5, 1, -1)
    return cfg, b, torch.tensor(ind), torch.tensor(boxes)


@pytest.mark.parametrize("dtype,rtol", [(torch.bfloat16, 2e-2), (torch.float16, 5e-3), (torch.float32, 1e-3)])
def test_sampler_512_to_224_vs_oracle(fg, dtype, rtol):
    """sample_fwd_tiled_kernel<T,3,STAGES,512> (the instance bench.py runs) against oracle.crop at 512 -> 224, boxes of
    the C5 distribution plus boxes over every border, upscaled, tiny, missing."""
    from oracle import crop as ocrop
    n = 16
    cfg, b, ind, boxes = _c5_like_inputs(n, dtype, 77)
    boxes[0] = torch.tensor([-120, -90, 530, 560]); boxes[1] = torch.tensor([300, 40, 700, 440]); boxes[2] = torch.tensor([100, 120, 250, 270])
    boxes[3] = torch.tensor([200, 210, 260, 270]); ind[:4] = True; ind[4] = False; boxes[4] = -1
    x = b["images"].float()
    chips_ref = ocrop.crop_faces(x, boxes, ind, 224, -1)
    small_ref = ocrop.resize_small(x, 224)
    chips, small = fg.ops.crop_resize_fwd(b["images"].to(DEV), boxes.to(DEV), ind.to(DEV), (224, 224), (224, 224), -1.0)
    atol = 1e-5 if dtype == torch.float32 else rtol
    np.testing.assert_allclose(chips.float().cpu().numpy(), chips_ref.numpy(), rtol=rtol, atol=atol)
    np.testing.assert_allclose(small.float().cpu().numpy(), small_ref.numpy(), rtol=rtol, atol=atol)
    # chip-only and resize-only launches of the same kernel
    c2, _ = fg.ops.crop_resize_fwd(b["images"].to(DEV), boxes.to(DEV), ind.to(DEV), (224, 224), None, -1.0)
    _, s2 = fg.ops.crop_resize_fwd(b["images"].to(DEV), None, None, None, (224, 224), -1.0)
    assert torch.equal(c2, chips) and torch.equal(s2, small)


@pytest.mark.parametrize("dtype,rtol", [(torch.bfloat16, 2e-2), (torch.float16, 5e-3), (torch.float32, 1e-3)])
@pytest.mark.parametrize("nsub", ["0", "4", "8", "16", "q4", "q16"])
def test_image_grad_512_vs_oracle(fg, dtype, rtol, nsub, monkeypatch):
    """image_grad_staged_kernel<T,NSUB> (16-bit) / the fp32 backward at the BASELINE shape against autograd through
    oracle.crop + oracle.hooks (the reference's crop, hook and Resize), every rows-per-CTA variant."""
    from oracle import crop as ocrop, hooks as ohooks
    if nsub.startswith("q"):                    # the second-generation kernel for 16-bit gradients too (fp32 always uses it)
        monkeypatch.setenv("FG_BWD_QUAD", "1")
        nsub = nsub[1:]
    if nsub != "0":
        monkeypatch.setenv("FG_BWD_NSUB", nsub)
    n = 16
    cfg, b, ind, boxes = _c5_like_inputs(n, dtype, 78)
    boxes[0] = torch.tensor([-120, -90, 530, 560]); boxes[1] = torch.tensor([300, 40, 700, 440]); boxes[2] = torch.tensor([100, 120, 250, 270])
    boxes[3] = torch.tensor([10, 300, 130, 420]); ind[:4] = True; ind[4] = False; boxes[4] = -1
    rng = np.random.default_rng(3)
    bbox_ori = torch.where(ind[:, None], boxes + b["bbox_jitter"], boxes); bbox_ori[5] = -1       # the [-1]*4 original-box quirk
    targets = [torch.tensor(rng.integers(-1, w, size=n)) for w in (2, 4, 2)]
    preds = b["preds_ori"]
    f2 = [0.2, 0.3, 0.3]
    x = b["images"].float().clone().requires_grad_(True)
    chips_ref = ocrop.crop_faces(x, boxes, ind, 224, -1)
    hooked = ohooks.apply_grad_hook_face(x, boxes, bbox_ori, targets, preds, f2, False)
    small_ref = ocrop.resize_small(hooked, 224)
    gc, gs = b["g_chips"] * 1e3, b["g_small"] * 1e3          # O(1) gradients
    ((chips_ref * gc.float()).sum() + (small_ref * gs.float()).sum()).backward()
    region, scale, _ = fg.ops.guidance_factors(None, boxes.to(DEV), bbox_ori.to(DEV), [t.to(DEV) for t in targets], [p.to(DEV) for p in preds],
                                               f2, None, False, 512, 512, want_weights=False)
    g = fg.ops.image_grad(gc.to(DEV), gs.to(DEV), boxes.to(DEV), ind.to(DEV), region, scale, (n, 3, 512, 512), dtype, torch.device(DEV))
    gref = x.grad.numpy()
    np.testing.assert_allclose(g.float().cpu().numpy(), gref, rtol=rtol, atol=rtol * float(np.abs(gref).max()))
    # each branch alone
    g_c = fg.ops.image_grad(gc.to(DEV), None, boxes.to(DEV), ind.to(DEV), None, None, (n, 3, 512, 512), dtype, torch.device(DEV))
    x2 = b["images"].float().clone().requires_grad_(True)
    (ocrop.crop_faces(x2, boxes, ind, 224, -1) * gc.float()).sum().backward()
    np.testing.assert_allclose(g_c.float().cpu().numpy(), x2.grad.numpy(), rtol=rtol, atol=rtol * float(x2.grad.abs().max()))


@pytest.mark.parametrize("kind,n,dtype,S", [("gender_race_age", 1024, torch.bfloat16, 100),      # BASELINE C5
                                            ("gender_race_age", 1024, torch.float32, 100),       # C3
                                            ("gender_race", 512, torch.float32, 100),            # C2 (all ranks' rows)
                                            ("gender", 64, torch.float32, 100),                  # C1
                                            ("gender_race", 512, torch.float16, 100)])           # the reference's own dtype
def test_whole_step_at_baseline_shapes(fg, kind, n, dtype, S):
    from tests import stepcheck
    rep = stepcheck.check_step_vs_oracle(kind=kind, n=n, dtype=dtype, n_grad=32, S=S, seed=5991, device="cuda:0", captured=True)
    assert min(rep["rows_with_target"]) > 0.3 * n, rep          # the threshold keeps a good share of the rows


def test_whole_step_c4_three_candidate_boxes(fg):
    """C4: up to 3 candidate boxes per image (largest-area selection), K = 8, 4096 rows in the assignment."""
    from tests import stepcheck
    rep = stepcheck.check_step_vs_oracle(kind="gender_race", n=4096, dtype=torch.float32, n_grad=24, S=16, seed=6001, device="cuda:0",
                                         max_faces=3)
    assert min(rep["rows_with_target"]) > 0.3 * 4096, rep
