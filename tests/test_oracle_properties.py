"""Property tests (hypothesis) of the CPU oracle, the checker every GPU parity test leans on (test-plan
item 2): random face boxes incl. boxes hanging over the border, degenerate sizes, and random transport problems."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import assign as oassign, boxes as oboxes, crop as ocrop, emd as oemd, hooks as ohooks
from ._golden import peaked_probs, procedural_image

IMG = torch.tensor(procedural_image(3, 3, 64, 80))


@settings(max_examples=60, deadline=None)
@given(x0=st.integers(-40, 70), y0=st.integers(-40, 60), w=st.integers(1, 90), h=st.integers(1, 90), o=st.sampled_from([7, 16, 33]))
def test_crop_face_equals_direct_sampler_for_any_box(x0, y0, w, h, o):
    """slice + pad(-1) + bilinear resize (E1:267-290) == direct clamped-bilinear sampling of the virtual padded box, for any
    box that overlaps the image (a box fully outside makes the reference's Resize fail on an empty slice)."""
    x1, y1 = x0 + w, y0 + h
    if x1 <= 0 or y1 <= 0 or x0 >= 80 or y0 >= 64:
        return
    chip = ocrop.crop_face(IMG, [x0, y0, x1, y1], [o, o], -1)
    direct = ocrop.direct_sampler(IMG, [x0, y0, x1, y1], (o, o), -1.0)
    assert chip.shape == (3, o, o)
    assert (direct - chip.double()).abs().max() < 5e-5


@settings(max_examples=100, deadline=None)
@given(cx=st.floats(0, 512), cy=st.floats(0, 512), w=st.floats(1, 400), h=st.floats(1, 400))
def test_expand_bbox_is_a_square_about_the_same_centre(cx, cy, w, h):
    """E1:238-265 with expand_coef 0.5, ratio 1: side ~ 1.5 max(w,h) and the centre is kept, up to the integer rounding."""
    box = np.array([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], dtype=np.float32)
    x0, y0, x1, y1 = oboxes.expand_bbox(box, 0.5, 1)
    side = 1.5 * max(float(box[2] - box[0]), float(box[3] - box[1]))
    assert abs((x1 - x0) - side) <= 1.001 and abs((y1 - y0) - side) <= 1.001
    assert abs((x0 + x1) / 2 - float(box[0] + box[2]) / 2) <= 1.001 and abs((y0 + y1) / 2 - float(box[1] + box[3]) / 2) <= 1.001


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 40), k=st.sampled_from([8, 16]), seed=st.integers(0, 10_000))
def test_transport_plan_invariants_and_optimality(n, k, seed):
    """The oracle's exact solver: one class per row, class sizes = demand, and no cheaper plan from the independent
    assignment solver (scipy linear_sum_assignment)."""
    rng = np.random.default_rng(seed)
    pg, pr = peaked_probs(rng, n, 2, 2.0), peaked_probs(rng, n, 4, 2.0)
    pa = peaked_probs(rng, n, 2, 2.0) if k == 16 else None
    M = oemd.cost_matrix_c(pg, pr, pa)
    b = rng.multinomial(n, rng.dirichlet(np.ones(k)))
    a = oemd.assign_c(M, b)
    assert a.shape == (n,) and np.array_equal(np.bincount(a, minlength=k), b)
    T = oemd.emd_lsa(np.ones(n), b, M)
    assert abs((T * M).sum() - M[np.arange(n), a].sum()) < 1e-9


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 14), miss=st.integers(0, 3), seed=st.integers(0, 10_000))
def test_race_assignment_invariants(n, miss, seed):
    """E6: rows without a face keep -1; the others get a class in 0..3 and an uncertainty in [0, 0.75]."""
    rng = np.random.default_rng(seed)
    p = torch.tensor(peaked_probs(rng, n + miss, 4, 1.5))
    if miss:
        p[rng.permutation(n + miss)[:miss]] = -1
    t, u = oassign.generate_dynamic_targets_race(p, True)
    valid = (p != -1).all(-1)
    assert (t[~valid] == -1).all() and (u[~valid] == -1).all()
    assert ((t[valid] >= 0) & (t[valid] <= 3)).all()
    assert ((u[valid] >= -1e-6) & (u[valid] <= 0.75 + 1e-6)).all()


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), quirk=st.booleans())
def test_hook_region_is_the_intersection_of_both_boxes_and_the_image(seed, quirk):
    """apply_grad_hook_face (E1:1584-1617): the gradient is scaled exactly inside bbox ∩ bbox_ori ∩ image, including the
    python-slice quirk of an original box of -1 (E1:1594-1597)."""
    rng = np.random.default_rng(seed)
    H = W = 48
    box = [int(v) for v in (rng.integers(-10, 30), rng.integers(-10, 30), rng.integers(20, 60), rng.integers(20, 60))]
    ori = [-1, -1, -1, -1] if quirk else [int(v) for v in (rng.integers(-10, 30), rng.integers(-10, 30), rng.integers(20, 60), rng.integers(20, 60))]
    images = torch.zeros(1, 3, H, W, requires_grad=True)
    out = ohooks.apply_grad_hook_face(images, torch.tensor([box]), torch.tensor([ori]), [torch.tensor([1])], [torch.tensor([0])],
                                      [0.25], e1_rule=True)
    out.sum().backward()
    g = images.grad[0, 0]
    assert set(np.unique(g.numpy()).tolist()) <= {0.25, 1.0}
    if quirk:       # python slice ends of -1: the region ignores the new box's right / bottom edge and stops one short of the border
        left, right, bottom, top = max(box[0], 0), W - 1, max(box[1], 0), H - 1
    else:
        left, right = max(box[0], ori[0], 0), min(box[2], ori[2], W)
        bottom, top = max(box[1], ori[1], 0), min(box[3], ori[3], H)
    want = torch.zeros(H, W, dtype=torch.bool)
    if right > left and top > bottom:
        want[bottom:top, left:right] = True
    assert torch.equal(g == 0.25, want)
