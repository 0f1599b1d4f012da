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


def test_resize_512_to_224_tap_pattern_is_static():
    """The 512 -> 224 resize is the fixed ratio 16 : 7.  With ATen's fp32 index formula (area_pixel_compute_source_index,
    align_corners=False) resized index o touches image indices i0(o) = floor((32 o + 9) / 14) and i0(o) + 1; inverted: image
    index p is a tap of exactly ONE resized index, o = 7 (p >> 4) + ((7 (p & 15) + 3) >> 4), and p & 15 in {4, 11} is a tap of
    none.  The backward kernels (fg_image_grad_quad.cuh, small_tab_512_224 in fg_image_grad_staged.cuh) build their tap tables
    from this pattern instead of searching; the weights still come from the formula."""
    f = np.float32
    ss = f(512) / f(224)

    def taps(o):
        src = f(f(ss * f(f(o) + f(0.5))) - f(0.5))
        src = max(src, f(0))
        i0 = min(int(src), 511)
        return i0, i0 + (1 if i0 < 511 else 0)

    owner = {}
    for o in range(224):
        i0, i1 = taps(o)
        assert i0 == (32 * o + 9) // 14 and i1 == i0 + 1
        for p in (i0, i1):
            assert p not in owner                      # no image index receives two resized indices (shrink by more than 2x)
            owner[p] = o
    for p in range(512):
        q = p & 15
        if q in (4, 11):
            assert p not in owner
        else:
            assert owner[p] == 7 * (p >> 4) + ((7 * q + 3) >> 4)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 10 ** 6))
def test_trust_region_rule_of_the_price_search(seed):
    """The rule by which ot_solve_kernel's price search (fg_assign.cu, price_search_hybrid) stops sweeping a row: with pivot
    prices p0 and per-class radii R, let z = cost - (p0 + R) and s = argmin z; the row is FROZEN when
    z_s + 2 R_s < min_{l != s} z_l.  A frozen row takes class s under EVERY price vector of the box |p - p0| <= R, so counting
    it once is exact.  (Integer keys like the kernel's, so that comparisons are exact.)"""
    rng = np.random.default_rng(seed)
    n, k = 400, 16
    cost = rng.integers(0, 1 << 22, size=(n, k)).astype(np.int64) * 16 + np.arange(k)      # class index in the low bits: no ties
    p0 = rng.integers(-(1 << 18), 1 << 18, size=k).astype(np.int64) * 16
    R = rng.integers(1, 1 << 17, size=k).astype(np.int64) * 16
    z = cost - (p0 + R)
    s = z.argmin(1)
    zs = np.sort(z, axis=1)
    frozen = zs[:, 0] + 2 * R[s] + 16 < zs[:, 1]
    assert frozen.any() and (~frozen).any()
    for _ in range(20):
        corner = rng.random() < 0.5
        d = np.where(rng.random(k) < 0.5, -R, R) if corner else (rng.integers(-(1 << 17), 1 << 17, size=k) * 16).clip(-R, R)
        assign = (cost - (p0 + d)).argmin(1)
        assert (assign[frozen] == s[frozen]).all()


@settings(max_examples=15, deadline=None)
@given(st.integers(1, 10 ** 6), st.sampled_from([1.0, 3.0]))
def test_tf32_candidate_rule_of_the_tensor_core_search_keeps_the_exact_argmax(seed, scale):
    """fg_face_search_top1_tc scores every database row once with TF32 products (both operands truncated to 10 mantissa bits)
    and re-scores exactly the rows whose approximate score is within 2 eps of the query's best, eps = 2.2e-3 |q| max|row|.
    The exact arg-max -- and any row tied with it -- must always be among them."""
    rng = np.random.default_rng(seed)
    D, d, m = 3000, 256, 8
    db = rng.standard_normal((D, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    db *= scale
    db[7] = db[1500]                                            # an exact tie
    q = rng.standard_normal((m, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[0] = db[1500] / scale                                     # the tied rows are this query's best
    trunc = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    approx = trunc(q.copy()) @ trunc(db.copy()).T
    exact = q.astype(np.float64) @ db.astype(np.float64).T
    eps = 2.2e-3 * np.linalg.norm(q, axis=1) * np.linalg.norm(db, axis=1).max()
    keep = approx >= approx.max(1, keepdims=True) - 2 * eps[:, None]
    best = exact.max(1, keepdims=True)
    assert (keep | (exact < best - 1e-12)).all()                # every exact maximiser is kept
    assert keep[0, 7] and keep[0, 1500]
    assert keep.sum(1).max() < 200                              # and the list stays short
