"""Next row f1: aligned 112x112 face chip (image_pipeline E1:292-312) and the face-realism loss
(E1:96-117, 1179-1190, 1917-1929).  skimage / kornia / sentence-transformers are not installed, so the oracle restates
their arithmetic (oracle/align.py); the CPU tests cross-check that restatement against OpenCV and an independent closed
form, and pin the reference's own glue code through the golden vectors of make_golden_next.py."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nextrows.npz"))
DEV = "cuda"


def smooth_images(n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    ph = torch.rand(n, 3, 1, 1, generator=g) * 6.28
    fr = 0.02 + 0.05 * torch.rand(n, 3, 1, 1, generator=g)
    img = 0.6 * torch.sin(fr * xx + ph) * torch.cos(fr * 0.7 * yy - ph) + (torch.rand(n, 3, H, W, generator=g) - 0.5) * 0.2
    return img.clamp_(-1, 1)


def landmarks_for(n, H, W, seed, scale_range=(0.8, 3.0), noise=1.5):
    """Five-point landmarks = the template under a random similarity (scale, rotation, centre) + detector noise; some
    faces hang over the image border, some are smaller than the 112 template (up-sampling)."""
    from oracle import align
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        s = rng.uniform(*scale_range)
        th = rng.uniform(-0.6, 0.6)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        c = rng.uniform([0.1 * W, 0.1 * H], [0.9 * W, 0.9 * H])
        out.append((align.TEMPLATE_112 - 56.0) @ (R.T * s) + c + rng.normal(size=(5, 2)) * noise)
    return np.asarray(out, dtype=np.float32)


# ----------------------------------------------------------------------------- CPU: the oracle's third-party restatements
def test_oracle_umeyama_vs_closed_form_and_opencv():
    import cv2
    from oracle import align
    lms = landmarks_for(6, 512, 512, 3)
    for lm in lms:
        T = align.umeyama(lm, align.TEMPLATE_112)
        src = lm.astype(np.float64)
        ms, md = src.mean(0), align.TEMPLATE_112.mean(0)
        sd, dd = src - ms, align.TEMPLATE_112 - md
        A = dd.T @ sd / 5
        a, b = A[0, 0] + A[1, 1], A[1, 0] - A[0, 1]
        Rm = np.array([[a, -b], [b, a]]) / ((sd ** 2).sum() / 5)
        np.testing.assert_allclose(T[:2, :2], Rm, rtol=1e-6)
        np.testing.assert_allclose(T[:2, 2], md - Rm @ ms, rtol=1e-6, atol=1e-4)
    # noise-free points: OpenCV's least-squares similarity agrees
    lm = landmarks_for(1, 512, 512, 4, noise=0.0)[0].astype(np.float64)
    T = align.umeyama(lm, align.TEMPLATE_112)
    Mcv, _ = cv2.estimateAffinePartial2D(lm, align.TEMPLATE_112, method=cv2.LMEDS)
    np.testing.assert_allclose(T[:2], Mcv, rtol=1e-5, atol=1e-4)


def test_oracle_warp_affine_vs_opencv():
    """kornia's warp with align_corners=True is the plain pixel-space affine warp OpenCV computes (fixed-point weights
    there: 1/32 px); the reference's align_corners=False call differs from it by a position-dependent sub-pixel shift."""
    import cv2
    from oracle import align
    img = smooth_images(1, 256, 256, 5)[0]
    lm = landmarks_for(1, 256, 256, 6, scale_range=(1.0, 1.6))[0]
    M = align.similarity_matrix(lm)
    out = align.warp_affine(img.unsqueeze(0), torch.tensor(M).unsqueeze(0).float(), (112, 112), align_corners=True)[0].numpy()
    ref = np.stack([cv2.warpAffine(c, M, (112, 112), flags=cv2.INTER_LINEAR) for c in img.numpy()])
    assert np.abs(out - ref).max() < 5e-3
    out2 = align.warp_affine(img.unsqueeze(0), torch.tensor(M).unsqueeze(0).float(), (112, 112), align_corners=False)[0].numpy()
    assert np.abs(out2 - ref).max() > np.abs(out - ref).max()


def test_oracle_image_pipeline_golden():
    """The reference's image_pipeline body (lifted with ast; skimage / kornia stubbed with the oracle's restatements)."""
    from oracle import align
    for c in range(int(GOLD["align_n_cases"])):
        img = torch.tensor(GOLD[f"align_img_{c}"])
        got = align.image_pipeline(img, GOLD[f"align_lm_{c}"])
        np.testing.assert_allclose(got.numpy(), GOLD[f"align_out_{c}"], rtol=0, atol=1e-6)


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_similarity_matrices_vs_oracle():
    import fairguide as fg
    from oracle import align
    lms = landmarks_for(64, 512, 512, 11)
    ind = torch.ones(64, dtype=torch.bool); ind[5] = False
    M = fg.similarity_matrices(torch.tensor(lms).to(DEV), ind.to(DEV)).cpu().numpy()
    for i in range(64):
        if not ind[i]:
            assert (M[i] == -1).all()
            continue
        want = align.similarity_matrix(lms[i]).astype(np.float32)       # .to(img.dtype), E1:307
        np.testing.assert_allclose(M[i], want, rtol=2e-6, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("H,W", [(512, 512), (200, 256)])
def test_gpu_aligned_chips_fwd_bwd_vs_oracle(dtype, tol, H, W):
    import fairguide as fg
    from oracle import align
    n = 10
    imgs = smooth_images(n, H, W, 21).to(dtype)
    lms = landmarks_for(n, H, W, 22, scale_range=(0.6, 3.0) if H == 512 else (0.5, 1.6))
    ind = torch.ones(n, dtype=torch.bool); ind[3] = False
    x_ref = imgs.float().clone().requires_grad_(True)
    ref = align.aligned_face_chips(x_ref, lms, ind)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    ref.backward(g)
    x = imgs.to(DEV).requires_grad_(True)
    out = fg.aligned_face_chips(x, torch.tensor(lms).to(DEV), ind.to(DEV))
    assert out.dtype == dtype and tuple(out.shape) == (n, 3, 112, 112)
    out.backward(g.to(DEV, dtype))
    np.testing.assert_allclose(out.detach().float().cpu().numpy(), ref.detach().numpy(), rtol=tol, atol=tol)
    gref = x_ref.grad.numpy()
    np.testing.assert_allclose(x.grad.float().cpu().numpy(), gref, rtol=tol, atol=tol * np.abs(gref).max())
    assert (out[3].detach() == -1).all() and (x.grad[3] == 0).all()
    # single-image entry point with the reference's signature
    one = fg.image_pipeline(imgs[0].to(DEV), lms[0])
    assert torch.equal(one, out[0].detach())


@pytest.mark.gpu
def test_gpu_aligned_warp_adjoint_and_accumulate_full_size():
    """At the BASELINE image size: <A x, g> == <x, A^T g> for the linear part of the warp (the -1 padding is its constant
    part), and accumulate mode == overwrite mode + what was there."""
    import fairguide as fg
    n = 64
    g0 = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(n, 3, 512, 512, generator=g0, device=DEV)
    lms = torch.tensor(landmarks_for(n, 512, 512, 31)).to(DEV)
    ind = torch.rand(n, generator=g0, device=DEV) > 0.1
    params = fg.ops.align_matrices(lms, ind, (512, 512))
    Ax = fg.ops.aligned_warp_fwd(x, params, ind) - fg.ops.aligned_warp_fwd(torch.zeros_like(x), params, ind)
    g = torch.randn(n, 3, 112, 112, generator=g0, device=DEV)
    Atg = fg.ops.aligned_warp_bwd(g, params, ind, tuple(x.shape))
    lhs, rhs = (Ax.double() * g.double()).sum().item(), (x.double() * Atg.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0), (lhs, rhs)
    base = torch.randn_like(x)
    acc = fg.ops.aligned_warp_bwd(g, params, ind, tuple(x.shape), g_images=base.clone())
    np.testing.assert_allclose(acc.cpu().numpy(), (base + Atg).cpu().numpy(), rtol=0, atol=1e-5)


# ----------------------------------------------------------------------------- face-realism loss
def _face_case(n, d, D, n_attr, seed):
    g = torch.Generator().manual_seed(seed)
    db = torch.nn.functional.normalize(torch.randn(D, d, generator=g), dim=-1)
    raw = torch.randn(n, d, generator=g) * 3.0
    feats_ori = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    face = torch.rand(n, generator=g) > 0.15
    widths = [2, 4, 2][:n_attr]
    targets = [torch.randint(-1, w, (n,), generator=g) for w in widths]
    preds = [torch.where(torch.rand(n, generator=g) < 0.9, t.clamp(min=0), torch.randint(0, w, (n,), generator=g)) for t, w in zip(targets, widths)]
    probs = [torch.softmax(torch.randn(n, w, generator=g) * 4.0, -1) for w in widths]
    return db, raw, feats_ori, face, targets, preds, probs


def test_oracle_semantic_search_is_top1():
    from oracle import align
    db, raw, _, face, *_ = _face_case(9, 64, 300, 1, 2)
    q = torch.nn.functional.normalize(raw, dim=-1)
    t, s = align.semantic_search(db, q, face, return_similarity=True)
    scores = q @ db.T
    for i in range(9):
        if face[i]:
            assert torch.equal(t[i], db[scores[i].argmax()]) and s[i] == scores[i].max()
        else:
            assert (t[i] == -1).all() and s[i] == -1


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,D", [(4, 512, 5000), (40, 512, 777), (70, 128, 300)])
def test_gpu_semantic_search_vs_oracle(n, d, D):
    import fairguide as fg
    from oracle import align
    db, raw, _, face, *_ = _face_case(n, d, D, 1, n)
    q = torch.nn.functional.normalize(raw, dim=-1)
    model = fg.FaceFeatsModel(db.to(DEV))
    t, s = model.semantic_search(q.to(DEV), face.to(DEV), return_similarity=True)
    rt, rs = align.semantic_search(db, q, face, return_similarity=True)
    np.testing.assert_allclose(s.cpu().numpy(), rs.numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(t.cpu().numpy(), rt.numpy(), rtol=0, atol=1e-6)          # same database rows
    assert torch.equal(model.semantic_search(q.to(DEV), face.to(DEV)), t)


@pytest.mark.gpu
@pytest.mark.parametrize("m,d,D", [(16, 512, 128), (33, 512, 1000), (128, 512, 20000), (200, 256, 4099), (64, 512, 100000)])
def test_gpu_tensor_core_search_equals_exact_search(m, d, D):
    """fg_face_search_top1_tc (one TF32 tcgen05 pass + exact re-score of the rows within the TF32 error bound) returns the
    SAME rows and the SAME fp32 scores as the streaming search -- with a selector, ragged last tiles, duplicated database
    rows (exact ties: the lowest row wins) and near-duplicates closer together than the TF32 resolution."""
    import fairguide as fg
    g = torch.Generator().manual_seed(m * 7 + D)
    db = torch.nn.functional.normalize(torch.randn(D, d, generator=g), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(m, d, generator=g), dim=-1)
    q[1] = db[D // 2]                                    # an exact hit ...
    db[D // 3] = db[D // 2]                              # ... that exists twice (tie -> row D // 3)
    db[D - 1] = torch.nn.functional.normalize(db[D // 2] + 1e-5 * torch.randn(d, generator=g), dim=-1)   # and a near-duplicate
    sel = torch.rand(m, generator=g) > 0.2
    sel[1] = True
    db, q, sel = db.to(DEV).contiguous(), q.to(DEV), sel.to(DEV)
    b0, s0 = fg.ops.face_search_top1(q, sel, db)
    b1, s1 = fg.ops.face_search_top1(q, sel, db, db_norm_bound=float(db.norm(dim=1).max()))
    assert torch.equal(b0, b1) and torch.equal(s0, s1)
    # the duplicated row: the lower index wins the tie (the near-duplicate may round to the same fp32 score or above)
    assert int(b1[1]) in (D // 3, D - 1) and (b1[~sel] == -1).all()
    ref = (q.double() @ db.double().T).max(1).values
    assert (ref[sel] - s1[sel].double()).abs().max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("n_attr", [1, 2, 3])
def test_gpu_face_realism_loss_fwd_bwd_vs_oracle(n_attr, dtype, tol):
    import fairguide as fg
    from oracle import align
    n, d, D = 37, 512, 2000
    db, raw, feats_ori, face, targets, preds, probs = _face_case(n, d, D, n_attr, 40 + n_attr)
    raw = raw.to(dtype)
    probs = [p.to(dtype) for p in probs]
    level = 0.75
    x_ref = raw.float().clone().requires_grad_(True)
    f_ref = torch.nn.functional.normalize(x_ref.to(torch.float), dim=-1)                 # get_face_feats tail, E1:1186-1189
    ref = align.face_loss(f_ref, db, face, targets, preds, probs, feats_ori, level)
    gl = torch.rand(n, generator=torch.Generator().manual_seed(5)) + 0.5
    (ref * gl).sum().backward()
    x = raw.to(DEV).requires_grad_(True)
    attr = []
    for a in range(n_attr):
        attr += [targets[a].to(DEV), preds[a].to(DEV), probs[a].to(DEV)]
    model = fg.FaceFeatsModel(db.to(DEV))
    loss = fg.face_realism_loss(x, feats_ori.to(DEV), model, face.to(DEV), *attr, confidence_level=level)
    assert loss.dtype == dtype
    (loss.float() * gl.to(DEV)).sum().backward()
    np.testing.assert_allclose(loss.detach().float().cpu().numpy(), ref.detach().numpy(), rtol=tol, atol=tol)
    gref = x_ref.grad.numpy()
    np.testing.assert_allclose(x.grad.float().cpu().numpy(), gref, rtol=tol, atol=tol * np.abs(gref).max())
    # the three kinds of rows all occur
    _, ws = fg.ops.face_loss_fwd(raw.to(DEV), feats_ori.to(DEV), model.face_feats.data, face.to(DEV), [t.to(DEV) for t in targets],
                                 [p.to(DEV) for p in preds], [p.to(DEV) for p in probs], level, n_attr == 1)
    rows = fg.ops.face_loss_target_rows(ws, n, d).cpu()
    from_ori = face.clone()
    for t, p_, q in zip(targets, preds, probs):
        from_ori &= (t != -1) & (t == p_) & (q.max(dim=-1).values >= level)
    assert torch.equal(rows == -2, from_ori) and ((rows == -1) == (ref.detach() == -1)).all()
    assert (rows == -1).any() and (rows >= 0).any() and (from_ori.any() or n_attr == 3)


@pytest.mark.gpu
def test_gpu_get_face_feats_vs_oracle():
    import fairguide as fg
    from oracle import align
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 5, stride=4), torch.nn.Flatten(), torch.nn.Linear(8 * 27 * 27, 512))
    data = torch.randn(6, 3, 112, 112)
    ref_in = data.clone().requires_grad_(True)
    ref = align.get_face_feats(net, ref_in)
    w = torch.randn(6, 512)
    (ref * w).sum().backward()
    net_d = net.to(DEV)
    x = data.to(DEV).requires_grad_(True)
    out = fg.get_face_feats(net_d, x)
    (out * w.to(DEV)).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-3, atol=1e-5)
    g = ref_in.grad.numpy()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g, rtol=1e-3, atol=1e-3 * np.abs(g).max())


# ----------------------------------------------------------------------------- oracle pinned to the reference's own statements
def test_oracle_face_db_and_search_golden():
    """FaceFeatsModel.__init__ / semantic_search (E1:82-117), executed from the reference source by make_golden_next.py."""
    from oracle import align
    db = torch.nn.functional.normalize(torch.tensor(GOLD["face_db_raw"]), dim=-1)                  # E1:88
    assert np.array_equal(db.numpy(), GOLD["face_db"])
    t, s = align.semantic_search(db, torch.tensor(GOLD["search_q"]), torch.tensor(GOLD["search_sel"]), return_similarity=True)
    assert np.array_equal(t.numpy(), GOLD["search_target"])
    np.testing.assert_allclose(s.numpy(), GOLD["search_sim"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("n_attr", [1, 2, 3])
def test_oracle_face_loss_golden(n_attr):
    """The loss_face_ij block of the training loop (E1:1917-1929 / E3:2124-2143 / E4:2253-2272), executed from the reference
    source: which rows take the original features, which the database search, which stay -1."""
    from oracle import align
    db = torch.tensor(GOLD["face_db"])
    f, fo, face = (torch.tensor(GOLD[f"face_{k}_{n_attr}"]) for k in ("feats", "feats_ori", "ind"))
    ts = [torch.tensor(GOLD[f"face_t{k}_{n_attr}"]) for k in range(n_attr)]
    ps = [torch.tensor(GOLD[f"face_p{k}_{n_attr}"]) for k in range(n_attr)]
    qs = [torch.tensor(GOLD[f"face_q{k}_{n_attr}"]) for k in range(n_attr)]
    got = align.face_loss(f, db, face, ts, ps, qs, fo, 0.75)
    want = GOLD[f"face_loss_{n_attr}"]
    assert np.array_equal(got.numpy() == -1, want == -1)
    np.testing.assert_allclose(got.numpy(), want, rtol=0, atol=1e-6)
    assert (want == -1).any() and (want != -1).any()


@pytest.mark.gpu
@pytest.mark.parametrize("n_attr", [1, 2, 3])
def test_gpu_face_loss_golden(n_attr):
    """The CUDA path against the same golden vectors (features already normalised: the kernel's normalisation is then the
    identity up to rounding)."""
    import fairguide as fg
    db = torch.tensor(GOLD["face_db"]).to(DEV)
    f, fo, face = (torch.tensor(GOLD[f"face_{k}_{n_attr}"]).to(DEV) for k in ("feats", "feats_ori", "ind"))
    attr = []
    for k in range(n_attr):
        attr += [torch.tensor(GOLD[f"face_t{k}_{n_attr}"]).to(DEV), torch.tensor(GOLD[f"face_p{k}_{n_attr}"]).to(DEV),
                 torch.tensor(GOLD[f"face_q{k}_{n_attr}"]).to(DEV)]
    loss = fg.face_realism_loss(f, fo, db, face, *attr, confidence_level=0.75)
    want = GOLD[f"face_loss_{n_attr}"]
    assert np.array_equal(loss.cpu().numpy() == -1, want == -1)
    np.testing.assert_allclose(loss.cpu().numpy(), want, rtol=1e-3, atol=1e-5)


def test_oracle_warp_equals_the_closed_form_map_the_kernel_uses():
    """kornia's normalise -> invert -> affine_grid -> grid_sample(align_corners=False) chain collapses to one affine map per
    image, `output pixel -> source pixel`:  u = (j+0.5)(Wd-1)/Wd, (x,y) = M^-1 (u,v,1), xs = x Ws/(Ws-1) - 0.5  (csrc/fg_align.cu).
    Checked here in numpy against the oracle's torch chain, including taps outside the image (zero padding)."""
    from oracle import align
    Hs, Ws, Hd, Wd = 96, 128, 112, 112
    img = smooth_images(1, Hs, Ws, 9)[0]
    for lm in landmarks_for(4, Hs, Ws, 10, scale_range=(0.4, 1.2)):
        M = align.similarity_matrix(lm).astype(np.float32).astype(np.float64)
        ref = align.warp_affine(img.unsqueeze(0), torch.tensor(M).unsqueeze(0).float(), (Hd, Wd), align_corners=False)[0].numpy()
        Minv = np.linalg.inv(np.vstack([M, [0, 0, 1]]))
        jj, ii = np.meshgrid(np.arange(Wd), np.arange(Hd))
        u, v = (jj + 0.5) * (Wd - 1) / Wd, (ii + 0.5) * (Hd - 1) / Hd
        xs = (Minv[0, 0] * u + Minv[0, 1] * v + Minv[0, 2]) * Ws / (Ws - 1) - 0.5
        ys = (Minv[1, 0] * u + Minv[1, 1] * v + Minv[1, 2]) * Hs / (Hs - 1) - 0.5
        x0, y0 = np.floor(xs).astype(int), np.floor(ys).astype(int)
        fx, fy = xs - x0, ys - y0

        def tap(c, y, x):
            ok = (x >= 0) & (x < Ws) & (y >= 0) & (y < Hs)
            return np.where(ok, c[np.clip(y, 0, Hs - 1), np.clip(x, 0, Ws - 1)], 0.0)

        mine = np.stack([(tap(c, y0, x0) * (1 - fx) + tap(c, y0, x0 + 1) * fx) * (1 - fy)
                         + (tap(c, y0 + 1, x0) * (1 - fx) + tap(c, y0 + 1, x0 + 1) * fx) * fy for c in img.numpy().astype(np.float64)])
        assert np.abs(mine - ref).max() < 2e-5
