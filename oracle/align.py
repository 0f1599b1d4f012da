"""ORACLE (test infrastructure): aligned 112x112 face chip and the face-realism loss.

Restates
    image_pipeline            E1:292-312     (img+1)/2*255 -> similarity warp to the 5-point template -> back to [-1,1]
    get_face_feats            E1:1179-1190   net(x) + net(flip(x)), float, L2 normalise          (the net is external)
    FaceFeatsModel.semantic_search  E1:96-117  top-1 dot-product search in the normalised face database
    face-realism loss         E1:1917-1929   1 - <f, target>, target from the original image or from the search

Third-party arithmetic that is NOT under /root/reference and NOT installed here (parity unpinned for it):
  * scikit-image 0.22.0 (environment.yml) ``transform.SimilarityTransform.estimate`` = ``_umeyama(src, dst, True)``:
    restated in ``umeyama`` from the published algorithm (Umeyama 1991, eq. 34-43), with numpy's SVD like skimage.
  * kornia 0.7.1 ``geometry.transform.warp_affine(..., align_corners=False)``: restated in ``warp_affine`` -- it is
    ``normalize_homography`` (pixel <-> [-1,1] with 2/(size-1), whatever the flag) + inverse + ``F.affine_grid`` +
    ``F.grid_sample`` with the flag passed through; torch's two functions are called directly as kornia does.
  * sentence-transformers ``util.semantic_search(q, corpus, score_function=dot_score, top_k=1)``: a dot-product matrix
    and torch.topk; restated as matmul + max.
Cross-checks in tests/test_align.py: ``umeyama`` against an independent closed-form 2-D solution and against
cv2.estimateAffinePartial2D on noise-free points; ``warp_affine(align_corners=True)`` against cv2.warpAffine.
"""
import numpy as np
import torch
import torch.nn.functional as F

# E1:296-302
TEMPLATE_112 = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655], [70.7299, 92.2041]])


def umeyama(src, dst, estimate_scale=True):
    """Least-squares similarity dst ~ s R src + t as a 3x3 matrix (float64), skimage 0.22 ``_umeyama`` semantics:
    means / variances in the inputs' own dtypes, covariance and SVD in float64."""
    src, dst = np.asarray(src), np.asarray(dst)
    num, dim = src.shape
    src_mean, dst_mean = src.mean(axis=0), dst.mean(axis=0)
    src_demean, dst_demean = src - src_mean, dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    if rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d) if estimate_scale else 1.0
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


def _normal_transform_pixel(height, width, dtype):
    # kornia 0.7.1 normal_transform_pixel: pixel -> [-1,1] with 2/(size-1)
    tr = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=dtype)
    tr[0, 0] = tr[0, 0] * 2.0 / (1e-14 if width == 1 else width - 1.0)
    tr[1, 1] = tr[1, 1] * 2.0 / (1e-14 if height == 1 else height - 1.0)
    return tr.unsqueeze(0)


def _inverse_cast(m):
    dt = m.dtype if m.dtype in (torch.float32, torch.float64) else torch.float32
    return torch.inverse(m.to(dt)).to(m.dtype)


def warp_affine(src, M, dsize, align_corners=False):
    """kornia 0.7.1 warp_affine(mode='bilinear', padding_mode='zeros'): src [B,C,H,W], M [B,2,3] (src pixel -> dst pixel)."""
    B, C, H, W = src.shape
    M3 = torch.cat([M, torch.tensor([[[0.0, 0.0, 1.0]]], dtype=M.dtype).expand(B, 1, 3)], dim=1)
    src_norm_trans_src_pix = _normal_transform_pixel(H, W, M.dtype)
    dst_norm_trans_dst_pix = _normal_transform_pixel(dsize[0], dsize[1], M.dtype)
    dst_norm_trans_src_norm = dst_norm_trans_dst_pix @ (M3 @ _inverse_cast(src_norm_trans_src_pix))
    src_norm_trans_dst_norm = _inverse_cast(dst_norm_trans_src_norm)
    grid = F.affine_grid(src_norm_trans_dst_norm[:, :2, :], [B, C, dsize[0], dsize[1]], align_corners=align_corners)
    return F.grid_sample(src, grid, align_corners=align_corners, mode="bilinear", padding_mode="zeros")


def similarity_matrix(landmarks):
    """tform.params[0:2] of E1:304-307 for one face: landmarks [5,2] -> float64 [2,3]."""
    return umeyama(np.asarray(landmarks), TEMPLATE_112, True)[0:2, :]


def image_pipeline(img, tgz_landmark):
    """E1:292-312.  img [3,H,W] in [-1,1] (gradient flows to it), tgz_landmark numpy [5,2] -> [3,112,112]."""
    img = (img + 1) / 2.0 * 255
    M = torch.tensor(similarity_matrix(tgz_landmark)).unsqueeze(dim=0).to(img.dtype)
    face = warp_affine(img.unsqueeze(dim=0), M, (112, 112), align_corners=False).squeeze()
    return (face / 255.0) * 2 - 1


def aligned_face_chips(images, landmarks, indicators, fill_value=-1):
    """The aligned-chip column of get_face_app (E1:1324-1351): per image, -1 filled without a face."""
    out = []
    for i in range(images.shape[0]):
        if bool(indicators[i]):
            out.append(image_pipeline(images[i], np.asarray(landmarks[i])).unsqueeze(0))
        else:
            out.append(torch.ones([1, 3, 112, 112], dtype=images.dtype) * fill_value)
    return torch.cat(out, dim=0)


# ----------------------------------------------------------------------------- face-realism loss
def get_face_feats(net, data, flip=True, normalize=True, to_high_precision=True):
    """E1:1179-1190."""
    feats = net(data)
    if flip:
        data = torch.flip(data, [3])
        feats = feats + net(data)
    if to_high_precision:
        feats = feats.to(torch.float)
    if normalize:
        feats = F.normalize(feats, dim=-1)
    return feats


def semantic_search(face_db, query_embeddings, selector, return_similarity=False):
    """FaceFeatsModel.semantic_search E1:96-117; ``face_db`` = the L2-normalised database [D,d] (E1:88)."""
    target = torch.ones_like(query_embeddings) * (-1)
    sims = torch.ones([query_embeddings.shape[0]], dtype=query_embeddings.dtype) * (-1)
    if selector.sum() > 0:
        scores = query_embeddings[selector] @ face_db.T            # util.dot_score
        best = scores.max(dim=1)                                   # top_k = 1
        target[selector] = face_db[best.indices]
        sims[selector] = best.values
    if return_similarity:
        return target.detach().clone(), sims
    return target.detach().clone()


def face_loss(face_feats, face_db, face_indicators, targets, preds_ori, probs_ori, face_feats_ori, confidence_level,
              search_needs_target=None, fill_value=-1.0):
    """E1:1917-1929 (one attribute) / E3:2124-2143 / E4:2253-2272 (lists of two / three) with the features already
    extracted for every row (the reference extracts them per index list): a row whose targets all equal the original
    predictions with confidence >= ``confidence_level`` takes ``1 - <f, f_ori>``; the other rows with a face (E1: and with
    a target) take ``1 - <f, nearest database entry>``; -1 elsewhere."""
    if torch.is_tensor(targets):
        targets, preds_ori, probs_ori = [targets], [preds_ori], [probs_ori]
    if search_needs_target is None:
        search_needs_target = len(targets) == 1
    n = face_feats.shape[0]
    loss = torch.ones(n, dtype=face_feats.dtype) * fill_value
    from_ori = (face_indicators == True)                                                        # noqa: E712
    for t, p, q in zip(targets, preds_ori, probs_ori):
        from_ori = from_ori * (t != -1) * (t == p) * (q.max(dim=-1).values >= confidence_level)
    idx1 = from_ori.nonzero().view([-1]).tolist()
    if len(idx1) > 0:
        loss[idx1] = 1 - (face_feats[idx1] * face_feats_ori[idx1]).sum(dim=-1)
    base = (face_indicators == True)                                                            # noqa: E712
    if search_needs_target:
        base = base * (targets[0] != -1)
    idx2 = sorted(set(base.nonzero().view([-1]).tolist()) - set(idx1))
    if len(idx2) > 0:
        f2 = face_feats[idx2]
        t2 = semantic_search(face_db, f2, face_indicators[idx2])
        loss[idx2] = 1 - (f2 * t2).sum(dim=-1)
    return loss
