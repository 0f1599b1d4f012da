"""Python call surface of the guidance path, mirroring exp-*/1-main-debias.py.

Same names, argument order, return tuples and -1 sentinels as the reference closures
(SURVEY.md section 8b).  The reference's closures capture ``gender_classifier`` /
``accelerator`` from ``main``; here they are explicit keyword arguments, and ``bind()`` builds
closures with the reference's exact signatures.  Every function below runs hand-written sm_100a
kernels through the C ABI (include/fairguide.h); there is no CPU or eager-PyTorch fallback.

E1 = exp-1-debias-gender/1-main-debias.py, E3 = exp-3-debias-gender-race/1-main-debias.py,
E4 = exp-4-debias-gender-race-age/1-main-debias.py.
"""
import functools
import math
import types

import numpy as np

import torch

from . import autograd as ag
from . import dist as fdist
from . import ops

_HEADS = {
    # kind: (col_start, width)
    "gender": ([40], [2]),                    # E1:1370  logits.view(B,-1,2)[:,20,:]
    "gender_race": ([0, 2], [2, 4]),          # E3:1404-1407
    "gender_race_age": ([0, 2, 6], [2, 4, 2]),  # E4:1398-1403
    "race": ([2], [4]),                       # exp-6-debias-race/1-main-debias.py:1381  logits[:,2:] of the 6-way head
}


# ----------------------------------------------------------------------------- boxes / crop
def expand_bbox(bbox, expand_coef, target_ratio):
    """E1:238-265.  ``bbox`` = 4 numbers (numpy float32 scalars in the reference) -> list of 4 ints.

    Four numbers in, four out, called once per detected face from a host loop (E1:1335): this is host work in the
    reference and stays host work here (a kernel launch plus a blocking read-back costs ~50x the ten flops).  The
    arithmetic is the one select_expand_kernel (csrc/fg_boxes.cu, the batched entry ``select_and_expand``) performs, spelled
    with explicit float32 / float64 casts: width, height and their ratio are float32 (numpy float32 scalar op scalar),
    everything that touches ``expand_coef`` / ``target_ratio`` / 0.5 (Python numbers) is float64 under the reference's
    pinned NumPy 1.26, and ``round`` is half-to-even."""
    f32, f64 = np.float32, np.float64
    b = [f32(v) for v in bbox]
    w, h = f32(b[2] - b[0]), f32(b[3] - b[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = f32(h / w)
    if f64(ratio) > f64(target_ratio):
        extra_h = f64(h) * f64(expand_coef)
        extra_w = (f64(h) + extra_h) / f64(target_ratio) - f64(w)
    else:
        extra_w = f64(w) * f64(expand_coef)
        extra_h = (f64(w) + extra_w) * f64(target_ratio) - f64(h)
    hw, hh = extra_w * f64(0.5), extra_h * f64(0.5)
    return [int(np.rint(f64(b[0]) - hw)), int(np.rint(f64(b[1]) - hh)), int(np.rint(f64(b[2]) + hw)), int(np.rint(f64(b[3]) + hh))]


def select_and_expand(boxes, counts, dim_max, expand_coef=0.5, target_ratio=1, fill_value=-1):
    """Batched get_largest_face_app (E1:1292-1304) + expand_bbox for [n,F,4] candidate boxes.
    -> (face_indicators bool [n], face_bboxs int64 [n,4])."""
    return ops.select_expand_boxes(boxes, counts, dim_max, expand_coef, target_ratio, fill_value)


def crop_face(img_tensor, bbox_new, target_size, fill_value):
    """E1:267-290.  ``img_tensor`` [3,H,W] (may require grad) -> [3,h,w]."""
    box = torch.as_tensor([list(bbox_new)], dtype=torch.int64, device=img_tensor.device)
    chips = ag.CropResize.apply(img_tensor.unsqueeze(0), box, None, None, None,
                                (int(target_size[0]), int(target_size[1])), None, float(fill_value))
    return chips[0]


def crop_faces(images, face_bboxs, face_indicators, size_face=224, fill_value=-1):
    """The crop part of get_face_app's per-image loop (E1:1324-1345) in one launch:
    chips [n,3,size,size]; rows without a face are all ``fill_value`` (E1:1328-1332)."""
    return ag.CropResize.apply(images, face_bboxs, face_indicators, None, None, (size_face, size_face), None, float(fill_value))


def resize_small(images, img_size_small=224):
    """transforms.Resize(img_size_small) on a square batch (E1:1860, E1:1905)."""
    return ag.CropResize.apply(images, None, None, None, None, None, (img_size_small, img_size_small), -1.0)


def crop_and_resize(images, face_bboxs, face_indicators, region=None, scale=None, size_face=224, img_size_small=224,
                    fill_value=-1):
    """Fused entry point the reference lacks: (face_chips, images_small) from one pass over the
    images; in the backward the gradient arriving through ``images_small`` is scaled by ``scale``
    inside ``region`` (= apply_grad_hook_face placed before the resize, E3:2106-2107) while the
    gradient arriving through ``face_chips`` is not (crops are taken before the hook, E3:2103)."""
    return ag.CropResize.apply(images, face_bboxs, face_indicators, region, scale, (size_face, size_face),
                               (img_size_small, img_size_small), float(fill_value))


# ----------------------------------------------------------------------------- classifier heads
def _split_classifier(classifier):
    """torchvision MobileNetV3: (backbone callable -> pooled [m,960], W1, b1, W2, b2); else None."""
    feats = getattr(classifier, "features", None)
    cls = getattr(classifier, "classifier", None)
    if feats is None or cls is None or len(cls) != 4:
        return None
    l1, l2 = cls[0], cls[3]
    if not isinstance(l1, torch.nn.Linear) or not isinstance(l2, torch.nn.Linear):
        return None
    if classifier.training:
        raise RuntimeError("fairguide head implements the eval-mode classifier (Dropout = identity)")

    def backbone(x):
        return torch.flatten(classifier.avgpool(feats(x)), 1)

    return backbone, l1.weight, l1.bias, l2.weight, l2.bias


def classifier_logits(classifier, chips):
    """logits [m,k_head] float32 with autograd back to ``chips``: backbone in torch/cuDNN, the dense
    head (classifier[0..3]) in our kernel.  A classifier without the MobileNetV3 layout is treated
    as an opaque logits producer."""
    parts = _split_classifier(classifier)
    if parts is None:
        return classifier(chips).float()
    backbone, w1, b1, w2, b2 = parts
    return ag.Head.apply(backbone(chips), w1, b1, w2, b2)


class _HeadAttributes(torch.autograd.Function):
    """slice -> softmax -> argmax -> scatter(fill) for every attribute, one launch."""

    @staticmethod
    def forward(ctx, logits, src_row, selector, n, kind, fill, dtype):
        cs, ws = _HEADS[kind]
        preds, probs, louts, flat = ops.head_attributes(logits, src_row, selector, n, cs, ws, fill, dtype, return_flat=True)
        ctx.kind, ctx.shape, ctx.n = kind, tuple(logits.shape), n
        ctx.save_for_backward(src_row, selector, flat)
        out = []
        for a in range(len(ws)):
            out += [preds[a], probs[a], louts[a]]
        ctx.mark_non_differentiable(*[preds[a] for a in range(len(ws))])
        return tuple(out)

    @staticmethod
    def backward(ctx, *grads):
        # one launch (fg_head_attributes_bwd): sliced-logit gradients plus the softmax backward of the probability gradients
        cs, ws = _HEADS[ctx.kind]
        src_row, selector, flat = ctx.saved_tensors
        A = len(ws)
        g_full = ops.head_attributes_bwd(flat, [grads[3 * a + 1] for a in range(A)], [grads[3 * a + 2] for a in range(A)],
                                         src_row, selector, ctx.n, ctx.shape[0], ctx.shape[1], cs, ws)
        return g_full, None, None, None, None, None, None


def _heads(kind, classifier, face_chips, selector, fill_value):
    n_attr = len(_HEADS[kind][1])
    chips = face_chips[selector] if selector is not None else face_chips
    n = selector.shape[0] if selector is not None else face_chips.shape[0]
    if chips.shape[0] == 0:
        logits = torch.zeros((0, _HEADS[kind][0][-1] + _HEADS[kind][1][-1]), dtype=torch.float32, device=face_chips.device)
    else:
        logits = classifier_logits(classifier, chips)
    if selector is not None:
        src_row = (torch.cumsum(selector.to(torch.int32), 0, dtype=torch.int32) - 1)
        out = _HeadAttributes.apply(logits, src_row, selector, n, kind, float(fill_value), face_chips.dtype)
    else:
        out = _HeadAttributes.apply(logits, None, None, n, kind, float(fill_value), face_chips.dtype)
    if kind == "gender_race_age" and selector is None:
        return out[:6]          # the reference returns only gender+race here (E4:1475); kept
    return out[:3 * n_attr]


def get_face_gender(face_chips, selector=None, fill_value=-1, *, gender_classifier):
    """E1:1355-1401 -> (preds_gender, probs_gender, logits_gender)."""
    return _heads("gender", gender_classifier, face_chips, selector, fill_value)


def get_face_gender_race(face_chips, selector=None, fill_value=-1, *, gender_race_classifier):
    """E3:1387-1457 -> (preds_g, probs_g, logits_g, preds_r, probs_r, logits_r)."""
    return _heads("gender_race", gender_race_classifier, face_chips, selector, fill_value)


def get_face_gender_race_age(face_chips, selector=None, fill_value=-1, *, gender_race_age_classifier):
    """E4:1378-1475 -> 9-tuple (6-tuple when selector is None, like the reference)."""
    return _heads("gender_race_age", gender_race_age_classifier, face_chips, selector, fill_value)


def get_face_race(face_chips, selector=None, fill_value=-1, *, race_classifier):
    """exp-6-debias-race/1-main-debias.py:1365-1411 -> (preds_race, probs_race, logits_race): columns 2..5 of the 6-way
    gender+race head."""
    return _heads("race", race_classifier, face_chips, selector, fill_value)


# ----------------------------------------------------------------------------- assignment
@torch.no_grad()
def generate_dynamic_targets(probs, target_ratio=0.5, w_uncertainty=False, *, uncertainty_threshold=None):
    """E1:1403-1447.  ``uncertainty_threshold`` (not in the reference signature) fuses E1:1835."""
    thr = -1.0 if uncertainty_threshold is None else float(uncertainty_threshold)
    t, u = ops.assign_rank_binom(probs, target_ratio, thr, w_uncertainty)
    return (t, u) if w_uncertainty else t


@torch.no_grad()
def _mc_targets(probs, w_uncertainty, num_samples_per_device, rand_tensors, num_valid, uncertainty_threshold, group,
                return_counts=False, check_status=True):
    pg, pr = probs[0], probs[1]
    pa = probs[2] if len(probs) == 3 else None
    K = 16 if pa is not None else 8
    n_all = pg.shape[0]
    if num_valid is None:
        # the reference has the same device->host read: ``if idxs_2_rank.sum() == 0`` (E3:1476)
        num_valid = int(((pg != -1).all(dim=-1) * (pr != -1).all(dim=-1)).sum().item())
    S = num_samples_per_device
    ws = ops.OtWorkspace(n_all, K, S, pg.device)
    if num_valid > 0 and rand_tensors is None:
        # same call order, shape, dtype and device as E3:1491-1492 / E4:1503-1505
        rand_tensors = tuple(torch.rand([S, num_valid], dtype=pg.dtype, device=pg.device) for _ in range(len(probs)))
    counts = ops.ot_plan_counts(pg, pr, pa, rand_tensors or (), num_valid, ws)
    if num_valid > 0:
        fdist.all_reduce_counts(counts, group)            # E3:1535, on exact int32 counts
    thr = -1.0 if uncertainty_threshold is None else float(uncertainty_threshold)
    ts, us = ops.ot_targets(counts, pg, pr, num_valid, ws, thr, w_uncertainty)
    if check_status:
        # one 16-byte read: a wrong ``num_valid`` or an inexact plan raises instead of returning wrong targets (the
        # reference's caller synchronises right after this call anyway: E3:2026 prints ``.sum().item()``)
        ws.check(num_valid)
    out = []
    for a in range(len(ts)):
        out.append(ts[a])
        if w_uncertainty:
            out.append(us[a])
    if return_counts:
        return tuple(out), counts, ws
    return tuple(out)


def generate_dynamic_targets_gender_race(probs_gender, probs_race, w_uncertainty=False, num_samples_per_device=100, *,
                                         rand_tensors=None, num_valid=None, uncertainty_threshold=None, group=None):
    """E3:1459-1569 -> (t_g, u_g, t_r, u_r) or (t_g, t_r).  Keyword-only extras: ``rand_tensors`` =
    the (gender, race) uniform draws [S,N] (default: torch.rand in the reference's order);
    ``num_valid`` = N if the caller already knows it (skips one device->host read);
    ``uncertainty_threshold`` fuses E3:2022-2023; ``group`` = process group of the all-reduce."""
    return _mc_targets((probs_gender, probs_race), w_uncertainty, num_samples_per_device, rand_tensors, num_valid,
                       uncertainty_threshold, group)


def generate_dynamic_targets_gender_race_age(probs_gender, probs_race, probs_age, w_uncertainty=False,
                                             num_samples_per_device=100, *, rand_tensors=None, num_valid=None,
                                             uncertainty_threshold=None, group=None):
    """E4:1477-1615 -> (t_g, u_g, t_r, u_r, t_a, u_a) or targets only."""
    return _mc_targets((probs_gender, probs_race, probs_age), w_uncertainty, num_samples_per_device, rand_tensors,
                       num_valid, uncertainty_threshold, group)


@functools.lru_cache(maxsize=64)
def race_compositions(N, mass=0.95):
    """Host side of E6:1438-1459 (host code in the reference as well; depends only on N): every composition
    (n1,n2,n3,n4) of N with its multinomial coefficient, normalised to sum 1, sorted by decreasing weight with the same
    numpy calls as the reference (so equal weights order the same way under the same numpy), cut after the cumulative
    weight first exceeds ``mass``.  -> (combs int64 [S,4], weights float64 [S])."""
    combs, coefs = [], []
    for n1 in range(N + 1):
        c1 = math.comb(N, n1)
        for n2 in range(N - n1 + 1):
            c12 = c1 * math.comb(N - n1, n2)
            for n3 in range(N - n1 - n2 + 1):
                combs.append([n1, n2, n3, N - n1 - n2 - n3])
                coefs.append(c12 * math.comb(N - n1 - n2, n3))
    combs, w = np.array(combs), np.array(coefs)
    w = w / np.linalg.norm(w, ord=1)
    order = np.flip(w.argsort())
    acc, keep = 0, len(order)
    for k, j in enumerate(order):
        acc += w[j]
        if acc > mass:
            keep = k + 1
            break
    order = order[:keep]
    return combs[order].astype(np.int64), w[order].astype(np.float64)


def generate_dynamic_targets_race(probs, w_uncertainty=False, *, num_valid=None, uncertainty_threshold=None):
    """exp-6-debias-race/1-main-debias.py:1413-1482 -> targets_all or (targets_all, uncertainty_all).
    ``num_valid`` = number of rows with a face if the caller knows it (the reference reads it from the device as well,
    E6:1424); ``uncertainty_threshold`` fuses the caller's ``targets[uncertainty > thr] = -1``."""
    if num_valid is None:
        num_valid = int((probs != -1).all(dim=-1).sum().item())
    demands = weights = None
    if num_valid > 0:
        combs, w = race_compositions(num_valid)
        d16 = np.zeros((combs.shape[0], 16), dtype=np.int32)
        d16[:, :4] = combs
        demands = torch.from_numpy(d16).to(probs.device)
        weights = torch.from_numpy(w).to(probs.device)
    thr = -1.0 if uncertainty_threshold is None else float(uncertainty_threshold)
    targets, unc, ws = ops.assign_race_enumerated(probs, num_valid, demands, weights, thr, w_uncertainty)
    if probs.shape[0] > 0:
        ops.check_ot_status(ws[:16].view(torch.int32).tolist(), num_valid)
    return (targets, unc) if w_uncertainty else targets


def threshold_and_slice(targets_all, uncertainty_all, uncertainty_threshold, n_local, local_process_index):
    """E3:2022-2025: ``targets_all[uncertainty_all > thr] = -1`` (in place) and this rank's rows."""
    targets_all[uncertainty_all > uncertainty_threshold] = -1
    return targets_all[n_local * local_process_index:n_local * (local_process_index + 1)]


# ----------------------------------------------------------------------------- f1: aligned chip
def similarity_matrices(face_landmarks, face_indicators=None, image_hw=(512, 512)):
    """``tform.params[0:2]`` of E1:304-307 for a batch of five-point landmarks [n,5,2] -> float64 [n,2,3]."""
    lm = torch.as_tensor(face_landmarks)
    return ops.align_matrices(lm, face_indicators, image_hw, want_matrix=True)[1]


def aligned_face_chips(images, face_landmarks, face_indicators=None, size_aligned_face=112, fill_value=-1):
    """The aligned-chip column of get_face_app (E1:1324-1351) in one call: image_pipeline (E1:292-312) for every image
    with a face, ``fill_value`` elsewhere.  images [n,3,H,W] (gradient flows to them), face_landmarks [n,5,2]."""
    lm = torch.as_tensor(face_landmarks, device=images.device)
    params = ops.align_matrices(lm, face_indicators, images.shape[-2:], (size_aligned_face,) * 2)
    return ag.AlignedWarp.apply(images, params, face_indicators, (size_aligned_face,) * 2, float(fill_value))


def image_pipeline(img, tgz_landmark):
    """E1:292-312: one image [3,H,W] in [-1,1] and its landmarks (numpy or tensor [5,2]) -> aligned chip [3,112,112]."""
    lm = torch.as_tensor(np.asarray(tgz_landmark) if not torch.is_tensor(tgz_landmark) else tgz_landmark)
    return aligned_face_chips(img.unsqueeze(0), lm.to(img.device).unsqueeze(0)).squeeze(0)


def get_face_feats(net, data, flip=True, normalize=True, to_high_precision=True):
    """E1:1179-1190.  ``net`` (SFNet in the reference) is an external callable; the float cast + L2 normalisation run in
    one kernel (with its backward)."""
    feats = net(data)
    if flip:
        feats = feats + net(torch.flip(data, [3]))
    if normalize:
        return ag.FeatsNormalize.apply(feats)          # computes in and returns float32, like to_high_precision=True
    return feats.to(torch.float) if to_high_precision else feats


class FaceFeatsModel(torch.nn.Module):
    """E1:82-117: the face database as a frozen, L2-normalised parameter and its top-1 dot-product search.  Built from
    the feature matrix [D,d] (the reference unpickles it from ``face_feats_path``)."""

    def __init__(self, face_feats):
        super().__init__()
        # normalised on the device by the same kernel as the queries (a host tensor is moved first; no CPU arithmetic here)
        feats = face_feats if face_feats.is_cuda else face_feats.cuda()
        self.face_feats = torch.nn.Parameter(ops.feats_normalize_fwd(feats)[0], requires_grad=False)
        # rows are unit vectors up to fp32 rounding: the bound the tensor-core search scales its TF32 error margin with
        self.db_norm_bound = 1.0 + 1e-5

    def forward(self, x):
        return None

    @torch.no_grad()
    def semantic_search(self, query_embeddings, selector=None, return_similarity=False):
        q = query_embeddings
        best, sim = ops.face_search_top1(q, selector, self.face_feats.data, db_norm_bound=self.db_norm_bound)
        hit = best >= 0
        target = torch.ones_like(q) * (-1)
        target[hit] = self.face_feats.data[best[hit]].to(q.dtype)
        if return_similarity:
            return target, sim.to(q.dtype)
        return target


def face_realism_loss(raw_feats, face_feats_ori, face_feats_model, face_indicators, *attr_args, confidence_level,
                      search_needs_target=None, fill_value=-1):
    """loss_face_ij of E1:1917-1929 (one attribute: ``targets, preds_ori, probs_ori``), E3:2124-2143 (two) and
    E4:2253-2272 (three) for the whole micro-batch, as one differentiable op on the RAW summed features
    ``net(x) + net(flip x)`` [b,d].  ``search_needs_target``: E1 searches only rows that have a target (default for one
    attribute), E3 / E4 search every row with a face."""
    n_attr = len(attr_args) // 3
    targets = [attr_args[3 * a] for a in range(n_attr)]
    preds = [attr_args[3 * a + 1] for a in range(n_attr)]
    probs = [attr_args[3 * a + 2] for a in range(n_attr)]
    if search_needs_target is None:
        search_needs_target = n_attr == 1
    db = face_feats_model.face_feats.data if isinstance(face_feats_model, torch.nn.Module) else face_feats_model
    return ag.FaceLoss.apply(raw_feats, face_feats_ori.to(torch.float32), db, face_indicators, float(confidence_level),
                             bool(search_needs_target), float(fill_value), n_attr, *targets, *preds, *probs)


# ----------------------------------------------------------------------------- hooks / weights
def _as_lists(args, n_attr):
    """(t0, pred0, prob0, t1, pred1, prob1, ...) -> ([t...], [pred...])"""
    return [args[3 * a] for a in range(n_attr)], [args[3 * a + 1] for a in range(n_attr)]


def _split_factors(attr_args, factors, names, defaults):
    """The reference's signatures end in ``factor...=default`` parameters that a caller may also pass positionally
    (E1:1584, E3:1751, E4:1823): trailing plain numbers of ``attr_args`` are those factors, in declaration order."""
    attr_args = list(attr_args)
    pos = []
    while attr_args and not torch.is_tensor(attr_args[-1]) and isinstance(attr_args[-1], (int, float)):
        pos.insert(0, attr_args.pop())
    if len(attr_args) % 3 != 0 or not 1 <= len(attr_args) // 3 <= 3:
        raise TypeError("expected (targets, preds_ori, probs_ori) per attribute")
    n_attr = len(attr_args) // 3
    names, defaults = names[n_attr], defaults[n_attr]
    if len(pos) > n_attr:
        raise TypeError(f"too many positional factors: {len(pos)} for {n_attr} attribute(s)")
    unknown = set(factors) - set(names)
    if unknown:
        raise TypeError(f"unexpected keyword argument(s) {sorted(unknown)}")
    vals = []
    for k, (nm, d) in enumerate(zip(names, defaults)):
        if k < len(pos):
            if nm in factors:
                raise TypeError(f"got multiple values for argument '{nm}'")
            vals.append(pos[k])
        else:
            vals.append(factors.get(nm, d))
    return attr_args, n_attr, vals


_FACTOR_NAMES = {1: ["factor"], 2: ["factor_gender", "factor_race"], 3: ["factor_gender", "factor_race", "factor_age"]}


def _hook(images, face_bboxs, face_bboxs_ori, targets, preds_ori, factors, e1_rule):
    H, W = images.shape[-2:]
    region, scale, _ = ops.guidance_factors(None, face_bboxs, face_bboxs_ori, targets, preds_ori, factors, None, e1_rule,
                                            H, W, want_region=True, want_weights=False)
    return ag.RegionScale.apply(images, region, scale)


def _weights(face_indicators, targets, preds_ori, factors, e1_rule, dtype):
    _, _, w = ops.guidance_factors(face_indicators, None, None, targets, preds_ori, None, factors, e1_rule, 0, 0,
                                   want_region=False, want_weights=True)
    return w.to(dtype)


def apply_grad_hook_face(images, face_bboxs, face_bboxs_ori, *attr_args, **factors):
    """E1:1584-1617 (``factor=``), E3:1751-1784 (``factor_gender=, factor_race=``), E4:1823-1867
    (``+ factor_age=``).  ``attr_args`` = (targets, preds_ori, probs_ori) per attribute, in the
    reference's positional order.  Forward is the identity; backward scales the face region."""
    attr_args, n_attr, f = _split_factors(attr_args, factors, _FACTOR_NAMES, {1: [0.1], 2: [0.3, 0.3], 3: [0.2, 0.3, 0.3]})
    targets, preds = _as_lists(attr_args, n_attr)
    return _hook(images, face_bboxs, face_bboxs_ori, targets, preds, f, n_attr == 1)


def gen_dynamic_weights(face_indicators, *attr_args, **factors):
    """E1:1619-1633 (``factor=``), E3:1787-1803, E4:1870-1895 -> weights [b] in probs_ori's dtype."""
    attr_args, n_attr, f = _split_factors(attr_args, factors, _FACTOR_NAMES, {1: [0.2], 2: [0.3, 0.6], 3: [0.2, 0.6, 0.6]})
    targets, preds = _as_lists(attr_args, n_attr)
    return _weights(face_indicators, targets, preds, f, n_attr == 1, attr_args[2].dtype)


# ----------------------------------------------------------------------------- loss
def fairness_ce_loss(logits, targets, face_indicators, fill_value=-1):
    """The three reference lines ``loss = ones*(-1); idx = (face * (targets != -1)).nonzero();
    loss[idx] = CE_loss(logits[idx], targets[idx])`` (E3:2114-2117) without the host sync."""
    return ag.FairCE.apply(logits, targets, face_indicators, float(fill_value))


# ----------------------------------------------------------------------------- collectives
def customized_all_gather(tensor, accelerator=None, return_tensor_other_processes=False):
    """E1:222-235 with one all_gather_into_tensor instead of a list gather + cat."""
    return fdist.customized_all_gather(tensor, accelerator, return_tensor_other_processes)


# ----------------------------------------------------------------------------- next rows
def stage_detector_input(images, to_host=True):
    """The detector's input array of ``get_face_app`` (E1:1317 + the BGR swap of E1:1326): uint8 ``[n,H,W,3]`` BGR.
    The conversion runs on the device; with ``to_host`` the result is copied into pinned host memory (3 bytes per
    pixel instead of the reference's float tensor) and returned as a numpy array, one row per ``face_app.get`` call."""
    staged = ops.stage_detector_input(images.detach())
    if not to_host:
        return staged
    host = torch.empty(staged.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(staged, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def get_evaluate_metrics(probs_gender_all, probs_race_all=None, probs_age_all=None):
    """E3:1716-1749 (5 numbers) / E4:1780-1821 (9 numbers with ``probs_age_all``) as python floats; one launch and one
    device-to-host read instead of one blocking ``.item()`` per number.  Called with ONE tensor it is exp-6's
    ``get_evaluate_metrics(probs_race_all)`` (exp-6-debias-race/1-main-debias.py:1624-1638): race0..3_freq, race_gap,
    race_pred_below_08."""
    if probs_race_all is None:
        return tuple(ops.bias_metrics(None, probs_gender_all, None).tolist())
    return tuple(ops.bias_metrics(probs_gender_all, probs_race_all, probs_age_all).tolist())


def bind(gender_classifier=None, gender_race_classifier=None, gender_race_age_classifier=None, accelerator=None,
         race_classifier=None):
    """Closures with the reference's exact signatures (the reference captures these objects from
    ``main``): ``fg = bind(gender_race_classifier=clf, accelerator=acc); fg.get_face_gender_race(chips, sel)``."""
    ns = types.SimpleNamespace(
        expand_bbox=expand_bbox, crop_face=crop_face, crop_faces=crop_faces, resize_small=resize_small,
        crop_and_resize=crop_and_resize, select_and_expand=select_and_expand,
        generate_dynamic_targets=generate_dynamic_targets,
        generate_dynamic_targets_gender_race=generate_dynamic_targets_gender_race,
        generate_dynamic_targets_gender_race_age=generate_dynamic_targets_gender_race_age,
        apply_grad_hook_face=apply_grad_hook_face, gen_dynamic_weights=gen_dynamic_weights,
        fairness_ce_loss=fairness_ce_loss, threshold_and_slice=threshold_and_slice)
    ns.get_face_gender = lambda face_chips, selector=None, fill_value=-1: get_face_gender(
        face_chips, selector, fill_value, gender_classifier=gender_classifier)
    ns.get_face_gender_race = lambda face_chips, selector=None, fill_value=-1: get_face_gender_race(
        face_chips, selector, fill_value, gender_race_classifier=gender_race_classifier)
    ns.get_face_gender_race_age = lambda face_chips, selector=None, fill_value=-1: get_face_gender_race_age(
        face_chips, selector, fill_value, gender_race_age_classifier=gender_race_age_classifier)
    ns.get_face_race = lambda face_chips, selector=None, fill_value=-1: get_face_race(
        face_chips, selector, fill_value, race_classifier=race_classifier)                                   # exp-6
    ns.customized_all_gather = lambda tensor, acc=accelerator, return_tensor_other_processes=False: customized_all_gather(
        tensor, acc, return_tensor_other_processes)
    ns.stage_detector_input = stage_detector_input
    ns.get_evaluate_metrics = get_evaluate_metrics
    ns.generate_dynamic_targets_race = generate_dynamic_targets_race                 # exp-6
    ns.image_pipeline, ns.aligned_face_chips = image_pipeline, aligned_face_chips   # E1:292
    ns.get_face_feats, ns.FaceFeatsModel, ns.face_realism_loss = get_face_feats, FaceFeatsModel, face_realism_loss
    from . import sync
    ns.make_grad_hook, ns.adjusted_dft_grad_coefs = sync.make_grad_hook, sync.adjusted_dft_grad_coefs
    ns.allreduce_average_gradients = sync.allreduce_average_gradients
    return ns
