"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers and the
current CUDA stream, raise on a non-zero return code.  No arithmetic happens here."""
import ctypes
import os

import torch

from . import _lib
from ._lib import check

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"fairguide: unsupported dtype {t.dtype}") from None


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fairguide ops run on CUDA tensors only (no CPU fallback)")


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c(t):
    return None if t is None else t.contiguous()


def _u8(t):
    if t is None:
        return None
    return (t.to(torch.uint8) if t.dtype != torch.bool else t.view(torch.uint8)).contiguous()


def _farr(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def _iarr(vals):
    return (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])


# ------------------------------------------------------------------ boxes
def select_expand_boxes(boxes, counts, dim_max, expand_coef=0.5, target_ratio=1.0, fill=-1):
    """boxes [n,F,4] float32, counts [n] int32|None -> (indicators bool [n], boxes int64 [n,4])."""
    _cuda(boxes, counts)
    boxes = boxes.to(torch.float32).contiguous()
    n, F = boxes.shape[0], boxes.shape[1]
    counts = None if counts is None else counts.to(torch.int32).contiguous()
    out = torch.empty((n, 4), dtype=torch.int64, device=boxes.device)
    ind = torch.empty((n,), dtype=torch.uint8, device=boxes.device)
    check(_lib.lib().fg_select_expand_boxes(_p(boxes), _p(counts), n, F, int(dim_max), float(expand_coef),
                                            float(target_ratio), int(fill), _p(out), _p(ind), _stream()),
          "fg_select_expand_boxes")
    return ind.view(torch.bool), out


# ------------------------------------------------------------------ crop / resize
def crop_resize_fwd(images, boxes, indicators, chip_hw, small_hw, fill_value=-1.0):
    """images [n,C,H,W]; returns (chips|None, small|None)."""
    _cuda(images, boxes, indicators)
    images = images.contiguous()
    n, C, H, W = images.shape
    chips = small = None
    ch = cw = sh = sw = 0
    if chip_hw is not None:
        ch, cw = chip_hw
        chips = torch.empty((n, C, ch, cw), dtype=images.dtype, device=images.device)
        boxes = boxes.to(torch.int64).contiguous()
    if small_hw is not None:
        sh, sw = small_hw
        small = torch.empty((n, C, sh, sw), dtype=images.dtype, device=images.device)
    ind = _u8(indicators)
    check(_lib.lib().fg_crop_resize_fwd(_p(images), n, C, H, W, _p(boxes) if chips is not None else None, _p(ind),
                                        _p(chips), ch, cw, _p(small), sh, sw, float(fill_value), _dt(images), _stream()),
          "fg_crop_resize_fwd")
    return chips, small


def image_grad(g_chips, g_small, boxes, indicators, region, scale, image_shape, dtype, device):
    n, C, H, W = image_shape
    _cuda(g_chips, g_small, boxes, indicators, region, scale)
    g_chips, g_small = _c(g_chips), _c(g_small)
    ch = cw = sh = sw = 0
    if g_chips is not None:
        ch, cw = g_chips.shape[-2:]
        boxes = boxes.to(torch.int64).contiguous()
    if g_small is not None:
        sh, sw = g_small.shape[-2:]
    out = torch.empty((n, C, H, W), dtype=dtype, device=device)
    ind = _u8(indicators)
    region, scale = _region_scale(region, scale)
    check(_lib.lib().fg_image_grad(_p(g_chips), _p(g_small), _p(boxes) if g_chips is not None else None, _p(ind),
                                   _p(region), _p(scale), _p(out), n, C, H, W, ch, cw, sh, sw, _DT[dtype], _stream()),
          "fg_image_grad")
    return out


def _region_scale(region, scale):
    """region int32 [n,4] / scale float32 [n], whatever integer / floating dtype the caller holds them in."""
    if (region is None) != (scale is None):
        raise RuntimeError("fairguide: region and scale go together")
    if region is None:
        return None, None
    return region.to(torch.int32).contiguous(), scale.to(torch.float32).contiguous()


def region_scale(g, region, scale):
    _cuda(g, region, scale)
    region, scale = _region_scale(region, scale)
    g = g.contiguous()
    n, C, H, W = g.shape
    out = torch.empty_like(g)
    check(_lib.lib().fg_region_scale(_p(g), _p(region), _p(scale), _p(out), n, C, H, W, _dt(g), _stream()), "fg_region_scale")
    return out


def guidance_factors(face_indicators, bbox, bbox_ori, targets, preds_ori, hook_factors, weight_factors, e1_rule, H, W,
                     want_region=True, want_weights=True):
    """targets / preds_ori: lists (1..3) of int64 [n].  Returns (region int32 [n,4], scale f32 [n], weights f32 [n])."""
    n_attr = len(targets)
    dev = targets[0].device
    n = targets[0].shape[0]
    _cuda(face_indicators, bbox, bbox_ori, *targets, *preds_ori)
    t = [x.to(torch.int64).contiguous() for x in targets] + [None] * (3 - n_attr)
    p = [x.to(torch.int64).contiguous() for x in preds_ori] + [None] * (3 - n_attr)
    region = scale = weights = None
    if want_region:
        region = torch.empty((n, 4), dtype=torch.int32, device=dev)
        scale = torch.empty((n,), dtype=torch.float32, device=dev)
        bbox = bbox.to(torch.int64).contiguous()
        bbox_ori = bbox_ori.to(torch.int64).contiguous()
    if want_weights:
        weights = torch.empty((n,), dtype=torch.float32, device=dev)
    face = _u8(face_indicators)
    hf = _farr(hook_factors) if hook_factors is not None else None
    wf = _farr(weight_factors) if weight_factors is not None else None
    check(_lib.lib().fg_guidance_factors(_p(face), _p(bbox) if want_region else None, _p(bbox_ori) if want_region else None,
                                         _p(t[0]), _p(t[1]), _p(t[2]), _p(p[0]), _p(p[1]), _p(p[2]),
                                         hf, wf, n_attr, 1 if e1_rule else 0, n, H, W, _p(region), _p(scale), _p(weights), _stream()),
          "fg_guidance_factors")
    return region, scale, weights


# ------------------------------------------------------------------ head
def head_fwd(pooled, w1, b1, w2, b2):
    _cuda(pooled, w1, b1, w2, b2)
    pooled = pooled.contiguous()
    m, d_in = pooled.shape
    d_hid, k_head = w1.shape[0], w2.shape[0]
    dt = pooled.dtype
    w1, b1, w2, b2 = (x.to(dt).contiguous() for x in (w1, b1, w2, b2))
    hidden = torch.empty((m, d_hid), dtype=dt, device=pooled.device)
    logits = torch.empty((m, k_head), dtype=torch.float32, device=pooled.device)
    nbytes = _lib.lib().fg_head_workspace_bytes(m, d_in, d_hid, k_head, _dt(pooled))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=pooled.device)
    check(_lib.lib().fg_head_fwd(_p(pooled), _p(w1), _p(b1), _p(w2), _p(b2), m, d_in, d_hid, k_head, _p(hidden), _p(logits),
                                 _p(ws), nbytes, _dt(pooled), _stream()), "fg_head_fwd")
    return logits, hidden


def head_bwd(g_logits, hidden, w1, w2):
    _cuda(g_logits, hidden, w1, w2)
    g_logits = g_logits.to(torch.float32).contiguous()
    m, d_hid = hidden.shape
    d_in, k_head = w1.shape[1], w2.shape[0]
    dt = hidden.dtype
    w1, w2 = w1.to(dt).contiguous(), w2.to(dt).contiguous()
    nbytes = _lib.lib().fg_head_workspace_bytes(m, d_in, d_hid, k_head, _dt(hidden))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=hidden.device)
    g_pooled = torch.empty((m, d_in), dtype=dt, device=hidden.device)
    check(_lib.lib().fg_head_bwd(_p(g_logits), _p(hidden), _p(w1), _p(w2), m, d_in, d_hid, k_head, _p(g_pooled), _p(ws), nbytes,
                                 _dt(hidden), _stream()), "fg_head_bwd")
    return g_pooled


def head_attributes(logits, src_row, selector, n, col_start, width, fill, dtype, return_flat=False):
    """-> (preds [A,n] int64, probs list of [n,w_a], logits list of [n,w_a]); attribute-major flat storage."""
    _cuda(logits, src_row, selector)
    dev = logits.device
    logits = logits.to(torch.float32).contiguous()
    A = len(width)
    tot = sum(width)
    preds = torch.empty((A, n), dtype=torch.int64, device=dev)
    probs = torch.empty((n * tot,), dtype=dtype, device=dev)
    lout = torch.empty((n * tot,), dtype=dtype, device=dev)
    sel = _u8(selector)
    src = None if src_row is None else src_row.to(torch.int32).contiguous()
    check(_lib.lib().fg_head_attributes(_p(logits), logits.shape[0], logits.shape[1], _p(src), _p(sel), n, A,
                                        _iarr(col_start), _iarr(width), float(fill), _p(preds), _p(probs), _p(lout),
                                        _DT[dtype], _stream()), "fg_head_attributes")
    ps, ls, off = [], [], 0
    for w in width:
        ps.append(probs[off:off + n * w].view(n, w))
        ls.append(lout[off:off + n * w].view(n, w))
        off += n * w
    if return_flat:
        return preds, ps, ls, probs
    return preds, ps, ls


def head_attributes_bwd(probs_flat, g_probs, g_logits_attr, src_row, selector, n, m, k_head, col_start, width):
    """Backward of head_attributes in one launch -> g_logits_full float32 [m,k_head].  ``probs_flat``: the forward's flat
    attribute-major probs storage; g_probs / g_logits_attr: lists (one entry per attribute) of [n,w_a] tensors or None."""
    A = len(width)
    dt = probs_flat.dtype
    gp = [None if g is None else g.to(dt).contiguous() for g in g_probs]
    gl = [None if g is None else g.to(dt).contiguous() for g in g_logits_attr]
    _cuda(probs_flat, src_row, selector, *gp, *gl)
    g_full = torch.empty((m, k_head), dtype=torch.float32, device=probs_flat.device)
    arr = lambda ts: (ctypes.c_void_p * A)(*[None if t is None else t.data_ptr() for t in ts])
    sel = _u8(selector)
    src = None if src_row is None else src_row.to(torch.int32).contiguous()
    check(_lib.lib().fg_head_attributes_bwd(_p(probs_flat), arr(gp), arr(gl), _p(src), _p(sel), n, m, k_head, A, _iarr(col_start),
                                            _iarr(width), _p(g_full), _DT[dt], _stream()), "fg_head_attributes_bwd")
    return g_full


# ------------------------------------------------------------------ loss
def fair_ce_fwd(logits, targets, face, fill=-1.0):
    _cuda(logits, targets, face)
    logits = logits.contiguous()
    n, k = logits.shape
    loss = torch.empty((n,), dtype=logits.dtype, device=logits.device)
    check(_lib.lib().fg_fair_ce_fwd(_p(logits), _p(targets.to(torch.int64).contiguous()), _p(_u8(face)), n, k, float(fill),
                                    _p(loss), _dt(logits), _stream()), "fg_fair_ce_fwd")
    return loss


def fair_ce_bwd(logits, targets, face, g_loss):
    logits = logits.contiguous()
    n, k = logits.shape
    g = torch.empty_like(logits)
    check(_lib.lib().fg_fair_ce_bwd(_p(logits), _p(targets.to(torch.int64).contiguous()), _p(_u8(face)),
                                    _p(g_loss.to(logits.dtype).contiguous()), n, k, _p(g), _dt(logits), _stream()), "fg_fair_ce_bwd")
    return g


def fair_loss_fused(logits_attr, targets, col_start, k_head, face, g_coef, dyn_w=None, loss_clip=None, loss_dino=None,
                    loss_face=None, weight_img=0.0, weight_face=0.0, fill=-1.0, want_grad=True):
    """CE of every attribute + loss assembly + gradient wrt the head logits in one launch (E3:2114-2147).
    logits_attr / targets: lists of [n,w_a] (dtype) and [n] int64.  Returns (loss_fair list, loss f32 [n], g_logits f32)."""
    import ctypes
    A = len(logits_attr)
    lg = [t.contiguous() for t in logits_attr]
    tg = [t.to(torch.int64).contiguous() for t in targets]
    _cuda(*lg, *tg, face)
    n, dt, dev = lg[0].shape[0], lg[0].dtype, lg[0].device
    cast = lambda t: None if t is None else t.to(dt).contiguous()
    lc, ld, lf = cast(loss_clip), cast(loss_dino), cast(loss_face)
    loss_fair = torch.empty((A, n), dtype=dt, device=dev)
    loss = torch.empty((n,), dtype=torch.float32, device=dev)
    g = torch.empty((n, k_head), dtype=torch.float32, device=dev) if want_grad else None
    ptrs = lambda ts: (ctypes.c_void_p * A)(*[t.data_ptr() for t in ts])
    ints = lambda xs: (ctypes.c_int32 * A)(*[int(x) for x in xs])
    check(_lib.lib().fg_fair_loss_fused(ptrs(lg), ptrs(tg), ints([t.shape[1] for t in lg]), ints(col_start), A, _p(_u8(face)), n, int(k_head),
                                        float(fill), float(g_coef), _p(None if dyn_w is None else dyn_w.float().contiguous()),
                                        _p(lc), _p(ld), _p(lf), float(weight_img), float(weight_face),
                                        _p(loss_fair), _p(loss), _p(g), _dt(lg[0]), _stream()), "fg_fair_loss_fused")
    return list(loss_fair.unbind(0)), loss, g


# ------------------------------------------------------------------ assignment
def assign_rank_binom(probs, target_ratio=0.5, threshold=-1.0, w_uncertainty=True):
    _cuda(probs)
    probs = probs.contiguous()
    n_all = probs.shape[0]
    targets = torch.empty((n_all,), dtype=torch.int64, device=probs.device)
    unc = torch.empty((n_all,), dtype=probs.dtype, device=probs.device) if w_uncertainty else None
    nbytes = _lib.lib().fg_rank_binom_workspace_bytes(n_all)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=probs.device)
    check(_lib.lib().fg_assign_rank_binom(_p(probs), n_all, float(target_ratio), float(threshold), _p(targets), _p(unc),
                                          _p(ws), nbytes, _dt(probs), _stream()), "fg_assign_rank_binom")
    return targets, unc


class OtWorkspace:
    """Device scratch shared by ot_plan_counts and ot_targets of one assignment."""

    def __init__(self, n_all, K, S, device):
        self.n_all, self.K, self.S = n_all, K, S
        self.nbytes = _lib.lib().fg_ot_workspace_bytes(n_all, K, S)
        self.buf = torch.empty((max(self.nbytes, 16),), dtype=torch.uint8, device=device)

    def status(self):
        """[status bits, n_valid seen on the device, base augmentations, draw augmentations] (synchronises)."""
        return self.buf[:16].view(torch.int32).tolist()

    def status_tensor(self):
        """The four status words as a device view (copy it to the host together with the step's outputs)."""
        return self.buf[:16].view(torch.int32)

    def check(self, n_valid=None):
        """Raise if the assignment kernels flagged a problem (one 16-byte device-to-host read)."""
        check_ot_status(self.status(), n_valid)


OT_STATUS_BITS = {1: "n_valid passed by the caller differs from the rows with a face found on the device",
                  2: "no augmenting path (infeasible demand)", 4: "augmenting path too long", 8: "repair-step cap reached (plan inexact)",
                  16: "class demands do not sum to the number of rows"}


def check_ot_status(words, n_valid=None):
    bits = int(words[0])
    if bits:
        why = "; ".join(msg for b, msg in OT_STATUS_BITS.items() if bits & b)
        raise RuntimeError(f"fairguide assignment failed (status {bits}: {why}; n_valid given {n_valid}, on device {int(words[1])})")


def ot_plan_counts(probs_gender, probs_race, probs_age, rands, n_valid, ws):
    """This rank's per-(row,class) plan counts, int32 [n_valid,K]."""
    _cuda(probs_gender, probs_race, probs_age, *rands)
    K = ws.K
    dev = probs_gender.device
    pg, pr = probs_gender.contiguous(), probs_race.contiguous()
    pa = None if probs_age is None else probs_age.contiguous()
    S = rands[0].shape[0] if (n_valid > 0 and len(rands) > 0) else 0
    r = [x.to(pg.dtype).contiguous() for x in rands] + [None] * (3 - len(rands))
    counts = torch.empty((n_valid, K), dtype=torch.int32, device=dev)
    check(_lib.lib().fg_ot_plan_counts(_p(pg), _p(pr), _p(pa), pg.shape[0], _p(r[0]), _p(r[1]), _p(r[2]), S, n_valid,
                                       _p(counts), _p(ws.buf), ws.nbytes, _dt(pg), _stream()), "fg_ot_plan_counts")
    return counts


def ot_targets(counts, probs_gender, probs_race, n_valid, ws, threshold=-1.0, w_uncertainty=True):
    K = ws.K
    dev = probs_gender.device
    n_all = probs_gender.shape[0]
    dt = probs_gender.dtype
    A = 3 if K == 16 else 2
    ts = [torch.empty((n_all,), dtype=torch.int64, device=dev) for _ in range(A)] + [None] * (3 - A)
    us = ([torch.empty((n_all,), dtype=dt, device=dev) for _ in range(A)] if w_uncertainty else [None] * A) + [None] * (3 - A)
    check(_lib.lib().fg_ot_targets(_p(counts), _p(probs_gender), _p(probs_race), n_all, n_valid, K, float(threshold),
                                   _p(ts[0]), _p(us[0]), _p(ts[1]), _p(us[1]), _p(ts[2]), _p(us[2]),
                                   _p(ws.buf), ws.nbytes, _DT[dt], _stream()), "fg_ot_targets")
    return ts[:A], us[:A]


def assign_race_enumerated(probs_race, n_valid, demands, weights, threshold=-1.0, w_uncertainty=True):
    """E6:1413-1482 on the device: exact plan per enumerated composition, weighted accumulation in the given order.
    demands int32 [S,16] (classes 4..15 zero), weights float64 [S] -> (targets int64 [n_all], uncertainty or None);
    also returns the workspace (status words)."""
    _cuda(probs_race, demands, weights)
    pr = probs_race.contiguous()
    n_all, dev = pr.shape[0], pr.device
    S = int(demands.shape[0]) if n_valid > 0 else 0
    assert n_valid == 0 or (demands.dtype == torch.int32 and tuple(demands.shape) == (S, 16) and weights.dtype == torch.float64)
    nbytes = _lib.lib().fg_race_workspace_bytes(n_all, S)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    targets = torch.empty((n_all,), dtype=torch.int64, device=dev)
    unc = torch.empty((n_all,), dtype=pr.dtype, device=dev) if w_uncertainty else None
    check(_lib.lib().fg_assign_race_enumerated(_p(pr), n_all, n_valid, _p(demands.contiguous()) if S else None,
                                               _p(weights.contiguous()) if S else None, S, float(threshold), _p(targets), _p(unc),
                                               _p(ws), nbytes, _dt(pr), _stream()), "fg_assign_race_enumerated")
    return targets, unc, ws


def race_cost_matrix(probs_race, n_valid):
    """Test hook: the E6 cost matrix [n_valid,4] float64."""
    _cuda(probs_race)
    pr = probs_race.contiguous()
    nbytes = _lib.lib().fg_race_workspace_bytes(pr.shape[0], 0)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=pr.device)
    M = torch.empty((n_valid, 4), dtype=torch.float64, device=pr.device)
    check(_lib.lib().fg_race_cost_matrix(_p(pr), pr.shape[0], n_valid, _p(M), _p(ws), nbytes, _dt(pr), _stream()), "fg_race_cost_matrix")
    return M


def ot_solve_single(M, b):
    """Test hook: exact assignment int32 [n] of the rows of M [n,K] float64 to classes with sizes b."""
    _cuda(M)
    M = M.to(torch.float64).contiguous()
    n, K = M.shape
    ws = OtWorkspace(n, K, 0, M.device)
    out = torch.empty((n,), dtype=torch.int32, device=M.device)
    barr = (ctypes.c_int64 * K)(*[int(v) for v in b])
    check(_lib.lib().fg_ot_solve_single(_p(M), n, K, barr, _p(out), _p(ws.buf), ws.nbytes, _stream()), "fg_ot_solve_single")
    return out, ws


def ot_cost_matrix(probs_gender, probs_race, probs_age, n_valid):
    _cuda(probs_gender, probs_race, probs_age)
    K = 8 if probs_age is None else 16
    n_all = probs_gender.shape[0]
    ws = OtWorkspace(n_all, K, 0, probs_gender.device)
    M = torch.empty((n_valid, K), dtype=torch.float64, device=probs_gender.device)
    pa = None if probs_age is None else probs_age.contiguous()
    check(_lib.lib().fg_ot_cost_matrix(_p(probs_gender.contiguous()), _p(probs_race.contiguous()), _p(pa), n_all, n_valid, _p(M),
                                       _p(ws.buf), ws.nbytes, _dt(probs_gender), _stream()), "fg_ot_cost_matrix")
    return M


# ------------------------------------------------------------------ next rows
def stage_detector_input(images):
    """[n,3,H,W] in [-1,1] -> uint8 [n,H,W,3] BGR on the device (E1:1317 + 1326): what face_app.get() is fed."""
    _cuda(images)
    images = images.contiguous()
    n, C, H, W = images.shape
    out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=images.device)
    check(_lib.lib().fg_stage_detector_input(_p(images), n, C, H, W, _p(out), _dt(images), _stream()), "fg_stage_detector_input")
    return out


def bias_metrics(probs_gender, probs_race, probs_age=None):
    """get_evaluate_metrics (E3:1716 / E4:1780) as one launch; returns a DEVICE fp64 tensor of 5 (or 9) numbers.
    ``probs_gender=None``: exp-6's race-only form (E6:1624-1638), 6 numbers."""
    _cuda(probs_gender, probs_race, probs_age)
    pr = probs_race.contiguous()
    pg = None if probs_gender is None else probs_gender.to(pr.dtype).contiguous()
    pa = None if probs_age is None else probs_age.to(pr.dtype).contiguous()
    out = torch.empty((6 if pg is None else 9 if pa is not None else 5,), dtype=torch.float64, device=pr.device)
    check(_lib.lib().fg_bias_metrics(_p(pg), _p(pr), _p(pa), pr.shape[0], _p(out), _dt(pr), _stream()), "fg_bias_metrics")
    return out


# ------------------------------------------------------------------ f1: aligned 112x112 chip
def align_matrices(landmarks, indicators, src_hw, dst_hw=(112, 112), want_matrix=False):
    """landmarks [n,5,2] float32 -> params [n,16] float32 (for the two warp kernels) and, optionally, the similarity
    matrices [n,2,3] float64 of E1:304-307."""
    _cuda(landmarks, indicators)
    lm = landmarks.to(torch.float32).contiguous()
    n = lm.shape[0]
    params = torch.empty((max(n, 1), 16), dtype=torch.float32, device=lm.device)
    M = torch.empty((n, 2, 3), dtype=torch.float64, device=lm.device) if want_matrix else None
    check(_lib.lib().fg_align_matrices(_p(lm), _p(_u8(indicators)), n, src_hw[0], src_hw[1], dst_hw[0], dst_hw[1],
                                       _p(M), _p(params), _stream()), "fg_align_matrices")
    return (params, M) if want_matrix else params


def aligned_warp_fwd(images, params, indicators, dst_hw=(112, 112), fill=-1.0):
    _cuda(images, params, indicators)
    images = images.contiguous()
    n, C, H, W = images.shape
    out = torch.empty((n, C, dst_hw[0], dst_hw[1]), dtype=images.dtype, device=images.device)
    check(_lib.lib().fg_aligned_warp_fwd(_p(images), n, C, H, W, _p(params), _p(_u8(indicators)), _p(out), dst_hw[0], dst_hw[1],
                                         float(fill), _dt(images), _stream()), "fg_aligned_warp_fwd")
    return out


def aligned_warp_bwd(g_out, params, indicators, image_shape, g_images=None):
    """Gradient of aligned_warp_fwd wrt the images; with ``g_images`` given it is accumulated in place."""
    _cuda(g_out, params, indicators, g_images)
    g_out = g_out.contiguous()
    n, C, Hd, Wd = g_out.shape
    acc = g_images is not None
    if acc:
        assert g_images.is_contiguous() and tuple(g_images.shape) == tuple(image_shape) and g_images.dtype == g_out.dtype
    else:
        g_images = torch.empty(tuple(image_shape), dtype=g_out.dtype, device=g_out.device)
    check(_lib.lib().fg_aligned_warp_bwd(_p(g_out), n, C, Hd, Wd, _p(params), _p(_u8(indicators)), _p(g_images), image_shape[2],
                                         image_shape[3], 1 if acc else 0, _dt(g_out), _stream()), "fg_aligned_warp_bwd")
    return g_images


# ------------------------------------------------------------------ f1: face-realism loss
def feats_normalize_fwd(raw):
    _cuda(raw)
    raw = raw.contiguous()
    n, d = raw.shape
    f = torch.empty((n, d), dtype=torch.float32, device=raw.device)
    inv = torch.empty((n,), dtype=torch.float32, device=raw.device)
    check(_lib.lib().fg_feats_normalize_fwd(_p(raw), n, d, _p(f), _p(inv), _dt(raw), _stream()), "fg_feats_normalize_fwd")
    return f, inv


def feats_normalize_bwd(g_f, f, inv, dtype):
    n, d = f.shape
    g_raw = torch.empty((n, d), dtype=dtype, device=f.device)
    check(_lib.lib().fg_feats_normalize_bwd(_p(g_f.to(torch.float32).contiguous()), _p(f), _p(inv), n, d, _p(g_raw), _DT[dtype], _stream()),
          "fg_feats_normalize_bwd")
    return g_raw


SEARCH_TC_MIN_QUERIES = 16     # below this the streaming search is HBM-bound on the database and as fast as anything can be


def face_search_top1(queries, selector, db, db_norm_bound=None):
    """-> (best_row int64 [m] (-1 where not selected), similarity float32 [m]).
    ``db_norm_bound`` (>= the largest L2 norm of a database row; computed once per database by the caller) enables the
    tensor-core search for batches of >= SEARCH_TC_MIN_QUERIES queries; the results are identical either way."""
    _cuda(queries, selector, db)
    q = queries.to(torch.float32).contiguous()
    m, d = q.shape
    assert db.dtype == torch.float32 and db.is_contiguous() and db.shape[1] == d
    best = torch.empty((m,), dtype=torch.int64, device=q.device)
    sim = torch.empty((m,), dtype=torch.float32, device=q.device)
    if db_norm_bound is not None and m >= SEARCH_TC_MIN_QUERIES and d % 32 == 0 and db.shape[0] >= 128 and os.environ.get("FG_SEARCH_EXACT") is None:
        nbytes = _lib.lib().fg_face_search_tc_workspace_bytes(m, db.shape[0])
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=q.device)
        check(_lib.lib().fg_face_search_top1_tc(_p(q), _p(_u8(selector)), m, _p(db), db.shape[0], d, float(db_norm_bound), _p(best), _p(sim),
                                                _p(ws), nbytes, _stream()), "fg_face_search_top1_tc")
        return best, sim
    nbytes = _lib.lib().fg_face_search_workspace_bytes(m)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=q.device)
    check(_lib.lib().fg_face_search_top1(_p(q), _p(_u8(selector)), m, _p(db), db.shape[0], d, _p(best), _p(sim), _p(ws), nbytes,
                                         _stream()), "fg_face_search_top1")
    return best, sim


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def face_loss_fwd(raw_feats, feats_ori, db, face_indicators, targets, preds_ori, probs_ori, confidence_level,
                  search_needs_target, fill=-1.0):
    """-> (loss [n] in raw_feats.dtype, workspace) ; the workspace carries the state face_loss_bwd needs."""
    _cuda(raw_feats, feats_ori, db, face_indicators, *targets, *preds_ori, *probs_ori)
    raw = raw_feats.contiguous()
    n, d = raw.shape
    assert feats_ori.dtype == torch.float32 and db.dtype == torch.float32 and db.is_contiguous()
    fo = feats_ori.contiguous()
    ts = [t.to(torch.int64).contiguous() for t in targets]
    ps = [t.to(torch.int64).contiguous() for t in preds_ori]
    qs = [t.to(raw.dtype).contiguous() for t in probs_ori]
    widths = _iarr([q.shape[1] for q in qs])
    nbytes = _lib.lib().fg_face_loss_workspace_bytes(n, d)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=raw.device)
    loss = torch.empty((n,), dtype=raw.dtype, device=raw.device)
    ind = _u8(face_indicators)
    check(_lib.lib().fg_face_loss_fwd(_p(raw), _p(fo), _p(db), n, d, db.shape[0], _p(ind), _ptr_array(ts), _ptr_array(ps),
                                      _ptr_array(qs), widths, len(ts), float(confidence_level), 1 if search_needs_target else 0,
                                      float(fill), _p(loss), _p(ws), nbytes, _dt(raw), _stream()), "fg_face_loss_fwd")
    return loss, ws


def face_loss_bwd(g_loss, feats_ori, db, ws, n, d, dtype):
    g_raw = torch.empty((n, d), dtype=dtype, device=db.device)
    check(_lib.lib().fg_face_loss_bwd(_p(g_loss.to(dtype).contiguous()), _p(feats_ori.contiguous()), _p(db), n, d, _p(g_raw), _p(ws),
                                      ws.numel(), _DT[dtype], _stream()), "fg_face_loss_bwd")
    return g_raw


def face_loss_target_rows(ws, n, d):
    out = torch.empty((n,), dtype=torch.int64, device=ws.device)
    check(_lib.lib().fg_face_loss_target_rows(_p(ws), n, d, _p(out), _stream()), "fg_face_loss_target_rows")
    return out
