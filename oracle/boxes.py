"""ORACLE (test infrastructure): face-box selection and expansion.

Restates get_largest_face_app (E1:1292-1304) and expand_bbox (E1:238-265).
Boxes are ``[x0, y0, x1, y1]`` float32 as insightface returns them.

Scalar promotion.  The reference pins NumPy 1.26.4 (environment.yml:102).  There, arithmetic
between two numpy float32 SCALARS stays float32, but a float32 scalar combined with a Python
int/float is promoted to float64.  (NumPy 2, installed here, keeps float32 -- NEP 50.)  The
box corners are rounded to integers, so the promotion decides rare half-way cases; this file
spells the pinned behaviour out with explicit casts so it does not depend on the installed
NumPy.  tests/golden/make_golden.py reproduces the same behaviour by running the reference
functions on scalar wrappers (LegacyF32).
"""
import numpy as np

F32 = np.float32
F64 = np.float64


def _clip_hi(v, dim_max):
    """min(v, dim_max) as Python evaluates it: the int wins only if it is strictly smaller.
    Returns (value, is_f32)."""
    return (F64(dim_max), False) if dim_max < v else (v, True)


def _clip_lo(v, dim_min):
    return (F64(dim_min), False) if dim_min > v else (v, True)


def _sub(a, b):
    (av, af), (bv, bf) = a, b
    if af and bf:
        return F32(av) - F32(bv), True
    return F64(av) - F64(bv), False


def _mul(a, b):
    (av, af), (bv, bf) = a, b
    if af and bf:
        return F32(av) * F32(bv), True
    return F64(av) * F64(bv), False


def largest_face_index(boxes, dim_max, dim_min=0):
    """Index of the detection with the largest area clipped to [dim_min, dim_max]^2.

    E1:1292-1304 -- a single detection is returned as is; otherwise a strict ``>`` against a
    running maximum that starts at 0, so the first maximum wins and a list whose clipped
    areas are all <= 0 yields index 0."""
    boxes = np.asarray(boxes, dtype=np.float32)
    if boxes.shape[0] == 1:
        return 0
    best_area = 0.0
    best = 0
    for k in range(boxes.shape[0]):
        bb = boxes[k]
        w = _sub(_clip_hi(bb[2], dim_max), _clip_lo(bb[0], dim_min))
        h = _sub(_clip_hi(bb[3], dim_max), _clip_lo(bb[1], dim_min))
        area = float(_mul(w, h)[0])
        if area > best_area:
            best_area = area
            best = k
    return best


def expand_bbox(bbox, expand_coef, target_ratio):
    """Square-ish expansion about the same centre, rounded half-to-even (E1:238-265).

    width, height and their ratio are float32 (scalar op scalar); everything that involves
    ``expand_coef`` / ``target_ratio`` / 0.5 (Python numbers) is float64."""
    b = [F32(v) for v in bbox]
    w = b[2] - b[0]
    h = b[3] - b[1]
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = h / w
    if F64(ratio) > target_ratio:
        extra_h = F64(h) * expand_coef
        extra_w = (F64(h) + extra_h) / target_ratio - F64(w)
    else:
        extra_w = F64(w) * expand_coef
        extra_h = (F64(w) + extra_w) * target_ratio - F64(h)
    out = [0, 0, 0, 0]
    out[0] = int(np.rint(F64(b[0]) - extra_w * 0.5))
    out[2] = int(np.rint(F64(b[2]) + extra_w * 0.5))
    out[1] = int(np.rint(F64(b[1]) - extra_h * 0.5))
    out[3] = int(np.rint(F64(b[3]) + extra_h * 0.5))
    return out


def select_and_expand(boxes, counts, dim_max, expand_coef=0.5, target_ratio=1, fill_value=-1):
    """Batched form used by the tests: ``boxes`` [n, F, 4] float32, ``counts`` [n] ints.

    Mirrors the per-image branch of get_face_app (E1:1324-1345): no detection -> indicator
    False and a ``[fill]*4`` box, else largest face -> expand_bbox(coef 0.5, ratio 1)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    out = np.full((n, 4), fill_value, dtype=np.int64)
    ind = np.zeros((n,), dtype=bool)
    for i in range(n):
        c = int(counts[i])
        if c == 0:
            continue
        k = largest_face_index(boxes[i, :c], dim_max)
        out[i] = expand_bbox(boxes[i, k], expand_coef, target_ratio)
        ind[i] = True
    return ind, out
