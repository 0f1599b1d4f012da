"""ORACLE (test infrastructure): face-box selection and expansion.

Restates get_largest_face_app (E1:1292-1304) and expand_bbox (E1:238-265).
Boxes are ``[x0, y0, x1, y1]`` float32 as insightface returns them.
"""
import numpy as np


def largest_face_index(boxes, dim_max, dim_min=0):
    """Index of the detection with the largest area clipped to [dim_min, dim_max]^2.

    E1:1292-1304 -- a single detection is returned as is; otherwise a strict ``>``
    against a running maximum that starts at 0, so the first maximum wins and a list whose
    clipped areas are all <= 0 yields index 0.  Arithmetic stays in the boxes' float32.
    """
    boxes = np.asarray(boxes, dtype=np.float32)
    if boxes.shape[0] == 1:
        return 0
    best_area = 0
    best = 0
    for k in range(boxes.shape[0]):
        bb = boxes[k]
        w = min(bb[2], dim_max) - max(bb[0], dim_min)
        h = min(bb[3], dim_max) - max(bb[1], dim_min)
        area = w * h
        if area > best_area:
            best_area = area
            best = k
    return best


def expand_bbox(bbox, expand_coef, target_ratio):
    """Square-ish expansion about the same centre, rounded half-to-even (E1:238-265).

    ``bbox`` holds numpy float32 scalars in the reference (insightface output), so every
    product with a Python float stays float32; ``round`` on a numpy float32 is rint.
    """
    w = bbox[2] - bbox[0]
    h = bbox[3] - bbox[1]
    ratio = h / w
    if ratio > target_ratio:
        extra_h = h * expand_coef
        extra_w = (h + extra_h) / target_ratio - w
    elif ratio <= target_ratio:
        extra_w = w * expand_coef
        extra_h = (w + extra_w) * target_ratio - h
    out = [0, 0, 0, 0]
    out[0] = int(round(bbox[0] - extra_w * 0.5))
    out[2] = int(round(bbox[2] + extra_w * 0.5))
    out[1] = int(round(bbox[1] - extra_h * 0.5))
    out[3] = int(round(bbox[3] + extra_h * 0.5))
    return out


def select_and_expand(boxes, counts, dim_max, expand_coef=0.5, target_ratio=1, fill_value=-1):
    """Batched form used by the tests: ``boxes`` [n, F, 4] float32, ``counts`` [n] ints.

    Mirrors the per-image branch of get_face_app (E1:1324-1345): no detection -> indicator
    False and a ``[fill]*4`` box, else largest face -> expand_bbox(coef 0.5, ratio 1).
    """
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
