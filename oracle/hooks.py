"""ORACLE (test infrastructure): face-region gradient scaling and dynamic loss weights.

Restates apply_grad_hook_face (E1:1584-1617, E3:1751-1784, E4:1823-1867) and
gen_dynamic_weights (E1:1619-1633, E3:1787-1803, E4:1870-1895).

One function per family serves all three experiments: ``targets`` / ``preds_ori`` /
``factors`` are lists with one entry per attribute (1, 2 or 3 entries); ``e1_rule=True``
selects E1's three-way branch in which a thresholded target (-1) always takes the factor
(E1:1599-1604); the E3/E4 rule compares target with the original prediction only.
"""
import itertools

import torch


def _factor_for(ts, ps, factors, e1_rule):
    if e1_rule:
        t, p = ts[0], ps[0]
        if t == -1:
            return factors[0]
        elif t == p:
            return 1
        return factors[0]
    wrong = [f for t, p, f in zip(ts, ps, factors) if t != p]
    return min(wrong) if wrong else 1


def apply_grad_hook_face(images, face_bboxs, face_bboxs_ori, targets, preds_ori, factors, e1_rule=False):
    """Forward identity; backward multiplies the gradient inside
    ``bbox ^ bbox_ori ^ image`` by the per-image factor.  Only the NEW box is tested for the
    -1 sentinel (E1:1589); a ``[-1]*4`` original box therefore produces python slice ends of
    -1, i.e. the region stops one pixel short of the border -- reproduced, not fixed."""
    out = []
    for i in range(images.shape[0]):
        image, bb, bo = images[i], face_bboxs[i], face_bboxs_ori[i]
        if (bb == -1).all():
            out.append(image.unsqueeze(dim=0))
            continue
        img_width, img_height = image.shape[1:]          # names swapped in the reference, kept
        left = max(bb[0], bo[0], 0)
        right = min(bb[2], bo[2], img_width)
        bottom = max(bb[1], bo[1], 0)
        top = min(bb[3], bo[3], img_height)
        face = image[:, bottom:top, left:right].clone()
        coef = _factor_for([t[i] for t in targets], [p[i] for p in preds_ori], factors, e1_rule)
        face.register_hook(lambda g, c=coef: c * g)
        add = torch.zeros_like(image)
        add[:, bottom:top, left:right] = face
        mask = torch.zeros_like(image)
        mask[:, bottom:top, left:right] = 1
        out.append((mask * add + (1 - mask) * image).unsqueeze(dim=0))
    return torch.cat(out)


def gen_dynamic_weights(face_indicators, targets, preds_ori, factors, dtype, e1_rule=False):
    """Per-image weight on the image-semantics loss.  No face: 1 under the E1 rule
    (E1:1622-1623), min(factors) otherwise (E3:1790-1791, E4:1880-1881)."""
    w = []
    for i in range(face_indicators.shape[0]):
        if (face_indicators[i] == False).all():  # noqa: E712  (reference spelling)
            w.append(1 if e1_rule else min(factors))
        else:
            w.append(_factor_for([t[i] for t in targets], [p[i] for p in preds_ori], factors, e1_rule))
    return torch.tensor(w, dtype=dtype, device=face_indicators.device)
