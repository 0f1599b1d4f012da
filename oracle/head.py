"""ORACLE (test infrastructure): attribute classifier heads.

Restates get_face_gender (E1:1355-1401), get_face_gender_race (E3:1387-1457) and
get_face_gender_race_age (E4:1378-1475).  The classifier (a closure variable in the
reference) is the first argument here.  Logit slices: E1 ``view(B,-1,2)[:,20,:]`` of an
80-way CelebA head; E3 ``[:, :2], [:, 2:]`` of 6; E4 ``[:, :2], [:, 2:6], [:, 6:]`` of 8.
"""
import torch

_SLICES = {
    "gender": lambda lg: [lg.view([lg.shape[0], -1, 2])[:, 20, :]],
    "gender_race": lambda lg: [lg[:, :2], lg[:, 2:]],
    "gender_race_age": lambda lg: [lg[:, :2], lg[:, 2:6], lg[:, 6:]],
    "race": lambda lg: [lg[:, 2:]],                   # exp-6-debias-race/1-main-debias.py:1381
}
_WIDTHS = {"gender": [2], "gender_race": [2, 4], "gender_race_age": [2, 4, 2], "race": [4]}


def _scatter(values, selector, fill_value):
    full = torch.ones([selector.shape[0]] + list(values.shape[1:]), dtype=values.dtype, device=values.device) * fill_value
    full[selector] = values
    return full


def _heads(kind, classifier, face_chips, selector, fill_value):
    chips = face_chips[selector] if selector is not None else face_chips
    widths = _WIDTHS[kind]
    if chips.shape[0] == 0:
        logits = [torch.empty([0, w], dtype=face_chips.dtype, device=face_chips.device) for w in widths]
        probs = [torch.empty([0, w], dtype=face_chips.dtype, device=face_chips.device) for w in widths]
        preds = [torch.empty([0], dtype=torch.int64, device=face_chips.device) for _ in widths]
    else:
        out = classifier(chips)
        logits = _SLICES[kind](out)
        probs = [torch.softmax(lg, dim=-1) for lg in logits]
        preds = [p.max(dim=-1).indices for p in probs]
    res = []
    for pd, pb, lg in zip(preds, probs, logits):
        if selector is not None:
            res += [_scatter(pd, selector, fill_value), _scatter(pb, selector, fill_value), _scatter(lg, selector, fill_value)]
        else:
            res += [pd, pb, lg]
    return tuple(res)


def get_face_gender(classifier, face_chips, selector=None, fill_value=-1):
    """-> (preds, probs, logits)"""
    return _heads("gender", classifier, face_chips, selector, fill_value)


def get_face_gender_race(classifier, face_chips, selector=None, fill_value=-1):
    """-> (preds_g, probs_g, logits_g, preds_r, probs_r, logits_r)"""
    return _heads("gender_race", classifier, face_chips, selector, fill_value)


def get_face_gender_race_age(classifier, face_chips, selector=None, fill_value=-1):
    """-> 9-tuple, gender / race / age.  With ``selector=None`` the reference returns only
    the gender and race entries (E4:1475) -- pinned by the golden vectors and kept."""
    out = _heads("gender_race_age", classifier, face_chips, selector, fill_value)
    return out if selector is not None else out[:6]


def get_face_race(classifier, face_chips, selector=None, fill_value=-1):
    """exp-6-debias-race/1-main-debias.py:1365-1411 -> (preds_race, probs_race, logits_race)."""
    return _heads("race", classifier, face_chips, selector, fill_value)


def mobilenet_head_reference(pooled, w1, b1, w2, b2):
    """torch fp32 reference of torchvision MobileNetV3 ``classifier`` in eval mode:
    Linear(960,1280) -> Hardswish -> Dropout(identity) -> Linear(1280,K)."""
    h = torch.nn.functional.hardswish(torch.nn.functional.linear(pooled, w1, b1))
    return torch.nn.functional.linear(h, w2, b2)
