"""Oracle for the two "next" rows: detector staging and the bias-gap metrics.  Test infrastructure only.

    stage_detector_input   E1:1317 (the uint8 conversion) + E1:1326 (the BGR swap), same lines in E3 / E4
    evaluate_metrics       get_evaluate_metrics, E3:1716-1749 and E4:1780-1821

torch-CPU restatements; pinned by tests/golden/nextrows.npz, which holds the outputs of the reference's own statements
executed by tests/golden/make_golden_next.py.
"""
import numpy as np
import torch


def stage_detector_input(images):
    """images [n,3,H,W] (any float dtype) -> uint8 [n,H,W,3], channels B,G,R.
    The three arithmetic steps run in the tensor's dtype (one rounding each), as in the eager expression."""
    scaled = images.detach() * 0.5
    scaled = scaled + 0.5
    scaled = scaled * 255
    hwc = scaled.cpu().permute(0, 2, 3, 1).float().numpy().astype(np.uint8)
    return np.ascontiguousarray(hwc[..., ::-1])


def _valid(p):
    return p[(p != -1).all(dim=-1)]


def _freqs(hit_masks):
    return torch.stack([m.float().mean() for m in hit_masks])


def _mean_offdiag_l1(freq):
    k = freq.shape[0]
    d = (freq[:, None] - freq[None, :]).abs()
    return (d.sum() / (k * (k - 1))).item()          # the diagonal is zero


def evaluate_metrics(probs_gender_all, probs_race_all, probs_age_all=None):
    g = _valid(probs_gender_all); r = _valid(probs_race_all)
    pg, pr = g.argmax(dim=-1), r.argmax(dim=-1)
    fg = _freqs([pg == q for q in range(2)])
    fr = _freqs([pr == q for q in range(4)])
    fgr = _freqs([(pg == a) * (pr == b) for a in range(2) for b in range(4)])
    out = [abs(fg[1] - fg[0]).item(), (g.max(dim=-1).values < 0.8).float().mean().item(), _mean_offdiag_l1(fr),
           (r.max(dim=-1).values < 0.8).float().mean().item(), _mean_offdiag_l1(fgr)]
    if probs_age_all is not None:
        a = _valid(probs_age_all)
        pa = a.argmax(dim=-1)
        a0, a1 = (pa == 0).float().mean().item(), (pa == 1).float().mean().item()
        out += [a0, a1, (a.max(dim=-1).values < 0.8).float().mean().item(), (abs(a0 - 0.75) + abs(a1 - 0.25)) / 2]
    return tuple(out)


def evaluate_metrics_race(probs_race_all):
    """exp-6-debias-race/1-main-debias.py:1624-1638: the four race frequencies, their mean pairwise gap and the share of
    faces whose top probability is below 0.8."""
    r = _valid(probs_race_all)
    pr = r.argmax(dim=-1)
    fr = _freqs([pr == q for q in range(4)])
    return tuple(f.item() for f in fr) + (_mean_offdiag_l1(fr), (r.max(dim=-1).values < 0.8).float().mean().item())


# ----------------------------------------------------------------------------- f4: consumer side
def adjusted_dft_grad_coefs(alphas_cumprod, alphas, timesteps):
    """E1:1104-1109 (inside generate_image_w_gradient)."""
    import math
    grad_coefs = []
    for t in timesteps:
        grad_coefs.append(alphas_cumprod[t].sqrt().item() * (1 - alphas_cumprod[t]).sqrt().item() / (1 - alphas[t].item()))
    grad_coefs = np.array(grad_coefs)
    grad_coefs /= (math.prod(grad_coefs) ** (1 / len(grad_coefs)))
    return grad_coefs


def allreduce_average_gradients(grads_per_rank, n_backward):
    """E1:1999-2011 for a simulated world: ``grads_per_rank[r][k]`` = rank r's gradient of parameter k.  Returns the list
    every rank ends up with and the per-rank ``grad_is_finite`` flags (checked BEFORE the all-reduce, like the reference)."""
    world = len(grads_per_rank)
    finite = [all(bool(torch.isfinite(g).all()) for g in grads) for grads in grads_per_rank]
    out = []
    for k in range(len(grads_per_rank[0])):
        s = grads_per_rank[0][k].clone()
        for r in range(1, world):
            s = s + grads_per_rank[r][k]                 # all_reduce(SUM)
        out.append(s / world / n_backward)
    return out, finite
