"""ORACLE (test infrastructure): fairness cross-entropy and loss assembly.

Restates the inline loss code of the training loop: E1:1912-1933, E3:2114-2147,
E4:2238-2283.  ``CE_loss`` is ``nn.CrossEntropyLoss(reduction="none")`` (E1:968).
"""
import torch


def fairness_ce(logits, targets, face_indicators, dtype=None):
    """Per-image CE on rows with a face and a kept target, -1 elsewhere (E3:2114-2117)."""
    n = targets.shape[0]
    dtype = dtype or logits.dtype
    loss = torch.ones(n, dtype=dtype, device=logits.device) * (-1)
    idx = ((face_indicators == True) * (targets != -1)).nonzero().view([-1])  # noqa: E712
    ce = torch.nn.functional.cross_entropy(logits[idx], targets[idx], reduction="none")
    loss[idx] = ce.to(dtype)
    return loss


def assemble(loss_fair_terms, dynamic_weights, loss_clip, loss_dino, loss_face, weight_loss_img, weight_loss_face):
    """``loss_ij`` of E3:2146 / E4:2282 (E1:1932 with a single fairness term); the caller
    backpropagates ``loss_ij.mean()`` (E3:2147)."""
    total = loss_fair_terms[0]
    for t in loss_fair_terms[1:]:
        total = total + t
    return total + weight_loss_img * dynamic_weights * (loss_clip + loss_dino) + weight_loss_face * loss_face


def cosine_loss(feats, feats_ori):
    """``- (f * f_ori).sum(-1) + 1`` (E1:1909-1910)."""
    return -(feats * feats_ori).sum(dim=-1) + 1
