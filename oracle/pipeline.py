"""ORACLE (test infrastructure): the whole guidance pass on the CPU, stage by stage with the
restated reference functions (per-image Python loops and all).  Mirrors
finetune-fair-diffusion_b200/pipeline.py::GuidancePath.step so the two can be compared key by
key; also the body that bench.py times as the CPU baseline / ``--impl reference`` arm.
"""
import time

import numpy as np
import torch

from . import assign as oassign
from . import boxes as oboxes
from . import crop as ocrop
from . import head as ohead
from . import hooks as ohooks
from . import loss as oloss

_FN = {
    "gender": ohead.get_face_gender,
    "gender_race": ohead.get_face_gender_race,
    "gender_race_age": ohead.get_face_gender_race_age,
}


def step(batch, cfg, head, rand_tensors=None, world_rand=None, literal=True, emd=None, timings=None, backbone=None):
    """``batch``: host tensors from pipeline.synth_batch(host=True); ``head`` = (w1,b1,w2,b2) fp32 CPU.
    Single rank (``world_rand`` may simulate the other ranks' Monte-Carlo draws)."""
    def tick(name, t0):
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + (time.perf_counter() - t0)

    n_attr = cfg.n_attr
    e1_rule = cfg.kind == "gender"
    images = batch["images"].detach().float().clone().requires_grad_(True)
    n, _, H, W = images.shape

    t0 = time.perf_counter()
    ind_np, boxes_np = oboxes.select_and_expand(batch["cand_boxes"].numpy(), batch["counts"].numpy(), H, cfg.expand_coef, 1, -1)
    ind, boxes = torch.tensor(ind_np), torch.tensor(boxes_np)
    tick("boxes", t0)

    t0 = time.perf_counter()
    chips = ocrop.crop_faces(images, boxes, ind, cfg.size_face, cfg.fill_value)
    tick("crop_fwd", t0)

    t0 = time.perf_counter()
    w1, b1, w2, b2 = [t.float() for t in head]
    pooled = batch["pooled"].float().clone().requires_grad_(True)
    if backbone is not None:
        classifier = lambda x: ohead.mobilenet_head_reference(backbone(x), w1, b1, w2, b2)
    else:
        classifier = lambda x: ohead.mobilenet_head_reference(pooled[ind], w1, b1, w2, b2)
    outs = _FN[cfg.kind](classifier, chips, selector=ind, fill_value=-1)
    preds = [outs[3 * a] for a in range(n_attr)]
    probs = [outs[3 * a + 1] for a in range(n_attr)]
    logits = [outs[3 * a + 2] for a in range(n_attr)]
    tick("head_fwd", t0)

    t0 = time.perf_counter()
    det = [p.detach() for p in probs]
    if cfg.kind == "gender":
        t_all, u_all = oassign.generate_dynamic_targets(det[0], cfg.target_ratio, True)
        res = (t_all, u_all)
    elif cfg.kind == "gender_race":
        res = oassign.generate_dynamic_targets_gender_race(det[0], det[1], True, cfg.num_samples_per_device,
                                                           rand_tensors=rand_tensors, world_rand=world_rand, emd=emd, literal=literal)
    else:
        res = oassign.generate_dynamic_targets_gender_race_age(det[0], det[1], det[2], True, cfg.num_samples_per_device,
                                                               rand_tensors=rand_tensors, world_rand=world_rand, emd=emd, literal=literal)
    targets = []
    for a in range(n_attr):
        t, u = res[2 * a].clone(), res[2 * a + 1]
        t[u > cfg.uncertainty_threshold] = -1                      # E3:2022-2023
        targets.append(t)
    tick("assign", t0)

    t0 = time.perf_counter()
    loss_fair = [oloss.fairness_ce(logits[a], targets[a], ind, torch.float32) for a in range(n_attr)]
    bbox_ori = torch.where(ind.unsqueeze(1), boxes + batch["bbox_jitter"], boxes)
    f2, f1 = list(cfg.factors2[:n_attr]), list(cfg.factors1[:n_attr])
    hooked = ohooks.apply_grad_hook_face(images, boxes, bbox_ori, targets, batch["preds_ori"], f2, e1_rule)
    small = ocrop.resize_small(hooked, cfg.img_size_small)
    dyn_w = ohooks.gen_dynamic_weights(ind, targets, batch["preds_ori"], f1, torch.float32, e1_rule)
    loss = oloss.assemble(loss_fair, dyn_w, batch["loss_clip"].float(), batch["loss_dino"].float(), batch["loss_face"].float(),
                          cfg.weight_loss_img, cfg.weight_loss_face)
    tick("loss_fwd", t0)

    t0 = time.perf_counter()
    # backward: the classifier backbone and CLIP/DINO are external; their gradients w.r.t. the chips
    # and the resized images arrive as the stand-ins g_chips / g_small.
    surrogate = loss.mean() + (chips * batch["g_chips"].float()).sum() + (small * batch["g_small"].float()).sum()
    grads = torch.autograd.grad(surrogate, [images, pooled], allow_unused=True)
    g_images = grads[0]
    g_pooled = grads[1] if grads[1] is not None else torch.zeros_like(pooled)
    tick("backward", t0)

    return dict(indicators=ind, boxes=boxes, chips=chips.detach(), small=small.detach(), preds=preds, probs=det,
                logits=[l.detach() for l in logits], targets=targets, loss_fair=[l.detach() for l in loss_fair],
                dyn_weights=dyn_w, loss=loss.detach(), loss_mean=loss.mean().detach(), g_images=g_images.detach(),
                g_pooled=g_pooled.detach(), bbox_ori=bbox_ori)
