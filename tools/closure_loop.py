"""The reference's own step (exp-3 / exp-4 `1-main-debias.py`, "Step 1"-"Step 4": E3:1956-2147) written against the
closures ``fairguide.bind()`` returns -- the drop-in surface of INTEGRATION.md section 1 -- so that the swap a maintainer
would make is what gets measured (bench.py ``e2e`` and ``--surface closures``).

Third-party networks between our kernels enter as stand-ins, wired through autograd so that the loop is
the reference's:  the classifier backbone (torchvision MobileNetV3 ``features``, cuDNN) = a Function that returns the
stand-in pooled features and hands the stand-in chip gradient back; CLIP / DINO = a Function that returns the stand-in
losses and hands the stand-in resized-image gradient back.

``literal=True``  calls exactly the reference's sequence (crop, head, gather, assignment, threshold, slice; then per
                  micro-batch: crop with gradient, head, apply_grad_hook_face, Resize, CE, gen_dynamic_weights, backward).
``literal=False`` uses the batched / fused entry points INTEGRATION.md recommends on the same surface
                  (crop_and_resize: crop + Resize + hook backward in one forward and one backward kernel).
"""
import torch

KIND_FN = {"gender": ("get_face_gender", "generate_dynamic_targets"),
           "gender_race": ("get_face_gender_race", "generate_dynamic_targets_gender_race"),
           "gender_race_age": ("get_face_gender_race_age", "generate_dynamic_targets_gender_race_age")}
FACTOR_NAMES = {1: ["factor"], 2: ["factor_gender", "factor_race"], 3: ["factor_gender", "factor_race", "factor_age"]}


class _Backbone(torch.autograd.Function):
    """chips [m,3,224,224] -> pooled stand-in [m,960,1,1]; backward: the stand-in chip gradient (cuDNN stage excluded)."""

    @staticmethod
    def forward(ctx, chips, pooled, g_chips):
        ctx.save_for_backward(g_chips)
        return pooled.view(pooled.shape[0], pooled.shape[1], 1, 1)

    @staticmethod
    def backward(ctx, g):
        (g_chips,) = ctx.saved_tensors
        return g_chips, None, None


class _Semantics(torch.autograd.Function):
    """images_small -> (loss_CLIP, loss_DINO) stand-ins; backward: the stand-in gradient wrt the resized images."""

    @staticmethod
    def forward(ctx, small, loss_clip, loss_dino, g_small):
        ctx.save_for_backward(g_small)
        return loss_clip.clone(), loss_dino.clone()

    @staticmethod
    def backward(ctx, g1, g2):
        (g_small,) = ctx.saved_tensors
        return g_small, None, None, None


class StandInClassifier(torch.nn.Module):
    """MobileNetV3-shaped classifier whose ``features`` is the stand-in above; ``classifier`` holds the real head weights
    (Linear -> Hardswish -> Dropout -> Linear), which fairguide runs in its own kernel."""

    def __init__(self, head, dtype):
        super().__init__()
        w1, b1, w2, b2 = head
        l1 = torch.nn.Linear(w1.shape[1], w1.shape[0]); l2 = torch.nn.Linear(w2.shape[1], w2.shape[0])
        self.classifier = torch.nn.Sequential(l1, torch.nn.Hardswish(), torch.nn.Dropout(0.2), l2).to(w1.device, dtype)
        with torch.no_grad():
            l1.weight.copy_(w1); l1.bias.copy_(b1); l2.weight.copy_(w2); l2.bias.copy_(b2)
        self.avgpool = torch.nn.Identity()
        self.rows = None            # (pooled rows, chip-gradient rows) of the chips the next call will see
        self.requires_grad_(False)
        self.eval()

    def features(self, chips):
        pooled, g_chips = self.rows
        return _Backbone.apply(chips, pooled, g_chips)


class ClosureStep:
    def __init__(self, fairguide, cfg, head, dtype, micro_batch=None, literal=True):
        from fairguide import pipeline
        self.cfg, self.dtype, self.micro_batch, self.literal = cfg, dtype, micro_batch, literal
        self.widths = pipeline.KINDS[cfg.kind][0]
        self.clf = StandInClassifier(head, dtype)
        kw = {"gender": "gender_classifier", "gender_race": "gender_race_classifier", "gender_race_age": "gender_race_age_classifier"}[cfg.kind]
        self.fg = fairguide.bind(**{kw: self.clf})
        self.head_fn = getattr(self.fg, KIND_FN[cfg.kind][0])
        self.assign_fn = getattr(self.fg, KIND_FN[cfg.kind][1])

    def _heads(self, chips, ind, rows):
        self.clf.rows = rows
        out = self.head_fn(chips, selector=ind, fill_value=-1)
        A = len(self.widths)
        return [out[3 * a] for a in range(A)], [out[3 * a + 1] for a in range(A)], [out[3 * a + 2] for a in range(A)]

    def __call__(self, batch, num_valid=None):
        fg, cfg = self.fg, self.cfg
        A = len(self.widths)
        images = batch["images"]
        n = images.shape[0]
        # ---- Steps 1-2 (E3:1956-2025): no gradient -- faces, probabilities, gather, assignment, threshold, local slice
        with torch.no_grad():
            ind, boxes = fg.select_and_expand(batch["cand_boxes"], batch["counts"], images.shape[-2], cfg.expand_coef, 1, -1)
            chips = fg.crop_faces(images, boxes, ind, cfg.size_face, cfg.fill_value)
            _, probs, _ = self._heads(chips, ind, (batch["pooled"][ind], None))
            probs_all = [fg.customized_all_gather(p) for p in probs]
            if cfg.kind == "gender":
                res = self.assign_fn(probs_all[0], cfg.target_ratio, True)
            else:
                res = self.assign_fn(*probs_all, True, cfg.num_samples_per_device, num_valid=num_valid)
            rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
            targets = [fg.threshold_and_slice(res[2 * a], res[2 * a + 1], cfg.uncertainty_threshold, n, rank) for a in range(A)]
        # ---- Step 4 (E3:2077-2147): per micro-batch, with gradient
        mb = self.micro_batch or n
        names = FACTOR_NAMES[A]
        f2 = dict(zip(names, cfg.factors2[:A])); f1 = dict(zip(names, cfg.factors1[:A]))
        g_images = torch.empty_like(images)
        loss_all = torch.empty(n, dtype=torch.float32, device=images.device)
        n_backward = (n + mb - 1) // mb
        for j in range(n_backward):
            sl = slice(j * mb, min((j + 1) * mb, n))
            x = images[sl].detach().requires_grad_(True)                   # stands for generate_image_w_gradient(...)
            ind_j, boxes_j = ind[sl], boxes[sl]
            t_j = [t[sl] for t in targets]
            trip = []
            for a in range(A):
                trip += [t_j[a], batch["preds_ori"][a][sl], batch["probs_ori"][a][sl]]
            if self.literal:
                chips_j = fg.crop_faces(x, boxes_j, ind_j, cfg.size_face, cfg.fill_value)               # get_face(images_ij)
                _, _, logits = self._heads(chips_j, ind_j, (batch["pooled"][sl][ind_j], batch["g_chips"][sl][ind_j]))
                x_h = fg.apply_grad_hook_face(x, boxes_j, batch["bbox_ori"][sl], *trip, **f2)
                small = fg.resize_small(x_h, cfg.img_size_small)                                            # transforms.Resize
            else:
                H, W = x.shape[-2:]
                from fairguide import ops
                region, scale, _ = ops.guidance_factors(None, boxes_j, batch["bbox_ori"][sl], t_j, [p[sl] for p in batch["preds_ori"]],
                                                        cfg.factors2[:A], None, A == 1, H, W, want_weights=False)
                chips_j, small = fg.crop_and_resize(x, boxes_j, ind_j, region, scale, cfg.size_face, cfg.img_size_small, cfg.fill_value)
                _, _, logits = self._heads(chips_j, ind_j, (batch["pooled"][sl][ind_j], batch["g_chips"][sl][ind_j]))
            loss_clip, loss_dino = _Semantics.apply(small, batch["loss_clip"][sl], batch["loss_dino"][sl], batch["g_small"][sl])
            loss = None
            for a in range(A):
                l = fg.fairness_ce_loss(logits[a], t_j[a], ind_j)
                loss = l if loss is None else loss + l
            dyn_w = fg.gen_dynamic_weights(ind_j, *trip, **f1)
            loss = loss + cfg.weight_loss_img * dyn_w * (loss_clip + loss_dino) + cfg.weight_loss_face * batch["loss_face"][sl]
            loss.mean().backward()
            g_images[sl] = x.grad
            loss_all[sl] = loss.detach().float()
        return dict(targets=targets, loss=loss_all, loss_mean=loss_all.mean(), g_images=g_images, n_backward=n_backward)
