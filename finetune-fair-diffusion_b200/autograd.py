"""torch.autograd.Function wrappers: each forward/backward is one call into the C ABI."""
import torch

from . import ops


class CropResize(torch.autograd.Function):
    """Fused crop_face (E1:267-290) for a batch + Resize(224) (E1:1905) + the backward of
    apply_grad_hook_face (E3:1751-1784) folded into the image gradient.

    forward(images, boxes, indicators, region, scale, chip_hw, small_hw, fill) -> (chips, small)
    backward: g_images = scale_in_region * resize_bwd(g_small) + crop_bwd(g_chips), one kernel."""

    @staticmethod
    def forward(ctx, images, boxes, indicators, region, scale, chip_hw, small_hw, fill_value):
        chips, small = ops.crop_resize_fwd(images, boxes, indicators, chip_hw, small_hw, fill_value)
        ctx.save_for_backward(boxes, indicators, region, scale)
        ctx.meta = (tuple(images.shape), images.dtype, images.device, chips is not None, small is not None)
        outs = tuple(o for o in (chips, small) if o is not None)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *grads):
        boxes, indicators, region, scale = ctx.saved_tensors
        shape, dtype, device, has_chips, has_small = ctx.meta
        grads = list(grads)
        g_chips = grads.pop(0) if has_chips else None
        g_small = grads.pop(0) if has_small else None
        if g_chips is None and g_small is None:
            return (None,) * 8
        g = ops.image_grad(g_chips, g_small, boxes, indicators, region, scale, shape, dtype, device)
        return (g,) + (None,) * 7


class RegionScale(torch.autograd.Function):
    """apply_grad_hook_face on its own: identity forward, region-scaled backward."""

    @staticmethod
    def forward(ctx, images, region, scale):
        ctx.save_for_backward(region, scale)
        return images.view_as(images)

    @staticmethod
    def backward(ctx, g):
        region, scale = ctx.saved_tensors
        return ops.region_scale(g, region, scale), None, None


class Head(torch.autograd.Function):
    """MobileNetV3 classifier head: Linear -> Hardswish -> Linear, frozen weights."""

    @staticmethod
    def forward(ctx, pooled, w1, b1, w2, b2):
        logits, hidden = ops.head_fwd(pooled, w1, b1, w2, b2)
        ctx.save_for_backward(hidden, w1, w2)
        return logits

    @staticmethod
    def backward(ctx, g_logits):
        hidden, w1, w2 = ctx.saved_tensors
        return ops.head_bwd(g_logits, hidden, w1, w2), None, None, None, None


class FairCE(torch.autograd.Function):
    """CE on face & target != -1, -1 elsewhere (E3:2114-2117)."""

    @staticmethod
    def forward(ctx, logits, targets, face_indicators, fill):
        ctx.save_for_backward(logits, targets, face_indicators)
        return ops.fair_ce_fwd(logits, targets, face_indicators, fill)

    @staticmethod
    def backward(ctx, g_loss):
        logits, targets, face = ctx.saved_tensors
        return ops.fair_ce_bwd(logits, targets, face, g_loss), None, None, None


class ScatterRows(torch.autograd.Function):
    """out[selector] = values with `fill` elsewhere, differentiable in `values` (used to keep the
    autograd link logits -> head when the head ran on the compacted rows)."""

    @staticmethod
    def forward(ctx, values, index, n, fill):
        out = torch.full((n,) + tuple(values.shape[1:]), fill, dtype=values.dtype, device=values.device)
        out[index] = values
        ctx.save_for_backward(index)
        return out

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        return g[index], None, None, None


class AlignedWarp(torch.autograd.Function):
    """image_pipeline (E1:292-312) for a batch: similarity warp of every image onto the 112x112 template."""

    @staticmethod
    def forward(ctx, images, params, indicators, dst_hw, fill_value):
        ctx.save_for_backward(params, indicators)
        ctx.image_shape = tuple(images.shape)
        return ops.aligned_warp_fwd(images, params, indicators, dst_hw, fill_value)

    @staticmethod
    def backward(ctx, g):
        params, indicators = ctx.saved_tensors
        return ops.aligned_warp_bwd(g, params, indicators, ctx.image_shape), None, None, None, None


class FeatsNormalize(torch.autograd.Function):
    """feats.to(float) -> F.normalize(dim=-1) (E1:1186-1189)."""

    @staticmethod
    def forward(ctx, raw):
        f, inv = ops.feats_normalize_fwd(raw)
        ctx.save_for_backward(f, inv)
        ctx.raw_dtype = raw.dtype
        return f

    @staticmethod
    def backward(ctx, g):
        f, inv = ctx.saved_tensors
        return ops.feats_normalize_bwd(g, f, inv, ctx.raw_dtype)


class FaceLoss(torch.autograd.Function):
    """loss_face (E1:1917-1929 and its E3 / E4 forms) from the raw summed features of the micro-batch."""

    @staticmethod
    def forward(ctx, raw_feats, feats_ori, db, face_indicators, confidence_level, search_needs_target, fill, n_attr, *attr):
        targets, preds, probs = attr[:n_attr], attr[n_attr:2 * n_attr], attr[2 * n_attr:]
        loss, ws = ops.face_loss_fwd(raw_feats, feats_ori, db, face_indicators, targets, preds, probs, confidence_level,
                                     search_needs_target, fill)
        ctx.save_for_backward(feats_ori, db, ws)
        ctx.meta = (raw_feats.shape[0], raw_feats.shape[1], raw_feats.dtype)
        ctx.n_inputs = 8 + 3 * n_attr
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        feats_ori, db, ws = ctx.saved_tensors
        n, d, dtype = ctx.meta
        return (ops.face_loss_bwd(g_loss, feats_ori, db, ws, n, d, dtype),) + (None,) * (ctx.n_inputs - 1)
