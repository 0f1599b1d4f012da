"""ORACLE (test infrastructure): 224x224 face crop and whole-image resize.

Restates crop_face (E1:267-290) with torch CPU ops, and, as an independent cross-check, a
direct clamped-bilinear sampler written from the ATen upsample_bilinear2d definition
(align_corners=False): ``src = max(scale*(dst+0.5)-0.5, 0)``, upper neighbour clamped to the
last row/column of the PADDED box, weights ``1-l, l``.
"""
import torch
import torch.nn.functional as F


def _resize_noaa(img, size):
    # torchvision 0.16.2 Resize on tensors == bilinear interpolate, align_corners=False,
    # antialias off (see oracle/__init__.py).  Works on [C,H,W] like the transform does.
    squeeze = img.dim() == 3
    x = img.unsqueeze(0) if squeeze else img
    y = F.interpolate(x, size=list(size), mode="bilinear", align_corners=False, antialias=False)
    return y.squeeze(0) if squeeze else y


def crop_face(img, box, target_size=(224, 224), fill_value=-1):
    """E1:267-290: clip box to the image, slice, pad the clipped part back with ``fill_value``,
    bilinear resize to ``target_size``.  ``img`` is [3,H,W]; gradient flows to ``img``."""
    H, W = img.shape[-2:]
    x0, y0, x1, y1 = int(box[0]), int(box[1]), int(box[2]), int(box[3])
    left, right = max(x0, 0), min(x1, W)
    top, bottom = max(y0, 0), min(y1, H)
    pl, pr = max(-x0, 0), max(x1 - W, 0)
    pt, pb = max(-y0, 0), max(y1 - H, 0)
    face = img[:, top:bottom, left:right]
    if pl > 0 or pt > 0 or pr > 0 or pb > 0:
        face = F.pad(face, [pl, pr, pt, pb], mode="constant", value=fill_value)
    return _resize_noaa(face, target_size)


def resize_small(images, size=224):
    """transforms.Resize(224) on a square batch (E1:1860, E1:1905)."""
    return _resize_noaa(images, (size, size))


def crop_faces(images, boxes, indicators, size=224, fill_value=-1):
    """Per-image loop of get_face_app (E1:1324-1345) reduced to the crop: rows without a
    face become all-``fill_value`` chips (E1:1328-1332)."""
    chips = []
    for i in range(images.shape[0]):
        if not bool(indicators[i]):
            chips.append(torch.ones([1, images.shape[1], size, size], dtype=images.dtype) * fill_value)
        else:
            chips.append(crop_face(images[i], boxes[i].tolist(), [size, size], fill_value).unsqueeze(0))
    return torch.cat(chips, dim=0)


def direct_sampler(img, box, out_hw=(224, 224), fill_value=-1.0):
    """Independent formulation: sample the virtual padded box directly (no slice/pad/resize).
    float64 interpolation of the float values; used only to cross-check ``crop_face``."""
    C, H, W = img.shape
    x0, y0, x1, y1 = [int(v) for v in box]
    bw, bh = x1 - x0, y1 - y0
    oh, ow = out_hw
    out = torch.empty((C, oh, ow), dtype=torch.float64)
    src = img.to(torch.float64)

    def axis(n_in, n_out):
        scale = torch.tensor(n_in, dtype=torch.float32) / torch.tensor(n_out, dtype=torch.float32)
        d = torch.arange(n_out, dtype=torch.float32)
        s = torch.clamp(scale * (d + 0.5) - 0.5, min=0.0)
        i0 = s.to(torch.int64)
        i1 = torch.clamp(i0 + 1, max=n_in - 1)
        lam = (s - i0.to(torch.float32)).to(torch.float64)
        return i0, i1, lam

    iy0, iy1, ly = axis(bh, oh)
    ix0, ix1, lx = axis(bw, ow)

    def fetch(py, px):
        yy = (y0 + py)[:, None].expand(oh, ow)
        xx = (x0 + px)[None, :].expand(oh, ow)
        inside = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = src[:, yy.clamp(0, H - 1), xx.clamp(0, W - 1)]
        return torch.where(inside[None], v, torch.tensor(float(fill_value), dtype=torch.float64))

    a, b = fetch(iy0, ix0), fetch(iy0, ix1)
    c, d = fetch(iy1, ix0), fetch(iy1, ix1)
    wx, wy = lx[None, None, :], ly[None, :, None]
    out = (1 - wy) * ((1 - wx) * a + wx * b) + wy * ((1 - wx) * c + wx * d)
    return out
