"""Helpers shared by the tests: golden-fixture loading and the procedural inputs that
tests/golden/make_golden.py used (same formulas, so the images need not be stored)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def procedural_image(seed, C, H, W):
    rng = np.random.Generator(np.random.PCG64(seed))
    y = np.arange(H, dtype=np.float64)[:, None]
    x = np.arange(W, dtype=np.float64)[None, :]
    chans = []
    for c in range(C):
        f = 0.55 * np.sin(0.031 * x + 0.017 * (c + 1) * y + 0.3 * seed) + 0.3 * np.cos(0.045 * y - 0.02 * x * (c + 1))
        chans.append(f)
    img = np.stack(chans) + rng.uniform(-0.1, 0.1, size=(C, H, W))
    return np.clip(img, -1, 1).astype(np.float32)


CROP_CASES_SMALL = [
    (1, 96, 96, (20, 24, 70, 74), 40),
    (2, 96, 96, (-10, -6, 50, 54), 40),
    (3, 96, 96, (50, 40, 110, 100), 40),
    (4, 96, 96, (-20, -20, 120, 120), 40),
    (5, 96, 96, (30, 30, 42, 42), 40),
    (6, 96, 96, (10, 10, 11, 60), 40),
    (7, 96, 80, (5, 9, 66, 70), 33),
]
CROP_CASES_FULL = [(21, 512, 512, (131, 97, 431, 397), 224), (22, 512, 512, (-37, 212, 339, 588), 224)]
CROP_CASES = CROP_CASES_SMALL + CROP_CASES_FULL


def peaked_probs(rng, n, k, sharp=4.0):
    z = rng.normal(size=(n, k)) * sharp
    z = z - z.max(axis=1, keepdims=True)
    p = np.exp(z)
    return (p / p.sum(axis=1, keepdims=True)).astype(np.float32)


def half_tensor(arr, dtype_name):
    """Decode a 16-bit fixture of half.npz (raw uint16 patterns; numpy has no bfloat16) into a torch tensor."""
    import torch
    if arr.dtype != np.uint16:
        return torch.tensor(arr)
    dt = {"f16": torch.float16, "bf16": torch.bfloat16}[dtype_name]
    return torch.tensor(arr.view(np.int16)).view(dt)


def half_equal(t, arr):
    """Bit-for-bit comparison of a 16-bit torch tensor with a raw-pattern fixture."""
    import torch
    got = t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    return got.shape == arr.shape and np.array_equal(got, arr)
