#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FUNCTION BODIES.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py [--ref /root/reference]

How: the reference scripts cannot be imported (diffusers, accelerate, insightface, POT, ...
are absent and the hot-path functions are closures inside ``main``).  This script parses
each ``1-main-debias.py`` with ``ast``, lifts the named ``FunctionDef`` nodes out UNMODIFIED,
compiles them, and runs them in a namespace that supplies their free variables (torch,
numpy, scipy, the classifier, ...).  Nothing from the reference is written into this repo;
only the functions' outputs on seeded inputs are stored.

Substitutions (the two third-party pieces that cannot run here; also listed in
oracle/__init__.py and DESIGN.md):
  * ``ot.emd``  -> exact assignment via scipy.optimize.linear_sum_assignment on the column-
    replicated cost matrix (POT 0.9.3 is not installed).  For the small cases the script
    asserts that scipy's HiGHS LP returns the same plan.
  * ``torchvision.transforms.Resize`` -> same class with ``antialias=False`` made the default,
    which is what the pinned torchvision 0.16.2 did for tensors (0.26 here defaults to True).
  * ``torch.rand`` is wrapped only to RECORD the tensors drawn, and
    ``torch.distributed.all_reduce`` is a two-pass stub that sums the simulated ranks.
"""
import argparse
import ast
import functools
import itertools
import math
import os
import sys
import types

import numpy as np
import scipy
import scipy.stats
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))

E1 = "exp-1-debias-gender/1-main-debias.py"
E3 = "exp-3-debias-gender-race/1-main-debias.py"
E4 = "exp-4-debias-gender-race-age/1-main-debias.py"


# --------------------------------------------------------------------------- lifting
def lift(path, names):
    tree = ast.parse(open(path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            found[node.name] = node
    missing = set(names) - set(found)
    assert not missing, f"{path}: missing {missing}"
    return found


class _Recorder:
    """Proxy for the ``torch`` module: forwards everything, records/replays torch.rand and
    exposes a stubbed ``distributed``."""

    def __init__(self):
        self.drawn = []
        self.replay = None
        self.distributed = types.SimpleNamespace(all_reduce=self._all_reduce, ReduceOp=types.SimpleNamespace(SUM="sum"))
        self.reduce_mode = "identity"
        self.captured = None
        self.reduced_value = None

    def __getattr__(self, name):
        return getattr(torch, name)

    def rand(self, *a, **k):
        if self.replay is not None:
            t = self.replay.pop(0)
        else:
            t = torch.rand(*a, **k)
        self.drawn.append(t.clone())
        return t

    class _Abort(Exception):
        pass

    def _all_reduce(self, tensor, op=None):
        if self.reduce_mode == "identity":
            return
        if self.reduce_mode == "capture":
            self.captured = tensor.clone()
            raise _Recorder._Abort()
        if self.reduce_mode == "inject":
            tensor.copy_(self.reduced_value)


def emd_stub(a, b, M):
    from scipy.optimize import linear_sum_assignment
    b = np.asarray(b, dtype=np.int64)
    M = np.asarray(M, dtype=np.float64)
    cols = np.repeat(np.arange(M.shape[1]), b)
    r, c = linear_sum_assignment(M[:, cols])
    T = np.zeros_like(M)
    T[r, cols[c]] = 1.0
    if M.shape[0] <= 48:
        from scipy.optimize import linprog
        N, K = M.shape
        A = np.zeros((N + K, N * K))
        for i in range(N):
            A[i, i * K:(i + 1) * K] = 1
        for j in range(K):
            A[N + j, j::K] = 1
        res = linprog(M.ravel(), A_eq=A, b_eq=np.concatenate([np.asarray(a, float), b.astype(float)]), bounds=(0, None), method="highs-ds")
        assert res.status == 0
        assert np.array_equal(np.rint(res.x.reshape(N, K)), T), "LSA and LP plans differ (tie in the costs?)"
    return T


def make_namespace(rec, extra=None):
    tv = types.SimpleNamespace(transforms=types.SimpleNamespace(
        Resize=functools.partial(torchvision.transforms.Resize, antialias=False),
        Pad=torchvision.transforms.Pad))
    ns = dict(torch=rec, torchvision=tv, transforms=tv.transforms, np=np, scipy=scipy, itertools=itertools, math=math,
              ot=types.SimpleNamespace(emd=emd_stub))
    if extra:
        ns.update(extra)
    return ns


def compile_into(ns, nodes):
    for node in nodes:
        mod = ast.Module(body=[node], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, "<reference>", "exec"), ns)


class LegacyF32:
    """A numpy float32 SCALAR as NumPy 1.26.4 (environment.yml:102, the reference's pin) treats it.

    NumPy 2 (NEP 50, installed here) keeps ``np.float32(x) * 0.5`` in float32; the pinned 1.26
    promotes scalar-with-Python-number arithmetic to float64 while float32-with-float32 stays
    float32.  expand_bbox / get_largest_face_app receive numpy float32 scalars from insightface,
    so the rounding of the box corners depends on this.  Feeding the reference functions these
    wrappers makes them compute what they compute in their own environment."""
    __array_ufunc__ = None
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = np.float32(v)

    @staticmethod
    def _bin(a, b, op):
        if isinstance(a, LegacyF32) and isinstance(b, LegacyF32):
            return LegacyF32(op(a.v, b.v))
        fa = np.float64(a.v) if isinstance(a, LegacyF32) else np.float64(a)
        fb = np.float64(b.v) if isinstance(b, LegacyF32) else np.float64(b)
        return op(fa, fb)

    def __sub__(self, o): return LegacyF32._bin(self, o, lambda x, y: x - y)
    def __rsub__(self, o): return LegacyF32._bin(o, self, lambda x, y: x - y)
    def __add__(self, o): return LegacyF32._bin(self, o, lambda x, y: x + y)
    def __radd__(self, o): return LegacyF32._bin(o, self, lambda x, y: x + y)
    def __mul__(self, o): return LegacyF32._bin(self, o, lambda x, y: x * y)
    def __rmul__(self, o): return LegacyF32._bin(o, self, lambda x, y: x * y)
    def __truediv__(self, o): return LegacyF32._bin(self, o, lambda x, y: x / y)
    def __rtruediv__(self, o): return LegacyF32._bin(o, self, lambda x, y: x / y)
    def _f(o): return float(o.v) if isinstance(o, LegacyF32) else float(o)
    def __lt__(self, o): return float(self.v) < LegacyF32._f(o)
    def __le__(self, o): return float(self.v) <= LegacyF32._f(o)
    def __gt__(self, o): return float(self.v) > LegacyF32._f(o)
    def __ge__(self, o): return float(self.v) >= LegacyF32._f(o)
    def __round__(self, nd=None): return int(np.rint(self.v))
    def __float__(self): return float(self.v)


class _LegacyNp:
    """``np`` as the MC-EMD functions see it: ``np.array(tensor)`` widens float32/float16 rows
    to float64.  Under the pinned NumPy 1.26 every expression those functions build from the rows
    (array minus int64 array at E3:1523, and scalar minus Python int at E4:1547-1557) is already
    promoted to float64; under NumPy 2 the scalar ones would stay float32.  Widening at the
    source reproduces the pinned behaviour exactly (float32 -> float64 is exact)."""

    def __getattr__(self, name):
        return getattr(np, name)

    def array(self, x, *a, **k):
        if torch.is_tensor(x) and x.dtype == torch.bfloat16:
            # EXTENSION beyond the reference: ``np.array(bf16_tensor)`` raises (numpy has no bfloat16), so the
            # reference's Monte-Carlo assignment only runs in fp16 / fp32.  Widening to fp32 (exact) is the one natural
            # continuation and is what the bf16 fixtures of gen_half() use.
            x = x.float()
        out = np.array(x, *a, **k)
        if out.dtype in (np.float32, np.float16):
            out = out.astype(np.float64)
        return out


def legacy_box(b):
    return [LegacyF32(x) for x in b]


# --------------------------------------------------------------------------- inputs
def procedural_image(seed, C, H, W):
    """Smooth field + small noise in [-1,1]; the tests rebuild it from the same formula."""
    rng = np.random.Generator(np.random.PCG64(seed))
    y = np.arange(H, dtype=np.float64)[:, None]
    x = np.arange(W, dtype=np.float64)[None, :]
    chans = []
    for c in range(C):
        f = 0.55 * np.sin(0.031 * x + 0.017 * (c + 1) * y + 0.3 * seed) + 0.3 * np.cos(0.045 * y - 0.02 * x * (c + 1))
        chans.append(f)
    img = np.stack(chans) + rng.uniform(-0.1, 0.1, size=(C, H, W))
    return np.clip(img, -1, 1).astype(np.float32)


def peaked_probs(rng, n, k, sharp=4.0):
    z = rng.normal(size=(n, k)) * sharp
    z = z - z.max(axis=1, keepdims=True)
    p = np.exp(z)
    return (p / p.sum(axis=1, keepdims=True)).astype(np.float32)


# --------------------------------------------------------------------------- generators
def gen_boxes(ref, out):
    fns = lift(os.path.join(ref, E1), ["expand_bbox", "get_largest_face_app"])
    ns = make_namespace(_Recorder())
    compile_into(ns, fns.values())
    rng = np.random.Generator(np.random.PCG64(11))
    n = 256
    ctr = rng.uniform(60, 452, size=(n, 2))
    wh = rng.uniform(20, 300, size=(n, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], axis=1).astype(np.float32)
    # rows that land exactly on .5 before rounding (half-to-even cases)
    boxes[:8] = np.array([[100, 100, 201, 201], [100, 100, 203, 203], [10.5, 20.5, 111.5, 121.5], [0, 0, 101, 51],
                          [3, 7, 12, 28], [250, 250, 261, 301], [400, 380, 505, 511], [1, 1, 4, 2]], dtype=np.float32)
    res = {}
    for tag, coef, ratio in (("c05_r1", 0.5, 1), ("c11_r1", 1.1, 1), ("c05_r12", 0.5, 1.2)):
        res["expanded_" + tag] = np.array([ns["expand_bbox"](legacy_box(b), expand_coef=coef, target_ratio=ratio) for b in boxes], dtype=np.int64)
        # what NumPy 2 would make of the same call (all-float32 arithmetic); stored to document the delta
        res["expanded_nep50_" + tag] = np.array([ns["expand_bbox"](b, expand_coef=coef, target_ratio=ratio) for b in boxes], dtype=np.int64)
    m, F = 96, 3
    multi = np.zeros((m, F, 4), dtype=np.float32)
    counts = rng.integers(1, F + 1, size=m)
    for i in range(m):
        c = rng.uniform(-40, 552, size=(F, 2))
        s = rng.uniform(10, 260, size=(F, 2))
        multi[i] = np.concatenate([c - s / 2, c + s / 2], axis=1)
    multi[0, :, :] = np.array([[600, 600, 700, 700], [-100, -100, -10, -10], [520, 0, 530, 40]], dtype=np.float32)  # all areas <= 0
    counts[0] = 3
    multi[1, 1] = multi[1, 0]   # exact tie -> first wins
    counts[1] = 2
    picked = []
    for i in range(m):
        faces = [{"bbox": legacy_box(multi[i, k])} for k in range(counts[i])]
        f = ns["get_largest_face_app"](faces, dim_max=512, dim_min=0)
        picked.append([k for k in range(counts[i]) if f is faces[k]][0])
    np.savez_compressed(os.path.join(out, "boxes.npz"), boxes=boxes, multi=multi, counts=counts.astype(np.int64),
                        picked=np.array(picked, dtype=np.int64), **res)


CROP_CASES_SMALL = [  # (seed, H, W, box, out)
    (1, 96, 96, (20, 24, 70, 74), 40),      # inside, downscale
    (2, 96, 96, (-10, -6, 50, 54), 40),     # pad left/top
    (3, 96, 96, (50, 40, 110, 100), 40),    # pad right/bottom
    (4, 96, 96, (-20, -20, 120, 120), 40),  # pad all sides
    (5, 96, 96, (30, 30, 42, 42), 40),      # upscale
    (6, 96, 96, (10, 10, 11, 60), 40),      # one pixel wide
    (7, 96, 80, (5, 9, 66, 70), 33),        # non-square image, odd output
]
CROP_CASES_FULL = [(21, 512, 512, (131, 97, 431, 397), 224), (22, 512, 512, (-37, 212, 339, 588), 224)]


def gen_crop(ref, out):
    fns = lift(os.path.join(ref, E1), ["crop_face"])
    ns = make_namespace(_Recorder())
    compile_into(ns, fns.values())
    res = {}
    for idx, (seed, H, W, box, o) in enumerate(CROP_CASES_SMALL + CROP_CASES_FULL):
        img = torch.tensor(procedural_image(seed, 3, H, W), requires_grad=True)
        chip = ns["crop_face"](img, list(box), [o, o], -1)
        g = torch.tensor(procedural_image(seed + 100, 3, o, o))
        (chip * g).sum().backward()
        res[f"chip_{idx}"] = chip.detach().numpy()
        grad = img.grad.numpy()
        if H <= 128:
            res[f"grad_{idx}"] = grad
        else:
            res[f"gradwin_{idx}"] = grad[:, 192:256, 128:192].copy()
            res[f"gradsum_{idx}"] = np.array([grad.astype(np.float64).sum(), np.abs(grad).astype(np.float64).sum()])
    # whole-image Resize(224) as called at E1:1905
    imgs = torch.tensor(procedural_image(31, 3, 512, 512)[None], requires_grad=True)
    small = ns["transforms"].Resize(224)(imgs)
    g = torch.tensor(procedural_image(131, 3, 224, 224)[None])
    (small * g).sum().backward()
    res["small"] = small.detach().numpy()[0]
    res["small_gradwin"] = imgs.grad.numpy()[0][:, 192:256, 128:192].copy()
    np.savez_compressed(os.path.join(out, "crop.npz"), **res)


def gen_heads(ref, out):
    rng = np.random.Generator(np.random.PCG64(5))
    res = {}
    n = 23
    selector = rng.uniform(size=n) > 0.25
    selector[3] = False
    m = int(selector.sum())
    chips = torch.zeros(n, 1)   # the stub classifier ignores pixel values
    for tag, path, fn, kh in (("e1", E1, "get_face_gender", 80), ("e3", E3, "get_face_gender_race", 6),
                              ("e4", E4, "get_face_gender_race_age", 8)):
        logits = (rng.normal(size=(m, kh)) * 2.0).astype(np.float32)
        full = (rng.normal(size=(n, kh)) * 2.0).astype(np.float32)
        state = {"next": None}
        clf = lambda x, s=state: torch.tensor(s["next"])
        ns = make_namespace(_Recorder(), {"gender_classifier": clf, "gender_race_classifier": clf,
                                          "gender_race_age_classifier": clf})
        compile_into(ns, lift(os.path.join(ref, path), [fn]).values())
        state["next"] = logits
        outs = ns[fn](chips, selector=torch.tensor(selector), fill_value=-1)
        state["next"] = full
        outs_nosel = ns[fn](chips, selector=None, fill_value=-1)
        empty = ns[fn](chips, selector=torch.zeros(n, dtype=torch.bool), fill_value=-1)
        res[f"{tag}_logits_in"] = logits
        res[f"{tag}_logits_in_full"] = full
        for k, o in enumerate(outs):
            res[f"{tag}_sel_{k}"] = o.numpy()
        for k, o in enumerate(outs_nosel):
            res[f"{tag}_nosel_{k}"] = o.numpy()
        for k, o in enumerate(empty):
            res[f"{tag}_empty_{k}"] = o.numpy()
    np.savez_compressed(os.path.join(out, "heads.npz"), selector=selector, **res)


def _with_missing(p, rng, frac):
    p = p.copy()
    miss = rng.uniform(size=p.shape[0]) < frac
    p[miss] = -1
    return p, miss


def gen_assign_e1(ref, out):
    rec = _Recorder()
    ns = make_namespace(rec)
    compile_into(ns, lift(os.path.join(ref, E1), ["generate_dynamic_targets"]).values())
    rng = np.random.Generator(np.random.PCG64(7))
    res = {}
    cases = [(64, 0.1, 0.5), (37, 0.2, 0.5), (200, 0.05, 0.5), (41, 0.0, 0.3), (5, 1.0, 0.5), (1, 0.0, 0.5)]
    for c, (n, frac, ratio) in enumerate(cases):
        p, miss = _with_missing(peaked_probs(rng, n, 2, 1.5), rng, frac)
        t, u = ns["generate_dynamic_targets"](torch.tensor(p), target_ratio=ratio, w_uncertainty=True)
        t_only = ns["generate_dynamic_targets"](torch.tensor(p), target_ratio=ratio, w_uncertainty=False)
        assert torch.equal(t, t_only)
        res[f"probs_{c}"] = p
        res[f"ratio_{c}"] = np.array(ratio)
        res[f"targets_{c}"] = t.numpy()
        res[f"unc_{c}"] = u.numpy()
    res["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(out, "assign_e1.npz"), **res)


def _run_world(ns, fn, rec, probs, world, S):
    """Execute the reference function once per simulated rank.  Pass 1 captures each rank's
    pre-all-reduce plan sum (and its draws); pass 2 replays the draws and injects the sum."""
    draws, partial = [], []
    for r in range(world):
        rec.drawn, rec.replay, rec.reduce_mode = [], None, "capture"
        try:
            ns[fn](*probs, w_uncertainty=True, num_samples_per_device=S)
            # N == 0 returns before the all-reduce
            rec.reduce_mode = "identity"
            return [], ns[fn](*probs, w_uncertainty=True, num_samples_per_device=S)
        except _Recorder._Abort:
            pass
        draws.append([d.clone() for d in rec.drawn])
        partial.append(rec.captured)
    total = partial[0].clone()
    for p in partial[1:]:
        total += p
    outs_per_rank = []
    for r in range(world):
        rec.drawn, rec.replay, rec.reduce_mode, rec.reduced_value = [], [d.clone() for d in draws[r]], "inject", total
        outs_per_rank.append(ns[fn](*probs, w_uncertainty=True, num_samples_per_device=S))
    for o in outs_per_rank[1:]:
        for a, b in zip(o, outs_per_rank[0]):
            assert torch.equal(a, b)
    return draws, outs_per_rank[0]


def gen_assign_mc(ref, out):
    rng = np.random.Generator(np.random.PCG64(9))
    torch.manual_seed(5991)
    for tag, path, fn, widths in (("e3", E3, "generate_dynamic_targets_gender_race", (2, 4)),
                                  ("e4", E4, "generate_dynamic_targets_gender_race_age", (2, 4, 2))):
        rec = _Recorder()
        ns = make_namespace(rec, {"np": _LegacyNp()})
        compile_into(ns, lift(os.path.join(ref, path), [fn]).values())
        res = {}
        cases = [(40, 0.0, 1, 100), (24, 0.2, 1, 100), (48, 0.1, 2, 50), (33, 0.1, 4, 25), (6, 1.0, 1, 10), (1, 0.0, 1, 10),
                 (30, 0.0, 1, 100)]
        for c, (n, frac, world, S) in enumerate(cases):
            sharp = 0.6 if c == 6 else 2.5     # case 6: flat probabilities -> many uncertain rows
            miss = rng.uniform(size=n) < frac
            probs = []
            for w in widths:
                p = peaked_probs(rng, n, w, sharp)
                p[miss] = -1
                probs.append(torch.tensor(p))
            draws, outs = _run_world(ns, fn, rec, probs, world, S)
            for k, p in enumerate(probs):
                res[f"probs{k}_{c}"] = p.numpy()
            res[f"world_{c}"] = np.array(world)
            res[f"S_{c}"] = np.array(S)
            for r, dr in enumerate(draws):
                for k, d in enumerate(dr):
                    res[f"rand{k}_r{r}_{c}"] = d.numpy()
            for k, o in enumerate(outs):
                res[f"out{k}_{c}"] = o.numpy()
        res["n_cases"] = np.array(len(cases))
        np.savez_compressed(os.path.join(out, f"assign_{tag}.npz"), **res)


def gen_hooks(ref, out):
    rng = np.random.Generator(np.random.PCG64(13))
    b, H = 10, 32
    imgs = np.stack([procedural_image(200 + i, 3, H, H) for i in range(b)])
    up = np.stack([procedural_image(300 + i, 3, H, H) for i in range(b)])
    box = np.zeros((b, 4), dtype=np.int64)
    box_ori = np.zeros((b, 4), dtype=np.int64)
    for i in range(b):
        c = rng.integers(8, 24, size=2)
        s = rng.integers(6, 30)
        box[i] = [c[0] - s // 2, c[1] - s // 2, c[0] + s // 2, c[1] + s // 2]
        box_ori[i] = box[i] + rng.integers(-4, 5, size=4)
    box[2] = -1                      # no face now
    box_ori[4] = -1                  # no face originally: slice-end quirk
    box_ori[6] = [28, 28, 31, 31]    # disjoint from the new box -> empty region
    box[6] = [2, 2, 12, 12]
    box[7] = [-5, -5, 40, 40]        # larger than the image
    res = dict(images=imgs, upstream=up, box=box, box_ori=box_ori)
    n_t = {"e1": 1, "e3": 2, "e4": 3}
    kmax = [2, 4, 2]
    factors2 = {"e1": [0.1], "e3": [0.2, 0.3], "e4": [0.2, 0.3, 0.35]}
    factors1 = {"e1": [0.2], "e3": [0.2, 0.6], "e4": [0.2, 0.6, 0.5]}
    for tag, path in (("e1", E1), ("e3", E3), ("e4", E4)):
        ns = make_namespace(_Recorder())
        compile_into(ns, lift(os.path.join(ref, path), ["make_grad_hook", "apply_grad_hook_face", "gen_dynamic_weights"]).values())
        A = n_t[tag]
        targets = [rng.integers(-1, kmax[a], size=b) for a in range(A)]
        preds = [rng.integers(0, kmax[a], size=b) for a in range(A)]
        for a in range(A):
            preds[a][4] = -1          # original image had no face
            targets[a][0] = preds[a][0]   # a full match
        probs = [np.full((b, kmax[a]), 0.5, dtype=np.float32) for a in range(A)]
        face_ind = ~(box == -1).all(axis=1)
        x = torch.tensor(imgs, requires_grad=True)
        args = [x, torch.tensor(box), torch.tensor(box_ori)]
        wargs = [torch.tensor(face_ind)]
        for a in range(A):
            args += [torch.tensor(targets[a]), torch.tensor(preds[a]), torch.tensor(probs[a])]
            wargs += [torch.tensor(targets[a]), torch.tensor(preds[a]), torch.tensor(probs[a])]
        if tag == "e1":
            y = ns["apply_grad_hook_face"](*args, factor=factors2[tag][0])
            w = ns["gen_dynamic_weights"](*wargs, factor=factors1[tag][0])
        elif tag == "e3":
            y = ns["apply_grad_hook_face"](*args, factor_gender=factors2[tag][0], factor_race=factors2[tag][1])
            w = ns["gen_dynamic_weights"](*wargs, factor_gender=factors1[tag][0], factor_race=factors1[tag][1])
        else:
            y = ns["apply_grad_hook_face"](*args, factor_gender=factors2[tag][0], factor_race=factors2[tag][1], factor_age=factors2[tag][2])
            w = ns["gen_dynamic_weights"](*wargs, factor_gender=factors1[tag][0], factor_race=factors1[tag][1], factor_age=factors1[tag][2])
        (y * torch.tensor(up)).sum().backward()
        res[f"{tag}_forward"] = y.detach().numpy()
        res[f"{tag}_grad"] = x.grad.numpy()
        res[f"{tag}_weights"] = w.numpy()
        res[f"{tag}_factors2"] = np.array(factors2[tag])
        res[f"{tag}_factors1"] = np.array(factors1[tag])
        for a in range(A):
            res[f"{tag}_targets{a}"] = targets[a].astype(np.int64)
            res[f"{tag}_preds{a}"] = preds[a].astype(np.int64)
    np.savez_compressed(os.path.join(out, "hooks.npz"), **res)


HALF_DTYPES = (("f16", torch.float16), ("bf16", torch.bfloat16))
THRESHOLD = 0.2          # every shipped YAML (e.g. exp-3-debias-gender-race/configs/debias-text-encoder.yaml:10)


def _bits(t):
    """16-bit tensors are stored as their raw uint16 patterns (numpy has no bfloat16)."""
    return t.contiguous().view(torch.int16).numpy().view(np.uint16) if t.dtype in (torch.float16, torch.bfloat16) else t.numpy()


def gen_half(ref, out):
    """The heads, the three assignments and the caller's thresholding statement (E3:2022-2023) on the 16-bit
    probability dtypes: fp16 is what the reference itself runs (weight_dtype, E1:933), bf16 is the dtype of the
    BASELINE headline config.  World sizes are simulated as in gen_assign_mc; cases keep world * S within the
    integers the dtype represents exactly (fp16: 2048, bf16: 256), because beyond that the reference's own
    all-reduce(SUM) in that dtype is order-dependent (NCCL ring vs tree) and has no single answer."""
    rng = np.random.Generator(np.random.PCG64(21))
    torch.manual_seed(5991)
    res = {}
    # ---- heads: softmax / argmax / scatter in the 16-bit dtype
    n = 29
    selector = rng.uniform(size=n) > 0.25
    m = int(selector.sum())
    chips = torch.zeros(n, 1)
    for dn, dt in HALF_DTYPES:
        for tag, path, fn, kh in (("e1", E1, "get_face_gender", 80), ("e3", E3, "get_face_gender_race", 6),
                                  ("e4", E4, "get_face_gender_race_age", 8)):
            logits = torch.tensor((rng.normal(size=(m, kh)) * 2.5).astype(np.float32)).to(dt)
            state = {"next": logits}
            clf = lambda x, s=state: s["next"].clone()
            ns = make_namespace(_Recorder(), {"gender_classifier": clf, "gender_race_classifier": clf,
                                              "gender_race_age_classifier": clf})
            compile_into(ns, lift(os.path.join(ref, path), [fn]).values())
            outs = ns[fn](chips, selector=torch.tensor(selector), fill_value=-1)
            res[f"heads_{dn}_{tag}_logits"] = _bits(logits)
            for k, o in enumerate(outs):
                res[f"heads_{dn}_{tag}_out{k}"] = _bits(o)
    res["heads_selector"] = selector
    # ---- E1 rank / binomial assignment (tie-free second column: argsort is unstable)
    ns = make_namespace(_Recorder())
    compile_into(ns, lift(os.path.join(ref, E1), ["generate_dynamic_targets"]).values())
    for dn, dt in HALF_DTYPES:
        for c, (n, frac, ratio) in enumerate([(64, 0.1, 0.5), (150, 0.05, 0.5), (41, 0.0, 0.3)]):
            lo, hi = (0.05, 0.95)
            vals = torch.linspace(lo, hi, 4 * n).to(dt).unique()          # distinct 16-bit values
            pick = torch.tensor(rng.permutation(vals.numel())[:n])
            p1 = vals[pick]
            p = torch.stack([(1 - p1.float()).to(dt), p1], dim=1)
            miss = torch.tensor(rng.uniform(size=n) < frac)
            p[miss] = -1
            t, u = ns["generate_dynamic_targets"](p, target_ratio=ratio, w_uncertainty=True)
            t_thr = t.clone(); t_thr[u > THRESHOLD] = -1                     # E1:1835
            res[f"e1_{dn}_probs_{c}"] = _bits(p); res[f"e1_{dn}_ratio_{c}"] = np.array(ratio)
            res[f"e1_{dn}_targets_{c}"] = t.numpy(); res[f"e1_{dn}_unc_{c}"] = _bits(u); res[f"e1_{dn}_thr_{c}"] = t_thr.numpy()
        res[f"e1_{dn}_n_cases"] = np.array(3)
    # ---- E3 / E4 Monte-Carlo assignment
    cases = {"f16": [(48, 0.1, 1, 100, 2.5), (40, 0.0, 2, 100, 2.5), (36, 0.1, 8, 100, 2.5), (30, 0.0, 1, 100, 0.6), (44, 0.0, 4, 100, 1.2)],
             "bf16": [(48, 0.1, 1, 100, 2.5), (40, 0.0, 2, 100, 2.5), (32, 0.0, 4, 60, 2.5), (30, 0.0, 1, 100, 0.6), (44, 0.1, 2, 128, 1.2)]}
    for tag, path, fn, widths in (("e3", E3, "generate_dynamic_targets_gender_race", (2, 4)),
                                  ("e4", E4, "generate_dynamic_targets_gender_race_age", (2, 4, 2))):
        rec = _Recorder()
        ns = make_namespace(rec, {"np": _LegacyNp()})
        compile_into(ns, lift(os.path.join(ref, path), [fn]).values())
        for dn, dt in HALF_DTYPES:
            for c, (n, frac, world, S, sharp) in enumerate(cases[dn]):
                miss = rng.uniform(size=n) < frac
                probs = []
                for w in widths:
                    p = torch.tensor(peaked_probs(rng, n, w, sharp)).to(dt)
                    p[torch.tensor(miss)] = -1
                    probs.append(p)
                draws, outs = _run_world(ns, fn, rec, probs, world, S)
                key = f"{tag}_{dn}"
                for k, p in enumerate(probs):
                    res[f"{key}_probs{k}_{c}"] = _bits(p)
                res[f"{key}_world_{c}"] = np.array(world); res[f"{key}_S_{c}"] = np.array(S)
                for r, dr in enumerate(draws):
                    for k, d in enumerate(dr):
                        res[f"{key}_rand{k}_r{r}_{c}"] = _bits(d)
                for k, o in enumerate(outs):
                    res[f"{key}_out{k}_{c}"] = _bits(o)
                for a in range(len(widths)):                               # E3:2022-2023 / E4:2129-2131
                    t = outs[2 * a].clone(); t[outs[2 * a + 1] > THRESHOLD] = -1
                    res[f"{key}_thr{a}_{c}"] = t.numpy()
            res[f"{tag}_{dn}_n_cases"] = np.array(len(cases[dn]))
    np.savez_compressed(os.path.join(out, "half.npz"), **res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        sys.exit(f"reference tree not found at {a.ref}")
    gen_boxes(a.ref, a.out)
    gen_crop(a.ref, a.out)
    gen_heads(a.ref, a.out)
    gen_assign_e1(a.ref, a.out)
    gen_assign_mc(a.ref, a.out)
    gen_hooks(a.ref, a.out)
    gen_half(a.ref, a.out)
    with open(os.path.join(a.out, "VERSIONS.txt"), "w") as f:
        f.write(f"torch {torch.__version__}\ntorchvision {torchvision.__version__}\nnumpy {np.__version__}\nscipy {scipy.__version__}\n")
    print("golden fixtures written to", a.out)


if __name__ == "__main__":
    main()
