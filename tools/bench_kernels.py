#!/usr/bin/env python
"""Micro-benchmark of the two HBM-bound kernels (fused crop+resize forward, fused image gradient) at the
BASELINE shape: CUDA events around single launches, inputs larger than L2, GB/s against algorithmic bytes.
usage: bench_kernels.py [n_images] [dtype]      (FG_LIB=<variant .so> selects a tuning build)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fairguide
from fairguide import pipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dtype = getattr(torch, sys.argv[2]) if len(sys.argv) > 2 else torch.bfloat16
dev = "cuda"
cfg = pipeline.GuidanceConfig(kind="gender_race_age")
b = pipeline.synth_batch_device(n, cfg, dtype, dev)
ind, boxes = fairguide.ops.select_expand_boxes(b["cand_boxes"], b["counts"], 512)
region = torch.stack([boxes[:, 0].clamp(0, 512), boxes[:, 1].clamp(0, 512), boxes[:, 2].clamp(0, 512), boxes[:, 3].clamp(0, 512)], 1).to(torch.int32)
scale = torch.full((n,), 0.3, device=dev)
e = torch.empty((), dtype=dtype).element_size()
fwd_b = (3 * 512 * 512 + 2 * 3 * 224 * 224) * e * n
bwd_b = (2 * 3 * 224 * 224 + 3 * 512 * 512) * e * n

def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    ts.sort()
    return ts[len(ts) // 2]

f = timeit(lambda: fairguide.ops.crop_resize_fwd(b["images"], boxes, ind, (224, 224), (224, 224), -1.0))
g = timeit(lambda: fairguide.ops.image_grad(b["g_chips"], b["g_small"], boxes, ind, region, scale, tuple(b["images"].shape), dtype, b["images"].device))
print(json.dumps({"lib": os.path.basename(fairguide._lib.LIB_PATH), "n": n, "dtype": str(dtype), "fwd_ms": round(f, 4), "fwd_GBps": round(fwd_b / f / 1e6, 1),
                  "bwd_ms": round(g, 4), "bwd_GBps": round(bwd_b / g / 1e6, 1)}))
