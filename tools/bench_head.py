#!/usr/bin/env python
"""Times fg_head_fwd / fg_head_bwd alone (CUDA events, L2 flushed between launches) for several inner sizes, to split the
kernel time into a per-launch part and a per-K-step part.  usage: bench_head.py [m] [dtype]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fairguide
m = int(sys.argv[1]) if len(sys.argv) > 1 else 965
dtype = getattr(torch, sys.argv[2]) if len(sys.argv) > 2 else torch.bfloat16
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=12):
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)
res = {}
for d_in, d_hid, k in ((64, 1280, 8), (320, 1280, 8), (960, 1280, 8), (960, 1280, 80), (960, 64, 8), (960, 640, 8)):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(m, d_in, generator=g).to(dtype).to(dev)
    w1 = (torch.randn(d_hid, d_in, generator=g) * 0.03).to(dtype).to(dev); b1 = torch.zeros(d_hid, dtype=dtype, device=dev)
    w2 = (torch.randn(k, d_hid, generator=g) * 0.03).to(dtype).to(dev); b2 = torch.zeros(k, dtype=dtype, device=dev)
    logits, hidden = fairguide.ops.head_fwd(x, w1, b1, w2, b2)
    gl = torch.randn(m, k, device=dev)
    res[f"in{d_in}_hid{d_hid}_k{k}"] = {"fwd_us": timeit(lambda: fairguide.ops.head_fwd(x, w1, b1, w2, b2)),
                                        "bwd_us": timeit(lambda: fairguide.ops.head_bwd(gl, hidden, w1, w2))}
print(json.dumps({"m": m, "dtype": str(dtype), **res}))
