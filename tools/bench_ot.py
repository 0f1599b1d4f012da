#!/usr/bin/env python
"""Times fg_ot_plan_counts (all kernels of the Monte-Carlo assignment before the all-reduce) and prints the
repair-step counters.  usage: bench_ot.py N K [S] [dtype]   (dtype bfloat16: probabilities with many exactly equal costs)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fairguide
N, K = int(sys.argv[1]), int(sys.argv[2]); S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dt = getattr(torch, sys.argv[4]) if len(sys.argv) > 4 else torch.float32
dev = "cuda"
g = torch.Generator().manual_seed(0)
def probs(w, sharp):
    z = torch.randn(N, w, generator=g) * sharp
    return torch.softmax(z, -1).to(dt).to(dev)
def head_probs():
    """Probabilities as the bench step produces them: the pipeline's random-init head on random pooled features -- peaked
    (logit std ~4) and class-biased (the hidden layer has a non-zero mean), in `dt`."""
    from fairguide import pipeline
    cfg = pipeline.GuidanceConfig(kind="gender_race_age" if K == 16 else "gender_race")
    widths, col_start, k_head = pipeline.KINDS[cfg.kind][0], pipeline.KINDS[cfg.kind][1], pipeline.KINDS[cfg.kind][2]
    head = pipeline.make_head_weights(cfg, dt, dev)
    pooled = torch.randn(N, cfg.d_in, generator=g).to(dt).to(dev)
    logits, _ = fairguide.ops.head_fwd(pooled, *head)
    ind = torch.ones(N, dtype=torch.bool, device=dev)
    _, pr_, _ = fairguide.ops.head_attributes(logits, None, ind, N, col_start, widths, -1.0, dt)
    return pr_


res = {}
for sharp in (2.0, 0.7, "head"):
    if sharp == "head":
        pp = head_probs(); pg, pr = pp[0], pp[1]; pa = pp[2] if K == 16 else None
    else:
        pg, pr = probs(2, sharp), probs(4, sharp); pa = probs(2, sharp) if K == 16 else None
    r = tuple(torch.rand(S, N, generator=g).to(dt).to(dev) for _ in range(3 if K == 16 else 2))
    ws = fairguide.ops.OtWorkspace(N, K, S, dev)
    ts = []
    for it in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); c = fairguide.ops.ot_plan_counts(pg, pr, pa, r, N, ws); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    st = ws.status()
    res[f"sharp{sharp}"] = {"ms": round(sorted(ts)[len(ts) // 2], 4), "status": st[0], "base_repairs": st[2], "draw_repairs_per_draw": round(st[3] / S, 2)}
print(json.dumps({"N": N, "K": K, "S": S, "dtype": str(dt), "env": {k: v for k, v in os.environ.items() if k.startswith("FG_OT")}, **res}))
