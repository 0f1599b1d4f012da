#!/usr/bin/env python
"""Times the kernels of the "next" rows at BASELINE-like sizes: CUDA events around single calls, median of
10 after warm-up.  f1: aligned 112x112 warp forward / backward on 1024 bf16 and fp32 images, face search over a
100k x 512 database, fused face loss; f3: E6 enumerated-composition assignment at N = 24 and 48 faces."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fairguide as fg

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    return sorted(ts)[len(ts) // 2]


res = {}
n = 1024
tmpl = torch.tensor([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655], [70.7299, 92.2041]], device=dev)
s = 1.2 + 1.2 * torch.rand(n, 1, 1, generator=g, device=dev)
th = (torch.rand(n, generator=g, device=dev) - 0.5) * 0.8
R = torch.stack([torch.stack([th.cos(), -th.sin()], -1), torch.stack([th.sin(), th.cos()], -1)], -2)
ctr = 160 + 192 * torch.rand(n, 1, 2, generator=g, device=dev)
lms = ((tmpl - 56.0) @ R.transpose(1, 2)) * s + ctr
ind = torch.rand(n, generator=g, device=dev) > 0.05
for dt in (torch.bfloat16, torch.float32):
    e = torch.empty((), dtype=dt).element_size()
    x = (torch.rand(n, 3, 512, 512, generator=g, device=dev) * 2 - 1).to(dt)
    params = fg.ops.align_matrices(lms, ind, (512, 512))
    go = torch.randn(n, 3, 112, 112, generator=g, device=dev).to(dt)
    gi = torch.zeros_like(x)
    face_px = float((s.flatten() * 112).pow(2).mean())           # source pixels under the chip, per image
    f = timeit(lambda: fg.ops.aligned_warp_fwd(x, params, ind))
    b = timeit(lambda: fg.ops.aligned_warp_bwd(go, params, ind, tuple(x.shape), g_images=gi))
    b0 = timeit(lambda: fg.ops.aligned_warp_bwd(go, params, ind, tuple(x.shape)))
    alg_f = n * 3 * (face_px + 112 * 112) * e
    alg_b = n * 3 * (2 * face_px + 112 * 112) * e               # accumulate: read + write the face region, read g_out
    res[f"aligned_warp_{str(dt).split('.')[-1]}"] = {
        "fwd_ms": round(f, 4), "fwd_GBps": round(alg_f / f / 1e6, 1), "bwd_accumulate_ms": round(b, 4), "bwd_accumulate_GBps": round(alg_b / b / 1e6, 1),
        "bwd_overwrite_ms": round(b0, 4), "bwd_overwrite_GBps": round(n * 3 * (512 * 512 + 112 * 112) * e / b0 / 1e6, 1)}
    del x, gi
res["align_matrices_ms"] = round(timeit(lambda: fg.ops.align_matrices(lms, ind, (512, 512))), 4)

D, d = 100_000, 512
db = torch.nn.functional.normalize(torch.randn(D, d, generator=g, device=dev), dim=-1)
dbound = float(db.norm(dim=1).max())        # once per database
for m in (4, 32, 128):
    q = torch.nn.functional.normalize(torch.randn(m, d, generator=g, device=dev), dim=-1)
    t = timeit(lambda: fg.ops.face_search_top1(q, None, db, db_norm_bound=dbound))
    t_exact = timeit(lambda: fg.ops.face_search_top1(q, None, db))
    b1, s1 = fg.ops.face_search_top1(q, None, db, db_norm_bound=dbound)
    b0, s0 = fg.ops.face_search_top1(q, None, db)
    assert torch.equal(b1, b0) and torch.equal(s1, s0), "tensor-core search differs from the exact search"
    t_torch = timeit(lambda: (q @ db.T).max(dim=1))
    res[f"face_search_m{m}"] = {"ms": round(t, 4), "db_GBps": round(D * d * 4 / t / 1e6, 1), "torch_matmul_max_ms": round(t_torch, 4), "streaming_kernel_ms": round(t_exact, 4),
                                "path": "tcgen05 TF32 + exact re-score" if m >= fg.ops.SEARCH_TC_MIN_QUERIES else "streaming (HBM-bound)"}
m = 32
raw = torch.randn(m, d, generator=g, device=dev)
fo = torch.nn.functional.normalize(torch.randn(m, d, generator=g, device=dev), dim=-1)
face = torch.ones(m, dtype=torch.bool, device=dev)
tg = [torch.randint(-1, 2, (m,), generator=g, device=dev)]
pr = [tg[0].clamp(min=0)]
pb = [torch.softmax(torch.randn(m, 2, generator=g, device=dev) * 3, -1)]
res["face_loss_fwd_m32_ms"] = round(timeit(lambda: fg.ops.face_loss_fwd(raw, fo, db, face, tg, pr, pb, 0.75, True)), 4)

for N in (24, 48):
    p = torch.softmax(torch.randn(N, 4, generator=g, device=dev) * 2, -1)
    combs, w = fg.api.race_compositions(N)
    t = timeit(lambda: fg.generate_dynamic_targets_race(p, True, num_valid=N), iters=5)
    res[f"race_enumerated_N{N}"] = {"ms": round(t, 4), "compositions_solved": int(len(w))}
print(json.dumps(res))
