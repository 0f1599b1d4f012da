#!/usr/bin/env python
"""Diagnostic: device plan counts vs the oracle's for one Monte-Carlo assignment case.  usage: diag_assign.py N kind S dtype"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fairguide as fg
from oracle import assign as oa, emd as oemd
from tests._golden import peaked_probs
N, kind, S, dtype = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), getattr(torch, sys.argv[4])
seed = int(sys.argv[5]) if len(sys.argv) > 5 else N + 1
n_attr = 2 if kind == "e3" else 3
rng = np.random.default_rng(seed)
probs = [torch.tensor(peaked_probs(rng, N, w, 1.5)).to(dtype) for w in ([2, 4] if n_attr == 2 else [2, 4, 2])]
miss = torch.tensor(rng.uniform(size=N) < 0.05)
for p in probs:
    p[miss] = -1
nv = int((~miss).sum())
gen = torch.Generator().manual_seed(N)
rands = tuple(torch.rand(S, nv, generator=gen).to(dtype) for _ in range(n_attr))
outs, counts, ws = fg.api._mc_targets(tuple(p.cuda() for p in probs), True, S, tuple(r.cuda() for r in rands), nv, None, None, return_counts=True)
print("status", ws.status())
v = [p[~miss] for p in probs]
M = oa.cost_matrix(v[0], v[1], v[2] if n_attr == 3 else None)
h = oa.draw_histograms(*rands)
ref = oa.plan_counts(M, h, oemd.emd_c)
dev = counts.cpu().numpy().astype(np.float64)
bad = np.nonzero((dev != ref).any(1))[0]
print("N", N, "valid", nv, "rows with different counts:", len(bad), "col sums equal:", np.array_equal(dev.sum(0), ref.sum(0)), "row sums S:", (dev.sum(1) == S).all())
for i in bad[:8]:
    print(i, dev[i], ref[i], "cost", np.round(M[i], 6))
# objective of the summed plan: sum_ij counts * M  (equal iff every draw is optimal, since the oracle's are)
print("objective dev", (dev * M).sum(), "ref", (ref * M).sum(), "diff", (dev * M).sum() - (ref * M).sum())
