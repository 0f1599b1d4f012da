"""GPU experiment: which association order do torch-CUDA reductions use for the tiny sums of the
reference's epilogue (E3:1538-1549, E4:1572-1589)?  Prints, per candidate order, the fraction of
rows that match torch bit-for-bit."""
import itertools
import numpy as np
import torch

rng = np.random.default_rng(0)
dev = "cuda"


def orders(n):
    a = lambda x: x
    c = {}
    c["seq"] = lambda v: functools_reduce(v)
    return c


def seq(v):
    s = v[0]
    for x in v[1:]:
        s = s + x
    return s


def pair(v):
    v = list(v)
    while len(v) > 1:
        v = [v[i] + v[i + 1] if i + 1 < len(v) else v[i] for i in range(0, len(v), 2)]
    return v[0]


def strided(v, k):
    # k accumulators, element j goes to accumulator j % k, accumulators then combined pairwise
    acc = [None] * k
    for j, x in enumerate(v):
        acc[j % k] = x if acc[j % k] is None else acc[j % k] + x
    acc = [a for a in acc if a is not None]
    return pair(acc)


def strided_seq(v, k):
    acc = [None] * k
    for j, x in enumerate(v):
        acc[j % k] = x if acc[j % k] is None else acc[j % k] + x
    acc = [a for a in acc if a is not None]
    return seq(acc)


for K, T in ((8, 100), (8, 800), (16, 100), (16, 800)):
    c = rng.multinomial(T, rng.dirichlet(np.ones(K) * 0.3, size=200000)).astype(np.float32)
    tp = torch.tensor(c, device=dev)
    tp = tp / tp[0, :].sum()
    cases = {}
    if K == 8:
        cases["g0 = tp[:, :4].sum"] = (tp[:, :4].sum(-1), [0, 1, 2, 3])
        cases["g1 = tp[:, 4:].sum"] = (tp[:, 4:].sum(-1), [4, 5, 6, 7])
        cases["r1 = tp[:, [1,5]].sum"] = (tp[:, [1, 5]].sum(-1), [1, 5])
        cases["row0 total"] = None
    else:
        cases["g0 = tp[:, :8].sum"] = (tp[:, :8].sum(-1), list(range(8)))
        cases["g1 = tp[:, 8:].sum"] = (tp[:, 8:].sum(-1), list(range(8, 16)))
        cases["r1 = tp[:, [2,3,10,11]].sum"] = (tp[:, [2, 3, 10, 11]].sum(-1), [2, 3, 10, 11])
        cases["a1 = tp[:, odd].sum"] = (tp[:, [1, 3, 5, 7, 9, 11, 13, 15]].sum(-1), [1, 3, 5, 7, 9, 11, 13, 15])
    for name, val in cases.items():
        if val is None:
            continue
        ref, cols = val
        v = [tp[:, j] for j in cols]
        res = {"seq": seq(v), "pair": pair(v)}
        for k in (2, 4):
            if len(cols) > k:
                res[f"strided{k}_pair"] = strided(v, k)
                res[f"strided{k}_seq"] = strided_seq(v, k)
        res["rev_seq"] = seq(v[::-1])
        line = ", ".join(f"{k}={float((r == ref).float().mean()):.4f}" for k, r in res.items())
        print(f"K={K} T={T} {name}: {line}")
    # CPU comparison of the same ops
    tpc = tp.cpu()
    if K == 8:
        print("   cpu vs cuda g0 equal frac:", float((tpc[:, :4].sum(-1) == tp[:, :4].sum(-1).cpu()).float().mean()))
    else:
        print("   cpu vs cuda g0 equal frac:", float((tpc[:, :8].sum(-1) == tp[:, :8].sum(-1).cpu()).float().mean()))
