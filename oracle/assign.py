"""ORACLE (test infrastructure): distributional-alignment target assignment.

Restates
    generate_dynamic_targets                  E1:1403-1447   rank split + binomial CDF
    generate_dynamic_targets_gender_race      E3:1459-1569   Monte-Carlo exact EMD, K=8
    generate_dynamic_targets_gender_race_age  E4:1477-1615   same, K=16, 75/25 age target
    thresholding + local slice                E3:2022-2025, E4:2129-2135, E1:1835-1837

Differences from the reference that are plumbing, not arithmetic:
  * ``rand_tensors`` lets the caller pass the uniform draws explicitly (the reference draws
    them with torch.rand inside, E3:1491-1492 / E4:1503-1505; CPU and CUDA generators give
    different streams, so parity tests hand the same draws to both sides);
  * ``world_rand`` simulates the other ranks of ``torch.distributed.all_reduce(target_probs,
    SUM)`` (E3:1535): a list with one tuple of draws per rank, whose plans are summed;
  * ``emd`` selects the exact solver standing in for ``ot.emd`` (oracle/emd.py);
  * ``literal=True`` keeps the reference's Python loops (0-d tensor histogram E3:1503-1507,
    per-element cost loop E3:1516-1525) -- that is the CPU baseline; ``literal=False`` runs
    the same arithmetic vectorised.
"""
import itertools
import math

import numpy as np
import scipy.stats
import torch

from . import emd as _emd


# ----------------------------------------------------------------------------- E1
@torch.no_grad()
def generate_dynamic_targets(probs, target_ratio=0.5, w_uncertainty=False):
    """E1:1403-1447.  ``probs`` [N,2] with -1 rows for images without a face."""
    has_face = (probs != -1).all(dim=-1)
    p = probs[has_face]
    rank = torch.argsort(torch.argsort(p[:, 1]))
    targets = (rank >= (rank.shape[0] * target_ratio)).long()
    targets_all = torch.ones([probs.shape[0]], dtype=torch.long, device=probs.device) * (-1)
    targets_all[has_face] = targets
    if not w_uncertainty:
        return targets_all
    n = p.shape[0]
    unc = torch.ones([n], dtype=probs.dtype, device=probs.device) * (-1)
    unc[targets == 1] = torch.tensor(
        1 - scipy.stats.binom.cdf(rank[targets == 1].cpu().numpy(), n, 1 - target_ratio)
    ).to(probs.dtype).to(probs.device)
    unc[targets == 0] = torch.tensor(
        scipy.stats.binom.cdf(rank[targets == 0].cpu().numpy(), n, target_ratio)
    ).to(probs.dtype).to(probs.device)
    unc_all = torch.ones([probs.shape[0]], dtype=probs.dtype, device=probs.device) * (-1)
    unc_all[has_face] = unc
    return targets_all, unc_all


# ----------------------------------------------------------------------- E3 / E4 shared
def _classes_from_rand(rg, rr, ra):
    """E3:1496-1501, E4:1510-1518."""
    cg = (rg > 0.5).int()
    cr = torch.zeros_like(cg)
    cr[(rr > 1 / 4) * (rr <= 2 / 4)] = 1
    cr[(rr > 2 / 4) * (rr <= 3 / 4)] = 2
    cr[(rr > 3 / 4)] = 3
    ca = None
    if ra is not None:
        ca = torch.zeros_like(cg)
        ca[(ra > 0.75)] = 1
    return cg, cr, ca


def draw_histograms(rg, rr, ra=None, literal=False):
    """Per-draw class-count vectors ``b`` ([S,8] or [S,16] python ints / int64)."""
    cg, cr, ca = _classes_from_rand(rg, rr, ra)
    K = 8 if ca is None else 16
    if literal:
        combs = []
        if ca is None:
            for gs, rs in itertools.zip_longest(cg, cr):
                freq = [0] * 8
                for g, r in itertools.zip_longest(gs, rs):
                    freq[g * 4 + r] += 1
                combs.append(freq)
        else:
            for gs, rs, as_ in itertools.zip_longest(cg, cr, ca):
                freq = [0] * 16
                for g, r, a in itertools.zip_longest(gs, rs, as_):
                    freq[g * 8 + r * 2 + a] += 1
                combs.append(freq)
        return np.array(combs, dtype=np.int64).reshape(len(combs), K)
    idx = (cg * 4 + cr) if ca is None else (cg * 8 + cr * 2 + ca)
    idx = idx.to(torch.int64)
    S = idx.shape[0]
    out = torch.zeros((S, K), dtype=torch.int64)
    out.scatter_add_(1, idx, torch.ones_like(idx))
    return out.numpy()


_G8 = np.array([[1, 0]] * 4 + [[0, 1]] * 4)
_R8 = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]] * 2)
_G16 = np.array([[1, 0]] * 8 + [[0, 1]] * 8)
_R16 = np.array(([[1, 0, 0, 0]] * 2 + [[0, 1, 0, 0]] * 2 + [[0, 0, 1, 0]] * 2 + [[0, 0, 0, 1]] * 2) * 2)
_A16 = np.array([[1, 0], [0, 1]] * 8)


def cost_matrix_literal(pg, pr, pa=None):
    """E3:1509-1526 / E4:1533-1560, element by element like the reference.

    The rows are widened to float64 where the reference creates them: under its pinned NumPy
    1.26 every expression below is promoted to float64 anyway (array - int64 array; float32
    scalar - Python int), whereas NumPy 2 would keep the age residual in float32."""
    M = []
    if pg.dtype == torch.bfloat16:      # numpy has no bfloat16 (the reference itself only runs fp16 / fp32 here); exact widening
        pg, pr, pa = pg.float(), pr.float(), (None if pa is None else pa.float())
    for i in range(pg.shape[0]):
        g = np.array(pg[i].cpu()).astype(np.float64)
        r = np.array(pr[i].cpu()).astype(np.float64)
        row = []
        if pa is None:
            for j in range(8):
                row.append((np.linalg.norm(g - _G8[j]) ** 2 + np.linalg.norm(r - _R8[j]) ** 2) ** 0.5)
        else:
            a = np.array(pa[i].cpu()).astype(np.float64)
            for j in range(16):
                if (_A16[j] == [1, 0]).all():
                    ca = math.sqrt((a[0] - 1) ** 2 + (a[1] - 0) ** 2)
                else:
                    ca = math.sqrt(((a[0] - 0) * 2) ** 2 + (a[1] - 1) ** 2)
                row.append((np.linalg.norm(g - _G16[j]) ** 2 + np.linalg.norm(r - _R16[j]) ** 2 + ca ** 2) ** 0.5)
        M.append(row)
    return np.array(M, dtype=np.float64).reshape(pg.shape[0], 8 if pa is None else 16)


def cost_matrix(pg, pr, pa=None, literal=False):
    if literal:
        return cost_matrix_literal(pg, pr, pa)
    f = lambda t: None if t is None else t.detach().cpu().to(torch.float32).numpy().astype(np.float64)
    return _emd.cost_matrix_c(f(pg), f(pr), f(pa))


def plan_counts(M, hists, emd=None):
    """Sum of the per-draw 0/1 plans (E3:1528-1532): float64 [N,K]."""
    solver = emd or _emd.emd_c
    N, K = M.shape
    acc = np.zeros([N, K])
    a = np.ones([N])
    for b in hists:
        acc += solver(a, b, M)
    return acc


def _marginals(tp, K):
    """E3:1538-1549 / E4:1572-1589, same ops on a probs-dtype tensor."""
    col = lambda idx: tp[:, idx].sum(dim=-1).unsqueeze(dim=-1)
    if K == 8:
        g = torch.cat([tp[:, :4].sum(dim=-1).unsqueeze(dim=-1), tp[:, 4:].sum(dim=-1).unsqueeze(dim=-1)], dim=-1)
        r = torch.cat([col([0, 4]), col([1, 5]), col([2, 6]), col([3, 7])], dim=-1)
        return [g, r]
    g = torch.cat([tp[:, :8].sum(dim=-1).unsqueeze(dim=-1), tp[:, 8:].sum(dim=-1).unsqueeze(dim=-1)], dim=-1)
    r = torch.cat([col([0, 1, 8, 9]), col([2, 3, 10, 11]), col([4, 5, 12, 13]), col([6, 7, 14, 15])], dim=-1)
    a = torch.cat([col([0, 2, 4, 6, 8, 10, 12, 14]), col([1, 3, 5, 7, 9, 11, 13, 15])], dim=-1)
    return [g, r, a]


def targets_from_counts(counts, valid, dtype, w_uncertainty=True):
    """Epilogue E3:1534-1569 from the (already rank-summed) plan counts.

    ``counts`` float64/int [N,K] numpy; ``valid`` bool [N_all].  Returns the flat tuple the
    reference returns: (t_g, u_g, t_r, u_r[, t_a, u_a]) or targets only."""
    K = counts.shape[1]
    tp = torch.tensor(np.asarray(counts, dtype=np.float64)).to(dtype)
    tp = tp / tp[0, :].sum()
    outs = []
    n_all = valid.shape[0]
    for m in _marginals(tp, K):
        t = m.argmax(axis=-1).to(torch.long)
        u = 1 - m.max(axis=-1).values
        t_all = torch.ones([n_all], dtype=torch.long) * (-1)
        t_all[valid] = t
        outs.append(t_all)
        if w_uncertainty:
            u_all = torch.ones([n_all], dtype=dtype) * (-1)
            u_all[valid] = u
            outs.append(u_all)
    return tuple(outs)


@torch.no_grad()
def _mc_emd(pg, pr, pa, w_uncertainty, num_samples_per_device, rand_tensors, world_rand, emd, literal):
    valid = (pg != -1).all(dim=-1) * (pr != -1).all(dim=-1)
    n_attr = 2 if pa is None else 3
    if valid.sum() == 0:
        outs = []
        for src in ([pg, pr] if pa is None else [pg, pr, pa]):
            outs.append(torch.ones([src.shape[0]], dtype=torch.long, device=src.device) * (-1))
            if w_uncertainty:
                outs.append(torch.ones([src.shape[0]], dtype=src.dtype, device=src.device) * (-1))
        return tuple(outs)
    g, r = pg[valid], pr[valid]
    a = pa[valid] if pa is not None else None
    N = g.shape[0]
    if world_rand is None:
        if rand_tensors is None:
            rand_tensors = tuple(torch.rand([num_samples_per_device, N], dtype=pg.dtype, device=pg.device)
                                 for _ in range(n_attr))
        world_rand = [rand_tensors]
    M = cost_matrix(g, r, a, literal=literal)
    total = np.zeros_like(M)
    for rt in world_rand:
        rt = tuple(rt) + (None,) * (3 - len(rt))
        hists = draw_histograms(rt[0], rt[1], rt[2] if pa is not None else None, literal=literal)
        # per rank: cast to the probs dtype before the all-reduce like E3:1534-1535
        total += torch.tensor(plan_counts(M, hists, emd)).to(pg.dtype).to(torch.float64).numpy()
    return targets_from_counts(total, valid, pg.dtype, w_uncertainty)


def generate_dynamic_targets_gender_race(probs_gender, probs_race, w_uncertainty=False, num_samples_per_device=100,
                                         rand_tensors=None, world_rand=None, emd=None, literal=True):
    """E3:1459-1569 -> (t_g, u_g, t_r, u_r) or (t_g, t_r)."""
    return _mc_emd(probs_gender, probs_race, None, w_uncertainty, num_samples_per_device, rand_tensors, world_rand, emd, literal)


def generate_dynamic_targets_gender_race_age(probs_gender, probs_race, probs_age, w_uncertainty=False,
                                             num_samples_per_device=100, rand_tensors=None, world_rand=None,
                                             emd=None, literal=True):
    """E4:1477-1615 -> (t_g, u_g, t_r, u_r, t_a, u_a) or targets only."""
    return _mc_emd(probs_gender, probs_race, probs_age, w_uncertainty, num_samples_per_device, rand_tensors, world_rand, emd, literal)


def threshold_and_slice(targets_all, uncertainty_all, threshold, n_local, rank):
    """E3:2022-2025: ``targets[unc > thr] = -1`` then this rank's rows."""
    t = targets_all.clone()
    t[uncertainty_all > threshold] = -1
    return t[n_local * rank:n_local * (rank + 1)], t


# ----------------------------------------------------------------------------- E6 (race only)
def pot_dist_euclidean(x1, x2):
    """``ot.dist(x1, x2, metric="euclidean")`` as POT 0.9.3 (environment.yml:172; not installed here) evaluates it on
    the numpy backend -- ot/utils.py ``euclidean_distances(X, Y, squared=False)``: squared norms by einsum in the
    inputs' own dtypes, the cross term ``-2 * X.Y^T``, both norms added in place, clamp at 0, square root.
    Call site: exp-6-debias-race/1-main-debias.py:1461."""
    a2 = np.einsum("ij,ij->i", x1, x1)
    b2 = np.einsum("ij,ij->i", x2, x2)
    c = -2 * np.dot(x1, x2.T)
    c += a2[:, None]
    c += b2[None, :]
    c = np.maximum(c, 0)
    return np.sqrt(c)


def race_compositions(N):
    """E6:1438-1459: all compositions of N into 4 parts, multinomial weights, 95 % head (statement by statement)."""
    all_combs, all_probs = [], []
    for n1 in range(N + 1):
        for n2 in range(N - n1 + 1):
            for n3 in range(N - n1 - n2 + 1):
                n4 = N - n1 - n2 - n3
                all_combs.append([n1, n2, n3, n4])
                all_probs.append(math.comb(N, n1) * math.comb(N - n1, n2) * math.comb(N - n1 - n2, n3))
    all_combs = np.array(all_combs)
    all_probs = np.array(all_probs)
    all_probs = all_probs / np.linalg.norm(all_probs, ord=1)
    idxs_sorted = np.flip(all_probs.argsort())
    prob_accumulate = 0
    for i_idx, idx in enumerate(idxs_sorted):
        prob_accumulate += all_probs[idx]
        if prob_accumulate > 0.95:
            break
    return all_combs[idxs_sorted[:i_idx + 1]], all_probs[idxs_sorted[:i_idx + 1]]


@torch.no_grad()
def generate_dynamic_targets_race(probs, w_uncertainty=False, emd=None):
    """exp-6-debias-race/1-main-debias.py:1413-1482 with ``ot.dist`` / ``ot.emd`` restated (pot_dist_euclidean, oracle/emd.py)."""
    solver = emd or _emd.emd_c
    idxs_2_rank = (probs != -1).all(dim=-1)
    probs_2_rank = probs[idxs_2_rank]
    N = probs_2_rank.shape[0]
    targets_all = torch.ones([probs.shape[0]], dtype=torch.long, device=probs.device) * (-1)
    uncertainty_all = torch.ones([probs.shape[0]], dtype=probs.dtype, device=probs.device) * (-1)
    if N > 0:        # (the reference divides 0/0 for N = 0 and scatters nothing)
        a = np.ones([N])
        target_points = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        all_combs, all_probs = race_compositions(N)
        M = pot_dist_euclidean(np.array(probs_2_rank.cpu()), target_points)
        target_probs = np.zeros([N, 4])
        for b, prob in itertools.zip_longest(all_combs, all_probs):
            T = solver(a, b, M)
            target_probs += T * prob
        target_probs = target_probs / np.expand_dims(np.linalg.norm(target_probs, axis=-1, ord=1), axis=-1)
        targets_all[idxs_2_rank] = torch.tensor(target_probs.argmax(axis=-1)).to(torch.long)
        uncertainty_all[idxs_2_rank] = torch.tensor(1 - target_probs.max(axis=-1)).to(probs.dtype)
    if w_uncertainty:
        return targets_all, uncertainty_all
    return targets_all
