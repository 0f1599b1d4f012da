#!/usr/bin/env python
"""CPU prototype of the solver's price search (fg_assign.cu, step 1 of ot_solve_kernel): the sign-based dual ascent on the
full problem versus the two-level schedule (coarse rounds on a strided sample of the rows, then full rounds inside a trust
region where only the rows with a small margin are swept).  Prints the count residual per schedule and the share of rows
that stay active.  Used to choose the round counts / radius before spending GPU time; not part of the product.
usage: proto_search.py [N] [sharp]"""
import sys
import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 7720
sharp = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
K = 16
rng = np.random.default_rng(0)


def softmax(z):
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(-1, keepdims=True)


def cost(pg, pr, pa):
    M = np.zeros((N, K))
    for j in range(K):
        gi, ri, ai = j >> 3, (j >> 1) & 3, j & 1
        ng2 = ((pg - np.eye(2)[gi]) ** 2).sum(-1)
        nr2 = ((pr - np.eye(4)[ri]) ** 2).sum(-1)
        ca2 = (pa[:, 0] - 1) ** 2 + pa[:, 1] ** 2 if ai == 0 else (2 * pa[:, 0]) ** 2 + (pa[:, 1] - 1) ** 2
        M[:, j] = np.sqrt(ng2 + nr2 + ca2)
    return M.astype(np.float32)


pg, pr, pa = (softmax(rng.standard_normal((N, w)) * sharp) for w in (2, 4, 2))
M = cost(pg, pr, pa)
q = np.array([0.5 * 0.25 * (0.25 if (j & 1) else 0.75) for j in range(K)])


def sweep(Msub, p):
    return np.bincount(np.argmin(Msub - p, axis=1), minlength=K)


def sign_search(Msub, b, p, step0, iters, trace=None):
    """b may be fractional (scaled demand of a sample)."""
    p = p.copy(); step = np.full(K, step0); prev = np.zeros(K); best = (1e18, p.copy())
    for it in range(iters + 1):
        cnt = sweep(Msub, p)
        err = b - cnt
        resid = np.abs(err).sum() / 2
        if trace is not None:
            trace.append(resid)
        if resid < best[0]:
            best = (resid, p.copy())
        if resid == 0 or it == iters:
            break
        sg = np.sign(err)
        step = np.where(sg * prev < 0, step * 0.5, np.where(sg * prev > 0, step * 1.2, step))
        prev = sg
        p = p + step * sg                     # too few rows -> larger price -> cheaper class (reduced cost = M - p)
    return best


# NOTE on sign: the kernel subtracts prices from costs (x - pk) and ADDS step*sg to the price when err > 0 (too few rows):
# a larger price makes the class cheaper.  Same convention here: reduced cost = M - p, p += step * sign(err).
def run(draws=8):
    out = {"full40": [], "two_level": [], "active": [], "dp": []}
    sample = (np.arange(2048) * N // 2048) if N > 2048 else np.arange(N)
    for d in range(draws):
        b = np.bincount(rng.choice(K, size=N, p=q), minlength=K)
        tr = []
        r_full, p_full = sign_search(M, b, np.zeros(K), 0.03, 40, tr)
        g = lambda i: tr[i] if i < len(tr) else 0
        out["full40"].append((r_full, len(tr) - 1, g(8), g(16), g(24), g(32)))
        for r1, r2, st2 in ((24, 16, 0.004), (28, 12, 0.004), (32, 12, 0.002), (32, 16, 0.002), (40, 12, 0.002)):
            frac = len(sample) / N
            r_c, p_c = sign_search(M[sample], b * frac, np.zeros(K), 0.03, r1)
            tr2 = []
            r_f, p_f = sign_search(M, b, p_c, st2, r2, tr2)
            out["two_level"].append((r1, r2, st2, round(r_c, 1), tr2[0], r_f, len(tr2) - 1))
            if (r1, r2) == (32, 12):
                dp = np.abs(p_f - p_c).max()
                out["dp"].append(dp)
                red = np.sort(M - p_c, axis=1)
                gap = red[:, 1] - red[:, 0]
                out["active"].append([float((gap < 2 * R).mean()) for R in (0.005, 0.01, 0.02, 0.04)])
    return out


o = run()
print("full 40 rounds: final resid, rounds used, resid@8, @16, @24, @32 per draw:", o["full40"])
print("two-level (r1, r2, step2, coarse resid, first full resid, final resid):")
for t in o["two_level"]:
    print("  ", t)
print("max |p_final - p_coarse| per draw:", [round(float(x), 4) for x in o["dp"]])
print("active share for R = 0.005, 0.01, 0.02, 0.04:", o["active"])


def trust_stats(draws=6):
    """After r0 full rounds from zero prices: how far do the prices still move, and which share of the rows has a margin
    below twice that distance (the rows a trust-region sweep must keep)?"""
    print("trust region: r0, resid@r0, max|p40 - p_r0|, active share at R = 1.25 * that")
    for d in range(draws):
        b = np.bincount(rng.choice(K, size=N, p=q), minlength=K)
        p = np.zeros(K); step = np.full(K, 0.03); prev = np.zeros(K); hist = []
        for it in range(41):
            cnt = sweep(M, p); err = b - cnt; hist.append((p.copy(), np.abs(err).sum() / 2))
            sg = np.sign(err)
            step = np.where(sg * prev < 0, step * 0.5, np.where(sg * prev > 0, step * 1.2, step)); prev = sg
            p = p + step * sg
        row = []
        for r0 in (6, 8, 10, 12, 16, 20):
            p0, res0 = hist[r0]
            dist = max(np.abs(h[0] - p0).max() for h in hist[r0:])
            red = np.sort(M - p0, axis=1); gap = red[:, 1] - red[:, 0]
            row.append((r0, res0, round(float(dist), 4), round(float((gap < 2.5 * dist).mean()), 3)))
        print("  ", row)


trust_stats()


def trust_sim(r0=12, c=4.0, draws=8, cap_share=0.5):
    """Full algorithm: r0 full rounds, then rounds inside a trust region of radius c * (largest current step); leaving the
    region re-pivots (one more full sweep).  Counts are exact either way, so the price trajectory equals the plain search."""
    tot = []
    for d in range(draws):
        b = np.bincount(rng.choice(K, size=N, p=q), minlength=K)
        p = np.zeros(K); step = np.full(K, 0.03); prev = np.zeros(K)
        pivots = []; p0 = None; R = 0; full = 0; cheap = 0; act = []
        for it in range(41):
            if it >= r0 and (p0 is None or np.abs(p - p0).max() > R):
                p0 = p.copy(); R = c * step.max(); full += 1
                red = np.sort(M - p0, axis=1); gap = red[:, 1] - red[:, 0]
                act.append(round(float((gap <= 2 * R).mean()), 3)); pivots.append(it)
            elif it >= r0:
                cheap += 1
            else:
                full += 1
            cnt = sweep(M, p); err = b - cnt
            if np.abs(err).sum() == 0:
                break
            sg = np.sign(err)
            step = np.where(sg * prev < 0, step * 0.5, np.where(sg * prev > 0, step * 1.2, step)); prev = sg
            p = p + step * sg
        tot.append((full, cheap, pivots, act))
    print(f"trust sim r0={r0} c={c}: (full sweeps, cheap sweeps, pivot rounds, active share at each pivot)")
    for t in tot:
        print("  ", t)


for r0, c in ((10, 4.0), (12, 4.0), (12, 6.0), (14, 4.0)):
    trust_sim(r0, c)


def trust_sim_pc(r0=12, c=3.0, draws=8, floor=0.0):
    """Per-class radius R_l = max(c * step_l, floor)."""
    tot = []
    for d in range(draws):
        b = np.bincount(rng.choice(K, size=N, p=q), minlength=K)
        p = np.zeros(K); step = np.full(K, 0.03); prev = np.zeros(K)
        pivots = []; p0 = None; R = None; full = 0; cheap = 0; act = []
        for it in range(41):
            if it >= r0 and (p0 is None or (np.abs(p - p0) > R).any()):
                p0 = p.copy(); R = np.maximum(c * step, floor); full += 1
                z = M - (p0 + R); s = np.argmin(z, axis=1)
                zs = np.sort(z, axis=1)
                frozen = zs[:, 0] + 2 * R[s] < zs[:, 1]
                act.append(round(float(1 - frozen.mean()), 3)); pivots.append(it)
            elif it >= r0:
                cheap += 1
            else:
                full += 1
            cnt = sweep(M, p); err = b - cnt
            if np.abs(err).sum() == 0:
                break
            sg = np.sign(err)
            step = np.where(sg * prev < 0, step * 0.5, np.where(sg * prev > 0, step * 1.2, step)); prev = sg
            p = p + step * sg
        tot.append((full, cheap, pivots, act))
    print(f"per-class trust sim r0={r0} c={c} floor={floor}:")
    for t in tot:
        print("  ", t)


for r0, c in ((8, 2.0), (10, 2.0), (10, 3.0), (12, 2.0), (12, 3.0)):
    trust_sim_pc(r0, c)


def schedule_cost(C, pivot, total, draws=10, c=3.0, P=4):
    """Cycle model (k cycles, measured on B200 at N = 7720): coarse round 2.5, full round 11, classification 15, trust round
    4.5, repair step 25.  C coarse rounds on every P-th row (same step state carried on), trust region from round `pivot`."""
    sample = np.arange(0, N, P)
    frac = len(sample) / N
    costs = []
    for d in range(draws):
        b = np.bincount(rng.choice(K, size=N, p=q), minlength=K)
        p = np.zeros(K); step = np.full(K, 0.03); prev = np.zeros(K)
        p0 = None; R = None; cyc = 0.0; best = 1e18
        for it in range(total + 1):
            coarse = it < C
            if coarse:
                cnt = sweep(M[sample], p); err = b * frac - cnt; cyc += 2.5
                err = np.where(np.abs(err) < 0.5, 0, err)
            else:
                if it >= pivot and (p0 is None or (np.abs(p - p0) > R).any()):
                    p0 = p.copy(); R = c * step; cyc += 15
                cyc += 4.5 if p0 is not None else 11
                cnt = sweep(M, p); err = b - cnt
                best = min(best, np.abs(err).sum() / 2)
                if best == 0:
                    break
            sg = np.sign(err)
            step = np.where(sg * prev < 0, step * 0.5, np.where(sg * prev > 0, step * 1.2, step)); prev = sg
            p = p + step * sg
        costs.append((cyc + 25 * best, best))
    cs = np.array([x[0] for x in costs]); rs = [x[1] for x in costs]
    print(f"C={C:2d} pivot={pivot:2d} total={total:2d}: search+repair mean {cs.mean():6.1f}k max {cs.max():6.1f}k  resid {rs}")


print("schedules (coarse rounds, pivot round, total rounds):")
for C, pv, tot in ((0, 12, 40), (0, 10, 40), (8, 12, 40), (8, 10, 40), (10, 12, 40), (10, 12, 44), (12, 14, 44), (12, 14, 40), (16, 18, 48), (0, 12, 32), (8, 12, 36)):
    schedule_cost(C, pv, tot)


def schedule_cost2(pivot, c, total=40, draws=10, shrink=1.0):
    """Refined cycle model: full round 11k; classification 20k; trust round 2.2k + 0.3k per row-slot (512 rows);
    repair 25k.  Active share measured per pivot (per-class radius c * step)."""
    reg = min(N, 2048)
    cs = []; npiv = []; acts = []
    for d in range(draws):
        b = np.bincount(rng.choice(K, size=N, p=q), minlength=K)
        p = np.zeros(K); step = np.full(K, 0.03); prev = np.zeros(K)
        p0 = None; R = None; cyc = 0.0; best = 1e18; act_rows = 0; piv = 0
        for it in range(total + 1):
            if it >= pivot and (p0 is None or (np.abs(p - p0) > R).any()):
                p0 = p.copy(); R = c * step; cyc += 20; piv += 1
                z = M[reg:] - (p0 + R); s = np.argmin(z, axis=1); zs = np.sort(z, axis=1)
                act_rows = int((~(zs[:, 0] + 2 * R[s] < zs[:, 1])).sum()); acts.append(act_rows)
            cyc += (2.2 + 0.3 * (reg + act_rows) / 512) if p0 is not None else 11
            cnt = sweep(M, p); err = b - cnt
            best = min(best, np.abs(err).sum() / 2)
            if best == 0:
                break
            sg = np.sign(err)
            step = np.where(sg * prev < 0, step * 0.5, np.where(sg * prev > 0, step * 1.2, step)); prev = sg
            p = p + step * sg
        cs.append(cyc + 25 * best); npiv.append(piv)
    cs = np.array(cs)
    print(f"pivot={pivot:2d} c={c:3.1f}: mean {cs.mean():6.1f}k max {cs.max():6.1f}k  pivots/draw {np.mean(npiv):.1f}  max active rows {max(acts)}")


print("refined model:")
for pv in (8, 10, 12, 14):
    for c in (2.0, 3.0, 4.0, 6.0):
        schedule_cost2(pv, c)
