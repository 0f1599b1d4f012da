"""CPU: the oracle's exact transport solver (C, row insertion) against two independent exact
solvers from scipy, plus plan invariants.  POT itself is not installable here (no network);
see oracle/emd.py."""
import numpy as np
import pytest

from oracle import emd as oemd
from tests._golden import peaked_probs


def _problem(seed, N, K):
    rng = np.random.Generator(np.random.PCG64(seed))
    pg, pr = peaked_probs(rng, N, 2, 2.0), peaked_probs(rng, N, 4, 2.0)
    pa = peaked_probs(rng, N, 2, 2.0) if K == 16 else None
    M = oemd.cost_matrix_c(pg, pr, pa)
    q = np.full(K, 1.0 / K) if K == 8 else np.tile([0.75, 0.25], 8) / 8
    b = rng.multinomial(N, q)
    return M, b


@pytest.mark.parametrize("N,K", [(1, 8), (7, 8), (32, 8), (64, 16), (257, 8), (512, 16)])
def test_c_solver_matches_lsa(N, K):
    for seed in range(3):
        M, b = _problem(seed * 17 + N, N, K)
        T = oemd.emd_c(np.ones(N), b, M)
        assert np.array_equal(T.sum(0).astype(np.int64), b) and np.all(T.sum(1) == 1)
        assert set(np.unique(T)) <= {0.0, 1.0}
        assert np.array_equal(T, oemd.emd_lsa(np.ones(N), b, M))


@pytest.mark.parametrize("N,K", [(24, 8), (40, 16)])
def test_c_solver_matches_lp(N, K):
    M, b = _problem(99 + N, N, K)
    assert np.array_equal(oemd.emd_c(np.ones(N), b, M), oemd.emd_lp(np.ones(N), b, M))


def test_extreme_demands():
    M, _ = _problem(5, 50, 8)
    b = np.zeros(8, dtype=np.int64)
    b[3] = 50
    assert np.array_equal(oemd.assign_c(M, b), np.full(50, 3))
    b = np.array([50, 0, 0, 0, 0, 0, 0, 0])
    assert np.array_equal(oemd.assign_c(M, b), np.zeros(50))


def test_cost_matrix_matches_literal_formula():
    import torch
    from oracle import assign as oassign
    rng = np.random.Generator(np.random.PCG64(3))
    pg, pr, pa = peaked_probs(rng, 33, 2), peaked_probs(rng, 33, 4), peaked_probs(rng, 33, 2)
    for a in (None, pa):
        lit = oassign.cost_matrix_literal(torch.tensor(pg), torch.tensor(pr), None if a is None else torch.tensor(a))
        fix = oemd.cost_matrix_c(pg, pr, a)
        np.testing.assert_allclose(fix, lit, rtol=4e-16, atol=0)   # <= 2 ulp: pow(x,.5) vs sqrt, BLAS dot order
