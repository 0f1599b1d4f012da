"""ORACLE (test infrastructure): exact transport solve standing in for POT ``ot.emd``.

Call sites in the reference: E3:1531, E4:1565 (``T = ot.emd(ones(N), b, M)``).  POT 0.9.3
(environment.yml:172) is absent from this image; see oracle/__init__.py.  Three exact
solvers, all returning the dense plan ``T`` [N,K] float64:

    emd_c      row-insertion solver in C (oracle/emd_rowinsert.c), used for large N
    emd_lsa    scipy.optimize.linear_sum_assignment on the column-replicated cost matrix
    emd_lp     scipy.optimize.linprog(method="highs-ds")
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "libfgoracle.so")
    src = os.path.join(_HERE, "emd_rowinsert.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.fg_oracle_emd_unit.restype = ctypes.c_int
        lib.fg_oracle_emd_unit.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p]
        lib.fg_oracle_cost_matrix.restype = None
        lib.fg_oracle_cost_matrix.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def assign_c(M, b):
    """Class index per row, exact optimum (int32 [N])."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.int64)
    N, K = M.shape
    out = np.empty((N,), dtype=np.int32)
    obj = ctypes.c_double(0.0)
    rc = _lib().fg_oracle_emd_unit(M.ctypes.data, b.ctypes.data, N, K, out.ctypes.data, ctypes.byref(obj))
    if rc != 0:
        raise RuntimeError(f"fg_oracle_emd_unit failed: {rc}")
    return out


def _plan(assign, K):
    T = np.zeros((assign.shape[0], K), dtype=np.float64)
    T[np.arange(assign.shape[0]), assign] = 1.0
    return T


def emd_c(a, b, M):
    a = np.asarray(a)
    assert np.all(a == 1), "oracle solver is specialised to unit supplies (E3:1529)"
    return _plan(assign_c(M, np.asarray(b)), M.shape[1])


def emd_lsa(a, b, M):
    from scipy.optimize import linear_sum_assignment
    b = np.asarray(b, dtype=np.int64)
    M = np.asarray(M, dtype=np.float64)
    cols = np.repeat(np.arange(M.shape[1]), b)
    r, c = linear_sum_assignment(M[:, cols])
    assign = np.empty((M.shape[0],), dtype=np.int64)
    assign[r] = cols[c]
    return _plan(assign, M.shape[1])


def emd_lp(a, b, M):
    from scipy.optimize import linprog
    from scipy.sparse import lil_matrix
    M = np.asarray(M, dtype=np.float64)
    N, K = M.shape
    A = lil_matrix((N + K, N * K))
    for i in range(N):
        A[i, i * K:(i + 1) * K] = 1
    for j in range(K):
        A[N + j, j::K] = 1
    rhs = np.concatenate([np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)])
    res = linprog(M.ravel(), A_eq=A.tocsr(), b_eq=rhs, bounds=(0, None), method="highs-ds")
    assert res.status == 0, res.message
    return np.rint(res.x.reshape(N, K))


def cost_matrix_c(pg, pr, pa=None):
    """Cost matrix in the fixed IEEE op order the CUDA kernel also uses (bit-comparable).
    Inputs are the probabilities widened exactly to float64."""
    pg = np.ascontiguousarray(pg, dtype=np.float64)
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    N = pg.shape[0]
    K = 8 if pa is None else 16
    if pa is not None:
        pa = np.ascontiguousarray(pa, dtype=np.float64)
    M = np.empty((N, K), dtype=np.float64)
    _lib().fg_oracle_cost_matrix(pg.ctypes.data, pr.ctypes.data, pa.ctypes.data if pa is not None else None, N, K, M.ctypes.data)
    return M
