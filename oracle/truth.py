"""High-precision (mpmath) ground truth for the linear min-derivative problem.

TEST INFRASTRUCTURE.  Independent of oracle.cpp and of the kernels' scaled-table formulation:
builds A(T), Q(T) literally (reference impl/polynomial_optimization_linear_impl.h:111-121,
:567-583), H = A^-T Q A^-1, R = C^T H C (:307-336), solves R_pp d_p = -R_pf d_f (:360-375) and
back-substitutes p = A^-1 C d (:262-283) -- all at `digits` decimal digits, rounded to fp64 at
the very end.  Used to show how far the reference-order fp64 arithmetic (oracle.cpp) and the
CUDA kernels each are from the exact answer.
"""
from math import factorial

import mpmath as mp
import numpy as np


def _B(k, j):
    return factorial(j) // factorial(j - k) if j >= k else 0


def layout(N, mask):
    """(slot_col [K*N], n_fixed, n_free) -- reference ordering (linear.h:287-295, impl :181-260)."""
    mask = np.asarray(mask, dtype=np.uint8)
    K = mask.shape[0] - 1
    h = N // 2
    flat = mask.reshape(-1)
    nf = int(flat.sum())
    col = np.zeros(flat.shape[0], dtype=np.int64)
    cf = cp = 0
    for i, mk in enumerate(flat):
        if mk:
            col[i] = cf
            cf += 1
        else:
            col[i] = nf + cp
            cp += 1
    slot = np.zeros(K * N, dtype=np.int64)
    for i in range(K):
        for s in range(N):
            v, k = (i, s) if s < h else (i + 1, s - h)
            slot[i * N + s] = col[v * h + k]
    return slot, nf, flat.shape[0] - nf


def solve(N, r, mask, values, times, digits=60):
    mp.mp.dps = digits
    mask = np.asarray(mask, dtype=np.uint8)
    values = np.asarray(values, dtype=np.float64)
    K = len(times)
    h = N // 2
    D = values.shape[2]
    slot, nf, npf = layout(N, mask)
    n = nf + npf
    R = mp.zeros(n, n)
    Ainvs = []
    for i in range(K):
        t = mp.mpf(float(times[i]))
        A = mp.zeros(N, N)
        Q = mp.zeros(N, N)
        for k in range(h):
            A[k, k] = _B(k, k)
            for j in range(k, N):
                A[h + k, j] = _B(k, j) * t ** (j - k)
        for a in range(r, N):
            for b in range(r, N):
                e = a + b - 2 * r + 1
                Q[a, b] = mp.mpf(2 * _B(r, a) * _B(r, b)) * t ** e / e
        Ai = mp.inverse(A)
        H = Ai.T * Q * Ai
        Ainvs.append(Ai)
        for a in range(N):
            for b in range(N):
                R[int(slot[i * N + a]), int(slot[i * N + b])] += H[a, b]
    d_all = mp.zeros(n, D)
    flat = mask.reshape(-1)
    cf = 0
    for idx, mk in enumerate(flat):
        if mk:
            v, k = divmod(idx, h)
            for d in range(D):
                d_all[cf, d] = mp.mpf(float(values[v, k, d]))
            cf += 1
    if npf > 0:
        Rpp = R[nf:, nf:]
        Rpf = R[nf:, :nf]
        rhs = -(Rpf * d_all[:nf, :])
        for d in range(D):
            x = mp.lu_solve(Rpp, rhs[:, d])
            for q in range(npf):
                d_all[nf + q, d] = x[q]
    coeffs = np.zeros((K, D, N))
    for i in range(K):
        for d in range(D):
            dv = mp.matrix([d_all[int(slot[i * N + s]), d] for s in range(N)])
            p = Ainvs[i] * dv
            coeffs[i, d, :] = [float(x) for x in p]
    d_free = np.array([[float(d_all[nf + q, d]) for q in range(npf)] for d in range(D)]).reshape(D, npf)
    return coeffs, d_free
