"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")
_SO_EXACT = os.path.join(_ROOT, "oracle", "libexact.so")
_lib = None
_lib_exact = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    stale = force
    for so, src in ((_SO, "oracle.cpp"), (_SO_EXACT, "exact.cpp")):
        src = os.path.join(_ROOT, "oracle", src)
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            stale = True
    if stale:
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_solve.restype = C.c_int
        L.oracle_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _dp, _dp, _dp,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_mapping_matrix.argtypes = [C.c_int, C.c_double, _dp]
        L.oracle_inverse_mapping_matrix.argtypes = [C.c_int, C.c_double, _dp]
        L.oracle_general_inverse.argtypes = [C.c_int, _dp, _dp]
        L.oracle_general_inverse.restype = C.c_int
        L.oracle_cost_matrix.argtypes = [C.c_int, C.c_int, C.c_double, _dp]
        L.oracle_base_coefficients.argtypes = [_dp]
        L.oracle_create_random_positions.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_uint64, _dp]
        L.oracle_nfabian.argtypes = [C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_double, _dp]
        L.oracle_solve_waypoint_batch.restype = C.c_double
        L.oracle_solve_waypoint_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _dp, _dp,
                                                  C.c_void_p, C.c_int, C.c_int]
        L.oracle_make_waypoint_batch.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_uint64,
                                                 C.c_double, C.c_double, _dp, _dp]
        L.oracle_cost_gradient_mellinger.restype = C.c_int
        L.oracle_cost_gradient_mellinger.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.oracle_hardware_threads.restype = C.c_int
        L.oracle_evaluate_range.restype = C.c_int
        L.oracle_evaluate_range.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                            C.c_int, C.c_int, _dp, _dp]
        L.oracle_cpu_info.restype = C.c_int
        L.oracle_cpu_info.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def lib_exact():
    global _lib_exact
    if _lib_exact is None:
        build()
        L = C.CDLL(_SO_EXACT)
        L.exact_solve_batch.restype = C.c_int
        L.exact_solve_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, _dp, _dp, _dp,
                                        C.c_void_p, C.c_void_p, C.c_int]
        _lib_exact = L
    return _lib_exact


def exact_solve_batch(N, r, times, d_fixed, mask=None, n_threads=None, want_free=False, want_cost=False):
    """binary128 solve (oracle/exact.cpp): times [B][K], d_fixed [B][D][n_fixed] (reference compact order) ->
    coeffs [B][K][D][N] rounded once from ~34-digit arithmetic; with want_free / want_cost returns the tuple
    (coeffs, d_free [B][D][n_free] or None, cost [B] or None)."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    d_fixed = np.ascontiguousarray(d_fixed, dtype=np.float64)
    B, K = times.shape
    D = d_fixed.shape[1]
    h = N // 2
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint8).reshape(K + 1, h)
        n_fixed = int(mask.sum())
    else:
        n_fixed = 2 * h + K - 1
    assert d_fixed.shape == (B, D, n_fixed), (d_fixed.shape, (B, D, n_fixed))
    n_free = (K + 1) * h - n_fixed
    coeffs = np.zeros((B, K, D, N))
    dfree = np.zeros((B, D, max(n_free, 1))) if want_free else None
    cost = np.zeros(B) if want_cost else None
    rc = lib_exact().exact_solve_batch(N, r, K, D, mask.ctypes.data if mask is not None else None, B, times, d_fixed,
                                       coeffs, dfree.ctypes.data if want_free else None,
                                       cost.ctypes.data if want_cost else None, n_threads or hardware_threads())
    if rc != 0:
        raise RuntimeError(f"exact_solve_batch: rc={rc}")
    if want_free or want_cost:
        return coeffs, (dfree[:, :, :n_free] if want_free else None), cost
    return coeffs


def cpu_info():
    """dict(hardware_concurrency, affinity, cgroup_quota_cpus (0 = unlimited), effective)."""
    a = (C.c_int32 * 4)()
    lib().oracle_cpu_info(C.addressof(a))
    return dict(hardware_concurrency=a[0], affinity=a[1], cgroup_quota_cpus=a[2], effective=a[3])


def solve(N, r, mask, values, times, want_cost=False):
    """mask [K+1][h] uint8, values [K+1][h][D], times [K] -> dict(coeffs[K][D][N], d_fixed, d_free, slot_col, ...)"""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    values = np.ascontiguousarray(values, dtype=np.float64)
    times = np.ascontiguousarray(times, dtype=np.float64)
    K = times.shape[0]
    h = N // 2
    D = values.shape[2]
    assert mask.shape == (K + 1, h) and values.shape == (K + 1, h, D)
    n_fixed = int(mask.sum())
    n_free = (K + 1) * h - n_fixed
    coeffs = np.zeros((K, D, N))
    d_fixed = np.zeros((D, n_fixed))
    d_free = np.zeros((D, max(n_free, 1)))
    slot_col = np.zeros(K * N, dtype=np.int32)
    cost = C.c_double(0.0)
    counts = (C.c_int32 * 3)()
    rc = lib().oracle_solve(N, r, K, D, mask, values, times, coeffs,
                            d_fixed.ctypes.data, d_free.ctypes.data, slot_col.ctypes.data,
                            C.addressof(cost), C.addressof(counts))
    if rc != 0:
        raise RuntimeError(f"oracle_solve failed rc={rc}")
    assert counts[1] == n_fixed and counts[2] == n_free
    return dict(coeffs=coeffs, d_fixed=d_fixed, d_free=d_free[:, :n_free], slot_col=slot_col,
                cost=cost.value, n_fixed=n_fixed, n_free=n_free)


def waypoint_problem(N, positions):
    """positions [K+1][D] -> (mask, values) of the createRandomVertices topology (vertex.cpp:27-82)."""
    positions = np.asarray(positions, dtype=np.float64)
    K1, D = positions.shape
    h = N // 2
    mask = np.zeros((K1, h), dtype=np.uint8)
    values = np.zeros((K1, h, D))
    mask[:, 0] = 1
    values[:, 0, :] = positions
    mask[0, :] = 1
    mask[-1, :] = 1
    return mask, values


def create_random_positions(K, D, lo, hi, seed):
    out = np.zeros((K + 1, D))
    lo = np.full(D, lo, dtype=np.float64) if np.isscalar(lo) else np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.full(D, hi, dtype=np.float64) if np.isscalar(hi) else np.ascontiguousarray(hi, dtype=np.float64)
    lib().oracle_create_random_positions(K, D, lo, hi, seed, out)
    return out


def nfabian(positions, v_max, a_max, magic=6.5):
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    K = positions.shape[0] - 1
    out = np.zeros(K)
    lib().oracle_nfabian(K, positions.shape[1], positions, v_max, a_max, magic, out)
    return out


def mapping_matrix(N, T):
    out = np.zeros((N, N))
    lib().oracle_mapping_matrix(N, T, out)
    return out


def inverse_mapping_matrix(N, T):
    out = np.zeros((N, N))
    lib().oracle_inverse_mapping_matrix(N, T, out)
    return out


def general_inverse(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.zeros_like(A)
    assert lib().oracle_general_inverse(A.shape[0], A, out) == 0
    return out


def cost_matrix(N, r, T):
    out = np.zeros((N, N))
    lib().oracle_cost_matrix(N, r, T, out)
    return out


def base_coefficients():
    out = np.zeros((22, 22))
    lib().oracle_base_coefficients(out)
    return out


def solve_waypoint_batch(N, r, positions, times, n_threads=1, mode=0, want_coeffs=True):
    """positions [B][K+1][D], times [B][K] -> (coeffs [B][K][D][N] or None, seconds)."""
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    times = np.ascontiguousarray(times, dtype=np.float64)
    B, K1, D = positions.shape
    K = K1 - 1
    coeffs = np.zeros((B, K, D, N)) if want_coeffs else None
    secs = lib().oracle_solve_waypoint_batch(N, r, K, D, B, positions, times,
                                             coeffs.ctypes.data if want_coeffs else None, n_threads, mode)
    if secs < 0:
        raise RuntimeError("oracle batch solve failed")
    return coeffs, secs


def hardware_threads():
    return lib().oracle_hardware_threads()


def make_waypoint_batch(K, D, B, lo=-10.0, hi=10.0, base_seed=1000, v_max=3.0, a_max=5.0):
    """The BASELINE.json fixture: positions [B][K+1][D], times [B][K] (seed = base_seed + b)."""
    pos = np.zeros((B, K + 1, D))
    times = np.zeros((B, K))
    lib().oracle_make_waypoint_batch(K, D, B, lo, hi, base_seed, v_max, a_max, pos, times)
    return pos, times


def waypoint_d_fixed(N, positions, start_derivs=None, end_derivs=None):
    """positions [B][K+1][D] -> d_fixed [B][D][n_fixed] in the reference's compact order
    (x_0,u_0(1..h-1), x_1..x_{K-1}, x_K,u_K(1..h-1)); end derivatives default to zero
    (Vertex::makeStartOrEnd, vertex.cpp:147-153)."""
    positions = np.asarray(positions)
    B, K1, D = positions.shape
    K = K1 - 1
    h = N // 2
    nf = 2 * h + K - 1
    out = np.zeros((B, D, nf))
    out[:, :, 0] = positions[:, 0, :]
    if K > 1:
        out[:, :, h:h + K - 1] = np.transpose(positions[:, 1:K, :], (0, 2, 1))
    out[:, :, h + K - 1] = positions[:, K, :]
    if start_derivs is not None:  # [B][h-1][D]
        out[:, :, 1:h] = np.transpose(start_derivs, (0, 2, 1))
    if end_derivs is not None:
        out[:, :, h + K:] = np.transpose(end_derivs, (0, 2, 1))
    return out


def cost_gradient_mellinger(N, r, positions, times):
    """One trajectory: (cost, grad[K]) as PolynomialOptimizationNonLinear::getCostAndGradientMellinger."""
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    times = np.ascontiguousarray(times, dtype=np.float64)
    K = times.shape[0]
    cost = np.zeros(1)
    grad = np.zeros(K)
    rc = lib().oracle_cost_gradient_mellinger(N, r, K, positions.shape[1], positions, times, cost, grad)
    if rc != 0:
        raise RuntimeError(f"oracle_cost_gradient_mellinger rc={rc}")
    return float(cost[0]), grad


def evaluate_range(times, coeffs, t_start, t_end, dt, derivative, max_samples):
    """One trajectory: Trajectory::evaluateRange restated (reference src/trajectory.cpp:81-141).
    times [K], coeffs [K][D][N] -> (n, out [max_samples][D], sampling_times [max_samples])."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    K, D, N = coeffs.shape
    out = np.zeros((max_samples, D))
    st = np.zeros(max_samples)
    n = lib().oracle_evaluate_range(N, K, D, times, coeffs, t_start, t_end, dt, derivative, max_samples, out, st)
    return n, out, st
