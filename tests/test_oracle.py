"""Pins the CPU oracle (oracle/oracle.cpp) to everything the reference's own tests hold for the
hot path (reference mav_trajectory_generation/test/test_polynomial_optimization.cpp), and to a
60-digit ground truth.  CPU only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import truth  # noqa: E402

N = 10

# OptimizationParams of the reference suite (test :790-867): D, derivative_to_optimize, K, seed, bounds, v, a
REF_PARAMS = [
    (1, 4, 1, 100, 10.0, 3.0, 5.0), (1, 4, 10, 102, 10.0, 3.0, 5.0), (1, 4, 50, 103, 10.0, 3.0, 5.0),
    (3, 4, 1, 104, 10.0, 3.0, 5.0), (3, 4, 10, 105, 10.0, 3.0, 5.0), (3, 4, 50, 106, 10.0, 3.0, 5.0),
    (1, 2, 5, 107, 10.0, 1.0, 2.0), (3, 2, 1, 108, 10.0, 1.0, 2.0), (3, 2, 5, 109, 10.0, 1.0, 2.0),
    (3, 3, 5, 110, 10.0, 1.0, 2.0),
]


def eval_poly(c, t, deriv):
    """Polynomial::evaluate (reference polynomial.h:134-149): Horner on the derivative coefficients."""
    n = len(c)
    acc = 0.0
    for j in range(n - 1, deriv - 1, -1):
        b = 1.0
        for k in range(deriv):
            b *= (j - k)
        acc = acc * t + b * c[j]
    return acc


def test_base_coefficients(oracle):
    """computeBaseCoefficients (polynomial.cpp:145-160): B(d,j) = j!/(j-d)!."""
    from math import factorial
    bc = oracle.base_coefficients()
    for d in range(12):
        for j in range(22):
            want = factorial(j) // factorial(j - d) if j >= d else 0
            assert bc[d, j] == float(want)


def test_mt19937_fixture(oracle):
    """createRandomVertices (vertex.cpp:27-82) with std::mt19937(105): first vertex."""
    pos = oracle.create_random_positions(10, 3, -10.0, 10.0, 105)
    np.testing.assert_array_equal(pos[0], [-3.4346932681103235, 7.0047000485169981, 2.9909075836621035])
    d = np.linalg.norm(np.diff(pos, axis=0), axis=1)
    assert (d > 0.2).all() and (np.abs(pos) <= 10.0).all()


def test_two_vertices_setup_matlab_golden(oracle):
    """TwoVerticesSetup (test :743-787): 1-D rest-to-rest 0 -> 5, T = 5, min snap; Matlab coefficients
    (:776-780).  CHECK_EIGEN_MATRIX_EQUAL_DOUBLE compares at fp64 test precision; the Matlab numbers
    are printed with 15 decimals, so 2e-14 absolute is the tightest meaningful bound."""
    mask = np.ones((2, 5), dtype=np.uint8)
    vals = np.zeros((2, 5, 1))
    vals[1, 0, 0] = 5.0
    res = oracle.solve(10, 4, mask, vals, np.array([5.0 * 2.0 / 2.0]))
    matlab = np.array([-0.000000000000004, 0.000000000000004, -0.000000000000006, 0.000000000000003,
                       -0.000000000000001, 0.201600000000015, -0.134400000000012, 0.034560000000004,
                       -0.004032000000000, 0.000179200000000])
    np.testing.assert_allclose(res["coeffs"][0, 0], matlab, rtol=0, atol=2e-14)
    assert res["n_free"] == 0 and res["n_fixed"] == 10


def test_a_matrix_inversion(oracle):
    """AMatrixInversion (test :731-741): structured inverse == A.inverse() to 1e-10, T = 1..60."""
    for t in range(1, 61):
        A = oracle.mapping_matrix(N, float(t))
        Ai = oracle.inverse_mapping_matrix(N, float(t))
        Ai_general = oracle.general_inverse(A)
        assert np.abs(Ai - Ai_general).max() <= 1e-10, t


def check_path(oracle, n_coeff, mask, values, times, coeffs, tol=1e-6):
    """checkPath (test :113-174): fixed constraints met at both segment ends and derivatives
    0..N/2-1 continuous at every interior vertex, to 1e-6."""
    K, D, _ = coeffs.shape
    h = n_coeff // 2
    for i in range(K):
        for (v, t) in ((i, 0.0), (i + 1, times[i])):
            for k in range(h):
                if mask[v, k]:
                    for d in range(D):
                        assert abs(eval_poly(coeffs[i, d], t, k) - values[v, k, d]) <= tol, (i, v, k, d)
        if i > 0:
            for k in range(h):
                for d in range(D):
                    a = eval_poly(coeffs[i - 1, d], times[i - 1], k)
                    b = eval_poly(coeffs[i, d], 0.0, k)
                    assert abs(a - b) <= tol, (i, k, d)


@pytest.mark.parametrize("D,r,K,seed,bounds,v_max,a_max", REF_PARAMS)
def test_unconstrained_linear_check_path(oracle, D, r, K, seed, bounds, v_max, a_max):
    """UnconstrainedLinearEstimateSegmentTimes (test :271-306): the reference's ten parameter sets."""
    pos = oracle.create_random_positions(K, D, -bounds, bounds, seed)
    times = oracle.nfabian(pos, v_max, a_max)
    mask, values = oracle.waypoint_problem(N, pos)  # vertices built with getHighestDerivativeFromN(N) (:75-77)
    res = oracle.solve(N, r, mask, values, times)
    check_path(oracle, N, mask, values, times, res["coeffs"])
    # computeCost vs numeric integration of the squared r-th derivative (checkCost, :176-197, 10 %)
    cost_numeric = 0.0
    for i in range(K):
        ts = np.linspace(0.0, times[i], 2001)
        for d in range(D):
            vals = np.array([eval_poly(res["coeffs"][i, d], t, r) for t in ts])
            cost_numeric += np.trapezoid(vals * vals, ts)
    assert abs(res["cost"] - cost_numeric) <= 0.1 * cost_numeric + 1e-12


@pytest.mark.parametrize("D,r,K,seed,bounds,v_max,a_max", REF_PARAMS)
def test_constraint_packing(oracle, D, r, K, seed, bounds, v_max, a_max):
    """ConstraintPacking (test :505-564): [d_f; d_p] -> p = A^-1 M d -> A p -> M^+ round trip (1e-6) and
    per-segment p == A^-1 M d; ends fixed only up to kMaxDerivative = r."""
    h = N // 2
    for i in range(5):
        pos = oracle.create_random_positions(K, D, -50.0, 50.0, 12345 + i)
        times = oracle.nfabian(pos, 3.0, 5.0)
        mask = np.zeros((K + 1, h), dtype=np.uint8)
        values = np.zeros((K + 1, h, D))
        mask[:, 0] = 1
        values[:, 0, :] = pos
        mask[0, : r + 1] = 1
        mask[-1, : r + 1] = 1
        res = oracle.solve(N, N // 2 - 1, mask, values, times)  # setupFromVertices default derivative (:526)
        slot = res["slot_col"]
        nf, npf = res["n_fixed"], res["n_free"]
        for d in range(D):
            d_all = np.concatenate([res["d_fixed"][d], res["d_free"][d]])
            recon = np.zeros(nf + npf)
            cnt = np.zeros(nf + npf)
            for s_i in range(K):
                A = oracle.mapping_matrix(N, times[s_i])
                Ai = oracle.inverse_mapping_matrix(N, times[s_i])
                p = Ai @ d_all[slot[s_i * N:(s_i + 1) * N]]
                np.testing.assert_allclose(p, res["coeffs"][s_i, d], rtol=0, atol=1e-6)
                d_un = A @ p
                np.add.at(recon, slot[s_i * N:(s_i + 1) * N], d_un)
                np.add.at(cnt, slot[s_i * N:(s_i + 1) * N], 1.0)
            np.testing.assert_allclose(recon / cnt, d_all, rtol=0, atol=1e-6)  # M_pinv = row-normalised M^T


def test_readme_example(oracle):
    """BASELINE config C1: README 3-vertex 3-D example (reference README.md:105-139), v = a = 2."""
    pos = np.array([[0.0, 0.0, 1.0], [1.0, 2.0, 3.0], [2.0, 1.0, 5.0]])
    times = oracle.nfabian(pos, 2.0, 2.0)
    np.testing.assert_allclose(times, [3.97084783, 3.82413014], rtol=0, atol=5e-9)
    mask, values = oracle.waypoint_problem(N, pos)
    res = oracle.solve(N, 4, mask, values, times)
    assert (res["n_fixed"], res["n_free"]) == (11, 4)
    want0 = [0, 0, 0, 0, 0, 1.339252819683e-02, -7.845546057916e-03, 1.954568943734e-03, -2.392908039981e-04,
             1.181394415329e-05]
    np.testing.assert_allclose(res["coeffs"][0, 0], want0, rtol=0, atol=2e-13)
    check_path(oracle, N, mask, values, times, res["coeffs"])


@pytest.mark.parametrize("n_coeff,r,K,seed", [(10, 4, 16, 1000), (10, 4, 8, 1001), (8, 3, 4, 1002), (10, 4, 2, 5)])
def test_oracle_vs_60_digit_truth(oracle, n_coeff, r, K, seed):
    """The reference-order fp64 arithmetic is within 1e-10 (global-relative) of the exact answer on
    the BASELINE fixtures; this is the floor any 1e-10 parity claim rests on."""
    pos = oracle.create_random_positions(K, 3, -10.0, 10.0, seed)
    times = oracle.nfabian(pos, 3.0, 5.0)
    mask, values = oracle.waypoint_problem(n_coeff, pos)
    res = oracle.solve(n_coeff, r, mask, values, times)
    tru, tru_free = truth.solve(n_coeff, r, mask, values, times)
    err = np.abs(res["coeffs"] - tru).max() / np.abs(tru).max()
    assert err <= 1e-10, err
    slot, nf, npf = truth.layout(n_coeff, mask)
    np.testing.assert_array_equal(slot, res["slot_col"])


def test_general_mask_vs_truth(oracle):
    """Arbitrary per-vertex masks (velocity fixed at an interior vertex, free end acceleration ...)."""
    rng = np.random.RandomState(7)
    K, D, h = 6, 2, 5
    pos = oracle.create_random_positions(K, D, -10.0, 10.0, 77)
    times = oracle.nfabian(pos, 3.0, 5.0)
    mask = np.zeros((K + 1, h), dtype=np.uint8)
    mask[:, 0] = 1
    mask[0, :] = 1
    mask[-1, :3] = 1
    mask[2, 1] = 1
    mask[4, 2] = 1
    values = rng.uniform(-1, 1, size=(K + 1, h, D)) * mask[:, :, None]
    values[:, 0, :] = pos
    res = oracle.solve(N, 4, mask, values, times)
    tru, _ = truth.solve(N, 4, mask, values, times)
    assert np.abs(res["coeffs"] - tru).max() / np.abs(tru).max() <= 1e-9
    check_path(oracle, N, mask, values, times, res["coeffs"])


@pytest.mark.parametrize("n_coeff,r,K,D,seed", [(10, 4, 16, 3, 1000), (10, 4, 8, 3, 1001), (8, 3, 4, 3, 1002), (12, 5, 6, 3, 3),
                                               (10, 2, 5, 1, 107), (6, 2, 5, 3, 9)])
def test_binary128_checker_vs_60_digit_truth(oracle, n_coeff, r, K, D, seed):
    """oracle/exact.cpp (binary128, the checker the GPU suite uses at scale) rounds to the same fp64 numbers
    as the 60-digit mpmath solve: <= 1 ulp of the largest coefficient, and its cost output agrees too."""
    B = 2
    pos, times = oracle.make_waypoint_batch(K, D, B, base_seed=seed)
    dfix = oracle.waypoint_d_fixed(n_coeff, pos)
    ex, ex_free, ex_cost = oracle.exact_solve_batch(n_coeff, r, times, dfix, n_threads=2, want_free=True, want_cost=True)
    for b in range(B):
        mask, values = oracle.waypoint_problem(n_coeff, pos[b])
        tru, tru_free = truth.solve(n_coeff, r, mask, values, times[b])
        assert np.abs(ex[b] - tru).max() <= 2.3e-16 * np.abs(tru).max()
        if tru_free.size:
            assert np.abs(ex_free[b] - tru_free).max() <= 2.3e-16 * np.abs(tru_free).max()
        res = oracle.solve(n_coeff, r, mask, values, times[b])
        assert abs(ex_cost[b] - res["cost"]) <= 1e-7 * abs(ex_cost[b])


def test_binary128_checker_general_mask(oracle):
    """Arbitrary masks (the generic kernel's checker): exact.cpp == truth.py to fp64 rounding."""
    rng = np.random.RandomState(4)
    n_coeff, K, D = 10, 4, 2
    h = n_coeff // 2
    mask = (rng.rand(K + 1, h) < 0.4).astype(np.uint8)
    mask[:, 0] = 1
    mask[0, :] = 1
    values = rng.uniform(-2, 2, size=(K + 1, h, D)) * mask[:, :, None]
    times = rng.uniform(2.0, 5.0, size=K)
    res = oracle.solve(n_coeff, h - 1, mask, values, times)
    ex = oracle.exact_solve_batch(n_coeff, h - 1, times[None], res["d_fixed"][None], mask=mask, n_threads=1)
    tru, _ = truth.solve(n_coeff, h - 1, mask, values, times)
    assert np.abs(ex[0] - tru).max() <= 2.3e-16 * np.abs(tru).max()


def test_cpu_info_is_sane(oracle):
    info = oracle.cpu_info()
    assert 1 <= info["effective"] <= info["affinity"] <= max(info["hardware_concurrency"], info["affinity"])
    assert oracle.hardware_threads() == info["effective"]
