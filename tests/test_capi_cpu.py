"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mtg_b200.h declares,
its host-only layout code matches the oracle's restatement of setupConstraintReorderingMatrix, and
it refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mav_trajectory_generation_b200 as m
from mav_trajectory_generation_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = m.load()
    header = open(os.path.join(ROOT, "include", "mtg_b200.h")).read()
    declared = set(re.findall(r"\b(mtg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mtg_version() >= 100


def test_layout_counts_baseline_configs():
    # n_all / n_fixed / n_free of C1..C4 and bytes per trajectory
    for (N, r, K, D), want, nbytes in [((10, 4, 2, 3), (20, 11, 4), 760), ((10, 4, 8, 3), (80, 17, 28), 2392),
                                       ((10, 4, 16, 3), (160, 25, 60), 4568), ((8, 3, 4, 3), (32, 11, 9), 1064)]:
        p = m.Problem(N, r, K, D)
        assert (p.n_all, p.n_fixed, p.n_free) == want
        assert p.bytes_per_trajectory == nbytes
        assert p.kernel == m.KERNEL_WAYPOINT


def test_layout_matches_oracle_for_random_masks(oracle):
    rng = np.random.RandomState(3)
    for trial in range(40):
        N = int(rng.choice([4, 6, 8, 10, 12]))
        h = N // 2
        K = int(rng.randint(1, 9))
        D = int(rng.randint(1, 4))
        mask = (rng.rand(K + 1, h) < 0.5).astype(np.uint8)
        mask[:, 0] = 1
        mask[0, :] = 1  # keeps R_pp positive definite
        values = rng.uniform(-1, 1, size=(K + 1, h, D)) * mask[:, :, None]
        times = rng.uniform(1.0, 4.0, size=K)
        res = oracle.solve(N, h - 1, mask, values, times)
        p = m.Problem(N, h - 1, K, D, fixed_mask=mask)
        assert (p.n_fixed, p.n_free) == (res["n_fixed"], res["n_free"])
        np.testing.assert_array_equal(p.slot_col, res["slot_col"])
        if res["n_free"] == 0:
            assert p.kernel == m.KERNEL_NOFREE


def test_routing():
    assert m.Problem(10, 4, 1, 3).kernel == m.KERNEL_NOFREE          # single fully constrained segment
    assert m.Problem(10, 4, 50, 3).kernel == m.KERNEL_WAYPOINT       # twisted sweep state still fits
    assert m.Problem(10, 4, 100, 3).kernel == m.KERNEL_WAYPOINT      # any K: chunked (checkpoint + recompute) kernel
    assert m.Problem(10, 4, 1000, 3).kernel == m.KERNEL_WAYPOINT
    assert m.Problem(10, 1, 16, 3).kernel == m.KERNEL_GENERIC        # (N, r) without a specialised kernel
    mask = np.zeros((5, 5), dtype=np.uint8)
    mask[:, 0] = 1
    mask[0, :] = 1
    mask[-1, :3] = 1
    assert m.Problem(10, 4, 4, 3, fixed_mask=mask).kernel == m.KERNEL_GENERIC
    wp = np.zeros((5, 5), dtype=np.uint8)
    wp[:, 0] = 1
    wp[0, :] = 1
    wp[-1, :] = 1
    assert m.Problem(10, 4, 4, 3, fixed_mask=wp).kernel == m.KERNEL_WAYPOINT  # explicit mask, same topology


@pytest.mark.parametrize("N,r,K,D", [(9, 3, 4, 3), (14, 4, 4, 3), (10, 5, 4, 3), (10, -1, 4, 3), (10, 4, 0, 3),
                                     (10, 4, 4, 0)])
def test_bad_problems_are_rejected_not_aborted(N, r, K, D):
    # the reference CHECK-aborts (impl/polynomial_optimization_linear_impl.h:60); the ABI returns a code
    with pytest.raises(ValueError):
        m.Problem(N, r, K, D)


def test_no_cpu_fallback():
    """Without a CUDA device mtg_create must fail loudly; with one, it must succeed."""
    import torch
    if torch.cuda.is_available():
        m.Solver(0).close()
    else:
        with pytest.raises(RuntimeError):
            m.Solver(0)
    lib = m.load()
    # compute entry points reject a null handle instead of computing anything on the host
    p = m.Problem(10, 4, 2, 3)
    assert lib.mtg_solve_linear_batch_f64(None, C.byref(p.c), 1, None, None, None, None, None, None) != 0
    assert lib.mtg_solve_linear_batch_host_f64(None, C.byref(p.c), 1, None, None, None, None, None) != 0


def test_option_constants_match_the_header():
    """capi.OPT_* are the MTG_OPT_* of include/mtg_b200.h (the Python binding restates them by hand)."""
    import re
    text = open(os.path.join(ROOT, "include", "mtg_b200.h")).read()
    header = {name: int(val) for name, val in re.findall(r"#define\s+MTG_OPT_(\w+)\s+(\d+)", text)}
    assert header, "no MTG_OPT_ definitions found"
    values = sorted(header.values())
    assert len(set(values)) == len(values), "duplicate option keys in the header"
    for name, val in header.items():
        assert getattr(capi, "OPT_" + name) == val, name
