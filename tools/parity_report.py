"""Error statistics of the CUDA path on the bit-exact mt19937 fixture of BASELINE.json (seeds 1000+b):
against the binary128 solve of the same equations (oracle/exact.cpp) and against the reference-order fp64
oracle (oracle/oracle.cpp), plus the oracle's own distance from exact.  Writes gpurun_out/r02_parity.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import mav_trajectory_generation_b200 as m  # noqa: E402
import oracle_lib as O  # noqa: E402

EDGES = [0, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 3e-11, 1e-10, 1e-9, 1]


def rel(a, b):
    n = a.shape[0]
    return np.abs(a - b).reshape(n, -1).max(1) / np.abs(b).reshape(n, -1).max(1)


def stats(e):
    hist, _ = np.histogram(e, bins=EDGES)
    return {"max": float(e.max()), "argmax": int(e.argmax()), "median": float(np.median(e)),
            "p99": float(np.quantile(e, 0.99)), "p999": float(np.quantile(e, 0.999)), "counts": hist.tolist()}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    s = m.Solver(0)
    out = {"bin_edges": EDGES}
    shapes = (("C3", 10, 4, 16, 3), ("C2", 10, 4, 8, 3), ("C4", 8, 3, 4, 3), ("C1shape", 10, 4, 2, 3),
              ("N12", 12, 5, 6, 3), ("D1", 10, 4, 16, 1), ("r3", 10, 3, 5, 3), ("K50", 10, 4, 50, 3))
    for name, N, r, K, D in shapes:
        nb = B if K <= 16 else max(256, B // 8)
        pos, times = O.make_waypoint_batch(K, D, nb, base_seed=1000)
        dfix = O.waypoint_d_fixed(N, pos)
        ref, _ = O.solve_waypoint_batch(N, r, pos, times, n_threads=O.hardware_threads())
        exact = O.exact_solve_batch(N, r, times, dfix)
        prob = m.Problem(N, r, K, D)
        got = s.solve_linear(prob, torch.from_numpy(times).cuda(), torch.from_numpy(dfix).cuda()).cpu().numpy()
        e_ge, e_go, e_oe = rel(got, exact), rel(got, ref), rel(ref, exact)
        ratio = e_ge / np.maximum(e_oe, 1e-300)
        worst = int(e_ge.argmax())
        out[name] = {"N": N, "r": r, "K": K, "D": D, "trajectories": nb, "T_min": float(times.min()),
                     "T_max": float(times.max()),
                     "gpu_vs_exact": stats(e_ge), "gpu_vs_oracle": stats(e_go), "oracle_vs_exact": stats(e_oe),
                     "gpu_closer_than_oracle_fraction": float((e_ge <= e_oe).mean()),
                     "max_ratio_gpu_err_over_oracle_err_where_gpu_above_1e-12":
                         float(ratio[e_ge > 1e-12].max()) if (e_ge > 1e-12).any() else 0.0,
                     "worst_trajectory": {"index": worst, "gpu_vs_exact": float(e_ge[worst]),
                                          "oracle_vs_exact": float(e_oe[worst]),
                                          "segment_times": [round(float(t), 3) for t in times[worst]]}}
        print(name, json.dumps(out[name]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
