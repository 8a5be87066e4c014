"""Error histogram of the CUDA path against the CPU oracle (and a 60-digit truth on a sample) on the
bit-exact mt19937 fixture of BASELINE.json (seeds 1000+b).  Writes profiles/r01_parity.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import mav_trajectory_generation_b200 as m  # noqa: E402
import oracle_lib as O  # noqa: E402
import truth  # noqa: E402


def main():
    s = m.Solver(0)
    out = {}
    for name, N, r, K, D, B in (("C3", 10, 4, 16, 3, 8192), ("C2", 10, 4, 8, 3, 8192), ("C4", 8, 3, 4, 3, 8192),
                                ("C1shape", 10, 4, 2, 3, 4096)):
        pos, times = O.make_waypoint_batch(K, D, B, base_seed=1000)
        ref, _ = O.solve_waypoint_batch(N, r, pos, times, n_threads=O.hardware_threads())
        prob = m.Problem(N, r, K, D)
        got = s.solve_linear(prob, torch.from_numpy(times).cuda(), torch.from_numpy(O.waypoint_d_fixed(N, pos)).cuda()).cpu().numpy()
        err = np.abs(got - ref).reshape(B, -1).max(1) / np.abs(ref).reshape(B, -1).max(1)
        edges = [0, 1e-14, 1e-13, 1e-12, 1e-11, 3e-11, 1e-10, 1e-9, 1]
        hist, _ = np.histogram(err, bins=edges)
        nt = 8
        e_gpu, e_orc = [], []
        for b in range(nt):
            mask, values = O.waypoint_problem(N, pos[b])
            tru, _ = truth.solve(N, r, mask, values, times[b])
            sc = np.abs(tru).max()
            e_gpu.append(float(np.abs(got[b] - tru).max() / sc))
            e_orc.append(float(np.abs(ref[b] - tru).max() / sc))
        out[name] = {"N": N, "r": r, "K": K, "D": D, "trajectories": B, "T_min": float(times.min()), "T_max": float(times.max()),
                     "gpu_vs_oracle": {"max": float(err.max()), "median": float(np.median(err)), "p99": float(np.quantile(err, 0.99)),
                                       "bin_edges": edges, "counts": hist.tolist()},
                     "vs_60_digit_truth_first_8": {"gpu_max": max(e_gpu), "oracle_max": max(e_orc)}}
        print(name, json.dumps(out[name]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
