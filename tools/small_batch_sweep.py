"""Developer experiment: the C5 shard sizes (1 048 576 / N trajectories x 16 segments) under the waypoint-kernel
variants and the tile-scheduling options -- what one rank runs in the compute phase of bench.py --gpus N."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mav_trajectory_generation_b200 as m
from tools.quick_bench import synth


def main():
    dev = torch.device("cuda:0")
    s = m.Solver(0)
    N, r, K, D = 10, 4, 16, 3
    prob = m.Problem(N, r, K, D)
    # (variant, dynamic tiles, CTA cap)
    variants = [(0, 0, 0), (0, 1, 0), (0, 2, 0), (3, 0, 0), (0, 1, 1), (0, 2, 1)]
    for B in (65536, 131072, 262144, 524288):
        times, dfix = synth(N, K, D, B, dev)
        out = torch.zeros((B, K, D, N), device=dev, dtype=torch.float64)
        for variant, dyn, cap in variants:
            s.set_option(1, variant)
            s.set_option(5, dyn)
            s.set_option(3, cap)
            for _ in range(5):
                s.solve_linear(prob, times, dfix, coeffs=out)
            torch.cuda.synchronize()
            reps = 40
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                s.solve_linear(prob, times, dfix, coeffs=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rate = B / (ms * 1e-3)
            print(json.dumps(dict(B=B, variant=variant, dyn=dyn, cap=cap, ms=round(ms, 4), traj_per_s=round(rate),
                                  frac_hbm=round(rate * prob.bytes_per_trajectory / 1e9 / 6575.4, 4))), flush=True)


if __name__ == "__main__":
    main()
