"""Developer experiment: HBM-roofline fraction against the number of segments K (N=10, r=4, D=3) for the default routing
and for the chunked kernel forced -- where the resident-factor kernels should hand over to checkpoint + recompute."""
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
    for K in (2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 34, 36, 40, 50, 64, 100):
        B = min(262144, (3 << 30) // (K * 240 + 400) // 18944 * 18944)
        prob = m.Problem(10, 4, K, 3)
        times, dfix = synth(10, K, 3, B, dev)
        out = torch.zeros((B, K, 3, 10), device=dev, dtype=torch.float64)
        row = dict(K=K, B=B)
        for label, variant, chunk in (("default", 0, 0), ("chunked", 5, 0), ("v3", 3, 0)):
            if label == "chunked" and K < 12:
                continue
            s.set_option(1, variant)
            s.set_option(6, chunk)
            try:
                for _ in range(3):
                    s.solve_linear(prob, times, dfix, coeffs=out)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    s.solve_linear(prob, times, dfix, coeffs=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                row[label] = round(B / (ms * 1e-3) * prob.bytes_per_trajectory / 1e9 / 6575.4, 4)
            except Exception as e:
                row[label] = "error: " + str(e)[:80]
        s.set_option(1, 0)
        print(json.dumps(row), flush=True)
        del out, times, dfix


if __name__ == "__main__":
    main()
