"""Developer experiment: lead (in outward-sweep steps) of the single-buffer tile refill of the TMA-input kernel."""
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
    for K, B in ((16, 262144), (16, 1048576), (14, 262144), (12, 262144)):
        prob = m.Problem(10, 4, K, 3)
        times, dfix = synth(10, K, 3, B, dev)
        out = torch.zeros((B, K, 3, 10), device=dev, dtype=torch.float64)
        ref = None
        for E in (-1, 0):
            s.set_option(m.capi.OPT_EARLY_REFILL, E)
            for _ in range(3):
                s.solve_linear(prob, times, dfix, coeffs=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                s.solve_linear(prob, times, dfix, coeffs=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(ref, out))
            print(json.dumps(dict(K=K, B=B, E=E, ms=round(ms, 4),
                                  frac=round(B / (ms * 1e-3) * prob.bytes_per_trajectory / 1e9 / 6575.4, 4), bitwise_equal_to_off=same)), flush=True)
        del out, times, dfix, ref


if __name__ == "__main__":
    main()
