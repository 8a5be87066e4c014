"""ncu target: a few launches of ONE kernel family on one workload.
usage: ncu_target.py NAME   with NAME in: c3 (K=16 default routing), c5x1, c2, c4, k50, k100, mask, c3v3 (per-tile kernel),
                                          c3v6 (TMA-input kernel), mellinger"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mav_trajectory_generation_b200 as m
from tools.quick_bench import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
shapes = {"c3": (10, 4, 16, 3, 262144), "c3v3": (10, 4, 16, 3, 262144), "c3v6": (10, 4, 16, 3, 262144),
          "c5x1": (10, 4, 16, 3, 1048576), "c2": (10, 4, 8, 3, 65536), "c4": (8, 3, 4, 3, 1048576),
          "k50": (10, 4, 50, 3, 65536), "k100": (10, 4, 100, 3, 32768), "mask": (10, 4, 16, 3, 65536),
          "mellinger": (10, 4, 16, 3, 16384)}
N, r, K, D, B = shapes[name]
dev = torch.device("cuda:0")
s = m.Solver(0)
if name == "c3v3":
    s.set_option(m.capi.OPT_WAYPOINT_VARIANT, 3)
if name == "c3v6":
    s.set_option(m.capi.OPT_WAYPOINT_VARIANT, 6)
mask = None
if name == "mask":
    h = N // 2
    mask = np.zeros((K + 1, h), dtype=np.uint8)
    mask[:, :2] = 1
    mask[0, :] = 1
    mask[-1, :] = 1
prob = m.Problem(N, r, K, D, fixed_mask=mask)
if mask is None:
    times, dfix = synth(N, K, D, B, dev)
else:
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    times = torch.rand((B, K), generator=g, device=dev, dtype=torch.float64) * 4 + 2
    dfix = torch.rand((B, D, prob.n_fixed), generator=g, device=dev, dtype=torch.float64) * 4 - 2
out = torch.empty((B, K, D, N), device=dev, dtype=torch.float64)
for _ in range(5):
    if name == "mellinger":
        s.cost_gradient_mellinger(prob, times, dfix)
    else:
        s.solve_linear(prob, times, dfix, coeffs=out)
torch.cuda.synchronize()
print("done", name)
