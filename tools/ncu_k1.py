"""Target for ncu: a few launches of one headline-kernel variant on C3 (262144 x 16 segments).
usage: ncu_k1.py VARIANT [RD CAP STAGGER_US DYN] [N r K D B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mav_trajectory_generation_b200 as m
from tools.quick_bench import synth

a = [int(x) for x in sys.argv[1:]]
variant = a[0] if a else 0
rd, cap, stag, dyn = (a[1:5] + [3, 0, 0, 0])[:4] if len(a) >= 5 else (3, 0, 0, 0)
N, r, K, D, B = a[5:10] if len(a) >= 10 else (10, 4, 16, 3, 262144)
s = m.Solver(0)
s.set_option(1, variant)
if variant == 4:
    s.set_option(2, rd)
    s.set_option(3, cap)
    s.set_option(4, stag)
    s.set_option(5, dyn)
dev = torch.device("cuda:0")
prob = m.Problem(N, r, K, D)
times, dfix = synth(N, K, D, B, dev)
out = torch.empty((B, K, D, N), device=dev, dtype=torch.float64)
for _ in range(5):
    s.solve_linear(prob, times, dfix, coeffs=out)
torch.cuda.synchronize()
print("done", variant, rd, cap, stag, dyn)
