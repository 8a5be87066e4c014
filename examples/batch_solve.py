"""Batched solve through the C-ABI from Python (plumbing used by the tests / bench): 100 000 random-waypoint
16-segment trajectories, positions in -> coefficients out, on cuda:0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mav_trajectory_generation_b200 as m

N, r, K, D, B = 10, 4, 16, 3, 100_000
solver = m.Solver(0)
positions = torch.rand((B, K + 1, D), dtype=torch.float64, device="cuda") * 20.0 - 10.0
times = torch.empty((B, K), dtype=torch.float64, device="cuda")
status = torch.empty((B,), dtype=torch.int32, device="cuda")
# fused time allocation (estimateSegmentTimesNfabian, v_max 3 m/s, a_max 5 m/s^2) + solve
coeffs = solver.solve_waypoints_nfabian(N, r, positions, 3.0, 5.0, seg_times_out=times, status=status)
torch.cuda.synchronize()
print("solved", int((status == 0).sum()), "of", B, "| coeffs", tuple(coeffs.shape))
# sample the positions of every trajectory on a 0.5 s grid
samples = solver.evaluate(times, coeffs, derivative=0, t_start=0.0, dt=0.5, n_samples=32)
print("samples", tuple(samples.shape), "first trajectory starts at", samples[0, 0].tolist())
