"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
headline TMEM kernel (plain + fused Nfabian), shared-memory twisted kernel, generic kernel through the
3-stream host pipeline with pinned buffers, back-substitution, cost, evaluate, Mellinger."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mav_trajectory_generation_b200 as m  # noqa: E402
import oracle_lib as O  # noqa: E402

NO_TMEM = "--no-tmem" in sys.argv  # synccheck of CUDA 12.9 mis-reads tcgen05.alloc as an uninitialised mbarrier
s = m.Solver(0)
N, r, K, D, B = 10, 4, 16, 3, 200
pos, times = O.make_waypoint_batch(K, D, B, base_seed=1000)
ref, _ = O.solve_waypoint_batch(N, r, pos, times, n_threads=4)
prob = m.Problem(N, r, K, D)
t_d, f_d = torch.from_numpy(times).cuda(), torch.from_numpy(O.waypoint_d_fixed(N, pos)).cuda()
for variant in ((2,) if NO_TMEM else (3, 4, 6, 5, 2)):
    s.set_option(m.capi.OPT_WAYPOINT_VARIANT, variant)
    st = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    out = s.solve_linear(prob, t_d, f_d, status=st)
    torch.cuda.synchronize()
    err = (np.abs(out.cpu().numpy() - ref).reshape(B, -1).max(1) / np.abs(ref).reshape(B, -1).max(1)).max()
    assert err < 1e-10 and bool((st == 0).all()), (variant, err)
s.set_option(m.capi.OPT_WAYPOINT_VARIANT, 2 if NO_TMEM else 0)
s.set_option(m.capi.OPT_MELLINGER_UNFUSED, 1 if NO_TMEM else 0)
if not NO_TMEM:
    # default routing on a multiple-of-16 batch larger than one pass of the persistent grid: the TMA-input kernel with
    # one tile buffer, its early refill (mbarrier + cp.async.bulk + parked TMEM slots) and the dynamic tile counter
    Bb = 40000
    tb = t_d[torch.arange(Bb, device="cuda") % B].contiguous()
    fb = f_d[torch.arange(Bb, device="cuda") % B].contiguous()
    ob = s.solve_linear(prob, tb, fb)
    torch.cuda.synchronize()
    assert torch.equal(ob[:B], ob[B:2 * B]) and float((ob[:B] - out).abs().max()) < 1e-6
    o_r, n_r, _ = s.evaluate_range(t_d, out, 0.0, float(t_d.sum(dim=1).min()), 0.37, derivs=(0, 1, 2), max_samples=64)
if not NO_TMEM:
    s.solve_waypoints_nfabian(N, r, torch.from_numpy(pos).cuda(), 3.0, 5.0, 6.5)
cost = s.compute_cost(prob, t_d, out)
ev = s.evaluate(t_d, out, 1, 0.0, 0.5, 33)
c2, g2 = s.cost_gradient_mellinger(prob, t_d[:7].contiguous(), f_d[:7].contiguous())
# K = 50 (large-K path), an odd K, K = 8 with a multiple-of-16 batch (TMA-input kernel, double buffered)
for K2 in (() if NO_TMEM else (50, 7, 8)):
    p2, t2 = O.make_waypoint_batch(K2, D, 48, base_seed=3)
    s.solve_linear(m.Problem(N, r, K2, D), torch.from_numpy(t2).cuda(), torch.from_numpy(O.waypoint_d_fixed(N, p2)).cuda())
# generic mask through the pinned 3-stream host pipeline
h = N // 2
mask = np.zeros((7, h), dtype=np.uint8)
mask[:, :2] = 1
mask[0, :] = 1
mask[-1, :] = 1
pg = m.Problem(N, r, 6, D, fixed_mask=mask)
Bg = 9001
rng = np.random.RandomState(0)
tg = torch.from_numpy(rng.uniform(2, 6, size=(Bg, 6))).pin_memory()
fg = torch.from_numpy(rng.uniform(-2, 2, size=(Bg, D, pg.n_fixed))).pin_memory()
hg = torch.zeros((Bg, 6, D, N), dtype=torch.float64).pin_memory()
s.solve_linear_host(pg, tg, fg, hg)
dg = s.solve_linear(pg, tg.cuda(), fg.cuda())
torch.cuda.synchronize()
assert torch.equal(dg.cpu(), hg)
s.close()
print("sanitize_target ok")
