#!/usr/bin/env python3
"""bench.py -- min-snap trajectories/sec of the batched solveLinear() hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one batch of synthetic random-waypoint trajectories:
BASELINE.json config C3 (262 144 trajectories, 16 segments, 3-D, N=10 min-snap, fp64) per GPU.
Multi-GPU = independent shards, no data-path collective ("scaling": "weak"); time is the max
over ranks.  Prints ONE JSON line (rank 0).

  value      whole-job trajectories/s, inputs/outputs resident in HBM (CUDA events on the
             launching stream, barrier + synchronize on both sides).
  e2e        same metric through the host-pointer C-ABI call (what
             PolynomialOptimization<N>::solveLinear() / BatchPolynomialOptimization calls):
             pinned HOST buffers, H2D + kernels + D2H inside the timed region.
  roofline   HBM: algorithmic bytes/launch (4568 B/trajectory) / kernel time,
             against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (restatement of the reference's Eigen path, oracle/oracle.cpp)
             on all host cores, on a bounded sample of the same workload (rank 0, N=1 only).

--impl reference times that CPU restatement itself (the reference needs Eigen/glog, absent
from this image, so oracle/_ref cannot exist): kind "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (N, r, K, D, batch per GPU)
    "C3": (10, 4, 16, 3, 262144),
    "C2": (10, 4, 8, 3, 65536),
    "C4": (8, 3, 4, 3, 1048576),
}
METRIC = "min-snap trajectories/sec (N=10, 16-seg, 3D)"
UNIT = "trajectories/s"


def workload_name(cfg, B):
    N, r, K, D, _ = CONFIGS[cfg]
    return f"{cfg}: batch {B} random-waypoint {K}-segment {D}D N={N} derivative_to_optimize={r} fp64 per GPU"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(cfg):
    """dram bytes per launch from the committed ncu --set full capture, or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(cfg)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region.  The region is tens of milliseconds,
    far shorter than nvidia-smi's sampling period, so NVML is polled from a thread (nvidia_ml_py); falls
    back to one nvidia-smi query if NVML is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80))

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.reason_bits = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._handle = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def _poll(self):
        p = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(p.nvmlDeviceGetClockInfo(self._handle, p.NVML_CLOCK_SM)))
                self.reason_bits |= int(p.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
            except Exception:
                try:
                    self.reason_bits |= int(p.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
                except Exception:
                    pass
            time.sleep(0.002)

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            sm = sorted(self.samples)
            reasons = [name for name, bit in self.REASONS if self.reason_bits & bit]
            if sm:
                return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                        "samples": len(sm), "source": "nvml"}
        try:  # fallback: a single nvidia-smi query right after the region
            out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                  "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b, c = [x.strip() for x in out.strip().split(",")[:3]]
            return {"sm_mhz": float(a), "sm_max_mhz": float(b), "reasons": [c], "samples": 1, "source": "nvidia-smi"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}


def synth_batch(torch, N, K, D, B, device, seed):
    """Random-waypoint batch with the distribution of createRandomVertices(box +-10) +
    estimateSegmentTimesNfabian(v=3, a=5) (reference vertex.cpp:27-82, :255-272), generated on the
    device with a counter-based generator (the bit-exact mt19937 fixture is what tests/ use)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pos = torch.rand((B, K + 1, D), generator=g, device=device, dtype=torch.float64) * 20.0 - 10.0
    dist = (pos[:, 1:] - pos[:, :-1]).norm(dim=2).clamp_min(0.2)
    v, a = 3.0, 5.0
    times = dist / v * 2 * (1.0 + 6.5 * v / a * torch.exp(-dist / v * 2))
    h = N // 2
    nf = 2 * h + K - 1
    dfix = torch.zeros((B, D, nf), device=device, dtype=torch.float64)
    dfix[:, :, 0] = pos[:, 0]
    dfix[:, :, h:h + K - 1] = pos[:, 1:K].transpose(1, 2)
    dfix[:, :, h + K - 1] = pos[:, K]
    return pos, times.contiguous(), dfix.contiguous()


def _best_threads(O, N, r, K, D):
    """Probe the oracle with all hardware threads and with half of them (SMT siblings often do not help
    this allocation-heavy code); return (threads, trajectories/s) of the faster setting."""
    hw = O.hardware_threads() or os.cpu_count() or 1
    best = (hw, 0.0)
    for threads in sorted({hw, max(1, hw // 2)}, reverse=True):
        probe = max(512, 64 * threads)
        pos, times = O.make_waypoint_batch(K, D, probe, base_seed=1000)
        O.solve_waypoint_batch(N, r, pos, times, n_threads=threads, mode=0, want_coeffs=False)  # warm
        _, s = O.solve_waypoint_batch(N, r, pos, times, n_threads=threads, mode=0, want_coeffs=False)
        rate = probe / max(s, 1e-9)
        if rate > best[1]:
            best = (threads, rate)
    return best


def cpu_baseline_sample(N, r, K, D, target_seconds=12.0, max_traj=2000000):
    """Times the CPU oracle (kind 'port') on a bounded sample, best of {all, half} host threads."""
    import oracle_lib as O
    O.build()
    threads, rate = _best_threads(O, N, r, K, D)
    n = int(min(max_traj, max(4096, rate * target_seconds)))
    pos, times = O.make_waypoint_batch(K, D, n, base_seed=1000)
    _, s = O.solve_waypoint_batch(N, r, pos, times, n_threads=threads, mode=0, want_coeffs=False)
    # the nonlinear optimiser's inner step only (updateSegmentTimes + solveLinear on a set-up object,
    # reference polynomial_optimization_nonlinear_impl.h:569-570): max over threads of the clocked time
    n1 = min(n, 200000)
    _, s1 = O.solve_waypoint_batch(N, r, pos[:n1], times[:n1], n_threads=threads, mode=1, want_coeffs=False)
    return {"value": n / s, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} trajectories of the workload (mt19937 fixture seeds 1000+b), construct+setupFromVertices+"
                      f"solveLinear per trajectory, {threads} host threads, {s:.2f} s wall",
            "update_times_plus_solve_only": n1 / s1}, n, s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N, r, K, D, B = CONFIGS[args.config]
    if args.batch:
        B = args.batch
    import oracle_lib as O
    O.build()
    threads, rate = _best_threads(O, N, r, K, D)
    # bounded sample per step: the whole --steps/--warmup run is sized to ~60 s of CPU work
    per_step_s = max(0.05, 60.0 / max(1, args.steps + args.warmup))
    per_step = int(max(2048, min(400000, rate * per_step_s)))
    pos, times = O.make_waypoint_batch(K, D, per_step, base_seed=1000)
    for _ in range(args.warmup):
        O.solve_waypoint_batch(N, r, pos, times, n_threads=threads, mode=0, want_coeffs=False)
    t = 0.0
    for _ in range(args.steps):
        _, s = O.solve_waypoint_batch(N, r, pos, times, n_threads=threads, mode=0, want_coeffs=False)
        t += s
    value = per_step * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, B), "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} trajectories per step (bounded sample of the workload), CPU restatement "
                                   f"of the reference Eigen path (Eigen/glog absent: reference itself unbuildable), "
                                   f"{threads} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_scatter(args, torch, dist, m, solver, prob, dev, world, rank):
    """BASELINE C5: the whole batch lives on rank 0; scatter inputs, solve shards, gather coefficients."""
    from mav_trajectory_generation_b200 import sharding
    N, r, K, D = prob.N, prob.r, prob.K, prob.D
    total = args.total
    if world == 1:
        dist_ok = False
    else:
        dist_ok = True
    if rank == 0:
        _, times_root, dfix_root = synth_batch(torch, N, K, D, total, dev, seed=99)
    else:
        times_root = dfix_root = None

    def solve_fn(t, f):
        return solver.solve_linear(prob, t, f)

    def step():
        if dist_ok:
            return sharding.solve_scattered(solve_fn, times_root, dfix_root, total, K, D, N, prob.n_fixed, dev)
        return solve_fn(times_root, dfix_root)

    for _ in range(max(args.warmup, 3)):
        out = step()
    if dist_ok:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    if dist_ok:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist_ok:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    # compute-only time of one shard (for the scatter/gather share)
    per = (total + world - 1) // world
    _, ts, fs = synth_batch(torch, N, K, D, per, dev, seed=7 + rank)
    buf = torch.empty((per, K, D, N), dtype=torch.float64, device=dev)
    for _ in range(3):
        solver.solve_linear(prob, ts, fs, coeffs=buf)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(10):
        solver.solve_linear(prob, ts, fs, coeffs=buf)
    c1.record()
    torch.cuda.synchronize()
    tc = torch.tensor([c0.elapsed_time(c1) / 10], dtype=torch.float64, device=dev)
    if dist_ok:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    if rank == 0:
        nbytes_link = (world - 1) / world * total * prob.bytes_per_trajectory
        line = {"metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C5: batch {total} random-waypoint {K}-segment {D}D N={N} fp64 sharded over "
                                       f"{world} GPU(s), NCCL scatter + solve + gather timed", "mode": "scatter"},
                "compute_only_ms": float(tc.item()), "compute_only_value": total / (float(tc.item()) * 1e-3),
                "nvlink_bytes_per_step": int(nbytes_link),
                "nvlink_GBps_root": nbytes_link / max(ms - float(tc.item()), 1e-9) / 1e6,
                "results_finite": bool(torch.isfinite(out).all().item())}
        print(json.dumps(line))
    if dist_ok:
        dist.destroy_process_group()
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import mav_trajectory_generation_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, r, K, D, B = CONFIGS[args.config]
    if args.batch:
        B = args.batch
    prob = m.Problem(N, r, K, D)
    solver = m.Solver(local)
    if args.mode == "scatter":
        return run_scatter(args, torch, dist, m, solver, prob, dev, world, rank)
    _, times, dfix = synth_batch(torch, N, K, D, B, dev, seed=1234 + rank)
    coeffs = torch.empty((B, K, D, N), dtype=torch.float64, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value")
    for _ in range(max(args.warmup, 3)):
        solver.solve_linear(prob, times, dfix, coeffs=coeffs, status=status)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = solver.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        starts[i].record()
        solver.solve_linear(prob, times, dfix, coeffs=coeffs, status=status)
        stops[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = solver.launch_count - launches0
    total_ms = starts[0].elapsed_time(stops[-1])
    kern_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    ok = bool((status == 0).all().item()) and bool(torch.isfinite(coeffs).all().item())
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---------------- end-to-end through the host-pointer C-ABI ("e2e")
    h_times = torch.empty((B, K), dtype=torch.float64).pin_memory()
    h_dfix = torch.empty((B, D, prob.n_fixed), dtype=torch.float64).pin_memory()
    h_coeffs = torch.empty((B, K, D, N), dtype=torch.float64).pin_memory()
    h_status = torch.empty((B,), dtype=torch.int32).pin_memory()
    h_times.copy_(times)
    h_dfix.copy_(dfix)
    e2e_steps = max(1, min(args.steps, 5))
    solver.solve_linear_host(prob, h_times, h_dfix, h_coeffs, status=h_status)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.solve_linear_host(prob, h_times, h_dfix, h_coeffs, status=h_status)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    e2e_value = world * B * e2e_steps / t_e2e
    e2e_ok = bool((h_status == 0).all().item()) and bool(torch.allclose(h_coeffs.to(dev), coeffs, rtol=0, atol=0))

    # ---------------- "next" row 8f-1: fused Nfabian time allocation + packing (positions in)
    fused_value = None
    try:
        pos_d = synth_batch(torch, N, K, D, B, dev, seed=1234 + rank)[0].contiguous()
        for _ in range(3):
            solver.solve_waypoints_nfabian(N, r, pos_d, 3.0, 5.0, 6.5, coeffs=coeffs)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nf_steps = max(1, min(args.steps, 20))
        f0.record()
        for _ in range(nf_steps):
            solver.solve_waypoints_nfabian(N, r, pos_d, 3.0, 5.0, 6.5, coeffs=coeffs)
        f1.record()
        torch.cuda.synchronize()
        fused_value = B * nf_steps / (f0.elapsed_time(f1) * 1e-3)
    except Exception as e:  # optional extra, never fails the bench
        fused_value = f"failed: {e}"

    if world > 1:
        dist.barrier()
    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_per_launch = prob.bytes_per_trajectory * B
        achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, B), "global_batch": world * B,
                       "parallelism": f"shard{world}" if world > 1 else "single",
                       "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2; no flush needed" %
                             (bytes_per_launch / 1e6),
                       "kernel": {1: "waypoint", 2: "generic", 3: "nofree"}[prob.kernel]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.config), "peak_source": peak_src,
                         "bytes_per_trajectory": prob.bytes_per_trajectory, "kernel_ms": kern_ms},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(8 * B * (K + D * prob.n_fixed)),
                    "d2h_bytes_per_step": int(8 * B * K * D * N + 4 * B), "steps": e2e_steps,
                    "bitwise_equal_to_device_path": e2e_ok},
            "gpu_launches": int(launches), "clocks": clocks, "results_ok": ok,
            "fused_waypoint_entry_traj_per_s_rank0": fused_value,
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, _, _ = cpu_baseline_sample(N, r, K, D)
                line["cpu_baseline"] = cb
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="resident", choices=["resident", "scatter"],
                    help="resident: every rank owns its shard in HBM (default, weak scaling, no collective); "
                         "scatter: BASELINE C5 -- rank 0 holds --total trajectories, NCCL scatter + solve + gather "
                         "inside the timed region (strong scaling)")
    ap.add_argument("--total", type=int, default=1048576, help="total trajectories for --mode scatter")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
