"""2-rank check of the fused solve + gather over NVLink peer memory (sharding.peer_solve_into_root) and of the NCCL
scatter/solve/gather pipeline: both must reproduce, bit for bit, a single-GPU solve of the whole batch on the root.
Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_gather_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import mav_trajectory_generation_b200 as m
from mav_trajectory_generation_b200 import sharding
from tools.quick_bench import synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    solver = m.Solver(local)
    peer_variant = int(os.environ.get("MTG_PEER_VARIANT", "-1"))  # diagnostic: force a kernel for the peer-store path
    ok = True
    for N, r, K, D, total in ((10, 4, 16, 3, 20011), (10, 4, 8, 3, 4096), (10, 4, 50, 3, 1000)):
        prob = m.Problem(N, r, K, D)
        if rank == 0:
            times, dfix = synth(N, K, D, total, dev, seed=3)
            out_peer = torch.zeros((total, K, D, N), dtype=torch.float64, device=dev)
            out_nccl = torch.zeros((total, K, D, N), dtype=torch.float64, device=dev)
            out_dma = torch.zeros((total, K, D, N), dtype=torch.float64, device=dev)
            want = solver.solve_linear(prob, times, dfix)
        else:
            times = dfix = out_peer = out_nccl = out_dma = want = None

        def solve_fn(t, f, c):
            solver.solve_linear(prob, t, f, coeffs=c)

        if peer_variant >= 0:
            solver.set_option(m.capi.OPT_WAYPOINT_VARIANT, peer_variant)
        t_pb = sharding.PeerBuffer(solver, times, 0)
        f_pb = sharding.PeerBuffer(solver, dfix, 0)
        o_pb = sharding.PeerBuffer(solver, out_peer, 0)
        sharding.peer_solve_into_root(solver, prob, t_pb, f_pb, o_pb, total, dev)
        torch.cuda.synchronize()
        solver.set_option(m.capi.OPT_WAYPOINT_VARIANT, 0)
        d_pb = sharding.PeerBuffer(solver, out_dma, 0)
        dbufs = {}
        for _ in range(2):  # twice: the second step reuses the buffers and streams of the first
            sharding.peer_dma_solve_gather(solver, prob, t_pb, f_pb, d_pb, total, dev, chunks=3, local_buffers=dbufs)
        torch.cuda.synchronize()
        if rank == 1:
            print(f"rank 1: peer-store solve K={K} completed (variant {peer_variant})", flush=True)
        dist.barrier()
        sharding.scatter_solve_gather(solve_fn, times, dfix, out_nccl, total, K, D, N, prob.n_fixed, dev, chunks=3)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            torch.cuda.synchronize()
            a, b, c = bool(torch.equal(out_peer, want)), bool(torch.equal(out_nccl, want)), bool(torch.equal(out_dma, want))
            print(f"K={K} total={total}: peer-store path bitwise {a}, NCCL path bitwise {b}, peer-DMA pipeline bitwise {c}")
            ok = ok and a and b and c
        for pb in (t_pb, f_pb, o_pb, d_pb):
            pb.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if not bool(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("peer_gather_check ok")


if __name__ == "__main__":
    main()
