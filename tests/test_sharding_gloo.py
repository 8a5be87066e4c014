"""world_size-2 gloo test (CPU) of the multi-GPU bookkeeping: shard bounds, scatter of the inputs,
gather of the outputs, ragged totals.  The local 'solve' is a deterministic stand-in computed with
torch on CPU -- what is under test is that trajectory b's inputs reach exactly one rank and its
outputs come back at row b (the GPU kernels themselves are covered by the -m gpu tests)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stand_in(times, dfix, K, D, N):
    # a per-trajectory function of the inputs only (no cross-trajectory coupling)
    B = times.shape[0]
    base = times.sum(dim=1, keepdim=True) + dfix.reshape(B, -1).sum(dim=1, keepdim=True)  # [B][1]
    grid = torch.arange(K * D * N, dtype=torch.float64).reshape(1, K, D, N)
    return base.reshape(B, 1, 1, 1) * (1.0 + grid)


def _worker(rank, world, port, total, out_q):
    sys.path.insert(0, ROOT)
    from mav_trajectory_generation_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, D, N, nf = 4, 3, 10, 13
    g = torch.Generator().manual_seed(5)
    times = torch.rand((total, K), generator=g, dtype=torch.float64) + 1.0
    dfix = torch.rand((total, D, nf), generator=g, dtype=torch.float64)
    want = _stand_in(times, dfix, K, D, N)
    got = sharding.solve_scattered(lambda t, f: _stand_in(t, f, K, D, N), times if rank == 0 else None,
                                   dfix if rank == 0 else None, total, K, D, N, nf, torch.device("cpu"))
    ok = True
    if rank == 0:
        ok = got.shape == want.shape and torch.equal(got, want)
    b = sharding.shard_bounds(total, world)
    ok = ok and b[0] == 0 and b[-1] == total and all(b[i] <= b[i + 1] for i in range(world))
    out_q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 101, 3])
def test_scatter_solve_gather_world2(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


def test_shard_bounds_balanced():
    sys.path.insert(0, ROOT)
    from mav_trajectory_generation_b200 import sharding
    for total in (0, 1, 7, 8, 1048576, 262145):
        for world in (1, 2, 4, 8):
            b = sharding.shard_bounds(total, world)
            sizes = [b[i + 1] - b[i] for i in range(world)]
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
