"""world_size-2 and -3 gloo tests (CPU) of the multi-GPU bookkeeping: shard bounds, the chunked full-duplex
scatter / solve / gather pipeline (grouped send/recv straight into slices of the root's tensors), ragged totals
and totals smaller than world x chunks.  The local 'solve' is a deterministic stand-in computed with
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


def _worker(rank, world, port, total, chunks, out_q):
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
    calls = []

    def solve_fn(t, f, c):
        calls.append(t.shape[0])
        c.copy_(_stand_in(t, f, K, D, N))

    got = sharding.solve_scattered(solve_fn, times if rank == 0 else None, dfix if rank == 0 else None, total, K, D, N,
                                   nf, torch.device("cpu"), chunks=chunks)
    ok = True
    if rank == 0:
        ok = got.shape == want.shape and torch.equal(got, want)
    else:
        ok = got is None
    b0 = sharding.shard_bounds(total, world)
    ok = ok and sum(calls) == b0[rank + 1] - b0[rank]   # every owned trajectory solved exactly once
    b = sharding.shard_bounds(total, world)
    ok = ok and b[0] == 0 and b[-1] == total and all(b[i] <= b[i + 1] for i in range(world))
    out_q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total,chunks", [(2, 64, 4), (2, 101, 3), (2, 3, 4), (3, 50, 2), (2, 1, 1)])
def test_scatter_solve_gather(world, total, chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, chunks, q)) for r in range(world)]
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
            for r in range(world):
                c = sharding.chunk_bounds(b[r], b[r + 1], 4)
                assert c[0] == b[r] and c[-1] == b[r + 1] and all(c[i] <= c[i + 1] for i in range(4))
