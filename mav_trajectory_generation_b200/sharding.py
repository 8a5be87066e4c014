"""Multi-GPU sharding of a trajectory batch (one process per GPU, torch.distributed).

The batch is embarrassingly parallel (each trajectory is a closed problem; reference
impl/polynomial_optimization_linear_impl.h:338-379 touches only its own members), so ranks own
contiguous slices and the data path needs NO collective.  Collectives appear only when one rank
holds the whole batch: `scatter_inputs` / `gather_outputs` move (seg_times, d_fixed) out and the
coefficients back with NCCL (grouped send/recv under torch.distributed.scatter / gather; NVLink
on a B200 box).  The same code runs on the gloo backend with CPU tensors, which is how the
bookkeeping is tested without GPUs (tests/test_sharding_gloo.py).
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world):
    """Contiguous, balanced slices: rank r owns [bounds[r], bounds[r+1])."""
    return [total * r // world for r in range(world + 1)]


def _padded(total, world):
    return (total + world - 1) // world


def scatter_inputs(times_root, dfix_root, total, K, D, n_fixed, device, src=0):
    """Rank `src` passes the full [total][K] / [total][D][n_fixed] tensors (others pass None).
    Returns this rank's (times, d_fixed, count) where count <= rows are meaningful."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per = _padded(total, world)
    t_loc = torch.empty((per, K), dtype=torch.float64, device=device)
    f_loc = torch.empty((per, D, n_fixed), dtype=torch.float64, device=device)
    if rank == src:
        pad = per * world - total
        if pad:
            # pad with copies of the last trajectory so every rank solves a well-posed problem
            times_root = torch.cat([times_root, times_root[-1:].expand(pad, -1)])
            dfix_root = torch.cat([dfix_root, dfix_root[-1:].expand(pad, -1, -1)])
        t_list = [c.contiguous() for c in times_root.chunk(world)]
        f_list = [c.contiguous() for c in dfix_root.chunk(world)]
    else:
        t_list = f_list = None
    dist.scatter(t_loc, t_list, src=src)
    dist.scatter(f_loc, f_list, src=src)
    count = max(0, min(per, total - rank * per))
    return t_loc, f_loc, count


def gather_outputs(coeffs_loc, total, dst=0):
    """Inverse of scatter_inputs for the [per][K][D][N] coefficient tensors; returns the
    [total][K][D][N] tensor on `dst` (None elsewhere)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst:
        parts = [torch.empty_like(coeffs_loc) for _ in range(world)]
    else:
        parts = None
    dist.gather(coeffs_loc, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat(parts)[:total]


def solve_scattered(solve_fn, times_root, dfix_root, total, K, D, N, n_fixed, device, root=0):
    """scatter -> local solve -> gather.  solve_fn(times, d_fixed) -> coeffs [per][K][D][N]."""
    t_loc, f_loc, _ = scatter_inputs(times_root, dfix_root, total, K, D, n_fixed, device, src=root)
    c_loc = solve_fn(t_loc, f_loc)
    return gather_outputs(c_loc, total, dst=root)
