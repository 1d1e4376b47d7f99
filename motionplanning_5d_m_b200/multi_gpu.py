"""Multi-GPU plumbing for the CFS hot path: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests).

The path shards with NO data-path collective: every (start, goal) problem, PSGCFS noise sample and RRT seed is
independent.  The only exchange is at the end:

  gather_results  -- all-gather of per-problem {cost, iters, status, u, x} when every rank needs the whole batch;
  best_of         -- the GPU analogue of `[~,id] = min(routeL)` over parfor workers (Lib/functions/s_Parallel_rrt.m:27):
                     ranks hold alternative solutions of the SAME problems (different noise streams / seeds / routes);
                     all-gather of (cost, status) (16 B per problem), argmin over ranks, then one all-reduce in which only
                     the winning rank contributes its trajectory (x + 0 is exact; the one value that does not survive is the
                     sign of a zero: -0.0 + 0.0 = +0.0).  A candidate whose cost is NaN or infinite is never admissible.
"""
import math

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(B, rank, world_size):
    """Contiguous block of ceil(B/G) problems per rank (the last ranks may get fewer or none)."""
    per = math.ceil(B / world_size) if world_size > 0 else B
    lo = min(rank * per, B)
    return lo, min(lo + per, B)


def gather_results(local, B, group=None):
    """local: dict name -> tensor whose first axis is this rank's shard (shard_bounds order).  Returns the same dict with
    first axis B on every rank.  Shards are padded to ceil(B/G) rows so that one all_gather_into_tensor per field suffices."""
    rank, G = world()
    if G == 1:
        return dict(local)
    per = math.ceil(B / G)
    out = {}
    for name, t in local.items():
        pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        full = torch.empty((G * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, pad.contiguous(), group=group)
        out[name] = full[:B]
    return out


def final_cost(cost_hist, iters):
    """cost of the last completed outer iteration per problem (cost_all(end)); +inf where no iteration completed."""
    it = iters.long()
    last = torch.gather(cost_hist, 1, (it.clamp(min=1) - 1)[:, None])[:, 0]
    return torch.where(it > 0, last, torch.full_like(last, float("inf")))


def best_of(cost, status, payload=None, group=None):
    """cost (P,), status (P,) int: this rank's result for each of P problems.  A candidate is admissible when its status
    low byte is 0 (converged) or 1 (MAX_ITER).  Returns (winner_rank (P,) long [-1: no admissible candidate], best_cost (P,),
    winner_payload dict or None); ties go to the lowest rank, like MATLAB's min()."""
    rank, G = world()
    ok = ((status.to(torch.int64) & 0xFF) < 2) & torch.isfinite(cost)
    c = torch.where(ok, cost, torch.full_like(cost, float("inf")))
    if G == 1:
        win = torch.where(ok, torch.zeros_like(status, dtype=torch.int64), torch.full_like(status, -1, dtype=torch.int64))
        return win, c, (dict(payload) if payload is not None else None)
    flat = torch.empty((G * c.shape[0],), dtype=c.dtype, device=c.device)
    dist.all_gather_into_tensor(flat, c.contiguous(), group=group)
    best, win = flat.view(G, c.shape[0]).min(dim=0)  # first minimum = lowest rank
    win = torch.where(torch.isinf(best), torch.full_like(win, -1), win)
    out = None
    if payload is not None:
        out = {}
        mine = win == rank
        for name, t in payload.items():
            mask = mine.view((-1,) + (1,) * (t.dim() - 1))
            contrib = torch.where(mask, t, torch.zeros_like(t))
            dist.all_reduce(contrib, op=dist.ReduceOp.SUM, group=group)
            out[name] = contrib
    return win, best, out
