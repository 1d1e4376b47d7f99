"""-m "not gpu": the N>1 host logic over gloo, world_size 2 (shards == single process; best-of picks the argmin)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import common


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from motionplanning_5d_m_b200 import multi_gpu
    O.build()
    B = 11  # ragged on purpose: shards of 6 and 5
    cfg = common.batch_m16ib(O, B, horizon=12)
    s = cfg["sys_info"]
    P = common.oracle_problem(O, "M16iB", cfg["obs"], s)
    lo, hi = multi_gpu.shard_bounds(B, rank, world)
    # stand-in for the CUDA solve of this rank's shard (the host logic under test does not care who solved it)
    r = P.solve_batch(cfg["x0"][lo:hi], cfg["ff"][lo:hi], cfg["caug"][lo:hi], cfg["xref"][lo:hi])
    local = {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ("u", "x", "cost_hist", "iters", "status")}
    full = multi_gpu.gather_results(local, B)
    # best-of: both ranks hold candidates for the same 7 problems
    rng = np.random.default_rng(100 + rank)
    cost = torch.from_numpy(rng.random(7))
    status = torch.tensor([0, 1, 2, 0, 2, 0, 256], dtype=torch.int32) if rank == 0 else torch.tensor(
        [0, 0, 0, 2, 2, 0, 1], dtype=torch.int32)
    traj = torch.from_numpy(rng.random((7, 5))) + 10 * rank
    win, best, pay = multi_gpu.best_of(cost, status, {"x": traj})
    if rank == 0:
        q.put({k: v.numpy() for k, v in full.items()})
        q.put((win.numpy(), best.numpy(), pay["x"].numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_equals_single_process_and_best_of(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=240)
    win, best, pay = q.get(timeout=60)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B = 11
    cfg = common.batch_m16ib(oracle, B, horizon=12)
    s = cfg["sys_info"]
    P = common.oracle_problem(oracle, "M16iB", cfg["obs"], s)
    ref = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(full[k], ref[k]), k  # N shards == 1 shard, bit for bit
    assert np.array_equal(np.isnan(full["cost_hist"]), np.isnan(ref["cost_hist"]))
    assert np.array_equal(np.nan_to_num(full["cost_hist"]), np.nan_to_num(ref["cost_hist"]))
    # best-of expectations
    c0, c1 = np.random.default_rng(100).random(7), np.random.default_rng(101).random(7)
    t0, t1 = np.random.default_rng(100), np.random.default_rng(101)
    t0.random(7), t1.random(7)
    x0, x1 = t0.random((7, 5)), t1.random((7, 5)) + 10
    ok0 = np.array([1, 1, 0, 1, 0, 1, 1], bool)
    ok1 = np.array([1, 1, 1, 0, 0, 1, 1], bool)
    e0, e1 = np.where(ok0, c0, np.inf), np.where(ok1, c1, np.inf)
    exp_win = np.where(np.isinf(np.minimum(e0, e1)), -1, np.where(e1 < e0, 1, 0))
    assert np.array_equal(win, exp_win)
    assert np.array_equal(best, np.minimum(e0, e1))
    for i in range(7):
        exp = x0[i] if exp_win[i] == 0 else x1[i] if exp_win[i] == 1 else np.zeros(5)
        assert np.array_equal(pay[i], exp)


def test_shard_bounds_cover_batch():
    from motionplanning_5d_m_b200 import multi_gpu
    for B in (0, 1, 7, 8, 4096, 4097):
        for G in (1, 2, 4, 8):
            spans = [multi_gpu.shard_bounds(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_final_cost_picks_last_iteration():
    from motionplanning_5d_m_b200 import multi_gpu
    ch = torch.tensor([[3.0, 2.0, float("nan")], [float("nan")] * 3, [5.0, 4.0, 1.0]])
    it = torch.tensor([2, 0, 3], dtype=torch.int32)
    fc = multi_gpu.final_cost(ch, it)
    assert fc[0] == 2.0 and torch.isinf(fc[1]) and fc[2] == 1.0


def test_best_of_ignores_non_finite_costs():
    """a NaN / Inf cost with an admissible status is never a winner (single process: the same masking runs before the gather)"""
    from motionplanning_5d_m_b200 import multi_gpu
    cost = torch.tensor([float("nan"), 1.0, float("inf"), 2.0], dtype=torch.float64)
    status = torch.tensor([0, 0, 1, 2], dtype=torch.int32)
    win, best, _ = multi_gpu.best_of(cost, status)
    assert win.tolist() == [-1, 0, -1, -1] and best[1] == 1.0 and torch.isinf(best[[0, 2, 3]]).all()
