"""The MEX gateway matlab/cfs_mex.cpp, compiled against a stand-in mex.h (tests/mex_stub/) and called with the arguments the
drop-in classdefs matlab/CFS_FANUC.m / PSGCFS_FANUC.m pass: MATLAB is not installed here, so this is how the gateway's
marshalling (robot / obs / sys_info structs, batched columns, outputs) gets compiled and executed at all."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib
from tests import common

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "mex_stub")], stdout=subprocess.DEVNULL)
    h = C.CDLL(os.path.join(HERE, "mex_stub", "libmexharness.so"))
    h.mexh_last_error.restype = C.c_char_p
    h.mexh_cfs.restype = C.c_int
    yield h
    h.mexh_shutdown()


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def call_mex(h, solver, grad, ROBOT, robot, obs, s, x0, ff, caug, xref, noise=None, drop=()):
    """mirrors cfs_mex(solver, grad, ROBOT, obs, sys_info [, noise]); x0 (B,2nj) ff (B,n) caug (B,) xref (B,2n) noise (B,K,n)"""
    nj, H = int(s["njoint"]), int(s["H"])
    n, B, K = nj * H, x0.shape[0], int(s["MAX_O_ITER"])
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    DH = np.asfortranarray(robot["DH"], dtype=np.float64)
    cap = np.asfortranarray(np.stack([np.asarray(robot["cap"][i]["p"], dtype=np.float64)[:, :2] for i in range(nj)], axis=2))
    T = None if robot.get("T") is None else np.asfortranarray(robot["T"], dtype=np.float64)
    seg = np.asfortranarray(np.stack([np.asarray(o["l"], dtype=np.float64) for o in obs], axis=2))
    D, eps = f([o["D"] for o in obs]), f([o["epsilon"] for o in obs])
    lim = None if (s.get("lim") is None or "lim" in drop) else f(s["lim"])
    mi = None if (s.get("MAX_input") is None or "MAX_input" in drop) else f(s["MAX_input"])
    out = dict(u=np.zeros((B, n)), x=np.zeros((B, 2 * n)), cost_hist=np.zeros((B, K)), e_u_hist=np.zeros((B, K)),
               iters=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32))
    qp = C.c_double(0.0)
    keep = [DH, cap, T, seg, D, eps, lim, mi, f(s["QQ"].T), f(x0), f(ff), f(caug), f(xref), None if noise is None else f(noise),
            f(np.asarray(robot["base"]).reshape(-1))]
    rc = h.mexh_cfs(solver.encode(), grad.encode(), ROBOT.encode(), C.c_int(nj), C.c_int(H), C.c_int(B), C.c_int(DH.shape[0]),
                    _dp(DH), _dp(keep[14]), _dp(cap), _dp(T), C.c_double(robot["delta_t"]), C.c_int(len(obs)), _dp(seg), _dp(D),
                    _dp(eps), _dp(keep[8]), _dp(lim), _dp(mi), _dp(keep[9]), _dp(keep[10]), _dp(keep[11]), _dp(keep[12]),
                    C.c_double(s["epsilon_O"]), C.c_int(K), C.c_double(s.get("alpha", 0.0)), _dp(keep[13]), _dp(out["u"]),
                    _dp(out["x"]), _dp(out["cost_hist"]), _dp(out["e_u_hist"]), _dp(out["iters"]), _dp(out["status"]), C.byref(qp))
    if rc:
        raise RuntimeError(h.mexh_last_error().decode())
    out["qp_steps"] = qp.value
    return out


def test_gateway_compiles_and_fails_loudly_without_a_gpu(harness):
    """not gpu: the gateway builds against the stub, links libcfs_b200.so, and without a usable device raises the MATLAB
    error id the classdefs document (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    ROBOT, robot, obs, s = common.main_fanuc_config()
    with pytest.raises(RuntimeError, match="cfs:cuda"):
        call_mex(harness, "CFS", "num_jac", ROBOT, robot, obs, s, s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]),
                 s["x_"][None])


@pytest.mark.gpu
def test_gateway_matches_the_ctypes_path(harness, ctx, oracle):
    # main_FANUC.m's configuration (M200i, H = 30, one problem), as CFS_FANUC.optimizer passes it
    ROBOT, robot, obs, s = common.main_fanuc_config()
    args = (s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]), s["x_"][None])
    r = dict(robot)
    r["name"] = ROBOT
    ctx.set_robot(r, 5)
    ctx.set_obstacles(obs)
    ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
    ref = ctx.solve_batch(*args, s["epsilon_O"], s["MAX_O_ITER"])
    out = call_mex(harness, "CFS", "num_jac", ROBOT, robot, obs, s, *args)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(out[k], ref[k]), k
    it = int(ref["iters"][0])
    assert np.array_equal(out["cost_hist"][0, :it], ref["cost_hist"][0, :it]) and out["qp_steps"] > 0
    # a batch (sys_info.xR / ff / caug / x_ with one column per problem), num_jac and derivest, M16iB
    cfg = common.batch_m16ib(oracle, 12, horizon=20)
    sb = dict(cfg["sys_info"])
    rb = dict(cfg["robot"])
    rb["name"] = "M16iB"
    ctx.set_robot(rb, 5)
    ctx.set_obstacles(cfg["obs"])
    ctx.set_cost(sb["H"], sb["QQ"], sb["lim"], sb["MAX_input"])
    bargs = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
    for grad, gm in (("num_jac", _lib.GRAD_NUMJAC), ("derivest", _lib.GRAD_DERIVEST)):
        ref = ctx.solve_batch(*bargs, sb["epsilon_O"], sb["MAX_O_ITER"], grad=gm)
        out = call_mex(harness, "CFS", grad, "M16iB", cfg["robot"], cfg["obs"], sb, *bargs)
        for k in ("u", "x", "iters", "status"):
            assert np.array_equal(out[k], ref[k]), (grad, k)
    # PSGCFS with the host-drawn noise (PSGCFS_FANUC.m:109), no bounds in the projection (:120)
    sb["alpha"] = 1.0 / np.linalg.svd(sb["QQ"], compute_uv=False).max()
    sb["MAX_O_ITER"] = 5
    noise = np.random.default_rng(4).normal(0.0, 0.1, size=(12, 5, sb["H"] * 5))
    ctx.set_cost(sb["H"], sb["QQ"], sb["lim"], None)
    ref = ctx.solve_batch(*bargs, sb["epsilon_O"], 5, solver=_lib.SOLVER_PSGCFS, noise=noise, alpha=sb["alpha"])
    out = call_mex(harness, "PSGCFS", "num_jac", "M16iB", cfg["robot"], cfg["obs"], sb, *bargs, noise=noise)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(out[k], ref[k]), ("psgcfs", k)
    # the script path of M16iB/main_CFS.m has no velocity rows: sys_info without .lim
    out = call_mex(harness, "CFS", "num_jac", "M16iB", cfg["robot"], cfg["obs"], cfg["sys_info"], *bargs, drop=("lim",))
    ctx.set_cost(sb["H"], sb["QQ"], None, cfg["sys_info"]["MAX_input"])
    ref = ctx.solve_batch(*bargs, sb["epsilon_O"], cfg["sys_info"]["MAX_O_ITER"])
    assert np.array_equal(out["u"], ref["u"]) and np.array_equal(out["status"], ref["status"])


@pytest.mark.gpu
def test_chomp_through_the_gateway(harness, oracle):
    """cfs_mex('solve', 'CHOMP', 'derivest', ROBOT, obs, sys_info, uu) -- the call matlab/CHOMP_FANUC.m makes -- returns what the
    ctypes path (cfs_chomp_batch) returns"""
    H, K, B = 20, 6, 5
    cfg = common.batch_m16ib(oracle, B, horizon=H, seed=3)
    s = dict(cfg["sys_info"], MAX_O_ITER=K)
    uu = 0.02 * np.random.default_rng(1).standard_normal((B, H * 5))
    c = M.Context(0)
    rb = dict(cfg["robot"])
    rb["name"] = "M16iB"
    c.set_robot(rb, 5)
    c.set_obstacles(cfg["obs"])
    c.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    ref = c.chomp_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], uu, float(s["alpha"]), K)
    out = call_mex(harness, "CHOMP", "derivest", "M16iB", cfg["robot"], cfg["obs"], s, cfg["x0"], cfg["ff"], cfg["caug"],
                   cfg["xref"], noise=uu)
    for k in ("u", "x", "cost_hist", "iters", "status"):
        assert np.array_equal(out[k], ref[k]), k
    with pytest.raises(RuntimeError, match="uu"):
        call_mex(harness, "CHOMP", "derivest", "M16iB", cfg["robot"], cfg["obs"], s, cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])


def _robot_args(robot, nj):
    DH = np.asfortranarray(robot["DH"], dtype=np.float64)
    cap = np.asfortranarray(np.stack([np.asarray(robot["cap"][i]["p"], dtype=np.float64)[:, :2] for i in range(nj)], axis=2))
    base = np.ascontiguousarray(np.asarray(robot["base"], dtype=np.float64).reshape(-1))
    return DH, cap, base


def _obs_args(obs):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    seg = np.asfortranarray(np.stack([np.asarray(o["l"], dtype=np.float64) for o in obs], axis=2))
    return seg, f([o["D"] for o in obs]), f([o["epsilon"] for o in obs])


def call_mex_rrt(h, ROBOT, SOLVER, robot, sc, rnd, max_iter=400):
    """mirrors cfs_mex('rrt', ROBOT, SOLVER, obs, sys_info, goal, region_g, region_s, sample_off, rnd): rnd (S, nrnd)"""
    nj, S, nrnd, cap = 5, rnd.shape[0], rnd.shape[1], max_iter + 2
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    DH, capp, base = _robot_args(robot, nj)
    seg, D, eps = _obs_args(sc["obs"])
    routes = np.zeros((S, cap, nj))
    ints = [np.zeros(S, dtype=np.int32) for _ in range(4)]
    parent, tot = np.zeros((S, cap), dtype=np.int32), np.zeros((S, cap))
    keep = [f(sc["x0"]), f(sc["goal"]), f(sc["ratial"]), f(sc["region_g"]), f(sc["region_s"]), f(sc["sample_off"]), f(rnd)]
    rc = h.mexh_rrt(ROBOT.encode(), SOLVER.encode(), C.c_int(nj), C.c_int(DH.shape[0]), _dp(DH), _dp(base), _dp(capp),
                    C.c_double(robot["delta_t"]), C.c_int(len(sc["obs"])), _dp(seg), _dp(D), _dp(eps), _dp(keep[0]), _dp(keep[1]),
                    _dp(keep[1]), _dp(keep[2]), _dp(keep[3]), _dp(keep[4]), _dp(keep[5]), C.c_int(nrnd), C.c_int(S), _dp(keep[6]),
                    C.c_int(max_iter), _dp(routes), _dp(ints[0]), _dp(ints[1]), _dp(ints[2]), _dp(ints[3]), _dp(parent), _dp(tot))
    if rc:
        raise RuntimeError(h.mexh_last_error().decode())
    return dict(routes=routes, route_len=ints[0], n_nodes=ints[1], fail=ints[2].astype(bool), rnd_used=ints[3], parent=parent,
                total_dis=tot)


def call_mex_routes(h, ROBOT, robot, obs, H, routes, route_len, Q, Rblk, r_scale, lim, max_input, eps_outer=0.1, max_outer=20):
    """mirrors cfs_mex('routes', ROBOT, obs, sys_info, routes, route_len, Q, Rblk, r_scale): routes (B, W, nj)"""
    nj, B, W = 5, routes.shape[0], routes.shape[1]
    n = nj * H
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    DH, capp, base = _robot_args(robot, nj)
    seg, D, eps = _obs_args(obs)
    out = dict(u=np.zeros((B, n)), x=np.zeros((B, 2 * n)), cost_hist=np.zeros((B, max_outer)), e_u_hist=np.zeros((B, max_outer)),
               iters=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32))
    keep = [f(routes), f(route_len), f(np.asarray(Q).T), f(np.asarray(Rblk).T), f(lim), f(max_input)]
    rc = h.mexh_routes(ROBOT.encode(), C.c_int(nj), C.c_int(H), C.c_int(B), C.c_int(W), C.c_int(DH.shape[0]), _dp(DH), _dp(base),
                       _dp(capp), C.c_double(robot["delta_t"]), C.c_int(len(obs)), _dp(seg), _dp(D), _dp(eps), _dp(keep[4]),
                       _dp(keep[5]), C.c_double(eps_outer), C.c_int(max_outer), _dp(keep[0]), _dp(keep[1]), _dp(keep[2]), _dp(keep[3]),
                       C.c_double(r_scale), _dp(out["u"]), _dp(out["x"]), _dp(out["cost_hist"]), _dp(out["e_u_hist"]),
                       _dp(out["iters"]), _dp(out["status"]))
    if rc:
        raise RuntimeError(h.mexh_last_error().decode())
    return out


def test_gateway_validates_its_inputs(harness):
    """not gpu: malformed arguments raise cfs:arg errors (no crash, no device needed: validation comes first or the create
    error is raised) -- unknown robot, PSGCFS without noise, empty batch, wrong sizes."""
    ROBOT, robot, obs, s = common.main_fanuc_config()
    args = (s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]), s["x_"][None])
    with pytest.raises(RuntimeError, match="cfs:arg.*unknown command"):
        call_mex(harness, "STOMP", "num_jac", ROBOT, robot, obs, s, *args)
    with pytest.raises(RuntimeError, match="cfs:arg.*grad"):
        call_mex(harness, "CFS", "hessian", ROBOT, robot, obs, s, *args)
    bad = dict(s)
    bad["H"] = 0
    with pytest.raises(RuntimeError, match="cfs:arg"):
        call_mex(harness, "CFS", "num_jac", ROBOT, robot, obs, bad, *args)
    import torch
    if not torch.cuda.is_available():
        return
    with pytest.raises(RuntimeError, match="cfs:arg.*unknown ROBOT"):
        call_mex(harness, "CFS", "num_jac", "M900", robot, obs, s, *args)
    with pytest.raises(RuntimeError, match="cfs:arg.*PSGCFS needs"):
        call_mex(harness, "PSGCFS", "num_jac", ROBOT, robot, obs, s, *args)


@pytest.mark.gpu
def test_rrtstar_cfs_through_the_gateway_only(harness, ctx, oracle):
    """RRTstar_CFS.m as matlab/s_Parallel_rrt.m + matlab/RRT_FANUC.m + matlab/CFS_FANUC.m drive it: cfs_mex('rrt') for all
    seeds, min(routeL) in the host, cfs_mex('routes') for the CFS stage -- nothing but mexFunction; checked against the ctypes
    path and, for the winner, against orc_rrt_find_route + the oracle's CFS."""
    from motionplanning_5d_m_b200 import problem, rrt
    sc = rrt.SCENE_RRTSTAR
    robot = M.robotproperty2("M200i")
    S, H = 12, 40
    rnd = np.random.default_rng(31).random((S, rrt.NRND_DEFAULT))
    assert harness.mexh_device(0) == 0
    out = call_mex_rrt(harness, "M200i", "RRT", robot, sc, rnd)
    r = dict(robot)
    r["name"] = "M200i"
    ctx.set_robot(r, 5)
    ctx.set_obstacles(sc["obs"])
    tile = lambda v: np.tile(np.asarray(v, dtype=np.float64)[None], (S, 1))
    ref = ctx.rrt_find_routes(tile(sc["x0"]), tile(sc["goal"]), tile(sc["goal"]), sc["region_g"], sc["region_s"], sc["sample_off"],
                              sc["ratial"], rnd, star=False, want_tree=True)
    assert np.array_equal(out["route_len"], ref["route_len"]) and np.array_equal(out["fail"], ref["fail"])
    assert np.array_equal(out["n_nodes"], ref["n_nodes"]) and np.array_equal(out["rnd_used"], ref["rnd_used"])
    for k in range(S):
        nn = int(ref["n_nodes"][k])
        assert np.array_equal(out["parent"][k, :nn], ref["parent"][k, :nn]) and np.array_equal(out["total_dis"][k, :nn], ref["total_dis"][k, :nn])
    routeL = np.where(out["fail"] | (out["route_len"] < 0), 1000, out["route_len"])   # s_Parallel_rrt.m:14,21
    best = int(np.argmin(routeL))                                                      # :27
    assert routeL[best] < 1000
    for k in range(S):
        assert np.array_equal(out["routes"][k, :max(out["route_len"][k], 0)], ref["routes"][k])
    # CFS stage for every seed through cfs_mex('routes'): failed seeds come back as NO_ROUTE
    W = int(routeL[routeL < 1000].max())
    lim, mi = np.ones(5), np.tile(np.array([1, 1, np.pi, np.pi, np.pi]) * robot["delta_t"], H)
    rl = np.where(routeL < 1000, routeL, 0).astype(np.float64)
    sol = call_mex_routes(harness, "M200i", robot, sc["obs"], H, np.ascontiguousarray(out["routes"][:, :W]), rl, problem.Q_RRTSTAR,
                          problem.R_MAIN_FANUC, 10.0, lim, mi)
    assert ((sol["status"][routeL == 1000] & 0xFF) == _lib.STATUS_NO_ROUTE).all()
    ctx.set_cost_blocks(H, problem.Q_RRTSTAR, problem.R_MAIN_FANUC, 10.0, lim, mi)
    ref_sol = ctx.solve_routes_var(np.ascontiguousarray(out["routes"][:, :W]), rl.astype(np.int32), 0.1, 20)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(sol[k], ref_sol[k]), k
    # the winner against the CPU restatements
    o = oracle.rrt_find_route(oracle.robot("M200i"), [q["l"] for q in sc["obs"]], [q["D"] for q in sc["obs"]], sc["x0"], sc["goal"],
                              sc["region_g"], sc["region_s"], sc["sample_off"], sc["goal"], sc["ratial"], rnd[best], star=False)
    assert np.abs(o["route"] - out["routes"][best, :routeL[best]]).max() < 1e-13
    _, _, obs_, sb = common.rrtstar_route_config(o["route"].T)
    P = common.oracle_problem(oracle, "M200i", obs_, sb)
    orc = P.solve_batch(sb["xR"][:, 0][None], sb["ff"][None], np.array([sb["caug"]]), sb["x_"][None])
    assert int(sol["status"][best]) == int(orc["status"][0]) and int(sol["iters"][best]) == int(orc["iters"][0])
    if (int(orc["status"][0]) & 0xFF) < 2:
        assert np.abs(sol["x"][best] - orc["x"][0]).max() < 1e-6
    # the same call again: robot / obstacles / cost are served from the gateway's content-hash cache, identical answers
    again = call_mex_routes(harness, "M200i", robot, sc["obs"], H, np.ascontiguousarray(out["routes"][:, :W]), rl, problem.Q_RRTSTAR,
                            problem.R_MAIN_FANUC, 10.0, lim, mi)
    assert np.array_equal(again["x"], sol["x"])
