"""-m gpu : C-ABI behaviour around the hot path -- edge cases, error paths, the asynchronous entry, option switches."""
import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib
from tests import common

pytestmark = pytest.mark.gpu


def _setup(ctx, cfg):
    s = cfg["sys_info"]
    r = dict(cfg["robot"])
    r["name"] = "M16iB"
    ctx.set_robot(r, 5)
    ctx.set_obstacles(cfg["obs"])
    ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
    return s


def test_empty_batch_and_single_problem(ctx, oracle):
    cfg = common.batch_m16ib(oracle, 3, horizon=12)
    s = _setup(ctx, cfg)
    out = ctx.solve_batch(cfg["x0"][:0], cfg["ff"][:0], cfg["caug"][:0], cfg["xref"][:0], s["epsilon_O"], s["MAX_O_ITER"])
    assert out["u"].shape == (0, 60) and out["status"].shape == (0,)
    one = ctx.solve_batch(cfg["x0"][:1], cfg["ff"][:1], cfg["caug"][:1], cfg["xref"][:1], s["epsilon_O"], s["MAX_O_ITER"])
    allb = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
    # a problem's result does not depend on what else is in the batch (bit for bit)
    assert np.array_equal(one["u"][0], allb["u"][0]) and np.array_equal(one["x"][0], allb["x"][0])
    assert one["iters"][0] == allb["iters"][0] and one["status"][0] == allb["status"][0]


def test_no_obstacles_is_the_unconstrained_or_bounded_qp(ctx, oracle):
    """obs = {} : get_con contributes nothing; with lim and MAX_input the QP still has velocity rows and bounds."""
    cfg = common.batch_m16ib(oracle, 4, horizon=10)
    s = cfg["sys_info"]
    r = dict(cfg["robot"])
    r["name"] = "M16iB"
    ctx.set_robot(r, 5)
    ctx.set_obstacles([])
    ctx.set_cost(s["H"], s["QQ"], None, None)
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], 3)
    u_unc = -np.linalg.solve(s["QQ"], cfg["ff"].T).T
    assert np.abs(out["u"] - u_unc).max() < 1e-9 * max(1.0, np.abs(u_unc).max())
    assert ((out["status"] & 0xFF) <= 1).all()


def test_max_outer_zero_and_one(ctx, oracle):
    cfg = common.batch_m16ib(oracle, 5, horizon=12)
    s = _setup(ctx, cfg)
    P0 = common.oracle_problem(oracle, "M16iB", cfg["obs"], dict(s, MAX_O_ITER=1))
    ref = P0.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], 1)
    assert (out["iters"] == ref["iters"]).all() and (out["status"] == ref["status"]).all()
    out0 = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], 0)
    assert (out0["iters"] == 0).all() and ((out0["status"] & 0xFF) == _lib.STATUS_MAX_ITER).all()
    assert np.array_equal(out0["x"], cfg["xref"]) and not out0["u"].any()


def test_async_entry_matches_blocking_entry(ctx, oracle):
    import torch
    cfg = common.batch_m16ib(oracle, 40, horizon=20)
    s = _setup(ctx, cfg)
    B, n, N, K = 40, 100, 200, int(s["MAX_O_ITER"])
    ref = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], K)
    hin = {k: torch.from_numpy(cfg[k]).pin_memory() for k in ("x0", "ff", "caug", "xref")}
    mk = lambda *sh, dt=torch.float64: torch.zeros(sh, dtype=dt).pin_memory()
    o = dict(u=mk(B, n), x=mk(B, N), cost=mk(B, K), eu=mk(B, K), iters=mk(B, dt=torch.int32), status=mk(B, dt=torch.int32))
    ctx.solve_batch_ptr(B, hin["x0"].data_ptr(), hin["ff"].data_ptr(), hin["caug"].data_ptr(), hin["xref"].data_ptr(),
                        s["epsilon_O"], K, o["u"].data_ptr(), o["x"].data_ptr(), o["cost"].data_ptr(), o["eu"].data_ptr(),
                        o["iters"].data_ptr(), o["status"].data_ptr(), device=False, sync=False)
    ctx.wait()
    assert np.array_equal(o["u"].numpy(), ref["u"]) and np.array_equal(o["x"].numpy(), ref["x"])
    assert np.array_equal(o["iters"].numpy(), ref["iters"]) and np.array_equal(o["status"].numpy(), ref["status"])
    assert ctx.stats()["problem_iters"] == int(ref["iters"].sum())


def test_escalation_threshold_does_not_change_results(ctx, oracle):
    """The heavy tier resumes escalated problems: forcing (nearly) every QP through it must give the same answers.
    The two tiers recover the primal point by different (mathematically identical) routes -- cached directions and prefix
    sums in the bulk tier, the Gram operator in the heavy tier -- so they agree to the parity tolerance, not bit for bit."""
    cfg = common.batch_m16ib(oracle, 64, horizon=30)
    s = _setup(ctx, cfg)
    a = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
    ctx.set_option("esc_steps", 2)
    b = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
    ctx.set_option("esc_steps", 48)
    assert (a["status"] == b["status"]).all() and (a["iters"] == b["iters"]).all()
    ok = (a["status"] & 0xFF) < 2
    dx, du = np.abs(a["x"][ok] - b["x"][ok]).max(), np.abs(a["u"][ok] - b["u"][ok]).max()
    assert dx < 1e-6 and du < 1e-6, (dx, du)


def test_error_paths(ctx, oracle):
    c2 = M.Context(0)
    with pytest.raises(M.CfsError, match="robot not set"):
        c2.dist_grad(np.zeros((1, 5)))
    cfg = common.batch_m16ib(oracle, 2, horizon=8)
    r = dict(cfg["robot"])
    r["name"] = "M16iB"
    c2.set_robot(r, 5)
    c2.set_obstacles(cfg["obs"])
    s = cfg["sys_info"]
    with pytest.raises(M.CfsError, match="cost not set"):
        c2.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 3)
    bad = -np.eye(40)
    with pytest.raises(M.CfsError, match="not positive definite"):
        c2.set_cost(8, bad, None, None)
    with pytest.raises(M.CfsError, match="unknown option"):
        c2.set_option("no_such_option", 1)
    c2.close()


def test_many_obstacles_and_large_horizon(ctx, oracle):
    """CFS_MAX_OBS obstacles and a horizon the fused kernel does not cover (falls back to the launch-per-iteration path)."""
    rng = np.random.default_rng(3)
    robot = M.robotproperty2("M16iB")
    r = dict(robot)
    r["name"] = "M16iB"
    ctx.set_robot(r, 5)
    obs = [{"l": np.array([[3.9 + 0.3 * rng.random(), 3.9 + 0.3 * rng.random()], [8.3, 8.3], [0.0, 1.5 + rng.random()]]),
            "D": 0.05, "epsilon": 0.05} for _ in range(32)]
    ctx.set_obstacles(obs)
    th = common.sampling_box(rng, 200)
    dist, lid, g, flags = ctx.dist_grad(th)
    rr = oracle.robot("M16iB")
    for j in (0, 13, 31):
        o6 = oracle.obs6(obs[j]["l"])
        dref = np.array([oracle.dist_arm(rr, t, o6)[0] for t in th])
        assert np.abs(dist[:, j] - dref).max() < 1e-12
    with pytest.raises(M.CfsError):
        ctx.set_obstacles(obs + obs)  # > CFS_MAX_OBS


def test_device_side_setup_matches_host_builder(ctx, oracle):
    """cfs_set_cost_blocks + cfs_solve_start_goal (main_FANUC.m:38-103 on the device) against the host-built problem."""
    from motionplanning_5d_m_b200 import problem
    cfg = common.batch_m16ib(oracle, 48, horizon=30)
    s = cfg["sys_info"]
    r = dict(cfg["robot"])
    r["name"] = "M16iB"
    ctx.set_robot(r, 5)
    ctx.set_obstacles(cfg["obs"])
    ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
    ref = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
    ctx.set_cost_blocks(s["H"], problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0, s["lim"], s["MAX_input"])
    out = ctx.solve_start_goal(cfg["theta0"], cfg["thetag"], s["epsilon_O"], s["MAX_O_ITER"])
    assert (out["status"] == ref["status"]).all() and (out["iters"] == ref["iters"]).all()
    ok = (ref["status"] & 0xFF) < 2
    assert np.abs(out["x"][ok] - ref["x"][ok]).max() < 1e-6 and np.abs(out["u"][ok] - ref["u"][ok]).max() < 1e-6
    for b in np.where(ok)[0]:
        it = ref["iters"][b]
        assert np.all(np.abs(out["cost_hist"][b, :it] - ref["cost_hist"][b, :it]) <= 1e-6 * np.abs(ref["cost_hist"][b, :it]))
    # and against the oracle fed with the host-built arrays
    P = common.oracle_problem(oracle, "M16iB", cfg["obs"], s)
    orc = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=8)
    assert (out["status"] == orc["status"]).all() and (out["iters"] == orc["iters"]).all()
    assert np.abs(out["x"][ok] - orc["x"][ok]).max() < 1e-6
    # x may be skipped
    nox = ctx.solve_start_goal(cfg["theta0"], cfg["thetag"], s["epsilon_O"], s["MAX_O_ITER"], want_x=False)
    assert nox["x"] is None and np.array_equal(nox["u"], out["u"])
    with pytest.raises(M.CfsError, match="cfs_set_cost_blocks"):
        ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
        ctx.solve_start_goal(cfg["theta0"], cfg["thetag"], s["epsilon_O"], s["MAX_O_ITER"])


def test_device_route_resampling_and_solve_routes(ctx, oracle):
    """cubicpolytraj resampling (RRTstar_CFS.m:96-100) + the CFS stage set-up on the device, against the host
    restatement problem.cubicpolytraj, the golden rrtstar_cfs case (shipped route data/200i_xori.mat) and the oracle on
    perturbed routes."""
    from motionplanning_5d_m_b200 import problem
    from tests.test_oracle import check_against_golden
    inp = common.golden("inputs.npz")
    route = np.asarray(inp["route_wp"], dtype=np.float64)            # (5, 16)
    ROBOT, robot, obs, s = common.rrtstar_route_config(route)
    r = dict(robot)
    r["name"] = ROBOT
    ctx.set_robot(r, 5)
    ctx.set_obstacles(obs)
    H = s["H"]
    rng = np.random.default_rng(11)
    routes = np.stack([route.T] + [route.T + 0.02 * rng.standard_normal(route.T.shape) * np.linspace(0, 1, 16)[:, None] *
                                   np.linspace(1, 0, 16)[:, None] * 4 for _ in range(7)])    # (8, W, 5), same end points
    # resampling alone (also with a horizon that does not divide the route, and W = 2)
    for Hs, rt in ((H, routes), (23, routes[:3]), (7, routes[:2, :2])):
        got = ctx.resample_routes(rt, Hs)
        W = rt.shape[1]
        for b in range(rt.shape[0]):
            ref = problem.cubicpolytraj(rt[b].T, np.arange(W) * 0.5, np.linspace(0, (W - 1) * 0.5, Hs + 1))
            assert np.abs(got[b].T - ref).max() < 1e-12
    assert np.array_equal(ctx.resample_routes(routes, H)[:, 0], routes[:, 0]) and \
        np.array_equal(ctx.resample_routes(routes, H)[:, -1], routes[:, -1])
    # solve: device glue against the golden case and against the oracle on host-built arrays
    ctx.set_cost_blocks(H, problem.Q_RRTSTAR, problem.R_MAIN_FANUC, 10.0, s["lim"], s["MAX_input"])
    out = ctx.solve_routes(routes, s["epsilon_O"], s["MAX_O_ITER"])
    check_against_golden("rrtstar_cfs", {k: v[:1] for k, v in out.items()}, common.golden("cases.npz"))
    for b in range(routes.shape[0]):
        _, _, _, sb = common.rrtstar_route_config(routes[b].T)
        P = common.oracle_problem(oracle, ROBOT, obs, sb)
        orc = P.solve_batch(sb["xR"][:, 0][None], sb["ff"][None], np.array([sb["caug"]]), sb["x_"][None])
        assert int(out["status"][b]) == int(orc["status"][0]) and int(out["iters"][b]) == int(orc["iters"][0])
        if (int(orc["status"][0]) & 0xFF) < 2:
            assert np.abs(out["x"][b] - orc["x"][0]).max() < 1e-6
            it = int(orc["iters"][0])
            assert np.all(np.abs(out["cost_hist"][b, :it] - orc["cost_hist"][0, :it]) <= 1e-6 * np.abs(orc["cost_hist"][0, :it]))
    assert ctx.solve_routes(routes[:0], s["epsilon_O"], s["MAX_O_ITER"])["u"].shape == (0, H * 5)


def test_batched_rrt_tree_growth_matches_the_restatement(ctx, oracle):
    """RRT_FANUC.find_route for many seeds at once (one CTA per seed) against the C restatement fed with the same
    uniform random streams: RRTstar_CFS.m's scene (M200i, two obstacles, :29-55), RRT* and RRT, plus a seed whose stream runs dry."""
    O = oracle
    ROBOT, robot, obs, s = common.rrtstar_cfs_config(np.zeros((5, 41)))
    r = dict(robot)
    r["name"] = ROBOT
    ctx.set_robot(r, 5)
    ctx.set_obstacles(obs)
    x0 = np.array([0.421, 0, -0.0092, -0.0010, -1.5786])                       # RRTstar_CFS.m:30
    goal = np.array([-1.4090, 0.8873, 0.4008, 0.0, 0.4430])                    # :33
    region_g = np.array([np.pi / 20, np.pi / 20, np.pi / 10, np.pi / 2, np.pi / 2])
    region_s = np.array([np.pi / 2, np.pi / 2, np.pi / 2, np.pi / 1.5, np.pi / 1.5])
    ratial, off = np.array([1, 1, 0.5, 0.1, 0.1]), np.zeros(5)
    S, R = 48, 12000
    rnd = np.random.default_rng(2026).random((S, R))
    rnd[S - 1, 40:] = 0.99      # this seed only ever samples the goal after 40 numbers: it stalls against the obstacle ...
    X0, G = np.tile(x0, (S, 1)), np.tile(goal, (S, 1))
    X0[1] = goal                 # ... and this one starts inside the goal box: route = x0, no sample drawn
    rob = O.robot(ROBOT)
    for star in (True, False):
        out = ctx.rrt_find_routes(X0, G, G, region_g, region_s, off, ratial, rnd, star=star, want_tree=True)
        assert out["route_len"][1] == 1 and out["rnd_used"][1] == 0 and np.array_equal(out["routes"][1][0], goal)
        n_ok = 0
        for k in range(S):
            ref = O.rrt_find_route(rob, [o["l"] for o in obs], [o["D"] for o in obs], X0[k], G[k], region_g, region_s, off, G[k],
                                   ratial, rnd[k], star=star)
            if ref is None:
                assert out["route_len"][k] == -1
                continue
            n_ok += 1
            assert out["route_len"][k] == len(ref["route"]) and out["n_nodes"][k] == ref["n_nodes"]
            assert bool(out["fail"][k]) == ref["fail"] and out["rnd_used"][k] == ref["rnd_used"]
            nn = ref["n_nodes"]
            assert np.array_equal(out["parent"][k, :nn], ref["parent"])
            assert np.abs(out["nodes"][k, :nn] - ref["nodes"]).max() < 1e-13
            assert np.abs(out["total_dis"][k, :nn] - ref["total_dis"]).max() < 1e-12
            assert np.abs(out["routes"][k] - ref["route"]).max() < 1e-13
        assert n_ok >= S - 2
        found = ~out["fail"] & (out["route_len"] > 0)
        assert found.any() and out["fail"].any()
        # s_Parallel_rrt.m:27: [~, id] = min(routeL) over the seeds that found a path
        rl = np.where(found, out["route_len"], 1000)
        assert rl.min() == min(len(rt) for rt, f in zip(out["routes"], found) if f)


def test_rrt_to_cfs_pipeline_on_the_device(ctx, oracle):
    """RRTstar_CFS.m end to end (s_Parallel_rrt -> min(routeL) -> cubicpolytraj -> CFS_FANUC.optimizer) with every stage on
    the GPU, checked stage by stage against the CPU restatements."""
    from motionplanning_5d_m_b200 import problem
    O = oracle
    ROBOT, robot, obs, s = common.rrtstar_cfs_config(np.zeros((5, 41)))
    r = dict(robot)
    r["name"] = ROBOT
    ctx.set_robot(r, 5)
    ctx.set_obstacles(obs)
    x0 = np.array([0.421, 0, -0.0092, -0.0010, -1.5786])
    goal = np.array([-1.4090, 0.8873, 0.4008, 0.0, 0.4430])
    region_g = np.array([np.pi / 20, np.pi / 20, np.pi / 10, np.pi / 2, np.pi / 2])
    region_s = np.array([np.pi / 2, np.pi / 2, np.pi / 2, np.pi / 1.5, np.pi / 1.5])
    S = 32
    rnd = np.random.default_rng(77).random((S, 12000))
    out = ctx.rrt_find_routes(np.tile(x0, (S, 1)), np.tile(goal, (S, 1)), np.tile(goal, (S, 1)), region_g, region_s, np.zeros(5),
                              np.array([1, 1, 0.5, 0.1, 0.1]), rnd, star=False)       # s_Parallel_rrt.m:16 uses 'RRT'
    rl = np.where(out["fail"] | (out["route_len"] < 0), 1000, out["route_len"])       # s_Parallel_rrt.m:14,21
    best = int(np.argmin(rl))                                                          # :27
    assert rl[best] < 1000
    route = out["routes"][best]                                                        # (W, 5)
    H = s["H"]
    ctx.set_cost_blocks(H, problem.Q_RRTSTAR, problem.R_MAIN_FANUC, 10.0, s["lim"], s["MAX_input"])
    sol = ctx.solve_routes(route[None], s["epsilon_O"], s["MAX_O_ITER"])
    _, _, _, sb = common.rrtstar_route_config(route.T)
    P = common.oracle_problem(O, ROBOT, obs, sb)
    orc = P.solve_batch(sb["xR"][:, 0][None], sb["ff"][None], np.array([sb["caug"]]), sb["x_"][None])
    assert int(sol["status"][0]) == int(orc["status"][0]) and int(sol["iters"][0]) == int(orc["iters"][0])
    if (int(orc["status"][0]) & 0xFF) < 2:
        assert np.abs(sol["x"][0] - orc["x"][0]).max() < 1e-6


@pytest.mark.parametrize("H", [40, 100])
def test_two_link_arm_long_horizon(ctx, oracle, H):
    """main_2L.m's planar arm at its own horizon and at H = 100 (> 64 waypoints: the prefix-sum primal recovery of the fused
    bulk tier runs in two chunks), a few perturbed goals, against the oracle."""
    robot = M.robotproperty2("2L")
    nj = 2
    goals = [[np.pi / 2, 0.0], [np.pi / 3, 0.4], [1.2, -0.3], [0.9, 0.8]]
    obs = [{"l": np.array([[0.3, 0.3], [0.3, 0.3], [0.0, 0.0]]), "D": 0.05, "epsilon": 0.05}]
    cfgs = [M.make_sys_info(robot, nj, H, [0.0, 0.0], g, Q=problem_mod().Q_2L, Rblk=problem_mod().R_2L, r_scale=0.1, lim=[0.1, 0.2],
                            max_input=np.tile(np.array([1.0, 1.0]) * 0.5 * robot["delta_t"], H), epsilon_O=1e-6,
                            MAX_O_ITER=40, x_ref=np.tile(np.array([0.0, 0.0, 0.0, 0.0]), H)) for g in goals]
    s = cfgs[0]
    r = dict(robot)
    r["name"] = "2L"
    ctx.set_robot(r, nj)
    ctx.set_obstacles(obs)
    ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    x0 = np.stack([c["xR"][:, 0] for c in cfgs])
    ff = np.stack([c["ff"] for c in cfgs])
    caug = np.array([c["caug"] for c in cfgs])
    xref = np.stack([c["x_"] for c in cfgs])
    out = ctx.solve_batch(x0, ff, caug, xref, s["epsilon_O"], s["MAX_O_ITER"])
    assert ctx.stats()["launches"] <= 12                     # the fused solver took it (lock-step: 2+ launches per iteration)
    P = common.oracle_problem(oracle, "2L", obs, s)
    ref = P.solve_batch(x0, ff, caug, xref)
    assert (out["status"] == ref["status"]).all() and (out["iters"] == ref["iters"]).all()
    ok = (ref["status"] & 0xFF) < 2
    assert ok.any() and np.abs(out["x"][ok] - ref["x"][ok]).max() < 1e-6 and np.abs(out["u"][ok] - ref["u"][ok]).max() < 1e-6


def problem_mod():
    from motionplanning_5d_m_b200 import problem
    return problem


def test_five_joint_horizon_sweep(ctx, oracle):
    """Horizons around the shared-memory limits of the fused tiers (whichever path the library picks must match the oracle)."""
    for H in (8, 33, 60, 72):
        cfg = common.batch_m16ib(oracle, 12, horizon=H)
        s = _setup(ctx, cfg)
        out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
        P = common.oracle_problem(oracle, "M16iB", cfg["obs"], s)
        ref = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=8)
        assert (out["status"] == ref["status"]).all() and (out["iters"] == ref["iters"]).all(), H
        ok = (ref["status"] & 0xFF) < 2
        if ok.any():
            assert np.abs(out["x"][ok] - ref["x"][ok]).max() < 1e-6, H


@pytest.mark.timeout(120)
def test_non_finite_inputs_terminate(ctx, oracle):
    """NaN / Inf in a problem's inputs must neither hang the persistent kernels nor disturb the other problems of the batch."""
    cfg = common.batch_m16ib(oracle, 16, horizon=20)
    s = _setup(ctx, cfg)
    ref = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
    x0, ff, caug, xref = (cfg[k].copy() for k in ("x0", "ff", "caug", "xref"))
    ff[3, 7] = np.nan
    xref[5, 12] = np.inf
    x0[9, 1] = np.nan
    for fused in (1, 0):
        ctx.set_option("fused", fused)
        out = ctx.solve_batch(x0, ff, caug, xref, s["epsilon_O"], s["MAX_O_ITER"])
        ctx.set_option("fused", 1)
        good = np.setdiff1d(np.arange(16), [3, 5, 9])
        assert np.array_equal(out["status"][good], ref["status"][good]) and np.array_equal(out["iters"][good], ref["iters"][good])
        assert np.abs(out["x"][good] - ref["x"][good]).max() < (1e-12 if fused else 1e-6)   # lock-step: parity tolerance
        assert ((out["status"][[3, 5, 9]] & 0xFF) <= 3).all()


def test_scheduling_options_never_change_results(ctx, oracle):
    """lpt (longest-expected-first work order), bulk_grid / heavy_grid caps and heavy_prio only change WHEN a problem is solved."""
    cfg = common.batch_m16ib(oracle, 160, horizon=30)
    s = _setup(ctx, cfg)
    args = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
    ref = ctx.solve_batch(*args, s["epsilon_O"], s["MAX_O_ITER"])
    for name, val, back in (("lpt", 0, 1), ("bulk_grid", 7, 0), ("heavy_grid", 3, 0), ("heavy_prio", 0, 1)):
        ctx.set_option(name, val)
        out = ctx.solve_batch(*args, s["epsilon_O"], s["MAX_O_ITER"])
        ctx.set_option(name, back)
        for k in ("u", "x", "iters", "status"):
            assert np.array_equal(out[k], ref[k]), (name, k)
    with pytest.raises(M.CfsError, match="unknown option"):
        ctx.set_option("no_such_option", 1)


def test_profiled_instantiation_returns_the_same_results(ctx, oracle):
    """Timing level 3 runs the profiled instantiations (clock64 split of the warp tier: cfs_get_warp_profile, and of the QP core:
    cfs_get_qp_profile).  Same results bit for bit, plausible counters: gradient + QP cycles inside the per-problem total, one
    gradient pass per problem-iteration (a failed QP consumed its gradient pass too)."""
    cfg = common.batch_m16ib(oracle, 96)
    s = cfg["sys_info"]
    r = dict(cfg["robot"])
    r["name"] = "M16iB"
    ctx.set_robot(r, 5)
    ctx.set_obstacles(cfg["obs"])
    ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
    args = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
    ref = ctx.solve_batch(*args)
    ctx.set_timing(3)
    try:
        out = ctx.solve_batch(*args)
        wp = ctx.warp_profile()
    finally:
        ctx.set_timing(1)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(out[k], ref[k]), k
    grad, qp, total, passes, resident, sms = (int(v) for v in wp)
    assert grad > 0 and qp > 0 and grad + qp < total and resident >= sms > 0
    st = out["status"] & 0xFF
    expect = int(out["iters"].sum() + (st >= 2).sum())   # problem-iterations incl. the failed last QP of a problem
    assert 0.5 * expect <= passes <= expect + 96          # (iterations the heavy tier ran are not the warp tier's passes)
