"""CHOMP_FANUC (Lib/CHOMP_FANUC.m): the C oracle against an independent numpy restatement that uses the
literal Baug row slices of the reference (CPU), and the CUDA path (cfs_chomp_batch) against the oracle (GPU)."""
import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
import oracle as O
from tests import common, np_restatement as NP


def _small_config(ROBOT, H, K, through_obstacle=True):
    robot = M.robotproperty2(ROBOT)
    if ROBOT == "M200i":
        x0 = [0.7825, 0.0284, 0.2172, 0.1444, -1.1779]
        xg = [-0.7825, 0.0284, 0.2172, 0.1444, -1.1779]
        obs = [{"l": np.array([[3.806, 3.606], [8.413, 8.413], [0.001, 1.038]]), "D": 0.2, "epsilon": 0.25}]
    else:
        x0 = [1.2, 0.1, 0.2, 0.1, -0.5]
        xg = [-0.8, 0.3, 0.1, 0.2, 0.4]
        obs = [dict(M.synthetic.OBS_M16IB)]
    s = M.make_sys_info(robot, 5, H, x0, xg, MAX_O_ITER=K)
    return robot, obs, s


@pytest.mark.parametrize("ROBOT", ["M16iB", "M200i"])
def test_oracle_chomp_against_an_independent_restatement(ROBOT):
    H, K = 10, 4
    robot, obs, s = _small_config(ROBOT, H, K)
    n = H * 5
    uu = 0.05 * np.sin(np.arange(n))
    u_np, x_np, cost_np, eu_np = NP.chomp_optimizer(s, robot, obs, ROBOT, uu)
    P = common.oracle_problem(O, ROBOT, obs, s)
    ref = P.chomp_batch([o["D"] for o in obs], [o["epsilon"] for o in obs], s["xR"][:, 0][None], s["ff"][None],
                        np.array([s["caug"]]), np.asarray(s["x_"])[None], uu[None], nthreads=1)
    assert int(ref["iters"][0]) == K and (int(ref["status"][0]) & 0xFF) == 1
    np.testing.assert_allclose(ref["u"][0], u_np, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ref["x"][0], x_np, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ref["cost_hist"][0], cost_np, rtol=1e-10)
    np.testing.assert_allclose(ref["e_u_hist"][0], eu_np, rtol=1e-9, atol=1e-12)
    # the obstacle term is live in this configuration (otherwise the test would only cover the quadratic part)
    quad_only = uu - s["alpha"] * 3 * (s["QQ"] @ uu + s["ff"])
    assert np.abs(quad_only - NP.chomp_optimizer(dict(s, MAX_O_ITER=1), robot, obs, ROBOT, uu)[0]).max() > 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("ROBOT", ["M16iB", "M200i"])
def test_gpu_chomp_batch_against_the_oracle(ROBOT):
    H, K, B = 30, 20, 24
    robot, obs, s = _small_config(ROBOT, H, K)
    rng = np.random.default_rng(7)
    n = H * 5
    if ROBOT == "M16iB":
        cfg = M.synthetic.batch_config_m16ib(B, common.oracle_feasible_fn(O, ROBOT, obs), H, seed=11)
        s = cfg["sys_info"]
        x0, ff, caug, xref = cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"]
    else:
        cfg = M.synthetic.batch_config_m200i_psgcfs(B, common.oracle_feasible_fn(O, ROBOT, obs), horizon=H, seed=11)
        s = cfg["sys_info"]
        x0, ff, caug, xref = cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"]
    s = dict(s, MAX_O_ITER=K)
    uu = 0.02 * rng.standard_normal((B, n))
    P = common.oracle_problem(O, ROBOT, obs, s)
    ref = P.chomp_batch([o["D"] for o in obs], [o["epsilon"] for o in obs], x0, ff, caug, xref, uu)
    ctx = M.Context(0)
    r = dict(robot)
    r["name"] = ROBOT
    ctx.set_robot(r, 5)
    ctx.set_obstacles(obs)
    ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    out = ctx.chomp_batch(x0, ff, caug, xref, uu, float(s["alpha"]), K)
    assert (out["iters"] == K).all() and ((out["status"] & 0xFF) == 1).all()
    # touch flag: the kernel runs derivest for every (waypoint, obstacle) pair, the reference only where the potential is live,
    # so the device flag can only be a superset of the oracle's
    assert (((out["status"] & 0x100) | (ref["status"] & 0x100)) == (out["status"] & 0x100)).all()
    np.testing.assert_allclose(out["u"], ref["u"], rtol=1e-9, atol=1e-8)  # relative: CHOMP with the 2000 x gain can leave the unit box
    np.testing.assert_allclose(out["x"], ref["x"], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(out["cost_hist"], ref["cost_hist"], rtol=1e-9)
    np.testing.assert_allclose(out["e_u_hist"], ref["e_u_hist"], rtol=1e-7, atol=1e-12)
    assert ctx.stats()["launches"] == 2 + 4 * K
    # the class mirror (B = 1) goes through the same entry
    s1 = dict(s, xR=np.tile(x0[0][:, None], (1, H + 1)), ff=ff[0], paug=ff[0], caug=float(caug[0]), x_=xref[0])
    sol = M.CHOMP_FANUC(obs, s1, uu[0], ROBOT, ctx=ctx).optimizer()
    np.testing.assert_array_equal(sol.u, out["u"][0])
    assert sol.iter_O == K + 1 and sol.eval.cost_all.shape == (K,)

