"""the PSGCFS problems whose first projection QP the oracle (and an LP) call infeasible: what does the library return?"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M, oracle as O
from motionplanning_5d_m_b200 import synthetic, _lib
import bench
O.build()
B, H = 2048, 30
ctx = M.Context(0)
robot = dict(M.robotproperty2("M200i")); robot["name"] = "M200i"
ctx.set_robot(robot, 5); ctx.set_obstacles([synthetic.OBS_M200I])
c0 = synthetic.batch_config_m200i_psgcfs(B, lambda c: ctx.nodes_feasible(c)[0], horizon=H, seed=synthetic.SEED)
s = c0["sys_info"]; n = H * 5
ctx.set_cost(H, s["QQ"], s["lim"], None)
P = bench.make_oracle_problem(O, c0, 0, solver=1)
sel = np.array([646, 652, 1325, 1376, 1521, 1651, 1706, 2021, 0, 1])
for lock in (0, 1):
    ctx.set_option("warp_lockstep", lock)
    out = ctx.solve_batch(c0["x0"][sel], c0["ff"][sel], c0["caug"][sel], c0["xref"][sel], float(s["epsilon_O"]), 1, solver=_lib.SOLVER_PSGCFS, noise=c0["noise"][sel][:, :1], alpha=float(s["alpha"]))
    for k, b in enumerate(sel):
        A_, b_, dist, lid, grad, t_ = P.get_con(c0["x0"][b], c0["xref"][b], np.zeros(n))
        viol = (A_ @ out["u"][k] - b_)
        print("lockstep_warp=%d problem %d: status %d iters %d, max row violation of the returned u: %.3e (row %d of %d), |u| %.3g" % (lock, b, out["status"][k] & 255, out["iters"][k], viol.max(), int(viol.argmax()), len(b_), np.abs(out["u"][k]).max()))
st = ctx.stats(); print(st)
print("---- rows: library (cfs_get_con, margin D) vs oracle ----")
ctx.set_option("warp_lockstep", 0)
for b in (652, 1325, 0):
    A_, b_, dist, lid, grad, t_ = P.get_con(c0["x0"][b], c0["xref"][b], np.zeros(n))
    Ag, bg = ctx.get_con(c0["x0"][b], c0["xref"][b], np.zeros(n), margin_is_D=True)
    out = ctx.solve_batch(c0["x0"][[b]], c0["ff"][[b]], c0["caug"][[b]], c0["xref"][[b]], float(s["epsilon_O"]), 1, solver=_lib.SOLVER_PSGCFS, noise=c0["noise"][[b]][:, :1], alpha=float(s["alpha"]))
    u = out["u"][0]
    print("problem", b, "rows equal:", np.abs(Ag - A_).max(), np.abs(bg - b_).max(), "| violation of returned u on library rows %.3e" % (Ag @ u - bg).max(),
          "steps", int(ctx.problem_steps(1)[0]), "max_active", ctx.stats()["max_active"], "touch", t_)
    # is u the unprojected PSG point?
    up = -float(s["alpha"]) * (c0["ff"][b] + 10 * c0["noise"][b][0] / 2.0)
    print("   |u - psg_point| %.3e   |psg point| %.3g" % (np.abs(u - up).max(), np.abs(up).max()))
