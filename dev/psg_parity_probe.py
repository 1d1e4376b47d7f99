"""PSGCFS bench batch: where does the library differ from the oracle, and how do the oracle's own twin builds behave there?"""
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
s = c0["sys_info"]; K = int(s["MAX_O_ITER"])
ctx.set_cost(H, s["QQ"], s["lim"], None)
out = ctx.solve_batch(c0["x0"], c0["ff"], c0["caug"], c0["xref"], float(s["epsilon_O"]), K, solver=_lib.SOLVER_PSGCFS, noise=c0["noise"], alpha=float(s["alpha"]))
class A: grad = "numjac"
P = bench.make_oracle_problem(O, c0, 0, solver=1)
ref = P.solve_batch(c0["x0"], c0["ff"], c0["caug"], c0["xref"], noise=c0["noise"])
twin = P.solve_batch(c0["x0"], c0["ff"], c0["caug"], c0["xref"], noise=c0["noise"], use_twin=True)
sens = np.abs(twin["x"] - ref["x"]).max(axis=1)
dx = np.abs(out["x"] - ref["x"]).max(axis=1)
same = ((ref["status"] & 0xFF) == (out["status"] & 0xFF)) & (ref["iters"] == out["iters"])
cond = (sens < 1e-8) & (twin["status"] == ref["status"]) & (twin["iters"] == ref["iters"])
print("status hist gpu", np.bincount(out["status"] & 0xFF), "ref", np.bincount(ref["status"] & 0xFF), "twin", np.bincount(twin["status"] & 0xFF))
print("not same:", int((~same).sum()), " cond & ~same:", int((cond & ~same).sum()), " ~cond:", int((~cond).sum()))
for b in np.where(~same)[0][:20]:
    print("  problem %d: gpu status %d iters %d | ref %d %d | twin %d %d | twin dx %.2e | gpu dx %.2e" % (b, out["status"][b] & 255, out["iters"][b], ref["status"][b] & 255, ref["iters"][b], twin["status"][b] & 255, twin["iters"][b], sens[b], dx[b]))
# per-iteration growth on the worst "well conditioned" problems
wc = np.where(cond & same)[0]
worst = wc[np.argsort(-dx[wc])[:5]]
print("largest gpu dx among well-conditioned & same:", [(int(b), float(dx[b]), float(sens[b])) for b in worst])
print("cost hist diff of those:", [float(np.nanmax(np.abs(out["cost_hist"][b] - ref["cost_hist"][b]) / np.abs(ref["cost_hist"][b]))) for b in worst])
