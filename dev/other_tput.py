"""throughput of the launch-per-iteration paths (PSGCFS, DERIVEST) on a 4096-problem M16iB batch"""
import sys; sys.path.insert(0, '.')
import numpy as np, time
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic, _lib
B, H, K = 4096, 50, 20
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0], horizon=H)
s = cfg["sys_info"]
alpha = 1.0 / np.linalg.svd(s["QQ"], compute_uv=False).max()
noise = np.random.default_rng(5).normal(0.0, 0.1, size=(B, K, H * 5))
args = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
for name, kw, bounds in (("CFS fused", dict(), True), ("CFS lock-step", dict(), True), ("CFS DERIVEST", dict(grad=_lib.GRAD_DERIVEST), True),
                         ("PSGCFS", dict(solver=_lib.SOLVER_PSGCFS, noise=noise, alpha=alpha), False)):
    ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"] if bounds else None)
    ctx.set_option("fused", 0 if name == "CFS lock-step" else 1)
    for rep in range(2):
        out = ctx.solve_batch(*args, 0.1, K, **kw)
    st = ctx.stats()
    print("%-14s total %.2f ms  launches %d  problem_iters %d  -> %.3f M traj/s  status %s" % (name, st["ms_total"], st["launches"], st["problem_iters"], B / st["ms_total"] / 1e3, np.bincount(out["status"] & 0xff, minlength=4)))
