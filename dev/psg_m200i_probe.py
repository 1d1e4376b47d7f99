"""PSGCFS M200i batch: per-iteration-count difference GPU vs oracle for the worst problems, and the oracle's own sensitivity."""
import sys; sys.path.insert(0, '.')
import numpy as np, oracle as O
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib, synthetic
from tests import common
O.build()
ctx = M.Context(0)
r = dict(M.robotproperty2("M200i")); r["name"] = "M200i"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M200I])
B = 128
cfg = synthetic.batch_config_m200i_psgcfs(B, lambda c: ctx.nodes_feasible(c)[0])
s = dict(cfg["sys_info"]); ctx.set_cost(s["H"], s["QQ"], s["lim"], None)
args = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
def both(k, nz):
    s["MAX_O_ITER"] = k
    P = common.oracle_problem(O, "M200i", cfg["obs"], s, solver=1)
    ref = P.solve_batch(*args, noise=nz[:, :k], nthreads=8)
    out = ctx.solve_batch(*args, s["epsilon_O"], k, solver=_lib.SOLVER_PSGCFS, noise=np.ascontiguousarray(nz[:, :k]), alpha=s["alpha"])
    return P, ref, out
for k in (1, 2, 3, 4, 6, 8, 12, 16, 20):
    P, ref, out = both(k, cfg["noise"])
    ok = (ref["status"] & 0xFF) < 2
    dx = np.abs(out["x"] - ref["x"]).max(axis=1); dx[~ok] = 0
    p1 = P.solve_batch(*args, noise=cfg["noise"][:, :k] * (1 + 1e-9), nthreads=8)
    p2 = P.solve_batch(cfg["x0"], cfg["ff"] * (1 + 1e-12), cfg["caug"], cfg["xref"], noise=cfg["noise"][:, :k], nthreads=8)
    s1 = np.abs(p1["x"] - ref["x"]).max(axis=1); s2 = np.abs(p2["x"] - ref["x"]).max(axis=1)
    w = np.argsort(-dx)[:4]
    print("k=%2d worst:" % k, [(int(b), "%.1e" % dx[b], "noise-sens %.1e" % s1[b], "ff-sens %.1e" % s2[b]) for b in w], flush=True)
