import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib
from tests import common
cfg = common.batch_m16ib(O, 96, horizon=30)
s = dict(cfg["sys_info"]); s["alpha"] = 1.0 / np.linalg.svd(s["QQ"], compute_uv=False).max(); s["MAX_O_ITER"] = 8
ctx = M.Context(0); r=dict(cfg["robot"]); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles(cfg["obs"]); ctx.set_cost(s["H"], s["QQ"], s["lim"], None)
noise = np.random.default_rng(5).normal(0.0, 0.1, size=(96, 8, 150))
np.set_printoptions(precision=10, linewidth=220)
for K in range(1, 9):
    s["MAX_O_ITER"] = K
    P = common.oracle_problem(O, "M16iB", cfg["obs"], s, solver=1)
    nz = np.ascontiguousarray(noise[:, :K, :])
    ref = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], noise=nz, nthreads=8)
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], K, solver=_lib.SOLVER_PSGCFS, noise=nz, alpha=s["alpha"])
    ok = ((ref["status"] & 0xFF) < 2) & (ref["status"] == out["status"])
    du = np.abs(out["u"] - ref["u"]).max(1); dx = np.abs(out["x"] - ref["x"]).max(1)
    du[~ok] = 0; dx[~ok] = 0
    b = int(np.argmax(du))
    print("K", K, "status mismatch", int((ref["status"] != out["status"]).sum()), "max du %.3e (b=%d) max dx %.3e" % (du.max(), b, dx.max()),
          "qp_max_active", ref["qp_max_active"][b], "qp_iters", ref["qp_iters"][b])
    print("   cost gpu", out["cost_hist"][b][:K]); print("   cost ref", ref["cost_hist"][b][:K])
    print("   top du problems", np.argsort(-du)[:5], np.sort(du)[::-1][:5])
