import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, motionplanning_5d_m_b200 as M
from tests import common
cfg = common.batch_m16ib(O, 192); s = cfg["sys_info"]
ctx = M.Context(0); r=dict(cfg["robot"]); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles(cfg["obs"]); ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
P = common.oracle_problem(O, "M16iB", cfg["obs"], s)
ref = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=8)
out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], s["MAX_O_ITER"])
np.set_printoptions(precision=6, linewidth=200)
bad = np.where(out["status"]!=ref["status"])[0]
print("stats", ctx.stats())
for b in bad:
    print(b, "gpu", out["status"][b], out["iters"][b], "ref", ref["status"][b], ref["iters"][b], ref["qp_iters"][b], ref["qp_max_active"][b])
    print(" gpu cost", out["cost_hist"][b]); print(" ref cost", ref["cost_hist"][b])
    print(" gpu e_u", out["e_u_hist"][b])
ok = (ref["status"]&0xff)<2
ok &= out["status"]==ref["status"]
print("max dx", np.abs(out["x"][ok]-ref["x"][ok]).max(), "max du", np.abs(out["u"][ok]-ref["u"][ok]).max())
print("iters equal", (out["iters"][ok]==ref["iters"][ok]).all())
