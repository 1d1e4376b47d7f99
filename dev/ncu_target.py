"""Short single-GPU target for ncu: the headline batch (M16iB, H=50) solved twice (first pass warms up)."""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
grad = _lib.GRAD_DERIVEST if (len(sys.argv) > 2 and sys.argv[2] == "derivest") else _lib.GRAD_NUMJAC
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0])
s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"])
for rep in range(reps):
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20, grad=grad)
print("stats", ctx.stats())
print("status", np.bincount(out["status"] & 0xff))
