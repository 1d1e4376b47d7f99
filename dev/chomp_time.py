"""CHOMP_FANUC batch timing: python dev/chomp_time.py [B] [H]"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0], horizon=H)
s = cfg["sys_info"]; ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
uu = np.zeros((B, H * 5))
for rep in range(3):
    out = ctx.chomp_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], uu, float(s["alpha"]), 20)
st = ctx.stats()
print("CHOMP_FANUC: %d problems, H = %d, 20 iterations: %.2f ms on the device (%d launches) -> %.0f trajectories/s, %.3f ms per iteration; "
      "cost first/last of problem 0: %.4g / %.4g" % (B, H, st["ms_total"], st["launches"], B / st["ms_total"] * 1e3, st["ms_total"] / 20,
                                                     out["cost_hist"][0, 0], out["cost_hist"][0, -1]))
