import sys; sys.path.insert(0,'.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B=4096
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0])
s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"])
out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
ps = ctx.problem_steps(B)
idx=np.argsort(-ps)[:24]
np.savez("gpurun_out/steps_dump.npz", idx=idx, steps=ps[idx], status=out["status"][idx], iters=out["iters"][idx], theta0=cfg["theta0"][idx], thetag=cfg["thetag"][idx], x=out["x"][idx], u=out["u"][idx])
print(list(zip(idx.tolist(), ps[idx].tolist(), (out["status"][idx]&0xff).tolist(), out["iters"][idx].tolist())))
