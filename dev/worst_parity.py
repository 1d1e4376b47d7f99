import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
O.build()
B=4096
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0])
s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"])
out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
P = O.Problem(O.robot("M16iB"), 50, [synthetic.OBS_M16IB["l"]], [0.2], s["QQ"], s["lim"], s["MAX_input"], 0.1, 20)
ref = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=16)
ok=((ref["status"]&0xff)<2)&(ref["status"]==out["status"])
dx=np.abs(out["x"]-ref["x"]).max(1); dx[~ok]=0
du=np.abs(out["u"]-ref["u"]).max(1); du[~ok]=0
top=np.argsort(-dx)[:8]
print("top dx", [(int(b), float(dx[b]), float(du[b]), int(ref["iters"][b]), int(ref["status"][b])) for b in top])
print("count dx>1e-7:", int((dx>1e-7).sum()), "dx>1e-6:", int((dx>1e-6).sum()))
pert = P.solve_batch(cfg["x0"], cfg["ff"]*(1+1e-12), cfg["caug"], cfg["xref"], nthreads=16)
sens=np.abs(pert["x"]-ref["x"]).max(1)
print("oracle sensitivity (ff*(1+1e-12)) on those:", [float(sens[b]) for b in top], "median", float(np.median(sens)))
b=int(top[0]); it=int(ref["iters"][b])
print("cost gpu", out["cost_hist"][b][:it]); print("cost ref", ref["cost_hist"][b][:it])
