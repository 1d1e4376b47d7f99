import sys; sys.path.insert(0,'.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B=4096
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0])
s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"])
outs=[]; steps=[]
for rep in range(5):
    if rep==3: ctx.set_timing(2)
    o = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
    outs.append(o); steps.append(ctx.problem_steps(B).copy()); print(rep, ctx.stats()["qp_steps"], ctx.stats()["max_active"], ctx.stats()["ms_total"])
for rep in range(1,5):
    du=np.abs(outs[rep]["u"]-outs[0]["u"]).max(1); ds=(steps[rep]!=steps[0])
    print(rep, "bitwise equal u:", np.array_equal(outs[rep]["u"],outs[0]["u"]), "max du", du.max(), "problems with different steps", np.where(ds)[0][:10], steps[0][ds][:10], steps[rep][ds][:10], "status diff", (outs[rep]["status"]!=outs[0]["status"]).sum())
