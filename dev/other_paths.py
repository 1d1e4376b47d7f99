import sys, time; sys.path.insert(0,'.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic, _lib
B=int(sys.argv[1]) if len(sys.argv)>1 else 4096
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0])
s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"])
alpha = 1.0/np.linalg.svd(s["QQ"], compute_uv=False).max()
noise = np.random.default_rng(1).normal(0,0.1,size=(B,20,250))
for name, kw in (("CFS numjac fused", {}), ("CFS numjac lockstep", {"fused0":1}), ("CFS derivest", dict(grad=_lib.GRAD_DERIVEST)), ("PSGCFS", dict(solver=_lib.SOLVER_PSGCFS, noise=noise, alpha=alpha))):
    if "fused0" in kw: ctx.set_option("fused",0); kw={}
    for rep in range(2):
        out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20, **kw)
    st=ctx.stats(); ctx.set_option("fused",1)
    print("%-22s device ms %.2f  launches %d  problem_iters %d  status %s" % (name, st["ms_total"], st["launches"], st["problem_iters"], np.bincount(out["status"]&0xff)))
