"""per-outer-iteration kernel times + problem statistics for the headline batch"""
import sys; sys.path.insert(0,'.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv)>1 else 4096
fused = int(sys.argv[2]) if len(sys.argv)>2 else 1
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0])
s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"])
ctx.set_option("fused", fused)
ctx.set_timing(2)
for rep in range(3):
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
g, q = ctx.iter_times()
np.set_printoptions(precision=3, linewidth=200, suppress=True)
print("stats", ctx.stats())
print("grad ms per iter", g)
print("qp ms per iter  ", q)
st = out["status"] & 0xff
print("status counts", np.bincount(st), "iters hist", np.bincount(out["iters"]))
print("infeasible by iter", np.bincount(out["iters"][st==2]))

ps = ctx.problem_steps(B)
inf1 = (st==2)&(out["iters"]==0)
print("steps of iter-1-infeasible problems: mean %.1f median %d p90 %d p99 %d max %d" % (ps[inf1].mean(), np.median(ps[inf1]), np.percentile(ps[inf1],90), np.percentile(ps[inf1],99), ps[inf1].max()))
print("steps hist (iter-1 infeasible)", np.histogram(ps[inf1], bins=[0,5,10,20,40,80,160,320,640,100000])[0])
print("steps feasible problems: mean %.2f max %d" % (ps[st<2].mean(), ps[st<2].max()))

ctx.set_timing(3)
out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
pf = ctx.qp_profile()
names = ["setup+grad","refresh","scan","gram+solve","update","epilogue"]
for tier, o in (("bulk", 0), ("heavy", 8)):
    p8 = pf[o:o+8]
    print("qp profile %s: problems %d outer steps %d total ticks %d" % (tier, p8[6], p8[7], p8[:6].sum()))
    for k,nm in enumerate(names): print("  %-10s %12d ticks  %8.1f per problem  %8.1f per outer step" % (nm, p8[k], p8[k]/max(p8[6],1), p8[k]/max(p8[7],1)))
