"""worst-case single-batch latency over many seeds (which seeds carry long heavy-tier chains?)"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B, H = 4096, 50
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
ctx.set_timing(2)
res = []
for base in (0, 1000, 2000, 7000):
    for c in range(24 if base in (0, 7000) else 8):
        cfg = synthetic.batch_config_m16ib(B, lambda cand: ctx.nodes_feasible(cand)[0], horizon=H, seed=synthetic.SEED + base + c)
        s = cfg["sys_info"]
        if not res: ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
        for rep in range(2):
            out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
        st = ctx.stats(); ps = ctx.problem_steps(B)
        res.append((base + c, st["ms_total"], st["ms_bulk"], st["ms_heavy"], int(ps.max())))
res = np.array(res)
print("seeds %d: total ms  mean %.2f  median %.2f  p90 %.2f  max %.2f (seed %d)" % (len(res), res[:,1].mean(), np.median(res[:,1]), np.percentile(res[:,1], 90), res[:,1].max(), res[res[:,1].argmax(),0]))
print("heavy ms: mean %.2f max %.2f; bulk ms mean %.2f max %.2f; max steps in one problem %d" % (res[:,3].mean(), res[:,3].max(), res[:,2].mean(), res[:,2].max(), res[:,4].max()))
print(np.round(res[np.argsort(-res[:,1])[:8]], 2))
