"""which problems of the seeded batches reach the heavy tier: steps, status, iters"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B, H = 4096, 50
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
for c in range(8):
    cfg = synthetic.batch_config_m16ib(B, lambda cand: ctx.nodes_feasible(cand)[0], horizon=H, seed=synthetic.SEED + c)
    s = cfg["sys_info"]
    if c == 0: ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
    ps = ctx.problem_steps(B)
    st = out["status"] & 0xff
    idx = np.argsort(-ps)[:12]
    print("batch", c, "status counts", np.bincount(st, minlength=4), "problems > 48 steps:", int((ps > 48).sum()), "sum steps", ps.sum())
    print("   top:", [(int(i), int(ps[i]), int(st[i]), int(out["iters"][i])) for i in idx])
    if c == 4:
        np.savez("gpurun_out/heavy_batch4.npz", ps=ps, st=st, iters=out["iters"], idx=idx)
