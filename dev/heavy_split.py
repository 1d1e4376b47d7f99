"""how much of the heavy tier is (a) first-iteration chains vs (b) late escalations: counts + serialized tier times per option set"""
import os, sys
sys.path.insert(0, '.')
import numpy as np
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
B, H, nj, K = 4096, 50, 5, 20
robot = dict(M.robotproperty2("M16iB")); robot["name"] = "M16iB"
ctx = M.Context(0); ctx.set_robot(robot, nj); ctx.set_obstacles([synthetic.OBS_M16IB])
for seed in range(3):
    cfg = synthetic.batch_config_m16ib(B, lambda c: ctx.nodes_feasible(c)[0], horizon=H, seed=synthetic.SEED + seed)
    s = cfg["sys_info"]; ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    args = [cfg[k] for k in ("x0", "ff", "caug", "xref")]
    for opts in (dict(warp_qcap=15), dict(warp_qcap=20), dict(warp_qcap=24), dict(warp_qcap=31)):
        for k, v in opts.items(): ctx.set_option(k, v)
        ctx.set_timing(2)
        out = ctx.solve_batch(*args, 0.1, K); out = ctx.solve_batch(*args, 0.1, K)
        st = ctx.stats()
        print("seed", seed, opts, "bulk %.2f heavy %.2f ms" % (st["ms_bulk"], st["ms_heavy"]), flush=True)
