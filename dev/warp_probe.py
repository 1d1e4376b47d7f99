"""bulk-tier probe: one 4096-problem batch per option set -- tier times (CUDA events inside the library), escalations, and the
largest difference to the CTA-per-problem bulk tier.   python dev/warp_probe.py [B] [seeds]"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
SEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, nj, K = 50, 5, 20
robot = dict(M.robotproperty2("M16iB")); robot["name"] = "M16iB"
ctx = M.Context(0)
ctx.set_robot(robot, nj); ctx.set_obstacles([synthetic.OBS_M16IB])
MODES = [("cta", dict(warp=0)), ("warp q15", dict(warp=1, warp_cfg=3, warp_qcap=15)), ("warp q24", dict(warp=1, warp_cfg=3, warp_qcap=24)),
         ("warp q32", dict(warp=1, warp_cfg=3, warp_qcap=32)), ("warp q48", dict(warp=1, warp_cfg=3, warp_qcap=48)),
         ("warp q48 esc96", dict(warp=1, warp_cfg=3, warp_qcap=48, esc_steps=96))]
LEVEL = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for seed in range(SEEDS):
    cfg = synthetic.batch_config_m16ib(B, lambda cand: ctx.nodes_feasible(cand)[0], horizon=H, seed=synthetic.SEED + seed)
    s = cfg["sys_info"]
    ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    args = [cfg[k] for k in ("x0", "ff", "caug", "xref")]
    base = None
    for name, opts in MODES:
        ctx.set_option("esc_steps", 48); ctx.set_option("heavy_cfg", 0); ctx.set_option("heavy_grid", 48); ctx.set_option("screen", 1)
        for k, v in opts.items(): ctx.set_option(k, v)
        ctx.set_timing(LEVEL)
        best = None
        for rep in range(3):
            out = ctx.solve_batch(*args, s["epsilon_O"], K)
            st = ctx.stats()
            if best is None or st["ms_total"] < best["ms_total"]: best = st
        ps = ctx.problem_steps(B)
        line = "seed %d %-14s total %.3f bulk %.3f heavy %.3f  qp_steps %d max_active %d maxsteps %d" % (
            seed, name, best["ms_total"], best["ms_bulk"], best["ms_heavy"], best["qp_steps"], best["max_active"], ps.max())
        if base is None:
            base = out
        else:
            ok = ((base["status"] & 0xFF) < 2)
            same_st = int((out["status"] != base["status"]).sum()); same_it = int((out["iters"] != base["iters"]).sum())
            dx = np.abs(out["x"][ok] - base["x"][ok]).max(); du = np.abs(out["u"][ok] - base["u"][ok]).max()
            it = base["iters"]; sel = ok & (it > 0)
            dc = np.abs(out["cost_hist"][sel, it[sel] - 1] / base["cost_hist"][sel, it[sel] - 1] - 1).max()
            line += "  | vs cta: status diff %d iters diff %d max|dx| %.2e max|du| %.2e rel dcost %.2e" % (same_st, same_it, dx, du, dc)
        print(line, flush=True)
