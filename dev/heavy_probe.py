"""heavy-tier timing of single batches: python dev/heavy_probe.py [seed_offset,...] [opt=value ...]"""
import sys; sys.path.insert(0,'.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
offs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
opts = dict(kv.split("=") for kv in sys.argv[2:])
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"]="M16iB"; ctx.set_robot(r,5); ctx.set_obstacles([synthetic.OBS_M16IB])
for k, v in opts.items(): ctx.set_option(k, int(v))
first = True
for off in offs:
    cfg = synthetic.batch_config_m16ib(4096, lambda c: ctx.nodes_feasible(c)[0], seed=synthetic.SEED + off)
    if first:
        s = cfg["sys_info"]; ctx.set_cost(50, s["QQ"], s["lim"], s["MAX_input"]); first = False
    ctx.set_timing(2)
    for rep in range(2):
        out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20)
    st = ctx.stats(); ps = ctx.problem_steps(4096)
    ctx.set_timing(1); ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], 0.1, 20); lat = ctx.stats()["ms_total"]
    print(opts, "seed+%d: serialised total %.2f ms (bulk %.2f heavy %.2f), pipelined %.2f ms | qp_steps %d longest chain %d max_active %d | status sum %d" % (off, st["ms_total"], st["ms_bulk"], st["ms_heavy"], lat, st["qp_steps"], ps.max(), st["max_active"], int((out["status"] & 0xff).sum())), flush=True)
