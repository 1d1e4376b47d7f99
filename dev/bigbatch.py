"""one large batch on one context vs. many contexts: is the multi-context pipeline losing throughput?"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np, torch
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
mult = int(sys.argv[1]) if len(sys.argv) > 1 else 16
opts = dict(kv.split('=') for kv in sys.argv[2:])
B0, H, nj = 4096, 50, 5
n, N, K = H * nj, 2 * H * nj, 20
dev = torch.device("cuda", 0)
robot = dict(M.robotproperty2("M16iB")); robot["name"] = "M16iB"
ctx = M.Context(0); ctx.set_robot(robot, nj); ctx.set_obstacles([synthetic.OBS_M16IB])
cfgs = [synthetic.batch_config_m16ib(B0, lambda cand: ctx.nodes_feasible(cand)[0], horizon=H, seed=synthetic.SEED + c) for c in range(min(mult, 8))]
s = cfgs[0]["sys_info"]; ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
B = B0 * mult
cat = lambda k: np.concatenate([cfgs[c % len(cfgs)][k] for c in range(mult)], axis=0)
d_in = {k: torch.from_numpy(cat(k)).to(dev) for k in ("x0", "ff", "caug", "xref")}
mk = lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt, device=dev)
o = dict(u=mk(B, n), x=mk(B, N), cost=mk(B, K), eu=mk(B, K), iters=mk(B, dt=torch.int32), status=mk(B, dt=torch.int32))
ctx.set_timing(2)
for k_, v_ in opts.items(): ctx.set_option(k_, int(v_))
for rep in range(3):
    ctx.solve_batch_ptr(B, d_in["x0"].data_ptr(), d_in["ff"].data_ptr(), d_in["caug"].data_ptr(), d_in["xref"].data_ptr(), 0.1, K,
                        o["u"].data_ptr(), o["x"].data_ptr(), o["cost"].data_ptr(), o["eu"].data_ptr(), o["iters"].data_ptr(),
                        o["status"].data_ptr(), device=True, sync=False)
    ctx.wait(); st = ctx.stats()
    if rep == 2: print(opts, "B=%d: total %.3f ms (bulk %.3f heavy %.3f) -> %.3f ms per 4096, %.3f M traj/s" % (B, st["ms_total"], st["ms_bulk"], st["ms_heavy"], st["ms_total"] / mult, B / st["ms_total"] / 1e3))
