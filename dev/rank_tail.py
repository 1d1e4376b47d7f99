"""per-rank drain study: the 20 batches rank r of an 8-GPU bench run solves (seeds SEED+1000r+c), on one GPU.
python dev/rank_tail.py [ranks] [opt=value ...]"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np, torch
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
opts = dict(kv.split("=") for kv in sys.argv[2:])
NC, steps = 20, 20
B, H, nj = 4096, 50, 5
n, N, K = H * nj, 2 * H * nj, 20
dev = torch.device("cuda", 0)
robot = dict(M.robotproperty2("M16iB")); robot["name"] = "M16iB"
ctxs, streams = [], []
for c in range(NC):
    ctx = M.Context(0); st = torch.cuda.Stream(device=dev); ctx.set_stream(st.cuda_stream)
    ctx.set_robot(robot, nj); ctx.set_obstacles([synthetic.OBS_M16IB])
    for k, v in opts.items(): ctx.set_option(k, int(v))
    ctxs.append(ctx); streams.append(st)
mk = lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt, device=dev)
d_out = [dict(u=mk(B, n), x=mk(B, N), cost=mk(B, K), eu=mk(B, K), iters=mk(B, dt=torch.int32), status=mk(B, dt=torch.int32)) for _ in range(NC)]
main_stream = torch.cuda.Stream(device=dev)
names = ("x0", "ff", "caug", "xref")
first = True
for r in range(R):
    cfgs = [synthetic.batch_config_m16ib(B, lambda cand: ctxs[0].nodes_feasible(cand)[0], horizon=H, seed=synthetic.SEED + 1000 * r + c) for c in range(NC)]
    if first:
        s = cfgs[0]["sys_info"]
        for ctx in ctxs: ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
        first = False
    d_in = [{k: torch.from_numpy(cfgs[c][k]).to(dev) for k in names} for c in range(NC)]
    def issue(c):
        i, o = d_in[c], d_out[c]
        ctxs[c].solve_batch_ptr(B, i["x0"].data_ptr(), i["ff"].data_ptr(), i["caug"].data_ptr(), i["xref"].data_ptr(), 0.1, K,
                                o["u"].data_ptr(), o["x"].data_ptr(), o["cost"].data_ptr(), o["eu"].data_ptr(), o["iters"].data_ptr(),
                                o["status"].data_ptr(), device=True, sync=False)
    def run():
        tb, te = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tb.record(main_stream)
        for st in streams: st.wait_event(tb)
        for k in range(steps): issue(k % NC)
        for st in streams:
            e = torch.cuda.Event(); e.record(st); main_stream.wait_event(e)
        te.record(main_stream); torch.cuda.synchronize()
        for c in ctxs: c.wait()
        return tb.elapsed_time(te)
    run()
    ms = min(run() for _ in range(3))
    lat, mx = [], []
    for c in range(NC):
        ctxs[c].set_timing(2); torch.cuda.synchronize(); issue(c); ctxs[c].wait(); stt = ctxs[c].stats()
        lat.append(stt["ms_total"]); mx.append(int(ctxs[c].problem_steps(B).max())); ctxs[c].set_timing(0)
    print("rank %d: %.3f ms/step (%.1f ms total); single-batch latency mean %.2f max %.2f ms; longest chain per batch: max %d, top5 %s"
          % (r, ms / steps, ms, np.mean(lat), np.max(lat), max(mx), sorted(mx)[-5:]), flush=True)
