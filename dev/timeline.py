"""Where do the stages of 20 pipelined batches fall on the device's time axis?  python dev/timeline.py [contexts] [opt=value ...]"""
import os, sys, ctypes as C
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np, torch
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic, _lib
NC = int(sys.argv[1]) if len(sys.argv) > 1 else 20
opts = dict(kv.split("=") for kv in sys.argv[2:])
steps = 20
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
cfgs = [synthetic.batch_config_m16ib(B, lambda cand: ctxs[0].nodes_feasible(cand)[0], horizon=H, seed=synthetic.SEED + c) for c in range(NC)]
s = cfgs[0]["sys_info"]
for ctx in ctxs: ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
d_in = [{k: torch.from_numpy(cfgs[c][k]).to(dev) for k in ("x0", "ff", "caug", "xref")} for c in range(NC)]
mk = lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt, device=dev)
d_out = [dict(u=mk(B, n), x=mk(B, N), cost=mk(B, K), eu=mk(B, K), iters=mk(B, dt=torch.int32), status=mk(B, dt=torch.int32)) for _ in range(NC)]
main_stream = torch.cuda.Stream(device=dev)
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
run(); run()
ms = run()
lib = _lib.load()
print("contexts %d opts %s: %.2f ms for %d batches (%.3f ms/batch)" % (NC, opts, ms, steps, ms / steps))
print("batch:  start  screen-done  heavy1-done  all-done   (ms after the first batch's start)")
out = (C.c_double * 4)()
rows = []
for c in range(min(NC, steps)):
    lib.cfs_dev_timeline(ctxs[c]._h, ctxs[0]._h, out)
    rows.append(list(out))
    print("  %2d: %7.2f %9.2f %11.2f %10.2f" % (c, *out))
rows = np.array(rows)
print("last heavy1-done %.2f, last all-done %.2f, median (all-done - heavy1-done) %.2f" % (rows[:, 2].max(), rows[:, 3].max(), np.median(rows[:, 3] - rows[:, 2])))
