"""stand-alone K1d (DERIVEST) timing: python dev/k1d_time.py [waypoints]"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 51200
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
cfg = synthetic.batch_config_m16ib(max(N // 50, 1), lambda c: ctx.nodes_feasible(c)[0])
th = cfg["xref"].reshape(-1, 50, 10)[:, :, :5].reshape(-1, 5)[:N]
peak, mhz = ctx.measure_fp64_peak()
d, lid, g, fl = ctx.dist_grad(th, grad=_lib.GRAD_DERIVEST)
print("linkid histogram", np.bincount(lid.reshape(-1), minlength=6))
for rep in range(2):
    ms = ctx.time_dist_grad(th, grad=_lib.GRAD_DERIVEST, reps=10)
    tf = 162080.0 * th.shape[0] / (ms * 1e-3) / 1e12
    print("K1d %d waypoints: %.4f ms  %.2f TFLOP/s algorithmic = %.1f%% of measured DFMA peak %.2f" % (th.shape[0], ms, tf, 100 * tf / peak, peak))
