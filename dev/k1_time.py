"""stand-alone K1 timing: python dev/k1_time.py [waypoints]"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 204800
ctx = M.Context(0)
r = dict(M.robotproperty2("M16iB")); r["name"] = "M16iB"; ctx.set_robot(r, 5); ctx.set_obstacles([synthetic.OBS_M16IB])
rng = np.random.default_rng(1)
th = synthetic.SAMPLE_OFF + (rng.random((N, 5)) - 0.5) * 2 * synthetic.REGION_S
peak, mhz = ctx.measure_fp64_peak()
for rep in range(3):
    ms = ctx.time_dist_grad(th, reps=20)
    tf = 9680.0 * N / (ms * 1e-3) / 1e12
    print("K1 %d waypoints: %.4f ms  %.2f TFLOP/s algorithmic = %.1f%% of measured DFMA peak %.2f (%.0f MHz)" % (N, ms, tf, 100 * tf / peak, peak, mhz))
