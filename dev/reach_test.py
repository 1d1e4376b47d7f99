"""CPU study: how many first-iteration infeasible QPs of the seeded batch are caught by a closed-form single-row
reachability test (bang-bang double integrator per joint)?"""
import sys, time; sys.path.insert(0, '.')
import numpy as np, oracle as O
from motionplanning_5d_m_b200 import synthetic
seed_off = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B, H, nj = 4096, 50, 5
r = O.robot("M16iB"); o6 = O.obs6(synthetic.OBS_M16IB["l"])
feas = lambda cand: np.array([O.dist_arm(r, th, o6)[0] >= 0.2 for th in cand])
cfg = synthetic.batch_config_m16ib(B, feas, horizon=H, seed=synthetic.SEED + seed_off)
s = cfg["sys_info"]
P = O.Problem(r, H, [synthetic.OBS_M16IB["l"]], [0.2], s["QQ"], s["lim"], s["MAX_input"], 0.1, 20)
t0 = time.time(); ref = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=8); print("oracle", time.time() - t0)
st = ref["status"] & 0xff
inf0 = (st == 2) & (ref["iters"] == 0)
print("status", np.bincount(st), "infeasible at iter 1:", inf0.sum())
dt = 0.5
a = s["MAX_input"][:nj]; v = s["lim"]
def reach(w0, i):
    """max and min displacement Bθ u at waypoint i (1-based) per joint"""
    hi = np.zeros(nj); lo = np.zeros(nj)
    for k in range(nj):
        for sgn, out in ((1, hi), (-1, lo)):
            w = w0[k]; th = 0.0
            for j in range(i):
                wn = w + sgn * dt * a[k]
                wn = min(v[k], wn) if sgn > 0 else max(-v[k], wn)
                th += dt * (w + wn) / 2; w = wn
            out[k] = th - i * dt * w0[k]
    return hi, lo
RH = [reach(np.zeros(nj), i) for i in range(1, H + 1)]
caught = np.zeros(B, bool); worst = np.zeros(B)
for b in np.nonzero(st >= 0)[0]:
    A, bb, dist, lid, grad, t = P.get_con(cfg["x0"][b], cfg["xref"][b], np.zeros(H * nj))
    m = -1e30
    for i in range(H):
        g = grad[i]; I = dist[i] - 0.2
        hi, lo = RH[i]
        # row: -g' dtheta <= I  ; min over reachable of -g'dtheta = -sum max(g*hi, g*lo)
        mn = -np.sum(np.maximum(g * hi, g * lo))
        m = max(m, mn - I)
    worst[b] = m; caught[b] = m > 1e-9
print("caught & infeasible@1:", (caught & inf0).sum(), "of", inf0.sum(), " caught but oracle feasible at iter 1 (must be 0):", (caught & ~inf0).sum())
try:
    d = np.load("gpurun_out/heavy_batch4.npz") if seed_off == 4 else None
except Exception: d = None
if d is not None:
    ps = d["ps"]
    idx = np.argsort(-ps)[:40]
    for i in idx: print(i, "steps", ps[i], "status", st[i], "iters", ref["iters"][i], "caught", caught[i], "worst %.3e" % worst[i], "oracle qp iters", ref["qp_iters"][i])
    print("steps of uncaught infeasible@1: ", np.sort(ps[inf0 & ~caught])[-20:])
    print("sum steps all", ps.sum(), "sum steps caught", ps[caught].sum())
np.savez("/tmp/reach_%d.npz" % seed_off, caught=caught, worst=worst, st=st, iters=ref["iters"])
