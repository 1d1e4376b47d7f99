"""CPU study: does a closed-form PAIR certificate over the control box alone (velocity rows relaxed) prove the
first-iteration infeasible QPs infeasible?   max_mu  min_{|u|<=umax} [(mu a1 + (1-mu) a2).u - (mu b1 + (1-mu) b2)] > 0"""
import sys; sys.path.insert(0, '.')
import numpy as np, oracle as O
from motionplanning_5d_m_b200 import synthetic
seed_off = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B, H, nj = 4096, 50, 5
r = O.robot("M16iB"); o6 = O.obs6(synthetic.OBS_M16IB["l"])
feas = lambda cand: np.array([O.dist_arm(r, th, o6)[0] >= 0.2 for th in cand])
cfg = synthetic.batch_config_m16ib(B, feas, horizon=H, seed=synthetic.SEED + seed_off)
s = cfg["sys_info"]
P = O.Problem(r, H, [synthetic.OBS_M16IB["l"]], [0.2], s["QQ"], s["lim"], s["MAX_input"], 0.1, 20)
d = np.load("/tmp/reach_%d.npz" % seed_off); st, iters = d["st"], d["iters"]
inf0 = (st == 2) & (iters == 0)
umax = s["MAX_input"]
mus = np.linspace(0, 1, 33)
def cert(b, pairs="all"):
    A, bb, dist, lid, grad, t = P.get_con(cfg["x0"][b], cfg["xref"][b], np.zeros(H * nj))
    Ao, bo = A[0::11], bb[0::11]                      # obstacle rows
    viol = np.where(bo < 0)[0]                        # rows violated at u = 0 ... candidates
    best = -1e30; arg = None
    rows = range(H)
    for i1 in rows:
        for i2 in range(i1 + 1, H):
            if pairs == "consecutive" and i2 != i1 + 1: continue
            if bo[i1] >= 0 and bo[i2] >= 0: continue
            C = mus[:, None] * Ao[i1][None, :] + (1 - mus)[:, None] * Ao[i2][None, :]
            g = -(np.abs(C) * umax[None, :]).sum(1) - (mus * bo[i1] + (1 - mus) * bo[i2])
            if g.max() > best: best = g.max(); arg = (i1, i2, mus[g.argmax()])
    return best, arg
idx = np.where(inf0)[0]
try:
    ps = np.load("gpurun_out/heavy_batch4.npz")["ps"] if seed_off == 4 else None
except Exception: ps = None
order = idx if ps is None else idx[np.argsort(-ps[idx])]
ncert = 0; tested = 0
for b in order[:80]:
    best, arg = cert(b)
    tested += 1; ncert += best > 1e-9
    if tested <= 25: print("problem %d gpu steps %s: best pair margin %.3e at %s -> %s" % (b, None if ps is None else ps[b], best, arg, "CERTIFIED" if best > 1e-9 else "-"))
print("certified %d of %d tested infeasible-at-1 problems" % (ncert, tested))
# safety: feasible problems must never be certified
ok = np.where((st < 2))[0][:60]
bad = sum(cert(b)[0] > 1e-9 for b in ok)
print("false certificates on %d feasible problems: %d" % (len(ok), bad))
