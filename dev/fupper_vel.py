"""CPU study: how early does the dual objective of the long infeasible chains cross a TIGHTER upper bound of the cost, built
from QQ alone but over the velocity box as well as the control box?   python dev/fupper_vel.py [B]
  box bound:       1/2 ||QQ||_inf sum umax^2 + sum |ff| umax                               (k_v0, round 1)
  abs bound:       1/2 umax'|QQ|umax + |ff|'umax                                          (entrywise, same superset)
  velocity bound:  u = D w / dt (w = joint velocities, |w| <= lim):  1/2 l'|D'QQD|l/dt^2 + |D'ff|'l/dt + const(w0)"""
import sys; sys.path.insert(0, '.')
import numpy as np, oracle as O
from tests import common
from dev.qp_proto import Proto
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H = 50; nj = 5; n = 250; dt = 0.5
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
pr = Proto(s['QQ'], H, nj, dt)
QQ = 0.5 * (s['QQ'] + s['QQ'].T); umax = s['MAX_input']; lim = s['lim']
qqinf = np.abs(QQ).sum(1).max()
# D: u = D w / dt with w the stacked joint velocities (w0 = 0 in the headline batch): u_i = (w_i - w_{i-1}) / dt
D = np.eye(n) - np.eye(n, k=-nj)
Dt = D / dt
QV = Dt.T @ QQ @ Dt
lv = np.tile(lim, H)
quad_box = 0.5 * qqinf * (umax ** 2).sum(); quad_abs = 0.5 * umax @ np.abs(QQ) @ umax; quad_vel = 0.5 * lv @ np.abs(QV) @ lv
print("quadratic part: box %.4g  abs %.4g  velocity %.4g" % (quad_box, quad_abs, quad_vel))

def solve(self, ff, caug, ocoef, orhs, lim, w0, umax, fuppers, dep_tol=1e-8):
    n, H, nj = self.n, self.H, self.nj; G = self.G
    OH = len(orhs); m = OH + 4 * n
    u0 = -self.Hinv @ ff; v0 = self.P @ u0; v = v0.copy()
    fval = caug + 0.5 * ff @ u0
    E = np.zeros((m, 3 * n)); RHS = np.zeros(m)
    for cid in range(OH):
        i = cid % H; E[cid, i * nj:(i + 1) * nj] = ocoef[cid]; RHS[cid] = orhs[cid]
    for k in range(4 * n):
        idx = k >> 1; neg = k & 1
        E[OH + k, n + idx] = -1.0 if neg else 1.0
        RHS[OH + k] = (lim[idx % nj] + w0[idx % nj] if neg else lim[idx % nj] - w0[idx % nj]) if k < 2 * n else umax[idx - n]
    nrm = np.sqrt(np.einsum('ij,jk,ik->i', E, G, E))
    act = []; lam = np.zeros(0); M = np.zeros((0, 0)); steps = 0
    crossed = [None] * len(fuppers)
    while True:
        if len(act): v = v0 - G @ (E[act].T @ lam)
        sl = RHS - E @ v
        tol = 1e-11 * (1 + np.abs(RHS))
        viol = (sl < -tol); viol[act] = False
        if not viol.any(): return 0, steps, len(act), crossed
        val = np.where(viol, sl / nrm, 0.0); p = int(np.argmin(val)); ep = E[p]; lam_p = 0.0
        sp = sl[p]
        while True:
            steps += 1
            q = len(act)
            EW = E[act] if q else np.zeros((0, 3 * n))
            g = EW @ (G @ ep); sigma = ep @ G @ ep
            r = M @ g if q else np.zeros(0)
            delta = sigma - g @ r
            dependent = not (delta > dep_tol * sigma) or q >= n
            t1, l = np.inf, -1
            for w in range(q):
                if r[w] > 0 and lam[w] / r[w] < t1: t1, l = lam[w] / r[w], w
            t2 = np.inf if dependent else max(0.0, -sp / delta)
            if l < 0 and dependent: return 2, steps, len(act), crossed
            full = t2 <= t1; t = t2 if full else t1
            if not dependent:
                fval += t * delta * (0.5 * t + lam_p); sp += t * delta
            for j, fu in enumerate(fuppers):
                if crossed[j] is None and fval > fu: crossed[j] = steps
            if crossed[0] is not None: return 22, steps, len(act), crossed
            lam = lam - t * r; lam_p += t
            if full:
                idl = 1.0 / delta
                Mn = np.zeros((q + 1, q + 1)); Mn[:q, :q] = M + np.outer(r, r) * idl; Mn[:q, q] = -r * idl; Mn[q, :q] = -r * idl; Mn[q, q] = idl
                M = Mn; act.append(p); lam = np.append(lam, lam_p); break
            col = M[:, l].copy(); M = M - np.outer(col, col) / col[l]
            keep = [w for w in range(q) if w != l]
            M = M[np.ix_(keep, keep)]; act.pop(l); lam = lam[keep]

tot = np.zeros(3); cnt = 0
for b in range(B):
    ff, caug = cfg['ff'][b], cfg['caug'][b]
    A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    lin_box = (np.abs(ff) * umax).sum(); lin_vel = (np.abs(Dt.T @ ff) * lv).sum()
    f_box = (abs(caug) + quad_box + lin_box)
    f_abs = (abs(caug) + quad_abs + lin_box)
    f_vel = (abs(caug) + min(quad_abs, quad_vel) + min(lin_box, lin_vel))
    st, steps, q, crossed = solve(pr, ff, caug, -grad, dist - 0.2, lim, cfg['x0'][b][5:], umax, [f_box, f_abs, f_vel])
    if st >= 2 and steps > 20:
        c = [x if x is not None else steps for x in crossed]
        print("problem %d status %d steps %d q %d | box %.3g abs %.3g vel %.3g | crossed at %s" % (b, st, steps, q, f_box, f_abs, f_vel, c))
        tot += np.array(c); cnt += 1
print("long chains", cnt, "mean steps: box %.1f abs %.1f vel %.1f" % tuple(tot / max(cnt, 1)))
