"""numpy prototype of k_qp.cu's primitive-space dual active-set solver (development aid, not shipped logic)."""
import sys; sys.path.insert(0, '.')
import numpy as np, oracle as O
from tests import common

def build_P(H, nj, dt):
    n = H*nj
    P = np.zeros((3*n, n))
    for i in range(H):
        for j in range(i+1):
            for k in range(nj):
                P[i*nj+k, j*nj+k] = 0.5*dt*dt + ((i-j)*dt)*dt
                P[n+i*nj+k, j*nj+k] = dt
    P[2*n:] = np.eye(n)
    return P

class Proto:
    def __init__(self, QQ, H, nj, dt):
        self.n = n = H*nj; self.H=H; self.nj=nj
        self.P = build_P(H, nj, dt)
        L = np.linalg.cholesky(QQ)
        self.Y = np.linalg.solve(L, self.P.T)      # n x 3n
        self.G = self.Y.T @ self.Y
        self.Hinv = self.G[2*n:, 2*n:]
    def solve(self, ff, ocoef, orhs, lim, w0, umax, refine=True, robust=True, dep_tol=1e-12, verbose=False, polish=0):
        n, H, nj = self.n, self.H, self.nj; G = self.G; Y = self.Y
        OH = len(orhs); m = OH + 4*n
        u0 = -self.Hinv @ ff; v = self.P @ u0
        # constraint coefficient vectors in primitive space (dense for the prototype)
        def evec(cid):
            e = np.zeros(3*n)
            if cid < OH:
                i = cid % H; e[i*nj:(i+1)*nj] = ocoef[cid]
            else:
                k = cid-OH; e[n + (k>>1)] = -1.0 if (k&1) else 1.0
            return e
        def rhs(cid):
            if cid < OH: return orhs[cid]
            k = cid-OH; idx = k>>1; neg = k&1
            if k < 2*n:
                j = idx % nj
                return lim[j]+w0[j] if neg else lim[j]-w0[j]
            return umax[idx-n]
        E = np.array([evec(c) for c in range(m)])     # m x 3n
        RHS = np.array([rhs(c) for c in range(m)])
        nrm = np.sqrt(np.einsum('ij,jk,ik->i', E, G, E))
        act = []; lam = []; M = np.zeros((0,0)); S = np.zeros((0,0)); steps = 0
        while True:
            sl = RHS - E @ v
            tol = 1e-11*(1+np.abs(RHS))
            viol = (sl < -tol); viol[act] = False
            if not viol.any():
                if polish and len(act):
                    v0_ = self.P @ u0; EW = E[act]; lam = np.array(lam)
                    bvec = EW @ v0_ - RHS[act]
                    for _ in range(polish):
                        Sl = EW @ (G @ (EW.T @ lam))
                        lam = lam + M @ (bvec - Sl)
                    v = v0_ - G @ (EW.T @ lam)
                return 0, v[2*n:], steps, len(act), np.array(lam)
            val = np.where(viol, sl/nrm, 0.0); p = int(np.argmin(val)); ep = E[p]; lam_p = 0.0
            while True:
                steps += 1
                if steps > 5000: return 3, None, steps, len(act), None
                q = len(act)
                EW = E[act] if q else np.zeros((0,3*n))
                g = EW @ (G @ ep); sigma = ep @ G @ ep
                r = M @ g if q else np.zeros(0)
                if refine and q: r = r + M @ (g - S @ r)
                delta = sigma - g @ r
                if robust and delta < 1e-3*sigma:
                    y = Y @ (ep - (EW.T @ r if q else 0)); delta = y @ y
                dependent = not (delta > dep_tol*sigma)
                t1, l = np.inf, -1
                for w in range(q):
                    if r[w] > 0 and lam[w]/r[w] < t1: t1, l = lam[w]/r[w], w
                sp = RHS[p] - ep @ v
                t2 = np.inf if dependent else max(0.0, -sp/delta)
                if l < 0 and dependent: return 2, None, steps, len(act), None
                full = t2 <= t1; t = t2 if full else t1
                lam = list(np.array(lam) - t*r) if q else []
                lam_p += t
                if (not dependent) and t > 0:
                    v = v - t*(G @ (ep - (EW.T @ r if q else 0)))
                if verbose: print(steps, 'p',p,'q',q,'delta/sigma',delta/sigma,'t1',t1,'t2',t2,'lam_p',lam_p)
                if full:
                    idl = 1.0/delta
                    Mn = np.zeros((q+1,q+1)); Mn[:q,:q] = M + np.outer(r,r)*idl; Mn[:q,q] = -r*idl; Mn[q,:q] = -r*idl; Mn[q,q]=idl
                    Sn = np.zeros((q+1,q+1)); Sn[:q,:q] = S; Sn[:q,q]=g; Sn[q,:q]=g; Sn[q,q]=sigma
                    M, S = Mn, Sn; act.append(p); lam.append(lam_p); break
                col = M[:,l].copy(); M = M - np.outer(col,col)/col[l]
                keep = [w for w in range(q) if w != l]
                M = M[np.ix_(keep,keep)]; S = S[np.ix_(keep,keep)]
                act.pop(l); lam.pop(l)

if __name__ == '__main__':
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250
    P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
    ref = P.solve_batch(cfg['x0'], cfg['ff'], cfg['caug'], cfg['xref'], nthreads=8)
    pr = Proto(s['QQ'], H, nj, 0.5)
    sel = [int(a) for a in sys.argv[2:]] if len(sys.argv) > 2 else list(np.where((ref['status']&0xff)==2)[0])
    for b in sel:
        A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
        ocoef = -grad; orhs = dist - 0.2
        for kw in (dict(refine=False, robust=False), dict(refine=True, robust=True)):
            st, u, steps, q, lam = pr.solve(cfg['ff'][b], ocoef, orhs, s['lim'], cfg['x0'][b][5:], s['MAX_input'], **kw)
            print(b, kw, 'proto status', st, 'steps', steps, 'q', q, '| oracle', ref['status'][b]&0xff, ref['qp_iters'][b], ref['qp_max_active'][b],
                  '' if lam is None else 'max lam %.3g' % (lam.max() if len(lam) else 0))
