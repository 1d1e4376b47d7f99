import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O
from tests import common
from dev.qp_proto import Proto
B=int(sys.argv[1]); 
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250; dt=0.5
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
pr = Proto(s['QQ'], H, nj, dt)
QQ=s['QQ']; umax=s['MAX_input']
qqinf=np.abs(0.5*(QQ+QQ.T)).sum(1).max()
lmax=np.linalg.eigvalsh(QQ).max()
print("qq inf", qqinf, "lmax", lmax)
def solve(self, ff, caug, ocoef, orhs, lim, w0, umax, dep_tol=1e-8):
    n, H, nj = self.n, self.H, self.nj; G = self.G
    OH = len(orhs); m = OH + 4*n
    u0 = -self.Hinv @ ff; v0 = self.P @ u0; v=v0.copy()
    fval = caug + 0.5*ff@u0
    fupper = (abs(caug) + (0.5*qqinf*umax**2 + np.abs(ff)*umax).sum())
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
    E = np.array([evec(c) for c in range(m)]); RHS = np.array([rhs(c) for c in range(m)])
    nrm = np.sqrt(np.einsum('ij,jk,ik->i', E, G, E))
    act = []; lam = np.zeros(0); M = np.zeros((0,0)); steps = 0; trace=[]
    while True:
        if len(act): v = v0 - G @ (E[act].T @ lam)
        sl = RHS - E @ v
        tol = 1e-11*(1+np.abs(RHS))
        viol = (sl < -tol); viol[act] = False
        if not viol.any(): return 0, steps, len(act), trace, fupper
        val = np.where(viol, sl/nrm, 0.0); p = int(np.argmin(val)); ep = E[p]; lam_p = 0.0
        sp = sl[p]
        while True:
            steps += 1
            q = len(act)
            EW = E[act] if q else np.zeros((0,3*n))
            g = EW @ (G @ ep); sigma = ep @ G @ ep
            r = M @ g if q else np.zeros(0)
            delta = sigma - g @ r
            dependent = not (delta > dep_tol*sigma) or q>=n
            t1, l = np.inf, -1
            for w in range(q):
                if r[w] > 0 and lam[w]/r[w] < t1: t1, l = lam[w]/r[w], w
            t2 = np.inf if dependent else max(0.0, -sp/delta)
            if l < 0 and dependent: return 2, steps, len(act), trace, fupper
            full = t2 <= t1; t = t2 if full else t1
            if not dependent:
                fval += t*delta*(0.5*t+lam_p); sp += t*delta
            nobs=sum(1 for a in act if a<OH); nvel=sum(1 for a in act if OH<=a<OH+2*n); nb=q-nobs-nvel
            trace.append((steps,p if p<OH else ('v' if p<OH+2*n else 'b'),q,nobs,nvel,nb,fval,float(lam.max()) if q else 0, full))
            if fval > fupper: return 22, steps, len(act), trace, fupper
            lam = lam - t*r; lam_p += t
            if full:
                idl = 1.0/delta
                Mn = np.zeros((q+1,q+1)); Mn[:q,:q] = M + np.outer(r,r)*idl; Mn[:q,q] = -r*idl; Mn[q,:q] = -r*idl; Mn[q,q]=idl
                M = Mn; act.append(p); lam=np.append(lam,lam_p); break
            col = M[:,l].copy(); M = M - np.outer(col,col)/col[l]
            keep = [w for w in range(q) if w != l]
            M = M[np.ix_(keep,keep)]; act.pop(l); lam=lam[keep]
res=[]
for b in range(B):
    A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    st, steps, q, trace, fupper = solve(pr, cfg['ff'][b], cfg['caug'][b], -grad, dist-0.2, s['lim'], cfg['x0'][b][5:], umax)
    res.append((b,st,steps,q))
    if st>=2 and steps>30:
        print("problem",b,"status",st,"steps",steps,"q",q,"fupper %.3g"%fupper, "caug %.3g"%cfg['caug'][b])
        for tr in trace[::max(1,len(trace)//25)]: print("   step %d p %s q %d (obs %d vel %d bnd %d) fval %.4g lammax %.3g full %s"%tr)
inf=[r for r in res if r[1]>=2]
print("infeasible", len(inf), "steps mean", np.mean([r[2] for r in inf]), "max", max(r[2] for r in inf))
