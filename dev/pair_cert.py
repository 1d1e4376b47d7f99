"""Pairwise relaxed Farkas certificate: rows (i,i+1) + per-joint {reach(i), reach(i+1), velocity-limited difference}."""
import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O
from tests import common
B=int(sys.argv[1]); GAP=int(sys.argv[2]) if len(sys.argv)>2 else 1
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250; dt=0.5
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
P1 = O.Problem(O.robot('M16iB'), H, [o['l'] for o in cfg['obs']], [0.2], s['QQ'], s['lim'], s['MAX_input'], 0.1, 1)
ref = P1.solve_batch(cfg['x0'], cfg['ff'], cfg['caug'], cfg['xref'], nthreads=8)
umax = s['MAX_input'].reshape(H,nj); lim=s['lim']
def reach(w0):
    out=[]
    for sign in (+1,-1):
        d=np.zeros((H,nj))
        for k in range(nj):
            cap_hi=(lim[k]-w0[k])/dt; cap_lo=-(lim[k]+w0[k])/dt
            S=0.0; th=0.0; om=0.0
            for j in range(H):
                tgt = cap_hi if sign>0 else cap_lo
                u = min(max(tgt-S, -umax[j,k]), umax[j,k]); S+=u
                th = th + dt*om + 0.5*dt*dt*u; om = om + dt*u
                d[j,k]=th
        out.append(d)
    return out  # dmax, dmin
lams=np.linspace(0,1,33)
def poly_min(al,be,amin,amax,bmin,bmax,dmin,dmax):
    # min al*a+be*b s.t. a in [amin,amax], b in [bmin,bmax], b-a in [dmin,dmax]: vertices of the polygon
    best=np.inf
    cands=[]
    for a in (amin,amax):
        lo=max(bmin,a+dmin); hi=min(bmax,a+dmax)
        if lo<=hi+1e-15: cands+= [(a,lo),(a,hi)]
    for b in (bmin,bmax):
        lo=max(amin,b-dmax); hi=min(amax,b-dmin)
        if lo<=hi+1e-15: cands+= [(lo,b),(hi,b)]
    if not cands: return np.inf   # empty polygon -> infeasible anyway
    return min(al*a+be*b for a,b in cands)
def pair_cert(oc, orhs, w0, gap):
    dmax,dmin=reach(w0)
    for i in range(H-gap):
        i2=i+gap
        for lam in lams:
            y1,y2=1-lam,lam
            tot=0.0
            for k in range(nj):
                al=y1*oc[i,k]; be=y2*oc[i2,k]
                tot+=poly_min(al,be,dmin[i,k],dmax[i,k],dmin[i2,k],dmax[i2,k], gap*dt*(-lim[k]-w0[k]), gap*dt*(lim[k]-w0[k]))
            if tot > y1*orhs[i]+y2*orhs[i2] + 1e-9*(abs(orhs[i])+abs(orhs[i2])+1): return (i,lam)
    return None
caught=0; ninf=0; fp=0; missed=[]
for b in range(B):
    rs = ref['status'][b]&0xff
    A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    oc=-grad.reshape(H,nj); orhs=(dist-0.2).reshape(H)
    c=None
    for gap in range(1,GAP+1):
        c=pair_cert(oc,orhs,cfg['x0'][b][5:],gap)
        if c: break
    if rs==2:
        ninf+=1
        if c: caught+=1
        else: missed.append((b,int(ref['qp_iters'][b])))
    elif c: fp+=1
print("infeasible",ninf,"caught",caught,"false positives",fp,"missed",missed)
