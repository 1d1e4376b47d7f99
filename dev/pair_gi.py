"""GI steps needed to prove infeasibility when only ONE consecutive pair of obstacle rows is kept (plus velocity/bound rows)."""
import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, warnings
warnings.filterwarnings('ignore')
from tests import common
from dev.qp_proto import Proto
B=int(sys.argv[1])
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
pr = Proto(s['QQ'], H, nj, 0.5)
P1 = O.Problem(O.robot('M16iB'), H, [o['l'] for o in cfg['obs']], [0.2], s['QQ'], s['lim'], s['MAX_input'], 0.1, 1)
ref = P1.solve_batch(cfg['x0'], cfg['ff'], cfg['caug'], cfg['xref'], nthreads=8)
inf=[b for b in range(B) if (ref['status'][b]&0xff)==2]
out=[]
for b in inf:
    A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    oc=-grad; orhs=dist-0.2
    stF,_,stepsF,qF,_ = pr.solve(cfg['ff'][b], oc, orhs, s['lim'], cfg['x0'][b][5:], s['MAX_input'], refine=False, robust=False, dep_tol=1e-8)
    # heuristic pair: consecutive rows with the most negative normalised inner product of their coefficient vectors
    gn = oc/np.maximum(np.linalg.norm(oc,axis=1,keepdims=True),1e-300)
    cosv = (gn[:-1]*gn[1:]).sum(1)
    viol = orhs<0
    cand = [i for i in range(H-1) if viol[i] or viol[i+1]]
    i0 = min(cand, key=lambda i: cosv[i]) if cand else int(np.argmin(cosv))
    best=None
    for i in (i0,):
        rh = np.full(H, 1e30); rh[i]=orhs[i]; rh[i+1]=orhs[i+1]
        st,_,steps,q,_ = pr.solve(cfg['ff'][b], oc, rh, s['lim'], cfg['x0'][b][5:], s['MAX_input'], refine=False, robust=False, dep_tol=1e-8)
        best=(i,st,steps,q, float(cosv[i]))
    out.append((b,stepsF,qF,best))
    print(b,"full steps",stepsF,"q",qF,"| pair",best)
ok=[o for o in out if o[3][1]==2]
print("pair-restricted infeasible:",len(ok),"of",len(out),"; steps mean %.1f max %d"%(np.mean([o[3][2] for o in ok]),max(o[3][2] for o in ok)), "| full steps mean %.1f max %d"%(np.mean([o[1] for o in out]),max(o[1] for o in out)))
