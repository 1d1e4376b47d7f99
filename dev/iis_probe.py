"""Which small subsets of obstacle rows (with all velocity/bound rows) already make an iteration-1 QP infeasible?"""
import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O
from scipy.optimize import linprog
from tests import common
B=int(sys.argv[1])
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
P1 = O.Problem(O.robot('M16iB'), H, [o['l'] for o in cfg['obs']], [0.2], s['QQ'], s['lim'], s['MAX_input'], 0.1, 1)
ref = P1.solve_batch(cfg['x0'], cfg['ff'], cfg['caug'], cfg['xref'], nthreads=8)
inf=[b for b in range(B) if (ref['status'][b]&0xff)==2]
print("infeasible", len(inf))
bounds=list(zip(-s['MAX_input'], s['MAX_input']))
def feas(A,b):
    r=linprog(np.zeros(n), A_ub=A, b_ub=b, bounds=bounds, method="highs"); return r.status==0
stats={"pair":0,"triple":0,"window5":0,"none":0}
for b in inf[:40]:
    A,bb,dist,lid,grad,t = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    rows_obs=[i*11 for i in range(H)]
    vel=[r for r in range(A.shape[0]) if r%11!=0]
    viol=[i for i in range(H) if dist[i]<0.2]
    found=None
    for i in range(H-1):
        idx=[rows_obs[i],rows_obs[i+1]]+vel
        if not feas(A[idx],bb[idx]): found=("pair",i); break
    if not found:
        for i in range(H-2):
            idx=[rows_obs[i],rows_obs[i+1],rows_obs[i+2]]+vel
            if not feas(A[idx],bb[idx]): found=("triple",i); break
    if not found:
        for i in range(H-4):
            idx=[rows_obs[i+k] for k in range(5)]+vel
            if not feas(A[idx],bb[idx]): found=("window5",i); break
    stats[found[0] if found else "none"]+=1
    print(b, "qp_iters", ref['qp_iters'][b], "violated waypoints", (viol[0],viol[-1],len(viol)) if viol else None, found)
print(stats)
