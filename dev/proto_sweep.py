import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, warnings
warnings.filterwarnings('ignore')
from tests import common
from dev.qp_proto import Proto
B=int(sys.argv[1]); tol=float(sys.argv[2])
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
pr = Proto(s['QQ'], H, nj, 0.5)
# oracle first-QP verdict: run with max_outer=1
P1 = O.Problem(O.robot('M16iB'), H, [o['l'] for o in cfg['obs']], [0.2], s['QQ'], s['lim'], s['MAX_input'], 0.1, 1)
ref = P1.solve_batch(cfg['x0'], cfg['ff'], cfg['caug'], cfg['xref'], nthreads=8)
mis=0; steps_inf=[]; steps_ok=[]; qs=[]
for b in range(B):
    A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    st, u, steps, q, lam = pr.solve(cfg['ff'][b], -grad, dist-0.2, s['lim'], cfg['x0'][b][5:], s['MAX_input'], refine=False, robust=False, dep_tol=tol)
    rs = ref['status'][b]&0xff
    rs = 0 if rs==1 else rs
    if st != rs: mis+=1; print('MISMATCH', b, st, rs, steps, q)
    elif st==0:
        du = np.abs(u-ref['u'][b]).max(); steps_ok.append(steps)
        if du>1e-8: print('du', b, du)
    else: steps_inf.append(steps); qs.append(q)
print('mismatches', mis, 'feasible steps mean/max', np.mean(steps_ok), np.max(steps_ok), 'infeasible steps mean/max', np.mean(steps_inf), np.max(steps_inf), 'q mean/max', np.mean(qs), np.max(qs), 'n_inf', len(steps_inf))
