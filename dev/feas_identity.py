"""Does an identity-metric feasibility QP (min 1/2|u|^2 s.t. rows) detect infeasibility in fewer dual steps?"""
import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, warnings
warnings.filterwarnings('ignore')
from tests import common
from dev.qp_proto import Proto
B=int(sys.argv[1])
cfg = common.batch_m16ib(O, B); s = cfg['sys_info']; H=50; nj=5; n=250
P = common.oracle_problem(O, 'M16iB', cfg['obs'], s)
prQ = Proto(s['QQ'], H, nj, 0.5)
prI = Proto(np.eye(n), H, nj, 0.5)
P1 = O.Problem(O.robot('M16iB'), H, [o['l'] for o in cfg['obs']], [0.2], s['QQ'], s['lim'], s['MAX_input'], 0.1, 1)
ref = P1.solve_batch(cfg['x0'], cfg['ff'], cfg['caug'], cfg['xref'], nthreads=8)
res=[]
for b in range(B):
    rs = ref['status'][b]&0xff
    A_, b_, dist, lid, grad, t_ = P.get_con(cfg['x0'][b], cfg['xref'][b], np.zeros(n))
    stQ, u, stepsQ, qQ, lam = prQ.solve(cfg['ff'][b], -grad, dist-0.2, s['lim'], cfg['x0'][b][5:], s['MAX_input'], refine=False, robust=False, dep_tol=1e-8)
    stI, u, stepsI, qI, lam = prI.solve(np.zeros(n), -grad, dist-0.2, s['lim'], cfg['x0'][b][5:], s['MAX_input'], refine=False, robust=False, dep_tol=1e-8)
    res.append((b, rs, stQ, stepsQ, qQ, stI, stepsI, qI))
    if (stI==2) != (rs==2): print("VERDICT MISMATCH", res[-1])
inf=[r for r in res if r[1]==2]
print("infeasible", len(inf))
print("QQ metric  steps mean %.1f max %d" % (np.mean([r[3] for r in inf]), max(r[3] for r in inf)))
print("I  metric  steps mean %.1f max %d" % (np.mean([r[6] for r in inf]), max(r[6] for r in inf)))
print("worst (QQ steps, I steps, qQ, qI):", sorted([(r[3], r[6], r[4], r[7]) for r in inf], reverse=True)[:12])
fe=[r for r in res if r[1]!=2]
print("feasible: I-metric steps mean %.1f max %d ; QQ steps mean %.1f max %d" % (np.mean([r[6] for r in fe]), max(r[6] for r in fe), np.mean([r[3] for r in fe]), max(r[3] for r in fe)))
