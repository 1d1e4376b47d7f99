"""Accuracy of the Gram-form dual active-set (GPU algorithm, numpy emulation) on the PSGCFS projection QPs."""
import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O
from tests import common
from dev.qp_proto import Proto
cfg = common.batch_m16ib(O, 96, horizon=30)
s = dict(cfg["sys_info"]); s["alpha"] = 1.0 / np.linalg.svd(s["QQ"], compute_uv=False).max()
H=30; nj=5; n=150; K=8
noise = np.random.default_rng(5).normal(0.0, 0.1, size=(96, 8, 150))
pr = Proto(np.eye(n), H, nj, 0.5)
res={}
for k in range(K+1):
    s["MAX_O_ITER"]=k
    P = common.oracle_problem(O, "M16iB", cfg["obs"], s, solver=1)
    res[k] = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], noise=np.ascontiguousarray(noise[:, :max(k,1), :]), nthreads=8) if k>0 else dict(u=np.zeros((96,n)), x=cfg["xref"].copy(), iters=np.zeros(96,int), status=np.zeros(96,int))
worst=[]
for k in range(K):
    for b in range(96):
        if res[k+1]["iters"][b] != k+1: continue
        u=res[k]["u"][b]; x=res[k]["x"][b]
        A,bb,dist,lid,grad,t = P.get_con(cfg["x0"][b], x, u)
        it=k+1
        ubar = u - s["alpha"]*((s["QQ"]@u + cfg["ff"][b]) + 10*noise[b,k]/(it*it+1))
        # proto: minimise 1/2|u-ubar|^2 -> ff = -ubar with identity metric
        disp = (x.reshape(H,10)[:,:5] - (cfg["x0"][b][:5] + (np.arange(1,H+1)[:,None]*0.5)*cfg["x0"][b][5:])) if k>0 else np.zeros((H,5))
        orhs = (dist-0.2) - (grad*disp).sum(1)
        for kw in (dict(refine=False, robust=False), dict(refine=False, robust=False, polish=1), dict(refine=False, robust=False, polish=2)):
            st,uu,steps,q,lam = pr.solve(-ubar, -grad, orhs, s["lim"], cfg["x0"][b][5:], 1e30*np.ones(n), dep_tol=1e-8, **kw)
            if st==0:
                worst.append((np.abs(uu-res[k+1]["u"][b]).max(), k, b, q, steps, kw.get("polish",0)))
worst.sort(reverse=True)
print(worst[:12])
print([w for w in worst if w[5]==1][:6]); print([w for w in worst if w[5]==2][:6])
