"""Structure of the remaining iteration-1 stragglers: which obstacle-row subsets (with velocity/bound rows) are infeasible?"""
import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O
from scipy.optimize import linprog
from tests import common
from motionplanning_5d_m_b200 import problem, robot as R
d=np.load("gpurun_out/steps_dump.npz")
H=50; nj=5; n=250
rob=R.robotproperty2("M16iB")
Aaug,Baug,Qaug,QQ=problem.build_cost_matrices(rob,nj,H,problem.Q_MAIN_FANUC,problem.R_MAIN_FANUC,50.0)
obs=[dict(l=np.array([[3.906,3.906],[8.313,8.313],[0.001,1.938]]),D=0.2,epsilon=0.2)]
s=dict(H=H,QQ=QQ,lim=np.ones(5),MAX_input=np.tile(np.array([1,1,np.pi,np.pi,np.pi])*0.5,H),epsilon_O=0.1,MAX_O_ITER=20)
P = common.oracle_problem(O,"M16iB",obs,s)
bounds=list(zip(-s['MAX_input'], s['MAX_input']))
def feas(A,b):
    r=linprog(np.zeros(n), A_ub=A, b_ub=b, bounds=bounds, method="highs"); return r.status==0
for k in range(len(d["idx"])):
    if d["iters"][k]!=0 or (d["status"][k]&0xff)!=2: continue
    th0,thg=d["theta0"][k],d["thetag"][k]
    x0=np.concatenate([th0,np.zeros(5)]); xref=problem.straight_line_reference(th0,thg,H)[0]
    A,bb,dist,lid,grad,t=P.get_con(x0,xref,np.zeros(n))
    oc=-grad; gn=oc/np.linalg.norm(oc,axis=1,keepdims=True); cosv=(gn[:-1]*gn[1:]).sum(1)
    rows_obs=[i*11 for i in range(H)]; vel=[r for r in range(A.shape[0]) if r%11!=0]
    viol=[i for i in range(H) if dist[i]<0.2]
    res={}
    pairs=[(i,j) for i in range(H) for j in range(i+1,min(H,i+4)) if not feas(A[[rows_obs[i],rows_obs[j]]+vel], bb[[rows_obs[i],rows_obs[j]]+vel])]
    print("problem",int(d["idx"][k]),"steps",int(d["steps"][k]),"violated",(viol[0],viol[-1]) if viol else None,"min cos %.4f at %d"%(cosv.min(),cosv.argmin()),
          "lid changes", [(i,int(lid[i]),int(lid[i+1])) for i in range(H-1) if lid[i]!=lid[i+1]], "infeasible pairs (gap<=3):", pairs[:6],
          "cos of those:", [round(float(gn[i]@gn[j]),4) for i,j in pairs[:6]])
