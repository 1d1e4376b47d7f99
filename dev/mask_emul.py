import sys; sys.path.insert(0,'.')
import numpy as np, oracle as O, warnings
warnings.filterwarnings('ignore')
from tests import common
from dev.qp_proto import Proto
from motionplanning_5d_m_b200 import problem, robot as R
d=np.load("gpurun_out/steps_dump.npz")
H=50; nj=5; n=250
rob=R.robotproperty2("M16iB")
Aaug,Baug,Qaug,QQ=problem.build_cost_matrices(rob,nj,H,problem.Q_MAIN_FANUC,problem.R_MAIN_FANUC,50.0)
obs=[dict(l=np.array([[3.906,3.906],[8.313,8.313],[0.001,1.938]]),D=0.2,epsilon=0.2)]
s=dict(H=H,QQ=QQ,lim=np.ones(5),MAX_input=np.tile(np.array([1,1,np.pi,np.pi,np.pi])*0.5,H),epsilon_O=0.1,MAX_O_ITER=20)
P = common.oracle_problem(O,"M16iB",obs,s)
pr = Proto(QQ,H,nj,0.5)
for k in range(len(d["idx"])):
    if d["iters"][k]!=0 or (d["status"][k]&0xff)!=2: continue
    th0,thg=d["theta0"][k],d["thetag"][k]
    x0=np.concatenate([th0,np.zeros(5)]); xref=problem.straight_line_reference(th0,thg,H)[0]
    gaug=np.tile(np.concatenate([thg,np.zeros(5)]),H)
    ff,caug=problem.build_linear_term(Aaug,Baug,Qaug,x0,gaug)
    A,bb,dist,lid,grad,t=P.get_con(x0,xref,np.zeros(n))
    oc=-grad; orhs=dist-0.2
    aa=(oc*oc).sum(1); ab=(oc[:-1]*oc[1:]).sum(1)
    flag=(ab<0)&(ab*ab>0.81*aa[:-1]*aa[1:])
    keep=np.zeros(H,bool); keep[:-1]|=flag; keep[1:]|=flag
    rh=np.where(keep,orhs,1e30)
    st,_,steps,q,_=pr.solve(ff[0],oc,rh,s['lim'],x0[5:],s['MAX_input'],refine=False,robust=False,dep_tol=1e-8)
    print("problem",int(d["idx"][k]),"gpu steps",int(d["steps"][k]),"unmasked rows",np.where(keep)[0],"masked-phase GI: status",st,"steps",steps,"q",q, "orhs(unmasked)",np.round(orhs[keep],3))
