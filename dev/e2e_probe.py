import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import synthetic
NC=int(sys.argv[1]); steps=int(sys.argv[2]); mode=sys.argv[3]
dev=torch.device("cuda",0); B,H,nj=4096,50,5; n=250; N=500; K=20
robot=dict(M.robotproperty2("M16iB")); robot["name"]="M16iB"
ctxs=[]; 
for c in range(NC):
    ctx=M.Context(0); ctx.set_robot(robot,nj); ctx.set_obstacles([synthetic.OBS_M16IB]); ctxs.append(ctx)
cfg=synthetic.batch_config_m16ib(B, lambda cand: ctxs[0].nodes_feasible(cand)[0])
s=cfg["sys_info"]
for ctx in ctxs: ctx.set_cost(H,s["QQ"],s["lim"],s["MAX_input"])
hin=[{k: torch.from_numpy(cfg[k]).pin_memory() for k in ("x0","ff","caug","xref")} for c in range(NC)]
mk=lambda *sh, dt=torch.float64: torch.empty(sh,dtype=dt).pin_memory()
hout=[dict(u=mk(B,n),x=mk(B,N),cost=mk(B,K),eu=mk(B,K),iters=mk(B,dt=torch.int32),status=mk(B,dt=torch.int32)) for c in range(NC)]
din=[{k: v.to(dev) for k,v in hin[c].items()} for c in range(NC)]
dout=[{k: v.to(dev) for k,v in hout[c].items()} for c in range(NC)]
def issue(c):
    if mode.startswith("dev"):
        i,o=din[c],dout[c]
        ctxs[c].solve_batch_ptr(B,i["x0"].data_ptr(),i["ff"].data_ptr(),i["caug"].data_ptr(),i["xref"].data_ptr(),0.1,K,o["u"].data_ptr(),o["x"].data_ptr(),o["cost"].data_ptr(),o["eu"].data_ptr(),o["iters"].data_ptr(),o["status"].data_ptr(),device=True,sync=False)
        return
    i,o=hin[c],hout[c]
    ctxs[c].solve_batch_ptr(B,i["x0"].data_ptr(),i["ff"].data_ptr(),i["caug"].data_ptr(),i["xref"].data_ptr(),0.1,K,o["u"].data_ptr(),o["x"].data_ptr(),o["cost"].data_ptr(),o["eu"].data_ptr(),o["iters"].data_ptr(),o["status"].data_ptr(),device=False,sync=False)
tw=[0.0]; ti=[0.0]
def run(steps):
    torch.cuda.synchronize(); t0=time.perf_counter(); tw[0]=0; ti[0]=0
    for k in range(steps):
        c=k%NC
        a=time.perf_counter()
        if mode in ("wait","devwait") and k>=NC: ctxs[c].wait()
        b=time.perf_counter(); tw[0]+=b-a
        issue(c)
        ti[0]+=time.perf_counter()-b
    for ctx in ctxs: ctx.wait()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)*1e3/steps
run(NC*2)
print(mode, "NC",NC,"ms/step", run(steps), run(steps), "host ms/step: wait %.3f issue %.3f"%(tw[0]*1e3/steps, ti[0]*1e3/steps))
for c in range(min(NC,4)): st=ctxs[c].stats(); print("  ctx",c,"h2d %.2f d2h %.2f total(solve) %.2f"%(st["ms_h2d"],st["ms_d2h"],st["ms_total"]))
