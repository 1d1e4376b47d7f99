"""small workload touching every kernel, for compute-sanitizer (memcheck / racecheck)"""
import sys; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M, oracle as O
from motionplanning_5d_m_b200 import synthetic, _lib, problem
from tests import common
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 48, 20
ctx = M.Context(0)
cfg = common.batch_m16ib(O, B, horizon=H)
s = cfg["sys_info"]; r = dict(cfg["robot"]); r["name"] = "M16iB"
ctx.set_robot(r, 5); ctx.set_obstacles(cfg["obs"]); ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
args = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
a = ctx.solve_batch(*args, 0.1, 20)                                     # fused: order, bulk, heavy
ctx.set_option("esc_steps", 2); b = ctx.solve_batch(*args, 0.1, 20); ctx.set_option("esc_steps", 48)   # heavy tier for most
ctx.set_option("fused", 0); c = ctx.solve_batch(*args, 0.1, 20); ctx.set_option("fused", 1)            # lock-step K1 + k_qp
d = ctx.solve_batch(*args, 0.1, 6, grad=_lib.GRAD_DERIVEST)             # K1d
ctx.set_cost(H, s["QQ"], s["lim"], None)
alpha = 1.0 / np.linalg.svd(s["QQ"], compute_uv=False).max()
e = ctx.solve_batch(*args, 0.1, 4, solver=_lib.SOLVER_PSGCFS, noise=np.random.default_rng(1).normal(0, .1, (B, 4, H * 5)), alpha=alpha)
ctx.set_cost_blocks(H, problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0, s["lim"], s["MAX_input"])
f = ctx.solve_start_goal(cfg["theta0"], cfg["thetag"], 0.1, 20)
routes = np.stack([np.linspace(cfg["theta0"][k], cfg["thetag"][k], 7) for k in range(8)])
g = ctx.solve_routes(routes, 0.1, 20)
ctx.get_con(cfg["x0"][0], cfg["xref"][0], np.zeros(H * 5))
ctx.dist_grad(cfg["theta0"]); ctx.dist_grad(cfg["theta0"], grad=_lib.GRAD_DERIVEST)
ctx.nodes_feasible(cfg["theta0"]); ctx.nearest_steer(cfg["theta0"], cfg["thetag"][:5], np.ones(5), 0.1)
rnd = np.random.default_rng(2).random((4, 6000))
h = ctx.rrt_find_routes(cfg["theta0"][:4], cfg["thetag"][:4], cfg["thetag"][:4], np.full(5, .3), synthetic.REGION_S, synthetic.SAMPLE_OFF, np.ones(5), rnd, max_iter=60)
print("statuses", np.bincount(a["status"] & 0xff, minlength=3), np.bincount(c["status"] & 0xff, minlength=3), "equal tiers", np.array_equal(a["status"], b["status"]), "rrt", h["n_nodes"])
