"""batched RRT tree growth: kernel time for S seeds vs the C restatement on the host cores"""
import sys, time; sys.path.insert(0, '.')
import numpy as np, motionplanning_5d_m_b200 as M, oracle as O
from tests import common
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ROBOT, robot, obs, s = common.rrtstar_cfs_config(np.zeros((5, 41)))
ctx = M.Context(0); r = dict(robot); r["name"] = ROBOT; ctx.set_robot(r, 5); ctx.set_obstacles(obs)
x0 = np.array([0.421, 0, -0.0092, -0.0010, -1.5786]); goal = np.array([-1.4090, 0.8873, 0.4008, 0.0, 0.4430])
rg = np.array([np.pi / 20, np.pi / 20, np.pi / 10, np.pi / 2, np.pi / 2]); rs = np.array([np.pi / 2, np.pi / 2, np.pi / 2, np.pi / 1.5, np.pi / 1.5])
rat = np.array([1, 1, 0.5, 0.1, 0.1]); rnd = np.random.default_rng(1).random((S, 8000))
X0, G = np.tile(x0, (S, 1)), np.tile(goal, (S, 1))
for star in (True, False):
    for rep in range(2):
        out = ctx.rrt_find_routes(X0, G, G, rg, rs, np.zeros(5), rat, rnd, star=star)
    print("%s: %d seeds in %.2f ms (kernel) = %.1f k trees/s; found %d, failed %d, exhausted %d, mean nodes %.0f" % (
        "RRT*" if star else "RRT", S, out["ms"], S / out["ms"], int((~out["fail"] & (out["route_len"] > 0)).sum()), int(out["fail"].sum()),
        int((out["route_len"] < 0).sum()), out["n_nodes"].mean()))
rob = O.robot(ROBOT); n = min(S, 256); t = time.time()
for k in range(n):
    O.rrt_find_route(rob, [o["l"] for o in obs], [o["D"] for o in obs], x0, goal, rg, rs, np.zeros(5), goal, rat, rnd[k], star=True)
dt = time.time() - t
print("C restatement, 1 thread: %d trees in %.2f s = %.2f k trees/s" % (n, dt, n / dt / 1e3))
