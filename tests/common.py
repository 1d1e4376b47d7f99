"""Shared problem builders for the parity tests (reference configurations)."""
import numpy as np

import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import problem, synthetic


def main_fanuc_config():
    """main_FANUC.m:13-60,106-127 : M200i, H=30, one capsule obstacle."""
    robot = M.robotproperty2("M200i")
    x0 = [0.7825, 0.0284, 0.2172, 0.1444, -1.1779]
    xg = [-0.7825, 0.0284, 0.2172, 0.1444, -1.1779]
    s = M.make_sys_info(robot, 5, 30, x0, xg)
    obs = [{"l": np.array([[3.806, 3.606], [8.413, 8.413], [0.001, 1.038]]), "D": 0.2, "epsilon": 0.25}]
    return "M200i", robot, obs, s


def main_2l_config():
    """main_2L.m:14-121 : two-link planar arm, H=40, point obstacle, stationary reference."""
    robot = M.robotproperty2("2L")
    nj, H = 2, 40
    x0, xg = [0.0, 0.0], [np.pi / 2, 0.0]
    s = M.make_sys_info(robot, nj, H, x0, xg, Q=problem.Q_2L, Rblk=problem.R_2L, r_scale=0.1, lim=[0.1, 0.2],
                        max_input=np.tile(np.array([1.0, 1.0]) * 0.5 * robot["delta_t"], H), epsilon_O=1e-6,
                        MAX_O_ITER=100, x_ref=np.tile(np.array([0.0, 0.0, 0.0, 0.0]), H))
    obs = [{"l": np.array([[0.3, 0.3], [0.3, 0.3], [0.0, 0.0]]), "D": 0.05, "epsilon": 0.05}]
    return "2L", robot, obs, s


def rrtstar_cfs_config(route):
    """RRTstar_CFS.m:40-50,96-187 : M200i, H=40, two obstacles, R*10, Q_v=[100,20,1,1,1]; route = 5 x (H+1) samples."""
    robot = M.robotproperty2("M200i")
    H = 40
    s = M.make_sys_info(robot, 5, H, route[:, 0], route[:, -1], Q=problem.Q_RRTSTAR, r_scale=10.0,
                        x_ref=np.concatenate([np.concatenate([route[:, i], np.zeros(5)]) for i in range(1, H + 1)]))
    obs = [{"l": np.array([[3.606, 3.606], [8.413, 8.413], [0.001, 1.038]]), "D": 0.2, "epsilon": 0.2},
           {"l": np.array([[3.406, 3.406], [7.813, 7.813], [0.800, 1.538]]), "D": 0.2, "epsilon": 0.2}]
    return "M200i", robot, obs, s


def oracle_problem(O, ROBOT, obs, s, solver=0, grad=0, margin=None, max_input="default"):
    if margin is None:
        margin = [o["epsilon"] if solver == 0 else o["D"] for o in obs]
    mi = s["MAX_input"] if max_input == "default" else max_input
    return O.Problem(O.robot(ROBOT), s["H"], [o["l"] for o in obs], margin, s["QQ"], s.get("lim"),
                     mi if solver == 0 else None, s["epsilon_O"], s["MAX_O_ITER"], solver=solver, grad=grad,
                     alpha=s.get("alpha", 0.0))


def oracle_feasible_fn(O, ROBOT, obs):
    r = O.robot(ROBOT)
    o6 = [O.obs6(o["l"]) for o in obs]

    def fn(cand):
        return np.array([all(O.dist_arm(r, th, o)[0] >= ob["D"] for o, ob in zip(o6, obs)) for th in cand])
    return fn


def batch_m16ib(O, B, horizon=50, seed=synthetic.SEED):
    return synthetic.batch_config_m16ib(B, oracle_feasible_fn(O, "M16iB", [synthetic.OBS_M16IB]), horizon, seed)


def sampling_box(rng, N):
    return synthetic.SAMPLE_OFF + (rng.random((N, 5)) - 0.5) * 2 * synthetic.REGION_S
