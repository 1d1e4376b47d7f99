"""Synthetic workloads named by BASELINE.json: random start/goal pairs for the batched CFS run.

batch_config_m16ib(): M16iB capsules, one obstacle segment [[3.906;8.313;0.001],[3.906;8.313;1.938]] with
D = epsilon = 0.2 (M16iB/main_CFS.m:57,103), horizon 50, cost exactly as main_FANUC.m:64-127; theta_0, theta_g uniform in
sample_off +- region_s (M16iB/RRT_FANUC_test.m:47-48), rejection-sampled until both ends satisfy dist_arm >= D
(the RRT feasibility rule, RRT_FANUC.m:172).  The feasibility test is a callback so that the bench uses the GPU
kernel (cfs_nodes_feasible) and the CPU tests use the oracle; both see the same Philox stream.
"""
import numpy as np

from . import problem
from .robot import robotproperty2

SEED = 20261018
OBS_M16IB = {"l": np.array([[3.906, 3.906], [8.313, 8.313], [0.001, 1.938]]), "D": 0.2, "epsilon": 0.2}
SAMPLE_OFF = np.array([0, np.pi / 2, 0, 0, 0])
REGION_S = np.array([np.pi / 2, np.pi / 2, np.pi / 2, np.pi / 1.5, np.pi / 1.5])


def sample_feasible(count, feasible_fn, rng, chunk=None, sample_off=None):
    """Draw `count` configurations uniformly in the sampling box that pass feasible_fn((M,5)) -> bool (M,)."""
    out = np.zeros((0, 5))
    chunk = chunk or max(64, int(count * 1.5))
    off = SAMPLE_OFF if sample_off is None else sample_off
    while out.shape[0] < count:
        cand = off + (rng.random((chunk, 5)) - 0.5) * 2 * REGION_S
        ok = np.asarray(feasible_fn(cand), dtype=bool)
        out = np.concatenate([out, cand[ok]], axis=0)
    return out[:count]


OBS_M200I = {"l": np.array([[3.806, 3.606], [8.413, 8.413], [0.001, 1.038]]), "D": 0.2, "epsilon": 0.25}  # main_FANUC.m:56-60
SAMPLE_OFF_M200I = np.zeros(5)                                                                              # RRTstar_CFS.m:54


def batch_config_m16ib(B, feasible_fn, horizon=50, seed=SEED):
    return batch_config("M16iB", B, feasible_fn, [dict(OBS_M16IB)], horizon, seed, SAMPLE_OFF)


def batch_config_m200i_psgcfs(B, feasible_fn, horizon=30, seed=SEED, max_outer=20):
    """BASELINE.json configs[3]: PSGCFS on the LR Mate 200iD (main_FANUC.m's robot, obstacle, horizon and weights; SOLVER =
    'PSGCFS', main_FANUC.m:22-25) for B random start/goal pairs, with the host-drawn normrnd(0,0.1,[nn,1]) of every outer
    iteration (PSGCFS_FANUC.m:109) as `noise` (B, MAX_O_ITER, n) and alpha = 1/max(svd(QQ)) (main_FANUC.m:120)."""
    cfg = batch_config("M200i", B, feasible_fn, [dict(OBS_M200I)], horizon, seed, SAMPLE_OFF_M200I)
    s = cfg["sys_info"]
    s["alpha"] = 1.0 / np.linalg.svd(s["QQ"], compute_uv=False).max()
    s["MAX_O_ITER"] = max_outer
    rng = np.random.Generator(np.random.Philox(seed + 7))
    cfg["noise"] = rng.normal(0.0, 0.1, size=(B, max_outer, horizon * 5))
    return cfg


def batch_config(ROBOT, B, feasible_fn, obs, horizon, seed, sample_off):
    robot = robotproperty2(ROBOT)
    nj = 5
    rng = np.random.Generator(np.random.Philox(seed))
    ends = sample_feasible(2 * B, feasible_fn, rng, chunk=max(64, 3 * B), sample_off=sample_off)
    th0, thg = ends[0::2], ends[1::2]
    Aaug, Baug, Qaug, QQ = problem.build_cost_matrices(robot, nj, horizon, problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0)
    x0 = np.concatenate([th0, np.zeros((B, nj))], axis=1)
    gaug = np.tile(np.concatenate([thg, np.zeros((B, nj))], axis=1), (1, horizon))
    ff, caug = problem.build_linear_term(Aaug, Baug, Qaug, x0, gaug)
    xref = problem.straight_line_reference(th0, thg, horizon)
    sys_info = dict(robot=robot, H=horizon, njoint=nj, nstate=2 * nj, nu=nj, QQ=QQ, Qaug=QQ, Aaug=Aaug, Baug=Baug,
                    lim=np.ones(nj), MAX_input=np.tile(np.array([1, 1, np.pi, np.pi, np.pi]) * robot["delta_t"], horizon),
                    epsilon_O=1e-1, MAX_O_ITER=20, alpha=0.0)
    return dict(robot=robot, ROBOT=ROBOT, obs=obs, sys_info=sys_info, x0=np.ascontiguousarray(x0),
                ff=np.ascontiguousarray(ff), caug=np.ascontiguousarray(caug), xref=np.ascontiguousarray(xref),
                theta0=th0, thetag=thg)
