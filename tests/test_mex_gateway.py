"""The MEX gateway matlab/cfs_mex.cpp, compiled against a stand-in mex.h (tests/mex_stub/) and called with the arguments the
drop-in classdefs matlab/CFS_FANUC.m / PSGCFS_FANUC.m pass: MATLAB is not installed here, so this is how the gateway's
marshalling (robot / obs / sys_info structs, batched columns, outputs) gets compiled and executed at all."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib
from tests import common

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "mex_stub")], stdout=subprocess.DEVNULL)
    h = C.CDLL(os.path.join(HERE, "mex_stub", "libmexharness.so"))
    h.mexh_last_error.restype = C.c_char_p
    h.mexh_cfs.restype = C.c_int
    yield h
    h.mexh_shutdown()


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def call_mex(h, solver, grad, ROBOT, robot, obs, s, x0, ff, caug, xref, noise=None, drop=()):
    """mirrors cfs_mex(solver, grad, ROBOT, obs, sys_info [, noise]); x0 (B,2nj) ff (B,n) caug (B,) xref (B,2n) noise (B,K,n)"""
    nj, H = int(s["njoint"]), int(s["H"])
    n, B, K = nj * H, x0.shape[0], int(s["MAX_O_ITER"])
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    DH = np.asfortranarray(robot["DH"], dtype=np.float64)
    cap = np.asfortranarray(np.stack([np.asarray(robot["cap"][i]["p"], dtype=np.float64)[:, :2] for i in range(nj)], axis=2))
    T = None if robot.get("T") is None else np.asfortranarray(robot["T"], dtype=np.float64)
    seg = np.asfortranarray(np.stack([np.asarray(o["l"], dtype=np.float64) for o in obs], axis=2))
    D, eps = f([o["D"] for o in obs]), f([o["epsilon"] for o in obs])
    lim = None if (s.get("lim") is None or "lim" in drop) else f(s["lim"])
    mi = None if (s.get("MAX_input") is None or "MAX_input" in drop) else f(s["MAX_input"])
    out = dict(u=np.zeros((B, n)), x=np.zeros((B, 2 * n)), cost_hist=np.zeros((B, K)), e_u_hist=np.zeros((B, K)),
               iters=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32))
    qp = C.c_double(0.0)
    keep = [DH, cap, T, seg, D, eps, lim, mi, f(s["QQ"].T), f(x0), f(ff), f(caug), f(xref), None if noise is None else f(noise),
            f(np.asarray(robot["base"]).reshape(-1))]
    rc = h.mexh_cfs(solver.encode(), grad.encode(), ROBOT.encode(), C.c_int(nj), C.c_int(H), C.c_int(B), C.c_int(DH.shape[0]),
                    _dp(DH), _dp(keep[14]), _dp(cap), _dp(T), C.c_double(robot["delta_t"]), C.c_int(len(obs)), _dp(seg), _dp(D),
                    _dp(eps), _dp(keep[8]), _dp(lim), _dp(mi), _dp(keep[9]), _dp(keep[10]), _dp(keep[11]), _dp(keep[12]),
                    C.c_double(s["epsilon_O"]), C.c_int(K), C.c_double(s.get("alpha", 0.0)), _dp(keep[13]), _dp(out["u"]),
                    _dp(out["x"]), _dp(out["cost_hist"]), _dp(out["e_u_hist"]), _dp(out["iters"]), _dp(out["status"]), C.byref(qp))
    if rc:
        raise RuntimeError(h.mexh_last_error().decode())
    out["qp_steps"] = qp.value
    return out


def test_gateway_compiles_and_fails_loudly_without_a_gpu(harness):
    """not gpu: the gateway builds against the stub, links libcfs_b200.so, and without a usable device raises the MATLAB
    error id the classdefs document (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    ROBOT, robot, obs, s = common.main_fanuc_config()
    with pytest.raises(RuntimeError, match="cfs:cuda"):
        call_mex(harness, "CFS", "num_jac", ROBOT, robot, obs, s, s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]),
                 s["x_"][None])


@pytest.mark.gpu
def test_gateway_matches_the_ctypes_path(harness, ctx, oracle):
    # main_FANUC.m's configuration (M200i, H = 30, one problem), as CFS_FANUC.optimizer passes it
    ROBOT, robot, obs, s = common.main_fanuc_config()
    args = (s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]), s["x_"][None])
    r = dict(robot)
    r["name"] = ROBOT
    ctx.set_robot(r, 5)
    ctx.set_obstacles(obs)
    ctx.set_cost(s["H"], s["QQ"], s["lim"], s["MAX_input"])
    ref = ctx.solve_batch(*args, s["epsilon_O"], s["MAX_O_ITER"])
    out = call_mex(harness, "CFS", "num_jac", ROBOT, robot, obs, s, *args)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(out[k], ref[k]), k
    it = int(ref["iters"][0])
    assert np.array_equal(out["cost_hist"][0, :it], ref["cost_hist"][0, :it]) and out["qp_steps"] > 0
    # a batch (sys_info.xR / ff / caug / x_ with one column per problem), num_jac and derivest, M16iB
    cfg = common.batch_m16ib(oracle, 12, horizon=20)
    sb = dict(cfg["sys_info"])
    rb = dict(cfg["robot"])
    rb["name"] = "M16iB"
    ctx.set_robot(rb, 5)
    ctx.set_obstacles(cfg["obs"])
    ctx.set_cost(sb["H"], sb["QQ"], sb["lim"], sb["MAX_input"])
    bargs = (cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"])
    for grad, gm in (("num_jac", _lib.GRAD_NUMJAC), ("derivest", _lib.GRAD_DERIVEST)):
        ref = ctx.solve_batch(*bargs, sb["epsilon_O"], sb["MAX_O_ITER"], grad=gm)
        out = call_mex(harness, "CFS", grad, "M16iB", cfg["robot"], cfg["obs"], sb, *bargs)
        for k in ("u", "x", "iters", "status"):
            assert np.array_equal(out[k], ref[k]), (grad, k)
    # PSGCFS with the host-drawn noise (PSGCFS_FANUC.m:109), no bounds in the projection (:120)
    sb["alpha"] = 1.0 / np.linalg.svd(sb["QQ"], compute_uv=False).max()
    sb["MAX_O_ITER"] = 5
    noise = np.random.default_rng(4).normal(0.0, 0.1, size=(12, 5, sb["H"] * 5))
    ctx.set_cost(sb["H"], sb["QQ"], sb["lim"], None)
    ref = ctx.solve_batch(*bargs, sb["epsilon_O"], 5, solver=_lib.SOLVER_PSGCFS, noise=noise, alpha=sb["alpha"])
    out = call_mex(harness, "PSGCFS", "num_jac", "M16iB", cfg["robot"], cfg["obs"], sb, *bargs, noise=noise)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(out[k], ref[k]), ("psgcfs", k)
    # the script path of M16iB/main_CFS.m has no velocity rows: sys_info without .lim
    out = call_mex(harness, "CFS", "num_jac", "M16iB", cfg["robot"], cfg["obs"], cfg["sys_info"], *bargs, drop=("lim",))
    ctx.set_cost(sb["H"], sb["QQ"], None, cfg["sys_info"]["MAX_input"])
    ref = ctx.solve_batch(*bargs, sb["epsilon_O"], cfg["sys_info"]["MAX_O_ITER"])
    assert np.array_equal(out["u"], ref["u"]) and np.array_equal(out["status"], ref["status"])
