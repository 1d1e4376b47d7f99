#!/usr/bin/env python
"""Freeze golden vectors for the parity tests.  Run in the build container (needs /root/reference for the .mat inputs):

    python tests/golden/make_golden.py

What is frozen, and from where:
  inputs.npz      data/good_xori.mat:xuori (250), data/M16_ref_2.mat:{xref,uref}, data/200i_xori.mat:route_wp (5x16) --
                  the reference's own INPUT fixtures, converted so that the GPU box (no /root/reference) can read them.
  kat.json        the reference's own known answers: distLinSeg.m:15-18 doc example; derivest.m:163-174 and
                  DERIVESTsuite/demo/derivest_demo.m:13,31,71,82 (values typed from those files, not computed).
  cases.npz       outputs of the CPU oracle (oracle/cfs_oracle.c) on the reference's shipped configurations.
                  MATLAB/Octave are not available offline, so these are ORACLE goldens: they pin the oracle against
                  regressions and give the CUDA path a fixed target; they are not MATLAB outputs (parity unpinned for
                  quadprog, see oracle/cfs_oracle.c header and DESIGN.md).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def solve(O, common, ROBOT, obs, s, solver=0, grad=0, noise=None):
    P = common.oracle_problem(O, ROBOT, obs, s, solver=solver, grad=grad)
    args = (s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]), s["x_"][None])
    r = P.solve_batch(*args, noise=noise)
    return {k: r[k][0] for k in ("u", "x", "cost_hist", "e_u_hist", "iters", "status")}


def main():
    import scipy.io as sio

    import oracle as O
    from tests import common
    O.build()
    xuori = sio.loadmat(os.path.join(REF, "data/good_xori.mat"))["xuori"].reshape(-1)
    m16 = sio.loadmat(os.path.join(REF, "data/M16_ref_2.mat"))
    route = sio.loadmat(os.path.join(REF, "data/200i_xori.mat"))["route_wp"]
    np.savez_compressed(os.path.join(HERE, "inputs.npz"), xuori=xuori, xref=m16["xref"].reshape(-1),
                        uref=m16["uref"].reshape(-1), route_wp=route)
    kat = {
        "distLinSeg": {"cite": "Lib/functions/distLinSeg.m:15-18", "p1": [0, 0], "p2": [1, 1], "p3": [1, 0], "p4": [2, 0],
                       "dist": 0.7071, "points": [[0.5, 0.5], [1, 0]], "digits": 4},
        "derivest": [
            {"cite": "DERIVESTsuite/DERIVESTsuite/derivest.m:163-174", "fun": "exp", "x0": 1.0, "der": 2.71828182845904,
             "digits": 14},
            {"cite": "demo/derivest_demo.m:13", "fun": "exp", "x0": 0.0, "der": 1.0, "digits": 13},
            {"cite": "demo/derivest_demo.m:71 + demo/html/derivest_demo.html", "fun": "sinh", "x0": 0.0, "der": 1.0,
             "errest": 1.0412e-15, "digits": 14},
            {"cite": "demo/derivest_demo.m:82", "fun": "log", "x0": 1e-3, "der": 1000.0, "digits": 9},
        ],
    }
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    out = {}

    def put(name, d):
        for k, v in d.items():
            out["%s.%s" % (name, k)] = np.asarray(v)

    # main_FANUC.m (M200i, H=30): CFS and PSGCFS (seeded normrnd stand-in)
    ROBOT, robot, obs, s = common.main_fanuc_config()
    put("main_fanuc_cfs", solve(O, common, ROBOT, obs, s))
    noise = np.random.default_rng(123).normal(0.0, 0.1, size=(1, s["MAX_O_ITER"], s["H"] * 5))
    put("main_fanuc_psgcfs", solve(O, common, ROBOT, obs, s, solver=1, noise=noise))
    # main_2L.m
    ROBOT, robot, obs, s = common.main_2l_config()
    put("main_2l_cfs", solve(O, common, ROBOT, obs, s))
    # M16iB/main_CFS.m as shipped (:57 obstacle -> first QP infeasible) and with the obstacle of its line :53
    ROBOT, robot, obs, s = common.main_cfs_m16ib_config(xuori)
    put("m16ib_script_derivest", solve(O, common, ROBOT, obs, s, grad=1))
    obs[0]["l"] = np.array(common.OBS_M16_SCRIPT_ALT)
    put("m16ib_script_alt_derivest", solve(O, common, ROBOT, obs, s, grad=1))
    put("m16ib_script_alt_numjac", solve(O, common, ROBOT, obs, s, grad=0))
    # RRTstar_CFS.m CFS stage on the shipped route data/200i_xori.mat
    ROBOT, robot, obs, s = common.rrtstar_route_config(route)
    put("rrtstar_cfs", solve(O, common, ROBOT, obs, s))
    # seeded random batch at the headline configuration (first 32 problems)
    cfg = common.batch_m16ib(O, 32)
    s = cfg["sys_info"]
    P = common.oracle_problem(O, "M16iB", cfg["obs"], s)
    r = P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=8)
    put("batch_m16ib_32", {k: r[k] for k in ("u", "x", "cost_hist", "iters", "status")})
    out["batch_m16ib_32.theta0"] = cfg["theta0"]
    out["batch_m16ib_32.thetag"] = cfg["thetag"]
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **out)
    for k in sorted(out):
        if k.endswith("iters") or k.endswith("status"):
            print(k, out[k])
    print("wrote", os.path.getsize(os.path.join(HERE, "cases.npz")), "bytes")


if __name__ == "__main__":
    main()
