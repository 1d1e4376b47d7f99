#!/usr/bin/env python
"""Write the golden CONFIGURATIONS as flat CFSB fixtures (motionplanning_5d_m_b200/fixture_io.py) that a MATLAB owner feeds to the
UNMODIFIED reference classes with matlab/make_reference_golden.m:

    python tests/golden/export_fixtures.py            # -> tests/golden/fixtures/<case>.bin   (inputs only)
    matlab -batch "addpath('<repo>/matlab'); make_reference_golden('<reference root>', '<repo>/tests/golden')"
                                                      # -> tests/golden/reference/<case>_ref.bin (u, x_, cost_all, iter_O)
    python -m pytest tests/test_reference_golden.py   # oracle (CPU) and CUDA path (-m gpu) against the MATLAB outputs

Each fixture holds exactly what the mains put into sys_info / obs (main_FANUC.m:56-60,106-127): robot id (0 M16iB, 1 M200i, 2 2L),
solver id (0 CFS, 1 PSGCFS), H, njoint, QQ, ff, caug, Aaug, Baug, x0 = xR(:,1), x_, lim, MAX_input, epsilon_O, MAX_O_ITER, alpha,
obs_l (3 x 2 x O), obs_D, obs_epsilon and, for PSGCFS, noise (nn x MAX_O_ITER: the normrnd draws in call order)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
ROBOT_ID = {"M16iB": 0, "M200i": 1, "2L": 2}


def case_arrays(ROBOT, obs, s, solver=0, noise=None):
    a = dict(robot_id=ROBOT_ID[ROBOT], solver_id=solver, H=s["H"], njoint=s["njoint"], QQ=s["QQ"], ff=s["ff"], caug=s["caug"],
             Aaug=s["Aaug"], Baug=s["Baug"], x0=s["xR"][:, 0], x_=s["x_"], epsilon_O=s["epsilon_O"], MAX_O_ITER=s["MAX_O_ITER"],
             alpha=s.get("alpha", 0.0), obs_l=np.stack([np.asarray(o["l"], dtype=np.float64) for o in obs], axis=2),
             obs_D=[o["D"] for o in obs], obs_epsilon=[o["epsilon"] for o in obs], has_lim=0 if s.get("lim") is None else 1)
    if s.get("lim") is not None:
        a["lim"] = s["lim"]
    a["MAX_input"] = s["MAX_input"]
    if noise is not None:
        a["noise"] = np.asarray(noise, dtype=np.float64).T          # nn x MAX_O_ITER
    return a


def cases():
    from tests import common
    g = common.golden("inputs.npz")
    out = {}
    ROBOT, robot, obs, s = common.main_fanuc_config()
    out["main_fanuc_cfs"] = case_arrays(ROBOT, obs, s)
    noise = np.random.default_rng(123).normal(0.0, 0.1, size=(s["MAX_O_ITER"], s["H"] * 5))
    out["main_fanuc_psgcfs"] = case_arrays(ROBOT, obs, s, solver=1, noise=noise)
    ROBOT, robot, obs, s = common.main_2l_config()
    out["main_2l_cfs"] = case_arrays(ROBOT, obs, s)
    ROBOT, robot, obs, s = common.rrtstar_route_config(g["route_wp"])
    out["rrtstar_cfs"] = case_arrays(ROBOT, obs, s)
    return out


def main():
    from motionplanning_5d_m_b200 import fixture_io
    for name, arrays in cases().items():
        path = os.path.join(HERE, "fixtures", name + ".bin")
        fixture_io.write_fixture(path, arrays)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
