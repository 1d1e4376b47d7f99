#!/usr/bin/env python
"""Golden case for box obstacles: the boxes derived from the reference's own STL map
(map/assembly line_Assem1.STL through MapFromSTL.m's transform, mm -> m: motionplanning_5d_m_b200/stl_boxes.py) near the M16iB,
and the oracle's distances / gradients / CFS solves against them.  Run in the build container (needs /root/reference):

    python tests/golden/make_box_golden.py      ->  tests/golden/stl_boxes.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
STL = "/root/reference/map/assembly line_Assem1.STL"
# start/goal pairs (found by search: the straight line passes within the margin of a box, the problem stays feasible) + one whose
# line passes through a box (first QP infeasible)
PAIRS = [([1.40478134, 3.0213851, 0.45941676, -0.92669453, 0.88538058], [-0.88979491, 1.01185889, 0.13111864, -0.41408637, -0.62402978]),
         ([-1.4485871, 2.4133454, -0.81677541, -0.70304219, 0.33231862], [1.56727673, 0.94146484, -0.0988879, 0.53386501, 0.76810843]),
         ([-0.55949604, 0.71614054, -0.9438738, -0.11634212, -0.66077556], [1.38298341, 1.77778794, 1.25703336, 1.77872228, 1.67607105])]


def main():
    import motionplanning_5d_m_b200 as M
    import oracle as O
    from motionplanning_5d_m_b200 import stl_boxes, synthetic
    from tests import common
    O.build()
    obs = stl_boxes.boxes_from_stl(STL, max_boxes=512, near=[3.25, 8.5, 0.8], radius=1.6, max_keep=12, max_size=1.2, D=0.1, epsilon=0.1)
    out = {"box_l": np.stack([o["l"] for o in obs], axis=2), "box_D": np.array([o["D"] for o in obs]),
           "box_epsilon": np.array([o["epsilon"] for o in obs])}
    r = O.robot("M16iB")
    rng = np.random.default_rng(20261018)
    th = synthetic.SAMPLE_OFF + (rng.random((256, 5)) - 0.5) * 2 * synthetic.REGION_S
    out["theta"] = th
    out["dist"] = np.array([[O.dist_arm(r, t, O.obs6(o))[0] for o in obs] for t in th])
    out["linkid"] = np.array([[O.dist_arm(r, t, O.obs6(o))[1] for o in obs] for t in th])
    out["grad"] = np.array([[O.num_jac(r, t, O.obs6(o)) for o in obs] for t in th])
    robot = M.robotproperty2("M16iB")
    H = 30
    for k, (t0, tg) in enumerate(PAIRS):
        s = M.make_sys_info(robot, 5, H, t0, tg)
        P = common.oracle_problem(O, "M16iB", obs, s)
        res = P.solve_batch(s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]), s["x_"][None])
        for key in ("u", "x", "cost_hist", "iters", "status"):
            out["solve%d.%s" % (k, key)] = res[key][0]
        out["solve%d.theta0" % k], out["solve%d.thetag" % k] = np.array(t0), np.array(tg)
        print("pair", k, "status", res["status"][0], "iters", res["iters"][0])
    np.savez_compressed(os.path.join(HERE, "stl_boxes.npz"), **out)
    print("boxes", out["box_l"].shape, "touching configurations", int((out["dist"] < 0).sum()), "of", out["dist"].size)


if __name__ == "__main__":
    main()
