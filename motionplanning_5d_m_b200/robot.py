"""robotproperty2 -- host-side mirror of Lib/functions/robotproperty2.m (robot constants only).

Returns a dict with the fields the reference's solvers read: name, nlink, delta_t, DH (6x4 [theta d a alpha]),
base (3,), cap (list of {'p': 3x2, 'r': radius}), A, B (double-integrator, robotproperty2.m:136-139) and, for '2L',
T (robot.T).  The literal DH constants 1.5708 / 3.1416 are kept (they are NOT pi/2, pi).
"""
import numpy as np


def robotproperty2(name):
    r = {"name": name, "delta_t": 0.5}
    if name == "M200i":  # robotproperty2.m:12-55
        r["nlink"] = 6
        r["DH"] = np.array([[0, 0, 0.050, -1.5708],
                            [-1.5708, 0, 0.440, 3.1416],
                            [0, 0, 0.035, -1.5708],
                            [0, -0.420, 0, 1.5708],
                            [0, 0, 0, -1.5708],
                            [0, -0.080, 0, 3.1416]], dtype=np.float64)
        caps = [([[0, 0], [0, 0], [0, 0]], 0.0),
                ([[-0.4, 0], [0, 0], [0, 0]], 0.13),
                ([[-0.03, -0.03], [0, 0], [0.05, 0.05]], 0.0),
                ([[0, 0], [0, 0.4], [0, 0]], 0.068),
                ([[0, 0], [0, 0], [-0.26, 0.01]], 0.01),
                ([[0.05, 0.18], [0, 0], [0.1107, 0.1107]], 0.06)]
        r["base"] = np.array([3150, 8500, 330], dtype=np.float64) / 1000
    elif name == "M16iB":  # robotproperty2.m:58-99
        r["nlink"] = 6
        r["DH"] = np.array([[0.5, 0.65, 0.15, 1.5708],
                            [1.5708, 0, 0.77, 0],
                            [0, 0, 0.1, 1.5708],
                            [0, 0.74, 0, -1.5708],
                            [-np.pi / 2, 0, 0, 1.5708],
                            [np.pi, 0.1, 0, 0]], dtype=np.float64)
        caps = [([[0, 0], [0, 0], [-0.1, 0.1]], 0.15),
                ([[-0.75, 0], [0, 0], [-0.15, -0.15]], 0.13),
                ([[-0.03, -0.03], [0, 0], [0.05, 0.05]], 0.22),
                ([[0, 0], [0, 0.55], [0, 0]], 0.11),
                ([[0, 0], [0, 0], [-0.05, 0.110]], 0.07),
                ([[-0.11, -0.11], [0, 0], [0.09, 0.09]], 0.11)]
        r["base"] = np.array([3250, 8500, 0], dtype=np.float64) / 1000
    elif name == "2L":  # robotproperty2.m:102-130
        r["nlink"] = 3
        r["DH"] = np.array([[0, 0, 0.3, 0], [0, 0, 0.2, 0], [0, 0, 0, 0]], dtype=np.float64)
        r["T"] = np.array([[0, 0, 0.3], [0, 0, 0], [0, 0, 0.0]], dtype=np.float64)
        caps = [([[0, 0.3], [0, 0], [0, 0]], 0.05), ([[0, 0.2], [0, 0], [0, 0]], 0.05)]
        r["base"] = np.zeros(3)
    else:
        raise ValueError("unknown robot %r (expected 'M200i', 'M16iB' or '2L')" % (name,))
    r["cap"] = [{"p": np.array(p, dtype=np.float64), "r": rad} for p, rad in caps]
    nl, dt = r["nlink"], r["delta_t"]
    I = np.eye(nl)
    r["A"] = np.block([[I, dt * I], [np.zeros((nl, nl)), I]])  # robotproperty2.m:136-137
    r["B"] = np.vstack([0.5 * dt * dt * I, dt * I])            # robotproperty2.m:138-139
    return r
