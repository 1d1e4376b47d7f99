"""-m "not gpu": pins the CPU oracle against the reference's own known answers and an independent numpy restatement."""
import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
from tests import common, np_restatement as NP


def test_distlinseg_doc_example(oracle):
    # Lib/functions/distLinSeg.m:15-18
    d, pts = oracle.dist_lin_seg([0, 0], [1, 1], [1, 0], [2, 0])
    assert abs(d - 0.7071) < 5e-5 and abs(d - np.sqrt(0.5)) < 1e-15
    assert np.allclose(pts, [[0.5, 0.5], [1, 0]], atol=0)


def test_distlinseg_branches(oracle):
    rng = np.random.default_rng(0)
    cases = []
    for _ in range(300):
        a, b, c, d = rng.normal(size=(4, 3))
        cases.append((a, b, c, d))
        cases.append((a, a, c, d))            # first segment is a point (distLinSeg.m:45-50)
        cases.append((a, b, c, c))            # second is a point (:39-44)
        cases.append((a, a, c, c))            # both points (:51-53)
        cases.append((a, b, c, c + 2.0 * (b - a)))  # parallel (:55-66)
    for a, b, c, d in cases:
        do, po = oracle.dist_lin_seg(a, b, c, d)
        dn, pn = NP.dist_lin_seg(a, b, c, d)
        assert abs(do - dn) < 1e-14
        assert np.abs(po.reshape(-1) - pn).max() < 1e-13


def test_derivest_known_answers(oracle):
    # derivest.m:163-174 : derivest(@exp,1) = 2.71828182845904 (15 digits shown)
    d, e, _ = oracle.derivest_named(0, 1.0)
    assert abs(d - 2.71828182845904) < 5e-14 and e < 1e-12
    # demo/derivest_demo.m:13 exp at 0 -> 1 ; :71 sinh central at 0 -> 1, err 1.0412e-15 ; :82 log at 1e-3 -> 1000
    assert abs(oracle.derivest_named(0, 0.0)[0] - 1.0) < 1e-13
    d, e, _ = oracle.derivest_named(2, 0.0)
    assert abs(d - 1.0) < 1e-14 and abs(e - 1.0412e-15) < 1e-19
    assert abs(oracle.derivest_named(3, 1e-3)[0] - 1000.0) < 1e-6
    # demo/derivest_demo.m:31 : sin at linspace(0,2*pi,13) -> cos
    for x in np.linspace(0, 2 * np.pi, 13):
        assert abs(oracle.derivest_named(1, x)[0] - np.cos(x)) < 1e-12


def test_derivest_vs_numpy_restatement(oracle):
    funs = {0: np.exp, 1: np.sin, 2: np.sinh}  # (log leaves the real line for x0-h*delta<0: MATLAB goes complex)
    rng = np.random.default_rng(1)
    for which, f in funs.items():
        for x in rng.uniform(0.05, 3.0, size=20):
            do, eo, fo = oracle.derivest_named(which, x)
            dn, en, fn = NP.derivest(lambda t: float(f(t)), x)
            # the error estimates are rounding noise, so the two implementations may select neighbouring steps:
            # the derivative itself agrees to ~1e-13 relative
            assert abs(do - dn) <= 1e-11 * max(1, abs(dn))


@pytest.mark.parametrize("ROBOT", ["M16iB", "M200i"])
def test_fk_dist_numjac_vs_numpy_restatement(oracle, ROBOT):
    rng = np.random.default_rng(2)
    robot = M.robotproperty2(ROBOT)
    r = oracle.robot(ROBOT)
    obs = np.array([[3.906, 3.906], [8.313, 8.313], [0.001, 1.938]])
    o6 = oracle.obs6(obs)
    for th in common.sampling_box(rng, 200):
        DH = robot["DH"][:5].copy()
        DH[:, 0] = th
        if ROBOT == "M200i":
            DH[1, 0] -= np.pi / 2
        pn = NP.cap_pos(robot["base"], DH, [c["p"] for c in robot["cap"]])
        po = oracle.cap_pos(r, th)
        for i in range(5):
            assert np.abs(po[i].T - pn[i]).max() < 1e-14
        dn, ln = NP.dist_arm(th, robot, obs, ROBOT)
        do, lo, _ = oracle.dist_arm(r, th, o6)
        assert abs(do - dn) < 1e-14 and lo == ln
        gn = NP.num_jac(lambda t: NP.dist_arm(t, robot, obs, ROBOT)[0], th)
        go = oracle.num_jac(r, th, o6)
        assert np.abs(go - gn).max() < 1e-9


def test_derivest_gradient_vs_numpy_restatement(oracle):
    rng = np.random.default_rng(3)
    robot = M.robotproperty2("M16iB")
    r = oracle.robot("M16iB")
    obs = np.array([[3.906, 3.906], [8.313, 8.313], [0.001, 1.938]])
    o6 = oracle.obs6(obs)

    def dist_link(th, lid):
        DH = robot["DH"][:5].copy()
        DH[:, 0] = th
        pos = NP.cap_pos(robot["base"], DH, [c["p"] for c in robot["cap"]])
        dis, pts = NP.dist_lin_seg(pos[lid - 1][:, 0], pos[lid - 1][:, 1], obs[:, 0], obs[:, 1])
        return dis if abs(dis) >= 1e-4 else -np.linalg.norm(pts[:3] - pos[lid - 1][:, 1])

    for th in common.sampling_box(rng, 12):
        _, lid, _ = oracle.dist_arm(r, th, o6)
        go = oracle.derivest_grad(r, th, o6, lid)
        for s in range(5):
            f = lambda x: dist_link(np.concatenate([th[:s], [x], th[s + 1:]]), lid)
            gn = NP.derivest(f, th[s])[0]
            assert abs(go[s] - gn) < 1e-9 * max(1.0, abs(gn)), (s, go[s], gn)


def test_qp_kkt_random(oracle):
    rng = np.random.default_rng(4)
    for trial in range(25):
        n, m = int(rng.integers(3, 30)), int(rng.integers(1, 60))
        Mx = rng.normal(size=(n, n))
        G = Mx @ Mx.T + 0.1 * np.eye(n)
        a = rng.normal(size=n) * 3
        Cm = rng.normal(size=(m, n))
        xf = rng.normal(size=n)
        d = Cm @ xf + rng.uniform(0.0, 1.0, size=m)  # xf strictly feasible
        x, lam, rc, it, kkt = oracle.qp_gi(G, a, Cm, d)
        assert rc == 0 and kkt < 1e-8 * (1 + np.abs(a).max()), (trial, rc, kkt)


def test_qp_infeasible_detected(oracle):
    G = np.eye(3)
    a = np.zeros(3)
    Cm = np.array([[1.0, 0, 0], [-1.0, 0, 0]])
    d = np.array([-1.0, -1.0])  # x1 <= -1 and x1 >= 1
    assert oracle.qp_gi(G, a, Cm, d)[2] == 2


def test_main_fanuc_oracle_golden(oracle):
    """Frozen oracle golden for main_FANUC.m's configuration (SURVEY.md section 6 scratch: 10 iterations, cost
    1.22532e5, clearance = epsilon) and the sanity pin against the reference's own logged RRT*-CFS costs."""
    ROBOT, robot, obs, s = common.main_fanuc_config()
    P = common.oracle_problem(oracle, ROBOT, obs, s)
    ref = P.solve_batch(s["xR"][:, 0][None], s["ff"][None], np.array([s["caug"]]), s["x_"][None])
    it = int(ref["iters"][0])
    assert it == 10 and (ref["status"][0] & 0xFF) == 0
    assert abs(ref["cost_hist"][0, it - 1] - 122532.0927) < 1e-3
    r = oracle.robot(ROBOT)
    x = ref["x"][0].reshape(30, 10)
    dmin = min(oracle.dist_arm(r, x[i, :5], oracle.obs6(obs[0]["l"]))[0] for i in range(30))
    assert abs(dmin - 0.25) < 1e-4
    # re-linearised at the converged trajectory the constraints hold up to the clearance error seen above
    A, b, *_ = P.get_con(s["xR"][:, 0], x.reshape(-1), ref["u"][0])
    assert (A @ ref["u"][0] - b).max() < 1e-4
