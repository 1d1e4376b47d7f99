"""Host-side problem set-up: mirrors of the cost construction in the reference's main scripts.

build_cost_matrices  <- main_FANUC.m:64-97  (Aaug, Baug, Qaug, R -> QQ);  RRTstar_CFS.m:124-160; main_2L.m:69-92
build_linear_term    <- main_FANUC.m:98-103 (gaug, ff, caug)
straight_line_reference <- main_FANUC.m:38-49
make_sys_info        <- main_FANUC.m:106-127 (the struct handed to CFS_FANUC / PSGCFS_FANUC)
cubicpolytraj        <- RRTstar_CFS.m:96-100 (Robotics System Toolbox call; default zero waypoint velocities)
"""
import numpy as np

Q_MAIN_FANUC = np.block([[np.diag([10, 10, 1, 1, 1.0]), 0.1 * np.eye(5)], [0.1 * np.eye(5), np.diag([10, 10, 1, 1, 1.0])]])
Q_RRTSTAR = np.block([[np.diag([10, 10, 1, 1, 1.0]), 0.1 * np.eye(5)], [0.1 * np.eye(5), np.diag([100, 20, 1, 1, 1.0])]])
R_MAIN_FANUC = np.array([[10, 0, 0, 0, 0], [0, 10, 1, 0, 0], [0, 1, 2, 0, 0], [0, 0, 0, 2, 0], [0, 0, 0, 0, 1.0]])
Q_M16_SCRIPT = np.block([[np.diag([10, 1, 1, 1, 1.0]), 0.1 * np.eye(5)], [0.1 * np.eye(5), np.diag([10, 1, 1, 1, 1.0])]])  # M16iB/main_CFS.m:69-80
Q_2L = np.block([[np.diag([10, 1.0]), 0.1 * np.eye(2)], [0.1 * np.eye(2), np.diag([10, 1.0])]])
R_2L = np.array([[5, 0], [0, 4.0]])


def joint_dynamics(robot, njoint):
    """robot.A([1:nj, nlink+1:nlink+nj],...) , robot.B(same rows, 1:nj)   (CFS_FANUC.m:91)"""
    nl = robot["nlink"]
    idx = list(range(njoint)) + list(range(nl, nl + njoint))
    return robot["A"][np.ix_(idx, idx)], robot["B"][np.ix_(idx, list(range(njoint)))]


def build_cost_matrices(robot, njoint, horizon, Q, Rblk, r_scale, stage_w=0.1, term_w=10000.0):
    A, Bm = joint_dynamics(robot, njoint)
    ns, nu, H = 2 * njoint, njoint, horizon
    Aaug = np.vstack([np.linalg.matrix_power(A, i) for i in range(1, H + 1)])
    Baug = np.zeros((H * ns, H * nu))
    Qaug = np.zeros((H * ns, H * ns))
    Apow = [np.linalg.matrix_power(A, k) @ Bm for k in range(H)]
    for i in range(1, H + 1):
        Qaug[(i - 1) * ns:i * ns, (i - 1) * ns:i * ns] = Q * (term_w if i == H else stage_w)
        for j in range(1, i + 1):
            Baug[(i - 1) * ns:i * ns, (j - 1) * nu:j * nu] = Apow[i - j]
    R = np.eye(H * nu)
    for i in range(H):
        R[i * nu:(i + 1) * nu, i * nu:(i + 1) * nu] = Rblk
    R = R + R.T
    QQ = Baug.T @ Qaug @ Baug + R * r_scale
    return Aaug, Baug, Qaug, QQ


def build_linear_term(Aaug, Baug, Qaug, x0, gaug):
    """ff = ((Aaug*x0-gaug)'*Qaug*Baug)' ; caug = (Aaug*x0-gaug)'*Qaug*(Aaug*x0-gaug).  x0, gaug may be batched (B,.)"""
    x0 = np.atleast_2d(x0)
    gaug = np.atleast_2d(gaug)
    e = x0 @ Aaug.T - gaug
    eq = e @ Qaug
    ff = eq @ Baug
    caug = np.einsum("bi,bi->b", eq, e)
    return ff, caug


def straight_line_reference(x0_theta, xg_theta, horizon):
    """x_ of main_FANUC.m:38-49: linspace in joint space, zero velocity rows, step 0 dropped; batched (B,nj) ok."""
    t0 = np.atleast_2d(x0_theta)
    tg = np.atleast_2d(xg_theta)
    B, nj = t0.shape
    # MATLAB linspace: y = d1 + (0:n1).*(d2-d1)./n1 with y(end) = d2
    th = t0[:, None, :] + (np.arange(1, horizon + 1)[None, :, None] * (tg - t0)[:, None, :]) / horizon
    th[:, -1, :] = tg
    x = np.concatenate([th, np.zeros((B, horizon, nj))], axis=2)
    return x.reshape(B, horizon * 2 * nj)


def make_sys_info(robot, njoint, horizon, x0_theta, xg_theta, Q=None, Rblk=None, r_scale=50.0, lim=None,
                  max_input=None, epsilon_O=1e-1, MAX_O_ITER=20, x_ref=None):
    """The sys_info struct of main_FANUC.m:106-127 as a dict (single problem)."""
    Q = Q_MAIN_FANUC if Q is None else Q
    Rblk = R_MAIN_FANUC if Rblk is None else Rblk
    Aaug, Baug, Qaug, QQ = build_cost_matrices(robot, njoint, horizon, Q, Rblk, r_scale)
    x0 = np.concatenate([np.asarray(x0_theta, dtype=np.float64), np.zeros(njoint)])
    gaug = np.tile(np.concatenate([np.asarray(xg_theta, dtype=np.float64), np.zeros(njoint)]), horizon)
    ff, caug = build_linear_term(Aaug, Baug, Qaug, x0, gaug)
    if x_ref is None:
        x_ref = straight_line_reference(x0_theta, xg_theta, horizon)[0]
    if lim is None:
        lim = np.ones(njoint)
    if max_input is None:
        max_input = np.tile(np.array([1, 1, np.pi, np.pi, np.pi][:njoint]) * robot["delta_t"], horizon)
    return dict(Aaug=Aaug, Baug=Baug, QQ=QQ, ff=ff[0], Qaug=QQ, paug=ff[0], caug=float(caug[0]), robot=robot, H=horizon,
                nstate=2 * njoint, njoint=njoint, xR=x0.reshape(-1, 1), nu=njoint, x_=np.asarray(x_ref, dtype=np.float64),
                alpha=1.0 / np.linalg.svd(QQ, compute_uv=False).max(), lim=np.asarray(lim, dtype=np.float64),
                epsilon_O=epsilon_O, MAX_O_ITER=MAX_O_ITER, MAX_input=np.asarray(max_input, dtype=np.float64))


def cubicpolytraj(waypoints, wp_times, traj_times):
    """sampled_route = cubicpolytraj(route, wpTimes, trajTimes) of RRTstar_CFS.m:100 with the toolbox defaults
    (zero velocity at every waypoint): on segment k, q = q_k + (3 s^2 - 2 s^3)(q_{k+1} - q_k), s = (t - t_k)/(t_{k+1} - t_k).
    waypoints (nj, W) -> (nj, len(traj_times)).  The toolbox is not part of the reference tree: parity unpinned."""
    wp = np.asarray(waypoints, dtype=np.float64)
    tw = np.asarray(wp_times, dtype=np.float64)
    tt = np.asarray(traj_times, dtype=np.float64)
    seg = np.clip(np.searchsorted(tw, tt, side="right") - 1, 0, len(tw) - 2)
    s = (tt - tw[seg]) / (tw[seg + 1] - tw[seg])
    blend = 3 * s * s - 2 * s * s * s
    return wp[:, seg] + blend[None, :] * (wp[:, seg + 1] - wp[:, seg])
