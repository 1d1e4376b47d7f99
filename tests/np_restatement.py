"""A second, independent numpy restatement of the reference's geometry / calculus functions, written straight from the
.m files (vectorised MATLAB idioms kept: cell arrays -> lists, qr(.,0) -> numpy.linalg.qr, sort -> stable argsort).
Used only to cross-check the C oracle; it shares no code with it."""
import numpy as np


def cap_pos(base, DH, cap):
    """Lib/functions/CapPos.m:8-23"""
    nlink = DH.shape[0]
    M = np.eye(4)
    pos = []
    for i in range(nlink):
        th, d, a, al = DH[i]
        R = np.array([[np.cos(th), -np.sin(th) * np.cos(al), np.sin(th) * np.sin(al)],
                      [np.sin(th), np.cos(th) * np.cos(al), -np.cos(th) * np.sin(al)],
                      [0, np.sin(al), np.cos(al)]])
        T = np.array([a * np.cos(th), a * np.sin(th), d])
        X = np.eye(4)
        X[:3, :3] = R
        X[:3, 3] = T
        M = M @ X
        pos.append(np.stack([M[:3, :3] @ cap[i][:, k] + M[:3, 3] + base for k in range(2)], axis=1))
    return pos


def dist_lin_seg(p1s, p1e, p2s, p2e):
    """Lib/functions/distLinSeg.m:23-101"""
    fix = lambda v: 0.0 if v < 0 else (1.0 if v > 1 else v)
    d1, d2, d12 = p1e - p1s, p2e - p2s, p2s - p1s
    D1, D2 = np.sum(d1 ** 2), np.sum(d2 ** 2)
    S1, S2, R = np.sum(d1 * d12), np.sum(d2 * d12), np.sum(d1 * d2)
    den = D1 * D2 - R ** 2
    if D1 == 0 or D2 == 0:
        if D1 != 0:
            u, t = 0.0, fix(S1 / D1)
        elif D2 != 0:
            t, u = 0.0, fix(-S2 / D2)
        else:
            t = u = 0.0
    elif den == 0:
        t, u = 0.0, -S2 / D2
        uf = fix(u)
        if uf != u:
            t = fix((uf * R + S1) / D1)
            u = uf
    else:
        t = fix((S1 * D2 - S2 * R) / den)
        u = (t * R - S2) / D2
        uf = fix(u)
        if uf != u:
            t = fix((uf * R + S1) / D1)
            u = uf
    return np.linalg.norm(d1 * t - d2 * u - d12), np.concatenate([p1s + d1 * t, p2s + d2 * u])


def dist_arm(theta, robot, obs, kind):
    """dist_arm_3D_Heu_2.m / dist_arm_3D_200i_2.m (points(1:3) form)"""
    n = len(theta)
    DH = robot["DH"][:n].copy()
    DH[:, 0] = theta
    if kind == "M200i":
        DH[1, 0] = DH[1, 0] - np.pi / 2
    pos = cap_pos(robot["base"], DH, [c["p"] for c in robot["cap"]])
    d, lid = np.inf, 0
    for i in range(n):
        dis, pts = dist_lin_seg(pos[i][:, 0], pos[i][:, 1], obs[:, 0], obs[:, 1])
        if abs(dis) < 0.0001:
            dis = -np.linalg.norm(pts[:3] - pos[i][:, 1])
        if dis < d:
            d, lid = dis, i + 1
    return d, lid


def num_jac(f, x, eps=1e-5):
    """Lib/functions/num_jac.m:1-17 (xp(i) is not reset)"""
    f(x)
    xp = np.array(x, dtype=np.float64)
    g = np.zeros(len(x))
    for i in range(len(x)):
        xp[i] = x[i] + eps / 2
        yhi = f(xp)
        xp[i] = x[i] - eps / 2
        ylo = f(xp)
        g[i] = (yhi - ylo) / eps
    return g


def _vec2mat(vec, n, m):
    i, j = np.meshgrid(np.arange(1, n + 1), np.arange(0, m), indexing="ij")
    return vec[i + j - 1]


def derivest(fun, x0):
    """derivest.m:193-468 with defaults + 'Vectorized','no' (scalar x0)"""
    sr = 2.0000001
    h = max(x0, 0.02)
    delta = 100 * sr ** np.arange(0, -26, -1.0)
    srinv = 1.0 / sr
    # fdamat(sr,1,2)
    i, j = np.meshgrid(np.arange(1, 3), np.arange(1, 3), indexing="ij")
    c = 1.0 / np.array([1.0, 6.0])
    mat = c[j - 1] * srinv ** ((i - 1) * (2 * j - 1))
    fdarule = np.linalg.solve(mat.T, np.array([1.0, 0.0]))  # [1 0]/mat
    fp = np.array([fun(x0 + h * dl) for dl in delta])
    fm = np.array([fun(x0 - h * dl) for dl in delta])
    f_del = (fp - fm) / 2
    ne = 26 + 1 - 2 - 2
    der_init = _vec2mat(f_del, ne, 2) @ fdarule
    der_init = der_init / (h * delta[:ne])
    rombexpon = np.array([4.0, 6.0])
    rmat = np.ones((4, 3))
    rmat[1, 1:] = srinv ** rombexpon
    rmat[2, 1:] = srinv ** (2 * rombexpon)
    rmat[3, 1:] = srinv ** (3 * rombexpon)
    q, r = np.linalg.qr(rmat)
    rhs = _vec2mat(der_init, 4, max(1, ne - 4))
    coefs = np.linalg.solve(r, q.T @ rhs)
    der_romb = coefs[0]
    s = np.sqrt(np.sum((rhs - rmat @ coefs) ** 2, axis=0))
    rinv = np.linalg.solve(r, np.eye(3))
    cov1 = np.sum(rinv ** 2, axis=1)
    errors = s * 12.7062047361747 * np.sqrt(cov1[0])
    nest = len(der_romb)
    tags = np.argsort(der_romb, kind="stable")
    keep = np.ones(nest, dtype=bool)
    keep[[0, 1, nest - 2, nest - 1]] = False
    tags = tags[keep]
    err = errors[tags]
    ind = int(np.argmin(err))
    return der_romb[tags][ind], err[ind], h * delta[tags][ind]


# ---------------------------------------------------------------------------------------------------------------------------
# The solver classes, restated independently of the C oracle (own row assembly, own QP): Lib/CFS_FANUC.m, Lib/PSGCFS_FANUC.m,
# Lib/EVAL.m.  quadprog is replaced by a primal-dual interior-point method (the algorithm family quadprog's default
# 'interior-point-convex' belongs to) followed by an active-set polish to KKT <= 1e-10; the oracle uses a dual active-set
# (Goldfarb-Idnani) solver, so agreement of the two pins get_con / the loop / the stop rule above the leaf functions.
# ---------------------------------------------------------------------------------------------------------------------------
def cap_pos2(base, theta, robot):
    """Lib/2L/CapPos2.m:16-29 (planar chain: R = Rz(theta), T = robot.T(:,i+1))"""
    M = np.eye(4)
    M[:3, 3] = robot["T"][:, 0]
    pos = []
    for i in range(len(theta)):
        c, s = np.cos(theta[i]), np.sin(theta[i])
        X = np.eye(4)
        X[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
        X[:3, 3] = robot["T"][:, i + 1] if i + 1 < robot["T"].shape[1] else robot["T"][:, -1]
        M = M @ X
        cp = robot["cap"][i]["p"]
        pos.append(np.stack([M[:3, :3] @ cp[:, k] + M[:3, 3] + base for k in range(2)], axis=1))
    return pos


def dist_arm_all(theta, robot, obs, kind):
    if kind != "2L":
        return dist_arm(theta, robot, obs, kind)
    raise NotImplementedError("the 2L chain is cross-checked at leaf level only (tests/test_oracle.py)")


def qp_ipm(Hm, f, A, b, tol=1e-11, max_iter=200):
    """min 1/2 x'Hx + f'x  s.t.  A x <= b   (H symmetric positive definite): Mehrotra predictor-corrector, then the active set
    is read off the multipliers and the equality-constrained KKT system is solved exactly (polish).  Returns x, kkt, n_active;
    raises ValueError when the iteration does not converge (infeasible problem)."""
    Hm = 0.5 * (Hm + Hm.T)
    Ab = np.unique(np.column_stack([A, b]), axis=0)             # the velocity rows are duplicated once per obstacle (CFS_FANUC.m:110)
    A, b = Ab[:, :-1], Ab[:, -1]
    m, n = A.shape
    sc = np.abs(Hm).max()                                       # scale the objective: multipliers of order one
    Hs, fs = Hm / sc, f / sc
    x = np.linalg.solve(Hs, -fs)
    s = np.maximum(b - A @ x, 1.0)
    z = np.ones(m)
    ok = False
    for it in range(max_iter):
        rd = Hs @ x + fs + A.T @ z
        rp = A @ x + s - b
        mu = s @ z / m
        if max(np.abs(rd).max(), np.abs(rp).max()) < 1e-9 and mu < 1e-10:
            ok = True                                           # moderately converged: the polish below makes it exact
            break
        d = z / s
        K = Hs + A.T @ (d[:, None] * A)

        def solve(rc):
            rhs = -rd - A.T @ (d * rp) + A.T @ (rc / s)          # eliminate ds, dz
            dx = np.linalg.solve(K, rhs)
            ds = -rp - A @ dx
            dz = -(rc + z * ds) / s
            return dx, ds, dz
        dxa, dsa, dza = solve(s * z)                            # predictor (mu = 0)

        def step(v, dv):
            neg = dv < 0
            return min(1.0, (-v[neg] / dv[neg]).min()) if neg.any() else 1.0
        aa = min(step(s, dsa), step(z, dza))
        mu_aff = (s + aa * dsa) @ (z + aa * dza) / m
        sigma = (mu_aff / mu) ** 3
        dx, ds, dz = solve(s * z + dsa * dza - sigma * mu)      # corrector
        a = 0.995 * min(step(s, ds), step(z, dz))
        x, s, z = x + a * dx, s + a * ds, z + a * dz
        if not (np.isfinite(x).all() and np.isfinite(z).all()) or np.abs(z).max() > 1e14:
            break
    if not ok:
        raise ValueError("interior-point iteration did not converge (QP infeasible?)")
    # polish: active set = rows whose multiplier dominates their slack
    act = np.where(z > s)[0]
    Lc = np.linalg.cholesky(Hm)
    hinv = lambda v: np.linalg.solve(Lc.T, np.linalg.solve(Lc, v))
    xu = hinv(-f)
    for _ in range(400):                                        # one change per pass: most violated row in, else worst multiplier out
        if len(act):                                             # Schur complement on the active rows: (Aa H^-1 Aa') lam = Aa xu - ba
            Aa = A[act]
            Y = hinv(Aa.T)
            lam = np.linalg.solve(Aa @ Y, Aa @ xu - b[act])
            xp = xu - Y @ lam
        else:
            xp, lam = xu, np.zeros(0)
        viol = (A @ xp - b) / (1 + np.abs(b))
        viol[act] = 0.0
        if viol.max() > 1e-10:
            act = np.append(act, int(np.argmax(viol)))
        elif len(act) and lam.min() < -1e-9 * max(np.abs(lam).max(), 1e-300):
            act = np.delete(act, int(np.argmin(lam)))
        else:
            kkt = np.abs(Hm @ xp + f + (A[act].T @ lam if len(act) else 0)).max() / np.abs(Hm).max()
            return xp, kkt, len(act)
    raise ValueError("active-set polish did not settle")


def get_con(x_, u, x0, robot, obs_list, kind, Aaug, Baug, lim, H, nj, margin_key):
    """Lib/CFS_FANUC.m:101-135 (margin obs.epsilon) / Lib/PSGCFS_FANUC.m:145-184 (margin obs.D): per obstacle j and step i one
    obstacle row, then nj rows +Baug_w and nj rows -Baug_w -- the velocity rows are appended INSIDE the obstacle loop, so
    they are duplicated once per obstacle (CFS_FANUC.m:110,126-129)."""
    ns = 2 * nj
    Ls, Ss = [], []
    for obs in obs_list:
        for i in range(H):
            theta = x_[ns * i:ns * i + nj]
            f = lambda th: dist_arm(th, robot, obs["l"], kind)[0]
            d, _ = dist_arm(theta, robot, obs["l"], kind)
            g = num_jac(f, theta)
            Bj = Baug[ns * i:ns * (i + 1), :]
            Ls.append(-g @ Bj[:nj, :])
            Ss.append((d - obs[margin_key]) - g @ Bj[:nj, :] @ u)
            if lim is not None:
                Aw = Aaug[ns * i + nj:ns * (i + 1), :] @ x0
                for k in range(nj):
                    Ls.append(Bj[nj + k, :])
                    Ss.append(lim[k] - Aw[k])
                for k in range(nj):
                    Ls.append(-Bj[nj + k, :])
                    Ss.append(lim[k] + Aw[k])
    return np.array(Ls), np.array(Ss)


def cfs_optimizer(s, robot, obs_list, kind, psg=False, noise=None):
    """CFS_FANUC.optimizer (Lib/CFS_FANUC.m:62-98) / PSGCFS_FANUC.optimizer (Lib/PSGCFS_FANUC.m:65-128) with EVAL's stop rule
    (Lib/EVAL.m:61-73: ||x_ - x_old|| < epsilon_O or iter_O > MAX_O_ITER, x_old initialised to ones and, in PSGCFS, never
    updated).  Returns u, x_, cost_all, iters."""
    H, nj = s["H"], s["njoint"]
    n = H * nj
    A, Bm = s["Aaug"], s["Baug"]
    QQ, ff, caug = s["QQ"], s["ff"], s["caug"]
    x0 = s["xR"][:, 0]
    x_ = np.array(s["x_"], dtype=np.float64)
    u = np.zeros(n)                                              # CFS_FANUC.m:56 (iteration 1 linearises at x_ but around u = 0)
    x_old = np.ones_like(x_)                                     # EVAL.m:47
    cost_all = []
    cost_old, cost_new = 100000.0, caug                          # EVAL.m:29, PSGCFS_FANUC.m:66
    iter_O = 1
    get_cost = lambda v: 0.5 * v @ QQ @ v + ff @ v + caug        # EVAL.m:51-53
    while not (np.linalg.norm(x_ - x_old) < s["epsilon_O"] or iter_O > s["MAX_O_ITER"]):
        Ainq, binq = get_con(x_, u, x0, robot, obs_list, kind, A, Bm, s.get("lim"), H, nj, "D" if psg else "epsilon")
        if psg:
            if not abs(cost_new - cost_old) < 1e-4:              # stop_inner, PSGCFS_FANUC.m:136-142 (MAX_I_ITER = 1)
                cost_old = cost_new
                nz = noise[iter_O - 1] if noise is not None else np.zeros(n)
                u_ = u - s["alpha"] * ((QQ @ u + ff) + 10 * nz / (iter_O ** 2 + 1))     # PSG_update_arm, :106-112
                u, _, _ = qp_ipm(np.eye(n), -u_, Ainq, binq)                          # Projection, :115-128 (no bounds)
        else:
            mi = s["MAX_input"]
            Aall = np.vstack([Ainq, np.eye(n), -np.eye(n)])                          # lb / ub of CFS_FANUC.m:85
            ball = np.concatenate([binq, mi, mi])
            x_old = x_.copy()                                                         # CFS_FANUC.m:88
            u, _, _ = qp_ipm(QQ, ff, Aall, ball)
        x_ = A @ x0 + Bm @ u                                     # roll-out, CFS_FANUC.m:90-94
        cost_new = get_cost(u)
        cost_all.append(cost_new)
        iter_O += 1
    return u, x_, np.array(cost_all), iter_O - 1


def dist_link(theta, robot, obs, kind, linkid):
    """dist_link_Heu.m:1-26 / dist_link_200i.m:1-25 (the 200i variant subtracts pi/2 from joint 2, :8)"""
    n = len(theta)
    DH = robot["DH"][:n].copy()
    DH[:, 0] = theta
    if kind == "M200i":
        DH[1, 0] = DH[1, 0] - np.pi / 2
    pos = cap_pos(robot["base"], DH, [c["p"] for c in robot["cap"]])
    i = linkid - 1
    dis, pts = dist_lin_seg(pos[i][:, 0], pos[i][:, 1], obs[:, 0], obs[:, 1])
    if abs(dis) < 0.0001:
        dis = -np.linalg.norm(pts[:3] - pos[i][:, 1])
    return dis


def chomp_dm(theta, robot, ob):
    """CHOMP_FANUC.dm_f (Lib/CHOMP_FANUC.m:115-134): DH(i,1) = theta(i) and nothing else, per-link distance minus obs.D"""
    n = len(theta)
    DH = robot["DH"][:n].copy()
    DH[:, 0] = theta
    pos = cap_pos(robot["base"], DH, [c["p"] for c in robot["cap"]])
    d = np.zeros(n)
    for i in range(n):
        dis, pts = dist_lin_seg(pos[i][:, 0], pos[i][:, 1], ob["l"][:, 0], ob["l"][:, 1])
        if abs(dis) < 0.0001:
            dis = -np.linalg.norm(pts[:3] - pos[i][:, 1])
        d[i] = dis - ob["D"]
    return d


def chomp_optimizer(s, robot, obs_list, kind, uu):
    """CHOMP_FANUC.optimizer (Lib/CHOMP_FANUC.m:54-165) with the literal Baug row slices of :153/:158.  eval.x_ / eval.x_old are
    never touched by the class, so the loop runs MAX_O_ITER times.  Returns u, x_, cost_all, e_u_all."""
    H, nj, ns = s["H"], s["njoint"], 2 * s["njoint"]
    n = H * nj
    A, Bm = s["Aaug"], s["Baug"]
    QQ, ff, caug = s["QQ"], s["ff"], s["caug"]
    x0 = s["xR"][:, 0]
    x_ = np.array(s["x_"], dtype=np.float64)
    u = np.array(uu, dtype=np.float64)
    cost_all, e_u_all = [], []
    for _ in range(int(s["MAX_O_ITER"])):
        u_old = u.copy()
        dc_all = np.zeros(n)                                     # dcostObs_f, :137-165
        for i in range(1, H + 1):
            theta = x_[ns * (i - 1): ns * (i - 1) + nj]
            for ob in obs_list:
                Dfx = chomp_dm(theta, robot, ob)
                lid = int(np.argmin(Dfx)) + 1
                rows = Bm[(i - 1) * nj: i * nj, :]                # Baug((i-1)*njoint+1:i*njoint,:)  -- as written
                if Dfx[lid - 1] < 0 or Dfx[lid - 1] <= ob["epsilon"]:
                    dD = np.zeros(nj)
                    for sj in range(nj):
                        def f(xv, sj=sj):
                            th = theta.copy()
                            th[sj] = xv
                            return dist_link(th, robot, ob["l"], kind, lid)
                        dD[sj] = derivest(f, theta[sj])[0]
                    if Dfx[lid - 1] < 0:
                        dc_all += -(dD @ rows)
                    else:
                        dc_all += (1.0 / ob["epsilon"]) * (Dfx[lid - 1] - ob["epsilon"]) * (dD @ rows)
        u = u_old - s["alpha"] * 3 * ((QQ @ u_old + ff) + 2000 * dc_all)          # :75
        x_ = A @ x0 + Bm @ u                                                       # :77-82
        fobs = 0.0                                                                 # fobs_m, :91-112
        for i in range(1, H + 1):
            theta = x_[ns * (i - 1): ns * (i - 1) + nj]
            for ob in obs_list:
                for dv in chomp_dm(theta, robot, ob):
                    if dv < 0:
                        fobs += -dv + 0.5 * ob["epsilon"]
                    elif dv <= ob["epsilon"]:
                        fobs += (1.0 / (2 * ob["epsilon"])) * (dv - ob["epsilon"]) ** 2
        cost_all.append(0.5 * u @ QQ @ u + ff @ u + caug + fobs)
        e_u_all.append(np.linalg.norm(u_old - u))
    return u, x_, np.array(cost_all), np.array(e_u_all)


def rrt_find_route(robot, kind, obs_list, x0, goal, goal_th, region_g, region_s, sample_off, ratial, rnd, bi=0.5, max_iter=400,
                   star=False):
    """RRT_FANUC.find_route (Lib/RRT_FANUC.m:63-207) with MATLAB's rand replaced by the stream `rnd`, consumed in the reference's
    order (pp = rand; rand(nstate,1) when pp < bi -- :108,111).  1-based parents as in all_nodes(1,:); -1 for the root.
    Returns route (nj, len), all_nodes (1+nj, node_num), total_dis, fail, numbers of the stream consumed."""
    nj = len(x0)
    it = iter(np.asarray(rnd, dtype=np.float64))
    used = [0]

    def rand():
        used[0] += 1
        return next(it)

    def feasible(theta):                                         # :146-181
        for ob in obs_list:
            DH = robot["DH"][:nj].copy()
            DH[:, 0] = theta
            if kind == "M200i":
                DH[1, 0] = DH[1, 0] - np.pi / 2
            pos = cap_pos(robot["base"], DH, [c["p"] for c in robot["cap"]])
            for i in range(nj):
                dis, pts = dist_lin_seg(pos[i][:, 0], pos[i][:, 1], ob["l"][:, 0], ob["l"][:, 1])
                if abs(dis) < 0.0001:
                    dis = -np.linalg.norm(pts[:3] - pos[i][:, 1])
                if dis < ob["D"]:
                    return False
        return True

    new = np.array(x0, dtype=np.float64)
    all_nodes = np.concatenate([[-1.0], new])[:, None]
    total_dis = [0.0]
    node_num, fail, parent = 1, False, 1
    to_dis = []

    def reached():                                               # :193-207
        nonlocal fail
        ok = bool(np.all(goal - region_g < new) and np.all(new < goal + region_g))
        if node_num > max_iter:
            fail = True
            ok = True
        return ok

    done = reached()
    while not done:
        while True:                                              # getNode, :97-104
            pp = rand()
            if pp < bi:
                sample = (np.array([rand() for _ in range(nj)]) - 0.5) * region_s * 2 + sample_off
            else:
                sample = np.array(goal_th, dtype=np.float64)
            to_dis = [np.linalg.norm((all_nodes[1:, i] - sample) * ratial) for i in range(node_num)]
            parent = int(np.argmin(to_dis)) + 1                  # strict <: the first minimum (:120-127)
            pn = all_nodes[1:, parent - 1]
            new = pn + (sample - pn) * 0.1 / np.linalg.norm(pn - sample)
            if feasible(new):
                break
        all_nodes = np.hstack([all_nodes, np.concatenate([[parent], new])[:, None]])   # addNode, :184-190
        total_dis.append(total_dis[parent - 1] + to_dis[parent - 1])
        node_num += 1
        if star:                                                 # arrangeNode, :134-142
            for i in range(len(to_dis)):
                if to_dis[i] < 0.2 and total_dis[i] > total_dis[-1] + to_dis[i]:
                    all_nodes[0, i] = node_num
                    total_dis[i] = total_dis[-1] + to_dis[i]
        done = reached()
    route = new[:, None]
    p = parent if node_num > 1 else -1
    if node_num == 1:
        p = -1
    guard = 0
    while p != -1 and guard <= node_num:
        route = np.hstack([all_nodes[1:, int(p) - 1][:, None], route])
        p = all_nodes[0, int(p) - 1]
        guard += 1
    return route, all_nodes, np.array(total_dis), fail, used[0]
