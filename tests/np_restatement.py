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
