"""ctypes binding of the CPU ORACLE (oracle/cfs_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (motionplanning_5d_m_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcfs_oracle.so")
_SO_TWIN = os.path.join(_HERE, "libcfs_oracle_fma.so")

M16IB, M200I, R2L = 0, 1, 2
KIND = {"M16iB": M16IB, "M200i": M200I, "2L": R2L}
MAXL = 6


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("cfs_oracle.c", "cfs_oracle.h")]
    if force or not os.path.exists(_SO) or not os.path.exists(_SO_TWIN) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Robot(C.Structure):
    _fields_ = [("kind", C.c_int), ("nj", C.c_int), ("DH", (C.c_double * 4) * MAXL), ("base", C.c_double * 3),
                ("cap", ((C.c_double * 3) * 2) * MAXL), ("T2L", (C.c_double * 3) * 3), ("dt", C.c_double)]


class Cfg(C.Structure):
    _fields_ = [("H", C.c_int), ("nobs", C.c_int), ("obs", C.c_void_p), ("margin", C.c_void_p), ("QQ", C.c_void_p),
                ("lim", C.c_void_p), ("max_input", C.c_void_p), ("eps_outer", C.c_double), ("max_outer", C.c_int),
                ("solver", C.c_int), ("grad", C.c_int), ("alpha", C.c_double)]


_lib = None
_twin = None


def twin():
    """The same restatement compiled with FMA contraction (oracle/Makefile): conditioning probe only."""
    global _twin
    if _twin is None:
        build()
        with open("/proc/cpuinfo") as f:
            if " fma" not in f.read():
                raise RuntimeError("this CPU has no FMA: the twin oracle cannot run")
        _twin = C.CDLL(_SO_TWIN)
    return _twin


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_dist_lin_seg.restype = C.c_double
        _lib.orc_dist_arm.restype = C.c_double
        _lib.orc_dist_link.restype = C.c_double
        _lib.orc_kkt_residual.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def robot(name):
    r = Robot()
    lib().orc_robot_init(C.byref(r), KIND[name])
    return r


def cap_pos(r, theta):
    pos = np.zeros((r.nj, 2, 3))
    lib().orc_cap_pos(C.byref(r), _p(_f64(theta)), _p(pos))
    return pos


def dist_lin_seg(p1s, p1e, p2s, p2e):
    p1s, p1e, p2s, p2e = map(_f64, (p1s, p1e, p2s, p2e))
    dim = p1s.size
    pts = np.zeros((2, dim))
    d = lib().orc_dist_lin_seg(_p(p1s), _p(p1e), _p(p2s), _p(p2e), C.c_int(dim), _p(pts))
    return d, pts


def obs6(l, box=False):
    """obs{j}.l (3x2, columns = endpoints, or min / max corner of a box) -> obstacle record [l(:,1), l(:,2), kind] (7 doubles)"""
    if isinstance(l, dict):
        l, box = l["l"], l.get("shape") == "box"
    l = np.asarray(l, dtype=np.float64)
    return _f64(np.concatenate([l[:, 0], l[:, 1], [1.0 if box else 0.0]]))


def dist_seg_box(ps, pe, lo, hi):
    """N3 extension: (distance, closest point of the segment) between [ps, pe] and the solid box [lo, hi]"""
    pt = np.zeros(3)
    f = lib().orc_dist_seg_box
    f.restype = C.c_double
    d = f(_p(_f64(ps)), _p(_f64(pe)), _p(_f64(lo)), _p(_f64(hi)), _p(pt))
    return d, pt


def dist_arm(r, theta, o6):
    lid, t = C.c_int(0), C.c_int(0)
    d = lib().orc_dist_arm(C.byref(r), _p(_f64(theta)), _p(_f64(o6)), C.byref(lid), C.byref(t))
    return d, lid.value, t.value


def dist_link(r, theta, o6, linkid):
    t = C.c_int(0)
    return lib().orc_dist_link(C.byref(r), _p(_f64(theta)), _p(_f64(o6)), C.c_int(linkid), C.byref(t))


def num_jac(r, theta, o6):
    g = np.zeros(r.nj)
    t = C.c_int(0)
    lib().orc_num_jac(C.byref(r), _p(_f64(theta)), _p(_f64(o6)), _p(g), C.byref(t))
    return g


def derivest_named(which, x0):
    d, e, f = C.c_double(), C.c_double(), C.c_double()
    lib().orc_derivest_named(C.c_int(which), C.c_double(x0), C.byref(d), C.byref(e), C.byref(f))
    return d.value, e.value, f.value


def derivest_grad(r, theta, o6, linkid):
    g = np.zeros(r.nj)
    t = C.c_int(0)
    lib().orc_derivest_grad(C.byref(r), _p(_f64(theta)), _p(_f64(o6)), C.c_int(linkid), _p(g), C.byref(t))
    return g


def build_cost(nj, H, dt, Q, Rblk, r_scale, stage_w=0.1, term_w=10000.0):
    ns, n, N = 2 * nj, nj * H, 2 * nj * H
    Q = np.asfortranarray(Q, dtype=np.float64)
    Rblk = np.asfortranarray(Rblk, dtype=np.float64)
    Aaug = np.zeros((N, ns), order="F")
    Baug = np.zeros((N, n), order="F")
    QQ = np.zeros((n, n), order="F")
    lib().orc_build_cost(C.c_int(nj), C.c_int(H), C.c_double(dt), _p(Q), _p(Rblk), C.c_double(r_scale),
                         C.c_double(stage_w), C.c_double(term_w), _p(Aaug), _p(Baug), None, _p(QQ))
    return Aaug, Baug, QQ


def build_ff(nj, H, Q, Aaug, Baug, x0, gaug, stage_w=0.1, term_w=10000.0):
    n = nj * H
    ff = np.zeros(n)
    caug = C.c_double()
    lib().orc_build_ff(C.c_int(nj), C.c_int(H), _p(np.asfortranarray(Q, dtype=np.float64)), C.c_double(stage_w),
                       C.c_double(term_w), _p(np.asfortranarray(Aaug)), _p(np.asfortranarray(Baug)), _p(_f64(x0)),
                       _p(_f64(gaug)), _p(ff), C.byref(caug))
    return ff, caug.value


class Problem:
    """Holds the arrays an orc_cfg points to (keeps them alive)."""

    def __init__(self, r, H, obs_list, margin, QQ, lim, max_input, eps_outer, max_outer, solver=0, grad=0, alpha=0.0):
        self.r = r
        self.obs = _f64(np.concatenate([obs6(o) for o in obs_list]))
        self.margin = _f64(margin)
        self.QQ = np.asfortranarray(QQ, dtype=np.float64)
        self.lim = None if lim is None else _f64(lim)
        self.max_input = None if max_input is None else _f64(max_input)
        self.cfg = Cfg(H, len(obs_list), _p(self.obs), _p(self.margin), _p(self.QQ), _p(self.lim), _p(self.max_input),
                       eps_outer, max_outer, solver, grad, alpha)
        self.H, self.nj = H, r.nj
        self.n, self.N = r.nj * H, 2 * r.nj * H

    def get_con(self, x0, xcur, u):
        c = self.cfg
        rows = c.nobs * self.H * ((1 + 2 * self.nj) if self.lim is not None else 1)
        A = np.zeros((rows, self.n))
        b = np.zeros(rows)
        dist = np.zeros(c.nobs * self.H)
        lid = np.zeros(c.nobs * self.H, dtype=np.int32)
        grad = np.zeros((c.nobs * self.H, self.nj))
        t = C.c_int(0)
        lib().orc_get_con(C.byref(self.r), C.byref(c), _p(_f64(x0)), _p(_f64(xcur)), _p(_f64(u)), _p(A), _p(b),
                          _p(dist), _p(lid), _p(grad), C.byref(t))
        return A, b, dist, lid, grad, t.value

    def solve_batch(self, x0, ff, caug, xref, noise=None, nthreads=0, use_twin=False):
        """x0 (B,2nj) ff (B,n) caug (B,) xref (B,N) noise (B,max_outer,n) -> dict.  use_twin: the FMA-contracted build."""
        x0, ff, caug, xref = map(_f64, (x0, ff, caug, xref))
        B = x0.shape[0]
        K = self.cfg.max_outer
        u = np.zeros((B, self.n))
        x = np.zeros((B, self.N))
        cost = np.zeros((B, K))
        e_u = np.zeros((B, K))
        iters = np.zeros(B, dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        nz = None if noise is None else _f64(noise)
        qp = np.zeros((B, 2), dtype=np.int32)
        (twin() if use_twin else lib()).orc_cfs_solve_batch2(C.byref(self.r), C.byref(self.cfg), C.c_int(B), C.c_int(nthreads), _p(x0), _p(ff),
                                   _p(caug), _p(xref), _p(nz), _p(u), _p(x), _p(cost), _p(e_u), _p(iters), _p(status),
                                   _p(qp))
        return dict(u=u, x=x, cost_hist=cost, e_u_hist=e_u, iters=iters, status=status, qp_iters=qp[:, 0],
                    qp_max_active=qp[:, 1])

    def chomp_batch(self, D, eps, x0, ff, caug, xref, u_init, nthreads=0, use_twin=False):
        """CHOMP_FANUC.optimizer for B problems: D / eps = obs{j}.D / obs{j}.epsilon, u_init (B,n) = the constructor's uu."""
        x0, ff, caug, xref, u_init = map(_f64, (x0, ff, caug, xref, u_init))
        D, eps = _f64(np.atleast_1d(D)), _f64(np.atleast_1d(eps))
        B = x0.shape[0]
        K = self.cfg.max_outer
        u = np.zeros((B, self.n))
        x = np.zeros((B, self.N))
        cost = np.zeros((B, K))
        e_u = np.zeros((B, K))
        iters = np.zeros(B, dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        (twin() if use_twin else lib()).orc_chomp_solve_batch(C.byref(self.r), C.byref(self.cfg), _p(D), _p(eps), C.c_int(B),
                                                               C.c_int(nthreads), _p(x0), _p(ff), _p(caug), _p(xref), _p(u_init),
                                                               _p(u), _p(x), _p(cost), _p(e_u), _p(iters), _p(status))
        return dict(u=u, x=x, cost_hist=cost, e_u_hist=e_u, iters=iters, status=status)


def chol_J0(G):
    n = G.shape[0]
    J0 = np.zeros((n, n), order="F")
    rc = lib().orc_chol_J0(C.c_int(n), _p(np.asfortranarray(G, dtype=np.float64)), _p(J0))
    assert rc == 0
    return J0


def qp_gi(G, a, Cm, d):
    """min 1/2 x'Gx + a'x s.t. Cm x <= d ; returns x, lam, rc, iters, kkt"""
    G = np.asfortranarray(G, dtype=np.float64)
    a, d = _f64(a), _f64(d)
    Cm = _f64(Cm)
    n, m = G.shape[0], Cm.shape[0]
    J0 = chol_J0(G)
    xunc = _f64(-np.linalg.solve(G, a))
    x = np.zeros(n)
    lam = np.zeros(m)
    it = C.c_int(0)
    rc = lib().orc_qp_gi(C.c_int(n), C.c_int(m), _p(J0), _p(xunc), _p(Cm), _p(d), _p(x), _p(lam), C.byref(it))
    kkt = lib().orc_kkt_residual(C.c_int(n), C.c_int(m), _p(G), _p(a), _p(Cm), _p(d), _p(x), _p(lam)) if rc == 0 else np.inf
    return x, lam, rc, it.value, kkt


def rrt_feasible(r, theta, obs_list, D):
    o = _f64(np.concatenate([obs6(o) for o in obs_list]))
    dm, t = C.c_double(), C.c_int(0)
    f = lib().orc_rrt_feasible(C.byref(r), _p(_f64(theta)), C.c_int(len(obs_list)), _p(o), _p(_f64(D)), C.byref(dm),
                               C.byref(t))
    return bool(f), dm.value


def rrt_nearest(nodes, sample, ratial):
    nodes = _f64(nodes)
    nn, nj = nodes.shape
    d = np.zeros(nn)
    lib().orc_rrt_nearest.restype = C.c_int
    p = lib().orc_rrt_nearest(C.c_int(nj), C.c_int(nn), _p(nodes), _p(_f64(sample)), _p(_f64(ratial)), _p(d))
    return p, d


def rrt_find_route(r, obs_list, D, x0, goal, region_g, region_s, sample_off, goal_th, ratial, rnd, bi=0.5, max_iter=400,
                   star=True):
    """RRT_FANUC.find_route (Lib/RRT_FANUC.m:63-207) with the uniform random stream `rnd` consumed in MATLAB's order.
    Returns dict(route (len, nj), n_nodes, fail, rnd_used, nodes, parent, total_dis) or None if rnd ran out."""
    o = _f64(np.concatenate([obs6(o) for o in obs_list]))
    nj = r.nj
    cap = max_iter + 2
    nodes = np.zeros((cap, nj))
    parent = np.zeros(cap, dtype=np.int32)
    tot = np.zeros(cap)
    route = np.zeros((cap, nj))
    nn, fl, ru = C.c_int(0), C.c_int(0), C.c_int(0)
    rnd = _f64(rnd)
    f = lib().orc_rrt_find_route
    f.restype = C.c_int
    ln = f(C.byref(r), C.c_int(len(obs_list)), _p(o), _p(_f64(D)), _p(_f64(x0)), _p(_f64(goal)), _p(_f64(region_g)),
           _p(_f64(region_s)), _p(_f64(sample_off)), _p(_f64(goal_th)), _p(_f64(ratial)), C.c_double(bi), C.c_int(max_iter),
           C.c_int(1 if star else 0), _p(rnd), C.c_int(rnd.size), C.c_int(cap), _p(nodes), _p(parent), _p(tot), _p(route),
           C.byref(nn), C.byref(fl), C.byref(ru))
    if ln < 0:
        return None
    return dict(route=route[:ln].copy(), n_nodes=nn.value, fail=bool(fl.value), rnd_used=ru.value, nodes=nodes[:nn.value].copy(),
                parent=parent[:nn.value].copy(), total_dis=tot[:nn.value].copy())
