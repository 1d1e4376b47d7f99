"""-m gpu : the headline configuration at FULL size (4096 problems, M16iB, H = 50) through size-independent properties --
the oracle needs ~3 s per pass there (bench.py's parity_sample does that comparison on every run); these checks need no oracle."""
import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib, problem, synthetic

pytestmark = pytest.mark.gpu

B, H, NJ, K = 4096, 50, 5, 20


@pytest.fixture(scope="module")
def full(ctx):
    robot = dict(M.robotproperty2("M16iB"))
    robot["name"] = "M16iB"
    ctx.set_robot(robot, NJ)
    ctx.set_obstacles([synthetic.OBS_M16IB])
    cfg = synthetic.batch_config_m16ib(B, lambda cand: ctx.nodes_feasible(cand)[0], horizon=H)
    s = cfg["sys_info"]
    ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    out = ctx.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], s["epsilon_O"], K)
    return cfg, s, out


def test_full_batch_statuses_and_histories(ctx, full):
    cfg, s, out = full
    st, it = out["status"] & 0xFF, out["iters"]
    assert set(np.unique(st)) <= {0, 1, 2} and (it >= 0).all() and (it <= K).all()
    assert (st == 0).sum() > 0.7 * B and (st == 2).any() and (st == 1).any()     # the batch exercises every outcome
    assert (it[st == 1] == K).all()
    for b in range(B):                                                           # cost_all has exactly iters entries
        assert np.isfinite(out["cost_hist"][b, :it[b]]).all() and np.isnan(out["cost_hist"][b, it[b]:]).all()
    # an infeasible first QP leaves the initial iterate (u = 0, x_ = the reference line)
    first = (st == 2) & (it == 0)
    assert first.any() and not out["u"][first].any() and np.array_equal(out["x"][first], cfg["xref"][first])


def test_full_batch_rollout_bounds_and_cost(ctx, full):
    """x_ is the roll-out of u (CFS_FANUC.m:90-94); u respects MAX_input and the velocity rows (:85, :126-129); the cost the
    kernel reports by duality equals EVAL.get_cost = 1/2 u'QQu + ff'u + caug (EVAL.m:51-53) evaluated directly."""
    cfg, s, out = full
    ok = ((out["status"] & 0xFF) < 2) & (out["iters"] > 0)
    u, x = out["u"][ok], out["x"][ok]
    x0 = cfg["x0"][ok]
    xr = x0 @ s["Aaug"].T + u @ s["Baug"].T
    assert np.abs(xr - x).max() < 1e-11
    assert (np.abs(u) <= s["MAX_input"][None, :] * (1 + 1e-9) + 1e-12).all()
    om = x.reshape(-1, H, 2 * NJ)[:, :, NJ:]
    assert (np.abs(om) <= s["lim"][None, None, :] + 1e-8).all()
    cost = 0.5 * np.einsum("bi,ij,bj->b", u, s["QQ"], u) + np.einsum("bi,bi->b", cfg["ff"][ok], u) + cfg["caug"][ok]
    last = out["cost_hist"][ok, out["iters"][ok] - 1]
    assert np.all(np.abs(last - cost) <= 1e-8 * np.abs(cost) + 1e-9)


def test_full_batch_is_deterministic_and_order_independent(ctx, full):
    """Every problem is solved independently of the device work queue: a second run, a permuted batch and single-problem
    calls give bit-identical answers."""
    cfg, s, out = full
    args = [cfg[k] for k in ("x0", "ff", "caug", "xref")]
    again = ctx.solve_batch(*args, s["epsilon_O"], K)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(again[k], out[k]), k
    perm = np.random.default_rng(3).permutation(B)
    pout = ctx.solve_batch(*[a[perm] for a in args], s["epsilon_O"], K)
    for k in ("u", "x", "iters", "status"):
        assert np.array_equal(pout[k], out[k][perm]), k
    hard = np.argsort(-out["iters"])[:3]
    for b in list(hard) + [0, 1]:
        one = ctx.solve_batch(*[a[b:b + 1] for a in args], s["epsilon_O"], K)
        assert np.array_equal(one["u"][0], out["u"][b]) and int(one["status"][0]) == int(out["status"][b])


def test_full_size_distance_gradient_properties(ctx, full):
    """K1 at the bench size (204 800 waypoints): link ids in range, distances of the rejection-sampled end points >= D,
    gradients finite, and the first-order model d(theta + delta) ~ d + grad'delta holds to O(|delta|^2) away from kinks."""
    cfg, s, out = full
    th = cfg["xref"].reshape(B, H, 2 * NJ)[:, :, :NJ].reshape(-1, NJ)
    d, lid, g, fl = ctx.dist_grad(th)
    d, lid, g = d.reshape(-1), lid.reshape(-1), g.reshape(-1, NJ)
    assert ((lid >= 1) & (lid <= NJ)).all() and np.isfinite(d).all() and np.isfinite(g).all()
    ends = d.reshape(B, H)[:, -1]
    assert (ends >= synthetic.OBS_M16IB["D"] - 1e-12).all()                      # goals were sampled feasible (RRT_FANUC.m:172)
    delta = 1e-4 * np.random.default_rng(7).standard_normal(th.shape)
    d2, lid2, _, _ = ctx.dist_grad(th + delta)
    d2, lid2 = d2.reshape(-1), lid2.reshape(-1)
    same = (lid2 == lid) & (d > 1e-3)
    err = np.abs(d2 - (d + np.einsum("ij,ij->i", g, delta)))[same]
    assert same.mean() > 0.95 and np.percentile(err, 99) < 1e-6
