"""-m "not gpu": host-side logic and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import motionplanning_5d_m_b200 as M
from motionplanning_5d_m_b200 import _lib, problem
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cfs_b200.h")).read()
    declared = set(re.findall(r"\b(cfs_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in ctypes.c_char_p(ctypes.cast(lib.cfs_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    try:
        M.Context(0)
    except M.CfsError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context creation must fail without a GPU")


def test_cost_builder_matches_oracle(oracle):
    robot = M.robotproperty2("M200i")
    H = 12
    Aaug, Baug, Qaug, QQ = problem.build_cost_matrices(robot, 5, H, problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0)
    Ao, Bo, QQo = oracle.build_cost(5, H, 0.5, problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0)
    assert np.abs(Aaug - Ao).max() == 0 and np.abs(Baug - Bo).max() == 0
    assert np.abs(QQ - QQo).max() <= 1e-12 * np.abs(QQ).max()
    rng = np.random.default_rng(0)
    x0 = np.concatenate([rng.normal(size=5), np.zeros(5)])
    gaug = np.tile(np.concatenate([rng.normal(size=5), np.zeros(5)]), H)
    ff, caug = problem.build_linear_term(Aaug, Baug, Qaug, x0, gaug)
    ffo, co = oracle.build_ff(5, H, problem.Q_MAIN_FANUC, Ao, Bo, x0, gaug)
    assert np.abs(ff[0] - ffo).max() <= 1e-12 * np.abs(ffo).max() and abs(caug[0] - co) <= 1e-12 * abs(co)


def test_closed_form_baug_blocks():
    """B_theta(i,j) = (0.5+(i-j))dt^2, B_omega(i,j) = dt for j<=i: the structure the CUDA path builds on."""
    robot = M.robotproperty2("M16iB")
    H, dt = 7, 0.5
    _, Baug, _, _ = problem.build_cost_matrices(robot, 5, H, problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0)
    for i in range(H):
        for j in range(H):
            blk = Baug[i * 10:(i + 1) * 10, j * 5:(j + 1) * 5]
            if j <= i:
                assert np.array_equal(blk[:5], (0.5 + (i - j)) * dt * dt * np.eye(5))
                assert np.array_equal(blk[5:], dt * np.eye(5))
            else:
                assert not blk.any()


def test_straight_line_reference_matches_main_fanuc():
    x0 = np.array([0.7825, 0.0284, 0.2172, 0.1444, -1.1779])
    xg = np.array([-0.7825, 0.0284, 0.2172, 0.1444, -1.1779])
    x = problem.straight_line_reference(x0, xg, 30)[0].reshape(30, 10)
    assert np.array_equal(x[-1, :5], xg) and not x[:, 5:].any()
    assert np.allclose(x[0, :5], x0 + (xg - x0) / 30, atol=1e-15)


def test_robotproperty2_constants(oracle):
    for name in ("M16iB", "M200i", "2L"):
        r = M.robotproperty2(name)
        ro = oracle.robot(name)
        nj = 2 if name == "2L" else 5
        for i in range(nj):
            assert np.array_equal(np.array(ro.DH[i][:]), r["DH"][i])
            for k in range(2):
                assert np.array_equal(np.array(ro.cap[i][k][:]), r["cap"][i]["p"][:, k])
        assert np.array_equal(np.array(ro.base[:]), r["base"])
    assert M.robotproperty2("M16iB")["DH"][0, 3] == 1.5708  # literal, not pi/2


def test_synthetic_batch_is_deterministic(oracle):
    a = common.batch_m16ib(oracle, 6, horizon=10)
    b = common.batch_m16ib(oracle, 6, horizon=10)
    for k in ("x0", "ff", "caug", "xref"):
        assert np.array_equal(a[k], b[k])
    r = oracle.robot("M16iB")
    o6 = oracle.obs6(a["obs"][0]["l"])
    for th in np.concatenate([a["theta0"], a["thetag"]]):
        assert oracle.dist_arm(r, th, o6)[0] >= 0.2


def test_bench_reference_arm_contract():
    """bench.py --impl reference: one JSON line with the contract's keys (the oracle port on the host cores; no GPU needed),
    and under a multi-rank launch only rank 0 prints."""
    import json
    import subprocess
    import sys
    bench = os.path.join(ROOT, "bench.py")
    cmd = [sys.executable, bench, "--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "24", "--horizon", "12"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cfs_trajectories_per_sec" and d["unit"] == "trajectories/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert other.returncode == 0 and not [l for l in other.stdout.splitlines() if l.startswith("{")]


def test_bench_helpers_degrade_without_a_gpu():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    import shutil
    if shutil.which("nvidia-smi") is None:
        assert b.bind_to_gpu_numa(0) is None
        smp = b.ClockSampler(0)
        smp.start()
        assert smp.stop()["sm_mhz"] is None
    assert b.cpu_count() >= 1


def test_flat_fixture_round_trip(oracle, tmp_path):
    """the flat little-endian fixture MATLAB / Python / C share (matlab/read_cfs_fixture.m reads it)."""
    from motionplanning_5d_m_b200 import fixture_io
    cfg = common.batch_m16ib(oracle, 6, horizon=10)
    p = str(tmp_path / "batch.bin")
    fixture_io.write_batch_fixture(p, cfg)
    back = fixture_io.read_fixture(p)
    assert np.array_equal(back["QQ"], cfg["sys_info"]["QQ"]) and np.array_equal(back["xref"], cfg["xref"])
    assert np.array_equal(back["ff"], cfg["ff"]) and back["ff"].shape == (6, 50) and float(back["H"]) == 10
    raw = open(p, "rb").read()
    assert raw[:4] == b"CFSB" and raw[12:14] == b"H\0"
    # column-major on disk: the first 8 values of QQ in the file are its first COLUMN
    off = raw.index(b"QQ".ljust(32, b"\0")) + 32 + 4 + 16
    assert np.array_equal(np.frombuffer(raw[off:off + 64], dtype="<f8"), cfg["sys_info"]["QQ"][:8, 0])


def test_cubicpolytraj_against_a_hermite_spline():
    """problem.cubicpolytraj (the toolbox default RRTstar_CFS.m:100 relies on: zero velocity at every waypoint) against scipy's
    cubic Hermite interpolant with zero end slopes on every segment -- an independent construction of the same piecewise cubic."""
    interp = pytest.importorskip("scipy.interpolate")
    from motionplanning_5d_m_b200 import problem
    rng = np.random.default_rng(5)
    for W, T in ((2, 41), (7, 41), (23, 31)):
        wp = rng.standard_normal((5, W))
        tw = np.linspace(0.0, 3.7, W)
        tt = np.linspace(0.0, 3.7, T)
        ours = problem.cubicpolytraj(wp, tw, tt)
        ref = interp.CubicHermiteSpline(tw, wp, np.zeros_like(wp), axis=1)(tt)
        np.testing.assert_allclose(ours, ref, rtol=0, atol=1e-12)
        assert np.array_equal(ours[:, 0], wp[:, 0]) and np.abs(ours[:, -1] - wp[:, -1]).max() < 1e-15
