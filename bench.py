#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: 5-DoF CFS trajectories/sec (M16iB, batch 4096, horizon 50).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU port (oracle/) on host cores

One "step" = one pass of the hot path over one batch: CFS_FANUC.optimizer run to the reference's stop rule for every
problem of a 4096-problem synthetic batch (random start/goal pairs).  `value` is measured with
the inputs resident in HBM (cfs_solve_batch_device); `e2e` is the same batch through the host-pointer C-ABI call
(cfs_solve_batch) with pinned host buffers, H2D and D2H inside the timed region.  Multi-GPU: every rank solves its own
batch (weak scaling, no data-path collective) and the per-problem costs are all-gathered over NCCL for best-of
selection (the GPU analogue of min(routeL), Lib/functions/s_Parallel_rrt.m:27).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_WAYPOINT_NUMJAC = 9680.0   # algorithmic FLOPs of one num_jac gradient (11 dist_arm evaluations), BASELINE.md section 3
F_WAYPOINT_DERIVEST = 162080.0
HBM_BYTES_PER_PROBLEM_ITER = 14.4e3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--grad", default="numjac", choices=["numjac", "derivest"])
    ap.add_argument("--cpu-sample", type=int, default=1024, help="problems in the bounded CPU baseline sample")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_count():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_oracle_problem(O, cfg, grad):
    s = cfg["sys_info"]
    return O.Problem(O.robot(cfg["ROBOT"]), s["H"], [o["l"] for o in cfg["obs"]], [o["epsilon"] for o in cfg["obs"]],
                     s["QQ"], s["lim"], s["MAX_input"], s["epsilon_O"], s["MAX_O_ITER"], solver=0, grad=grad)


def run_reference(args, rank, world):
    """The reference algorithm on the host CPU: the oracle port (MATLAB/Octave are not installed)."""
    if rank != 0:
        return
    import oracle as O
    from motionplanning_5d_m_b200 import synthetic
    O.build()
    cores = cpu_count()
    r = O.robot("M16iB")
    o6 = O.obs6(synthetic.OBS_M16IB["l"])
    feas = lambda cand: np.array([O.dist_arm(r, th, o6)[0] >= synthetic.OBS_M16IB["D"] for th in cand])
    S = min(args.batch, 512)
    cfg = synthetic.batch_config_m16ib(S, feas, horizon=args.horizon)
    P = make_oracle_problem(O, cfg, 1 if args.grad == "derivest" else 0)
    run = lambda: P.solve_batch(cfg["x0"], cfg["ff"], cfg["caug"], cfg["xref"], nthreads=cores)
    for _ in range(max(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = run()
    dt = time.perf_counter() - t0
    val = S * args.steps / dt
    line = {"impl": "reference", "metric": "cfs_trajectories_per_sec", "value": val, "unit": "trajectories/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched CFS: random start/goal pairs, M16iB capsules, horizon %d, 1 obstacle, "
                                   "num_jac gradients" % args.horizon if args.grad == "numjac" else
                       "batched CFS, DERIVEST gradients", "batch": args.batch, "horizon": args.horizon,
                       "sample_per_step": S},
            "cpu_baseline": {"value": val, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": "%d problems/step of the same seeded batch, OpenMP over problems" % S},
            "e2e": {"value": val, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_per_cfs_iter": 1e3 * dt / args.steps / max(int(out["iters"].max()), 1),
            "note": "MATLAB/Octave unavailable offline: the C port of the reference algorithm (oracle/) is timed"}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import motionplanning_5d_m_b200 as M
    from motionplanning_5d_m_b200 import _lib, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = M.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    B, H, nj = args.batch, args.horizon, 5
    n, N = H * nj, 2 * H * nj
    grad_mode = _lib.GRAD_DERIVEST if args.grad == "derivest" else _lib.GRAD_NUMJAC

    # ---- synthetic batch (untimed): endpoints rejection-sampled with the GPU feasibility kernel --------------------
    robot = M.robotproperty2("M16iB")
    r = dict(robot)
    r["name"] = "M16iB"
    ctx.set_robot(r, nj)
    ctx.set_obstacles([synthetic.OBS_M16IB])
    cfg = synthetic.batch_config_m16ib(B, lambda cand: ctx.nodes_feasible(cand)[0], horizon=H,
                                       seed=synthetic.SEED + rank)
    s = cfg["sys_info"]
    ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    eps_o, K = float(s["epsilon_O"]), int(s["MAX_O_ITER"])
    setup_ms = ctx.stats()["ms_setup"]

    with torch.cuda.stream(stream):
        d_in = {k: torch.from_numpy(cfg[k]).to(dev) for k in ("x0", "ff", "caug", "xref")}
        d_out = dict(u=torch.empty((B, n), dtype=torch.float64, device=dev),
                     x=torch.empty((B, N), dtype=torch.float64, device=dev),
                     cost=torch.empty((B, K), dtype=torch.float64, device=dev),
                     eu=torch.empty((B, K), dtype=torch.float64, device=dev),
                     iters=torch.empty(B, dtype=torch.int32, device=dev),
                     status=torch.empty(B, dtype=torch.int32, device=dev))
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2
        gathered = torch.empty((world, B, 2), dtype=torch.float64, device=dev) if world > 1 else None
    h_in = {k: torch.from_numpy(cfg[k]).pin_memory() for k in ("x0", "ff", "caug", "xref")}
    h_out = dict(u=torch.empty((B, n), dtype=torch.float64).pin_memory(),
                 x=torch.empty((B, N), dtype=torch.float64).pin_memory(),
                 cost=torch.empty((B, K), dtype=torch.float64).pin_memory(),
                 eu=torch.empty((B, K), dtype=torch.float64).pin_memory(),
                 iters=torch.empty(B, dtype=torch.int32).pin_memory(),
                 status=torch.empty(B, dtype=torch.int32).pin_memory())

    def step_device(sync=False):
        ctx.solve_batch_ptr(B, d_in["x0"].data_ptr(), d_in["ff"].data_ptr(), d_in["caug"].data_ptr(),
                            d_in["xref"].data_ptr(), eps_o, K, d_out["u"].data_ptr(), d_out["x"].data_ptr(),
                            d_out["cost"].data_ptr(), d_out["eu"].data_ptr(), d_out["iters"].data_ptr(),
                            d_out["status"].data_ptr(), grad=grad_mode, device=True, sync=sync)
        if world > 1:
            # best-of selection: all-gather (final cost, status) of every problem, argmin over ranks (s_Parallel_rrt.m:27)
            it = d_out["iters"].clamp(min=1).long() - 1
            last = torch.gather(d_out["cost"], 1, it[:, None])[:, 0]
            mine = torch.stack([last, d_out["status"].double()], dim=1)
            dist.all_gather_into_tensor(gathered.view(-1, 2), mine)
            cost_all = torch.where(gathered[:, :, 1] == 0, gathered[:, :, 0], torch.full_like(gathered[:, :, 0], np.inf))
            return cost_all.argmin(dim=0)
        return None

    def step_host():
        ctx.solve_batch_ptr(B, h_in["x0"].data_ptr(), h_in["ff"].data_ptr(), h_in["caug"].data_ptr(),
                            h_in["xref"].data_ptr(), eps_o, K, h_out["u"].data_ptr(), h_out["x"].data_ptr(),
                            h_out["cost"].data_ptr(), h_out["eu"].data_ptr(), h_out["iters"].data_ptr(),
                            h_out["status"].data_ptr(), grad=grad_mode, device=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(stream):
            for a, b in evs:
                flush.zero_()        # L2 flush between timed iterations (excluded from the timed region)
                a.record(stream)
                fn()
                b.record(stream)
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs]

    # ---- warm-up --------------------------------------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
        step_host()
    barrier()

    # ---- timed: device-resident ----------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    barrier()
    ms_dev = timed(step_device, args.steps)
    barrier()
    # ---- timed: end to end through the host-pointer C ABI -----------------------------------------------------------
    ms_e2e = timed(step_host, args.steps)
    barrier()
    clocks = sampler.stop()

    tot_dev = torch.tensor([sum(ms_dev), sum(ms_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_dev, op=dist.ReduceOp.MAX)
    tot_dev_ms, tot_e2e_ms = float(tot_dev[0]), float(tot_dev[1])

    # ---- instrumented pass: per-kernel CUDA events inside the library -----------------------------------------------
    ctx.set_timing(2)
    with torch.cuda.stream(stream):
        flush.zero_()
        step_device(sync=True)
    st = ctx.stats()
    ctx.set_timing(1)
    iters = d_out["iters"].cpu().numpy()
    status = d_out["status"].cpu().numpy()
    n_grad_launches = int(min(K, (iters + ((status & 0xFF) >= 2)).max()))
    fp64_tf, fp64_mhz = ctx.measure_fp64_peak()
    f_wp = F_WAYPOINT_DERIVEST if args.grad == "derivest" else F_WAYPOINT_NUMJAC
    grad_flops = f_wp * st["grad_waypoints"]
    ach_tf = grad_flops / (st["ms_grad"] * 1e-3) / 1e12 if st["ms_grad"] > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = HBM_BYTES_PER_PROBLEM_ITER * st["problem_iters"] / (st["ms_total"] * 1e-3) / 1e9

    if rank == 0:
        h2d = sum(int(h_in[k].numel() * h_in[k].element_size()) for k in h_in)
        d2h = sum(int(h_out[k].numel() * h_out[k].element_size()) for k in h_out)
        line = {
            "metric": "cfs_trajectories_per_sec", "value": world * B * args.steps / (tot_dev_ms * 1e-3),
            "unit": "trajectories/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": tot_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched CFS (BASELINE.json configs[2]): %d random start/goal pairs per GPU, M16iB "
                                   "capsules, horizon %d, 1 obstacle capsule, %s gradients, run to the reference stop "
                                   "rule (eps 0.1, <=20 outer iterations)" % (B, H, args.grad),
                       "batch_per_gpu": B, "horizon": H, "l2": "256 MB flush between timed steps (outside the timed events)",
                       "seed": synthetic.SEED, "parallelism": "independent problems sharded over %d GPU(s); NCCL "
                       "all-gather of (cost,status) for best-of" % world},
            "e2e": {"value": world * B * args.steps / (tot_e2e_ms * 1e-3), "unit": "trajectories/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": tot_e2e_ms / args.steps,
                    "api": "cfs_solve_batch (host pointers, pinned), H2D+solve+D2H inside the timed events"},
            "gpu_launches": int(st["launches"]) * args.steps,
            "clocks": clocks,
            "ms_per_cfs_iter": tot_dev_ms / args.steps / max(int(iters.max()), 1),
            "problem_iters_per_sec": float(st["problem_iters"]) / (st["ms_total"] * 1e-3),
            "roofline": {"bound": "fp64", "kernel": "k_grad_%s" % args.grad, "achieved": ach_tf, "peak": fp64_tf,
                         "unit": "TFLOP/s", "frac": ach_tf / fp64_tf if fp64_tf else None, "traffic": None,
                         "peak_source": "measured here by cfs_measure_fp64_peak (DFMA micro-benchmark; "
                                        "MEASURED_PEAKS.json has no FP64 entry), implied SM clock %.0f MHz" % fp64_mhz,
                         "algorithmic_flops_per_launch": grad_flops / max(n_grad_launches, 1),
                         "launches": n_grad_launches, "avg_launch_ms": st["ms_grad"] / max(n_grad_launches, 1),
                         "share_of_step": st["ms_grad"] / st["ms_total"] if st["ms_total"] else None},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_ach / hbm_peak, "note": "14.4 KB algorithmic bytes per problem-iteration; the "
                             "path is FP64-bound (SURVEY.md section 8d)"},
            "breakdown_ms": {"total": st["ms_total"], "grad": st["ms_grad"], "qp_rollout": st["ms_qp"],
                             "setup_once": setup_ms},
            "solve_stats": {"converged": int(((status & 0xFF) == 0).sum()), "max_iter": int(((status & 0xFF) == 1).sum()),
                            "infeasible": int(((status & 0xFF) == 2).sum()), "numerical": int(((status & 0xFF) == 3).sum()),
                            "mean_iters": float(iters.mean()), "max_iters": int(iters.max()),
                            "qp_steps": int(st["qp_steps"]), "max_working_set": int(st["max_active"])},
        }
        if world == 1:
            # bounded CPU sample of the same workload: the oracle port on all host cores, and a parity spot check
            import oracle as O
            O.build()
            S = min(args.cpu_sample, B)
            P = make_oracle_problem(O, cfg, 1 if args.grad == "derivest" else 0)
            cores = cpu_count()
            t0 = time.perf_counter()
            ref = P.solve_batch(cfg["x0"][:S], cfg["ff"][:S], cfg["caug"][:S], cfg["xref"][:S], nthreads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": S / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
                                    "sample": "first %d problems of the same batch, C port of the reference algorithm "
                                              "(oracle/), OpenMP over problems" % S, "seconds": dt}
            xg = h_out["x"].numpy()[:S]
            ok = ((ref["status"] & 0xFF) < 2) & (ref["status"] == h_out["status"].numpy()[:S])
            line["parity_sample"] = {"problems": S, "status_equal": bool((ref["status"] == h_out["status"].numpy()[:S]).all()),
                                     "iters_equal": bool((ref["iters"] == h_out["iters"].numpy()[:S]).all()),
                                     "max_abs_dx": float(np.abs(xg[ok] - ref["x"][ok]).max()) if ok.any() else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
