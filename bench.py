#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: 5-DoF CFS trajectories/sec (M16iB, batch 4096, horizon 50).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU port (oracle/) on host cores

One "step" = one pass of the hot path over one batch: CFS_FANUC.optimizer run to the reference's stop rule for every
problem of a 4096-problem synthetic batch (random start/goal pairs).

  value  whole-job throughput with the inputs resident in HBM (cfs_solve_batch_device).  The K timed steps are issued
         round-robin on --contexts library contexts (each its own stream and buffer set), so the rare heavy-tier
         stragglers of batch k overlap the bulk of batch k+1 -- the GPU analogue of the reference's parfor workers;
         `latency_ms_single_batch` is one batch alone on an idle GPU.
  e2e    the same K steps through the host-pointer C-ABI call (cfs_solve_batch_async / cfs_wait) with pinned host
         buffers: H2D of every input and D2H of every result inside the timed region, copies of batch k+1 overlapping
         the kernels of batch k.
Multi-GPU: every rank solves its own batches (weak scaling, no data-path collective) and after every step the
per-problem (cost, status) are all-gathered over NCCL for best-of selection (motionplanning_5d_m_b200.multi_gpu.best_of,
the GPU analogue of min(routeL), Lib/functions/s_Parallel_rrt.m:27).
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
# One hardware queue per stream: with the default of 8 connections, the copies of one context queue behind the persistent
# kernels of another (false dependencies) and the e2e pipeline loses 40 % (measured: 2.75 -> 1.98 ms per step).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

F_WAYPOINT_NUMJAC = 9680.0   # algorithmic FLOPs of one num_jac gradient (11 dist_arm evaluations), SURVEY.md section 8d
F_WAYPOINT_DERIVEST = 162080.0
METRIC = "cfs_trajectories_per_sec"
FUSED_DRAM_BYTES_PER_LAUNCH = 43.13e6 + 14.30e6  # ncu capture of one 4096 x H=50 batch (profiles/README.md)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--grad", default="numjac", choices=["numjac", "derivest"])
    ap.add_argument("--contexts", type=int, default=24, help="library contexts (streams + buffer sets) the steps rotate over")
    ap.add_argument("--cpu-reps", type=int, default=3, help="CPU baseline: passes of the oracle over the same batch")
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
        if os.environ.get("BENCH_NO_CLOCKS"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def mark(self):
        """start of the timed region: samples taken before it (warm-up) are only used if none fall inside it"""
        self.t_mark = time.perf_counter()

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
        inside = [r[1:] for r in self.rows if r[0] >= getattr(self, "t_mark", 0.0)]
        window = "timed region" if inside else "warm-up + timed region (no sample fell inside the timed region)"
        for r in (inside or [r[1:] for r in self.rows]):
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def bind_to_gpu_numa(index):
    """Multi-rank runs: keep this rank's threads (and hence its first-touch pinned staging buffers) on the CPU cores local to
    its GPU (sysfs local_cpulist of the GPU's PCI function).  Returns a short description for the JSON line, or None."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True, timeout=20).strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "%s: %d local cores" % (bus, len(cpus))
    except Exception:
        return None


def cpu_count():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_text(args):
    return ("batched CFS (BASELINE.json configs[2]): %d random start/goal pairs per GPU, M16iB capsules, horizon %d, "
            "1 obstacle capsule, %s gradients, every problem run to the reference stop rule (eps 0.1, <= 20 outer "
            "iterations)" % (args.batch, args.horizon, "num_jac" if args.grad == "numjac" else "DERIVEST"))


def make_oracle_problem(O, cfg, grad):
    s = cfg["sys_info"]
    return O.Problem(O.robot(cfg["ROBOT"]), s["H"], [o["l"] for o in cfg["obs"]], [o["epsilon"] for o in cfg["obs"]],
                     s["QQ"], s["lim"], s["MAX_input"], s["epsilon_O"], s["MAX_O_ITER"], solver=0, grad=grad)


def oracle_batch(args, O):
    from motionplanning_5d_m_b200 import synthetic
    r = O.robot("M16iB")
    o6 = O.obs6(synthetic.OBS_M16IB["l"])
    feas = lambda cand: np.array([O.dist_arm(r, th, o6)[0] >= synthetic.OBS_M16IB["D"] for th in cand])
    return synthetic.batch_config_m16ib(args.batch, feas, horizon=args.horizon)


def run_reference(args, rank):
    """The reference algorithm on the host CPU.  MATLAB / Octave are not installed and the reference is pure MATLAB
    (nothing to pip-install), so this arm times the C port of the same algorithm (oracle/) on every host core."""
    if rank != 0:
        return
    import oracle as O
    O.build()
    cores = cpu_count()
    cfg = oracle_batch(args, O)
    P = make_oracle_problem(O, cfg, 1 if args.grad == "derivest" else 0)
    # bounded sample per step: the first S problems of the seeded batch, sized so that K steps take about a minute on
    # 16 host cores (the port solves ~1.2 k num_jac / ~0.6 k DERIVEST trajectories per second)
    budget = 65536 if args.grad == "numjac" else 16384
    S = min(args.batch, max(64, budget // max(args.steps, 1)))
    run = lambda: P.solve_batch(cfg["x0"][:S], cfg["ff"][:S], cfg["caug"][:S], cfg["xref"][:S], nthreads=cores)
    for _ in range(min(args.warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = run()
    dt = time.perf_counter() - t0
    val = S * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args), "batch_per_gpu": args.batch, "horizon": args.horizon,
                       "sample_per_step": S},
            "cpu_baseline": {"value": val, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": "%d problems/step of the same seeded batch, %d steps, OpenMP over problems" % (S, args.steps)},
            "e2e": {"value": val, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_per_cfs_iter": 1e3 * dt / args.steps / max(int(out["iters"].max()), 1),
            "note": "MATLAB/Octave unavailable offline and the reference is MATLAB-only: the C port of the reference "
                    "algorithm (oracle/cfs_oracle.c) is timed on all host cores"}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import motionplanning_5d_m_b200 as M
    from motionplanning_5d_m_b200 import _lib, multi_gpu, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else None  # N = 1 keeps every core for the CPU baseline leg
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, H, nj, NC = args.batch, args.horizon, 5, max(1, args.contexts)
    n, N = H * nj, 2 * H * nj
    grad_mode = _lib.GRAD_DERIVEST if args.grad == "derivest" else _lib.GRAD_NUMJAC
    main_stream = torch.cuda.Stream(device=dev)

    # ---- contexts: one stream + one buffer set each ------------------------------------------------------------------
    robot = dict(M.robotproperty2("M16iB"))
    robot["name"] = "M16iB"
    ctxs, streams = [], []
    for c in range(NC):
        ctx = M.Context(local_rank)
        st = torch.cuda.Stream(device=dev)
        ctx.set_stream(st.cuda_stream)
        ctx.set_robot(robot, nj)
        ctx.set_obstacles([synthetic.OBS_M16IB])
        ctxs.append(ctx)
        streams.append(st)
    # ---- synthetic batches (untimed): one seeded batch per buffer set, endpoints rejection-sampled on the GPU -------------
    cfgs = [synthetic.batch_config_m16ib(B, lambda cand: ctxs[0].nodes_feasible(cand)[0], horizon=H,
                                         seed=synthetic.SEED + 1000 * rank + c) for c in range(NC)]
    s = cfgs[0]["sys_info"]
    eps_o, K = float(s["epsilon_O"]), int(s["MAX_O_ITER"])
    for ctx in ctxs:
        ctx.set_cost(H, s["QQ"], s["lim"], s["MAX_input"])
    setup_ms = ctxs[0].stats()["ms_setup"]

    names_in = ("x0", "ff", "caug", "xref")

    def out_set(pin):
        mk = (lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt).pin_memory()) if pin else \
            (lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt, device=dev))
        return dict(u=mk(B, n), x=mk(B, N), cost=mk(B, K), eu=mk(B, K), iters=mk(B, dt=torch.int32),
                    status=mk(B, dt=torch.int32))

    d_in = [{k: torch.from_numpy(cfgs[c][k]).to(dev) for k in names_in} for c in range(NC)]
    d_out = [out_set(False) for _ in range(NC)]
    h_in = [{k: torch.from_numpy(cfgs[c][k]).pin_memory() for k in names_in} for c in range(NC)]
    h_out = [out_set(True) for _ in range(NC)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2
    torch.cuda.synchronize(dev)

    def issue(c, host):
        i, o = (h_in[c], h_out[c]) if host else (d_in[c], d_out[c])
        ctxs[c].solve_batch_ptr(B, i["x0"].data_ptr(), i["ff"].data_ptr(), i["caug"].data_ptr(), i["xref"].data_ptr(),
                                eps_o, K, o["u"].data_ptr(), o["x"].data_ptr(), o["cost"].data_ptr(), o["eu"].data_ptr(),
                                o["iters"].data_ptr(), o["status"].data_ptr(), grad=grad_mode, device=not host, sync=False)

    h_sg = [dict(t0=torch.from_numpy(np.ascontiguousarray(cfgs[c]["theta0"])).pin_memory(),
                 tg=torch.from_numpy(np.ascontiguousarray(cfgs[c]["thetag"])).pin_memory()) for c in range(NC)]

    def issue_sg(c):
        """start/goal pairs in (80 B per problem), problem set-up on the device, u + cost history + iters + status out"""
        o = h_out[c]
        ctxs[c].solve_start_goal_ptr(B, h_sg[c]["t0"].data_ptr(), h_sg[c]["tg"].data_ptr(), eps_o, K, o["u"].data_ptr(), 0,
                                     o["cost"].data_ptr(), 0, o["iters"].data_ptr(), o["status"].data_ptr(), grad=grad_mode,
                                     sync=False)

    def best_of(c):
        """per-step exchange of the multi-GPU job: all-gather (cost, status), argmin over ranks (s_Parallel_rrt.m:27)"""
        with torch.cuda.stream(streams[c]):
            fc = multi_gpu.final_cost(d_out[c]["cost"], d_out[c]["iters"])
            return multi_gpu.best_of(fc, d_out[c]["status"])[0]

    def run_steps(steps, host):
        """issue `steps` batches round-robin over the contexts; returns device ms between the first issue and the last end"""
        t_begin = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_begin.record(main_stream)
        for st in streams:
            st.wait_event(t_begin)
        for k in range(steps):
            c = k % NC
            if host and k >= NC:
                ctxs[c].wait()  # the host buffers of this context are about to be reused
            if host == "sg":
                issue_sg(c)
            else:
                issue(c, host)
            if world > 1 and not host:
                best_of(c)
        for c, st in enumerate(streams):
            e = torch.cuda.Event()
            e.record(st)
            main_stream.wait_event(e)
        t_end.record(main_stream)
        torch.cuda.synchronize(dev)
        for ctx in ctxs:
            ctx.wait()
        return t_begin.elapsed_time(t_end)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up -------------------------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs up to a second to deliver its first row on an 8-GPU box: started before the warm-up
    W = max(args.warmup, 3)
    run_steps(max(W, NC), False)
    run_steps(max(W, NC), True)
    barrier()
    # ---- timed: device-resident -------------------------------------------------------------------------------------------
    flush.zero_()
    barrier()
    sampler.mark()
    ms_dev = run_steps(args.steps, False)
    barrier()
    # algorithmic work of the timed region: every context solved its own batch; count its gradient waypoints
    wp_ctx = [ctxs[c].stats()["grad_waypoints"] for c in range(min(NC, args.steps))]
    wp_timed = sum(wp_ctx[k % NC] for k in range(args.steps))
    # ---- timed: end to end through the host-pointer C ABI -----------------------------------------------------------------
    flush.zero_()
    barrier()
    ms_e2e = run_steps(args.steps, True)
    barrier()
    clocks = sampler.stop()
    e2e_out = {k: h_out[0][k].numpy().copy() for k in ("x", "status", "iters")}  # results of the array-path e2e leg
    tot = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(tot[0]), float(tot[1])

    # ---- one batch alone, per-tier CUDA events inside the library (timing level 2) ---------------------------------------
    ctx0 = ctxs[0]
    ctx0.set_timing(2)
    lat = []
    for _ in range(3):
        flush.zero_()
        torch.cuda.synchronize(dev)
        issue(0, False)
        ctx0.wait()
        lat.append(ctx0.stats())
    st = min(lat, key=lambda d: d["ms_total"])
    ctx0.set_timing(1)
    iters = d_out[0]["iters"].cpu().numpy()
    status = d_out[0]["status"].cpu().numpy()
    fused = st["launches"] <= 6
    fp64_tf, fp64_mhz = ctx0.measure_fp64_peak()
    f_wp = F_WAYPOINT_DERIVEST if args.grad == "derivest" else F_WAYPOINT_NUMJAC
    grad_flops = f_wp * st["grad_waypoints"]
    # stand-alone distance/gradient kernel on the same number of waypoints as the first outer iteration of the batch
    th_all = cfgs[0]["xref"].reshape(B, H, 2 * nj)[:, :, :nj].reshape(-1, nj)
    k1_ms = ctx0.time_dist_grad(th_all, grad=grad_mode, reps=10)
    k1_tf = f_wp * th_all.shape[0] / (k1_ms * 1e-3) / 1e12
    # ---- e2e from start/goal pairs: the mains' problem set-up (main_FANUC.m:38-103) done on the device ---------------------
    from motionplanning_5d_m_b200 import problem
    for ctx in ctxs:
        ctx.set_cost_blocks(H, problem.Q_MAIN_FANUC, problem.R_MAIN_FANUC, 50.0, s["lim"], s["MAX_input"])
    run_steps(NC, "sg")
    barrier()
    ms_sg = run_steps(args.steps, "sg")
    barrier()
    sg_status_equal = bool((h_out[0]["status"].numpy() == e2e_out["status"]).all())
    tot_sg = torch.tensor([ms_sg], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_sg, op=dist.ReduceOp.MAX)
    ms_sg = float(tot_sg[0])
    dom_ms = (st["ms_bulk"] + st["ms_heavy"]) if fused else st["ms_grad"]
    ach_single_tf = grad_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    share = dom_ms / st["ms_total"] if st["ms_total"] else 1.0
    # K launches overlap on NC streams in the timed region: the duration of one launch there is the timed region / K
    # (x the kernel's share of a step, measured on one batch alone and checked against the ncu launch list)
    amort_ms = ms_dev / args.steps * share
    flops_per_launch = f_wp * wp_timed / args.steps
    ach_tf = flops_per_launch / (amort_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_per_problem = 8.0 * ((2 * nj + n + 1 + N) + 3 * n + (n + N + 2 * K) + 1)  # inputs + v0 + outputs
    hbm_ach = bytes_per_problem * B / (st["ms_total"] * 1e-3) / 1e9

    if rank == 0:
        h2d = sum(int(t.numel() * t.element_size()) for t in h_in[0].values())
        d2h = sum(int(t.numel() * t.element_size()) for t in h_out[0].values())
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms_dev * 1e-3), "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args), "batch_per_gpu": B, "horizon": H,
                       "l2": "%d rotating buffer sets (one seeded batch each, %.0f MB of inputs+state+outputs in total) "
                             "> 126 MB L2; 256 MB flush before each timed region" % (NC, NC * B * (bytes_per_problem + 8 * 4 * n) / 1e6),
                       "contexts": NC, "seed": synthetic.SEED, "numa_binding": numa,
                       "parallelism": "independent problems sharded over %d GPU(s), no data-path collective; per-step NCCL "
                                      "all-gather of (cost,status) for best-of" % world},
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "trajectories/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "cfs_solve_batch_async + cfs_wait (host pointers, pinned): H2D + solve + D2H of every step "
                           "inside the timed events, %d contexts in rotation" % NC},
            "e2e_start_goal": {"value": world * B * args.steps / (ms_sg * 1e-3), "unit": "trajectories/s",
                               "h2d_bytes_per_step": int(2 * B * nj * 8),
                               "d2h_bytes_per_step": int(B * (n + K) * 8 + 2 * B * 4), "ms_per_step": ms_sg / args.steps,
                               "status_equal_to_array_path": sg_status_equal,
                               "api": "cfs_set_cost_blocks + cfs_solve_start_goal_async: start/goal pairs in, the mains' set-up "
                                      "(straight-line reference, ff, caug: main_FANUC.m:38-103) built on the device, "
                                      "u + cost history + iters + status out"},
            "gpu_launches": int(st["launches"]) * args.steps * 2 + (int(st["launches"]) + 1) * args.steps,
            "clocks": clocks,
            "latency_ms_single_batch": st["ms_total"],
            "ms_per_cfs_iter": st["ms_total"] / max(int(iters.max()), 1),
            "problem_iters_per_sec": float(st["problem_iters"]) * args.steps / (ms_dev * 1e-3),
            "roofline": {"bound": "fp64", "kernel": "k_cfs_fused (bulk + heavy tier launches)" if fused else "k_grad",
                         "achieved": ach_tf, "peak": fp64_tf, "unit": "TFLOP/s", "frac": ach_tf / fp64_tf if fp64_tf else None,
                         "traffic": FUSED_DRAM_BYTES_PER_LAUNCH if (fused and B == 4096 and H == 50) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the bulk-tier launch, ncu --set "
                                           "full, profiles/r01_prof_fused_v6_raw.csv",
                         "peak_source": "measured here by cfs_measure_fp64_peak (DFMA micro-benchmark; MEASURED_PEAKS.json "
                                        "has no FP64 entry), implied SM clock %.0f MHz" % fp64_mhz,
                         "algorithmic_flops_per_launch": flops_per_launch, "avg_launch_ms": amort_ms,
                         "avg_launch_ms_basis": "timed region / steps x share_of_step: the %d timed launches overlap on %d "
                                                "streams, so this is the GPU time one launch costs in the timed region" % (args.steps, NC),
                         "single_launch": {"ms": dom_ms, "achieved": ach_single_tf,
                                           "frac": ach_single_tf / fp64_tf if fp64_tf else None,
                                           "note": "one batch alone on an idle GPU (bulk + heavy tier, CUDA events inside "
                                                   "the library): the tail of long problems is not overlapped"},
                         "note": "algorithmic FLOPs = %.0f per waypoint gradient x %.0f waypoint gradients per launch "
                                 "(SURVEY.md 8d); the fused kernel also runs the QP, roll-out and stop rule and is bound by "
                                 "dependent-issue latency and L2 latency, not by the FP64 pipe" % (f_wp, wp_timed / args.steps),
                         "share_of_step": share},
            "roofline_k1": {"bound": "fp64", "kernel": "k_grad_%s stand-alone" % args.grad, "achieved": k1_tf, "peak": fp64_tf,
                            "unit": "TFLOP/s", "frac": k1_tf / fp64_tf if fp64_tf else None, "waypoints": int(th_all.shape[0]),
                            "avg_launch_ms": k1_ms},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "note": "%.1f KB algorithmic bytes per problem (inputs + v0 + outputs, once per solve)" % (bytes_per_problem / 1e3)},
            "breakdown_ms": {"single_batch_total": st["ms_total"], "fused_bulk_tier": st["ms_bulk"],
                             "fused_heavy_tier": st["ms_heavy"], "grad_lockstep": st["ms_grad"], "qp_lockstep": st["ms_qp"],
                             "setup_once": setup_ms},
            "solve_stats": {"converged": int(((status & 0xFF) == 0).sum()), "max_iter": int(((status & 0xFF) == 1).sum()),
                            "infeasible": int(((status & 0xFF) == 2).sum()), "numerical": int(((status & 0xFF) == 3).sum()),
                            "mean_iters": float(iters.mean()), "max_iters": int(iters.max()),
                            "qp_steps": int(st["qp_steps"]), "max_working_set": int(st["max_active"])},
        }
        if world == 1:
            # bounded CPU sample of the same workload: the oracle port on all host cores, plus a parity spot check
            import oracle as O
            O.build()
            P = make_oracle_problem(O, cfgs[0], 1 if args.grad == "derivest" else 0)
            cores = cpu_count()
            S = B if args.grad == "numjac" else min(B, 256)
            c0 = cfgs[0]
            t0 = time.perf_counter()
            for _ in range(args.cpu_reps):
                ref = P.solve_batch(c0["x0"][:S], c0["ff"][:S], c0["caug"][:S], c0["xref"][:S], nthreads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": S * args.cpu_reps / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
                                    "sample": "%d passes over %d problems of the same batch, C port of the reference "
                                              "algorithm (oracle/), OpenMP over problems" % (args.cpu_reps, S), "seconds": dt}
            xg, sg, ig = (e2e_out[k][:S] for k in ("x", "status", "iters"))
            ok = ((ref["status"] & 0xFF) < 2) & (ref["status"] == sg)
            dxp = np.abs(xg - ref["x"]).max(axis=1)
            dxp[~ok] = 0.0
            # conditioning probe: the oracle's own answer when ff is perturbed by 1e-12 relative (a problem that does not
            # converge within MAX_O_ITER can amplify that a million-fold; DESIGN.md "parity noise floor")
            pert = P.solve_batch(c0["x0"][:S], c0["ff"][:S] * (1.0 + 1e-12), c0["caug"][:S], c0["xref"][:S], nthreads=cores)
            sens = np.abs(pert["x"] - ref["x"]).max(axis=1)
            well = ok & (sens < 1e-8)
            line["parity_sample"] = {"problems": S, "status_equal": bool((ref["status"] == sg).all()),
                                     "iters_equal": bool((ref["iters"] == ig).all()),
                                     "max_abs_dx": float(dxp.max()) if ok.any() else None,
                                     "max_abs_dx_well_conditioned": float(dxp[well].max()) if well.any() else None,
                                     "ill_conditioned_problems": int((ok & ~well).sum()),
                                     "ill_conditioned_rule": "oracle's own x moves by > 1e-8 when ff is scaled by (1 + 1e-12)",
                                     "problems_over_1e-6": int((dxp > 1e-6).sum())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
