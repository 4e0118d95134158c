#!/usr/bin/env python
"""bench.py -- batched MPC QPs/sec (A1, horizon 10) on N B200s, plus the CPU reference arm.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU solver on the host cores

A "step" is one pass of the hot path (SolveMPCKernel + GetMPCSolution equivalent: state -> condensed
QP -> solve -> forces) over one batch of 65536 synthetic A1 trot instances per GPU (weak scaling: the
per-GPU batch is fixed as N grows; instances are independent, there is no data-path collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched MPC QPs/sec (A1, horizon 10)"
HORIZON, DT_MPC, ROBOT, GAIT = 10, 0.03, "a1", "trot"
BATCH_PER_GPU = 65536
N_INPUT_SETS = 4            # 4 x 49 MB of inputs > 126 MB L2: every step reads inputs not resident in L2
F_ALG_FLOP = 4.38e6         # algorithmic flops per QP at h=10
HBM_ALG_BYTES = 188 * 4 + 48 + 4   # algorithmic HBM bytes per QP: input rows + 12 forces + status
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline, one process per host core (the reference keeps its state in file
# statics).  kind "reference": the reference's OWN qr_mpc_interface.cpp + qpOASES 3.2.0 compiled from
# /root/reference into oracle/_ref/libqr_mpc_ref.so (Eigen replaced by oracle/mini_eigen), driven through
# SetupProblem / SolveMPCKernel / GetMPCSolution.  kind "port" (only if that file is absent): the oracle's
# float32 restatement + the same qpOASES.
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    core, lo, hi, nwsr, seed = args
    try:
        os.sched_setaffinity(0, {core})
    except (AttributeError, OSError):
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import _pkg
    import oracle as O
    pkg = _pkg.load()
    batch = pkg.synth.make_mpc_batch(ROBOT, HORIZON, DT_MPC, hi, seed=seed, gait=GAIT)
    P = O.params_of(batch["robot"], HORIZON, DT_MPC)
    if O.ref_mpc_available() and nwsr == 100:
        sec, lat, _ = O.ref_mpc_time_batch(P, batch, lo, hi)
        return sec, lat.tolist(), -1, "reference"
    sec, lat, capped, _ = O.mpc_time_batch(P, batch, lo, hi, nwsr)
    return sec, lat.tolist(), capped, "port"


def _kind_text(kind):
    if kind == "reference":
        return ("the reference's own qr_mpc_interface.cpp (SetupProblem once, SolveMPCKernel + 12 GetMPCSolution per QP) "
                "and qpOASES 3.2.0 compiled from /root/reference into oracle/_ref (Eigen replaced by oracle/mini_eigen)")
    return ("float32 restatement of qr_mpc_interface.cpp + the reference's qpOASES 3.2.0 compiled from "
            "/root/reference (oracle/_ref)")


def cpu_reference_run(per_core: int, nwsr: int = 100, seed: int = 1234):
    """All host cores, one process each, `per_core` cold SolveMPC-equivalents per process.
    Returns dict(value QPs/s over the wall clock of the slowest process, cores, p50/p99 latency, capped)."""
    import multiprocessing as mp
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = list(range(os.cpu_count() or 1))
    jobs = [(c, i * per_core, (i + 1) * per_core, nwsr, seed) for i, c in enumerate(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(cores)) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall_all = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    lat = np.concatenate([np.asarray(r[1]) for r in res])
    total = per_core * len(cores)
    return dict(value=total / slowest, cores=len(cores), total=total, seconds=slowest, wall_with_setup=wall_all,
                p50_ms=float(np.percentile(lat, 50) * 1e3), p99_ms=float(np.percentile(lat, 99) * 1e3),
                capped=int(sum(r[2] for r in res)), nwsr=nwsr, kind=res[0][3])


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_core = 48
    for _ in range(args.warmup):
        cpu_reference_run(8)
    vals, times = [], []
    last = None
    for _ in range(args.steps):
        last = cpu_reference_run(per_core)
        vals.append(last["total"])
        times.append(last["seconds"])
    value = sum(vals) / sum(times)
    sample = (f"{last['total']} A1 h=10 trot QPs per step ({per_core} per core), cold QProblem + init per QP, "
              f"stock nWSR=100, {_kind_text(last['kind'])}, -O3, one pinned process per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "QP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "A1 convex MPC h=10 dt=0.03 trot (reference CPU solver, bounded sample per step)"},
        "cpu_baseline": {"value": value, "unit": "QP/s", "cores": last["cores"], "kind": last["kind"], "sample": sample,
                         "p50_ms": last["p50_ms"], "p99_ms": last["p99_ms"]},
        "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def _time_ms(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_wbc_and_full_step(pkg, capi, torch, stream):
    """Reported beside the headline metric: qr_wbc_kernel throughput (Lite3, batches 1024 and 65536) and one full
    control tick of BASELINE configs[1] (Lite3 trot: contact table + reference trajectory -> MPC -> leg torques,
    swing-foot parabola, WBIC with the MPC forces as Fr_des), batch 1024, all resident on the device."""
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    robot = pkg.robots.ROBOTS["lite3"]
    M = capi.wbc_model_of(robot)
    out = {"wbc": {"kernel": "qr_wbc_kernel", "robot": "lite3", "unit": "robots/s", "arith": "float64, one warp per robot"}}
    for B in (1024, 65536):
        wb = pkg.synth.make_wbc_batch("lite3", B, seed=6)
        state, cmd, contact = dev(wb["state"]), dev(wb["cmd"]), dev(wb["contact"])
        tau = torch.empty((B, 12), device="cuda")
        st = torch.empty(B, dtype=torch.int32, device="cuda")
        ms = _time_ms(torch, lambda: capi.wbc_solve_batch_device(M, state, cmd, contact, tau, stream, status=st), 10)
        out["wbc"][f"batch_{B}"] = {"value": B / ms * 1e3, "ms_per_step": ms, "status_nonzero": int((st != 0).sum())}
    # full tick, batch 1024
    B, h, dt = 1024, 10, 0.03
    mb = pkg.synth.make_mpc_batch("lite3", h, dt, B, seed=11, gait="trot")
    wb = pkg.synth.make_wbc_batch("lite3", B, seed=12)
    P = capi.params_of(robot, h, dt)
    gt = pkg.robots.GAITS["trot"]
    rng = np.random.default_rng(13)
    progress = np.mod(rng.uniform(0, 1, (B, 1)) + np.array(gt["offsets"])[None, :], 1.0).astype(np.float32)
    duty = np.full((B, 4), gt["duty"], np.float32)
    traj_init = np.zeros((B, 12), np.float32)
    traj_init[:, 2] = mb["rpy"][:, 2]; traj_init[:, 3:5] = mb["p"][:, :2]; traj_init[:, 5] = robot.body_height
    traj_init[:, 9] = 0.5
    d = {k: dev(mb[k]) for k in KEYS}
    d_prog, d_duty, d_init, d_xy = dev(progress), dev(duty), dev(traj_init), dev(mb["p"][:, :2])
    o = dict(grf=torch.empty((B, 12), device="cuda"), status=torch.empty(B, dtype=torch.int32, device="cuda"),
             iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
    state, cmd, contact = dev(wb["state"]), dev(wb["cmd"]), dev(wb["contact"])
    q = state[:, 13:25].contiguous()
    tau_mpc = torch.empty((B, 12), device="cuda")
    tau = torch.empty((B, 12), device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    sw_start, sw_end = dev(rng.uniform(-0.1, 0.1, (4 * B, 3)).astype(np.float32)), dev(rng.uniform(-0.1, 0.1, (4 * B, 3)).astype(np.float32))
    sw_h, sw_ph = dev(np.full(4 * B, 0.08, np.float32)), dev(rng.uniform(0, 1, 4 * B).astype(np.float32))
    sw_pos = torch.empty((4 * B, 3), device="cuda")
    nhl = pkg.synth.num_horizon_l(gt)

    def tick():
        capi.mpc_inputs_batch_device(h, nhl, dt, d_prog, d_duty, None, None, d_init, d_xy, d["gait"], d["traj"], stream)
        capi.mpc_solve_batch_device(P, d, o, stream)
        capi.mpc_leg_torque_batch_device(robot, d["quat"], q, o["grf"], None, tau_mpc, stream)
        capi.swing_parabola_batch_device(sw_start, sw_end, sw_h, sw_ph, False, sw_pos, None, stream)
        cmd[:, 51:63] = o["grf"]   # Fr_des of qrWbcCtrlData <- MPC forces (a torch copy kernel, not one of ours)
        capi.wbc_solve_batch_device(M, state, cmd, contact, tau, stream, status=st)

    ms = _time_ms(torch, tick, 20)
    out["full_step"] = {"workload": "Lite3 trot h=10 dt=0.03: contact table + reference trajectory -> MPC -> leg torques, "
                                    "swing parabola, WBIC (BASELINE configs[1]), batch 1024 on the device",
                        "value": B / ms * 1e3, "unit": "robot ticks/s", "ms_per_step": ms,
                        "mpc_not_converged": int((o["status"] != 0).sum()), "wbc_status_nonzero": int((st != 0).sum())}
    # BASELINE configs[3]: Aliengo, gait drawn per instance (trot / walk / gallop: different contact masks and numbers of
    # eliminated swing variables -> several size classes in one call), batch 16384
    B, h, dt = 16384, 10, 0.03
    mb = pkg.synth.make_mpc_batch("aliengo", h, dt, B, seed=15, gait="mixed")
    Pm = capi.params_of(pkg.robots.ROBOTS["aliengo"], h, dt)
    dm = {k: dev(mb[k]) for k in KEYS}
    om = dict(grf=torch.empty((B, 12), device="cuda"), status=torch.empty(B, dtype=torch.int32, device="cuda"),
              iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
    ms = _time_ms(torch, lambda: capi.mpc_solve_batch_device(Pm, dm, om, stream), 5, warm=2)
    nf = (mb["gait"] > 0).sum(1)
    out["mixed_gait"] = {"workload": "Aliengo convex MPC h=10, gait drawn per instance (trot / walk / gallop), batch 16384 "
                                     "(BASELINE configs[3])", "value": B / ms * 1e3, "unit": "QP/s", "ms_per_step": ms,
                         "not_converged": int((om["status"] != 0).sum()),
                         "stance_footsteps_min_mean_max": [int(nf.min()), float(nf.mean()), int(nf.max())],
                         "rounds_mean": float(om["iters"][:, 1].float().mean())}
    # BASELINE configs[4]: horizon-30 long-preview MPC (360 variables), batch 4096, A1 trot
    B, h, dt = 4096, 30, 0.03
    mb = pkg.synth.make_mpc_batch("a1", h, dt, B, seed=14, gait="trot")
    P30 = capi.params_of(pkg.robots.ROBOTS["a1"], h, dt)
    d30 = {k: dev(mb[k]) for k in KEYS}
    o30 = dict(grf=torch.empty((B, 12), device="cuda"), status=torch.empty(B, dtype=torch.int32, device="cuda"),
               iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
    ms = _time_ms(torch, lambda: capi.mpc_solve_batch_device(P30, d30, o30, stream), 3, warm=1)
    out["horizon30"] = {"workload": "A1 trot convex MPC h=30 dt=0.03 (360 variables, 72 stance foot-steps), batch 4096 "
                                    "(BASELINE configs[4]); beyond the reference's own K_MAX_GAIT_SEGMENTS = 16",
                        "value": B / ms * 1e3, "unit": "QP/s", "ms_per_step": ms,
                        "not_converged": int((o30["status"] != 0).sum()),
                        "rounds_mean": float(o30["iters"][:, 1].float().mean())}
    return out


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import _pkg
    pkg = _pkg.load()
    from quadruped_robot_b200 import build as qbuild, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.init(local_rank)
    B = args.batch
    h = HORIZON
    robot = pkg.robots.ROBOTS[ROBOT]
    P = capi.params_of(robot, h, DT_MPC)

    # synthetic inputs: N_INPUT_SETS different batches per rank, resident in HBM before timing
    sets_host = [pkg.synth.make_mpc_batch(ROBOT, h, DT_MPC, B, seed=1000 * rank + s, gait=GAIT) for s in range(N_INPUT_SETS)]
    sets_dev = [{k: torch.from_numpy(b[k]).cuda() for k in KEYS} for b in sets_host]
    out = dict(grf=torch.empty((B, 12), device="cuda"), status=torch.empty(B, dtype=torch.int32, device="cuda"),
               iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        capi.mpc_solve_batch_device(P, sets_dev[i % N_INPUT_SETS], out, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(args.warmup + i)
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    status = out["status"].cpu().numpy()
    iters = out["iters"].cpu().numpy()
    n_bad = int((status != 0).sum())

    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory, H2D + D2H timed)
    pinned = []
    for b in sets_host:
        pb = {}
        for k in KEYS:
            tns = torch.from_numpy(b[k]).pin_memory()
            pb[k] = tns.numpy()
            pb["_keep_" + k] = tns
        pinned.append(pb)
    grf_pin = torch.empty((B, 12), dtype=torch.float32).pin_memory()
    st_pin = torch.empty(B, dtype=torch.int32).pin_memory()
    lib = capi.lib()

    def e2e_step(i):
        pb = pinned[i % N_INPUT_SETS]
        rc = lib.qr_gpu_mpc_solve_batch_host(C.byref(P), None, B, *[C.c_void_p(pb[k].ctypes.data) for k in KEYS],
                                             None, None, C.c_void_p(grf_pin.data_ptr()), None,
                                             C.c_void_p(st_pin.data_ptr()), None)
        if rc != 0:
            raise RuntimeError(lib.qr_gpu_last_error().decode())

    for i in range(max(1, args.warmup)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d = B * sum(sets_host[0][k].shape[1] for k in KEYS) * 4
    d2h = B * (12 * 4 + 4)

    # ---- batch-1 latency (rank 0): p50 / p99 of a synchronous host-API call
    lat = None
    if rank == 0:
        # 128 DIFFERENT instances, visited in turn: the number of active-set rounds -- and with it the solve time --
        # varies from instance to instance, so a percentile over repetitions of one instance would hide most of the spread
        n_inst = 128
        ones = [{k: np.ascontiguousarray(sets_host[0][k][i:i + 1]) for k in KEYS} for i in range(n_inst)]
        ptrs = [[C.c_void_p(o[k].ctypes.data) for k in KEYS] for o in ones]
        g1 = np.empty((1, 12), np.float32)
        ts = []
        for i in range(128 + 1024):
            a = time.perf_counter()
            lib.qr_gpu_mpc_solve_batch_host(C.byref(P), None, 1, *ptrs[i % n_inst],
                                            None, None, C.c_void_p(g1.ctypes.data), None, None, None)
            ts.append(time.perf_counter() - a)
        ts = np.asarray(ts[128:]) * 1e6
        lat = {"batch": 1, "p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
               "mean_us": float(ts.mean()), "reps": 1024, "instances": n_inst}

    # ---- the other half of the hot path (rank 0, N = 1): the WBC kernel alone and BASELINE configs[1], one full
    # MPC + WBIC tick for a batch of robots, everything on the device
    extra = None
    if rank == 0 and world == 1:
        extra = bench_wbc_and_full_step(pkg, capi, torch, stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the one kernel in the step (FMA pipe: SURVEY.md section 8d), HBM figures beside it
    kern_s = float(np.mean(kern_ms)) * 1e-3
    peaks = {}
    try:
        pl = C.CDLL(qbuild.PEAKS_LIB)
        f64, f32 = C.c_double(), C.c_double()
        if pl.qr_peak_fma(C.byref(f64), C.byref(f32)) == 0:
            peaks = {"fp64_fma_tflops": f64.value, "fp32_fma_tflops": f32.value}
    except OSError:
        pass
    mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(mp_path):
        try:
            hbm_peak, hbm_src = float(json.load(open(mp_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (KeyError, ValueError):
            pass
    achieved_tf = B * F_ALG_FLOP / kern_s / 1e12
    fp64_peak = peaks.get("fp64_fma_tflops")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except ValueError:
            pass
    roofline = {
        "bound": "fp64_fma", "kernel": "qr_mpc_fused_kernel", "achieved": achieved_tf, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": (achieved_tf / fp64_peak) if fp64_peak else None, "traffic": traffic,
        "peak_source": "FP64 DFMA chain microbenchmark run in this process (csrc/peaks.cu); not in MEASURED_PEAKS.json",
        "algorithmic_flop_per_qp": F_ALG_FLOP, "kernel_ms": kern_s * 1e3, "fp32_fma_peak_tflops": peaks.get("fp32_fma_tflops"),
        "hbm": {"achieved": B * HBM_ALG_BYTES / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": B * HBM_ALG_BYTES / kern_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_qp": HBM_ALG_BYTES},
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        r = cpu_reference_run(96)
        cpu = {"value": r["value"], "unit": "QP/s", "cores": r["cores"], "kind": r["kind"],
               "sample": (f"{r['total']} A1 h=10 trot QPs ({r['seconds']:.1f} s, 96 per core), cold QProblem + init per QP, "
                          f"stock nWSR=100, {_kind_text(r['kind'])}, one pinned process per core"),
               "p50_ms": r["p50_ms"], "p99_ms": r["p99_ms"]}

    n_size_classes = (4 * h + 7) // 8
    occ = capi.occupancy(h, int(round(float((sets_host[0]['gait'] > 0).sum(axis=1).mean()))))
    line = {
        "metric": METRIC, "value": value, "unit": "QP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"A1 convex MPC h={h} dt={DT_MPC} trot, batch {B} per GPU (BASELINE configs[2] shape)",
                   "global_batch": world * B, "parallelism": f"dp{world} (independent instances, no collective on the solve path)",
                   "l2": f"{N_INPUT_SETS} rotating input sets ({N_INPUT_SETS * h2d / 1e6:.0f} MB) > 126 MB L2",
                   "arith": "float32 condensing (reference operation order) + float64 block active-set iteration (coarse move-blocked prediction, then full size; ends on verified KKT conditions), interior-point fallback",
                   "launch": occ, "not_converged": n_bad,
                   "ipm_iters_mean": float(iters[:, 0].mean()), "polish_rounds_mean": float(iters[:, 1].mean())},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "qr_gpu_mpc_solve_batch_host (pinned host buffers; per step H2D + kernels + D2H + stream sync, the batch cut into two chunks on two streams so that most of the upload runs under the kernels)"},
        "gpu_launches": args.steps * (1 + n_size_classes),   # per step: qr_mpc_classify_kernel + one qr_mpc_fused_kernel per size class
        "roofline": roofline,
        "cpu_baseline": cpu,
        "latency": lat,
    }
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="instances per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
