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
F_ALG_FLOP_H30 = 117.2e6    # the same count at h=30
WBC_F_ALG_FLOP = 0.3e6      # algorithmic flops per WBC tick
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
def _preload_reference_libs():
    """dlopen the reference-compiled libraries in THIS process before the per-core workers are forked, so that the
    driver's loaded-library record of the bench process shows what the CPU arm runs (the workers inherit the mapping)."""
    loaded = []
    for name in ("libqr_mpc_ref.so", "libqr_wbc_ref.so"):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(path):
            try:
                C.CDLL(path)
                loaded.append(os.path.join("oracle", "_ref", name))
            except OSError:
                pass
    return loaded


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
    _preload_reference_libs()
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


def _cpu_leg_worker(args):
    """One pinned worker of a CPU leg: kind in {"wbc", "tick", "mixed", "h30"}; returns (seconds, count, label)."""
    kind, core, per_core, seed = args
    try:
        os.sched_setaffinity(0, {core})
    except (AttributeError, OSError):
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import _pkg
    import oracle as O
    pkg = _pkg.load()
    if kind == "wbc":
        wb = pkg.synth.make_wbc_batch("lite3", per_core, seed=seed + core)
        sec, _, _ = O.ref_wbc_time_batch(O.wbc_model_of(wb["robot"]), wb["state"], wb["cmd"], wb["contact"])
        return sec, per_core, "reference"
    if kind == "tick":
        # one control tick per robot: SolveMPCKernel + GetMPCSolution, then the WBC tick fed with the MPC forces
        mb = pkg.synth.make_mpc_batch("lite3", 10, 0.03, per_core, seed=seed + core, gait="trot")
        wb = pkg.synth.make_wbc_batch("lite3", per_core, seed=seed + 100 + core)
        P = O.params_of(mb["robot"], 10, 0.03)
        sec, _, x12 = O.ref_mpc_time_batch(P, mb, 0, per_core, want_x=True)
        wb["cmd"][:, 51:63] = x12.astype(np.float32)
        sec2, _, _ = O.ref_wbc_time_batch(O.wbc_model_of(wb["robot"]), wb["state"], wb["cmd"], wb["contact"])
        return sec + sec2, per_core, "reference"
    if kind == "mixed":
        mb = pkg.synth.make_mpc_batch("aliengo", 10, 0.03, per_core, seed=seed + core, gait="mixed")
        P = O.params_of(mb["robot"], 10, 0.03)
        sec, _, _ = O.ref_mpc_time_batch(P, mb, 0, per_core)
        return sec, per_core, "reference"
    if kind == "h30":
        # beyond the reference's own K_MAX_GAIT_SEGMENTS = 16: the oracle's restatement + the reference's qpOASES, converged
        mb = pkg.synth.make_mpc_batch("a1", 30, 0.03, per_core, seed=seed + core, gait="trot")
        P = O.params_of(mb["robot"], 30, 0.03)
        sec, _, _, _ = O.mpc_time_batch(P, mb, 0, per_core, 100000)
        return sec, per_core, "port"
    raise ValueError(kind)


def cpu_leg(kind: str, per_core: int, seed: int = 4321):
    """A bounded CPU sample of one of the side workloads on all host cores (one pinned process per core)."""
    import multiprocessing as mp
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = list(range(os.cpu_count() or 1))
    _preload_reference_libs()
    with mp.get_context("fork").Pool(len(cores)) as pool:
        res = pool.map(_cpu_leg_worker, [(kind, c, per_core, seed) for c in cores])
    slowest = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return {"value": total / slowest, "cores": len(cores), "kind": res[0][2],
            "sample": f"{total} units ({per_core} per core, {slowest:.2f} s on the slowest core), one pinned process per core",
            "eigen": "oracle/mini_eigen (plain loops, not vectorised) in place of Eigen"}


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
                         "p50_ms": last["p50_ms"], "p99_ms": last["p99_ms"],
                         "eigen": "oracle/mini_eigen (plain loops, not vectorised) in place of Eigen: the condensing part "
                                  "(about 5 % of a solve) runs slower than a real Eigen build would",
                         "native_libs": _preload_reference_libs()},
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


def bench_wbc_and_full_step(pkg, capi, torch, stream, with_cpu=False, fp64_peak=None):
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
    # roofline of qr_wbc_kernel: FP64 FMA pipe, algorithmic 0.3 MFLOP per robot, batch 65536
    wbc_tf = out["wbc"]["batch_65536"]["value"] * WBC_F_ALG_FLOP / 1e12
    out["wbc"]["roofline"] = {"bound": "fp64_fma", "kernel": "qr_wbc_kernel", "achieved": wbc_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (wbc_tf / fp64_peak) if fp64_peak else None, "algorithmic_flop_per_robot": WBC_F_ALG_FLOP,
                              "hbm_algorithmic_bytes_per_robot": (37 + 66 + 4 + 12 + 1) * 4}
    if with_cpu:
        c = cpu_leg("wbc", 2048)
        c["unit"] = "robots/s"
        c["what"] = ("the reference's own FloatingBaseModel / qrSingleContact / task_set / qrMultitaskProjection / "
                     "qrWholeBodyImpulseCtrl + QuadProg++ compiled from /root/reference (oracle/_ref/libqr_wbc_ref.so), "
                     "controller objects built once, one recomputing tick per robot (qr_wbc_locomotion_controller.cpp:108-134)")
        out["wbc"]["cpu_baseline"] = c
    # full tick, batch 1024: every stage through the C ABI, nothing touched by the host in between
    B, h, dt = 1024, 10, 0.03
    mb = pkg.synth.make_mpc_batch("lite3", h, dt, B, seed=11, gait="trot")
    wb = pkg.synth.make_wbc_batch("lite3", B, seed=12)
    fh = pkg.synth.make_foothold_batch("lite3", B, seed=16)
    P = capi.params_of(robot, h, dt)
    G = capi.leg_geometry_of(robot)
    fP = capi.foothold_params_of(fh["params"])
    gt = pkg.robots.GAITS["trot"]
    rng = np.random.default_rng(13)
    F32 = np.float32
    duty, stance = F32(gt["duty"]), F32(gt["stance_duration"])
    period = F32(stance / duty)
    cfg = np.zeros((B, 4, 5), F32)
    cfg[:, :, 0] = np.array(gt["offsets"], F32)
    cfg[:, :, 1], cfg[:, :, 2], cfg[:, :, 3], cfg[:, :, 4] = period, duty, period - stance, duty
    istate = np.ones((B, 20), np.int32); istate[:, 16:] = 0
    fstate = np.zeros((B, 4), F32); fstate[:, 3] = 1.0
    traj_init = np.zeros((B, 12), F32)
    traj_init[:, 2] = mb["rpy"][:, 2]; traj_init[:, 3:5] = mb["p"][:, :2]; traj_init[:, 5] = robot.body_height
    traj_init[:, 9] = 0.5
    state_h = wb["state"].copy()
    state_h[:, :4], state_h[:, 4:7] = mb["quat"], mb["p"]
    d = {k: dev(mb[k]) for k in KEYS}
    d_time = dev(rng.uniform(0, 2, B).astype(F32))
    d_cfg, d_i, d_f = dev(cfg.reshape(B, 20)), dev(istate), dev(fstate)
    d_pf, d_np, d_sr = (torch.zeros((B, 4), device="cuda") for _ in range(3))
    d_allow, d_early, d_mask, d_stance = (torch.empty((B, 4), dtype=torch.int32, device="cuda") for _ in range(4))
    d_stance.fill_(0)   # measured contacts of a tick = the planned stance legs of the tick before (no early touch-downs)
    d_duty, d_init, d_xy = dev(np.full((B, 4), duty, F32)), dev(traj_init), dev(mb["p"][:, :2])
    o = dict(grf=torch.empty((B, 12), device="cuda"), status=torch.empty(B, dtype=torch.int32, device="cuda"),
             iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
    state, cmd = dev(state_h), dev(wb["cmd"])
    q = state[:, 13:25].contiguous()
    foot_base = torch.empty((B, 12), device="cuda")
    tau_mpc, ff = torch.empty((B, 12), device="cuda"), torch.empty((B, 12), device="cuda")
    tau = torch.empty((B, 12), device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    fd = {k: dev(v) for k, v in fh.items() if isinstance(v, np.ndarray)}
    fd["swing_remain"], fd["norm_phase"], fd["allow_switch"], fd["swing_mask"], fd["foot_base"] = d_sr, d_np, d_allow, d_mask, foot_base
    foothold, planner_phase = torch.zeros((B, 12), device="cuda"), torch.zeros((B, 4), device="cuda")
    switch_pos = dev((np.array(robot.hip_positions)[None] + np.array([0, 0, -robot.body_height]) + rng.uniform(-0.06, 0.06, (B, 4, 3))).astype(F32).reshape(B, 12))
    swing_dur = dev(cfg[:, :, 3].copy())
    nhl = pkg.synth.num_horizon_l(gt)
    TICK_KERNELS = 9 + (4 * h + 7) // 8   # gait, table/trajectory, FK, lever arms, classify + fused classes, foothold, swing targets, WBC

    def tick():
        capi.gait_update_batch_device(d_time, d_cfg, 0.1, d_stance, None, False, d_i, d_f, d_pf, d_np, d_sr, stream,
                                      allow=d_allow, early=d_early, swing_mask=d_mask, stance_mask=d_stance)
        capi.mpc_inputs_batch_device(h, nhl, dt, d_pf, d_duty, d_early, d_stance, d_init, d_xy, d["gait"], d["traj"], stream)
        capi.leg_kinematics_batch_device(G, q, None, foot_base, None, None, stream)
        capi.mpc_lever_arms_batch_device(robot, d["quat"], foot_base, d["r_feet"], stream)
        capi.mpc_solve_batch_device_ex(P, d, o, stream, robot, q=q, f_ff=ff, tau=tau_mpc, wbc_cmd=cmd)   # epilogue: f_ff, tau, Fr_des
        capi.foothold_heuristic_batch_device(fP, fd, foothold, planner_phase, stream)
        capi.swing_targets_batch_device(G, d["p"], d["quat"], d["v"], foothold, planner_phase, switch_pos, swing_dur, d_mask, True, cmd, stream)
        capi.wbc_solve_batch_device(M, state, cmd, d_stance, tau, stream, status=st)

    ms = _time_ms(torch, tick, 20)
    out["full_step"] = {"workload": "Lite3 trot h=10 dt=0.03, one control tick: gait phase -> contact table + reference trajectory -> leg FK -> "
                                    "lever arms -> MPC with the fused leg-force / torque / Fr_des epilogue -> foothold -> swing targets -> WBIC "
                                    "(BASELINE configs[1]), batch 1024 on the device, no host work between the stages",
                        "kernels_per_tick": TICK_KERNELS,
                        "value": B / ms * 1e3, "unit": "robot ticks/s", "ms_per_step": ms,
                        "mpc_not_converged": int((o["status"] != 0).sum()), "wbc_status_nonzero": int((st != 0).sum())}
    if with_cpu:
        c = cpu_leg("tick", 32)
        c["unit"] = "robot ticks/s"
        c["what"] = "reference MPC source build (SolveMPCKernel + GetMPCSolution, stock nWSR=100) followed by the reference WBC source build, per robot"
        out["full_step"]["cpu_baseline"] = c
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
    if with_cpu:
        c = cpu_leg("mixed", 48)
        c["unit"] = "QP/s"
        c["what"] = "reference MPC source build on Aliengo mixed-gait instances, stock nWSR=100"
        out["mixed_gait"]["cpu_baseline"] = c
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
                        "rounds_mean": float(o30["iters"][:, 1].float().mean()),
                        "factorisation": "reduced systems of >= 16 block columns: blocked Cholesky on 8x8 tiles with "
                                         "mma.sync.m8n8k4.f64 (csrc/chol8.h); smaller ones: scalar 3x3-block LDL'"}
    tf30 = out["horizon30"]["value"] * F_ALG_FLOP_H30 / 1e12
    out["horizon30"]["roofline"] = {"bound": "fp64_fma", "achieved": tf30, "peak": fp64_peak, "unit": "TFLOP/s",
                                    "frac": (tf30 / fp64_peak) if fp64_peak else None, "algorithmic_flop_per_qp": F_ALG_FLOP_H30}
    if with_cpu:
        c = cpu_leg("h30", 3)
        c["unit"] = "QP/s"
        c["what"] = ("h = 30 is beyond the reference's own arrays (K_MAX_GAIT_SEGMENTS = 16): the oracle's float32 restatement of "
                     "qr_mpc_interface.cpp + the reference's qpOASES 3.2.0 run to convergence")
        out["horizon30"]["cpu_baseline"] = c
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

    # ---- strong scaling (BASELINE configs[2]: "batch 65536 sharded across 8 B200"): the SAME 65536 instances cut into
    # contiguous shards, one per GPU, and the forces gathered.  Under torchrun every rank solves its shard on the device
    # and one ncclAllGather of 12*B/G floats per rank assembles the result on every GPU (inside the timed region).
    # In a single process that sees several GPUs the library's own qr_gpu_mpc_solve_batch_host_multi is timed instead
    # (host rows in, one host thread per GPU, results gathered into one host array).
    strong = None
    if world > 1:
        Bs = BATCH_PER_GPU // world
        full = pkg.synth.make_mpc_batch(ROBOT, h, DT_MPC, BATCH_PER_GPU, seed=777, gait=GAIT)
        shard = {k: torch.from_numpy(np.ascontiguousarray(full[k][rank * Bs:(rank + 1) * Bs])).cuda() for k in KEYS}
        o_s = dict(grf=torch.empty((Bs, 12), device="cuda"), status=torch.empty(Bs, dtype=torch.int32, device="cuda"))
        gathered = torch.empty((world * Bs, 12), device="cuda")

        def strong_step():
            capi.mpc_solve_batch_device(P, shard, o_s, stream)
            dist.all_gather_into_tensor(gathered, o_s["grf"])

        for _ in range(3):
            strong_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            strong_step()
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms = float(ts.item()) / args.steps
        strong = {"workload": f"A1 h={h} trot, {world * Bs} instances in total, {Bs} per GPU, forces all-gathered (NCCL) every step",
                  "value": world * Bs / ms * 1e3, "unit": "QP/s", "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                  "not_converged": int((o_s["status"] != 0).sum())}
    elif torch.cuda.device_count() > 1:
        G = torch.cuda.device_count()
        pb = pinned[0]
        host_rows = {k: pb[k] for k in KEYS}
        res = {"grf": grf_pin.numpy(), "status": st_pin.numpy()}
        for _ in range(2):
            capi.mpc_solve_batch_host_multi(P, host_rows, list(range(G)), want_info=True, out=res)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            capi.mpc_solve_batch_host_multi(P, host_rows, list(range(G)), want_info=True, out=res)
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        torch.cuda.set_device(local_rank)
        strong = {"workload": f"A1 h={h} trot, {B} instances in one host batch sharded over {G} GPUs of this process by "
                              "qr_gpu_mpc_solve_batch_host_multi (pinned host rows in, forces gathered into one host array)",
                  "value": B / ms * 1e3, "unit": "QP/s", "ms_per_step": ms, "n_gpus": G, "scaling": "strong",
                  "not_converged": int((res["status"] != 0).sum())}

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
    peaks = {}
    if rank == 0:
        try:
            pl = C.CDLL(qbuild.PEAKS_LIB)
            f64, f32 = C.c_double(), C.c_double()
            if pl.qr_peak_fma(C.byref(f64), C.byref(f32)) == 0:
                peaks = {"fp64_fma_tflops": f64.value, "fp32_fma_tflops": f32.value}
        except OSError:
            pass
    extra = None
    if rank == 0 and world == 1:
        extra = bench_wbc_and_full_step(pkg, capi, torch, stream, with_cpu=not args.no_cpu_baseline,
                                        fp64_peak=peaks.get("fp64_fma_tflops"))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the one kernel in the step (FMA pipe: SURVEY.md section 8d), HBM figures beside it
    kern_s = float(np.mean(kern_ms)) * 1e-3
    mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(mp_path):
        try:
            hbm_peak, hbm_src = float(json.load(open(mp_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (KeyError, ValueError):
            pass
    achieved_tf = B * F_ALG_FLOP / kern_s / 1e12
    fp64_peak = peaks.get("fp64_fma_tflops")
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except ValueError:
            pass
    roofline = {
        "bound": "fp64_fma", "kernel": "qr_mpc_fused_kernel", "achieved": achieved_tf, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": (achieved_tf / fp64_peak) if fp64_peak else None, "traffic": traffic,
        "traffic_source": traffic_src,
        "peak_source": "FP64 DFMA chain microbenchmark run in this process (csrc/peaks.cu); not in MEASURED_PEAKS.json",
        "algorithmic_flop_per_qp": F_ALG_FLOP, "kernel_ms": kern_s * 1e3, "fp32_fma_peak_tflops": peaks.get("fp32_fma_tflops"),
        "hbm": {"achieved": B * HBM_ALG_BYTES / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": B * HBM_ALG_BYTES / kern_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_qp": HBM_ALG_BYTES},
    }

    cpu = None
    parity_rec = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        r = cpu_reference_run(96)
        cpu = {"value": r["value"], "unit": "QP/s", "cores": r["cores"], "kind": r["kind"],
               "sample": (f"{r['total']} A1 h=10 trot QPs ({r['seconds']:.1f} s, 96 per core), cold QProblem + init per QP, "
                          f"stock nWSR=100, {_kind_text(r['kind'])}, one pinned process per core"),
               "p50_ms": r["p50_ms"], "p99_ms": r["p99_ms"],
               "eigen": "oracle/mini_eigen (plain loops, not vectorised) in place of Eigen: the condensing part (about 5 % "
                        "of a solve) runs slower than a real Eigen build would"}
        # parity of this run's GPU forces on a sample of the timed workload, checked by the same CPU leg: element-wise
        # against the certified exact optimum of the reference's QP (converged qpOASES + extended-precision KKT solve)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import bulk
        n_par = 1024
        sub = {k: (np.ascontiguousarray(v[:n_par]) if isinstance(v, np.ndarray) and v.shape[:1] == (B,) else v)
               for k, v in sets_host[0].items()}
        gsub = capi.mpc_solve_batch_host(P, sub, want_u=True)
        osub = bulk.run(sub, h, DT_MPC, np.arange(n_par))
        eot = bulk.err_over_tol(gsub["u"], osub["x_star"]).max(axis=1)
        parity_rec = {"n": n_par, "worst_err_over_tol": float(eot.max()), "n_fail": int((eot > 1.0).sum()),
                      "tolerance": "|f - x*| <= 1e-4 |x*| + 1e-5 element-wise over all 12h forces; x* = certified exact optimum "
                                   "of the reference's QP (oracle build -> converged qpOASES -> extended-precision KKT solve)",
                      "qpoases_worst_err_over_tol": float(bulk.err_over_tol(osub["x_conv"], osub["x_star"]).max()),
                      "status_nonzero": int((gsub["status"] != 0).sum())}

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
        "parity": parity_rec,
        "latency": lat,
        "strong": strong,
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
