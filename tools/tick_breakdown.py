"""Per-stage timing of the bench's full control tick (BASELINE configs[1]); python tools/tick_breakdown.py [NAME]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import build as Bd, capi
if len(sys.argv) > 1 and sys.argv[1] != "main":
    Bd.LIB = os.path.join(ROOT, "scratch", f"libqr_{sys.argv[1]}.so")
import torch
capi.init(0)
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")
stream = torch.cuda.current_stream().cuda_stream
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
robot = pkg.robots.ROBOTS["lite3"]
M = capi.wbc_model_of(robot)
if True:
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


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("tick", timeit(tick))
stages = {
 "gait": lambda: capi.gait_update_batch_device(d_time, d_cfg, 0.1, d_stance, None, False, d_i, d_f, d_pf, d_np, d_sr, stream, allow=d_allow, early=d_early, swing_mask=d_mask, stance_mask=d_stance),
 "inputs": lambda: capi.mpc_inputs_batch_device(h, nhl, dt, d_pf, d_duty, d_early, d_stance, d_init, d_xy, d["gait"], d["traj"], stream),
 "fk": lambda: capi.leg_kinematics_batch_device(G, q, None, foot_base, None, None, stream),
 "lever": lambda: capi.mpc_lever_arms_batch_device(robot, d["quat"], foot_base, d["r_feet"], stream),
 "mpc_ex": lambda: capi.mpc_solve_batch_device_ex(P, d, o, stream, robot, q=q, f_ff=ff, tau=tau_mpc, wbc_cmd=cmd),
 "mpc_plain": lambda: capi.mpc_solve_batch_device(P, d, o, stream),
 "foothold": lambda: capi.foothold_heuristic_batch_device(fP, fd, foothold, planner_phase, stream),
 "swing": lambda: capi.swing_targets_batch_device(G, d["p"], d["quat"], d["v"], foothold, planner_phase, switch_pos, swing_dur, d_mask, True, cmd, stream),
 "wbc": lambda: capi.wbc_solve_batch_device(M, state, cmd, d_stance, tau, stream, status=st),
}
for k, f in stages.items():
    print(f"{k:10s} {timeit(f)*1e3:8.1f} us")
print("mpc rounds mean", float(o["iters"][:,1].float().mean()), "max", int(o["iters"][:,1].max()), "ipm inst", int((o["iters"][:,0]>0).sum()), "nf mean", float((d["gait"]>0).sum(1).float().mean()))
print("stance legs mean", float(d_stance.float().sum(1).mean()), "wbc status nonzero", int((st!=0).sum()))
