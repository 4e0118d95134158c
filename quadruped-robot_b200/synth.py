"""Seeded synthetic MPC workloads and the host-side contact-table /
reference-trajectory builders.

Everything is float32 and laid out as one contiguous row per robot instance ("problem rows"):
    p[B,3] v[B,3] quat[B,4] w[B,3] r_feet[B,12] rpy[B,3] traj[B,12h] gait[B,4h] mu[B] f_max[B]
which is the layout `qr_gpu_mpc_solve_batch` consumes (one CTA per problem reads its rows
with coalesced loads).

contact_table / reference_traj restate
    quadruped/src/controllers/mpc/qr_mpc_stance_leg_controller.cpp:282-303 and :345-376
in float32 so that masks are bit-exact with the reference's x86-64 build (no FMA contraction).
"""
from __future__ import annotations

import numpy as np

from .robots import GAITS, ROBOTS, RobotMPC

F32 = np.float32


def num_horizon_l(gait: dict) -> int:
    """numHorizonL = max(2, int(fullCyclePeriod / 0.4)), qr_mpc_stance_leg_controller.cpp:50."""
    full_cycle = F32(gait["stance_duration"]) / F32(gait["duty"])
    return max(2, int(float(full_cycle) / 0.4))


def contact_table(horizon: int, n_horizon_l: int, progress: np.ndarray, duty: np.ndarray,
                  early_contact: np.ndarray | None = None,
                  contacts: np.ndarray | None = None) -> np.ndarray:
    """mpcTable [B, h, 4] float32 of 0/1 (qr_mpc_stance_leg_controller.cpp:282-303).

    progress, duty: [B,4] float32.  early_contact / contacts: [B,4] bool or None.
    """
    progress = np.asarray(progress, F32)
    duty = np.asarray(duty, F32)
    B = progress.shape[0]
    d_phase = F32(1.0 / (n_horizon_l * horizon))  # double division, narrowed
    table = np.empty((B, horizon, 4), F32)
    for i in range(horizon):
        ph = progress + F32(i) * d_phase  # float32 multiply, float32 add (two roundings)
        # `while (ph > 1.0) ph -= 1.0` -- at most a few wraps for progress in [0,1]
        for _ in range(4):
            over = ph > F32(1.0)
            if not over.any():
                break
            ph = np.where(over, ph - F32(1.0), ph).astype(F32)
        st = ph < duty
        if early_contact is not None:
            st = st | np.asarray(early_contact, bool)
        table[:, i, :] = st
    if contacts is not None:
        table[:, 0, :] = np.asarray(contacts, bool)
    return table


def reference_traj(horizon: int, dt_mpc: float, init: np.ndarray, pos_xy: np.ndarray) -> np.ndarray:
    """trajAll [B, 12h] float32 (qr_mpc_stance_leg_controller.cpp:345-376)."""
    init = np.array(init, F32, copy=True)
    pos_xy = np.asarray(pos_xy, F32)
    dt = F32(dt_mpc)
    for a in range(2):
        lo = pos_xy[:, a] - F32(0.1)
        hi = pos_xy[:, a] + F32(0.1)
        init[:, 3 + a] = np.minimum(np.maximum(init[:, 3 + a], lo), hi)
    B = init.shape[0]
    traj = np.empty((B, horizon, 12), F32)
    traj[:, :, :] = init[:, None, :]
    yaw_rate, vx, vy = init[:, 8], init[:, 9], init[:, 10]
    for i in range(1, horizon):
        traj[:, i, 2] = traj[:, i - 1, 2] + dt * yaw_rate
        traj[:, i, 3] = traj[:, i - 1, 3] + dt * vx
        traj[:, i, 4] = traj[:, i - 1, 4] + dt * vy
    return traj.reshape(B, 12 * horizon)


def _rot_zyx(rpy: np.ndarray) -> np.ndarray:
    """R = Rz(yaw) Ry(pitch) Rx(roll), float64 [B,3,3]."""
    r, p, y = rpy[:, 0], rpy[:, 1], rpy[:, 2]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.empty((rpy.shape[0], 3, 3))
    R[:, 0, 0] = cy * cp; R[:, 0, 1] = cy * sp * sr - sy * cr; R[:, 0, 2] = cy * sp * cr + sy * sr
    R[:, 1, 0] = sy * cp; R[:, 1, 1] = sy * sp * sr + cy * cr; R[:, 1, 2] = sy * sp * cr - cy * sr
    R[:, 2, 0] = -sp;     R[:, 2, 1] = cp * sr;                R[:, 2, 2] = cp * cr
    return R


def _quat_from_rpy(rpy: np.ndarray) -> np.ndarray:
    """(w,x,y,z) of Rz Ry Rx, float64 [B,4]."""
    hr, hp, hy = rpy[:, 0] / 2, rpy[:, 1] / 2, rpy[:, 2] / 2
    cr, sr, cp, sp, cy, sy = np.cos(hr), np.sin(hr), np.cos(hp), np.sin(hp), np.cos(hy), np.sin(hy)
    q = np.stack([cy * cp * cr + sy * sp * sr,
                  cy * cp * sr - sy * sp * cr,
                  cy * sp * cr + sy * cp * sr,
                  sy * cp * cr - cy * sp * sr], axis=1)
    return q


def make_mpc_batch(robot: str | RobotMPC = "a1", horizon: int = 10, dt: float = 0.03,
                   batch: int = 1024, seed: int = 0, gait: str = "trot",
                   mu_sweep: bool = False) -> dict:
    """Randomised states per SURVEY.md section 8d.  gait = trot | walk | gallop | stand | mixed."""
    rb = ROBOTS[robot] if isinstance(robot, str) else robot
    rng = np.random.default_rng(seed)
    B, h = batch, horizon
    U = rng.uniform
    rpy = np.stack([U(-0.1, 0.1, B), U(-0.1, 0.1, B), U(-np.pi, np.pi, B)], axis=1)
    p = np.stack([U(-1, 1, B), U(-1, 1, B), rb.body_height + U(-0.02, 0.02, B)], axis=1)
    w = U(-0.5, 0.5, (B, 3))
    v = np.stack([U(-0.5, 1.0, B), U(-0.3, 0.3, B), U(-0.1, 0.1, B)], axis=1)
    R = _rot_zyx(rpy)
    quat = _quat_from_rpy(rpy)
    hips = np.array(rb.hip_positions, float)  # [4,3]
    feet_base = hips[None, :, :] + np.array([0, 0, -rb.body_height]) + U(-0.05, 0.05, (B, 4, 3))
    feet_base = feet_base - np.array(rb.com_offset)
    r_feet = np.einsum("bij,blj->bli", R, feet_base)  # world-aligned lever arms [B,4,3]
    cmd = np.stack([U(-0.5, 1.0, B), U(-0.3, 0.3, B), U(-0.5, 0.5, B)], axis=1)  # vx, vy, yawRate
    v_des_w = np.einsum("bij,bj->bi", R, np.stack([cmd[:, 0], cmd[:, 1], np.zeros(B)], axis=1))
    init = np.zeros((B, 12))
    init[:, 2] = rpy[:, 2]
    init[:, 3:5] = p[:, :2] + U(-0.15, 0.15, (B, 2))  # desired xy, clipped to +-0.1 of actual
    init[:, 5] = rb.body_height
    init[:, 8] = cmd[:, 2]
    init[:, 9:11] = v_des_w[:, :2]
    traj = reference_traj(h, dt, init.astype(F32), p[:, :2].astype(F32))

    names = ["trot", "walk", "gallop"] if gait == "mixed" else [gait]
    pick = rng.integers(0, len(names), B)
    phase0 = U(0, 1, B)
    table = np.empty((B, h, 4), F32)
    for gi, gname in enumerate(names):
        sel = np.nonzero(pick == gi)[0]
        if sel.size == 0:
            continue
        gt = GAITS[gname]
        progress = np.mod(phase0[sel, None] + np.array(gt["offsets"])[None, :], 1.0).astype(F32)
        duty = np.full((sel.size, 4), gt["duty"], F32)
        table[sel] = contact_table(h, num_horizon_l(gt), progress, duty)
    mu = U(0.2, 1.0, B) if mu_sweep else np.full(B, rb.mu)
    out = dict(
        p=p, v=v, quat=quat, w=w, r_feet=r_feet.reshape(B, 12), rpy=rpy, traj=traj,
        gait=table.reshape(B, 4 * h), mu=mu, f_max=np.full(B, rb.f_max))
    out = {k: np.ascontiguousarray(a, dtype=F32) for k, a in out.items()}
    out["robot"] = rb
    out["horizon"] = h
    out["dt"] = dt
    return out


# ------------------------------------------------------------------------------------------------
# Whole-body-control workloads
# ------------------------------------------------------------------------------------------------
def _rx(a):
    c, s = np.cos(a), np.sin(a)
    o, z = np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([o, z, z], -1), np.stack([z, c, -s], -1), np.stack([z, s, c], -1)], -2)


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    o, z = np.ones_like(a), np.zeros_like(a)
    return np.stack([np.stack([c, z, s], -1), np.stack([z, o, z], -1), np.stack([-s, z, c], -1)], -2)


def foot_positions_world(rb: RobotMPC, rpy: np.ndarray, pos: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Analytic forward kinematics of the four feet [B,4,3] for the tree of BuildDynamicModel
    (src/robots/qr_robot_a1_sim.cpp:176-345): abad about x at (+-0.1805, +-0.047, 0), hip about y at
    (0, +-hip_len, 0), knee about y at (0, 0, -upper_len), foot at (0, -+0.004, -lower_len)."""
    B = q.shape[0]
    R = _rot_zyx(rpy)
    out = np.empty((B, 4, 3))
    for leg in range(4):
        sx = 1.0 if leg < 2 else -1.0
        sy = -1.0 if leg % 2 == 0 else 1.0
        t_abad = np.array([sx * 0.1805, sy * 0.047, 0.0])
        t_hip = np.array([0.0, sy * rb.hip_len, 0.0])
        t_knee = np.array([0.0, 0.0, -rb.upper_len])
        t_foot = np.array([0.0, -sy * 0.004, -rb.lower_len])
        q0, q1, q2 = q[:, 3 * leg], q[:, 3 * leg + 1], q[:, 3 * leg + 2]
        p = t_knee + np.einsum("bij,j->bi", _ry(q2), t_foot)
        p = t_hip + np.einsum("bij,bj->bi", _ry(q1), p)
        p = t_abad + np.einsum("bij,bj->bi", _rx(q0), p)
        out[:, leg] = pos + np.einsum("bij,bj->bi", R, p)
    return out


def make_wbc_batch(robot: str | RobotMPC = "lite3", batch: int = 1024, seed: int = 0) -> dict:
    """Randomised WBC inputs: state[B,37] (quat pos twist q qd), cmd[B,66] (the qrWbcCtrlData fields +
    the previous orientation-velocity command), contact[B,4] int32.  Trot-like contact patterns
    (a diagonal pair, three or four feet in stance)."""
    rb = ROBOTS[robot] if isinstance(robot, str) else robot
    rng = np.random.default_rng(seed)
    B = batch
    U = rng.uniform
    rpy = np.stack([U(-0.1, 0.1, B), U(-0.1, 0.1, B), U(-np.pi, np.pi, B)], 1)
    pos = np.stack([U(-1, 1, B), U(-1, 1, B), rb.body_height + U(-0.02, 0.02, B)], 1)
    twist = np.concatenate([U(-0.5, 0.5, (B, 3)), np.stack([U(-0.5, 1.0, B), U(-0.3, 0.3, B), U(-0.1, 0.1, B)], 1)], 1)
    stand = np.tile(np.array([0.0, 0.9, -1.8]), 4)
    q = stand[None, :] + U(-0.2, 0.2, (B, 12))
    qd = U(-1, 1, (B, 12))
    state = np.concatenate([_quat_from_rpy(rpy), pos, twist, q, qd], 1)
    patterns = np.array([[1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1], [1, 1, 0, 1], [0, 1, 1, 1], [1, 0, 1, 1], [1, 1, 1, 0]], np.int32)
    contact = patterns[rng.integers(0, len(patterns), B)]
    feet = foot_positions_world(rb, rpy, pos, q)
    nc = contact.sum(1)
    fr = np.zeros((B, 4, 3))
    fr[..., 0] = U(-10, 10, (B, 4))
    fr[..., 1] = U(-10, 10, (B, 4))
    fr[..., 2] = (rb.mass * 9.81 / nc)[:, None] + U(-10, 10, (B, 4))
    fr *= contact[..., None]
    cmd = np.concatenate([
        pos + U(-0.03, 0.03, (B, 3)),                                    # pBody_des
        np.stack([U(-0.5, 1.0, B), U(-0.3, 0.3, B), np.zeros(B)], 1),    # vBody_des
        np.zeros((B, 3)),                                                # aBody_des
        rpy + np.stack([U(-0.05, 0.05, B), U(-0.05, 0.05, B), U(-0.1, 0.1, B)], 1),  # pBody_RPY_des
        np.stack([np.zeros(B), np.zeros(B), U(-0.5, 0.5, B)], 1),        # vBody_Ori_des
        (feet + U(-0.1, 0.1, (B, 4, 3))).reshape(B, 12),                 # pFoot_des
        U(-0.5, 0.5, (B, 12)),                                           # vFoot_des
        U(-2, 2, (B, 12)),                                               # aFoot_des
        fr.reshape(B, 12),                                               # Fr_des (from the MPC)
        np.stack([np.zeros(B), np.zeros(B), U(-0.5, 0.5, B)], 1),        # previous vBody_Ori_des
    ], 1)
    return dict(state=np.ascontiguousarray(state, F32), cmd=np.ascontiguousarray(cmd, F32),
                contact=np.ascontiguousarray(contact, np.int32), rpy=rpy, robot=rb)


# ------------------------------------------------------------------------------------------------
# Force-balance stance controller workloads
# ------------------------------------------------------------------------------------------------
def make_fb_batch(robot: str | RobotMPC = "a1", batch: int = 256, seed: int = 0, world_frame: bool = False,
                  tilted: bool = False) -> dict:
    """Inputs of ComputeContactForce for `batch` robots: foot[B,12] (4x3 foot positions in the frame of the
    computation), acc[B,6] desired acceleration, contact[B,4]; optional per-robot inertia[B,9], gravity[B,3] and
    frame[B,9] (normal, tangent1, tangent2) for a tilted control frame.  acc_weight / ratios are the values of
    config/a1_sim/stance_leg_controller.yaml and qr_torque_stance_leg_controller.cpp:98-108."""
    rb = ROBOTS[robot] if isinstance(robot, str) else robot
    rng = np.random.default_rng(seed)
    B = batch
    U = rng.uniform
    hips = np.array(rb.hip_positions, float)
    foot = hips[None] + np.array([0, 0, -rb.body_height]) + U(-0.05, 0.05, (B, 4, 3))
    acc = np.concatenate([U(-2, 2, (B, 2)), U(-1.5, 1.5, (B, 1)), U(-4, 4, (B, 3))], 1)
    patterns = np.array([[1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1], [1, 1, 0, 1], [0, 1, 1, 1], [1, 0, 1, 1], [1, 1, 1, 0]], np.int32)
    contact = patterns[rng.integers(0, len(patterns), B)]
    params = dict(mass=rb.mass, inertia=np.diag(rb.inertia).astype(F32), acc_weight=(1.0, 1.0, 1.0, 10.0, 10.0, 1.0),
                  reg_weight=1e-4, mu=0.45 if world_frame else 0.5, fmin_ratio=(0.01,) * 4, fmax_ratio=(10.0,) * 4,
                  world_frame=int(world_frame))
    out = dict(foot=np.ascontiguousarray(foot.reshape(B, 12), F32), acc=np.ascontiguousarray(acc, F32),
               contact=np.ascontiguousarray(contact), params=params, robot=rb, inertia=None, gravity=None, frame=None)
    if tilted:
        pitch = U(-0.3, 0.3, B)
        rpy = np.stack([U(-0.1, 0.1, B), pitch, np.zeros(B)], 1)
        R = _rot_zyx(rpy)
        I = np.diag(rb.inertia)
        out["inertia"] = np.ascontiguousarray(np.einsum("bij,jk,blk->bil", R, I, R).reshape(B, 9), F32)
        out["gravity"] = np.ascontiguousarray(np.einsum("bji,j->bi", R, np.array([0, 0, 9.8])), F32)
        n = np.stack([-np.sin(pitch), np.zeros(B), np.cos(pitch)], 1)
        t2 = np.tile(np.array([0.0, 1.0, 0.0]), (B, 1))
        t1 = np.cross(t2, n)
        out["frame"] = np.ascontiguousarray(np.concatenate([n, t1, t2], 1), F32)
    return out


# ------------------------------------------------------------------------------------------------
# Swing-leg workloads
# ------------------------------------------------------------------------------------------------
def make_swing_batch(batch: int = 256, seed: int = 0) -> dict:
    """Feet for the B-spline generator: start / target positions (steps up and down), apex height, duration 1 (the
    generator is driven with the swing phase), sample times inside and just outside [0, duration]."""
    rng = np.random.default_rng(seed)
    B = batch
    U = rng.uniform
    ip = np.stack([U(-0.3, 0.3, B), U(-0.2, 0.2, B), U(-0.32, -0.25, B)], 1)
    tp = ip + np.stack([U(-0.15, 0.25, B), U(-0.08, 0.08, B), U(-0.06, 0.06, B)], 1)
    tp[: B // 8, :2] = ip[: B // 8, :2]          # purely vertical steps (atan2(0, 0))
    t = U(-0.002, 1.002, B)
    t[:8] = [0.0, 1.0, 0.5, 0.05, -0.0015, 1.0015, -0.0005, 1.0005]   # incl. two samples the generator rejects
    return dict(initial_pos=ip.astype(F32), target_pos=tp.astype(F32), height=U(0.04, 0.12, B).astype(F32),
                duration=np.ones(B, F32), initial_time=np.zeros(B, F32), time=t.astype(F32))


def make_foothold_batch(robot: str | RobotMPC = "a1", batch: int = 256, seed: int = 0) -> dict:
    """Inputs of ComputeHeuristicFootHold for `batch` robots (see qr_gpu_foothold_heuristic_batch)."""
    rb = ROBOTS[robot] if isinstance(robot, str) else robot
    rng = np.random.default_rng(seed)
    B = batch
    U = rng.uniform
    rpy = np.stack([U(-0.15, 0.15, B), U(-0.15, 0.15, B), U(-np.pi, np.pi, B)], 1)
    base_R = _rot_zyx(rpy)
    dR = _rot_zyx(np.stack([rpy[:, 0], rpy[:, 1], np.zeros(B)], 1))
    hips = np.array(rb.hip_positions, float)                       # [4,3]
    abad = hips.copy(); abad[:, 1] -= np.sign(hips[:, 1]) * rb.hip_len
    foot_base = hips[None] + np.array([0, 0, -rb.body_height]) + U(-0.06, 0.06, (B, 4, 3))
    swing = rng.integers(0, 2, (B, 4)).astype(np.int32)
    swing[:, 0] = 1
    params = dict(hip_offset=abad.astype(F32).reshape(12), hip_pos=hips.astype(F32).reshape(12), hip_len=rb.hip_len,
                  swing_kp=(0.03, 0.03, 0.03))
    return dict(com_vel=U(-0.6, 1.0, (B, 3)).astype(F32), rpy_rate=U(-0.8, 0.8, (B, 3)).astype(F32),
                dR=np.ascontiguousarray(dR.reshape(B, 9), F32), base_R=np.ascontiguousarray(base_R.reshape(B, 9), F32),
                rpy=rpy.astype(F32), foot_base=np.ascontiguousarray(foot_base.reshape(B, 12), F32),
                des_speed=np.stack([U(-0.5, 1.0, B), U(-0.3, 0.3, B), np.zeros(B)], 1).astype(F32),
                des_twist=U(-0.6, 0.6, B).astype(F32), des_height=np.full(B, rb.body_height - 0.01, F32),
                swing_remain=U(0.0, 0.25, (B, 4)).astype(F32), norm_phase=U(0, 1, (B, 4)).astype(F32),
                allow_switch=rng.integers(0, 2, (B, 4)).astype(np.int32), swing_mask=swing, params=params, robot=rb)
