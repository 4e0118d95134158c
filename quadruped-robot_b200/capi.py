"""ctypes binding of libqr_gpu.so (include/qr_gpu.h) -- the host-side mirror of the reference's
MPC interface (SetupProblem / SolveMPCKernel / GetMPCSolution) for batches.

There is no fallback: if the library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class MpcParams(C.Structure):
    """qr_mpc_params (include/qr_gpu.h)."""
    _fields_ = [("horizon", C.c_int32), ("dt", C.c_float), ("mu", C.c_float), ("f_max", C.c_float),
                ("mass", C.c_float), ("inertia", C.c_float * 3), ("weights", C.c_float * 12),
                ("alpha", C.c_float)]


class QpOptions(C.Structure):
    """qr_qp_options (include/qr_gpu.h)."""
    _fields_ = [("max_as_rounds", C.c_int32), ("max_ipm_iter", C.c_int32), ("max_polish_rounds", C.c_int32),
                ("flags", C.c_int32), ("ipm_tol", C.c_double),
                ("act_kappa", C.c_double), ("feas_tol", C.c_double), ("mult_tol", C.c_double)]


QP_NO_PREDICTION = 1
QP_SCALAR_FACTOR = 2


def default_options() -> QpOptions:
    return QpOptions(32, 40, 12, 0, 1e-7, 1e3, 1e-9, 1e-11)


EXPORTS = ["qr_gpu_init", "qr_gpu_shutdown", "qr_gpu_last_error", "qr_gpu_mpc_occupancy",
           "qr_gpu_mpc_solve_batch", "qr_gpu_mpc_solve_batch_host", "qr_gpu_mpc_condense_batch",
           "qr_gpu_qp_solve_batch", "qr_gpu_wbc_solve_batch", "qr_gpu_wbc_solve_batch_f64",
           "qr_gpu_swing_parabola_batch", "qr_gpu_mpc_inputs_batch", "qr_gpu_mpc_leg_torque_batch", "qr_gpu_force_balance_batch", "qr_gpu_wbc_solve_batch_host", "qr_gpu_swing_bspline_batch",
           "qr_gpu_foothold_heuristic_batch", "qr_gpu_mpc_solve_batch_host_multi", "qr_gpu_mpc_solve_batch_ex",
           "qr_gpu_mpc_lever_arms_batch", "qr_gpu_leg_kinematics_batch", "qr_gpu_leg_ik_batch", "qr_gpu_swing_targets_batch",
           "qr_gpu_gait_update_batch"]


class WbcModel(C.Structure):
    """qr_wbc_model (include/qr_gpu.h)."""
    _fields_ = [("body_size", C.c_float * 3), ("hip_len", C.c_float), ("upper_len", C.c_float), ("lower_len", C.c_float)]


def wbc_model_of(robot) -> WbcModel:
    m = WbcModel()
    m.body_size[:] = robot.body_size
    m.hip_len, m.upper_len, m.lower_len = robot.hip_len, robot.upper_len, robot.lower_len
    return m


def lib():
    """Load libqr_gpu.so (must have been built by build.build(); never built implicitly on import)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_build.LIB):
            raise RuntimeError(f"{_build.LIB} is missing: run __graft_entry__.build() first")
        _LIB = C.CDLL(_build.LIB)
        _LIB.qr_gpu_last_error.restype = C.c_char_p
    return _LIB


class QrGpuError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise QrGpuError(f"{what} failed with {rc}: {lib().qr_gpu_last_error().decode()}")


def init(device: int = -1):
    _check(lib().qr_gpu_init(device), "qr_gpu_init")


def params_of(robot, horizon: int, dt: float, mu: float | None = None, f_max: float | None = None) -> MpcParams:
    """SetupProblem(dt, horizon, mu, fMax, mass, inertia, weights, alpha) -> qr_mpc_params."""
    P = MpcParams()
    P.horizon = horizon
    P.dt = dt
    P.mu = robot.mu if mu is None else mu
    P.f_max = robot.f_max if f_max is None else f_max
    P.mass = robot.mass
    P.inertia[:] = robot.inertia
    P.weights[:] = robot.weights
    P.alpha = robot.alpha
    return P


def _vp(x):
    """void* of a numpy array (host) or anything with data_ptr() (torch CUDA tensor); None -> NULL."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


_KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")


def mpc_solve_batch_host(P: MpcParams, batch: dict, opt: QpOptions | None = None, per_instance_mu=False,
                         want_u=False, want_info=True):
    """Host buffers in, host buffers out (numpy).  Returns dict(grf, u, status, iters)."""
    B = batch["p"].shape[0]
    h = P.horizon
    grf = np.empty((B, 12), np.float32)
    u = np.empty((B, 12 * h), np.float32) if want_u else None
    status = np.empty(B, np.int32) if want_info else None
    iters = np.empty((B, 2), np.int32) if want_info else None
    rc = lib().qr_gpu_mpc_solve_batch_host(
        C.byref(P), C.byref(opt) if opt is not None else None, B, *[_vp(batch[k]) for k in _KEYS],
        _vp(batch["mu"]) if per_instance_mu else None, None, _vp(grf), _vp(u), _vp(status), _vp(iters))
    _check(rc, "qr_gpu_mpc_solve_batch_host")
    return dict(grf=grf, u=u, status=status, iters=iters)


def mpc_solve_batch_host_multi(P: MpcParams, batch: dict, devices, opt: QpOptions | None = None, per_instance_mu=False,
                               want_u=False, want_info=True, out: dict | None = None):
    """One host batch sharded over `devices` (one host thread and one contiguous shard per GPU of this process);
    results gathered into one set of host arrays.  `out` may supply preallocated (e.g. pinned) result arrays."""
    B = batch["p"].shape[0]
    h = P.horizon
    out = out or {}
    grf = out.get("grf") if out.get("grf") is not None else np.empty((B, 12), np.float32)
    u = out.get("u") if want_u and out.get("u") is not None else (np.empty((B, 12 * h), np.float32) if want_u else None)
    status = out.get("status") if out.get("status") is not None else (np.empty(B, np.int32) if want_info else None)
    iters = out.get("iters") if out.get("iters") is not None else (np.empty((B, 2), np.int32) if want_info else None)
    devs = (C.c_int * len(devices))(*devices)
    rc = lib().qr_gpu_mpc_solve_batch_host_multi(
        len(devices), devs, C.byref(P), C.byref(opt) if opt is not None else None, B, *[_vp(batch[k]) for k in _KEYS],
        _vp(batch["mu"]) if per_instance_mu else None, None, _vp(grf), _vp(u), _vp(status), _vp(iters))
    _check(rc, "qr_gpu_mpc_solve_batch_host_multi")
    return dict(grf=grf, u=u, status=status, iters=iters)


def mpc_solve_batch_device(P: MpcParams, dev: dict, out: dict, stream_ptr: int, opt: QpOptions | None = None,
                           per_instance_mu=False):
    """Device tensors in `dev` (torch), outputs in `out` (grf, optional u/status/iters); asynchronous."""
    B = dev["p"].shape[0]
    rc = lib().qr_gpu_mpc_solve_batch(
        C.byref(P), C.byref(opt) if opt is not None else None, B, *[_vp(dev[k]) for k in _KEYS],
        _vp(dev["mu"]) if per_instance_mu else None, None, _vp(out["grf"]), _vp(out.get("u")),
        _vp(out.get("status")), _vp(out.get("iters")), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_mpc_solve_batch")


def mpc_condense_batch_device(P: MpcParams, dev: dict, H, g, ub, stream_ptr: int):
    B = dev["p"].shape[0]
    rc = lib().qr_gpu_mpc_condense_batch(C.byref(P), B, *[_vp(dev[k]) for k in _KEYS], None,
                                         _vp(H), _vp(g), _vp(ub), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_mpc_condense_batch")


def qp_solve_batch_device(horizon: int, mu: float, H, g, ub, x32, x64, status, iters, stream_ptr: int,
                          opt: QpOptions | None = None, mu_i=None):
    B = H.shape[0]
    rc = lib().qr_gpu_qp_solve_batch(horizon, C.c_float(mu), C.byref(opt) if opt is not None else None, B,
                                     _vp(H), _vp(g), _vp(ub), _vp(mu_i), _vp(x32), _vp(x64), _vp(status),
                                     _vp(iters), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_qp_solve_batch")


def occupancy(horizon: int, stance_footsteps: int):
    sm, per, thr, smem = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib().qr_gpu_mpc_occupancy(horizon, stance_footsteps, C.byref(sm), C.byref(per), C.byref(thr), C.byref(smem)),
           "qr_gpu_mpc_occupancy")
    return dict(sm_count=sm.value, ctas_per_sm=per.value, threads_per_cta=thr.value, smem_bytes=smem.value)


def wbc_solve_batch_device(model: WbcModel, state, cmd, contact, tau, stream_ptr: int, fr=None, qdes=None, qddes=None,
                           status=None):
    """float32 outputs (torch CUDA tensors); asynchronous on the stream."""
    rc = lib().qr_gpu_wbc_solve_batch(C.byref(model), state.shape[0], _vp(state), _vp(cmd), _vp(contact), _vp(tau),
                                      _vp(fr), _vp(qdes), _vp(qddes), _vp(status), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_wbc_solve_batch")


def wbc_solve_batch_device_f64(model: WbcModel, state, cmd, contact, tau, stream_ptr: int, fr=None, qdes=None,
                               qddes=None, dbg=None, status=None):
    rc = lib().qr_gpu_wbc_solve_batch_f64(C.byref(model), state.shape[0], _vp(state), _vp(cmd), _vp(contact), _vp(tau),
                                          _vp(fr), _vp(qdes), _vp(qddes), _vp(dbg), _vp(status), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_wbc_solve_batch_f64")


def swing_parabola_batch_device(start, end, height, phase, phase_module: bool, pos, valid, stream_ptr: int):
    rc = lib().qr_gpu_swing_parabola_batch(start.shape[0], _vp(start), _vp(end), _vp(height), _vp(phase),
                                           int(phase_module), _vp(pos), _vp(valid), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_swing_parabola_batch")


def mpc_inputs_batch_device(horizon, num_horizon_l, dt_mpc, progress, duty, early, contacts, traj_init, pos_xy,
                            gait_out, traj_out, stream_ptr: int):
    rc = lib().qr_gpu_mpc_inputs_batch(horizon, num_horizon_l, C.c_float(dt_mpc), progress.shape[0], _vp(progress), _vp(duty),
                                       _vp(early), _vp(contacts), _vp(traj_init), _vp(pos_xy), _vp(gait_out), _vp(traj_out),
                                       C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_mpc_inputs_batch")


def mpc_leg_torque_batch_device(robot, quat, q, grf, f_ff, tau, stream_ptr: int):
    rc = lib().qr_gpu_mpc_leg_torque_batch(C.c_float(robot.hip_len), C.c_float(robot.upper_len), C.c_float(robot.lower_len),
                                           quat.shape[0], _vp(quat), _vp(q), _vp(grf), _vp(f_ff), _vp(tau), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_mpc_leg_torque_batch")


class FbParams(C.Structure):
    """qr_fb_params of include/qr_gpu.h."""
    _fields_ = [("mass", C.c_float), ("inertia", C.c_float * 9), ("acc_weight", C.c_float * 6), ("reg_weight", C.c_float),
                ("mu", C.c_float), ("fmin_ratio", C.c_float * 4), ("fmax_ratio", C.c_float * 4), ("world_frame", C.c_int32)]


def fb_params_of(p: dict) -> FbParams:
    import numpy as np
    P = FbParams()
    P.mass = p["mass"]
    P.inertia[:] = [float(v) for v in np.asarray(p["inertia"], np.float32).reshape(9)]
    P.acc_weight[:] = [float(v) for v in p["acc_weight"]]
    P.reg_weight, P.mu = p["reg_weight"], p["mu"]
    P.fmin_ratio[:] = [float(v) for v in p["fmin_ratio"]]
    P.fmax_ratio[:] = [float(v) for v in p["fmax_ratio"]]
    P.world_frame = int(p["world_frame"])
    return P


def force_balance_batch_device(P: FbParams, foot, acc, contact, force, stream_ptr: int, inertia=None, gravity=None,
                               frame=None, status=None, iters=None):
    """torch CUDA tensors; asynchronous on the stream."""
    rc = lib().qr_gpu_force_balance_batch(C.byref(P), foot.shape[0], _vp(inertia), _vp(foot), _vp(acc), _vp(contact),
                                          _vp(gravity), _vp(frame), _vp(force), _vp(status), _vp(iters), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_force_balance_batch")


def wbc_solve_batch_host(model: WbcModel, state, cmd, contact):
    """numpy in / numpy out through the host-buffer entry point: dict(tau, fr, qdes, qddes, status)."""
    import numpy as np
    B = state.shape[0]
    f = lambda a: np.ascontiguousarray(a, np.float32)
    state, cmd, contact = f(state), f(cmd), np.ascontiguousarray(contact, np.int32)
    out = {k: np.zeros((B, 12), np.float32) for k in ("tau", "fr", "qdes", "qddes")}
    st = np.zeros(B, np.int32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().qr_gpu_wbc_solve_batch_host(C.byref(model), B, fp(state), fp(cmd), fp(contact), fp(out["tau"]), fp(out["fr"]),
                                           fp(out["qdes"]), fp(out["qddes"]), fp(st))
    _check(rc, "qr_gpu_wbc_solve_batch_host")
    out["status"] = st
    return out


class FootholdParams(C.Structure):
    """qr_foothold_params of include/qr_gpu.h."""
    _fields_ = [("hip_offset", C.c_float * 12), ("hip_pos", C.c_float * 12), ("hip_len", C.c_float), ("swing_kp", C.c_float * 3)]


def foothold_params_of(p: dict) -> FootholdParams:
    import numpy as np
    P = FootholdParams()
    P.hip_offset[:] = [float(v) for v in np.asarray(p["hip_offset"], np.float32).reshape(12)]
    P.hip_pos[:] = [float(v) for v in np.asarray(p["hip_pos"], np.float32).reshape(12)]
    P.hip_len = p["hip_len"]
    P.swing_kp[:] = [float(v) for v in p["swing_kp"]]
    return P


def swing_bspline_batch_device(initial_pos, target_pos, height, duration, initial_time, time, pos, vel, valid, stream_ptr: int):
    rc = lib().qr_gpu_swing_bspline_batch(initial_pos.shape[0], _vp(initial_pos), _vp(target_pos), _vp(height), _vp(duration),
                                          _vp(initial_time), _vp(time), _vp(pos), _vp(vel), _vp(valid), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_swing_bspline_batch")


def foothold_heuristic_batch_device(P: FootholdParams, d: dict, foothold, phase, stream_ptr: int):
    """d: torch CUDA tensors named like make_foothold_batch's keys."""
    rc = lib().qr_gpu_foothold_heuristic_batch(
        C.byref(P), d["com_vel"].shape[0], _vp(d["com_vel"]), _vp(d["rpy_rate"]), _vp(d["dR"]), _vp(d["base_R"]), _vp(d["rpy"]),
        _vp(d["foot_base"]), _vp(d["des_speed"]), _vp(d["des_twist"]), _vp(d["des_height"]), _vp(d["swing_remain"]),
        _vp(d["norm_phase"]), _vp(d["allow_switch"]), _vp(d["swing_mask"]), _vp(foothold), _vp(phase), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_foothold_heuristic_batch")


class MpcEpilogue(C.Structure):
    """qr_mpc_epilogue of include/qr_gpu.h."""
    _fields_ = [("hip_len", C.c_float), ("upper_len", C.c_float), ("lower_len", C.c_float), ("q", C.c_void_p),
                ("f_ff_out", C.c_void_p), ("tau_out", C.c_void_p), ("wbc_cmd_io", C.c_void_p)]


def _ptr(x):
    return None if x is None else x.data_ptr()


def mpc_solve_batch_device_ex(P: MpcParams, dev: dict, out: dict, stream_ptr: int, robot, q=None, f_ff=None, tau=None,
                              wbc_cmd=None, opt: QpOptions | None = None, per_instance_mu=False):
    """qr_gpu_mpc_solve_batch_ex: the solve with its fused post-processing (leg forces, joint torques, Fr_des rows)."""
    ep = MpcEpilogue(robot.hip_len, robot.upper_len, robot.lower_len, _ptr(q), _ptr(f_ff), _ptr(tau), _ptr(wbc_cmd))
    B = dev["p"].shape[0]
    rc = lib().qr_gpu_mpc_solve_batch_ex(
        C.byref(P), C.byref(opt) if opt is not None else None, B, *[_vp(dev[k]) for k in _KEYS],
        _vp(dev["mu"]) if per_instance_mu else None, None, _vp(out["grf"]), _vp(out.get("u")),
        _vp(out.get("status")), _vp(out.get("iters")), C.byref(ep), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_mpc_solve_batch_ex")


class LegGeometry(C.Structure):
    """qr_leg_geometry of include/qr_gpu.h."""
    _fields_ = [("hip_len", C.c_float), ("upper_len", C.c_float), ("lower_len", C.c_float), ("hip_offset", C.c_float * 12)]


def leg_geometry_of(robot) -> LegGeometry:
    """hipOffset = the abad joint positions (hip positions moved inwards by the hip length), 3x4 column-major."""
    g = LegGeometry()
    g.hip_len, g.upper_len, g.lower_len = robot.hip_len, robot.upper_len, robot.lower_len
    hips = np.array(robot.hip_positions, np.float64)
    abad = hips.copy()
    abad[:, 1] -= np.sign(hips[:, 1]) * robot.hip_len
    g.hip_offset[:] = [float(v) for v in abad.astype(np.float32).reshape(12)]
    return g


def mpc_lever_arms_batch_device(robot, quat, foot_base, r_feet, stream_ptr: int):
    com = (C.c_float * 3)(*[float(v) for v in robot.com_offset])
    rc = lib().qr_gpu_mpc_lever_arms_batch(quat.shape[0], _vp(quat), _vp(foot_base), com, _vp(r_feet), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_mpc_lever_arms_batch")


def leg_kinematics_batch_device(G: LegGeometry, q, qd, foot_base, jac, foot_vel, stream_ptr: int):
    rc = lib().qr_gpu_leg_kinematics_batch(C.byref(G), q.shape[0], _vp(q), _vp(qd), _vp(foot_base), _vp(jac), _vp(foot_vel),
                                           C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_leg_kinematics_batch")


def leg_ik_batch_device(G: LegGeometry, foot_base, foot_vel, leg_mask, q_out, qd_out, stream_ptr: int):
    rc = lib().qr_gpu_leg_ik_batch(C.byref(G), foot_base.shape[0], _vp(foot_base), _vp(foot_vel), _vp(leg_mask), _vp(q_out),
                                   _vp(qd_out), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_leg_ik_batch")


def swing_targets_batch_device(G: LegGeometry, base_pos, quat, v_world, foothold, planner_phase, switch_pos, swing_duration,
                               swing_mask, horizontal_terrain: bool, wbc_cmd, stream_ptr: int, foot_base_des=None, q_des=None,
                               qd_des=None, valid=None):
    rc = lib().qr_gpu_swing_targets_batch(C.byref(G), base_pos.shape[0], _vp(base_pos), _vp(quat), _vp(v_world), _vp(foothold),
                                          _vp(planner_phase), _vp(switch_pos), _vp(swing_duration), _vp(swing_mask),
                                          int(horizontal_terrain), _vp(wbc_cmd), _vp(foot_base_des), _vp(q_des), _vp(qd_des),
                                          _vp(valid), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_swing_targets_batch")


def gait_update_batch_device(time, cfg, contact_threshold: float, contacts, stop, advanced_trot: bool, istate, fstate,
                             phase_full, norm_phase, swing_remain, stream_ptr: int, allow=None, early=None, swing_mask=None,
                             stance_mask=None):
    rc = lib().qr_gpu_gait_update_batch(time.shape[0], _vp(time), _vp(cfg), C.c_float(contact_threshold), _vp(contacts), _vp(stop),
                                        int(advanced_trot), _vp(istate), _vp(fstate), _vp(phase_full), _vp(norm_phase),
                                        _vp(swing_remain), _vp(allow), _vp(early), _vp(swing_mask), _vp(stance_mask), C.c_void_p(stream_ptr))
    _check(rc, "qr_gpu_gait_update_batch")
