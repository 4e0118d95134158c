"""ctypes binding of the CPU oracle (oracle/libqr_oracle.so) + an extended-precision KKT polish.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under quadruped-robot_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RET_MAX_NWSR_REACHED = 64  # qpOASES MessageHandling.hpp


class MpcParams(C.Structure):
    _fields_ = [("horizon", C.c_int), ("dt", C.c_float), ("mu", C.c_float), ("f_max", C.c_float),
                ("mass", C.c_float), ("inertia", C.c_float * 3), ("weights", C.c_float * 12),
                ("alpha", C.c_float)]


def build(force: bool = False) -> str:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "libqr_oracle.so")
    have_ref = os.path.isdir("/root/reference/quadruped/extern/qpOASES/src")
    if force or not os.path.exists(so) or have_ref:
        if have_ref:
            subprocess.run(["make", "-s", "-j8", "-C", _HERE, "all", "refmpc", "refwbc", "refctl"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        elif not os.path.exists(so):
            raise RuntimeError("oracle/libqr_oracle.so missing and /root/reference not available to build it")
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libqr_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.qro_mpc_time_batch.restype = C.c_double
    return _LIB


def params_of(robot, horizon: int, dt: float, mu: float | None = None, f_max: float | None = None) -> MpcParams:
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


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def mpc_build(P: MpcParams, batch: dict, i: int, want_ab: bool = False):
    """Float32 (H, g, ub) of problem i of a synth batch."""
    h = P.horizon
    n, m = 12 * h, 20 * h
    H = np.empty((n, n), np.float32)
    g = np.empty(n, np.float32)
    ub = np.empty(m, np.float32)
    Aqp = np.empty((13 * h, 13), np.float32) if want_ab else None
    Bqp = np.empty((13 * h, n), np.float32) if want_ab else None
    rows = [np.ascontiguousarray(batch[k][i]) for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")]
    rc = lib().qro_mpc_build(C.byref(P), *[_fp(r) for r in rows], _fp(H), _fp(g), _fp(ub),
                             _fp(Aqp) if want_ab else None, _fp(Bqp) if want_ab else None)
    assert rc == 0
    return (H, g, ub, Aqp, Bqp) if want_ab else (H, g, ub)



_REF = None


def ref_mpc_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libqr_mpc_ref.so"))


def ref_mpc_solve(P: MpcParams, batch: dict, i: int):
    """The reference's OWN qr_mpc_interface.cpp (oracle/_ref/libqr_mpc_ref.so, compiled unmodified
    against oracle/mini_eigen): SetupProblem + SolveMPCKernel + GetMPCSolution on problem i.
    Returns (H, g, ub, x) read back from the reference's qpOASES buffers (float64)."""
    global _REF
    if _REF is None:
        _REF = C.CDLL(os.path.join(_HERE, "_ref", "libqr_mpc_ref.so"))
    h = P.horizon
    n, m = 12 * h, 20 * h
    H, g, ub, x = np.empty((n, n)), np.empty(n), np.empty(m), np.empty(n)
    params = np.array([P.dt, P.mu, P.f_max, P.mass, P.alpha], np.float64)
    inertia = np.array(P.inertia[:], np.float32)
    weights = np.array(P.weights[:], np.float32)
    rows = [np.ascontiguousarray(batch[k][i], np.float32) for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")]
    rc = _REF.qr_ref_mpc_solve(h, _dp(params), _fp(inertia), _fp(weights), *[_fp(r) for r in rows],
                               _dp(H), _dp(g), _dp(ub), _dp(x))
    assert rc == 0
    return H, g, ub, x


def ref_mpc_time_batch(P: MpcParams, batch: dict, lo: int, hi: int, want_x: bool = False):
    """Times the reference's own SolveMPCKernel + GetMPCSolution(0..11) (oracle/_ref/libqr_mpc_ref.so) on
    problems [lo, hi) in this process, SetupProblem once.  Returns (seconds, lat, x12)."""
    global _REF
    if _REF is None:
        _REF = C.CDLL(os.path.join(_HERE, "_ref", "libqr_mpc_ref.so"))
    _REF.qr_ref_mpc_time_batch.restype = C.c_double
    cnt = hi - lo
    lat = np.empty(cnt)
    x = np.empty((cnt, 12)) if want_x else None
    params = np.array([P.dt, P.mu, P.f_max, P.mass, P.alpha], np.float64)
    inertia = np.array(P.inertia[:], np.float32)
    weights = np.array(P.weights[:], np.float32)
    arrs = [np.ascontiguousarray(batch[k][lo:hi], np.float32) for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")]
    sec = _REF.qr_ref_mpc_time_batch(P.horizon, _dp(params), _fp(inertia), _fp(weights), cnt, *[_fp(a) for a in arrs],
                                     _dp(x) if want_x else None, _dp(lat))
    return sec, lat, x


def mpc_qpoases(h: int, mu: float, H, g, ub, nWSR: int = 100000):
    """The reference's qpOASES call on float32 QP data.  Returns x, info(rval, nWSR), kkt, cstat."""
    n, m = 12 * h, 20 * h
    x = np.empty(n)
    info = np.zeros(2, np.int32)
    kkt = np.zeros(3)
    cstat = np.zeros(m, np.int32)
    H = np.ascontiguousarray(H, np.float32)
    g = np.ascontiguousarray(g, np.float32)
    ub = np.ascontiguousarray(ub, np.float32)
    lib().qro_mpc_qpoases(h, C.c_float(mu), _fp(H), _fp(g), _fp(ub), nWSR, _dp(x), _ip(info), _dp(kkt), _ip(cstat))
    return x, info, kkt, cstat


def qpoases_dense(H, g, A, lbA, ubA, nWSR: int = 100000):
    n, m = H.shape[0], A.shape[0]
    H, g, A = (np.ascontiguousarray(a, np.float64) for a in (H, g, A))
    lbA, ubA = (np.ascontiguousarray(a, np.float64) for a in (lbA, ubA))
    x = np.empty(n)
    info = np.zeros(2, np.int32)
    kkt = np.zeros(3)
    cstat = np.zeros(m, np.int32)
    lib().qro_qpoases_dense(n, m, _dp(H), _dp(g), _dp(A), _dp(lbA), _dp(ubA), nWSR, _dp(x), _ip(info), _dp(kkt), _ip(cstat))
    return x, info, kkt, cstat


def mpc_solve(P: MpcParams, batch: dict, i: int, nWSR: int = 100000):
    n = 12 * P.horizon
    x = np.empty(n)
    info = np.zeros(2, np.int32)
    rows = [np.ascontiguousarray(batch[k][i]) for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")]
    lib().qro_mpc_solve(C.byref(P), *[_fp(r) for r in rows], nWSR, _dp(x), _ip(info))
    return x, info


def mpc_time_batch(P: MpcParams, batch: dict, lo: int, hi: int, nWSR: int = 100, want_x: bool = False):
    """Times SolveMPC-equivalents for problems [lo, hi) in this process.  Returns (seconds, lat, capped, x)."""
    cnt = hi - lo
    n = 12 * P.horizon
    lat = np.empty(cnt)
    capped = C.c_int(0)
    x = np.empty((cnt, n)) if want_x else None
    arrs = [np.ascontiguousarray(batch[k][lo:hi]) for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")]
    sec = lib().qro_mpc_time_batch(C.byref(P), cnt, *[_fp(a) for a in arrs], nWSR,
                                   _dp(x) if want_x else None, _dp(lat), C.byref(capped))
    return sec, lat, capped.value, x


def contact_table(h, n_horizon_l, progress, duty, early=None, contacts=None):
    progress = np.ascontiguousarray(progress, np.float32)
    duty = np.ascontiguousarray(duty, np.float32)
    out = np.empty((h, 4), np.float32)
    e = None if early is None else np.ascontiguousarray(early, np.int32)
    c = None if contacts is None else np.ascontiguousarray(contacts, np.int32)
    lib().qro_mpc_contact_table(h, n_horizon_l, _fp(progress), _fp(duty), None if e is None else _ip(e),
                                None if c is None else _ip(c), _fp(out))
    return out


def reference_traj(h, dt, init, pos_xy):
    init = np.ascontiguousarray(init, np.float32)
    pos_xy = np.ascontiguousarray(pos_xy, np.float32)
    out = np.empty(12 * h, np.float32)
    lib().qro_mpc_reference_traj(h, C.c_float(dt), _fp(init), _fp(pos_xy), _fp(out))
    return out


# ---------------------------------------------------------------------------------------------
# Exact optimum of the reference's QP by an extended-precision active-set solve ("polish").
# ---------------------------------------------------------------------------------------------

def constraint_rows(h: int, mu: float):
    """Dense friction matrix (20h x 12h) as ResizeQPMats builds it (qr_mpc_interface.cpp:230-240)."""
    n, m = 12 * h, 20 * h
    A = np.zeros((m, n))
    mu_ = float(np.float32(1.0) / np.float32(mu))
    blk = np.array([[mu_, 0, 1], [-mu_, 0, 1], [0, mu_, 1], [0, -mu_, 1], [0, 0, 1]], float)
    for k in range(4 * h):
        A[5 * k:5 * k + 5, 3 * k:3 * k + 3] = blk
    return A


def polish_from_working_set(H, g, A, lb, ub, cstat, refine: int = 8):
    """x* = exact minimiser of the reference's QP on qpOASES' own final working set.

    The quadratic form x'Hx of the (slightly asymmetric, float32-built) H is that of Hs = (H+H')/2.
    Builds K = [[Hs, -Aa'],[Aa, 0]] from the active rows (cstat != 0; qpOASES keeps them linearly
    independent), solves in float64 and refines the residual in numpy.longdouble.
    Returns (x, multipliers of the active rows).
    """
    LD = np.longdouble
    Hs = (np.asarray(H, LD) + np.asarray(H, LD).T) / 2
    rows = np.nonzero(cstat)[0]
    Ar = np.asarray(A, LD)[rows]
    rhs = np.where(np.asarray(cstat)[rows] > 0, np.asarray(ub, LD)[rows], np.asarray(lb, LD)[rows])
    n, k = Hs.shape[0], len(rows)
    K = np.zeros((n + k, n + k), LD)
    K[:n, :n] = Hs
    K[:n, n:] = -Ar.T
    K[n:, :n] = Ar
    r = np.concatenate([-np.asarray(g, LD), rhs])
    Kd = K.astype(float)
    sol = np.linalg.lstsq(Kd, r.astype(float), rcond=1e-14)[0].astype(LD)
    for _ in range(refine):
        sol = sol + np.linalg.lstsq(Kd, (r - K @ sol).astype(float), rcond=1e-14)[0].astype(LD)
    return sol[:n].astype(float), sol[n:].astype(float)


def polish_fast(H, g, A, lb, ub, cstat, refine: int = 4):
    """Same x* as polish_from_working_set, for bulk use: ONE float64 LU factorisation of the KKT matrix of the
    working set (qpOASES keeps its rows linearly independent, H is positive definite, so K is non-singular) and
    residual refinement in numpy.longdouble against that factorisation.  Falls back to the least-squares version
    when the refinement does not contract (a dependent working set)."""
    import scipy.linalg as sla
    LD = np.longdouble
    Hs = (np.asarray(H, LD) + np.asarray(H, LD).T) / 2
    rows = np.nonzero(cstat)[0]
    Ar = np.asarray(A, LD)[rows]
    rhs = np.where(np.asarray(cstat)[rows] > 0, np.asarray(ub, LD)[rows], np.asarray(lb, LD)[rows])
    n, k = Hs.shape[0], len(rows)
    K = np.zeros((n + k, n + k), LD)
    K[:n, :n] = Hs
    K[:n, n:] = -Ar.T
    K[n:, :n] = Ar
    r = np.concatenate([-np.asarray(g, LD), rhs])
    try:
        lu = sla.lu_factor(K.astype(float), check_finite=False)
        sol = sla.lu_solve(lu, r.astype(float), check_finite=False).astype(LD)
        res0 = None
        for _ in range(refine):
            res = r - K @ sol
            sol = sol + sla.lu_solve(lu, res.astype(float), check_finite=False).astype(LD)
            res0 = float(np.abs(res).max()) if res0 is None else res0
        final = float(np.abs(r - K @ sol).max())
        if np.isfinite(final) and final <= 1e-12 * max(1.0, float(np.abs(r).max())):
            return sol[:n].astype(float), sol[n:].astype(float)
    except (ValueError, sla.LinAlgError):
        pass
    return polish_from_working_set(H, g, A, lb, ub, cstat)


def exact_optimum(H, g, A, lb, ub, cstat, max_changes: int = 60, tol: float = 1e-9):
    """x* = THE minimiser of the reference's QP (strictly convex, so unique), certified.

    Starts from converged qpOASES' final working set (cstat) and its extended-precision KKT solve (polish_fast).  That
    point is the exact optimum only if the working set is the optimal one; qpOASES stops on its homotopy-length
    tolerance, and on about one instance in a thousand the working set it stops with still misses a row that the exact
    KKT point violates by ~1e-3 N (or holds a row whose exact multiplier is slightly negative).  So the working set is
    corrected here by a plain primal-dual active-set continuation in extended precision -- add the most violated row,
    else drop the row with the most wrong-signed multiplier, re-solve -- until every row is feasible to `tol` and every
    multiplier has the right sign.  Returns (x, working set, changes made)."""
    cstat = np.array(cstat, np.int32, copy=True)
    lb = np.asarray(lb, float)
    ub = np.asarray(ub, float)
    A = np.asarray(A, float)
    eq = lb == ub
    for change in range(max_changes + 1):
        x, lam = polish_fast(H, g, A, lb, ub, cstat)
        rows = np.nonzero(cstat)[0]
        ax = A @ x
        viol_lo = np.where(cstat == 0, lb - ax, -np.inf)
        viol_up = np.where(cstat == 0, ax - ub, -np.inf)
        worst_lo, worst_up = int(viol_lo.argmax()), int(viol_up.argmax())
        if max(viol_lo[worst_lo], viol_up[worst_up]) > tol:
            if viol_lo[worst_lo] >= viol_up[worst_up]:
                cstat[worst_lo] = -1
            else:
                cstat[worst_up] = 1
            continue
        # Hs x + g = Ar' lam: a row at its lower bound needs lam >= 0, at its upper bound lam <= 0
        wrong = np.where(eq[rows], 0.0, np.where(cstat[rows] < 0, -lam, lam))
        k = int(wrong.argmax()) if len(rows) else -1
        if k >= 0 and wrong[k] > tol:
            cstat[rows[k]] = 0
            continue
        return x, cstat, change
    raise RuntimeError("exact_optimum: the working set did not settle")


def kkt_certificate(H, g, A, lb, ub, x, act_tol: float = 1e-7):
    """Independent optimality check of a candidate x for min 1/2 x'Hs x + g'x, lb <= Ax <= ub.

    Returns (stationarity, feasibility) where stationarity = min over multipliers of the right sign
    on the rows active at x of ||Hs x + g - A_act' y||_inf (non-negative least squares), and
    feasibility = the largest bound violation.  For a strictly convex QP both ~0 proves optimality:
    ||x - x*|| <= stationarity / lambda_min(Hs).
    """
    from scipy.optimize import nnls
    Hs = (np.asarray(H, float) + np.asarray(H, float).T) / 2
    x = np.asarray(x, float)
    ax = A @ x
    feas = float(max(0.0, (lb - ax).max(), (ax - ub).max()))
    scale = max(1.0, float(np.abs(ax).max()))
    lo = np.nonzero(ax - lb <= act_tol * scale)[0]
    up = np.nonzero(ub - ax <= act_tol * scale)[0]
    grad = Hs @ x + g
    # grad = sum_lo y_i a_i - sum_up z_i a_i with y, z >= 0
    M = np.concatenate([A[lo].T, -A[up].T], axis=1)
    if M.shape[1] == 0:
        return float(np.abs(grad).max()), feas
    y, _ = nnls(M, grad, maxiter=50 * M.shape[1])
    return float(np.abs(M @ y - grad).max()), feas


# ---------------------------------------------------------------------------------------------
# Whole-body control oracle
# ---------------------------------------------------------------------------------------------
class WbcModel(C.Structure):
    _fields_ = [("body_size", C.c_float * 3), ("hip_len", C.c_float), ("upper_len", C.c_float), ("lower_len", C.c_float)]


def wbc_model_of(robot) -> WbcModel:
    m = WbcModel()
    m.body_size[:] = robot.body_size
    m.hip_len, m.upper_len, m.lower_len = robot.hip_len, robot.upper_len, robot.lower_len
    return m


_REFWBC = None


def ref_wbc_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libqr_wbc_ref.so"))


def _ref_wbc():
    global _REFWBC
    if _REFWBC is None:
        _REFWBC = C.CDLL(os.path.join(_HERE, "_ref", "libqr_wbc_ref.so"))
    return _REFWBC


def wbc_step(model: WbcModel, state, cmd, contact, precision: str = "f64"):
    """One WBC tick for one robot (precision f64 | f32 = the restatement, ref = the reference's own
    classes compiled from /root/reference).  Returns dict(tau, fr, qdes, qddes, H, G, C, Jc, Jcdqd, pGC, vGC, qdd, rc)."""
    dt = np.float64 if precision == "f64" else np.float32
    ptr = _dp if precision == "f64" else _fp
    tau, fr, qdes, qddes = (np.zeros(12, dt) for _ in range(4))
    dbg = np.zeros(324 + 18 + 18 + 216 + 12 + 12 + 12 + 18, dt)
    state = np.ascontiguousarray(state, np.float32)
    cmd = np.ascontiguousarray(cmd, np.float32)
    contact = np.ascontiguousarray(contact, np.int32)
    if precision == "ref":   # the reference's own WBC classes (oracle/_ref/libqr_wbc_ref.so), float32
        fn = _ref_wbc().qr_ref_wbc_step
    else:
        fn = lib().qro_wbc_step_f64 if precision == "f64" else lib().qro_wbc_step_f32
    rc = fn(C.byref(model), _fp(state), _fp(cmd), _ip(contact), ptr(tau), ptr(fr), ptr(qdes), ptr(qddes), ptr(dbg))
    o = 0
    out = dict(tau=tau, fr=fr, qdes=qdes, qddes=qddes, rc=rc)
    for name, n, shape in (("H", 324, (18, 18)), ("G", 18, (18,)), ("C", 18, (18,)), ("Jc", 216, (4, 3, 18)),
                           ("Jcdqd", 12, (4, 3)), ("pGC", 12, (4, 3)), ("vGC", 12, (4, 3)), ("qdd", 18, (18,))):
        out[name] = dbg[o:o + n].reshape(shape).copy()
        o += n
    return out


def ref_wbc_time_batch(model: WbcModel, state, cmd, contact, want_tau: bool = False):
    """Times the reference's own WBC classes (oracle/_ref/libqr_wbc_ref.so: controller objects built once, then per
    robot UpdateModel + ContactTaskUpdate + FindConfiguration + MakeTorque, qr_wbc_locomotion_controller.cpp:108-134)
    on the rows of state / cmd / contact in this process.  Returns (seconds, lat[count], tau[count,12] or None)."""
    fn = _ref_wbc().qr_ref_wbc_time_batch
    fn.restype = C.c_double
    state = np.ascontiguousarray(state, np.float32)
    cmd = np.ascontiguousarray(cmd, np.float32)
    contact = np.ascontiguousarray(contact, np.int32)
    cnt = state.shape[0]
    lat = np.empty(cnt)
    tau = np.empty((cnt, 12), np.float32) if want_tau else None
    sec = fn(C.byref(model), cnt, _fp(state), _fp(cmd), _ip(contact), _fp(tau) if want_tau else None, _dp(lat))
    return sec, lat, tau


def swing_parabola(start, end, height, t, phase_module=False):
    start = np.ascontiguousarray(start, np.float32)
    end = np.ascontiguousarray(end, np.float32)
    pos = np.zeros(3, np.float32)
    ok = lib().qro_swing_parabola(_fp(start), _fp(end), C.c_float(height), C.c_float(t), int(phase_module), _fp(pos))
    return pos, bool(ok)


def grf_to_torque(robot, quat, q, f_world):
    quat = np.ascontiguousarray(quat, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    f_world = np.ascontiguousarray(f_world, np.float32)
    ff = np.zeros(12, np.float32)
    tau = np.zeros(12, np.float32)
    lib().qro_mpc_grf_to_torque(C.c_float(robot.hip_len), C.c_float(robot.upper_len), C.c_float(robot.lower_len),
                                _fp(quat), _fp(q), _fp(f_world), _fp(ff), _fp(tau))
    return ff, tau


# ------------------------------------------------------------------------------------------------
# Force-balance stance QP
# ------------------------------------------------------------------------------------------------
class FbParams(C.Structure):
    _fields_ = [("mass", C.c_float), ("inertia", C.c_float * 9), ("acc_weight", C.c_float * 6), ("reg_weight", C.c_float),
                ("mu", C.c_float), ("fmin_ratio", C.c_float * 4), ("fmax_ratio", C.c_float * 4), ("world_frame", C.c_int)]


def fb_params_of(p: dict) -> FbParams:
    P = FbParams()
    P.mass = p["mass"]
    P.inertia[:] = [float(v) for v in np.asarray(p["inertia"], np.float32).reshape(9)]
    P.acc_weight[:] = [float(v) for v in p["acc_weight"]]
    P.reg_weight, P.mu = p["reg_weight"], p["mu"]
    P.fmin_ratio[:] = [float(v) for v in p["fmin_ratio"]]
    P.fmax_ratio[:] = [float(v) for v in p["fmax_ratio"]]
    P.world_frame = int(p["world_frame"])
    return P


def force_balance(P: FbParams, foot, acc, contact, inertia=None, gravity=None, frame=None):
    """One robot through the restated ComputeContactForce + the reference's QuadProg++.
    Returns dict(force[12], status, cost, G[12,12], a[12], C[24,12], lb[24])."""
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    foot, acc, inertia, gravity, frame = f32(foot), f32(acc), f32(inertia), f32(gravity), f32(frame)
    contact = np.ascontiguousarray(contact, np.int32)
    force = np.zeros(12, np.float32)
    G, a, Cm, lb = np.zeros((12, 12), np.float32), np.zeros(12, np.float32), np.zeros((24, 12), np.float32), np.zeros(24, np.float32)
    cost = C.c_double()
    opt = lambda a: None if a is None else _fp(a)
    st = lib().qro_force_balance(C.byref(P), opt(inertia), _fp(foot), _fp(acc), _ip(contact), opt(gravity), opt(frame),
                                 _fp(force), _fp(G), _fp(a), _fp(Cm), _fp(lb), C.byref(cost))
    return dict(force=force, status=st, cost=cost.value, G=G, a=a, C=Cm, lb=lb)


# ------------------------------------------------------------------------------------------------
# WALK-mode swing trajectory (B-spline via the reference's tinynurbs) and heuristic foothold
# ------------------------------------------------------------------------------------------------
class FootholdParams(C.Structure):
    _fields_ = [("hip_offset", C.c_float * 12), ("hip_pos", C.c_float * 12), ("hip_len", C.c_float), ("swing_kp", C.c_float * 3)]


def foothold_params_of(p: dict) -> FootholdParams:
    P = FootholdParams()
    P.hip_offset[:] = [float(v) for v in np.asarray(p["hip_offset"], np.float32).reshape(12)]
    P.hip_pos[:] = [float(v) for v in np.asarray(p["hip_pos"], np.float32).reshape(12)]
    P.hip_len = p["hip_len"]
    P.swing_kp[:] = [float(v) for v in p["swing_kp"]]
    return P


def swing_bspline(initial_pos, target_pos, height, duration, initial_time, time):
    ip, tp = np.ascontiguousarray(initial_pos, np.float32), np.ascontiguousarray(target_pos, np.float32)
    pos, vel = np.zeros(3, np.float32), np.zeros(3, np.float32)
    ok = lib().qro_swing_bspline(_fp(ip), _fp(tp), C.c_float(height), C.c_float(duration), C.c_float(initial_time),
                                 C.c_float(time), _fp(pos), _fp(vel))
    return pos, vel, bool(ok)


def foothold(P: FootholdParams, leg, b: dict, i: int):
    """Robot i, one leg, of a make_foothold_batch dict: (foothold[3], phase)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    fh, ph = np.zeros(12, np.float32), C.c_float()
    lib().qro_foothold(C.byref(P), int(leg), _fp(f32(b["com_vel"][i])), _fp(f32(b["rpy_rate"][i])), _fp(f32(b["dR"][i])),
                       _fp(f32(b["base_R"][i])), _fp(f32(b["rpy"][i])), _fp(f32(b["foot_base"][i])), _fp(f32(b["des_speed"][i])),
                       C.c_float(b["des_twist"][i]), C.c_float(b["des_height"][i]), C.c_float(b["swing_remain"][i, leg]),
                       int(b["allow_switch"][i, leg]), C.c_float(b["norm_phase"][i, leg]), _fp(fh), C.byref(ph))
    return fh[3 * leg:3 * leg + 3].copy(), ph.value


def quadprog_ineq(G, g0, Cm, c0):
    """min 1/2 x'Gx + g0'x s.t. Cm x + c0 >= 0 through the reference's QuadProg++: (x, cost)."""
    G, g0, Cm, c0 = (np.ascontiguousarray(a, np.float64) for a in (G, g0, Cm, c0))
    n, m = G.shape[0], Cm.shape[0]
    x = np.zeros(n)
    f = lib().qro_quadprog_ineq
    f.restype = C.c_double
    cost = f(n, m, _dp(G), _dp(g0), _dp(Cm), _dp(c0), _dp(x))
    return x, cost


# ------------------------------------------------------------------------------------------------
# Controller code compiled from the reference itself (oracle/_ref/libqr_ctl_ref.so, ref_ctl_shim.cpp): leg kinematics,
# contact table / reference trajectory, SolveDenseMPC (lever arms, f_ff), swing parabola, foothold heuristic,
# MPC-mode swing targets, open-loop gait phase, force-balance QP.
# ------------------------------------------------------------------------------------------------
_REFCTL = None


def ref_ctl_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libqr_ctl_ref.so"))


def _ref_ctl():
    global _REFCTL
    if _REFCTL is None:
        _REFCTL = C.CDLL(os.path.join(_HERE, "_ref", "libqr_ctl_ref.so"))
    return _REFCTL


def robot_geom(robot) -> np.ndarray:
    """geom[18] of ref_ctl_shim.cpp: hip_len, upper_len, lower_len, hipOffset (3x4 column-major = abad positions), comOffset."""
    hips = np.array(robot.hip_positions, np.float64)
    abad = hips.copy()
    abad[:, 1] -= np.sign(hips[:, 1]) * robot.hip_len
    return np.concatenate([[robot.hip_len, robot.upper_len, robot.lower_len], abad.reshape(12), robot.com_offset]).astype(np.float32)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, np.int32)


def _opt_fp(a):
    return None if a is None else _fp(a)


def ref_leg_kinematics(robot, q, qd=None, f_leg=None):
    """qrRobot leg kinematics (src/robots/qr_robot.cpp:106-251) for one robot: dict(foot_base[12] column-major 3x4, jac[4,3,3],
    foot_vel[12], tau[12], ik_q[12], ik_qd[12])."""
    geom = robot_geom(robot)
    q, qd, f_leg = _f32(q), _f32(qd if qd is not None else np.zeros(12)), _f32(f_leg if f_leg is not None else np.zeros(12))
    out = {k: np.zeros(n, np.float32) for k, n in (("foot_base", 12), ("jac", 36), ("foot_vel", 12), ("tau", 12), ("ik_q", 12), ("ik_qd", 12))}
    _ref_ctl().qr_ref_leg_kinematics(_fp(geom), _fp(q), _fp(qd), _fp(f_leg), _fp(out["foot_base"]), _fp(out["jac"]),
                                     _fp(out["foot_vel"]), _fp(out["tau"]), _fp(out["ik_q"]), _fp(out["ik_qd"]))
    out["jac"] = out["jac"].reshape(4, 3, 3)
    return out


def ref_mpc_inputs(h, n_horizon_l, dt, progress, duty, leg_state=None, contacts=None, init=None, base_xy=None):
    """mpcTable [h,4] and trajAll [12h] from the reference's own lines (qr_mpc_stance_leg_controller.cpp:282-303, 344-376)."""
    progress, duty = _f32(progress), _f32(duty)
    ls, ct = _i32(leg_state), _i32(contacts)
    table = np.zeros(4 * h, np.float32)
    traj = np.zeros(12 * h, np.float32) if init is not None else None
    init, base_xy = _f32(init), _f32(base_xy)
    _ref_ctl().qr_ref_mpc_inputs(h, n_horizon_l, C.c_float(dt), _fp(progress), _fp(duty), None if ls is None else _ip(ls),
                                 None if ct is None else _ip(ct), _opt_fp(init), _opt_fp(base_xy), _fp(table), _opt_fp(traj))
    return table.reshape(h, 4), traj


def ref_solve_dense_mpc(P: MpcParams, robot, rpy, pos, quat, v_world, w_world, foot_base, traj, table):
    """SolveDenseMPC (qr_mpc_stance_leg_controller.cpp:385-410) through the reference's own SolveMPCKernel / GetMPCSolution:
    dict(lever[12] column-major 3x4, f[12], f_ff[12], fr_des[12])."""
    setup = np.array([P.dt, P.mu, P.f_max, P.mass, P.alpha] + list(P.inertia[:]) + list(P.weights[:]), np.float64)
    geom = robot_geom(robot)
    args = [_f32(a) for a in (rpy, pos, quat, v_world, w_world, foot_base, traj, np.asarray(table).reshape(-1))]
    out = {k: np.zeros(12, np.float32) for k in ("lever", "f", "f_ff", "fr_des")}
    _ref_ctl().qr_ref_solve_dense_mpc(P.horizon, _dp(setup), _fp(geom), *[_fp(a) for a in args], _fp(out["lever"]), _fp(out["f"]),
                                      _fp(out["f_ff"]), _fp(out["fr_des"]))
    return out


def ref_swing_parabola(start, end, height, t, phase_module=False):
    start, end = _f32(start), _f32(end)
    pos, vel, acc = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    ok = _ref_ctl().qr_ref_swing_parabola(_fp(start), _fp(end), C.c_float(height), C.c_float(t), int(phase_module), _fp(pos), _fp(vel), _fp(acc))
    return pos, vel, acc, bool(ok)


def ref_foothold(robot, params: dict, b: dict, i: int, foothold_io, phase_io, q=None):
    """ComputeHeuristicFootHold (qr_foothold_planner.cpp:112-240) for robot i of a make_foothold_batch dict."""
    geom = robot_geom(robot)
    geom[3:15] = np.asarray(params["hip_offset"], np.float32)
    state_des = np.zeros(12, np.float32)
    state_des[6:9] = b["des_speed"][i]
    state_des[11] = b["des_twist"][i]
    clearance = 0.01
    state_des[2] = b["des_height"][i] + np.float32(clearance)
    fh, ph = np.array(foothold_io, np.float32, copy=True), np.array(phase_io, np.float32, copy=True)
    _ref_ctl().qr_ref_foothold(_fp(geom), _fp(_f32(params["hip_pos"])), _fp(_f32(params["swing_kp"])), _fp(_f32(b["com_vel"][i])),
                               _fp(_f32(b["rpy_rate"][i])), _fp(_f32(b["dR"][i])), _fp(_f32(b["base_R"][i])), _fp(_f32(b["rpy"][i])),
                               _fp(_f32(b["foot_base"][i])), _opt_fp(_f32(q)), _fp(state_des), C.c_float(clearance),
                               _fp(_f32(b["swing_remain"][i])), _fp(_f32(b["norm_phase"][i])), _ip(_i32(b["allow_switch"][i])),
                               _ip(_i32(b["swing_mask"][i])), _fp(fh), _fp(ph))
    return fh, ph


def ref_swing_targets(robot, base_pos, quat, v_world, foothold, planner_phase, switch_pos, swing_duration, swing_mask,
                      horizontal_terrain=True):
    """MPC-mode swing targets (qr_swing_leg_controller.cpp:361-409, 417-420): dict(p_foot_des, v_foot_des, a_foot_des,
    foot_base_des, q_des, qd_des), [12] each, rows of stance legs left at zero."""
    geom = robot_geom(robot)
    out = {k: np.zeros(12, np.float32) for k in ("p_foot_des", "v_foot_des", "a_foot_des", "foot_base_des", "q_des", "qd_des")}
    _ref_ctl().qr_ref_swing_targets(_fp(geom), _fp(_f32(base_pos)), _fp(_f32(quat)), _fp(_f32(v_world)), _fp(_f32(foothold)),
                                    _fp(_f32(planner_phase)), _fp(_f32(switch_pos)), _fp(_f32(swing_duration)), _ip(_i32(swing_mask)),
                                    int(horizontal_terrain), *[_fp(out[k]) for k in ("p_foot_des", "v_foot_des", "a_foot_des", "foot_base_des", "q_des", "qd_des")])
    return out


def ref_gait_update(t, cfg, contacts, istate, fstate, out, contact_threshold=0.1, stop=False, advanced_trot=False):
    """One qrOpenLoopGaitGenerator::Update(t) (qr_openloop_gait_generator.cpp:126-247) on caller-held state; returns updated
    copies (istate[20] int32, fstate[4], out[12], allow[4])."""
    istate, fstate, out = np.array(istate, np.int32, copy=True), np.array(fstate, np.float32, copy=True), np.array(out, np.float32, copy=True)
    allow = np.zeros(4, np.int32)
    _ref_ctl().qr_ref_gait_update(C.c_float(t), _fp(_f32(cfg)), C.c_float(contact_threshold), _ip(_i32(contacts)), int(stop),
                                  int(advanced_trot), _ip(istate), _fp(fstate), _fp(out), _ip(allow))
    return istate, fstate, out, allow


def ref_contact_force_world(params: dict, quat, foot_base, acc, contact, n=(0, 0, 1), t1=(1, 0, 0), t2=(0, 1, 0)):
    """Quadruped::ComputeContactForce, world-frame overload (qr_qp_torque_optimizer.cpp:304-400) with the reference's own
    matrix builders and QuadProg++.  Returns the 3x4 result as [12] column-major (leg columns)."""
    out = np.zeros(12, np.float32)
    _ref_ctl().qr_ref_contact_force_world(C.c_float(params["mass"]), _fp(_f32(np.asarray(params["inertia"]).reshape(9))), _fp(_f32(quat)),
                                          _fp(_f32(foot_base)), _fp(_f32(acc)), _ip(_i32(contact)), _fp(_f32(n)), _fp(_f32(t1)),
                                          _fp(_f32(t2)), _fp(_f32(params["acc_weight"])), _fp(_f32(params["fmin_ratio"])),
                                          _fp(_f32(params["fmax_ratio"])), C.c_float(params["reg_weight"]), C.c_float(params["mu"]), _fp(out))
    return out


def ref_contact_force_control(params: dict, quat, foot_base, acc, contact, terrain_type=0, control_rpy=(0, 0, 0), aligned=None):
    """Quadruped::ComputeContactForce, control-frame overload (qr_qp_torque_optimizer.cpp:190-301) compiled from the reference.
    Returns (F[12] = (X Rcb)^T column-major, dict(Rcb[3,3], inertia[9], foot[12] rows of the 4x3 footPosition, gravity[3], normal[3]))."""
    out, der = np.zeros(12, np.float32), np.zeros(36, np.float32)
    aligned = np.eye(3, dtype=np.float32) if aligned is None else _f32(aligned)
    _ref_ctl().qr_ref_contact_force_control(C.c_float(params["mass"]), _fp(_f32(np.asarray(params["inertia"]).reshape(9))), _fp(_f32(quat)),
                                            _fp(_f32(foot_base)), _fp(_f32(acc)), _ip(_i32(contact)), int(terrain_type), _fp(_f32(control_rpy)),
                                            _fp(_f32(np.asarray(aligned).reshape(9))), _fp(_f32(params["acc_weight"])),
                                            C.c_float(params["fmin_ratio"][0]), C.c_float(params["fmax_ratio"][0]),
                                            C.c_float(params["reg_weight"]), C.c_float(params["mu"]), _fp(out), _fp(der))
    return out, dict(Rcb=der[:9].reshape(3, 3).copy(), inertia=der[9:18].copy(), foot=der[18:30].copy(), gravity=der[30:33].copy(),
                     normal=der[33:36].copy())
