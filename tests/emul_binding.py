"""Builds and binds tests/emul/libqr_emul.so -- the device sources compiled for the host (tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emul", "emul.cpp")
_SO = os.path.join(_HERE, "emul", "libqr_emul.so")
_CSRC = os.path.join(_HERE, "..", "quadruped-robot_b200", "csrc")
_KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")


def _stale():
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


class Emul:
    def __init__(self, lib):
        self.lib = lib

    @staticmethod
    def _fp(a):
        return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))

    def condense(self, P, batch):
        B, h = batch["p"].shape[0], P.horizon
        n = 12 * h
        H = np.zeros((B, n, n), np.float32)
        g = np.zeros((B, n), np.float32)
        ub = np.zeros((B, 20 * h), np.float32)
        self.lib.qr_emul_mpc_condense_batch(C.byref(P), B, *[self._fp(batch[k]) for k in _KEYS], None,
                                            self._fp(H), self._fp(g), self._fp(ub))
        return H, g, ub

    def condense_pair_mismatches(self, P, batch):
        B = batch["p"].shape[0]
        return int(self.lib.qr_emul_condense_pair_mismatches(C.byref(P), B, *[self._fp(batch[k]) for k in _KEYS]))

    def solve(self, P, batch, opt=None, per_instance_mu=False):
        B, h = batch["p"].shape[0], P.horizon
        n = 12 * h
        grf = np.zeros((B, 12), np.float32)
        u = np.zeros((B, n), np.float32)
        u64 = np.zeros((B, n))
        st = np.zeros(B, np.int32)
        it = np.zeros((B, 2), np.int32)
        self.lib.qr_emul_mpc_solve_batch(
            C.byref(P), C.byref(opt) if opt is not None else None, B, *[self._fp(batch[k]) for k in _KEYS],
            self._fp(batch["mu"]) if per_instance_mu else None, None, self._fp(grf), self._fp(u),
            u64.ctypes.data_as(C.POINTER(C.c_double)), st.ctypes.data_as(C.POINTER(C.c_int)),
            it.ctypes.data_as(C.POINTER(C.c_int)))
        return dict(grf=grf, u=u, u64=u64, status=st, iters=it)

    def qp_solve(self, h, mu, H, g, ub, opt=None):
        B, n = H.shape[0], 12 * h
        x64 = np.zeros((B, n))
        st = np.zeros(B, np.int32)
        it = np.zeros((B, 2), np.int32)
        H = np.ascontiguousarray(H, np.float32)
        g = np.ascontiguousarray(g, np.float32)
        ub = np.ascontiguousarray(ub, np.float32)
        self.lib.qr_emul_qp_solve_batch(h, C.c_float(mu), C.byref(opt) if opt is not None else None, B,
                                        self._fp(H), self._fp(g), self._fp(ub), None, None,
                                        x64.ctypes.data_as(C.POINTER(C.c_double)),
                                        st.ctypes.data_as(C.POINTER(C.c_int)), it.ctypes.data_as(C.POINTER(C.c_int)))
        return dict(x64=x64, status=st, iters=it)


class WbcModelC(C.Structure):
    _fields_ = [("body_size", C.c_float * 3), ("hip_len", C.c_float), ("upper_len", C.c_float), ("lower_len", C.c_float)]


def _wbc_model(robot):
    m = WbcModelC()
    m.body_size[:] = robot.body_size
    m.hip_len, m.upper_len, m.lower_len = robot.hip_len, robot.upper_len, robot.lower_len
    return m


def wbc_solve(self, batch):
    """Host emulation of the WBC device code: dict(tau, fr, qdes, qddes, dbg, status) in float64."""
    B = batch["state"].shape[0]
    tau, fr, qdes, qddes = (np.zeros((B, 12)) for _ in range(4))
    dbg = np.zeros((B, 630))
    st = np.zeros(B, np.int32)
    m = _wbc_model(batch["robot"])
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    self.lib.qr_emul_wbc_solve_batch(C.byref(m), B, self._fp(batch["state"]), self._fp(batch["cmd"]),
                                     batch["contact"].ctypes.data_as(C.POINTER(C.c_int)), dp(tau), dp(fr), dp(qdes),
                                     dp(qddes), dp(dbg), st.ctypes.data_as(C.POINTER(C.c_int)))
    return dict(tau=tau, fr=fr, qdes=qdes, qddes=qddes, dbg=dbg, status=st)


def swing_parabola(self, start, end, height, t, phase_module=False):
    start = np.ascontiguousarray(start, np.float32)
    end = np.ascontiguousarray(end, np.float32)
    pos = np.zeros(3, np.float32)
    ok = self.lib.qr_emul_swing_parabola(self._fp(start), self._fp(end), C.c_float(height), C.c_float(t),
                                         int(phase_module), self._fp(pos))
    return pos, bool(ok)


def force_balance(self, P, batch):
    """Host emulation of the force-balance device code: dict(force[B,12], status, iters)."""
    B = batch["foot"].shape[0]
    force = np.zeros((B, 12), np.float32)
    st, it = np.zeros(B, np.int32), np.zeros(B, np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    self.lib.qr_emul_force_balance_batch(C.byref(P), B, self._fp(batch.get("inertia")), self._fp(batch["foot"]),
                                         self._fp(batch["acc"]), ip(batch["contact"]), self._fp(batch.get("gravity")),
                                         self._fp(batch.get("frame")), self._fp(force), ip(st), ip(it))
    return dict(force=force, status=st, iters=it)


def fb_build(self, P, batch, i):
    """float32 QP data of robot i as the device code builds them: G[12,12], a[12], C[24,12], lb[24]."""
    G, a, Cm, lb = np.zeros((12, 12), np.float32), np.zeros(12, np.float32), np.zeros((24, 12), np.float32), np.zeros(24, np.float32)
    row = lambda k: None if batch.get(k) is None else np.ascontiguousarray(batch[k][i])
    self.lib.qr_emul_fb_build(C.byref(P), self._fp(row("inertia")), self._fp(row("foot")), self._fp(row("acc")),
                              row("contact").ctypes.data_as(C.POINTER(C.c_int)), self._fp(row("gravity")),
                              self._fp(row("frame")), self._fp(G), self._fp(a), self._fp(Cm), self._fp(lb))
    return G, a, Cm, lb


def swing_bspline(self, ip, tp, height, duration, t0, t):
    ip, tp = np.ascontiguousarray(ip, np.float32), np.ascontiguousarray(tp, np.float32)
    pos, vel = np.zeros(3, np.float32), np.zeros(3, np.float32)
    ok = self.lib.qr_emul_swing_bspline(self._fp(ip), self._fp(tp), C.c_float(height), C.c_float(duration), C.c_float(t0),
                                        C.c_float(t), self._fp(pos), self._fp(vel))
    return pos, vel, bool(ok)


def foothold(self, P, leg, b, i):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    fh, ph = np.zeros(12, np.float32), C.c_float()
    self.lib.qr_emul_foothold(C.byref(P), int(leg), self._fp(f32(b["com_vel"][i])), self._fp(f32(b["rpy_rate"][i])),
                              self._fp(f32(b["dR"][i])), self._fp(f32(b["base_R"][i])), self._fp(f32(b["rpy"][i])),
                              self._fp(f32(b["foot_base"][i])), self._fp(f32(b["des_speed"][i])), C.c_float(b["des_twist"][i]),
                              C.c_float(b["des_height"][i]), C.c_float(b["swing_remain"][i, leg]), int(b["allow_switch"][i, leg]),
                              C.c_float(b["norm_phase"][i, leg]), self._fp(fh), C.byref(ph))
    return fh[3 * leg:3 * leg + 3].copy(), ph.value


def small_qp(self, G, g0, Cm, c0):
    G, g0, Cm, c0 = (np.ascontiguousarray(a, np.float64) for a in (G, g0, Cm, c0))
    n, m = G.shape[0], Cm.shape[0]
    x = np.zeros(n)
    it = C.c_int()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    st = self.lib.qr_emul_small_qp(n, m, dp(G), dp(g0), dp(Cm), dp(c0), dp(x), C.byref(it))
    return x, st, it.value


def solve_ex(self, P, batch, robot, q, cmd):
    """The fused solve with the leg-force / torque / Fr_des epilogue: dict(grf, f_ff, tau, cmd, status)."""
    B = batch["p"].shape[0]
    grf, ff, tau = (np.zeros((B, 12), np.float32) for _ in range(3))
    st = np.zeros(B, np.int32)
    q = np.ascontiguousarray(q, np.float32)
    cmd = np.ascontiguousarray(cmd, np.float32)
    self.lib.qr_emul_mpc_solve_batch_ex(C.byref(P), B, *[self._fp(batch[k]) for k in _KEYS], C.c_float(robot.hip_len),
                                        C.c_float(robot.upper_len), C.c_float(robot.lower_len), self._fp(q), self._fp(grf),
                                        self._fp(ff), self._fp(tau), self._fp(cmd), st.ctypes.data_as(C.POINTER(C.c_int)))
    return dict(grf=grf, f_ff=ff, tau=tau, cmd=cmd, status=st)


Emul.solve_ex = solve_ex
Emul.small_qp = small_qp
Emul.swing_bspline = swing_bspline
Emul.foothold = foothold
Emul.wbc_solve = wbc_solve
Emul.swing_parabola = swing_parabola
Emul.force_balance = force_balance
Emul.fb_build = fb_build


def load():
    if _stale():
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-shared", "-o", _SO, _SRC],
                       check=True)
    return Emul(C.CDLL(_SO))
