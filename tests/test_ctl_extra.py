"""Controller arithmetic around the two solvers (csrc/ctl_extra.h): lever arms, leg FK / Jacobian / IK, the MPC-mode
swing targets written into the WBC command rows, the open-loop gait phase, and the fused leg-torque epilogue of the MPC
kernel.  The checker is the REFERENCE ITSELF: oracle/_ref/libqr_ctl_ref.so holds the functions compiled from
/root/reference (oracle/ref_ctl_shim.cpp lists file:line of each).  CPU tests run the host build of the device sources
(tests/emul); the GPU tests run the kernels through the C ABI."""
import ctypes as C

import numpy as np
import pytest

F32 = np.float32


def fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_ctl_available():
        pytest.skip("oracle/_ref/libqr_ctl_ref.so not built (needs /root/reference)")
    return oracle


def _geom(pkg, robot):
    from quadruped_robot_b200 import capi
    return capi.leg_geometry_of(robot)


def swing_inputs(pkg, robot, B, seed):
    """Robot states with planned footholds and lift-off positions for the swing legs."""
    rng = np.random.default_rng(seed)
    U = rng.uniform
    wb = pkg.synth.make_wbc_batch(robot, B, seed=seed)
    rb = wb["robot"]
    quat, pos = wb["state"][:, :4].copy(), wb["state"][:, 4:7].copy()
    v_world = U(-0.5, 1.0, (B, 3)).astype(F32)
    hips = np.array(rb.hip_positions, float)
    foothold = (hips[None] + np.array([0, 0, -rb.body_height]) + U(-0.1, 0.1, (B, 4, 3))).astype(F32).reshape(B, 12)
    switch = (hips[None] + np.array([0, 0, -rb.body_height]) + U(-0.08, 0.08, (B, 4, 3))).astype(F32).reshape(B, 12)
    phase = U(-0.002, 1.002, (B, 4)).astype(F32)
    phase[0] = [0.0, 1.0, 0.5, -0.0005]
    phase[1] = [-0.0015, 1.0015, 0.3, 1.0005]      # two phases the trajectory generator rejects
    dur = U(0.15, 0.3, (B, 4)).astype(F32)
    mask = rng.integers(0, 2, (B, 4)).astype(np.int32)
    mask[:, 0] = 1
    mask[:2] = 1
    return dict(robot=rb, quat=quat, pos=pos, v_world=v_world, foothold=foothold, switch=switch, phase=phase, dur=dur, mask=mask,
                cmd=wb["cmd"].copy())


def gait_cfg(pkg, B, seed, gait="trot"):
    rng = np.random.default_rng(seed)
    gt = pkg.robots.GAITS[gait]
    duty, stance = F32(gt["duty"]), F32(gt["stance_duration"])
    period = stance / duty
    cfg = np.zeros((B, 4, 5), F32)
    cfg[:, :, 0] = np.array(gt["offsets"], F32)[None, :]
    cfg[:, :, 1] = period
    cfg[:, :, 2] = duty
    cfg[:, :, 3] = period - stance
    cfg[:, :, 4] = duty
    istate = np.ones((B, 5, 4), np.int32)          # every leg starts in STANCE (initialLegState), flags 0
    istate[:, 4, :] = 0
    fstate = np.zeros((B, 4), F32)
    fstate[:, 3] = 1.0                              # waitTime
    out = np.zeros((B, 12), F32)
    t0 = rng.uniform(0, 0.2, B).astype(F32)
    return cfg.reshape(B, 20), istate.reshape(B, 20), fstate, out, t0


# ------------------------------------------------------------------------------------------------ CPU (host build)
def test_lever_arms_and_leg_kinematics_match_reference(ref, emul, pkg):
    rng = np.random.default_rng(70)
    for name in ("a1", "lite3"):
        rb = pkg.robots.ROBOTS[name]
        G = _geom(pkg, rb)
        b = pkg.synth.make_wbc_batch(name, 128, seed=71)
        com = np.array(rb.com_offset, F32)
        exact_fk = exact_ik = 0
        for i in range(128):
            q, qd, quat = b["state"][i, 13:25].copy(), b["state"][i, 25:37].copy(), b["state"][i, :4].copy()
            k = ref.ref_leg_kinematics(rb, q, qd)
            fb, jac, fv, iq, iqd = (np.zeros(n, F32) for n in (12, 36, 12, 12, 12))
            emul.lib.qr_emul_leg_kinematics(C.byref(G), fp(q), fp(qd), fp(fb), fp(jac), fp(fv), fp(iq), fp(iqd))
            np.testing.assert_allclose(fb, k["foot_base"], rtol=0, atol=3e-7)
            np.testing.assert_allclose(jac.reshape(4, 3, 3), k["jac"], rtol=0, atol=3e-7)
            np.testing.assert_allclose(fv, k["foot_vel"], rtol=0, atol=2e-6)
            np.testing.assert_allclose(iq, k["ik_q"], rtol=0, atol=3e-6)
            np.testing.assert_allclose(iqd, k["ik_qd"], rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(iq, q, rtol=0, atol=5e-6)     # IK(FK(q)) = q
            exact_fk += np.array_equal(fb, k["foot_base"])
            exact_ik += np.array_equal(iq, k["ik_q"])
            # lever arms against the reference's SolveDenseMPC expression (its table/trajectory do not matter here)
            r = np.zeros(12, F32)
            emul.lib.qr_emul_lever_arms(fp(quat), fp(fb), fp(com), fp(r))
            Rb = np.array(_base_rmat(quat), F32)
            want = np.zeros(12, F32)
            for leg in range(4):
                d = (fb[3 * leg:3 * leg + 3] - com).astype(F32)
                for a in range(3):
                    want[3 * leg + a] = F32(F32(F32(Rb[a, 0] * d[0]) + F32(Rb[a, 1] * d[1])) + F32(Rb[a, 2] * d[2]))
            assert np.array_equal(r, want)
        # sin / cos / sqrt: glibc's float functions are correctly rounded in all but rare ties, so the forward kinematics
        # agrees bit for bit almost always; its acosf / asinf / atan2f are only accurate to an ulp, so the inverse
        # kinematics agrees to the tolerance above (3e-6 rad), not bitwise
        assert exact_fk >= 100


def _base_rmat(q):
    e0, e1, e2, e3 = [F32(v) for v in q]
    two, one = F32(2), F32(1)
    return [[one - two * (e2 * e2 + e3 * e3), two * (e1 * e2 - e0 * e3), two * (e1 * e3 + e0 * e2)],
            [two * (e1 * e2 + e0 * e3), one - two * (e1 * e1 + e3 * e3), two * (e2 * e3 - e0 * e1)],
            [two * (e1 * e3 - e0 * e2), two * (e2 * e3 + e0 * e1), one - two * (e1 * e1 + e2 * e2)]]


def test_lever_arms_match_reference_solve_dense_mpc(ref, emul, pkg, monkeypatch):
    monkeypatch.setenv("MINI_EIGEN_EXP_NILPOTENT3", "1")
    h, dt = 5, 0.06
    rb = pkg.robots.ROBOTS["lite3"]
    b = pkg.synth.make_mpc_batch("lite3", h, dt, 4, seed=72, gait="trot")
    P = ref.params_of(rb, h, dt)
    rng = np.random.default_rng(73)
    com = np.array(rb.com_offset, F32)
    for i in range(4):
        foot_base = (np.array(rb.hip_positions) + np.array([0, 0, -rb.body_height]) + rng.uniform(-0.05, 0.05, (4, 3))).astype(F32).reshape(12)
        o = ref.ref_solve_dense_mpc(P, rb, b["rpy"][i], b["p"][i], b["quat"][i], b["v"][i], b["w"][i], foot_base, b["traj"][i], b["gait"][i])
        r = np.zeros(12, F32)
        emul.lib.qr_emul_lever_arms(fp(b["quat"][i]), fp(foot_base), fp(com), fp(r))
        assert np.array_equal(r, o["lever"])


def test_swing_targets_match_reference(ref, emul, pkg):
    for name, horizontal in (("a1", True), ("lite3", False)):
        s = swing_inputs(pkg, name, 96, 74)
        rb = s["robot"]
        G = _geom(pkg, rb)
        n_rejected = 0
        for i in range(96):
            cmd = s["cmd"][i].copy()
            fbd, qd_, qdd = (np.full(12, -7, F32) for _ in range(3))
            valid = np.zeros(4, np.int32)
            emul.lib.qr_emul_swing_targets(C.byref(G), fp(s["pos"][i]), fp(s["quat"][i]), fp(s["v_world"][i]), fp(s["foothold"][i]),
                                           fp(s["phase"][i]), fp(s["switch"][i]), fp(s["dur"][i]), ip(s["mask"][i]), int(horizontal),
                                           fp(cmd), fp(fbd), fp(qd_), fp(qdd), ip(valid))
            # the reference keeps going with stale locals when its generator rejects a phase; compare the accepted legs
            ok_mask = np.array([int(s["mask"][i, l] and -1e-3 <= s["phase"][i, l] < 1 + 1e-3) for l in range(4)], np.int32)
            assert np.array_equal(valid, ok_mask)
            n_rejected += int((s["mask"][i] > ok_mask).sum())
            o = ref.ref_swing_targets(rb, s["pos"][i], s["quat"][i], s["v_world"][i], s["foothold"][i], s["phase"][i], s["switch"][i],
                                      s["dur"][i], ok_mask, horizontal)
            for l in range(4):
                sl = slice(3 * l, 3 * l + 3)
                if ok_mask[l]:
                    np.testing.assert_allclose(cmd[15 + 3 * l:18 + 3 * l], o["p_foot_des"][sl], rtol=0, atol=2e-6)
                    assert np.array_equal(cmd[27 + 3 * l:30 + 3 * l], o["v_foot_des"][sl])
                    assert np.array_equal(cmd[39 + 3 * l:42 + 3 * l], o["a_foot_des"][sl])
                    np.testing.assert_allclose(fbd[sl], o["foot_base_des"][sl], rtol=0, atol=1e-6)
                    np.testing.assert_allclose(qd_[sl], o["q_des"][sl], rtol=0, atol=2e-5)
                    np.testing.assert_allclose(qdd[sl], o["qd_des"][sl], rtol=0, atol=1e-5)
                else:
                    assert np.array_equal(cmd[15 + 3 * l:18 + 3 * l], s["cmd"][i][15 + 3 * l:18 + 3 * l]) and (fbd[sl] == -7).all()
        assert n_rejected > 0


@pytest.mark.parametrize("advanced", [False, True])
def test_gait_update_matches_reference_over_many_ticks(ref, emul, pkg, advanced):
    B = 24
    cfg, istate, fstate, out, t0 = gait_cfg(pkg, B, 75)
    rng = np.random.default_rng(76)
    r_i, r_f, r_o = istate.copy(), fstate.copy(), out.copy()
    e_i, e_f, e_o = istate.copy(), fstate.copy(), out.copy()
    seen = set()
    for tick in range(400):
        for i in range(B):
            t = F32(t0[i] + F32(0.002) * F32(tick))
            # contacts follow the planned state most of the time, with early touch-downs and late lift-offs mixed in
            planned = (r_i[i, 8:12] == 1).astype(np.int32)
            noise = rng.uniform(size=4) < 0.08
            contacts = np.where(noise, 1 - planned, planned).astype(np.int32)
            stop = int(rng.uniform() < 0.02)
            ri, rf, ro, ra = ref.ref_gait_update(t, cfg[i], contacts, r_i[i], r_f[i], r_o[i], stop=bool(stop), advanced_trot=advanced)
            ei, ef, eo, ea = e_i[i].copy(), e_f[i].copy(), e_o[i].copy(), np.zeros(4, np.int32)
            emul.lib.qr_emul_gait_update(C.c_float(t), fp(cfg[i]), C.c_float(0.1), ip(contacts), stop, int(advanced), ip(ei), fp(ef), fp(eo), ip(ea))
            assert np.array_equal(ei, ri) and np.array_equal(ea, ra), (tick, i)
            assert np.array_equal(ef, rf) and np.array_equal(eo, ro), (tick, i)
            r_i[i], r_f[i], r_o[i] = ri, rf, ro
            e_i[i], e_f[i], e_o[i] = ei, ef, eo
            seen.update(int(v) for v in ri[12:16])
    assert {0, 1, 2} <= seen    # SWING, STANCE and EARLY_CONTACT all occurred


def test_fused_epilogue_equals_separate_post_processing(emul, oracle, pkg):
    """Leg forces / torques from the solve's own scatter phase (host build) = the separate restated post-processing."""
    h, dt, B = 10, 0.03, 6
    b = pkg.synth.make_mpc_batch("a1", h, dt, B, seed=77, gait="trot")
    wb = pkg.synth.make_wbc_batch("a1", B, seed=78)
    rb = b["robot"]
    from quadruped_robot_b200 import capi
    P = capi.params_of(rb, h, dt)
    q = np.ascontiguousarray(wb["state"][:, 13:25])
    r = emul.solve_ex(P, b, rb, q, wb["cmd"].copy())
    base = emul.solve(P, b)
    assert np.array_equal(r["grf"], base["grf"])
    for i in range(B):
        ff_o, tau_o = oracle.grf_to_torque(rb, b["quat"][i], q[i], r["grf"][i])
        assert np.array_equal(r["f_ff"][i], ff_o)
        np.testing.assert_allclose(r["tau"][i], tau_o, rtol=2e-6, atol=2e-6)
        assert np.array_equal(r["cmd"][i, 51:63], r["grf"][i])
        assert np.array_equal(np.delete(r["cmd"][i], np.s_[51:63]), np.delete(wb["cmd"][i], np.s_[51:63]))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_ctl_kernels_match_reference(gpu, oracle, pkg):
    import torch
    if not oracle.ref_ctl_available():
        pytest.skip("oracle/_ref/libqr_ctl_ref.so not built")
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    B = 4096
    for name, horizontal in (("a1", True), ("lite3", False)):
        rb = pkg.robots.ROBOTS[name]
        G = gpu.leg_geometry_of(rb)
        wb = pkg.synth.make_wbc_batch(name, B, seed=80)
        q, qd, quat = wb["state"][:, 13:25].copy(), wb["state"][:, 25:37].copy(), wb["state"][:, :4].copy()
        fb, jac, fv = torch.empty((B, 12), device="cuda"), torch.empty((B, 36), device="cuda"), torch.empty((B, 12), device="cuda")
        gpu.leg_kinematics_batch_device(G, dev(q), dev(qd), fb, jac, fv, st)
        iq, iqd = torch.empty((B, 12), device="cuda"), torch.empty((B, 12), device="cuda")
        gpu.leg_ik_batch_device(G, fb, fv, None, iq, iqd, st)
        r = torch.empty((B, 12), device="cuda")
        gpu.mpc_lever_arms_batch_device(rb, dev(quat), fb, r, st)
        torch.cuda.synchronize()
        fb_h, jac_h, fv_h, iq_h, iqd_h, r_h = (x.cpu().numpy() for x in (fb, jac, fv, iq, iqd, r))
        com = np.array(rb.com_offset, F32)
        for i in range(0, B, 16):
            k = oracle.ref_leg_kinematics(rb, q[i], qd[i])
            np.testing.assert_allclose(fb_h[i], k["foot_base"], rtol=0, atol=3e-7)
            np.testing.assert_allclose(jac_h[i].reshape(4, 3, 3), k["jac"], rtol=0, atol=3e-7)
            np.testing.assert_allclose(fv_h[i], k["foot_vel"], rtol=0, atol=2e-6)
            np.testing.assert_allclose(iq_h[i], k["ik_q"], rtol=0, atol=3e-6)
            np.testing.assert_allclose(iqd_h[i], k["ik_qd"], rtol=2e-5, atol=2e-5)
            Rb = np.array(_base_rmat(quat[i]), np.float64)
            want = (Rb @ (fb_h[i].reshape(4, 3) - com).T.astype(np.float64)).T.reshape(12)
            np.testing.assert_allclose(r_h[i], want, rtol=0, atol=2e-7)
        # swing targets written into the WBC command rows
        s = swing_inputs(pkg, name, B, 81)
        cmd = dev(s["cmd"])
        fbd, qdes, qddes = (torch.full((B, 12), -7.0, device="cuda") for _ in range(3))
        valid = torch.empty((B, 4), dtype=torch.int32, device="cuda")
        gpu.swing_targets_batch_device(G, dev(s["pos"]), dev(s["quat"]), dev(s["v_world"]), dev(s["foothold"]), dev(s["phase"]),
                                       dev(s["switch"]), dev(s["dur"]), dev(s["mask"]), horizontal, cmd, st, fbd, qdes, qddes, valid)
        torch.cuda.synchronize()
        cmd_h, fbd_h, q_h, qd_h, valid_h = (x.cpu().numpy() for x in (cmd, fbd, qdes, qddes, valid))
        for i in range(0, B, 16):
            ok_mask = np.array([int(s["mask"][i, l] and -1e-3 <= s["phase"][i, l] < 1 + 1e-3) for l in range(4)], np.int32)
            assert np.array_equal(valid_h[i], ok_mask)
            o = oracle.ref_swing_targets(rb, s["pos"][i], s["quat"][i], s["v_world"][i], s["foothold"][i], s["phase"][i], s["switch"][i],
                                         s["dur"][i], ok_mask, horizontal)
            for l in range(4):
                sl = slice(3 * l, 3 * l + 3)
                if ok_mask[l]:
                    np.testing.assert_allclose(cmd_h[i, 15 + 3 * l:18 + 3 * l], o["p_foot_des"][sl], rtol=0, atol=2e-6)
                    assert np.array_equal(cmd_h[i, 27 + 3 * l:30 + 3 * l], o["v_foot_des"][sl])
                    np.testing.assert_allclose(fbd_h[i, sl], o["foot_base_des"][sl], rtol=0, atol=1e-6)
                    np.testing.assert_allclose(q_h[i, sl], o["q_des"][sl], rtol=0, atol=2e-5)
                else:
                    assert np.array_equal(cmd_h[i, 15 + 3 * l:18 + 3 * l], s["cmd"][i, 15 + 3 * l:18 + 3 * l])


@pytest.mark.gpu
def test_gpu_gait_update_matches_reference(gpu, oracle, pkg):
    import torch
    if not oracle.ref_ctl_available():
        pytest.skip("oracle/_ref/libqr_ctl_ref.so not built")
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    B, ticks = 512, 300
    cfg, istate, fstate, out, t0 = gait_cfg(pkg, B, 82)
    rng = np.random.default_rng(83)
    d_cfg, d_i, d_f = dev(cfg), dev(istate), dev(fstate)
    d_pf, d_np, d_sr = dev(out[:, 0:4]), dev(out[:, 4:8]), dev(out[:, 8:12])
    d_allow, d_early, d_mask = (torch.empty((B, 4), dtype=torch.int32, device="cuda") for _ in range(3))
    contacts_all = (rng.uniform(size=(ticks, B, 4)) < 0.6).astype(np.int32)
    check = list(range(0, B, 32))
    r_state = {i: (istate[i].copy(), fstate[i].copy(), out[i].copy(), None) for i in check}
    for tick in range(ticks):
        t = (t0 + F32(0.002) * F32(tick)).astype(F32)
        gpu.gait_update_batch_device(dev(t), d_cfg, 0.1, dev(contacts_all[tick]), None, True, d_i, d_f, d_pf, d_np, d_sr, st,
                                     allow=d_allow, early=d_early, swing_mask=d_mask)
        for i in check:
            ri, rf, ro, _ = r_state[i]
            ri, rf, ro, ra = oracle.ref_gait_update(t[i], cfg[i], contacts_all[tick, i], ri, rf, ro, advanced_trot=True)
            r_state[i] = (ri, rf, ro, ra)
    torch.cuda.synchronize()
    gi, gf = d_i.cpu().numpy(), d_f.cpu().numpy()
    go = np.concatenate([d_pf.cpu().numpy(), d_np.cpu().numpy(), d_sr.cpu().numpy()], axis=1)
    ga, ge, gm = d_allow.cpu().numpy(), d_early.cpu().numpy(), d_mask.cpu().numpy()
    for i in check:
        ri, rf, ro, ra = r_state[i]
        assert np.array_equal(gi[i], ri) and np.array_equal(gf[i], rf) and np.array_equal(go[i], ro), i
        ls = ri[12:16]
        assert np.array_equal(ga[i], ra) and np.array_equal(ge[i], (ls == 2).astype(np.int32))
        assert np.array_equal(gm[i], (~(((ls == 1) & (ra == 1)) | (ls == 2))).astype(np.int32))


@pytest.mark.gpu
def test_gpu_fused_epilogue(gpu, oracle, pkg):
    """qr_gpu_mpc_solve_batch_ex: same forces as the plain call; f_ff, tau and the Fr_des rows equal the separate kernels."""
    import torch
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for B in (40, 3000):     # latency kernel and throughput kernels
        h, dt = 10, 0.03
        b = pkg.synth.make_mpc_batch("lite3", h, dt, B, seed=84, gait="trot")
        wb = pkg.synth.make_wbc_batch("lite3", B, seed=85)
        rb = b["robot"]
        P = gpu.params_of(rb, h, dt)
        d = {k: dev(b[k]) for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait", "mu")}
        q = dev(wb["state"][:, 13:25])
        cmd = dev(wb["cmd"])
        out = dict(grf=torch.empty((B, 12), device="cuda"), status=torch.empty(B, dtype=torch.int32, device="cuda"))
        ff, tau = torch.empty((B, 12), device="cuda"), torch.empty((B, 12), device="cuda")
        gpu.mpc_solve_batch_device_ex(P, d, out, st, rb, q=q, f_ff=ff, tau=tau, wbc_cmd=cmd)
        out2 = dict(grf=torch.empty((B, 12), device="cuda"))
        gpu.mpc_solve_batch_device(P, d, out2, st)
        ff2, tau2 = torch.empty((B, 12), device="cuda"), torch.empty((B, 12), device="cuda")
        gpu.mpc_leg_torque_batch_device(rb, d["quat"], q, out2["grf"], ff2, tau2, st)
        torch.cuda.synchronize()
        assert (out["status"] == 0).all()
        assert torch.equal(out["grf"], out2["grf"]) and torch.equal(ff, ff2) and torch.equal(tau, tau2)
        cmd_h = cmd.cpu().numpy()
        assert np.array_equal(cmd_h[:, 51:63], out["grf"].cpu().numpy())
        assert np.array_equal(np.delete(cmd_h, np.s_[51:63], axis=1), np.delete(wb["cmd"], np.s_[51:63], axis=1))
