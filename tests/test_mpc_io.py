"""MPC pre/post-processing rows: contact table + reference trajectory on the device (bit-exact with the
oracle) and GRF -> joint torques (analytic leg Jacobian)."""
import ctypes as C

import numpy as np
import pytest

F32 = np.float32


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def random_io(pkg, B, h, seed):
    rng = np.random.default_rng(seed)
    progress = rng.uniform(0, 1, (B, 4)).astype(F32)
    duty = rng.choice([0.4, 0.6, 0.75, 1.0], (B, 1)).astype(F32).repeat(4, 1)
    progress[0] = [0.0, 1.0, duty[0, 0], F32(duty[0, 0]) - F32(1e-7)]
    early = (rng.uniform(size=(B, 4)) < 0.1).astype(np.int32)
    contacts = rng.integers(0, 2, (B, 4)).astype(np.int32)
    init = rng.uniform(-1, 1, (B, 12)).astype(F32)
    pos = (init[:, 3:5] + rng.uniform(-0.3, 0.3, (B, 2))).astype(F32)
    return progress, duty, early, contacts, init, np.ascontiguousarray(pos)


def test_emul_inputs_bit_exact(emul, oracle, pkg):
    for h, nhl in ((5, 2), (10, 2), (16, 25)):
        progress, duty, early, contacts, init, pos = random_io(pkg, 64, h, 40 + h)
        for i in range(64):
            tab = np.zeros(4 * h, F32)
            emul.lib.qr_emul_mpc_contact_table(h, nhl, fp(progress[i]), fp(duty[i]), ip(early[i]), ip(contacts[i]), fp(tab))
            assert np.array_equal(tab.reshape(h, 4), oracle.contact_table(h, nhl, progress[i], duty[i], early[i], contacts[i]))
            tr = np.zeros(12 * h, F32)
            emul.lib.qr_emul_mpc_reference_traj(h, C.c_float(0.03), fp(init[i]), fp(pos[i]), fp(tr))
            assert np.array_equal(tr, oracle.reference_traj(h, 0.03, init[i], pos[i]))
        # and the host-side numpy mirror used by the workload generator
        mine = pkg.synth.contact_table(h, nhl, progress, duty, early.astype(bool), contacts.astype(bool))
        for i in range(64):
            assert np.array_equal(mine[i], oracle.contact_table(h, nhl, progress[i], duty[i], early[i], contacts[i]))


def test_emul_grf_to_torque(emul, oracle, pkg):
    rng = np.random.default_rng(50)
    for name in ("a1", "lite3"):
        rb = pkg.robots.ROBOTS[name]
        b = pkg.synth.make_wbc_batch(name, 64, seed=51)
        exact = 0
        for i in range(64):
            quat, q = b["state"][i, :4].copy(), b["state"][i, 13:25].copy()
            f = rng.uniform(-60, 130, 12).astype(F32)
            ff_o, tau_o = oracle.grf_to_torque(rb, quat, q, f)
            ff, tau = np.zeros(12, F32), np.zeros(12, F32)
            emul.lib.qr_emul_mpc_grf_to_torque(C.c_float(rb.hip_len), C.c_float(rb.upper_len), C.c_float(rb.lower_len),
                                               fp(quat), fp(q), fp(f), fp(ff), fp(tau))
            assert np.array_equal(ff, ff_o)                       # pure float32 arithmetic: bit-exact
            np.testing.assert_allclose(tau, tau_o, rtol=2e-6, atol=2e-6)   # sin/cos/sqrt: last-ulp differences at most
            exact += np.array_equal(tau, tau_o)
            # physics: tau = J^T f_ff with J = d(foot position in base frame)/dq  (finite differences of the FK)
        assert exact >= 56


def test_torque_matches_fk_jacobian(oracle, pkg):
    """Independent check of the analytic Jacobian by virtual work: tau . dq = f_ff . d(foot position), with the
    foot position from the reference's OTHER closed form, FootPositionInHipFrame (src/robots/qr_robot.cpp:125-145;
    the analytic leg model has no lateral foot offset, unlike the rigid-body tree of the WBC)."""
    rb = pkg.robots.ROBOTS["a1"]
    rng = np.random.default_rng(52)

    def foot_in_hip(q3, leg):
        sh = rb.hip_len * (-1.0) ** (leg + 1)
        l = np.sqrt(rb.upper_len ** 2 + rb.lower_len ** 2 + 2 * rb.upper_len * rb.lower_len * np.cos(q3[2]))
        eff = q3[1] + q3[2] / 2
        ox, ozh = -l * np.sin(eff), -l * np.cos(eff)
        return np.array([ox, np.cos(q3[0]) * sh - np.sin(q3[0]) * ozh, np.sin(q3[0]) * sh + np.cos(q3[0]) * ozh])

    for _ in range(5):
        q = (np.tile([0.0, 0.9, -1.8], 4) + rng.uniform(-0.2, 0.2, 12)).astype(F32)
        quat = np.array([1, 0, 0, 0], F32)
        f = rng.uniform(-50, 50, 12).astype(F32)
        ff, tau = oracle.grf_to_torque(rb, quat, q, f)
        eps = 1e-4
        for j in range(12):
            leg, a = j // 3, j % 3
            qp, qm = q[3 * leg:3 * leg + 3].astype(float), q[3 * leg:3 * leg + 3].astype(float)
            qp[a] += eps
            qm[a] -= eps
            dp = (foot_in_hip(qp, leg) - foot_in_hip(qm, leg)) / (2 * eps)
            assert abs(float(ff[3 * leg:3 * leg + 3].astype(float) @ dp) - float(tau[j])) < 1e-3 * max(1.0, abs(float(tau[j])))


@pytest.mark.gpu
def test_gpu_inputs_and_torque(gpu, oracle, pkg):
    import torch
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    h, nhl, B = 10, 2, 2048
    progress, duty, early, contacts, init, pos = random_io(pkg, B, h, 60)
    gait = torch.empty((B, 4 * h), device="cuda")
    traj = torch.empty((B, 12 * h), device="cuda")
    gpu.mpc_inputs_batch_device(h, nhl, 0.03, dev(progress), dev(duty), dev(early), dev(contacts), dev(init), dev(pos), gait, traj, st)
    torch.cuda.synchronize()
    gait, traj = gait.cpu().numpy(), traj.cpu().numpy()
    for i in range(0, B, 5):
        assert np.array_equal(gait[i].reshape(h, 4), oracle.contact_table(h, nhl, progress[i], duty[i], early[i], contacts[i]))
        assert np.array_equal(traj[i], oracle.reference_traj(h, 0.03, init[i], pos[i]))
    # torques
    rb = pkg.robots.ROBOTS["a1"]
    b = pkg.synth.make_wbc_batch("a1", B, seed=61)
    f = np.random.default_rng(62).uniform(-60, 130, (B, 12)).astype(F32)
    ff = torch.empty((B, 12), device="cuda")
    tau = torch.empty((B, 12), device="cuda")
    gpu.mpc_leg_torque_batch_device(rb, dev(b["state"][:, :4]), dev(b["state"][:, 13:25]), dev(f), ff, tau, st)
    torch.cuda.synchronize()
    ff, tau = ff.cpu().numpy(), tau.cpu().numpy()
    for i in range(0, B, 9):
        ff_o, tau_o = oracle.grf_to_torque(rb, b["state"][i, :4], b["state"][i, 13:25], f[i])
        assert np.array_equal(ff[i], ff_o)
        np.testing.assert_allclose(tau[i], tau_o, rtol=2e-6, atol=2e-6)
