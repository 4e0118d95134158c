"""Pins of the oracle's restatements around the MPC / WBC core against code COMPILED FROM THE REFERENCE
(oracle/_ref/libqr_ctl_ref.so, built by oracle/Makefile target `refctl`: whole reference translation units plus line
ranges of member functions cut out of the reference sources, see oracle/ref_ctl_shim.cpp).  Every comparison is bit
for bit: both sides are float32 code compiled by the same g++ without FMA contraction.

  row a8   contact table + reference trajectory  <- qr_mpc_stance_leg_controller.cpp:282-303, 344-376
  row a8   lever arms, f_ff, Fr_des              <- qr_mpc_stance_leg_controller.cpp:385-410 (SolveDenseMPC, real solve)
  row a9   GRF -> joint torques                  <- qr_robot.cpp:148-172, 241-251
  row a17  swing parabola                        <- qr_foot_trajectory_generator.cpp (whole file) + qr_geometry.cpp
  row f3   force-balance QP (world frame)        <- qr_qp_torque_optimizer.cpp (whole file) + QuadProg++
  row f4   foothold heuristic                    <- qr_foothold_planner.cpp:112-239
"""
import numpy as np
import pytest

F32 = np.float32


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_ctl_available():
        pytest.skip("oracle/_ref/libqr_ctl_ref.so not built (needs /root/reference)")
    return oracle


def test_contact_table_and_trajectory_restatement_is_the_reference(ref, pkg):
    rng = np.random.default_rng(0)
    for h, nhl in ((5, 2), (10, 2), (16, 25)):
        for _ in range(150):
            progress = rng.uniform(0, 1, 4).astype(F32)
            duty = np.full(4, rng.choice([0.4, 0.6, 0.75, 1.0]), F32)
            early = (rng.uniform(size=4) < 0.15).astype(np.int32)
            contacts = rng.integers(0, 2, 4).astype(np.int32)
            init = rng.uniform(-1, 1, 12).astype(F32)
            init[[6, 7, 11]] = 0   # {rollComp, pitchComp, yaw, x, y, z, 0, 0, yawRate, vx, vy, 0} (qr_mpc_stance_leg_controller.cpp:363-366)
            pos = (init[3:5] + rng.uniform(-0.3, 0.3, 2)).astype(F32)
            leg_state = np.where(early > 0, 2, rng.integers(0, 2, 4)).astype(np.int32)   # LegState::EARLY_CONTACT = 2
            tab, traj = ref.ref_mpc_inputs(h, nhl, 0.03, progress, duty, leg_state, contacts, init, pos)
            assert np.array_equal(tab, ref.contact_table(h, nhl, progress, duty, early, contacts))
            assert np.array_equal(traj, ref.reference_traj(h, 0.03, init, pos))
            # and the numpy mirror the workload generator uses
            mine = pkg.synth.contact_table(h, nhl, progress[None], duty[None], early[None].astype(bool), contacts[None].astype(bool))
            assert np.array_equal(mine[0], tab)


def test_swing_parabola_restatement_is_the_reference(ref):
    rng = np.random.default_rng(1)
    rejected = 0
    for _ in range(800):
        s = rng.uniform(-0.3, 0.3, 3).astype(F32)
        e = (s + rng.uniform(-0.2, 0.2, 3)).astype(F32)
        hgt, ph = F32(rng.uniform(0.03, 0.12)), F32(rng.uniform(-0.1, 1.1))
        for pm in (False, True):
            p, v, a, ok = ref.ref_swing_parabola(s, e, hgt, ph, pm)
            po, oko = ref.swing_parabola(s, e, hgt, ph, pm)
            assert ok == oko
            rejected += not ok
            if ok:
                assert np.array_equal(p, po)
            assert not v.any() and not a.any()   # the reference's parabola generator leaves velocity / acceleration at 0
    assert rejected > 0


def test_leg_torque_restatement_is_the_reference(ref, pkg):
    rng = np.random.default_rng(2)
    for name in ("a1", "lite3"):
        rb = pkg.robots.ROBOTS[name]
        b = pkg.synth.make_wbc_batch(name, 96, seed=51)
        for i in range(96):
            quat, q = b["state"][i, :4].copy(), b["state"][i, 13:25].copy()
            f = rng.uniform(-60, 130, 12).astype(F32)
            ff_o, tau_o = ref.grf_to_torque(rb, quat, q, f)
            k = ref.ref_leg_kinematics(rb, q, f_leg=ff_o)
            assert np.array_equal(k["tau"], tau_o)


def test_solve_dense_mpc_lever_arms_and_leg_forces(ref, pkg, monkeypatch):
    """The reference's SolveDenseMPC run for real (its own SolveMPCKernel + qpOASES): the lever arms it hands to the
    solver are what the workload generator / the device lever-arm kernel produce, f equals the reference MPC build's
    GetMPCSolution, and f_ff = -R_base^T f equals the restated post-processing."""
    monkeypatch.setenv("MINI_EIGEN_EXP_NILPOTENT3", "1")
    h, dt, B = 10, 0.03, 6
    rb = pkg.robots.ROBOTS["a1"]
    b = pkg.synth.make_mpc_batch("a1", h, dt, B, seed=9, gait="trot")
    P = ref.params_of(rb, h, dt)
    rng = np.random.default_rng(3)
    for i in range(B):
        quat = b["quat"][i]
        Rb = _rot_of_quat(quat.astype(np.float64))
        # feet in the base frame such that R (foot - comOffset) reproduces the batch's lever arms to float32 rounding
        foot_base = (Rb.T @ b["r_feet"][i].reshape(4, 3).T.astype(np.float64)).T + np.array(rb.com_offset)
        foot_base = foot_base.astype(F32).reshape(12)
        o = ref.ref_solve_dense_mpc(P, rb, b["rpy"][i], b["p"][i], quat, b["v"][i], b["w"][i], foot_base, b["traj"][i], b["gait"][i])
        assert np.abs(o["lever"] - b["r_feet"][i]).max() < 2e-6
        # the same solve through the MPC-only reference build on the lever arms the controller computed
        b2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        b2["r_feet"][i] = o["lever"]
        _, _, _, x = ref.ref_mpc_solve(P, b2, i)
        assert np.array_equal(o["f"], x[:12].astype(F32))
        assert np.array_equal(o["fr_des"], o["f"])
        ff_o, _ = ref.grf_to_torque(rb, quat, np.zeros(12, F32), o["f"])
        assert np.array_equal(o["f_ff"], ff_o)


def _rot_of_quat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_foothold_restatement_is_the_reference(ref, pkg):
    B = 300
    f = pkg.synth.make_foothold_batch("a1", B, seed=55)
    # vx_des >= 0: for vx_des < -0.01 the reference writes footTargetPosition(0,2) / (0,3) of a 3-vector, i.e. out of
    # range (qr_foothold_planner.cpp:219-222) -- undefined behaviour that neither the restatement nor the kernel mirrors
    f["des_speed"][:, 0] = np.abs(f["des_speed"][:, 0])
    Po = ref.foothold_params_of(f["params"])
    checked = 0
    for i in range(B):
        fh, ph = ref.ref_foothold(f["robot"], f["params"], f, i, np.full(12, -9, F32), np.full(4, -9, F32))
        for leg in range(4):
            if f["swing_mask"][i, leg]:
                fo, pho = ref.foothold(Po, leg, f, i)
                assert np.array_equal(fh[3 * leg:3 * leg + 3], fo) and ph[leg] == F32(pho)
                checked += 1
            else:
                assert (fh[3 * leg:3 * leg + 3] == -9).all() and ph[leg] == -9
    assert checked > B


def test_force_balance_restatement_is_the_reference(ref, pkg):
    """World-frame overload of ComputeContactForce (qr_qp_torque_optimizer.cpp:304-400) compiled from the reference with
    its own QuadProg++, against the restatement: forces bit for bit, including the legs in swing (status 1)."""
    fb = pkg.synth.make_fb_batch("a1", 200, seed=3, world_frame=True)
    P = ref.fb_params_of(fb["params"])
    seen_swing = 0
    for i in range(200):
        o = ref.force_balance(P, fb["foot"][i], fb["acc"][i], fb["contact"][i])
        r = ref.ref_contact_force_world(fb["params"], [1, 0, 0, 0], fb["foot"][i], fb["acc"][i], fb["contact"][i])
        assert np.array_equal(r, o["force"]), i
        seen_swing += int(fb["contact"][i].sum() < 4)
    assert seen_swing > 50


def test_force_balance_control_frame_overload(ref, pkg):
    """Control-frame overload of ComputeContactForce (qr_qp_torque_optimizer.cpp:190-301) compiled from the reference, fed
    through a ground-estimator stand-in: on flat ground (Rcb = I) the restatement's forces equal it bit for bit; on a slope
    the QP data (Rcb I Rcb', rotated feet, rotated gravity, tilted normal -- derived by the same expressions, :203-225)
    give the same X, compared after the reference's final X * Rcb in float32."""
    fb = pkg.synth.make_fb_batch("a1", 96, seed=4, world_frame=False)
    P = ref.fb_params_of(fb["params"])
    rng = np.random.default_rng(5)
    for i in range(96):
        F, d = ref.ref_contact_force_control(fb["params"], [1, 0, 0, 0], fb["foot"][i], fb["acc"][i], fb["contact"][i], terrain_type=0)
        o = ref.force_balance(P, d["foot"], fb["acc"][i], fb["contact"][i], inertia=d["inertia"], gravity=d["gravity"])
        assert np.array_equal(F.reshape(4, 3), o["force"].reshape(4, 3)), i
        pitch, roll = rng.uniform(-0.3, 0.3), rng.uniform(-0.1, 0.1)
        quat = pkg.synth._quat_from_rpy(np.array([[roll, pitch, 0.0]]))[0].astype(F32)
        cp = F32(pitch + rng.uniform(-0.05, 0.05))
        Rg = pkg.synth._rot_zyx(np.array([[0.0, float(cp), 0.0]]))[0].astype(F32)
        F, d = ref.ref_contact_force_control(fb["params"], quat, fb["foot"][i], fb["acc"][i], fb["contact"][i], terrain_type=3,
                                             control_rpy=(0, cp, 0), aligned=Rg)
        n, t2 = d["normal"], np.array([0, 1, 0], F32)
        t1 = np.cross(t2, n).astype(F32)          # :95-96
        o = ref.force_balance(P, d["foot"], fb["acc"][i], fb["contact"][i], inertia=d["inertia"], gravity=d["gravity"],
                              frame=np.concatenate([n, t1, t2]).astype(F32))
        want = (o["force"].reshape(4, 3).astype(F32) @ d["Rcb"]).astype(F32)
        assert np.abs(F.reshape(4, 3) - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), i
