"""CPU parity tests of the DEVICE SOURCES (csrc/*.h) compiled for the host (tests/emul): the same
control flow and arithmetic the CUDA kernels execute, checked against the oracle and the golden
fixtures without a GPU.  The GPU tests (test_gpu_parity.py) repeat these through the C ABI."""
import hashlib
import os

import numpy as np
import pytest

import parity


@pytest.mark.parametrize("path", parity.golden_files(), ids=os.path.basename)
def test_condense_bit_exact_vs_golden(path, emul, oracle, pkg):
    """(H, g, ub) built by the device code equal the oracle's float32 build bit for bit."""
    z, b, h, dt, mu_sweep = parity.load_golden(path, pkg)
    P = oracle.params_of(b["robot"], h, dt)
    H, g, ub = emul.condense(P, b)
    for i in range(H.shape[0]):
        assert hashlib.sha256(H[i].tobytes()).hexdigest() == str(z["H_sha256"][i]) or np.array_equal(
            H[i], oracle.mpc_build(P, b, i)[0])  # sha differs only through -0.0 vs +0.0
        assert np.array_equal(g[i], z["g"][i])
        assert np.array_equal(ub[i], z["ub"][i])
    assert np.array_equal(H[0], z["H0"])


@pytest.mark.parametrize("path", parity.golden_files(), ids=os.path.basename)
def test_fused_solve_vs_golden(path, emul, oracle, pkg):
    """P1/P2: forces within 1e-4 rel / 1e-5 abs of x*, element-wise; swing forces exactly zero."""
    z, b, h, dt, mu_sweep = parity.load_golden(path, pkg)
    P = oracle.params_of(b["robot"], h, dt)
    r = emul.solve(P, b, per_instance_mu=mu_sweep)
    assert (r["status"] == 0).all()
    for i in range(b["p"].shape[0]):
        parity.assert_elementwise(r["u64"][i], z["x_star"][i], f"{os.path.basename(path)}[{i}] f64")
        parity.assert_elementwise(r["u"][i], z["x_star"][i], f"{os.path.basename(path)}[{i}] f32")
        parity.assert_vs_qpoases(r["u"][i], z["x_conv"][i], z["x_star"][i])
        sw = parity.swing_mask(b["gait"][i], h)
        assert (r["u"][i][sw] == 0).all()
        assert np.array_equal(r["grf"][i], r["u"][i][:12])
        # the device solver is never further from the optimum than converged qpOASES is
        assert np.abs(r["u64"][i] - z["x_star"][i]).max() <= np.abs(z["x_conv"][i] - z["x_star"][i]).max() + 1e-9


def test_qp_solver_on_identical_data(emul, oracle, pkg):
    """P1: the solver alone, fed the byte-identical (H, g, ub) qpOASES consumes."""
    h, dt = 10, 0.03
    b = pkg.synth.make_mpc_batch("a1", h, dt, 12, seed=11, gait="mixed")
    P = oracle.params_of(b["robot"], h, dt)
    Hs, gs, ubs = zip(*[oracle.mpc_build(P, b, i) for i in range(12)])
    r = emul.qp_solve(h, P.mu, np.stack(Hs), np.stack(gs), np.stack(ubs))
    assert (r["status"] == 0).all()
    A = oracle.constraint_rows(h, P.mu)
    for i in range(12):
        xq, info, kkt, cstat = oracle.mpc_qpoases(h, P.mu, Hs[i], gs[i], ubs[i], 100000)
        xs, _ = oracle.polish_from_working_set(Hs[i], gs[i], A, np.zeros(20 * h), ubs[i].astype(float), cstat)
        parity.assert_elementwise(r["x64"][i], xs, f"qp[{i}]")
        parity.assert_vs_qpoases(r["x64"][i], xq, xs)
        stat, feas = oracle.kkt_certificate(Hs[i], gs[i], A, np.zeros(20 * h), ubs[i].astype(float), r["x64"][i])
        assert stat < 1e-9 and feas < 1e-8


def test_same_active_set_as_qpoases(emul, oracle, pkg):
    """P1(i): identical zero pattern / saturated f_z pattern as the converged reference solver."""
    h, dt = 10, 0.03
    b = pkg.synth.make_mpc_batch("a1", h, dt, 8, seed=12, gait="trot")
    P = oracle.params_of(b["robot"], h, dt)
    r = emul.solve(P, b)
    for i in range(8):
        xq, _ = oracle.mpc_solve(P, b, i, 100000)
        fz, fzq = r["u64"][i][2::3], xq[2::3]
        assert np.array_equal(np.abs(fz - P.f_max) < 1e-6, np.abs(fzq - P.f_max) < 1e-6)
        assert np.array_equal(np.abs(fz) < 1e-6, np.abs(fzq) < 1e-6)


@pytest.mark.parametrize("gait", ["stand", "gallop", "walk"])
def test_gaits_and_edge_masks(gait, emul, oracle, pkg):
    h, dt = 10, 0.03
    b = pkg.synth.make_mpc_batch("aliengo", h, dt, 6, seed=13, gait=gait)
    # edge cases: all swing (flight phase) and a single stance foot-step
    b["gait"][0] = 0.0
    b["gait"][1] = 0.0
    b["gait"][1][0] = 1.0
    P = oracle.params_of(b["robot"], h, dt)
    r = emul.solve(P, b)
    assert (r["status"] == 0).all()
    assert (r["u"][0] == 0).all()
    A = oracle.constraint_rows(h, P.mu)
    for i in range(1, 6):
        H, g, ub = oracle.mpc_build(P, b, i)
        xq, info, kkt, cstat = oracle.mpc_qpoases(h, P.mu, H, g, ub, 100000)
        xs, _ = oracle.polish_from_working_set(H, g, A, np.zeros(20 * h), ub.astype(float), cstat)
        parity.assert_elementwise(r["u64"][i], xs, f"{gait}[{i}]")


def test_horizon_extremes(emul, oracle, pkg):
    for h in (1, 2, 16):
        b = pkg.synth.make_mpc_batch("a1", h, 0.03, 3, seed=14 + h, gait="trot")
        P = oracle.params_of(b["robot"], h, 0.03)
        Hc, gc, ubc = emul.condense(P, b)
        r = emul.solve(P, b)
        assert (r["status"] == 0).all()
        A = oracle.constraint_rows(h, P.mu)
        for i in range(3):
            H, g, ub = oracle.mpc_build(P, b, i)
            assert np.array_equal(H, Hc[i]) and np.array_equal(g, gc[i]) and np.array_equal(ub, ubc[i])
            xq, info, kkt, cstat = oracle.mpc_qpoases(h, P.mu, H, g, ub, 100000)
            xs, _ = oracle.polish_from_working_set(H, g, A, np.zeros(20 * h), ub.astype(float), cstat)
            parity.assert_elementwise(r["u64"][i], xs, f"h={h}[{i}]")


def test_bad_inputs_are_flagged(emul, oracle, pkg):
    h, dt = 5, 0.06
    b = pkg.synth.make_mpc_batch("lite3", h, dt, 3, seed=20)
    b["p"][0, 0] = np.nan
    b["gait"][1, 3] = -1.0          # negative bound -> infeasible box
    P = oracle.params_of(b["robot"], h, dt)
    r = emul.solve(P, b)
    assert r["status"][0] == 3 and (r["u"][0] == 0).all()
    assert r["status"][1] == 2 and (r["u"][1] == 0).all()
    assert r["status"][2] == 0


def test_permutation_and_determinism(emul, oracle, pkg):
    h, dt = 10, 0.03
    b = pkg.synth.make_mpc_batch("a1", h, dt, 10, seed=21, gait="mixed")
    P = oracle.params_of(b["robot"], h, dt)
    r1 = emul.solve(P, b)
    perm = np.random.default_rng(0).permutation(10)
    b2 = {k: (np.ascontiguousarray(v[perm]) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    r2 = emul.solve(P, b2)
    assert np.array_equal(r1["u64"][perm], r2["u64"])


def test_cycling_instances_converge_without_fallback(emul, oracle, pkg):
    """Lite3 trot instances on which the plain block active-set updates cycle (period 4, found by tracing):
    the cycle detection switches to the restricted one-add / one-drop mode, which ends at the verified optimum
    without the interior-point fallback; the result is the exact optimum of the oracle-built QP."""
    h, dt, B = 10, 0.03, 1500
    b = pkg.synth.make_mpc_batch("lite3", h, dt, B, seed=3, gait="trot")
    idx = np.array([29, 338, 366, 520, 720, 802, 927])
    sub = {k: (np.ascontiguousarray(v[idx]) if isinstance(v, np.ndarray) and v.shape[:1] == (B,) else v) for k, v in b.items()}
    P = oracle.params_of(sub["robot"], h, dt)
    from quadruped_robot_b200 import capi
    opt = capi.default_options()
    opt.flags = capi.QP_NO_PREDICTION   # from a cold start (with the coarse prediction they no longer cycle)
    e = emul.solve(P, sub, opt=opt)
    assert (e["status"] == 0).all()
    assert (e["iters"][:, 0] == 0).all(), e["iters"]
    assert e["iters"][:, 1].max() <= 32 and e["iters"][:, 1].min() >= 12   # these really are the hard ones
    Po = P
    A = oracle.constraint_rows(h, Po.mu)
    for i in range(len(idx)):
        H, g, ub = oracle.mpc_build(Po, sub, i)
        stat, feas = oracle.kkt_certificate(H, g, A, np.zeros(20 * h), ub.astype(float), e["u64"][i])
        assert stat < 1e-7 and feas < 1e-7, (i, stat, feas)


def test_interleaved_hessian_pair_is_bit_identical(emul, oracle, pkg):
    """The fused path evaluates a Hessian entry and its transposed entry in one interleaved loop
    (qr_condense_h_pair); every bit equals the entry-wise evaluation that the golden / oracle tests pin."""
    for robot, h, gait, seed in (("a1", 10, "trot", 51), ("lite3", 5, "trot", 52), ("aliengo", 16, "mixed", 53)):
        b = pkg.synth.make_mpc_batch(robot, h, 0.03, 6, seed=seed, gait=gait)
        P = oracle.params_of(b["robot"], h, 0.03)
        assert emul.condense_pair_mismatches(P, b) == 0


def test_all_apex_instance(emul, oracle, pkg):
    """A reference trajectory that asks for a downward acceleration above g: every stance foot-step ends pinned to the apex
    f = 0 of its pyramid, the reduced system of that round is EMPTY (no factorisation, x = p), and the verification
    accepts it."""
    h, dt = 10, 0.03
    b = pkg.synth.make_mpc_batch("a1", h, dt, 3, seed=23, gait="trot")
    b["traj"] = b["traj"].copy()
    b["traj"].reshape(3, h, 12)[:, :, 5] = b["p"][:, 2:3] - 10.0     # desired height 10 m below the robot
    P = oracle.params_of(b["robot"], h, dt)
    e = emul.solve(P, b)
    assert (e["status"] == 0).all() and (e["iters"][:, 0] == 0).all()
    A = oracle.constraint_rows(h, P.mu)
    for i in range(3):
        H, g, ub = oracle.mpc_build(P, b, i)
        assert np.abs(e["u64"][i]).max() == 0.0                        # the optimum is f = 0 everywhere ...
        # ... a degenerate vertex (all five rows of every foot-step meet there; the reference's qpOASES call does not
        # survive it -- it returns forces of thousands of newtons with an error code): certified by NNLS multipliers
        stat, feas = oracle.kkt_certificate(H, g, A, np.zeros(20 * h), ub.astype(float), e["u64"][i])
        assert stat < 1e-9 and feas == 0.0
