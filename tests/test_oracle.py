"""CPU tests of the oracle itself: the vendored solvers behave, the restatement reproduces the golden
fixtures, and the host-side table / trajectory builders are bit-exact with it."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import parity


def test_qpoases_known_answer(oracle):
    """Data of the first QP of qpOASES' own testing/cpp/test_example1.cpp (variable bounds written as
    constraint rows, solved through the reference's call sequence with setToMPC options): the optimum
    is (0.5, -1.5) and the KKT residuals of SolutionAnalysis::getKktViolation are at rounding level."""
    H = np.array([[1.0, 0.0], [0.0, 0.5]])
    A = np.array([[1.0, 1.0]])
    g = np.array([1.5, 1.0])
    # example1 also has variable bounds lb=(0.5,-2), ub=(5,2); express them as constraint rows
    Aall = np.vstack([A, np.eye(2)])
    lbA = np.array([-1.0, 0.5, -2.0])
    ubA = np.array([2.0, 5.0, 2.0])
    x, info, kkt, _ = oracle.qpoases_dense(H, g, Aall, lbA, ubA, 10)
    assert info[0] == 0
    np.testing.assert_allclose(x, [0.5, -1.5], atol=1e-12)  # the optimum printed by example1
    assert kkt[0] < 1e-10 and kkt[1] < 1e-10 and kkt[2] < 1e-10


def test_quadprog_demo(oracle):
    """The hand-checked QP of extern/QuadProgpp/src/main.cc is covered in test_wbc_oracle once the WBC
    oracle lands; here we only require the archive to be present next to qpOASES."""
    ref = os.path.join(os.path.dirname(oracle.__file__), "_ref")
    assert os.path.exists(os.path.join(ref, "libqpOASES.a"))
    assert os.path.exists(os.path.join(ref, "libquadprog.a"))


@pytest.mark.parametrize("path", parity.golden_files(), ids=os.path.basename)
def test_oracle_reproduces_golden(path, oracle, pkg):
    z, b, h, dt, mu_sweep = parity.load_golden(path, pkg)
    B = b["p"].shape[0]
    for i in range(B):
        P = oracle.params_of(b["robot"], h, dt, mu=float(b["mu"][i]))
        H, g, ub = oracle.mpc_build(P, b, i)
        assert hashlib.sha256(H.tobytes()).hexdigest() == str(z["H_sha256"][i])
        assert np.array_equal(g, z["g"][i]) and np.array_equal(ub, z["ub"][i])
        if i == 0:
            assert np.array_equal(H, z["H0"])
        x, info, kkt, cstat = oracle.mpc_qpoases(h, P.mu, H, g, ub, 100000)
        np.testing.assert_allclose(x, z["x_conv"][i], rtol=0, atol=1e-9)
        assert np.array_equal(cstat, z["cstat"][i])
        # the exact optimum certifies itself
        A = oracle.constraint_rows(h, P.mu)
        stat, feas = oracle.kkt_certificate(H, g, A, np.zeros(20 * h), ub.astype(float), z["x_star"][i])
        assert stat < 1e-9 and feas < 1e-9
        # converged qpOASES is near it, but not within the element-wise tolerance in general
        assert np.abs(x - z["x_star"][i]).max() < 0.2


def test_contact_table_bit_exact(oracle, pkg):
    rng = np.random.default_rng(7)
    synth = pkg.synth
    for gname, gt in pkg.robots.GAITS.items():
        for h in (5, 10, 16):
            nh = synth.num_horizon_l(gt)
            progress = rng.uniform(0, 1, (64, 4)).astype(np.float32)
            # include exact boundary phases
            progress[0] = [0.0, 1.0, gt["duty"], np.float32(gt["duty"]) - np.float32(1e-7)]
            duty = np.full((64, 4), gt["duty"], np.float32)
            early = rng.integers(0, 2, (64, 4)) * (rng.uniform(size=(64, 4)) < 0.1)
            contacts = rng.integers(0, 2, (64, 4))
            mine = synth.contact_table(h, nh, progress, duty, early.astype(bool), contacts.astype(bool))
            for i in range(64):
                ref = oracle.contact_table(h, nh, progress[i], duty[i], early[i], contacts[i])
                assert np.array_equal(mine[i], ref), (gname, h, i)


def test_reference_traj_bit_exact(oracle, pkg):
    rng = np.random.default_rng(8)
    for h in (5, 10, 16):
        init = rng.uniform(-1, 1, (32, 12)).astype(np.float32)
        pos = (init[:, 3:5] + rng.uniform(-0.3, 0.3, (32, 2))).astype(np.float32)
        mine = pkg.synth.reference_traj(h, 0.03, init, pos)
        for i in range(32):
            ref = oracle.reference_traj(h, 0.03, init[i], pos[i])
            assert np.array_equal(mine[i], ref)


def test_stock_cap_is_reported(oracle, pkg):
    """P3 of the parity protocol: with the stock nWSR = 100 some instances return
    RET_MAX_NWSR_REACHED and a truncated iterate (the reference ignores the code)."""
    z, b, h, dt, _ = parity.load_golden(
        os.path.join(parity.HERE, "golden", "mpc_a1_h10_musweep.npz"), pkg)
    capped = z["info_stock"][:, 0] == oracle.RET_MAX_NWSR_REACHED
    assert capped.any()
    i = int(np.nonzero(capped)[0][0])
    assert np.abs(z["x_stock"][i] - z["x_star"][i]).max() > 1e-2
