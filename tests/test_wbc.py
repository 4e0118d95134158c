"""Whole-body control: oracle self-checks (physics), device sources vs oracle on CPU (host emulation) and
through the C ABI on the GPU, golden fixtures, swing parabola.

Tolerances.  The reference computes the WBC in float32; re-running its algorithm in float64 moves the
torques by ~1e-5 N m (tests/golden/wbc_*.npz hold both).  The engine computes in float64 and is compared
  * with the float64 oracle at 1e-9 (abs + rel), and
  * with the float32 oracle (the reference's arithmetic) at BASELINE's 1e-4 relative / 1e-5 absolute on the
    torque norm (element-wise the reference's own rounding noise is already ~1e-5).
"""
import glob
import os

import numpy as np
import pytest

import parity

FIELDS = ("tau", "fr", "qdes", "qddes")


def load_wbc_golden(path, pkg):
    z = np.load(path)
    b = dict(state=np.ascontiguousarray(z["state"]), cmd=np.ascontiguousarray(z["cmd"]),
             contact=np.ascontiguousarray(z["contact"]), robot=pkg.robots.ROBOTS[str(z["meta"][0])])
    return z, b


def unpack_dbg(d):
    return dict(H=d[:324].reshape(18, 18), G=d[324:342], C=d[342:360], Jc=d[360:576].reshape(4, 3, 18),
                Jcdqd=d[576:588].reshape(4, 3), pGC=d[588:600].reshape(4, 3), vGC=d[600:612].reshape(4, 3), qdd=d[612:630])


def wbc_goldens():
    return sorted(glob.glob(os.path.join(parity.HERE, "golden", "wbc_*.npz")))


def check_against_golden(r, z, B):
    for i in range(B):
        for f in FIELDS:
            np.testing.assert_allclose(r[f][i], z[f + "_f64"][i], rtol=1e-9, atol=1e-9, err_msg=f"{f}[{i}] vs float64 oracle")
        d = unpack_dbg(r["dbg"][i])
        for f in ("H", "G", "C", "Jc", "Jcdqd", "pGC", "vGC"):
            np.testing.assert_allclose(d[f], z[f + "_f64"][i], rtol=1e-10, atol=1e-10, err_msg=f"{f}[{i}]")
        # against the reference's float32 arithmetic: norm-wise BASELINE tolerance
        t32 = z["tau_f32"][i].astype(float)
        assert np.abs(r["tau"][i] - t32).max() <= 1e-4 * np.abs(t32).max() + 1e-5


# ---------------------------------------------------------------------------------------------
# oracle self-checks: the dynamics restatement against physics, not against itself
# ---------------------------------------------------------------------------------------------
def test_oracle_dynamics_physics(oracle, pkg):
    b = pkg.synth.make_wbc_batch("a1", 6, seed=300)
    M = oracle.wbc_model_of(b["robot"])
    for i in range(6):
        st = b["state"][i].astype(np.float64)
        r = oracle.wbc_step(M, st, b["cmd"][i], b["contact"][i], "f64")
        H = r["H"]
        assert np.abs(H - H.T).max() < 1e-12 and np.linalg.eigvalsh(H).min() > 1e-4
        np.testing.assert_allclose([H[3, 3], H[4, 4], H[5, 5]], 13.5, rtol=1e-7)   # 6 + 4 (0.696 + 1.013 + 0.166)
        qdot = np.concatenate([st[7:13], st[25:37]])
        np.testing.assert_allclose(r["vGC"], r["Jc"] @ qdot, atol=1e-6)           # foot velocity = Jc qdot
        feet = pkg.synth.foot_positions_world(b["robot"], b["rpy"][i:i + 1], st[None, 4:7], st[None, 13:25])[0]
        np.testing.assert_allclose(r["pGC"], feet, atol=1e-6)                      # independent analytic FK
        # Jacobian columns of the joints = finite differences of the foot position
        eps = 1e-4
        for j in (0, 4, 8, 11):
            sp, sm = st.copy(), st.copy()
            sp[13 + j] += eps
            sm[13 + j] -= eps
            pp = oracle.wbc_step(M, sp.astype(np.float32), b["cmd"][i], b["contact"][i], "f64")["pGC"]
            pm = oracle.wbc_step(M, sm.astype(np.float32), b["cmd"][i], b["contact"][i], "f64")["pGC"]
            dq = float(np.float32(sp[13 + j])) - float(np.float32(sm[13 + j]))
            np.testing.assert_allclose((pp - pm) / dq, r["Jc"][:, :, 6 + j], atol=2e-4)
        # gravity: with zero velocity, C = 0 and G[5] (vertical base force row in body frame) carries the weight
        s0 = st.copy()
        s0[7:13] = 0
        s0[25:37] = 0
        r0 = oracle.wbc_step(M, s0.astype(np.float32), b["cmd"][i], b["contact"][i], "f64")
        assert np.abs(r0["C"]).max() < 1e-9
        assert abs(np.linalg.norm(r0["G"][3:6]) - 13.5 * 9.81) < 1e-3


def test_oracle_quadprog_demo(oracle, pkg):
    """All four feet in stance with the MPC force already feasible and the dynamics consistent is not
    required: the oracle's QuadProg++ call must return a finite cost (rc 0) on every fixture instance."""
    for path in wbc_goldens():
        z, b = load_wbc_golden(path, pkg)
        M = oracle.wbc_model_of(b["robot"])
        for i in range(b["state"].shape[0]):
            assert oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f64")["rc"] == 0


@pytest.mark.parametrize("path", wbc_goldens(), ids=os.path.basename)
def test_oracle_reproduces_wbc_golden(path, oracle, pkg):
    z, b = load_wbc_golden(path, pkg)
    M = oracle.wbc_model_of(b["robot"])
    for i in range(b["state"].shape[0]):
        r = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f64")
        np.testing.assert_allclose(r["tau"], z["tau_f64"][i], rtol=0, atol=1e-10)
        r32 = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f32")
        np.testing.assert_allclose(r32["tau"], z["tau_f32"][i], rtol=0, atol=1e-4)


# ---------------------------------------------------------------------------------------------
# PIN: the reference's own WBC classes compiled unmodified from /root/reference (oracle/_ref)
# ---------------------------------------------------------------------------------------------
def _relmax(a, b):
    return np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / max(np.abs(np.asarray(b, float)).max(), 1e-9)


def _check_against_reference_build(oracle, M, state, cmd, contact, what):
    """Float32 reference build vs the float64 restatement: agreement at float32 rounding level.  Measured
    worst cases: dynamics 5e-7 of the largest entry, tau / fr 2e-6, qdes 3e-6, qddes 2e-5 (the kinematic
    WBC chains four float32 pseudo-inverses)."""
    ref = oracle.wbc_step(M, state, cmd, contact, "ref")
    f64 = oracle.wbc_step(M, state, cmd, contact, "f64")
    f32 = oracle.wbc_step(M, state, cmd, contact, "f32")
    for f in ("H", "G", "C", "Jc", "Jcdqd", "pGC", "vGC"):
        assert _relmax(ref[f], f64[f]) < 3e-6, (what, f, _relmax(ref[f], f64[f]))
    for f, tol in (("tau", 2e-5), ("fr", 2e-5), ("qdes", 2e-5), ("qddes", 1e-4)):   # BASELINE form: rel + 1e-5 abs
        r_ = ref[f].astype(float)
        bound = tol * np.abs(r_).max() + 1e-5
        assert np.abs(f64[f] - r_).max() <= bound, (what, f, np.abs(f64[f] - r_).max(), bound)
        # the float32 restatement is as close to the reference build as float32 allows
        assert np.abs(f32[f] - r_).max() <= bound, (what, f, "f32")
    swing = np.repeat(np.asarray(contact) == 0, 3)
    assert (ref["fr"][swing] == 0).all()
    return ref


@pytest.mark.parametrize("path", wbc_goldens(), ids=os.path.basename)
def test_wbc_restatement_matches_reference_source_on_golden(path, oracle, pkg):
    """FloatingBaseModel<float>, qrSingleContact, the three tasks, qrMultitaskProjection and
    qrWholeBodyImpulseCtrl (with QuadProg++) of the reference, compiled from their own source files, on the
    golden inputs; the committed golden torques are within the BASELINE tolerance of what the reference
    build returns."""
    if not oracle.ref_wbc_available():
        pytest.skip("oracle/_ref/libqr_wbc_ref.so not built (needs /root/reference at build time)")
    z, b = load_wbc_golden(path, pkg)
    M = oracle.wbc_model_of(b["robot"])
    for i in range(b["state"].shape[0]):
        ref = _check_against_reference_build(oracle, M, b["state"][i], b["cmd"][i], b["contact"][i], (path, i))
        t = ref["tau"].astype(float)
        assert np.abs(z["tau_f64"][i] - t).max() <= 1e-4 * np.abs(t).max() + 1e-5


def test_wbc_restatement_matches_reference_source_all_contact_patterns(oracle, pkg):
    if not oracle.ref_wbc_available():
        pytest.skip("oracle/_ref/libqr_wbc_ref.so not built (needs /root/reference at build time)")
    for robot, seed in (("a1", 311), ("lite3", 312)):
        b = pkg.synth.make_wbc_batch(robot, 32, seed=seed)
        pats = [[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [1, 1, 0, 0], [1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 0], [1, 1, 1, 1]]
        for i, p in enumerate(pats * 4):
            b["contact"][i] = p
            b["cmd"][i, 51:63] *= np.repeat(np.array(p, np.float32), 3)
        # a quarter of the instances ask for forces far outside the friction pyramid (active QP constraints)
        b["cmd"][24:, 51:63] *= np.float32(3.0)
        b["cmd"][24:, 51:63:3] += np.float32(40.0) * (b["cmd"][24:, 53:63:3] != 0)
        M = oracle.wbc_model_of(b["robot"])
        for i in range(32):
            _check_against_reference_build(oracle, M, b["state"][i], b["cmd"][i], b["contact"][i], (robot, i))


# ---------------------------------------------------------------------------------------------
# device sources on the CPU (host emulation)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", wbc_goldens(), ids=os.path.basename)
def test_emul_wbc_vs_golden(path, emul, pkg):
    z, b = load_wbc_golden(path, pkg)
    r = emul.wbc_solve(b)
    assert (r["status"] == 0).all()
    check_against_golden(r, z, b["state"].shape[0])


def test_emul_wbc_vs_oracle_all_contact_patterns(emul, oracle, pkg):
    b = pkg.synth.make_wbc_batch("lite3", 16, seed=301)
    pats = [[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [1, 1, 0, 0], [1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 0], [1, 1, 1, 1]]
    for i, p in enumerate(pats * 2):
        b["contact"][i] = p
        b["cmd"][i, 51:63] *= np.repeat(np.array(p, np.float32), 3)
    r = emul.wbc_solve(b)
    assert (r["status"] == 0).all()
    M = oracle.wbc_model_of(b["robot"])
    for i in range(16):
        o = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f64")
        for f in FIELDS:
            np.testing.assert_allclose(r[f][i], o[f], rtol=1e-9, atol=1e-9, err_msg=f"{f}[{i}] pattern {b['contact'][i]}")
        swing = np.repeat(np.array(b["contact"][i]) == 0, 3)
        assert (r["fr"][i][swing] == 0).all()


def test_emul_wbc_force_limits_active(emul, oracle, pkg):
    """Desired forces far outside the friction pyramid / above the normal-force limit: the QP must clip them
    exactly as QuadProg++ does."""
    b = pkg.synth.make_wbc_batch("a1", 8, seed=302)
    b["contact"][:] = 1
    b["cmd"][:, 51:63] = np.tile(np.array([60.0, -50.0, 30.0], np.float32), 4)
    b["cmd"][4:, 51:63] = np.tile(np.array([5.0, 5.0, 200.0], np.float32), 4)
    r = emul.wbc_solve(b)
    assert (r["status"] == 0).all()
    M = oracle.wbc_model_of(b["robot"])
    for i in range(8):
        o = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f64")
        np.testing.assert_allclose(r["fr"][i], o["fr"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(r["tau"][i], o["tau"], rtol=1e-8, atol=1e-8)
        f = r["fr"][i].reshape(4, 3)
        assert (np.abs(f[:, 0]) <= 0.4 * f[:, 2] + 1e-6).all() and (f[:, 2] <= 13.5 * 9.81 + 1e-3).all()


def test_swing_parabola_bit_exact(emul, oracle):
    rng = np.random.default_rng(5)
    for _ in range(400):
        s = rng.uniform(-0.3, 0.3, 3).astype(np.float32)
        e = (s + rng.uniform(-0.2, 0.2, 3)).astype(np.float32)
        h = np.float32(rng.uniform(0.03, 0.15))
        t = np.float32(rng.uniform(-0.01, 1.01))
        for pm in (False, True):
            po, oko = oracle.swing_parabola(s, e, h, t, pm)
            pe, oke = emul.swing_parabola(s, e, h, t, pm)
            assert oko == oke
            if oko:
                assert np.array_equal(po, pe), (s, e, h, t, pm, po, pe)
    # end points: phase 0 gives the start, phase 1 the target
    p0, _ = oracle.swing_parabola([0.1, 0.2, -0.3], [0.2, 0.1, -0.28], 0.08, 0.0)
    p1, _ = oracle.swing_parabola([0.1, 0.2, -0.3], [0.2, 0.1, -0.28], 0.08, 1.0)
    np.testing.assert_allclose(p0, [0.1, 0.2, -0.3], atol=1e-7)
    np.testing.assert_allclose(p1, [0.2, 0.1, -0.28], atol=1e-6)
    pm_, _ = oracle.swing_parabola([0.1, 0.2, -0.3], [0.2, 0.1, -0.28], 0.08, 0.5)
    np.testing.assert_allclose(pm_[2], -0.28 + 0.08, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# GPU, through the C ABI
# ---------------------------------------------------------------------------------------------
def gpu_wbc(gpu, b):
    import torch
    B = b["state"].shape[0]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = {k: torch.empty((B, 12), dtype=torch.float64, device="cuda") for k in FIELDS}
    dbg = torch.empty((B, 630), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    gpu.wbc_solve_batch_device_f64(gpu.wbc_model_of(b["robot"]), dev(b["state"]), dev(b["cmd"]), dev(b["contact"]),
                                   out["tau"], torch.cuda.current_stream().cuda_stream, fr=out["fr"], qdes=out["qdes"],
                                   qddes=out["qddes"], dbg=dbg, status=st)
    torch.cuda.synchronize()
    r = {k: v.cpu().numpy() for k, v in out.items()}
    r["dbg"] = dbg.cpu().numpy()
    r["status"] = st.cpu().numpy()
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("path", wbc_goldens(), ids=os.path.basename)
def test_gpu_wbc_vs_golden(path, gpu, pkg):
    z, b = load_wbc_golden(path, pkg)
    r = gpu_wbc(gpu, b)
    assert (r["status"] == 0).all()
    check_against_golden(r, z, b["state"].shape[0])


@pytest.mark.gpu
def test_gpu_wbc_batch_1024_vs_emulation_and_oracle(gpu, emul, oracle, pkg):
    """BASELINE config 2 shape: Lite3, 1024 randomised states."""
    import torch
    b = pkg.synth.make_wbc_batch("lite3", 1024, seed=303)
    r = gpu_wbc(gpu, b)
    assert (r["status"] == 0).all()
    e = emul.wbc_solve({k: (v[:128] if isinstance(v, np.ndarray) else v) for k, v in b.items()})
    for f in FIELDS:
        np.testing.assert_allclose(r[f][:128], e[f], rtol=1e-9, atol=1e-9)
    M = oracle.wbc_model_of(b["robot"])
    for i in range(0, 1024, 37):
        o = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f64")
        for f in FIELDS:
            np.testing.assert_allclose(r[f][i], o[f], rtol=1e-9, atol=1e-9)
        o32 = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "f32")
        assert np.abs(r["tau"][i] - o32["tau"]).max() <= 1e-4 * np.abs(o32["tau"]).max() + 1e-5
    # float32 entry point = rounded float64 result; determinism
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tau32 = torch.empty((1024, 12), device="cuda")
    gpu.wbc_solve_batch_device(gpu.wbc_model_of(b["robot"]), dev(b["state"]), dev(b["cmd"]), dev(b["contact"]), tau32,
                               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(tau32.cpu().numpy(), r["tau"].astype(np.float32))
    r2 = gpu_wbc(gpu, b)
    assert np.array_equal(r2["tau"], r["tau"])


@pytest.mark.gpu
def test_gpu_wbc_vs_reference_source_build(gpu, oracle, pkg):
    """The CUDA path against the reference's OWN WBC classes compiled from /root/reference
    (oracle/_ref/libqr_wbc_ref.so travels to the GPU box): BASELINE tolerance 1e-4 rel / 1e-5 abs, norm-wise."""
    if not oracle.ref_wbc_available():
        pytest.skip("oracle/_ref/libqr_wbc_ref.so not built")
    for robot, seed in (("a1", 321), ("lite3", 322)):
        b = pkg.synth.make_wbc_batch(robot, 256, seed=seed)
        r = gpu_wbc(gpu, b)
        assert (r["status"] == 0).all()
        M = oracle.wbc_model_of(b["robot"])
        for i in range(0, 256, 5):
            ref = oracle.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], "ref")
            for f in FIELDS:
                t = ref[f].astype(float)
                assert np.abs(r[f][i] - t).max() <= 1e-4 * np.abs(t).max() + 1e-5, (robot, i, f)


@pytest.mark.gpu
def test_gpu_swing_parabola(gpu, oracle):
    import torch
    rng = np.random.default_rng(6)
    B = 4096
    s = rng.uniform(-0.3, 0.3, (B, 3)).astype(np.float32)
    e = (s + rng.uniform(-0.2, 0.2, (B, 3))).astype(np.float32)
    h = rng.uniform(0.03, 0.15, B).astype(np.float32)
    t = rng.uniform(-0.01, 1.01, B).astype(np.float32)
    for pm in (False, True):
        pos = torch.empty((B, 3), device="cuda")
        valid = torch.empty(B, dtype=torch.int32, device="cuda")
        gpu.swing_parabola_batch_device(torch.from_numpy(s).cuda(), torch.from_numpy(e).cuda(), torch.from_numpy(h).cuda(),
                                        torch.from_numpy(t).cuda(), pm, pos, valid, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        pos, valid = pos.cpu().numpy(), valid.cpu().numpy()
        mism = 0
        for i in range(0, B, 7):
            po, ok = oracle.swing_parabola(s[i], e[i], h[i], t[i], pm)
            assert ok == bool(valid[i])
            if ok:
                if pm:   # sin() of the device math library may differ from glibc in the last ulp of a double
                    np.testing.assert_allclose(pos[i], po, rtol=3e-7, atol=1e-7)
                    mism += not np.array_equal(pos[i], po)
                else:
                    assert np.array_equal(pos[i], po)
        assert mism < 20
