"""Force-balance stance QP: the device code (host emulation on CPU, CUDA on a B200)
against the oracle = restated ComputeContactForce + the reference's own QuadProg++ (oracle/fb_oracle.cpp).

Tolerance: BASELINE.json's 1e-4 relative / 1e-5 absolute on the forces; the float32 QP data handed to the solver
must be bit-identical."""
import numpy as np
import pytest

RTOL, ATOL = 1e-4, 1e-5
CASES = [("a1", False, False, 21), ("a1", True, False, 22), ("lite3", False, True, 23), ("aliengo", True, True, 24)]


def _oracle_all(oracle, b):
    P = oracle.fb_params_of(b["params"])
    out = []
    for i in range(b["foot"].shape[0]):
        row = lambda k: None if b.get(k) is None else b[k][i]
        out.append(oracle.force_balance(P, b["foot"][i], b["acc"][i], b["contact"][i], row("inertia"), row("gravity"), row("frame")))
    return out


def _check(force, ref, tag):
    for i, r in enumerate(ref):
        err = np.abs(force[i].astype(np.float64) - r["force"].astype(np.float64))
        tol = RTOL * np.abs(r["force"]) + ATOL
        assert (err <= tol).all(), (tag, i, float((err / tol).max()), force[i], r["force"])


@pytest.mark.parametrize("robot,world,tilted,seed", CASES)
def test_emulation_matches_oracle(robot, world, tilted, seed, pkg, emul, oracle):
    from quadruped_robot_b200 import capi
    b = pkg.synth.make_fb_batch(robot, 96, seed=seed, world_frame=world, tilted=tilted)
    P = capi.fb_params_of(b["params"])
    ref = _oracle_all(oracle, b)
    # (1) the float32 QP data are bit-identical with the oracle's restatement of the reference build
    for i in range(0, 96, 7):
        G, a, Cm, lb = emul.fb_build(P, b, i)
        assert np.array_equal(G, ref[i]["G"]) and np.array_equal(a, ref[i]["a"])
        assert np.array_equal(Cm, ref[i]["C"]) and np.array_equal(lb, ref[i]["lb"])
    # (2) forces
    r = emul.force_balance(P, b)
    _check(r["force"], ref, f"{robot}/emul")
    # (3) the solver ends the way QuadProg++ does: swing legs make the reference's QP infeasible by 2e-7
    #     (status 1 <-> QuadProg++ returned inf), an all-stance robot is solved to optimality
    n_swing = (b["contact"] == 0).sum(1)
    ref_status = np.array([x["status"] for x in ref])
    assert (ref_status[n_swing == 0] == 0).all() and (r["status"][n_swing == 0] == 0).all()
    assert (r["status"] != 3).all() and (ref_status != 3).all()
    # swing legs carry (numerically) no force; X = -x is the force of the leg ON the ground, so stance legs have
    # -n.X = n.x in [fmin, fmax]
    f = r["force"].reshape(-1, 4, 3)
    n = b["frame"][:, :3] if b.get("frame") is not None else np.tile(np.array([0, 0, 1.0], np.float32), (96, 1))
    fn = np.einsum("blk,bk->bl", f, n)
    assert (np.abs(fn[b["contact"] == 0]) < 1e-5).all()
    fmin = b["params"]["fmin_ratio"][0] * b["params"]["mass"] * 9.8
    assert (-fn[b["contact"] == 1] >= fmin - 1e-3).all()


def test_golden_fixture(pkg, emul, oracle):
    """Committed vectors (tests/golden/fb_a1.npz, made by tests/golden/make_golden_fb.py from the oracle)."""
    import os
    from quadruped_robot_b200 import capi
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "fb_a1.npz"), allow_pickle=False)
    b = pkg.synth.make_fb_batch("a1", int(z["batch"]), seed=int(z["seed"]), world_frame=False, tilted=True)
    for k in ("foot", "acc", "contact", "inertia", "gravity", "frame"):
        assert np.array_equal(b[k], z[k]), k
    P = capi.fb_params_of(b["params"])
    r = emul.force_balance(P, b)
    ref = [dict(force=z["force"][i]) for i in range(int(z["batch"]))]
    _check(r["force"], ref, "golden")
    # and the oracle still reproduces its own fixture
    now = _oracle_all(oracle, b)
    assert all(np.array_equal(now[i]["force"], z["force"][i]) for i in range(int(z["batch"])))


@pytest.mark.gpu
@pytest.mark.parametrize("robot,world,tilted,seed", CASES[:3])
def test_gpu_matches_oracle(robot, world, tilted, seed, pkg, gpu, oracle, emul):
    import torch
    b = pkg.synth.make_fb_batch(robot, 512, seed=seed, world_frame=world, tilted=tilted)
    P = gpu.fb_params_of(b["params"])
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    force = torch.empty((512, 12), device="cuda")
    st = torch.empty(512, dtype=torch.int32, device="cuda")
    gpu.force_balance_batch_device(P, dev(b["foot"]), dev(b["acc"]), dev(b["contact"]), force,
                                   torch.cuda.current_stream().cuda_stream, inertia=dev(b["inertia"]),
                                   gravity=dev(b["gravity"]), frame=dev(b["frame"]), status=st)
    torch.cuda.synchronize()
    f = force.cpu().numpy()
    sub = {k: (v[:64] if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    _check(f[:64], _oracle_all(oracle, sub), f"{robot}/gpu")
    e = emul.force_balance(P, b)
    assert np.abs(f - e["force"]).max() < 2e-5
    assert (st.cpu().numpy() != 3).all()


def test_small_qp_solver_against_quadprog(emul, oracle):
    """csrc/small_qp.h alone against the reference's QuadProg++ on random strictly convex QPs: generic, heavily
    constrained, with duplicated / parallel rows (linearly dependent candidates) and with infeasible row pairs."""
    rng = np.random.default_rng(71)
    n_inf = 0
    for trial in range(300):
        n = int(rng.integers(2, 13))
        m = int(rng.integers(1, 25))
        M = rng.normal(size=(n, n))
        G = M @ M.T + 0.05 * np.eye(n)
        g0 = rng.normal(size=n) * 3
        Cm = rng.normal(size=(m, n))
        c0 = rng.normal(size=m) + (0.5 if trial % 3 else -0.5)
        kind = trial % 5
        if kind == 1 and m >= 4:      # duplicated and scaled rows
            Cm[1] = Cm[0]; c0[1] = c0[0]
            Cm[3] = 2.0 * Cm[2]; c0[3] = 2.0 * c0[2] + 0.1
        if kind == 2 and m >= 2:      # parallel rows forming a slab (feasible)
            Cm[1] = -Cm[0]; c0[1] = -c0[0] + 1.0
        if kind == 3 and m >= 2:      # contradictory pair (infeasible)
            Cm[1] = -Cm[0]; c0[0] = -1.0; c0[1] = -1.0
        xo, cost = oracle.quadprog_ineq(G, g0, Cm, c0)
        xe, st, it = emul.small_qp(G, g0, Cm, c0)
        if np.isinf(cost):
            n_inf += 1
            assert st == 1, (trial, st)
            continue
        assert st == 0, (trial, st, cost)
        assert np.abs(xe - xo).max() <= 1e-7 * (1 + np.abs(xo).max()), (trial, np.abs(xe - xo).max())
        assert (Cm @ xe + c0 >= -1e-8).all()
    assert n_inf >= 40
