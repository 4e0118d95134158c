"""Swing-leg targets of SURVEY.md section 8f rank 4: WALK-mode cubic B-spline trajectory and heuristic foothold.

Oracle = oracle/swing_oracle.cpp: the generator glue restated, the spline evaluated by the reference's OWN vendored
tinynurbs (compiled from /root/reference with oracle/mini_glm).  The device code issues the same float32 operations in
the same order, so the CPU build of the device sources must agree BIT FOR BIT (same libm); on the GPU the only
difference is the last-bit rounding of CUDA's atan2f / sinf / cosf against glibc's (tolerance 2e-6 m)."""
import numpy as np
import pytest


def test_bspline_emulation_bit_exact(pkg, emul, oracle):
    b = pkg.synth.make_swing_batch(400, seed=51)
    n_valid = 0
    for i in range(400):
        args = (b["initial_pos"][i], b["target_pos"][i], float(b["height"][i]), float(b["duration"][i]),
                float(b["initial_time"][i]), float(b["time"][i]))
        po, vo, oko = oracle.swing_bspline(*args)
        pe, ve, oke = emul.swing_bspline(*args)
        assert oke == oko, i
        if oko:
            n_valid += 1
            assert np.array_equal(pe, po) and np.array_equal(ve, vo), (i, pe, po, ve, vo)
    assert 350 < n_valid <= 398         # samples more than 1e-3 outside [0, duration] are rejected like the reference does
    # end points: the curve starts at the initial position and ends at the target
    p0, _, _ = oracle.swing_bspline(b["initial_pos"][0], b["target_pos"][0], 0.08, 1.0, 0.0, 0.0)
    p1, _, _ = oracle.swing_bspline(b["initial_pos"][0], b["target_pos"][0], 0.08, 1.0, 0.0, 1.0)
    assert np.abs(p0 - b["initial_pos"][0]).max() < 1e-6 and np.abs(p1 - b["target_pos"][0]).max() < 1e-6


def test_foothold_emulation_bit_exact(pkg, emul, oracle):
    from quadruped_robot_b200 import capi
    for robot, seed in (("a1", 52), ("lite3", 53)):
        b = pkg.synth.make_foothold_batch(robot, 200, seed=seed)
        Pe, Po = capi.foothold_params_of(b["params"]), oracle.foothold_params_of(b["params"])
        for i in range(200):
            for leg in range(4):
                fo, pho = oracle.foothold(Po, leg, b, i)
                fe, phe = emul.foothold(Pe, leg, b, i)
                assert np.array_equal(fe, fo) and phe == pho, (robot, i, leg, fe, fo)


@pytest.mark.gpu
def test_gpu_swing_and_foothold(pkg, gpu, oracle):
    import torch
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # ---- B-spline
    B = 4096
    b = pkg.synth.make_swing_batch(B, seed=54)
    pos = torch.full((B, 3), 7.0, device="cuda"); vel = torch.full((B, 3), 7.0, device="cuda")
    valid = torch.empty(B, dtype=torch.int32, device="cuda")
    gpu.swing_bspline_batch_device(dev(b["initial_pos"]), dev(b["target_pos"]), dev(b["height"]), dev(b["duration"]),
                                   dev(b["initial_time"]), dev(b["time"]), pos, vel, valid, st)
    torch.cuda.synchronize()
    pos, vel, valid = pos.cpu().numpy(), vel.cpu().numpy(), valid.cpu().numpy()
    for i in range(0, B, 13):
        po, vo, ok = oracle.swing_bspline(b["initial_pos"][i], b["target_pos"][i], float(b["height"][i]), 1.0, 0.0, float(b["time"][i]))
        assert bool(valid[i]) == ok
        if ok:
            assert np.abs(pos[i] - po).max() < 2e-6 and np.abs(vel[i] - vo).max() < 2e-5 * max(1.0, np.abs(vo).max())
        else:
            assert (pos[i] == 7.0).all()          # rejected samples leave the outputs untouched
    # ---- foothold
    B = 2048
    f = pkg.synth.make_foothold_batch("a1", B, seed=55)
    P = gpu.foothold_params_of(f["params"])
    d = {k: dev(v) for k, v in f.items() if isinstance(v, np.ndarray)}
    fh = torch.full((B, 12), -9.0, device="cuda"); ph = torch.full((B, 4), -9.0, device="cuda")
    gpu.foothold_heuristic_batch_device(P, d, fh, ph, st)
    torch.cuda.synchronize()
    fh, ph = fh.cpu().numpy(), ph.cpu().numpy()
    Po = oracle.foothold_params_of(f["params"])
    for i in range(0, B, 17):
        for leg in range(4):
            if f["swing_mask"][i, leg]:
                fo, pho = oracle.foothold(Po, leg, f, i)
                assert np.abs(fh[i, 3 * leg:3 * leg + 3] - fo).max() < 2e-6 and ph[i, leg] == np.float32(pho)
            else:
                assert (fh[i, 3 * leg:3 * leg + 3] == -9.0).all() and ph[i, leg] == -9.0
