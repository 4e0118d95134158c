"""BASELINE configs[1] end to end on the GPU (`pytest -m gpu`): one full control tick for a batch of Lite3 robots --
gait phase -> contact table + reference trajectory -> lever arms -> MPC (with the fused leg-force / torque / Fr_des
epilogue) -> foothold -> swing targets -> WBIC -- every stage through the C ABI, nothing copied by the host in
between, against THE REFERENCE'S OWN CODE chained on the CPU: the functions of oracle/_ref/libqr_ctl_ref.so (cut out
of / compiled from qr_mpc_stance_leg_controller.cpp, qr_robot.cpp, qr_swing_leg_controller.cpp, qr_foothold_planner.cpp
...), its SolveMPCKernel / GetMPCSolution with qpOASES, and its WBC classes with QuadProg++ (libqr_wbc_ref.so).

Tolerances: masks / tables / trajectories bit-exact; lever arms 1 ulp; forces 1e-4 rel / 1e-5 abs against the exact
optimum (and within the reference solver's own distance of the reference's forces where its stock nWSR = 100 run
finishes); WBC torques 1e-4 * max|tau| + 1e-5 against the reference WBC fed with the same Fr_des, and the element-wise
error distribution is recorded in gpurun_out/chain_parity.json."""
import json
import os

import numpy as np
import pytest

import bulk

pytestmark = pytest.mark.gpu
F32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_full_tick_chain_against_reference_builds(gpu, oracle, pkg, monkeypatch):
    import torch
    if not (oracle.ref_ctl_available() and oracle.ref_wbc_available() and oracle.ref_mpc_available()):
        pytest.skip("reference builds not available")
    monkeypatch.setenv("MINI_EIGEN_EXP_NILPOTENT3", "1")
    st = torch.cuda.current_stream().cuda_stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    B, h, dt = 256, 10, 0.03
    rb = pkg.robots.ROBOTS["lite3"]
    rng = np.random.default_rng(950)
    U = rng.uniform
    mb = pkg.synth.make_mpc_batch("lite3", h, dt, B, seed=951, gait="trot")
    wb = pkg.synth.make_wbc_batch("lite3", B, seed=952)
    # one consistent robot state for both halves of the tick
    quat, pos = mb["quat"], mb["p"]
    state = wb["state"].copy()
    state[:, :4], state[:, 4:7] = quat, pos
    q = np.ascontiguousarray(state[:, 13:25])
    G = gpu.leg_geometry_of(rb)
    P = gpu.params_of(rb, h, dt)
    gt = pkg.robots.GAITS["trot"]
    nhl = pkg.synth.num_horizon_l(gt)
    # gait state: every robot somewhere in its cycle
    duty, stance = F32(gt["duty"]), F32(gt["stance_duration"])
    period = F32(stance / duty)
    cfg = np.zeros((B, 4, 5), F32)
    cfg[:, :, 0] = np.array(gt["offsets"], F32)
    cfg[:, :, 1], cfg[:, :, 2], cfg[:, :, 3], cfg[:, :, 4] = period, duty, period - stance, duty
    cfg = cfg.reshape(B, 20)
    istate = np.ones((B, 20), np.int32); istate[:, 16:] = 0
    fstate = np.zeros((B, 4), F32); fstate[:, 3] = 1.0
    phases = np.zeros((B, 12), F32)
    t_now = U(0.0, 2.0, B).astype(F32)
    contacts = np.ones((B, 4), np.int32)
    init = np.zeros((B, 12), F32)
    init[:, 2], init[:, 3:5], init[:, 5] = mb["rpy"][:, 2], pos[:, :2], rb.body_height
    init[:, 8], init[:, 9:11] = U(-0.3, 0.3, B), U(-0.3, 0.6, (B, 2))
    fh_in = pkg.synth.make_foothold_batch("lite3", B, seed=953)
    fh_in["des_speed"][:, 0] = np.abs(fh_in["des_speed"][:, 0])
    switch_pos = (np.array(rb.hip_positions)[None] + np.array([0, 0, -rb.body_height]) + U(-0.06, 0.06, (B, 4, 3))).astype(F32).reshape(B, 12)
    v_world = mb["v"]
    swing_dur = np.tile(cfg.reshape(B, 4, 5)[:, :, 3], 1).astype(F32)

    # ---------------- GPU tick
    d_i, d_f = dev(istate), dev(fstate)
    d_pf, d_np, d_sr = dev(phases[:, :4]), dev(phases[:, 4:8]), dev(phases[:, 8:])
    d_allow, d_early, d_mask, contact_state = (torch.empty((B, 4), dtype=torch.int32, device="cuda") for _ in range(4))
    gpu.gait_update_batch_device(dev(t_now), dev(cfg), 0.1, dev(contacts), None, False, d_i, d_f, d_pf, d_np, d_sr, st,
                                 allow=d_allow, early=d_early, swing_mask=d_mask, stance_mask=contact_state)
    d_duty = dev(np.full((B, 4), duty, F32))
    gait = torch.empty((B, 4 * h), device="cuda"); traj = torch.empty((B, 12 * h), device="cuda")
    gpu.mpc_inputs_batch_device(h, nhl, dt, d_pf, d_duty, d_early, contact_state, dev(init), dev(pos[:, :2]), gait, traj, st)
    foot_base = torch.empty((B, 12), device="cuda")
    gpu.leg_kinematics_batch_device(G, dev(q), None, foot_base, None, None, st)
    r_feet = torch.empty((B, 12), device="cuda")
    gpu.mpc_lever_arms_batch_device(rb, dev(quat), foot_base, r_feet, st)
    d = dict(p=dev(pos), v=dev(mb["v"]), quat=dev(quat), w=dev(mb["w"]), r_feet=r_feet, rpy=dev(mb["rpy"]), traj=traj, gait=gait)
    cmd = dev(wb["cmd"])
    out = dict(grf=torch.empty((B, 12), device="cuda"), u=torch.empty((B, 12 * h), device="cuda"),
               status=torch.empty(B, dtype=torch.int32, device="cuda"))
    ff, tau_mpc = torch.empty((B, 12), device="cuda"), torch.empty((B, 12), device="cuda")
    gpu.mpc_solve_batch_device_ex(P, d, out, st, rb, q=dev(q), f_ff=ff, tau=tau_mpc, wbc_cmd=cmd)
    # swing side: foothold heuristic for the swing legs, then the WBC foot targets
    fP = gpu.foothold_params_of(fh_in["params"])
    fd = {k: dev(v) for k, v in fh_in.items() if isinstance(v, np.ndarray)}
    fd["swing_remain"], fd["norm_phase"], fd["allow_switch"], fd["swing_mask"] = d_sr, d_np, d_allow, d_mask
    fd["foot_base"] = foot_base
    foothold = torch.zeros((B, 12), device="cuda"); planner_phase = torch.zeros((B, 4), device="cuda")
    gpu.foothold_heuristic_batch_device(fP, fd, foothold, planner_phase, st)
    gpu.swing_targets_batch_device(G, dev(pos), dev(quat), dev(v_world), foothold, planner_phase, dev(switch_pos), dev(swing_dur),
                                   d_mask, True, cmd, st)
    tau = torch.empty((B, 12), dtype=torch.float64, device="cuda"); wst = torch.empty(B, dtype=torch.int32, device="cuda")
    gpu.wbc_solve_batch_device_f64(gpu.wbc_model_of(rb), dev(state), cmd, contact_state, tau, st, status=wst)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in dict(gait=gait, traj=traj, r_feet=r_feet, grf=out["grf"], u=out["u"], status=out["status"], ff=ff,
                                              tau_mpc=tau_mpc, cmd=cmd, tau=tau, wst=wst, mask=d_mask, early=d_early, pf=d_pf,
                                              foot_base=foot_base, foothold=foothold, pphase=planner_phase, contact=contact_state,
                                              allow=d_allow, np_=d_np, sr=d_sr).items()}
    assert (g["status"] == 0).all() and (g["wst"] == 0).all()

    # ---------------- the reference chained on the CPU
    Po = oracle.params_of(rb, h, dt)
    Mo = oracle.wbc_model_of(rb)
    mb_gpu = dict(mb); mb_gpu["traj"], mb_gpu["gait"], mb_gpu["r_feet"] = g["traj"], g["gait"], g["r_feet"]
    exact = bulk.run(mb_gpu, h, dt, np.arange(B))
    worst_f = bulk.err_over_tol(g["u"], exact["x_star"]).max()
    assert worst_f <= 1.0, worst_f
    tau_err, tau_err_chain, n_fin = [], [], 0
    for i in range(B):
        ri, rf, ro, ra = oracle.ref_gait_update(t_now[i], cfg[i], contacts[i], istate[i], fstate[i], phases[i])
        ls = ri[12:16]
        assert np.array_equal(g["pf"][i], ro[:4]) and np.array_equal(g["early"][i], (ls == 2).astype(np.int32))
        swing = (~(((ls == 1) & (ra == 1)) | (ls == 2))).astype(np.int32)
        assert np.array_equal(g["mask"][i], swing)
        tab, trj = oracle.ref_mpc_inputs(h, nhl, dt, ro[:4], np.full(4, duty, F32), ls, 1 - swing, init[i], pos[i, :2])
        assert np.array_equal(g["gait"][i].reshape(h, 4), tab) and np.array_equal(g["traj"][i], trj)
        k = oracle.ref_leg_kinematics(rb, q[i])
        np.testing.assert_allclose(g["foot_base"][i], k["foot_base"], rtol=0, atol=3e-7)
        o = oracle.ref_solve_dense_mpc(Po, rb, mb["rpy"][i], pos[i], quat[i], mb["v"][i], mb["w"][i], g["foot_base"][i], trj, tab)
        np.testing.assert_allclose(g["r_feet"][i], o["lever"], rtol=0, atol=1.2e-7)
        # forces: the reference's stock run against the exact optimum bounds how far it may be from the GPU's answer
        fin = exact["stock"][i] == 0
        if fin:
            n_fin += 1
            gap = np.abs(o["f"] - exact["x_star"][i, :12]).max()
            assert np.abs(g["grf"][i] - o["f"]).max() <= gap + 1e-4 * np.abs(o["f"]).max() + 1e-5, i
        ff_ref, tau_ref = oracle.grf_to_torque(rb, quat[i], q[i], g["grf"][i])     # = the reference's own lines (test_ref_pins)
        assert np.array_equal(g["ff"][i], ff_ref)
        np.testing.assert_allclose(g["tau_mpc"][i], tau_ref, rtol=2e-6, atol=2e-6)
        assert np.array_equal(g["cmd"][i, 51:63], g["grf"][i])
        # swing side
        fh, ph = oracle.ref_foothold(rb, fh_in["params"], dict(fh_in, swing_remain=g["sr"], norm_phase=g["np_"], allow_switch=g["allow"],
                                                             swing_mask=g["mask"], foot_base=g["foot_base"]), i,
                                     np.zeros(12, F32), np.zeros(4, F32))
        np.testing.assert_allclose(g["foothold"][i], fh, rtol=0, atol=2e-6)
        ok_mask = np.array([int(swing[l] and -1e-3 <= ph[l] < 1 + 1e-3) for l in range(4)], np.int32)
        sw = oracle.ref_swing_targets(rb, pos[i], quat[i], v_world[i], g["foothold"][i], g["pphase"][i], switch_pos[i], swing_dur[i], ok_mask, True)
        cmd_ref = wb["cmd"][i].copy()
        for l in range(4):
            if ok_mask[l]:
                cmd_ref[15 + 3 * l:18 + 3 * l] = sw["p_foot_des"][3 * l:3 * l + 3]
                cmd_ref[27 + 3 * l:30 + 3 * l] = sw["v_foot_des"][3 * l:3 * l + 3]
                cmd_ref[39 + 3 * l:42 + 3 * l] = sw["a_foot_des"][3 * l:3 * l + 3]
        np.testing.assert_allclose(g["cmd"][i, 15:51], cmd_ref[15:51], rtol=0, atol=3e-6)
        # WBC: the reference's own classes on the command rows the GPU chain produced ...
        wref = oracle.wbc_step(Mo, state[i], g["cmd"][i], g["contact"][i], "ref")
        t = wref["tau"].astype(float)
        assert np.abs(g["tau"][i] - t).max() <= 1e-4 * np.abs(t).max() + 1e-5, i
        tau_err.append(np.abs(g["tau"][i] - t) / (1e-4 * np.abs(t) + 1e-5))
        # ... and on the reference chain's own rows (its own forces as Fr_des) where its MPC run finished
        if fin:
            cmd_ref[51:63] = o["fr_des"]
            wchain = oracle.wbc_step(Mo, state[i], cmd_ref, g["contact"][i], "ref")
            tau_err_chain.append(np.abs(g["tau"][i] - wchain["tau"].astype(float)).max())
    tau_err = np.concatenate(tau_err)
    rec = dict(n=B, mpc_worst_err_over_tol_vs_exact_optimum=float(worst_f), reference_mpc_finished=n_fin,
               wbc_tau_elementwise_err_over_tol={"p50": float(np.percentile(tau_err, 50)), "p99": float(np.percentile(tau_err, 99)),
                                                 "max": float(tau_err.max()), "frac_above_1": float((tau_err > 1).mean())},
               wbc_tau_vs_full_reference_chain_max_Nm={"p50": float(np.percentile(tau_err_chain, 50)), "max": float(np.max(tau_err_chain))})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "chain_parity.json"), "w"), indent=1)
    assert n_fin >= B // 2
