"""Bulk oracle parity on the GPU box (`pytest -m gpu`): thousands of instances per BASELINE configuration solved by
the CUDA path through the C ABI and, on all host cores, by the oracle (float32 restatement -> the reference's own
qpOASES 3.2.0 run to convergence -> extended-precision optimum x* on its final working set, tests/bulk.py).

Checked on EVERY instance (north_star: "within the stated tolerance of qpOASES on every instance"):
  * |f_gpu - x*| <= 1e-4 |x*| + 1e-5 element-wise over all 12h forces (the BASELINE tolerance against the exact optimum
    of the reference's QP);
  * ||f_gpu - f_qpOASES||_inf <= max(1e-4 ||f||_inf, ||f_qpOASES - x*||_inf) + 1e-5 (the GPU is never further from
    converged qpOASES than qpOASES is from the truth);
  * swing-leg forces exactly 0 (identical contact / swing masks);
  * on the instances the stock nWSR = 100 run finishes, the reference's OWN source build (oracle/_ref/libqr_mpc_ref.so:
    SetupProblem / SolveMPCKernel / GetMPCSolution) is within its own distance to x* (+ tolerance) of the GPU forces.
The figures of every configuration are written to gpurun_out/bulk_parity.json.
"""
import json
import os

import numpy as np
import pytest

import bulk
import parity

pytestmark = pytest.mark.gpu
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait", "mu")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name, robot, horizon, dt, gait, mu sweep, instances, compare with the reference source build
CONFIGS = [
    ("a1_h10_trot", "a1", 10, 0.03, "trot", False, 2048, True),          # BASELINE configs[0] / [2]
    ("a1_h10_musweep", "a1", 10, 0.03, "trot", True, 2048, False),       # configs[2]: sweep over friction mu
    ("aliengo_h10_mixed", "aliengo", 10, 0.03, "mixed", False, 2048, True),   # configs[3]
    ("lite3_h5_trot", "lite3", 5, 0.06, "trot", False, 2048, True),      # the shipped Lite3 default
    ("lite3_h10_trot", "lite3", 10, 0.03, "trot", False, 2048, True),    # configs[1]
    ("a1_h16_trot", "a1", 16, 0.03, "trot", False, 2048, True),          # the reference's largest horizon
    # all four legs in stance at the reference's largest horizon: 64 stance foot-steps, i.e. the size class whose reduced
    # systems are factorised on the FP64 tensor cores (csrc/chol8.h) -- that path against the reference's own build
    ("a1_h16_stand", "a1", 16, 0.03, "stand", False, 512, True),
]


def _solve(gpu, P, b, opt=None, per_instance_mu=False):
    import torch
    dev = {k: torch.from_numpy(np.ascontiguousarray(b[k])).cuda() for k in KEYS}
    B, h = b["p"].shape[0], P.horizon
    out = dict(grf=torch.empty((B, 12), device="cuda"), u=torch.empty((B, 12 * h), device="cuda"),
               status=torch.empty(B, dtype=torch.int32, device="cuda"),
               iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
    gpu.mpc_solve_batch_device(P, dev, out, torch.cuda.current_stream().cuda_stream, opt=opt, per_instance_mu=per_instance_mu)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _record(name, rec):
    path = os.path.join(ROOT, "gpurun_out", "bulk_parity.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except ValueError:
            data = {}
    data[name] = rec
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def _compare(name, r, o, b, h, extra=None):
    u = r["u"].astype(float)
    eot = bulk.err_over_tol(u, o["x_star"])
    worst = eot.max(axis=1)
    d_conv = np.abs(u - o["x_conv"]).max(axis=1)
    bound = np.maximum(parity.RTOL * np.abs(o["x_conv"]).max(axis=1), np.abs(o["x_conv"] - o["x_star"]).max(axis=1)) + parity.ATOL
    swing = np.repeat(b["gait"].reshape(len(u), 4 * h) == 0, 3, axis=1)
    rec = dict(n=int(len(u)), n_fail=int((worst > 1.0).sum()), worst_err_over_tol=float(worst.max()),
               median_err_over_tol=float(np.median(worst)), max_abs_err_N=float(np.abs(u - o["x_star"]).max()),
               qpoases_worst_err_over_tol=float(bulk.err_over_tol(o["x_conv"], o["x_star"]).max()),
               qpoases_max_abs_err_N=float(np.abs(o["x_conv"] - o["x_star"]).max()),
               n_further_from_qpoases_than_bound=int((d_conv > bound).sum()),
               status_nonzero=int((r["status"] != 0).sum()), ipm_instances=int((r["iters"][:, 0] > 0).sum()),
               rounds_mean=float(r["iters"][:, 1].mean()), stock_capped=int((o["stock"] != 0).sum()),
               qpoases_nwsr_mean=float(o["nwsr"].mean()))
    if extra:
        rec.update(extra)
    _record(name, rec)
    assert (r["status"] == 0).all(), (name, np.bincount(r["status"]))
    assert rec["n_fail"] == 0, (name, rec)
    assert rec["n_further_from_qpoases_than_bound"] == 0, (name, rec)
    assert (u[swing] == 0).all(), name
    assert np.array_equal(r["grf"], r["u"][:, :12])
    return rec


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_bulk_parity_vs_exact_optimum(cfg, gpu, pkg, monkeypatch):
    name, robot, h, dt, gait, mu_sweep, B, with_ref = cfg
    monkeypatch.setenv("MINI_EIGEN_EXP_NILPOTENT3", "1")   # reference build evaluates exp() of the nilpotent matrix by its series
    b = pkg.synth.make_mpc_batch(robot, h, dt, B, seed=700 + len(name), gait=gait, mu_sweep=mu_sweep)
    P = gpu.params_of(b["robot"], h, dt)
    r = _solve(gpu, P, b, per_instance_mu=mu_sweep)
    o = bulk.run(b, h, dt, np.arange(B), want_ref=with_ref)
    extra = {}
    if with_ref and o["x_ref"] is not None:
        fin = o["stock"] == 0
        ref_gap = np.abs(o["x_ref"] - o["x_star"][:, :12]).max(axis=1)
        d = np.abs(r["grf"].astype(float) - o["x_ref"]).max(axis=1)
        ok = d <= ref_gap + parity.RTOL * np.abs(o["x_star"][:, :12]).max(axis=1) + parity.ATOL
        extra = dict(ref_source_build_finished=int(fin.sum()), ref_source_build_fail=int((~ok[fin]).sum()),
                     ref_source_build_max_gap_N=float(ref_gap[fin].max()) if fin.any() else 0.0,
                     ref_source_build_capped_max_gap_N=float(ref_gap[~fin].max()) if (~fin).any() else 0.0)
    rec = _compare(name, r, o, b, h, extra)
    if extra:
        assert extra["ref_source_build_finished"] >= B // 2
        assert extra["ref_source_build_fail"] == 0, rec
        # (ref_source_build_max_gap_N, the reference build's own distance from x* where its stock run finishes, is
        # recorded, not bounded: qpOASES' termination criterion leaves up to ~0.1 N on some Lite3 instances)


def test_bulk_parity_horizon30_and_forced_fallback(gpu, pkg):
    """BASELINE configs[4] (h = 30, 360 variables): 256 instances against x*, once on the default path and once with the
    block active-set iteration cut off after one round (max_as_rounds = 1), so that EVERY instance goes through the
    interior-point fallback and its verification rounds (qp_solver.h qr_ipm / stage 3)."""
    h, dt, B = 30, 0.03, 256
    b = pkg.synth.make_mpc_batch("a1", h, dt, B, seed=730, gait="trot")
    P = gpu.params_of(b["robot"], h, dt)
    o = bulk.run(b, h, dt, np.arange(B))
    r = _solve(gpu, P, b)
    _compare("a1_h30_trot", r, o, b, h)
    opt = gpu.default_options()
    opt.max_as_rounds = 1
    opt.flags = gpu.QP_NO_PREDICTION
    rf = _solve(gpu, P, b, opt=opt)
    rec = _compare("a1_h30_trot_forced_ipm", rf, o, b, h)
    assert rec["ipm_instances"] >= B - 2   # (an instance whose optimum has no active row at all needs no second round)


def test_forced_fallback_h10(gpu, pkg):
    """The same forced interior-point path at h = 10 on 2048 mixed-gait instances."""
    h, dt, B = 10, 0.03, 2048
    b = pkg.synth.make_mpc_batch("aliengo", h, dt, B, seed=731, gait="mixed")
    P = gpu.params_of(b["robot"], h, dt)
    o = bulk.run(b, h, dt, np.arange(B))
    opt = gpu.default_options()
    opt.max_as_rounds = 1
    opt.flags = gpu.QP_NO_PREDICTION
    rf = _solve(gpu, P, b, opt=opt)
    rec = _compare("aliengo_h10_mixed_forced_ipm", rf, o, b, h)
    assert rec["ipm_instances"] >= B * 0.9


def test_full_size_kkt_certificates(gpu, pkg):
    """BASELINE's batch of 65536 A1 trot instances: an independent NNLS KKT certificate (oracle-built QP, multipliers of
    the right sign on the rows active at the GPU's answer) on 1024 instances drawn from the whole batch."""
    h, dt, B = 10, 0.03, 65536
    b = pkg.synth.make_mpc_batch("a1", h, dt, B, seed=3, gait="trot")
    P = gpu.params_of(b["robot"], h, dt)
    r = _solve(gpu, P, b)
    assert (r["status"] == 0).all()
    idx = np.sort(np.random.default_rng(2).choice(B, 1024, replace=False))
    k = bulk.run(b, h, dt, idx, candidate=r["u"][idx], kkt_only=True)["kkt"]
    _record("a1_h10_trot_65536_kkt", dict(n=1024, stationarity_max=float(k[:, 0].max()), feasibility_max=float(k[:, 1].max())))
    # float32 output rounding (<= 8e-6 N) times ||H|| bounds the visible stationarity residual
    assert k[:, 0].max() < 5e-7 and k[:, 1].max() < 1e-4, k.max(axis=0)
