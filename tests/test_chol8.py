"""The tensor-core factorisation of the long-horizon size classes (csrc/chol8.h: 8x8 tiles, mma.sync.m8n8k4.f64).

* tests/cpp/chol8_test.cu drives qr_chol8_factor / qr_chol8_backward directly on random SPD systems of 47..216 variables
  (sizes that are no multiple of 3 or 8 included) next to the scalar 3x3-block LDL' of csrc/qp_solver.h and checks the
  residual |K x - b| of both on the host.
* Through the C ABI: 256 h = 30 instances solved with the default path and with QR_QP_SCALAR_FACTOR (every reduced system
  by the scalar code); both end on verified KKT conditions of the same QP, so the forces must agree to solve accuracy.
  (tests/test_gpu_bulk_parity.py compares the default path with the oracle's exact optimum.)"""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "chol8_test.cu")
EXE = os.path.join(ROOT, "tests", "cpp", "chol8_test")
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait", "mu")


def build_chol8_test():
    from quadruped_robot_b200 import build
    csrc = [os.path.join(build.CSRC, f) for f in ("chol8.h", "qp_solver.h", "qr_team.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(f) for f in csrc + [SRC]):
        subprocess.run([build.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                        "-o", EXE, SRC], check=True)
    return EXE


def test_chol8_harness_compiles(pkg):
    """CPU: nvcc cross-compiles the harness for sm_100a."""
    assert os.path.exists(build_chol8_test())


@pytest.mark.gpu
def test_chol8_residuals(pkg):
    out = subprocess.run([build_chol8_test()], check=True, capture_output=True, text=True, timeout=300).stdout
    rows = re.findall(r"nred\s+(\d+) .*?(chol8 DMMA|ldl 3x3)\s*: residual ([0-9.e+-]+) \(\|x\| ([0-9.e+-]+)\)\s+(\d+) cycles", out)
    assert len(rows) == 18, out
    for nred, which, res, nrm, cyc in rows:
        assert float(res) <= 1e-12 * max(1.0, float(nrm)), (nred, which, res)
    c = {(int(n), w): int(cy) for n, w, _, _, cy in rows}
    # the reason the path exists: at the h = 30 sizes it is the faster of the two
    assert c[(135, "chol8 DMMA")] < c[(135, "ldl 3x3")] and c[(216, "chol8 DMMA")] < c[(216, "ldl 3x3")], out


@pytest.mark.gpu
def test_tensor_core_path_agrees_with_scalar_path(gpu, pkg):
    import torch
    h, dt, B = 30, 0.03, 256
    b = pkg.synth.make_mpc_batch("a1", h, dt, B, seed=930, gait="trot")
    P = gpu.params_of(b["robot"], h, dt)
    dev = {k: torch.from_numpy(np.ascontiguousarray(b[k])).cuda() for k in KEYS}

    def solve(opt):
        nonlocal dev, B
        out = dict(grf=torch.empty((B, 12), device="cuda"), u=torch.empty((B, 12 * h), device="cuda"),
                   status=torch.empty(B, dtype=torch.int32, device="cuda"),
                   iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))
        gpu.mpc_solve_batch_device(P, dev, out, torch.cuda.current_stream().cuda_stream, opt=opt)
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    a = solve(None)
    # batches <= SM count take the latency kernel (same 256-thread team, same tensor-core path)
    dev_all, B_all = dev, B
    dev = {k: v[:64].contiguous() for k, v in dev_all.items()}
    B = 64
    small = solve(None)
    dev, B = dev_all, B_all
    # (its coarse levels run up to four rounds instead of two, so a weakly active row may end up on the other side:
    # same optimum, last float32 digit at most)
    assert np.abs(small["u"] - a["u"][:64]).max() <= 1e-6 * max(1.0, np.abs(a["u"]).max()), np.abs(small["u"] - a["u"][:64]).max()
    opt = gpu.default_options()
    opt.flags = gpu.QP_SCALAR_FACTOR
    s = solve(opt)
    assert (a["status"] == 0).all() and (s["status"] == 0).all()
    # both are the verified optimum of the same strictly convex QP: float32 outputs, 1e-6 N of slack for the rare
    # instance whose last digit rounds the other way
    assert np.abs(a["u"] - s["u"]).max() <= 1e-6 * max(1.0, np.abs(s["u"]).max())
    # (the 16-bit hashes of the two iterations may differ in the round count of near-degenerate instances only)
    assert abs(a["iters"][:, 1].mean() - s["iters"][:, 1].mean()) < 0.25
