"""The C++ drop-in adapter (include/qr_gpu_mpc_adapter.hpp) compiled with g++ against libqr_gpu.so and
driven with the reference controller's call sequence: SetupProblem -> SolveMPCKernel -> GetMPCSolution."""
import os
import struct
import subprocess

import numpy as np
import pytest

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_test")


def build_adapter_test(pkg):
    from quadruped_robot_b200 import build
    lib = build.build()
    libdir = os.path.dirname(lib)
    stale = (not os.path.exists(EXE) or os.path.getmtime(EXE) < max(
        os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "include", "qr_gpu_mpc_adapter.hpp")),
        os.path.getmtime(lib)))
    if stale:
        subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
                        "-L" + libdir, "-lqr_gpu", "-Wl,-rpath," + libdir], check=True)
    return EXE


def test_adapter_compiles_and_links(pkg):
    """CPU: the header compiles as C++17 without Eigen and links against the C ABI."""
    assert os.path.exists(build_adapter_test(pkg))


@pytest.mark.gpu
def test_adapter_matches_golden(pkg, gpu, tmp_path):
    z, b, h, dt, _ = parity.load_golden(os.path.join(parity.HERE, "golden", "mpc_a1_h10_trot.npz"), pkg)
    rb = b["robot"]
    B = b["p"].shape[0]
    blob = struct.pack("ii", B, h)
    cfg = [dt, rb.mu, rb.f_max, rb.mass, 0.0] + list(rb.inertia) + list(rb.weights) + [rb.alpha]
    blob += np.asarray(cfg, np.float32).tobytes()
    for i in range(B):
        for k in ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait"):
            blob += np.ascontiguousarray(b[k][i], np.float32).tobytes()
    path = tmp_path / "inst.bin"
    path.write_bytes(blob)
    out = subprocess.run([build_adapter_test(pkg), str(path)], check=True, capture_output=True, text=True).stdout
    rows = [ln.split() for ln in out.splitlines() if ln.startswith("F ")]
    assert len(rows) == B
    for row in rows:
        i, status = int(row[1]), int(row[2])
        f = np.array([float(x) for x in row[3:]])
        assert status == 0
        parity.assert_elementwise(f, z["x_star"][i][:12], f"adapter[{i}]")


WSRC = os.path.join(ROOT, "tests", "cpp", "wbc_adapter_test.cpp")
WEXE = os.path.join(ROOT, "tests", "cpp", "wbc_adapter_test")


def build_wbc_adapter_test(pkg):
    from quadruped_robot_b200 import build
    lib = build.build()
    libdir = os.path.dirname(lib)
    stale = (not os.path.exists(WEXE) or os.path.getmtime(WEXE) < max(
        os.path.getmtime(WSRC), os.path.getmtime(os.path.join(ROOT, "include", "qr_gpu_wbc_adapter.hpp")),
        os.path.getmtime(lib)))
    if stale:
        subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), WSRC, "-o", WEXE,
                        "-L" + libdir, "-lqr_gpu", "-Wl,-rpath," + libdir], check=True)
    return WEXE


def test_wbc_adapter_compiles_and_links(pkg):
    """CPU: include/qr_gpu_wbc_adapter.hpp compiles as C++17 without Eigen and links against the C ABI."""
    assert os.path.exists(build_wbc_adapter_test(pkg))


@pytest.mark.gpu
def test_wbc_adapter_matches_oracle(pkg, gpu, oracle, tmp_path):
    """UpdateModel -> Run(ctrlData) x3 through the C++ mirror of qrWbcLocomotionController: every-second-call gating,
    and the third tick (previous orientation-velocity command = the current one) against the float64 oracle."""
    B = 12
    wb = pkg.synth.make_wbc_batch("a1", B, seed=61)
    m = gpu.wbc_model_of(wb["robot"])
    blob = struct.pack("i", B) + bytes(m)
    cmd = wb["cmd"].copy()
    cmd[:, 63:66] = cmd[:, 12:15]          # what the adapter feeds on its third tick
    for i in range(B):
        blob += wb["state"][i].tobytes() + wb["cmd"][i].tobytes() + wb["contact"][i].tobytes()
    path = tmp_path / "wbc.bin"
    path.write_bytes(blob)
    out = subprocess.run([build_wbc_adapter_test(pkg), str(path)], check=True, capture_output=True, text=True).stdout
    rows = [ln.split() for ln in out.splitlines() if ln.startswith("W ")]
    assert len(rows) == B, out
    Mo = oracle.wbc_model_of(wb["robot"])
    for row in rows:
        i = int(row[1])
        assert [int(x) for x in row[2:5]] == [0, 0, 0] and int(row[5]) == 1
        vals = np.array([float(x) for x in row[6:]])
        o = oracle.wbc_step(Mo, wb["state"][i], cmd[i], wb["contact"][i], "f64")
        assert np.abs(vals[:12] - o["tau"]).max() < 1e-4 * np.abs(o["tau"]).max() + 1e-5
        assert np.abs(vals[12:] - o["qdes"]).max() < 1e-5
