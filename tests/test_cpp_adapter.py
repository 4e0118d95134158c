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
