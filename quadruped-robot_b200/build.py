"""Build recipe of the product library: nvcc -> quadruped-robot_b200/libqr_gpu.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libqr_gpu.so")
PEAKS_LIB = os.path.join(_HERE, "libqr_peaks.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["mpc_kernels.cu"]
HEADERS = ["qr_team.h", "mpc_condense.h", "qp_solver.h", "mpc_problem.h", "wbc_model.h", "wbc_problem.h", "mpc_io.h", "small_qp.h", "fb_problem.h", "swing_extra.h", "chol8.h", "ctl_extra.h"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(_HERE, "..", "include", "qr_gpu.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libqr_gpu.so (in-tree)."""
    if force or _stale():
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.run(cmd, check=True, cwd=CSRC)
    peaks_src = os.path.join(CSRC, "peaks.cu")
    if force or not os.path.exists(PEAKS_LIB) or os.path.getmtime(peaks_src) > os.path.getmtime(PEAKS_LIB):
        subprocess.run([NVCC] + FLAGS + ["-o", PEAKS_LIB, peaks_src], check=True, cwd=CSRC)
    return LIB
