"""The C-ABI library loads and exports every function include/qr_gpu.h declares (no compute calls:
this runs without a GPU), and refuses to work without one instead of falling back."""
import os
import re

import pytest


def _declared():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "qr_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qr_gpu_\w+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    from quadruped_robot_b200 import build, capi
    build.build()
    lib = capi.lib()
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qr_gpu.h but not exported"
    assert set(names) == set(capi.EXPORTS)


def test_no_cpu_fallback(pkg):
    """Without a device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from quadruped_robot_b200 import capi
    with pytest.raises(capi.QrGpuError):
        capi.init(0)
    import numpy as np
    b = pkg.synth.make_mpc_batch("a1", 10, 0.03, 2, seed=0)
    P = capi.params_of(b["robot"], 10, 0.03)
    with pytest.raises(capi.QrGpuError):
        capi.mpc_solve_batch_host(P, b)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkgdir = os.path.join(root, "quadruped-robot_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "qr_oracle.h" not in txt and "libqr_oracle" not in txt, f
