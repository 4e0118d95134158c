"""Phase-cycle profile of the fused kernel (debug build -DQR_PROFILE)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import build as B, capi
lib_path = os.path.join(ROOT, "scratch", "libqr_prof.so")
if "--build" in sys.argv:
    subprocess.run([B.NVCC] + B.FLAGS + ["-DQR_PROFILE"] + [a for a in sys.argv if a.startswith("-D")] + ["-o", lib_path, os.path.join(B.CSRC, "mpc_kernels.cu")], check=True)
    sys.exit(0)
B.LIB = lib_path
capi.init(0)
lib = capi.lib()
NAMES = {20: "stage", 21: "condense", 1: "bases", 3: "pack + H p", 4: "reduced matrix+rhs", 11: "LDL' + fwd solve", 13: "backward solve",
         5: "x = Z y + p", 6: "H x + g", 7: "verify", 22: "scatter", 23: "coarse problem build", 12: "fwd (fallback)"}
H = int(os.environ.get("QR_PROF_H", "10"))
for nb in ((148, 4096) if H <= 16 else (148, 1184)):
    batch = pkg.synth.make_mpc_batch("a1", H, 0.03, nb, seed=5, gait="trot")
    P = capi.params_of(batch["robot"], H, 0.03)
    capi.mpc_solve_batch_host(P, batch)
    tab = (C.c_ulonglong * 64)()
    lib.qr_gpu_debug_profile(tab)
    r = capi.mpc_solve_batch_host(P, batch)
    lib.qr_gpu_debug_profile(tab)
    rounds = r["iters"][:, 1].mean()
    tot = sum(tab)
    print(f"B {nb} cycles per QP {tot / nb:.0f}   rounds {rounds:.2f}")
    for k, name in NAMES.items():
        if tab[k]:
            per = tab[k] / nb
            extra = f"   per round {per / rounds:8.0f}" if k in (1, 3, 4, 11, 13, 5, 6, 7) else ""
            print(f"  {name:20s} {per:8.0f} cyc/QP  {100 * tab[k] / tot:5.1f}%{extra}")
