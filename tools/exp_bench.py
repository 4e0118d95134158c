"""Throughput of alternative builds of the fused kernel: python scratch/exp_bench.py --build NAME -D... | --run NAME"""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import build as B, capi
name = sys.argv[2]
lib_path = os.path.join(ROOT, "scratch", f"libqr_{name}.so")
if sys.argv[1] == "--build":
    cmd = [B.NVCC] + B.FLAGS + ["-Xptxas", "-v"] + [a for a in sys.argv[3:]] + ["-o", lib_path, os.path.join(B.CSRC, "mpc_kernels.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    lines = (r.stdout + r.stderr).split("\n")
    for i, l in enumerate(lines):
        if "fused_kernelILi24" in l and "Compiling" in l:
            print(name, lines[i + 2].strip(), "|", lines[i + 3].strip())
    sys.exit(r.returncode)
import torch
B.LIB = lib_path
capi.init(0)
h, dt, nb = 10, 0.03, 65536
sets = [pkg.synth.make_mpc_batch("a1", h, dt, nb, seed=s, gait="trot") for s in range(3)]
P = capi.params_of(sets[0]["robot"], h, dt)
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")
dev = [{k: torch.from_numpy(b[k]).cuda() for k in KEYS} for b in sets]
out = dict(grf=torch.empty((nb, 12), device="cuda"), status=torch.empty(nb, dtype=torch.int32, device="cuda"),
           iters=torch.empty((nb, 2), dtype=torch.int32, device="cuda"))
st = torch.cuda.current_stream().cuda_stream
for i in range(3):
    capi.mpc_solve_batch_device(P, dev[i % 3], out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(6):
    capi.mpc_solve_batch_device(P, dev[i % 3], out, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 6
one = {k: np.ascontiguousarray(sets[0][k][:1]) for k in KEYS}
ts = []
for i in range(600):
    a = time.perf_counter(); capi.mpc_solve_batch_host(P, one); ts.append(time.perf_counter() - a)
ts = np.asarray(ts[100:]) * 1e6
print(f"{name}: latency p50 {np.percentile(ts, 50):.1f} us p99 {np.percentile(ts, 99):.1f} us")
g = out['grf'].double()
print(f"{name}: checksum grf sum {float(g.sum()):.6f} abs-sum {float(g.abs().sum()):.6f} rounds {float(out['iters'][:,1].float().mean()):.4f}")
print(f"{name}: {nb / ms * 1e3 / 1e6:.3f} M QP/s  ({ms:.2f} ms/step) occupancy {capi.occupancy(h, 24)} bad {int((out['status'] != 0).sum())}")
