"""WBC kernel throughput of an alternative build: python tools/exp_wbc.py NAME   (scratch/libqr_NAME.so, or main)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import build as B, capi
name = sys.argv[1]
if name != "main":
    B.LIB = os.path.join(ROOT, "scratch", f"libqr_{name}.so")
import torch
capi.init(0)
st = torch.cuda.current_stream().cuda_stream
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
M = capi.wbc_model_of(pkg.robots.ROBOTS["lite3"])
for nb in (1024, 65536):
    wb = pkg.synth.make_wbc_batch("lite3", nb, seed=6)
    state, cmd, contact = dev(wb["state"]), dev(wb["cmd"]), dev(wb["contact"])
    tau = torch.empty((nb, 12), dtype=torch.float64, device="cuda")
    stt = torch.empty(nb, dtype=torch.int32, device="cuda")
    f = lambda: capi.wbc_solve_batch_device_f64(M, state, cmd, contact, tau, st, status=stt)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name} wbc B={nb}: {nb / ms * 1e3 / 1e6:.3f} M robots/s ({ms:.3f} ms) status!=0 {int((stt != 0).sum())} checksum {float(tau.sum()):.9f} {float(tau.abs().sum()):.6f}", flush=True)
