"""Throughput of alternative builds on other workloads: python scratch/exp2.py NAME  (h30 trot 4096, mixed aliengo 16384, a1 trot 65536)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import build as B, capi
name = sys.argv[1]
which = sys.argv[2:] or ["h30", "mixed", "trot"]
if name != "main":
    B.LIB = os.path.join(ROOT, "scratch", f"libqr_{name}.so")
import torch
capi.init(0)
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")
st = torch.cuda.current_stream().cuda_stream
def run(tag, robot, h, nb, gait, seed, reps):
    mb = pkg.synth.make_mpc_batch(robot, h, 0.03, nb, seed=seed, gait=gait)
    P = capi.params_of(pkg.robots.ROBOTS[robot], h, 0.03)
    d = {k: torch.from_numpy(mb[k]).cuda() for k in KEYS}
    out = dict(grf=torch.empty((nb, 12), device="cuda"), status=torch.empty(nb, dtype=torch.int32, device="cuda"),
               iters=torch.empty((nb, 2), dtype=torch.int32, device="cuda"))
    for i in range(2): capi.mpc_solve_batch_device(P, d, out, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): capi.mpc_solve_batch_device(P, d, out, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name} {tag}: {nb / ms * 1e3 / 1e3:.1f} k QP/s ({ms:.2f} ms) bad {int((out['status'] != 0).sum())} rounds {float(out['iters'][:,1].float().mean()):.2f} max {int(out['iters'][:,1].max())} ipm_inst {int((out['iters'][:,0]>0).sum())} ipm_max {int(out['iters'][:,0].max())} grf abs-sum {float(out['grf'].double().abs().sum()):.6f}", flush=True)
if "h30" in which: run("h30", "a1", 30, 4096, "trot", 14, 3)
if "mixed" in which: run("mixed", "aliengo", 10, 16384, "mixed", 13, 5)
if "trot" in which: run("trot", "a1", 10, 65536, "trot", 0, 5)
if "lite3h5" in which: run("lite3h5", "lite3", 5, 65536, "trot", 3, 5)
for nb in (1024, 2048, 4096):
    if f"l{nb}" in which: run(f"lite3 h10 B={nb}", "lite3", 10, nb, "trot", 3, 20)
