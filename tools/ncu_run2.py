"""One device-resident launch sequence for ncu: python tools/ncu_run2.py ROBOT HORIZON BATCH GAIT"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
capi.init(0)
robot, h, nb, gait = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
b = pkg.synth.make_mpc_batch(robot, h, 0.03, nb, seed=5, gait=gait)
P = capi.params_of(b["robot"], h, 0.03)
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")
d = {k: torch.from_numpy(b[k]).cuda() for k in KEYS}
out = dict(grf=torch.empty((nb, 12), device="cuda"), status=torch.empty(nb, dtype=torch.int32, device="cuda"),
           iters=torch.empty((nb, 2), dtype=torch.int32, device="cuda"))
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    capi.mpc_solve_batch_device(P, d, out, st)
torch.cuda.synchronize()
print("ok", int(out["status"].max()), float(out["iters"][:, 1].float().mean()))
