import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
import torch
capi.init(0)
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait")
st = torch.cuda.current_stream().cuda_stream
for (robot,h,nb,gait,seed) in (("lite3",10,4096,"trot",3),("a1",10,16384,"trot",0),("lite3",5,8192,"trot",4),("aliengo",10,8192,"mixed",13)):
    mb = pkg.synth.make_mpc_batch(robot, h, 0.03, nb, seed=seed, gait=gait)
    P = capi.params_of(pkg.robots.ROBOTS[robot], h, 0.03)
    d = {k: torch.from_numpy(mb[k]).cuda() for k in KEYS}
    out = dict(grf=torch.empty((nb, 12), device="cuda"), status=torch.empty(nb, dtype=torch.int32, device="cuda"),
               iters=torch.empty((nb, 2), dtype=torch.int32, device="cuda"))
    capi.mpc_solve_batch_device(P, d, out, st); torch.cuda.synchronize()
    it = out["iters"].cpu().numpy()
    idx = np.nonzero(it[:,0] > 0)[0]
    print(robot, h, gait, seed, nb, "ipm idx", idx.tolist(), "iters", it[idx].tolist(), "rounds hist", np.bincount(it[:,1]).tolist(), flush=True)
