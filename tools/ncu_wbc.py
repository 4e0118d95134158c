import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
capi.init(0)
st = torch.cuda.current_stream().cuda_stream
B = 2048
wb = pkg.synth.make_wbc_batch("lite3", B, seed=6)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
state, cmd, contact = dev(wb["state"]), dev(wb["cmd"]), dev(wb["contact"])
tau = torch.empty((B, 12), dtype=torch.float32, device="cuda")
for _ in range(2):
    capi.wbc_solve_batch_device(capi.wbc_model_of(wb["robot"]), state, cmd, contact, tau, st)
torch.cuda.synchronize()
print("ok")
