import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
capi.init(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
batch = pkg.synth.make_mpc_batch("a1", 10, 0.03, nb, seed=5, gait="trot")
P = capi.params_of(batch["robot"], 10, 0.03)
for _ in range(2):
    r = capi.mpc_solve_batch_host(P, batch)
print("ok", r["status"].max(), r["iters"][:, 1].mean())
