"""Small mixed workload for compute-sanitizer (memcheck / racecheck): python tools/sanitize.py [scale]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
capi.init(0)
s = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
for robot, h, nb, gait, mus in (("lite3", 10, 600, "trot", False), ("aliengo", 10, 600, "mixed", True), ("a1", 30, 200, "trot", False),
                                ("a1", 3, 300, "walk", False), ("a1", 10, 40, "trot", False), ("a1", 10, 9000, "trot", False)):
    nb = max(2, int(nb * s))
    b = pkg.synth.make_mpc_batch(robot, h, 0.03, nb, seed=7, gait=gait, mu_sweep=mus)
    P = capi.params_of(b["robot"], h, 0.03)
    r = capi.mpc_solve_batch_host(P, b, per_instance_mu=mus, want_u=True)
    print(robot, h, gait, nb, "status", np.bincount(r["status"]), "rounds", r["iters"][:, 1].mean(), flush=True)
