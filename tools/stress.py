"""GPU stress run across robots / horizons / gaits: default path (coarse prediction) vs the cold-start path
(QR_QP_NO_PREDICTION) vs the host build of the device sources; both GPU paths must report status 0 everywhere and agree
on the forces to float32 output rounding, because each ends on verified KKT conditions of the same strictly convex QP."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
import emul_binding
em = emul_binding.load()
capi.init(0)
cold = capi.default_options(); cold.flags = capi.QP_NO_PREDICTION
worst = 0
for robot, h, dt, gait, B, mus in (("a1", 10, 0.03, "trot", 20000, False), ("aliengo", 10, 0.03, "mixed", 12000, True), ("lite3", 5, 0.06, "trot", 9000, False),
                                   ("a1", 16, 0.03, "walk", 600, True), ("a1", 12, 0.03, "stand", 300, False), ("a1", 30, 0.03, "mixed", 96, False),
                                   ("lite3", 7, 0.05, "gallop", 2000, True), ("a1", 1, 0.03, "trot", 500, False), ("a1", 2, 0.03, "stand", 500, True),
                                   ("a1", 3, 0.03, "walk", 500, False), ("a1", 32, 0.03, "stand", 6, False), ("lite3", 10, 0.03, "trot", 8192, True)):
    b = pkg.synth.make_mpc_batch(robot, h, dt, B, seed=hash((robot, h, gait)) % 1000, gait=gait, mu_sweep=mus)
    P = capi.params_of(b["robot"], h, dt)
    r = capi.mpc_solve_batch_host(P, b, per_instance_mu=mus, want_u=True)
    c = capi.mpc_solve_batch_host(P, b, opt=cold, per_instance_mu=mus, want_u=True)
    ne = min(B, 400)
    sub = {k: (np.ascontiguousarray(v[:ne]) if isinstance(v, np.ndarray) and v.shape[:1] == (B,) else v) for k, v in b.items()}
    e = em.solve(P, sub, per_instance_mu=mus)
    d_cold = np.abs(r["u"] - c["u"]).max()
    d_emul = np.abs(r["u"][:ne] - e["u64"]).max()
    zeros_same = np.array_equal(r["u"] == 0, c["u"] == 0)
    print(f"{robot} h={h} {gait} B={B} mu_sweep={mus}: status {np.bincount(r['status'])} cold {np.bincount(c['status'])} emul {np.bincount(e['status'])} "
          f"max|pred-cold| {d_cold:.2e} max|gpu-emul| {d_emul:.2e} zeros {zeros_same} rounds {r['iters'][:,1].mean():.2f} (cold {c['iters'][:,1].mean():.2f}) "
          f"max {r['iters'][:,1].max()} ipm {int((r['iters'][:,0]>0).sum())} (cold {int((c['iters'][:,0]>0).sum())})", flush=True)
    worst = max(worst, d_cold, d_emul)
    assert (r["status"] == 0).all() and (c["status"] == 0).all() and zeros_same
print("worst", worst)
