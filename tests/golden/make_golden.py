"""Mint the golden fixtures of tests/golden/*.npz from the CPU oracle (run in the build container,
where /root/reference is mounted and oracle/_ref is built from it):

    python tests/golden/make_golden.py

Each fixture holds seeded synthetic inputs (quadruped-robot_b200/synth.py) and, per instance, what the
oracle produced: the float32 QP data (g, ub, sha256 of H; H itself for instance 0), the stock
qpOASES answer (nWSR = 100, with its return code), the converged qpOASES answer and working set, and
x* = the extended-precision optimum on that working set (oracle.polish_from_working_set).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkg  # noqa: E402
import oracle as O  # noqa: E402

pkg = _pkg.load()

CONFIGS = [
    dict(name="a1_h10_trot", robot="a1", h=10, dt=0.03, gait="trot", seed=100, B=6, mu_sweep=False),
    dict(name="lite3_h5_trot", robot="lite3", h=5, dt=0.06, gait="trot", seed=101, B=6, mu_sweep=False),
    dict(name="a1_h10_musweep", robot="a1", h=10, dt=0.03, gait="trot", seed=102, B=6, mu_sweep=True),
    dict(name="aliengo_h10_mixed", robot="aliengo", h=10, dt=0.03, gait="mixed", seed=103, B=6, mu_sweep=False),
    dict(name="a1_h16_stand", robot="a1", h=16, dt=0.03, gait="stand", seed=104, B=2, mu_sweep=False),
    # BASELINE.json configs[4]: long-preview MPC, 360 variables.  Beyond the reference's own K_MAX_GAIT_SEGMENTS = 16
    # (its arrays would overflow), so these are checked against the oracle's restatement + qpOASES only.
    dict(name="a1_h30_trot", robot="a1", h=30, dt=0.03, gait="trot", seed=105, B=2, mu_sweep=False),
    dict(name="a1_h30_stand", robot="a1", h=30, dt=0.03, gait="stand", seed=106, B=1, mu_sweep=False),
]
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait", "mu", "f_max")


def main():
    O.build()
    only = sys.argv[1:]
    for cfg in CONFIGS:
        if only and cfg["name"] not in only:
            continue
        b = pkg.synth.make_mpc_batch(cfg["robot"], cfg["h"], cfg["dt"], cfg["B"], seed=cfg["seed"],
                                     gait=cfg["gait"], mu_sweep=cfg["mu_sweep"])
        h, B = cfg["h"], cfg["B"]
        n, m = 12 * h, 20 * h
        out = {k: b[k] for k in KEYS}
        g_all = np.zeros((B, n), np.float32)
        ub_all = np.zeros((B, m), np.float32)
        sha = []
        x_stock = np.zeros((B, n)); info_stock = np.zeros((B, 2), np.int32)
        x_conv = np.zeros((B, n)); info_conv = np.zeros((B, 2), np.int32)
        x_star = np.zeros((B, n)); cstat = np.zeros((B, m), np.int32)
        H0 = None
        for i in range(B):
            P = O.params_of(b["robot"], h, cfg["dt"], mu=float(b["mu"][i]))
            H, g, ub = O.mpc_build(P, b, i)
            if i == 0:
                H0 = H.copy()
            sha.append(hashlib.sha256(H.tobytes()).hexdigest())
            g_all[i], ub_all[i] = g, ub
            x_stock[i], info_stock[i], _, _ = O.mpc_qpoases(h, P.mu, H, g, ub, 100)
            x_conv[i], info_conv[i], _, cstat[i] = O.mpc_qpoases(h, P.mu, H, g, ub, 100000)
            A = O.constraint_rows(h, P.mu)
            x_star[i], _ = O.polish_from_working_set(H, g, A, np.zeros(m), ub.astype(float), cstat[i])
        out.update(g=g_all, ub=ub_all, H_sha256=np.array(sha), H0=H0, x_stock=x_stock, info_stock=info_stock,
                   x_conv=x_conv, info_conv=info_conv, x_star=x_star, cstat=cstat,
                   meta=np.array([cfg["robot"], str(h), str(cfg["dt"]), cfg["gait"], str(cfg["seed"]),
                                  str(int(cfg["mu_sweep"]))]))
        path = os.path.join(HERE, f"mpc_{cfg['name']}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB; capped@100:",
              int((info_stock[:, 0] == O.RET_MAX_NWSR_REACHED).sum()), "of", B)


if __name__ == "__main__":
    main()
