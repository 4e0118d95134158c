"""Golden vectors of the force-balance stance QP: inputs from the seeded generator, forces from the oracle
(restated ComputeContactForce + the reference's own QuadProg++ compiled from /root/reference).
Run from the repo root: python tests/golden/make_golden_fb.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkg  # noqa: E402

pkg = _pkg.load()
import oracle as O  # noqa: E402

O.build()
B, seed = 48, 41
b = pkg.synth.make_fb_batch("a1", B, seed=seed, world_frame=False, tilted=True)
P = O.fb_params_of(b["params"])
force = np.stack([O.force_balance(P, b["foot"][i], b["acc"][i], b["contact"][i], b["inertia"][i], b["gravity"][i],
                                  b["frame"][i])["force"] for i in range(B)])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fb_a1.npz"), batch=B, seed=seed,
                    foot=b["foot"], acc=b["acc"], contact=b["contact"], inertia=b["inertia"], gravity=b["gravity"],
                    frame=b["frame"], force=force)
print("wrote fb_a1.npz", force.shape)
