"""Mint tests/golden/wbc_*.npz from the WBC oracle (float32 = the reference's arithmetic, float64 = the same
algorithm in double; both call the reference's QuadProg++ from oracle/_ref).  Run in the build container:
    python tests/golden/make_golden_wbc.py
"""
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
CONFIGS = [dict(name="lite3", robot="lite3", seed=200, B=12), dict(name="a1", robot="a1", seed=201, B=12)]
FIELDS = ("tau", "fr", "qdes", "qddes", "H", "G", "C", "Jc", "Jcdqd", "pGC", "vGC", "qdd")


def main():
    O.build()
    for cfg in CONFIGS:
        b = pkg.synth.make_wbc_batch(cfg["robot"], cfg["B"], seed=cfg["seed"])
        M = O.wbc_model_of(b["robot"])
        out = dict(state=b["state"], cmd=b["cmd"], contact=b["contact"], meta=np.array([cfg["robot"], str(cfg["seed"])]))
        for prec in ("f64", "f32"):
            rs = [O.wbc_step(M, b["state"][i], b["cmd"][i], b["contact"][i], prec) for i in range(cfg["B"])]
            for f in FIELDS:
                out[f"{f}_{prec}"] = np.stack([r[f] for r in rs])
        path = os.path.join(HERE, f"wbc_{cfg['name']}.npz")
        np.savez_compressed(path, **out)
        noise = np.abs(out["tau_f32"] - out["tau_f64"]).max()
        print(path, os.path.getsize(path) // 1024, "KiB; reference float32 noise on tau:", noise)


if __name__ == "__main__":
    main()
