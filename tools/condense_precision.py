"""What a tensor-core (FP64-accumulating) condensing GEMM would do to the answer: python tools/condense_precision.py [h] [count]

For `count` synthetic A1 trot instances the oracle's float32 Bqp (the reference's operation order) is condensed twice:
  H32 = the reference's float32 products and sums (oracle restatement of qr_mpc_interface.cpp:396-412),
  H64 = Bqp' diag(2w) Bqp + 2 alpha I with every product and sum in float64, rounded to float32 once at the end
        (what a DMMA condensing kernel would deliver; g is left as the float32 build's).
Both QPs go through converged qpOASES + the extended-precision optimum; printed is the distance of the two optima in units
of the parity tolerance 1e-4 |x| + 1e-5.  CPU only (test infrastructure: uses oracle/)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkg
pkg = _pkg.load()
import oracle as O
h = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dt = 0.03
b = pkg.synth.make_mpc_batch("a1", h, dt, cnt, seed=730, gait="trot")
P = O.params_of(b["robot"], h, dt)
A = O.constraint_rows(h, P.mu)
w = np.asarray(list(P.weights) + [0.0], np.float32)
worst, worst_abs, relH = 0.0, 0.0, 0.0
for i in range(cnt):
    H, g, ub, Aqp, Bqp = O.mpc_build(P, b, i, want_ab=True)
    w2 = np.tile(np.float32(2) * w, h).astype(np.float64)
    B64 = Bqp.astype(np.float64)
    H64 = (B64.T * w2) @ B64 + 2.0 * float(np.float32(P.alpha)) * np.eye(12 * h)
    H64 = H64.astype(np.float32)
    relH = max(relH, float(np.abs(H64 - H).max() / np.abs(H).max()))
    xs = []
    for Hm in (H, H64):
        xq, info, _, cstat = O.mpc_qpoases(h, P.mu, Hm, g, ub, 100000)
        assert info[0] == 0
        x, _, _ = O.exact_optimum(Hm, g, A, np.zeros(20 * h), ub.astype(float), cstat)
        xs.append(x)
    d = np.abs(xs[0] - xs[1])
    worst = max(worst, float((d / (1e-4 * np.abs(xs[0]) + 1e-5)).max()))
    worst_abs = max(worst_abs, float(d.max()))
print(f"h = {h}, {cnt} instances: max |H64 - H32| / max |H| = {relH:.2e}; optima differ by up to {worst_abs:.3e} N = {worst:.1f} x the parity tolerance")
