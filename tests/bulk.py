"""Bulk oracle runs on all host cores (TEST INFRASTRUCTURE: imports oracle/).

The reference keeps its MPC state in file statics and qpOASES has a process-global message handler, so the oracle is
run one process per core (fork), each worker pinned, BLAS limited to one thread.  For every requested instance a
worker returns
    x_star   the exact optimum of the reference's QP: oracle float32 build (restatement of qr_mpc_interface.cpp:359-412)
             -> converged qpOASES (the reference's vendored 3.2.0, setToMPC, cold init) -> extended-precision KKT solve
             on its final working set, the working set corrected until every row is feasible and every multiplier
             has the right sign (oracle.exact_optimum; `fixes` = rows added / dropped, 0 for all but ~0.1 % of instances)
    x_conv   the converged qpOASES answer itself
    stock    return code of the stock nWSR = 100 run (0: finished, 64: working-set cap hit)
    x_ref    (optional) GetMPCSolution(0..11) of the reference's OWN qr_mpc_interface.cpp build (oracle/_ref/
             libqr_mpc_ref.so) at stock nWSR = 100
    kkt      (optional) independent NNLS KKT certificate (stationarity, feasibility) of a candidate solution
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

_JOB = {}


def cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


def _worker(core, idx, cand):
    try:
        os.sched_setaffinity(0, {core})
    except (AttributeError, OSError):
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except ImportError:
        pass
    import oracle as O
    b, h, dt = _JOB["batch"], _JOB["h"], _JOB["dt"]
    want_ref, kkt_only = _JOB["want_ref"], _JOB["kkt_only"]
    n, m = 12 * h, 20 * h
    x_star = np.zeros((len(idx), n))
    x_conv = np.zeros((len(idx), n))
    stock = np.zeros(len(idx), np.int32)
    nwsr = np.zeros(len(idx), np.int32)
    fixes = np.zeros(len(idx), np.int32)
    x_ref = np.zeros((len(idx), 12)) if want_ref else None
    kkt = np.zeros((len(idx), 2)) if cand is not None else None
    A_cache = {}
    for k, i in enumerate(idx):
        mu = float(b["mu"][i])
        P = O.params_of(b["robot"], h, dt, mu=mu)
        if mu not in A_cache:
            A_cache[mu] = O.constraint_rows(h, P.mu)
        A = A_cache[mu]
        H, g, ub = O.mpc_build(P, b, int(i))
        if cand is not None:
            kkt[k] = O.kkt_certificate(H, g, A, np.zeros(m), ub.astype(float), np.asarray(cand[k], float))
            if kkt_only:
                continue
        xq, info, _, cstat = O.mpc_qpoases(h, P.mu, H, g, ub, 100000)
        assert info[0] == 0, ("converged qpOASES failed", i, info)
        x_conv[k] = xq
        nwsr[k] = info[1]
        x_star[k], _, fixes[k] = O.exact_optimum(H, g, A, np.zeros(m), ub.astype(float), cstat)
        _, info100, _, _ = O.mpc_qpoases(h, P.mu, H, g, ub, 100)
        stock[k] = info100[0]
        if want_ref:
            _, _, _, xr = O.ref_mpc_solve(P, b, int(i))
            x_ref[k] = xr[:12]
    return x_star, x_conv, stock, nwsr, x_ref, kkt, fixes


def run(batch, h, dt, idx, want_ref=False, candidate=None, kkt_only=False, nproc=None):
    """Oracle answers for instances `idx` of a synth batch, rows in the order of `idx`.  candidate: [len(idx)][12h]
    solutions to certify with the NNLS KKT check (kkt_only: skip the qpOASES part)."""
    import oracle as O
    O.build()
    O.lib()
    idx = np.asarray(idx, np.int64)
    cs = cores()[:nproc] if nproc else cores()
    bounds = np.linspace(0, len(idx), len(cs) + 1).astype(int)
    parts = [(cs[w], bounds[w], bounds[w + 1]) for w in range(len(cs)) if bounds[w + 1] > bounds[w]]
    _JOB.update(batch=batch, h=h, dt=dt, want_ref=bool(want_ref and O.ref_mpc_available() and h <= 16), kkt_only=kkt_only)
    ctx = mp.get_context("fork")
    with ctx.Pool(len(parts)) as pool:
        asyncs = [pool.apply_async(_worker, (c, idx[lo:hi], None if candidate is None else np.asarray(candidate[lo:hi])))
                  for c, lo, hi in parts]
        res = [a.get() for a in asyncs]
    cat = lambda k: None if res[0][k] is None else np.concatenate([r[k] for r in res])
    return dict(idx=idx, x_star=cat(0), x_conv=cat(1), stock=cat(2), nwsr=cat(3), x_ref=cat(4), kkt=cat(5), fixes=cat(6))


RTOL, ATOL = 1e-4, 1e-5


def err_over_tol(u, x_star):
    """Element-wise |u - x*| / (1e-4 |x*| + 1e-5): the BASELINE tolerance against the exact optimum; <= 1 passes."""
    u = np.asarray(u, float)
    return np.abs(u - x_star) / (RTOL * np.abs(x_star) + ATOL)
