"""Stream / thread contract of include/qr_gpu.h on the GPU (`pytest -m gpu`): device-pointer calls issued concurrently
on several streams and from several host threads, and *_host calls from several threads, must give bit for bit the
results of the same calls issued one after the other.  (Round 1 bound every launch to one process-global scratch and
work list; the calls below then raced on the L2-resident Hessian slices, counters and tickets.)"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
KEYS = ("p", "v", "quat", "w", "r_feet", "rpy", "traj", "gait", "mu")


def _dev(b):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(b[k])).cuda() for k in KEYS}


def _outs(B, h):
    import torch
    return dict(grf=torch.empty((B, 12), device="cuda"), u=torch.empty((B, 12 * h), device="cuda"),
                status=torch.empty(B, dtype=torch.int32, device="cuda"), iters=torch.empty((B, 2), dtype=torch.int32, device="cuda"))


WORK = [("a1", 10, 0.03, 6000, "trot"), ("lite3", 5, 0.06, 9000, "trot"), ("aliengo", 10, 0.03, 5000, "mixed"),
        ("a1", 16, 0.03, 1500, "trot")]


def test_two_streams_two_threads_match_serial(gpu, pkg):
    import torch
    jobs = []
    for k, (robot, h, dt, B, gait) in enumerate(WORK):
        b = pkg.synth.make_mpc_batch(robot, h, dt, B, seed=900 + k, gait=gait)
        jobs.append(dict(P=gpu.params_of(b["robot"], h, dt), dev=_dev(b), B=B, h=h))
    # serial reference: one stream, one call after the other
    serial = []
    for j in jobs:
        o = _outs(j["B"], j["h"])
        gpu.mpc_solve_batch_device(j["P"], j["dev"], o, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (o["status"] == 0).all()
        serial.append({k: v.cpu().numpy() for k, v in o.items()})
    # concurrent: every job on its own stream, issued from its own host thread, three rounds back to back
    streams = [torch.cuda.Stream() for _ in jobs]
    outs = [[_outs(j["B"], j["h"]) for _ in range(3)] for j in jobs]
    errors = []

    def issue(k):
        try:
            torch.cuda.set_device(0)
            gpu.init(0)
            for rep in range(3):
                gpu.mpc_solve_batch_device(jobs[k]["P"], jobs[k]["dev"], outs[k][rep], streams[k].cuda_stream)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=issue, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for k in range(len(jobs)):
        for rep in range(3):
            for key in ("u", "grf", "status", "iters"):
                assert np.array_equal(outs[k][rep][key].cpu().numpy(), serial[k][key]), (k, rep, key)


def test_mpc_and_qp_and_wbc_on_different_streams(gpu, pkg, oracle):
    """An MPC solve, a QP-only solve on caller-supplied data and a WBC batch in flight at once."""
    import torch
    h, dt = 10, 0.03
    b = pkg.synth.make_mpc_batch("a1", h, dt, 4000, seed=910, gait="trot")
    P = gpu.params_of(b["robot"], h, dt)
    dev = _dev(b)
    Bq = 256
    n = 12 * h
    H = torch.empty((Bq, n, n), device="cuda"); g = torch.empty((Bq, n), device="cuda"); ub = torch.empty((Bq, 20 * h), device="cuda")
    sub = {k: v[:Bq].contiguous() for k, v in dev.items()}
    gpu.mpc_condense_batch_device(P, sub, H, g, ub, torch.cuda.current_stream().cuda_stream)
    wb = pkg.synth.make_wbc_batch("lite3", 4096, seed=911)
    M = gpu.wbc_model_of(wb["robot"])
    M2 = gpu.wbc_model_of(pkg.robots.ROBOTS["a1"])
    wb2 = pkg.synth.make_wbc_batch("a1", 4096, seed=912)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ws, wc, wk = t(wb["state"]), t(wb["cmd"]), t(wb["contact"])
    ws2, wc2, wk2 = t(wb2["state"]), t(wb2["cmd"]), t(wb2["contact"])
    torch.cuda.synchronize()

    def run(streams):
        o = _outs(4000, h)
        x64 = torch.empty((Bq, n), dtype=torch.float64, device="cuda")
        tau = torch.empty((4096, 12), device="cuda"); tau2 = torch.empty((4096, 12), device="cuda")
        gpu.mpc_solve_batch_device(P, dev, o, streams[0])
        gpu.qp_solve_batch_device(h, P.mu, H, g, ub, None, x64, None, None, streams[1])
        gpu.wbc_solve_batch_device(M, ws, wc, wk, tau, streams[2])
        gpu.wbc_solve_batch_device(M2, ws2, wc2, wk2, tau2, streams[3])   # a second robot model while the first is in flight
        torch.cuda.synchronize()
        return [o["u"].cpu().numpy(), x64.cpu().numpy(), tau.cpu().numpy(), tau2.cpu().numpy()]

    cur = torch.cuda.current_stream().cuda_stream
    serial = run([cur] * 4)
    ss = [torch.cuda.Stream() for _ in range(4)]
    for _ in range(3):
        conc = run([s.cuda_stream for s in ss])
        for a, c in zip(serial, conc):
            assert np.array_equal(a, c)
    assert np.abs(serial[0][:Bq] - serial[1]).max() < 2e-5   # fused path = condense + QP-only path on the same data


def test_host_entry_points_from_threads(gpu, pkg):
    """qr_gpu_mpc_solve_batch_host (both its packed small-batch path and its two-chunk path) and
    qr_gpu_wbc_solve_batch_host called from four threads at once."""
    import torch
    specs = [("a1", 10, 0.03, 9000, "trot"), ("lite3", 5, 0.06, 40, "trot"), ("aliengo", 10, 0.03, 8300, "mixed"),
             ("a1", 10, 0.03, 3, "trot")]
    batches, params, serial = [], [], []
    for k, (robot, h, dt, B, gait) in enumerate(specs):
        b = pkg.synth.make_mpc_batch(robot, h, dt, B, seed=920 + k, gait=gait)
        P = gpu.params_of(b["robot"], h, dt)
        batches.append(b); params.append(P)
        serial.append(gpu.mpc_solve_batch_host(P, b, want_u=True))
    wb = pkg.synth.make_wbc_batch("lite3", 2000, seed=925)
    M = gpu.wbc_model_of(wb["robot"])
    wser = gpu.wbc_solve_batch_host(M, wb["state"], wb["cmd"], wb["contact"])
    results = [None] * len(specs)
    wres = [None]
    errors = []

    def mpc(k):
        try:
            gpu.init(0)
            for _ in range(3):
                results[k] = gpu.mpc_solve_batch_host(params[k], batches[k], want_u=True)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    def wbc():
        try:
            gpu.init(0)
            for _ in range(3):
                wres[0] = gpu.wbc_solve_batch_host(M, wb["state"], wb["cmd"], wb["contact"])
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=mpc, args=(k,)) for k in range(len(specs))] + [threading.Thread(target=wbc)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(len(specs)):
        for key in ("u", "grf", "status", "iters"):
            assert np.array_equal(results[k][key], serial[k][key]), (k, key)
    for key in ("tau", "fr", "qdes", "qddes", "status"):
        assert np.array_equal(wres[0][key], wser[key]), key
    torch.cuda.synchronize()


def test_uninitialised_device_is_an_error(gpu, pkg):
    """Calls on a device without a context fail with a code (no silent launch on another GPU's buffers)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    b = pkg.synth.make_mpc_batch("a1", 10, 0.03, 4, seed=1)
    P = gpu.params_of(b["robot"], 10, 0.03)
    errs = []

    def other():
        torch.cuda.set_device(1)
        try:
            gpu.mpc_solve_batch_host(P, b)
        except gpu.QrGpuError as e:
            errs.append(e)

    t = threading.Thread(target=other)
    t.start(); t.join()
    assert errs


def test_multi_gpu_host_call_matches_single_gpu(gpu, pkg):
    """qr_gpu_mpc_solve_batch_host_multi: one host batch sharded over every GPU of the box, gathered into the caller's
    arrays -- bit-equal to the unsharded call (instances are independent; there is no collective)."""
    import torch
    G = torch.cuda.device_count()
    h, dt, B = 10, 0.03, 20000 + 7
    b = pkg.synth.make_mpc_batch("aliengo", h, dt, B, seed=930, gait="mixed", mu_sweep=True)
    P = gpu.params_of(b["robot"], h, dt)
    one = gpu.mpc_solve_batch_host(P, b, per_instance_mu=True, want_u=True)
    assert (one["status"] == 0).all()
    for devs in ([0], list(range(G)), list(range(G))[::-1]):
        r = gpu.mpc_solve_batch_host_multi(P, b, devs, per_instance_mu=True, want_u=True)
        for key in ("u", "grf", "status", "iters"):
            assert np.array_equal(r[key], one[key]), (devs, key)
    torch.cuda.set_device(0)
    with pytest.raises(gpu.QrGpuError):
        gpu.mpc_solve_batch_host_multi(P, b, [0, 0])
