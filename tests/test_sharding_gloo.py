"""N > 1 host logic on CPU: world_size 2 over gloo.  Instances are independent, so ranks shard the batch
contiguously with no data-path collective; the only communication is the max-over-ranks of the elapsed
time (bench.py) and an optional gather of the forces.  Each rank 'solves' its shard with the host
emulation of the device code (tests only) and the gathered result must equal the unsharded one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_range(batch, rank, world):
    """Contiguous shard [lo, hi) of rank."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _worker(rank, world, port, batch, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _pkg
    import emul_binding
    pkg = _pkg.load()
    from quadruped_robot_b200 import capi
    E = emul_binding.load()
    full = pkg.synth.make_mpc_batch("a1", 10, 0.03, batch, seed=77, gait="mixed")
    lo, hi = shard_range(batch, rank, world)
    mine = {k: (np.ascontiguousarray(v[lo:hi]) if isinstance(v, np.ndarray) else v) for k, v in full.items()}
    P = capi.params_of(full["robot"], 10, 0.03)
    r = E.solve(P, mine)
    # the two collectives bench.py / a gathering caller use
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sizes = [b - a for a, b in (shard_range(batch, q, world) for q in range(world))]
    pad = max(sizes)                      # all_gather wants equal shapes: pad the shorter shards
    mine_padded = torch.zeros((pad, 12), dtype=torch.float32)
    mine_padded[:hi - lo] = torch.from_numpy(r["grf"])
    parts = [torch.zeros((pad, 12), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(parts, mine_padded)
    if rank == 0:
        grf = torch.cat([p[:n] for p, n in zip(parts, sizes)]).numpy()
        np.savez(out_path, grf=grf, tmax=t.numpy(), status=r["status"])
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for batch in (0, 1, 7, 64, 65536):
        for world in (1, 2, 4, 8):
            r = [shard_range(batch, q, world) for q in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_unsharded(tmp_path, pkg, emul):
    batch, world = 13, 2
    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, batch, out), nprocs=world, join=True)
    z = np.load(out)
    from quadruped_robot_b200 import capi
    full = pkg.synth.make_mpc_batch("a1", 10, 0.03, batch, seed=77, gait="mixed")
    P = capi.params_of(full["robot"], 10, 0.03)
    ref = emul.solve(P, full)
    assert np.array_equal(z["grf"], ref["grf"])
    assert float(z["tmax"][0]) == 0.2
