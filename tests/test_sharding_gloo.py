"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous instance shards, no data-path collective,
max-over-ranks timing and summed statistics.  The per-shard compute is done by the ORACLE here (this is a CPU
test of the plumbing; on the GPU box the same shards go through fmpc_step)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nb, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import cases
    import mpc_sensorlessao_b200  # noqa: F401
    from mpc_sensorlessao_b200.shard import reduce_stats, shard_range
    from oracle import fmpc_ref
    fmpc_ref.build()
    c = cases.small_problem(11, 6, 4, 5, nb, 0.4, warm=True)
    lo, hi = shard_range(nb, rank, world)
    sub = dict(c, nb=hi - lo)
    for k in ("x0", "x0_pre", "w", "X0", "U0", "nu0"):
        if c.get(k) is not None:
            sub[k] = c[k][lo:hi]
    ref = cases.ref_solve(fmpc_ref, sub, 5, 0.01, nthreads=1)
    np.save(os.path.join(outdir, f"z_{rank}.npy"), ref["z"])
    tmax, (iters, count) = reduce_stats(10.0 * (rank + 1), [int(ref["iters"].sum()), hi - lo])
    if rank == 0:
        np.save(os.path.join(outdir, "stats.npy"), np.array([tmax, iters, count]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    from mpc_sensorlessao_b200.shard import shard_range
    for nb in (0, 1, 7, 8, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_range(nb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= -(-nb // world)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_rank_sharded_solve_equals_whole_batch(tmp_path, fref):
    nb, world = 7, 2
    mp.spawn(_worker, args=(world, _free_port(), nb, str(tmp_path)), nprocs=world, join=True)
    c = cases.small_problem(11, 6, 4, 5, nb, 0.4, warm=True)
    whole = cases.ref_solve(fref, c, 5, 0.01, nthreads=1)
    z = np.concatenate([np.load(tmp_path / f"z_{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(z, whole["z"])            # bit-identical: no cross-instance coupling anywhere
    tmax, iters, count = np.load(tmp_path / "stats.npy")
    assert tmax == 20.0 and count == nb and iters == whole["iters"].sum()
