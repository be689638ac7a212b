"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous shards cover the batch exactly once,
per-shard synthetic inputs are independent of the world size, and the stats reduction (the ONLY
collective of the path) sums / maxes correctly."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    from biped_mpc_py_b200 import synth, MPC, Biped
    from biped_mpc_py_b200.shard import shard_slice, local_stats, reduce_stats
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mpc, biped = MPC(), Biped()
        full = synth.make_batch(n_total, shard_index=0, mpc=mpc, biped=biped)   # strong scaling: one batch, sliced
        sl = shard_slice(n_total, rank, world)
        mine = {k: v[sl] for k, v in full.items()}
        # stand-in solver outcome (host only): deterministic functions of the inputs
        iters = 8 + (np.abs(mine["x_fb"][:, 0]) * 10).astype(np.int32)
        status = (mine["gait"] == 0).astype(np.int32) * 0
        status[::7] = 1
        resid = np.stack([np.abs(mine["x_fb"][:, 1]), np.abs(mine["x_fb"][:, 2])], axis=1)
        out = reduce_stats(*local_stats(status, iters, resid))
        q.put((rank, sl.start, sl.stop, out, float(mine["x_fb"].sum())))
    finally:
        dist.destroy_process_group()


def test_shard_slices_partition():
    from biped_mpc_py_b200.shard import shard_slice
    for n in (0, 1, 7, 4096, 262144, 262145):
        for world in (1, 2, 3, 8):
            got = np.concatenate([np.arange(n)[shard_slice(n, r, world)] for r in range(world)])
            assert np.array_equal(got, np.arange(n))
            sizes = [shard_slice(n, r, world).stop - shard_slice(n, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_slice(10, 2, 2)


def test_weak_scaling_shards_are_independent_streams():
    """bench.py's weak-scaling shards use rng(SEED + rank): distinct, reproducible, world-size independent."""
    from biped_mpc_py_b200 import synth
    a0, a1 = synth.make_batch(64, shard_index=0), synth.make_batch(64, shard_index=1)
    assert not np.array_equal(a0["x_fb"], a1["x_fb"])
    assert np.array_equal(a0["x_fb"], synth.make_batch(64, shard_index=0)["x_fb"])


def test_stats_reduction_world2_gloo():
    import torch.multiprocessing as mp
    from biped_mpc_py_b200 import synth, MPC, Biped
    from biped_mpc_py_b200.shard import local_stats
    n_total, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # slices are contiguous and cover the batch once
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_total
    # every rank sees the same reduced stats, equal to the single-process stats of the whole batch
    full = synth.make_batch(n_total, shard_index=0, mpc=MPC(), biped=Biped())
    iters = 8 + (np.abs(full["x_fb"][:, 0]) * 10).astype(np.int32)
    status = np.zeros(n_total, np.int32)
    for r in range(world):
        s = np.zeros(res[r][2] - res[r][1], np.int32)
        s[::7] = 1
        status[res[r][1]:res[r][2]] = s
    resid = np.stack([np.abs(full["x_fb"][:, 1]), np.abs(full["x_fb"][:, 2])], axis=1)
    ssum, smax = local_stats(status, iters, resid)
    for r in range(world):
        out = res[r][3]
        assert out["instances"] == n_total
        assert out["iters_sum"] == ssum[1] and out["not_optimal"] == ssum[2]
        assert out["iters_max"] == smax[0] and out["mu_max"] == smax[1] and out["rd_max"] == smax[2]
    assert abs(res[0][4] + res[1][4] - float(full["x_fb"].sum())) < 1e-9
