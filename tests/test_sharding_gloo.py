"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch, regenerate their own slice of
the synthetic workload and gather the per-filter statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ekf_slam_b200.sharding as sharding
import ekf_slam_b200.synth as synth
from ekf_slam_b200._lib import STATS_FIELDS


def test_partition_covers_batch_disjointly():
    for n, w in [(4096, 8), (10, 3), (7, 8), (1, 1), (1024, 2)]:
        seen = []
        for r in range(w):
            b0, nb = sharding.partition(n, w, r)
            seen += list(range(b0, b0 + nb))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.partition(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b0, nb = sharding.partition(B, world, rank)
        seq = synth.SynthSequence(B=nb, N=5, T=2, seed=31, b_offset=b0)
        # stand-in statistics: deterministic functions of the global filter index and the shard's data
        stats = {k: (np.arange(b0, b0 + nb) * (i + 1)).astype(np.int32) for i, k in enumerate(STATS_FIELDS)}
        stats["n_ic"] = seq.has[1].sum(axis=1).astype(np.int32)
        allst = sharding.gather_stats(stats, n_filters_total=B)
        tmax = sharding.max_over_ranks(10.0 + rank)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), tmax=tmax, zc=seq.zc, **allst)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    B, world = 7, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    full = synth.SynthSequence(B=B, N=5, T=2, seed=31)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    # both ranks hold the same gathered statistics, in global filter order
    for k in STATS_FIELDS:
        assert np.array_equal(r0[k], r1[k])
    assert np.array_equal(r0["n_ic"], full.has[1].sum(axis=1))
    assert np.array_equal(r0["ransac_iters"], np.arange(B) * 2)
    assert r0["tmax"] == 11.0 and r1["tmax"] == 11.0
    # each rank regenerated exactly its slice of the global workload
    assert np.array_equal(np.concatenate([r0["zc"], r1["zc"]], axis=1), full.zc)
