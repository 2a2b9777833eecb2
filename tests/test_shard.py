"""CPU, world_size 2 over gloo: the multi-GPU frontend's host logic -- contiguous frame shards with no data-path
collective, and the single weight broadcast from rank 0 (ranks > 0 start from zero weights)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ffcnn_b200 as fb
from ffcnn_b200 import shard


def test_shard_ranges_cover_without_overlap():
    for n in (0, 1, 7, 256, 2048):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def _worker(rank, world, port, cfg, wts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = fb.Net(cfg, wts if rank == 0 else None, 0, 0, device=None)        # only rank 0 reads the weights file
    before = float(np.abs(net.packed_weights()).sum())
    nbytes = shard.broadcast_weights(net, dist)
    after = net.packed_weights()
    lo, hi = shard.shard_range(10, rank, world)
    counts = shard.gather_box_counts(list(range(lo, hi)), dist)
    q.put((rank, before, nbytes, float(after.astype(np.float64).sum()), counts))
    net.close()
    dist.destroy_process_group()


def test_weight_broadcast_and_gather_world2(assets):
    cfg, wts, _ = assets
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, wts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, n0, s0, c0), (r1, b1, n1, s1, c1) = res
    assert b0 > 0 and b1 == 0.0                     # rank 1 really started from zero weights
    assert n0 == n1 == 356576 * 4                   # one broadcast of the packed buffer, nothing else
    assert s0 == s1 != 0.0
    assert c0 == c1 == list(range(10))              # frame order restored across shards


def test_bind_near_gpu_is_best_effort():
    """Without NVML (this container) or when the GPU-local CPU set is everything the process may use, the helper leaves the
    affinity mask alone and says so."""
    import os
    from ffcnn_b200 import shard
    before = os.sched_getaffinity(0)
    got = shard.bind_near_gpu(0)
    try:
        assert got is None or (set(got) < before and len(got) >= 4)
        assert os.sched_getaffinity(0) == (before if got is None else set(got))
    finally:
        os.sched_setaffinity(0, before)


def test_weighted_shards_properties():
    """Rate-proportional contiguous shards (the end-to-end frontend of a box whose GPUs do not share the host path equally):
    contiguous cover of [0, n), multiples of the quantum, the cap respected, equal rates = the equal shards of shard_range,
    faster ranks never get fewer frames, bad arguments refused."""
    import random
    from ffcnn_b200.shard import shard_range, weighted_shards
    assert weighted_shards(2048, [1.0] * 8) == [shard_range(2048, r, 8) for r in range(8)]
    got = weighted_shards(2048, [23.4] * 4 + [35.7] * 4, quantum=8, max_per_rank=384)
    assert [hi - lo for lo, hi in got] == [200] * 4 + [312] * 4
    rng = random.Random(7)
    for _ in range(200):
        world = rng.randint(1, 8)
        q = rng.choice([1, 4, 8])
        n = rng.randint(0, 4096)
        rates = [rng.uniform(0.5, 4.0) for _ in range(world)]
        cap = rng.choice([None, (n + world - 1) // world + 8 * q + q])
        sh = weighted_shards(n, rates, quantum=q, max_per_rank=cap)
        assert sh[0][0] == 0 and sh[-1][1] == n and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        sizes = [hi - lo for lo, hi in sh]
        assert all(s % q == 0 for s in sizes[:-1]) and (sizes[-1] - n % q) % q == 0
        if cap is not None:
            assert max(sizes) <= cap
        else:                                            # without a cap a rank is within one quantum (+ the tail) of its exact share
            tot = sum(rates)
            assert all(abs(s - n * r / tot) <= q + n % q for s, r in zip(sizes, rates))
    for bad in ((10, [], 8), (10, [1.0, 0.0], 8), (-1, [1.0], 8), (10, [1.0], 0)):
        with pytest.raises(ValueError):
            weighted_shards(bad[0], bad[1], quantum=bad[2])
    with pytest.raises(ValueError):
        weighted_shards(100, [1.0, 1.0], quantum=1, max_per_rank=40)
