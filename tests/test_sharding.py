"""N>1 path on CPU: world_size-2 gloo run of the volume-sharding driver (no GPU, the per-volume
worker is the oracle) — placement covers every volume exactly once, timing is the max over ranks,
and the checksum-of-checksums equals the single-process one."""
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


def _oracle_worker(v, seed):
    from oracle import pyoracle as po
    dat = po.synth(0, 3 * 100 * 10 + 57, seed)
    shards = po.encode_dat_image(dat, buffer_size=50, large=1000, small=100)
    d = 0
    for s in shards[10:]:
        d = (d * 1000003 + int(np.frombuffer(s.tobytes(), dtype=np.uint8).astype(np.uint64).sum())) & ((1 << 64) - 1)
    return d, 1.0 + (v % 3)


def _rank_main(rank, world, port, n_volumes, q):
    import torch.distributed as dist
    from seaweedfs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, sharding.run_batch(n_volumes, _oracle_worker, dist=dist)))
    finally:
        dist.destroy_process_group()


def test_round_robin_placement():
    from seaweedfs_b200 import sharding
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in sharding.volumes_for_rank(256, world, r))
        assert seen == list(range(256))
        assert all(len(sharding.volumes_for_rank(256, world, r)) == 256 // world for r in range(world))
    assert sharding.volumes_for_rank(5, 2, 1) == [1, 3]
    with pytest.raises(ValueError):
        sharding.volumes_for_rank(4, 2, 2)


def test_two_rank_gloo_batch():
    import torch.multiprocessing as mp
    from seaweedfs_b200 import sharding
    n_volumes = 7
    single = sharding.run_batch(n_volumes, _oracle_worker)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, n_volumes, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0]["digest"] == results[1]["digest"] == single["digest"]
    # rank 0 holds volumes 0,2,4,6 → 1+3+2+1 = 7 ms ; rank 1 holds 1,3,5 → 2+1+3 = 6 ms
    assert results[0]["per_rank_ms"] == [7.0, 6.0] and results[0]["ms_max"] == 7.0
    assert single["ms_max"] == 13.0
