"""CPU (-m "not gpu"): the multi-GPU partition on world_size-2 `gloo`.

Each rank traces its contiguous ray-index shard and the fixed-point flux words
and counters are summed with ONE all-reduce -- the data path bench.py --gpus N
runs over NCCL.  On this CPU box the per-shard tracer is the oracle (the test
is about the host-side partition / reduction logic, not the kernels); the
`-m gpu` suite checks the same shard sums against the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from viennaray_b200 import distributed as vd


def test_shard_bounds_tile_the_range():
    for num in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            b = [vd.shard_bounds(num, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == num
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        vd.shard_bounds(10, 2, 2)


def test_step_shards_are_disjoint():
    seen = set()
    for step in range(3):
        for rank in range(4):
            s, e = vd.step_shard(step, 100, rank, 4)
            assert e - s == 100 and not (seen & set(range(s, e)))
            seen |= set(range(s, e))
    assert seen == set(range(1200))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_rays, seed, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from tests import common
    c = common.case("plane")
    c["sticking"] = 0.3  # fractional weights: order-independent only in fixed point
    orc = common.make_oracle(c)
    cfg = orc.config(num_rays, seed)
    begin, end = vd.shard_bounds(num_rays, rank, world)
    flux, info = orc.trace(common.oracle_particle(c), cfg, begin, end)
    d = info.as_dict()
    counters = [d[k] for k in ("numRays", "totalTraces", "nonGeoHits", "geoHits", "particleHits",
                               "boundaryHits", "reflections", "raysTerminated")]
    words = torch.from_numpy(np.concatenate([flux, np.asarray(counters, np.uint64)])
                             .view(np.int64).copy())
    vd.all_reduce_words(words)
    if rank == 0:
        np.save(out_path, words.numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_job(tmp_path):
    from oracle import pyoracle as po
    from tests import common
    num, seed, world = 20000, 4242, 2
    out = str(tmp_path / "words.npy")
    mp.spawn(_worker, args=(world, _free_port(), num, seed, out), nprocs=world, join=True)
    words = np.load(out)
    c = common.case("plane")
    c["sticking"] = 0.3
    orc = common.make_oracle(c)
    flux, info = orc.trace(common.oracle_particle(c), orc.config(num, seed))
    assert (words[:orc.n] == flux).all()
    d = info.as_dict()
    assert words[orc.n + 1] == d["totalTraces"] and words[orc.n + 3] == d["geoHits"]
    assert words[orc.n] == num  # numRays summed over the shards
