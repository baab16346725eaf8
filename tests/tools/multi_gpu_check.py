"""torchrun --nproc-per-node N tests/tools/multi_gpu_check.py
Every rank traces its contiguous ray-index shard of ONE job on its own GPU,
the fixed-point flux words + counters are summed with one NCCL all-reduce, and
rank 0 checks the result bit-for-bit against the CPU oracle's whole-job flux
(tests/ infrastructure: this script is a test, not a product path)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import common  # noqa: E402
from viennaray_b200 import capi, host  # noqa: E402
from viennaray_b200 import distributed as vd  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in
                      (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
num, seed = 400_000, 777
ok = True
for name in ("trench", "trench_ion", "triangle3D"):
    c = common.case(name)
    ctx, src, _ = common.make_gpu(c, device=local)
    begin, end = vd.shard_bounds(num, rank, world)
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(num, seed, begin, end), sync=True)
    ptr, words = ctx.flux_device()
    t = vd.as_int64_tensor(ptr, words, torch.device("cuda", local))
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    with torch.cuda.stream(stream):
        vd.all_reduce_words(t)
    ctx.synchronize()
    flux = ctx.flux_download_fixed()[0]
    _, info = ctx.flux_download()
    if rank == 0:
        orc = common.make_oracle(c)
        fo, io = orc.trace(common.oracle_particle(c), orc.config(num, seed))
        same = bool((flux == fo).all()) and info[0].totalRaysTraced == io.totalTraces and \
            info[0].geometryHits == io.geoHits
        print("%-12s world %d: all-reduced flux %s the oracle's whole-job flux (%d traces)" %
              (name, world, "==" if same else "!=", info[0].totalRaysTraced), flush=True)
        ok = ok and same
    ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
