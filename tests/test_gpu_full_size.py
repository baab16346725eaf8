"""-m gpu: BASELINE.json's full-size configs -- C4 (1M-disk trench, neutral + ion)
and C5 (4M-disk hole array, power-cosine source, reflective boundaries).

At these sizes the oracle still traces a few hundred thousand rays in seconds,
so a seeded sample is compared bit-for-bit; the 2e7-ray runs are checked through
size-independent properties: the ray-index shards of a job add up to the whole
job exactly (what the multi-GPU all-reduce relies on), equal seeds give equal
words, and the sky map (rays finished without a traversal) changes nothing."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import host

pytestmark = pytest.mark.gpu
SEED = 4242


def _fixed(ctx, src, part, cfg):
    ctx.trace_device(src, [part], cfg, sync=True)
    return ctx.flux_download_fixed()[0], ctx.flux_download()[1][0]


COUNTERS = ("totalRaysTraced", "nonGeometryHits", "geometryHits", "boundaryHits", "reflections",
            "raysTerminated")


@pytest.mark.parametrize("name", ["trench_full", "trench_full_ion", "holes_full"])
def test_full_size_config(name):
    c = common.case(name if name != "trench_full_ion" else "trench_full")
    if name == "trench_full_ion":
        c.update(kind=2, sticking=0.5, power=100.0, cone=float(np.deg2rad(85.0)))
    ctx, src, st = common.make_gpu(c)
    part = common.gpu_particle(c)
    n = len(c["points"])
    assert n >= 999_999

    # 1. seeded sample against the oracle, bit-exact
    orc = common.make_oracle(c)
    num = 300_000
    fo, io = orc.trace(common.oracle_particle(c), orc.config(num, SEED))
    fg, ig = _fixed(ctx, src, part, host.config(num, SEED))
    assert (fg == fo).all(), "%d of %d primitives differ" % (int((fg != fo).sum()), n)
    d = io.as_dict()
    assert (ig.totalRaysTraced, ig.geometryHits, ig.nonGeometryHits, ig.boundaryHits) == \
        (d["totalTraces"], d["geoHits"], d["nonGeoHits"], d["boundaryHits"])
    del orc

    # 2. properties at 2e7 rays
    big = 20_000_000
    whole, iw = _fixed(ctx, src, part, host.config(big, SEED))
    again, _ = _fixed(ctx, src, part, host.config(big, SEED))
    assert (whole == again).all()  # determinism, bitwise (tests/rngSeed)
    parts = np.zeros_like(whole)
    sums = dict.fromkeys(COUNTERS, 0)
    for b, e in ((0, 7_000_001), (7_000_001, big)):  # ragged shards
        f, i = _fixed(ctx, src, part, host.config(big, SEED, b, e))
        parts += f
        for k in COUNTERS:
            sums[k] += getattr(i, k)
    assert (parts == whole).all()
    assert all(sums[k] == getattr(iw, k) for k in COUNTERS)
    assert iw.numRays == big and iw.totalRaysTraced > big
    # every weight unit that lands is counted at least once: flux is non-trivial everywhere
    # the source can see
    assert (whole > 0).mean() > 0.9
    ctx.close()

    # 3. the sky map only skips traversals, never changes a result
    os.environ["VR_SKY_CELLS"] = "0"
    try:
        ctx2, src2, _ = common.make_gpu(c)
    finally:
        os.environ.pop("VR_SKY_CELLS")
    plain, ip = _fixed(ctx2, src2, part, host.config(big, SEED))
    ctx2.close()
    assert (plain == whole).all()
    assert all(getattr(ip, k) == getattr(iw, k) for k in COUNTERS)
