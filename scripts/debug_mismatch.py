import sys, numpy as np
sys.path.insert(0, '.')
from tests import common
from tests.test_gpu_parity import _first_bounce_rays, SEED
name = sys.argv[1] if len(sys.argv) > 1 else "trench_ion"
c = common.case(name); orc = common.make_oracle(c); ctx, src, st = common.make_gpu(c)
m = 200000
rays = orc.source_rays(common.oracle_particle(c), orc.config(10**7, SEED), 0, m)
np.set_printoptions(precision=9, suppress=False)
for leg in range(2):
    go, po_, to, _ = orc.intersect(rays)
    gg, pg, tg, cnt_g, nb_g = ctx.debug_intersect(rays, nb_cap=24)
    bad = np.nonzero((go != gg) | (po_ != pg))[0]
    print("leg", leg, "bad", len(bad))
    for i in bad[:10]:
        print(i, "ray", rays[i].tolist(), "oracle", go[i], po_[i], to[i], "gpu", gg[i], pg[i], tg[i])
        if go[i] == 1:
            p = c["points"][po_[i]]; print("   oracle prim at", p, c["normals"][po_[i]])
        if gg[i] == 1:
            p = c["points"][pg[i]]; print("   gpu prim at", p, c["normals"][pg[i]])
    keep = go == 1
    rays = _first_bounce_rays(orc, rays[keep], po_[keep], to[keep], c)
print("bbox", orc.bbox().tolist())
