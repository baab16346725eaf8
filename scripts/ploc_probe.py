import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
from viennaray_b200 import host, scenes

def morton(c, lo, inv):
    q = np.clip(((c - lo) * inv * 2097152.0), 0, 2097151).astype(np.uint64)
    def spread(v):
        v = v & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return (spread(q[:, 0]) << np.uint64(2)) | (spread(q[:, 1]) << np.uint64(1)) | spread(q[:, 2])

def area(lo, hi):
    d = hi - lo
    return 2 * (d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0])

def ploc(plo, phi, order, R=8):
    n = len(order)
    lo = plo[order].astype(np.float64); hi = phi[order].astype(np.float64)
    cnt = np.ones(n, np.int64)
    # SAH accumulators with leaf collapse at <= 4
    inner = 0.0; leaf = 0.0
    a = area(lo, hi)
    it = 0
    while len(cnt) > 1:
        m = len(cnt)
        best = np.full(m, np.inf); nn = np.full(m, -1, np.int64)
        for d in range(1, min(R, m - 1) + 1):
            ulo = np.minimum(lo[:-d], lo[d:]); uhi = np.maximum(hi[:-d], hi[d:])
            ua = area(ulo, uhi)
            # candidate for i (partner i+d) and for i+d (partner i)
            b = ua < best[:-d]
            idx = np.nonzero(b)[0]; best[idx] = ua[idx]; nn[idx] = idx + d
            b = ua < best[d:]
            idx = np.nonzero(b)[0]; best[idx + d] = ua[idx]; nn[idx + d] = idx
        i = np.arange(m)
        mutual = (nn[nn] == i) & (i < nn)
        li = i[mutual]; ri = nn[mutual]
        nlo = np.minimum(lo[li], lo[ri]); nhi = np.maximum(hi[li], hi[ri])
        ncnt = cnt[li] + cnt[ri]
        na = area(nlo, nhi)
        # accounting: merged node is inner if ncnt > 4; its children with cnt<=4 are leaves
        big = ncnt > 4
        inner += na[big].sum()
        for side in (li, ri):
            c = cnt[side][big]; aa = a[side][big]
            l = c <= 4
            leaf += (aa[l] * c[l]).sum()
        lo[li] = nlo; hi[li] = nhi; cnt[li] = ncnt; a[li] = na
        keep = np.ones(m, bool); keep[ri] = False
        lo = lo[keep]; hi = hi[keep]; cnt = cnt[keep]; a = a[keep]
        it += 1
    root = a[0]
    return inner / root, leaf / root, it

for name, gen in (("C4", scenes.trench), ("C5", scenes.hole_array)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]: continue
    points, normals, gd = gen()
    r = host.disk_radius(gd, 3)
    ext = r * np.sqrt(np.maximum(1 - normals.astype(np.float64) ** 2, 0))
    plo = points - ext; phi = points + ext
    slo = plo.min(0); shi = phi.max(0)
    for alpha in (1.0, 0.0):
        e = shi - slo
        inv = 1.0 / (e ** alpha * e.max() ** (1 - alpha))
        key = morton(0.5 * (plo + phi), slo, inv)
        order = np.argsort(key, kind='stable')
        for R in (8, 16):
            t = time.time()
            si, sl, it = ploc(plo, phi, order, R)
            print(name, "alpha", alpha, "R", R, "PLOC sah_inner %.2f sah_leaf %.2f cost %.2f iters %d (%.1fs)" % (si, sl, si + 0.4 * sl, it, time.time() - t), flush=True)
