// The hot path as a wavefront over a resident ray pool:
//
//   traverseKernel  persistent warps; every lane owns one ray at a time and
//                   replaces it from the slot cursor as soon as it finishes
//                   (warp-level compaction of the work), while-while traversal
//                   of the BVH with disk / triangle tests; boundary box first
//                   so its hit bounds the traversal.  Writes one hit per slot.
//   shadeKernel     one thread per slot: boundary handling, backface rule,
//                   neighbour spread, flux accumulation, particle reflection,
//                   Russian roulette (rayTraceKernel.hpp:172-333) and in-place
//                   regeneration of finished rays from the source
//                   (rayTraceKernel.hpp:120-143).
//
//   spreadKernel    optional (scenes larger than L2): the neighbour spread of the queued
//                   geometry hits as its own pass with every lane busy.
//
//   tailKernel      once the source is dry and few rays survive: one thread per
//                   ray runs traverse + shade in a loop to the ray's end, in one
//                   launch (replaces ~1000 iterations of tiny launches).
//
// The host alternates the two until few slots are alive.  Results are identical
// to running each ray to completion on its own (per-ray counter RNG, fixed
// point flux sums), which is what the CPU oracle does.
#include "vr_device.cuh"
#include <mutex>

namespace vr {

// Traversal stack.  The pushes carry no bounds check: the build measures the depth of the
// radix tree (vr_bvh.cu, depthKernel) and refuses a tree a traversal could overflow on.  A
// binary descent pushes at most one entry per level and the tree over 63-bit Morton keys,
// ties broken by the 32-bit index, is at most 95 levels deep, so VR_STACK entries always
// do; the optional 4-wide variant pushes up to three per (double) level and is only
// emitted when 3 * ceil(depth / 2) fits.
#ifndef VR_NODE_MIN
#define VR_NODE_MIN 1  // lanes at inner nodes needed to keep the warp in the node loop
#endif
#ifndef VR_REFILL_MIN
#define VR_REFILL_MIN 1  // idle lanes a warp waits for before it fetches new slots
#endif
#ifndef VR_TRAV_BLOCKS
#define VR_TRAV_BLOCKS 10  // resident blocks per SM asked of ptxas (caps registers at 48)
#endif
#ifndef VR_TRAV_THREADS
#define VR_TRAV_THREADS 128  // threads per traverse block (the resident warps per SM stay the same)
#endif
// threads per shade block: 512 for the default instantiation (+1.2 % on C4 over 256; 1024:
// -2 %), 256 for the one that queues every hit for spreadKernel (scenes beyond the L2: C5
// loses 0.8 % with 512)
#ifndef VR_SHADE_THREADS
#define VR_SHADE_THREADS 512
#endif
#ifndef VR_SHADE_THREADS_Q
#define VR_SHADE_THREADS_Q 256
#endif
#ifndef VR_TRAV_BLOCKS_WIDE
#define VR_TRAV_BLOCKS_WIDE 10  // the same for the 4-wide node variant
#endif
// the traverse kernel with the top of the tree in shared memory: fewer, larger blocks share
// one copy of the table (2 x 640 threads = the 40 warps per SM of 10 x 128)
#ifndef VR_TRAV_THREADS_TOP
#define VR_TRAV_THREADS_TOP 640
#endif
#ifndef VR_TRAV_BLOCKS_TOP
#define VR_TRAV_BLOCKS_TOP 2
#endif
// entries of the traversal stack kept in shared memory (0: all of it in local memory)
#ifndef VR_SMEM_STACK
#define VR_SMEM_STACK 0
#endif
#define VR_WDIST_CAP 64  // disks one ray can hit at once (hit disk + its neighbour list)
#ifndef VR_SHADE_BLOCKS_Q
#define VR_SHADE_BLOCKS_Q 4
#endif
#ifndef VR_SHADE_BLOCKS
#define VR_SHADE_BLOCKS 4
#endif

__device__ __forceinline__ bool slotEmpty(const float4 &od0) { return od0.w != od0.w; }

// ---------------------------------------------------------------------------
// boundary box (geomID 0).  The planes are axis aligned, so a cheap plane
// distance selects the candidate triangles; the decision and the reported t
// come from the exact triangle test.
// ---------------------------------------------------------------------------
// Not inlined and not unrolled on purpose: the shade kernel calls it from several
// places, and eight inlined triangle tests per call site pushed that kernel to
// 123 KB of SASS (instruction-fetch stalls were 23 % of its samples).
// (arguments and result by value: a real call that keeps everything in registers)
__device__ __noinline__ Hit boundaryTestGeneric(const DeviceScene &sc, const V3 org, const V3 dir) {
  Hit best;
  best.t = 3.402823466e+38f;
  best.geom = best.prim = best.orig = VR_INVALID_ID;
  // margin of the cheap rectangle check, in length units
  const float ext = fmaxf(fmaxf(sc.bbox[1][0] - sc.bbox[0][0], sc.bbox[1][1] - sc.bbox[0][1]),
                          sc.bbox[1][2] - sc.bbox[0][2]);
  // 1. all lanes together: which of the four planes can the ray hit inside the
  //    box face?  (usually one.)  Their plane distances select the order.  This is a
  //    pre-check with margins, so the plane distance comes from an approximate
  //    reciprocal; planes 0,1 bound the first lateral axis, planes 2,3 the second.
  float tpk[4];
  unsigned cand = 0u;
  const int a0 = sc.firstDir, a1 = sc.secondDir, a2 = 3 - a0 - a1;
  const float o[3] = {comp(org, a0), comp(org, a1), comp(org, a2)};
  const float d[3] = {comp(dir, a0), comp(dir, a1), comp(dir, a2)};
  const float lo[3] = {sc.bbox[0][a0], sc.bbox[0][a1], sc.bbox[0][a2]};
  const float hi[3] = {sc.bbox[1][a0], sc.bbox[1][a1], sc.bbox[1][a2]};
#pragma unroll
  for (int e = 0; e < 2; ++e) {  // e: the plane's axis, f: the other lateral axis
    const int f = 1 - e;
    tpk[2 * e] = tpk[2 * e + 1] = 3.402823466e+38f;
    if (d[e] == 0.f)
      continue;
    const float rd = __fdividef(1.f, d[e]);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      // (a ray that starts on the plane has a zero numerator and is skipped)
      const float tp = ((side ? hi[e] : lo[e]) - o[e]) * rd;
      if (!(tp >= 0.5f * VR_TNEAR && tp <= 3.402823466e+38f))
        continue;
      // the plane's two triangles tile the box face: skip them when the plane
      // point is clearly outside that rectangle
      const float m = 1e-4f * (ext + fabsf(tp));
      const float hf = o[f] + d[f] * tp, hu = o[2] + d[2] * tp;
      if (hf < lo[f] - m || hf > hi[f] + m || hu < lo[2] - m || hu > hi[2] + m)
        continue;
      tpk[2 * e + side] = tp;
      cand |= 1u << (2 * e + side);
    }
  }
  // 2. the exact triangle tests, nearest candidate plane first; the lanes of a warp run
  //    this loop side by side on their own planes instead of idling through a loop over
  //    all four (that loop ran the tests with ~2 active lanes)
  while (cand) {
    int k = 0;
    float tmin = 3.402823466e+38f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((cand >> q & 1u) && tpk[q] < tmin) {
        tmin = tpk[q];
        k = q;
      }
    cand &= ~(1u << k);
    if (!(tmin <= best.t * 1.00001f + 1e-5f))
      break;  // the remaining planes lie behind the hit that was found
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      const int i = 2 * k + j;
      V3 v0 = {sc.btri[i][0][0], sc.btri[i][0][1], sc.btri[i][0][2]};
      V3 v1 = {sc.btri[i][1][0], sc.btri[i][1][1], sc.btri[i][1][2]};
      V3 v2 = {sc.btri[i][2][0], sc.btri[i][2][1], sc.btri[i][2][2]};
      testTri(v0, v1, v2, 0u, (uint32_t)i, (uint32_t)i, org, dir, best, nullptr);
    }
  }
  return best;
}

// The same closest hit with a shortcut for the usual case (lateral axes x and y, sc.bFast).
// A ray that starts inside the box leaves it through one of the two planes it moves
// towards.  When plain plane distances show that (a) neither plane behind the ray is a
// candidate, (b) the nearer exit plane is crossed clearly inside one of its two triangles
// -- or clearly above / below the box, a miss -- the triangle is known and only its t is
// computed, with the very float operations of testTri: for an axis-aligned triangle
// Ng = cross(e2, e1) has one non-zero component N (sc.bN), the other products of
// dot(Ng, C) and dot(Ng, dir) are exact zeros, and t = fl(fl(N C_a) / fl(N d_a)) (a
// division is symmetric in the signs, so the den < 0 branch of testTri gives the same
// bits).  A miss is declared with exactly the candidate rule of boundaryTestGeneric (same
// expressions, same margin m), a hit with four times that margin; a ray inside a margin (corners, edges, the diagonal, a start on or
// behind a plane) takes boundaryTestGeneric.  Same results, about a fifth of the
// instructions (the boundary tests were 38 % of the shade kernel's).
// the shortcut inlined at its call sites (+0.75 % on C4; the generic path stays a call)
#ifndef VR_BOUNDARY_INLINE
#define VR_BOUNDARY_INLINE 1
#endif
#if VR_BOUNDARY_INLINE
__device__ __forceinline__ Hit boundaryTest(const DeviceScene &sc, const V3 org, const V3 dir) {
#else
__device__ __noinline__ Hit boundaryTest(const DeviceScene &sc, const V3 org, const V3 dir) {
#endif
  if (!sc.bFast)
    return boundaryTestGeneric(sc, org, dir);
  const float INF = __int_as_float(0x7f800000);
  const bool px = dir.x > 0.f, py = dir.y > 0.f;
  const float lox = sc.bbox[0][0], hix = sc.bbox[1][0], loy = sc.bbox[0][1], hiy = sc.bbox[1][1];
  const float loz = sc.bbox[0][2], hiz = sc.bbox[1][2];
  const float rdx = __fdividef(1.f, dir.x), rdy = __fdividef(1.f, dir.y);
  // distance to the plane ahead (exit) and to the plane behind (entry) per lateral axis
  float txe = ((px ? hix : lox) - org.x) * rdx, txn = ((px ? lox : hix) - org.x) * rdx;
  float tye = ((py ? hiy : loy) - org.y) * rdy, tyn = ((py ? loy : hiy) - org.y) * rdy;
  if (dir.x == 0.f) {
    txe = INF;
    txn = -INF;
  }
  if (dir.y == 0.f) {
    tye = INF;
    tyn = -INF;
  }
  const bool xFirst = txe < tye;
  const float t0 = xFirst ? txe : tye;
  // (NaN from 0 * inf fails the comparisons and takes the generic path)
  bool clear = txn < 0.25f * VR_TNEAR && tyn < 0.25f * VR_TNEAR && t0 >= 2.f * VR_TNEAR &&
               org.z >= loz && org.z <= hiz;
  Hit best;
  best.t = 3.402823466e+38f;
  best.geom = best.prim = best.orig = VR_INVALID_ID;
  if (clear && t0 == INF)
    return best;  // parallel to all four planes
  const float m = 1e-4f * (sc.bExt + t0);
  const float hu = org.z + dir.z * t0;
  if (clear && (hu < loz - m || hu > hiz + m))
    return best;  // leaves through the top or the bottom: every later plane point is farther out
  // the crossing point on the plane: f along the other lateral axis, u along z
  const float hf = xFirst ? org.y + dir.y * t0 : org.x + dir.x * t0;
  const float lof = xFirst ? loy : lox, hif = xFirst ? hiy : hix;
  // (the triangle test decides inside / outside to within ~1e-7 |C| / |d_a| of an edge: with
  // |d_a| >= 0.01 that is a tenth of the margin 4 m)
  const float mi = 4.f * m;
  clear = clear && hf > lof + mi && hf < hif - mi && hu > loz + mi && hu < hiz - mi &&
          fabsf(xFirst ? dir.x : dir.y) >= 0.01f;
  // side of the diagonal (lof, loz) - (hif, hiz)
  const float Lf = hif - lof, Lz = hiz - loz;
  const float a = (hf - lof) * Lz, b = (hu - loz) * Lf;
  clear = clear && fabsf(a - b) > 1e-3f * (Lf * Lz);
  if (!clear)
    return boundaryTestGeneric(sc, org, dir);
  // x planes: triangles 0, 1 (low), 2, 3 (high), the lower index below the diagonal;
  // y planes: 4, 5 (low), 6, 7 (high), the lower index above it (vr_scene_set_boundary)
  const bool upper = b > a;
  const int i = xFirst ? (px ? 2 : 0) + (upper ? 1 : 0) : (py ? 6 : 4) + (upper ? 0 : 1);
  const float N = sc.bN[i];
  const float T = N * (sc.bX[i] - (xFirst ? org.x : org.y));
  const float den = N * (xFirst ? dir.x : dir.y);
  best.t = T / den;
  best.geom = 0u;
  best.prim = best.orig = (uint32_t)i;
  return best;
}

// closest boundary hit of a fresh ray, stored as the traversal's initial best
__device__ __forceinline__ void storeBoundaryHit(const DeviceScene &sc, const RayPool &pool,
                                                 uint32_t s, const V3 &org, const V3 &dir) {
  const Hit best = boundaryTest(sc, org, dir);
  __stcs(&pool.hit[s],
         make_float4(best.t, __uint_as_float(best.prim), __uint_as_float(best.geom), 0.f));
}

// Sky test: may the ray leaving `org` towards the source be declared free of
// any primitive?  Yes when (a) its cell of the sky map is flat with its
// neighbourhood and the ray is steeper than every sight line from the cell's
// base height to the tops of all farther cells, (b) it starts at or above that
// base, and (c) it rises above the highest primitive before it reaches a
// lateral boundary (tBoundary), so that no periodic / mirrored image matters.
__device__ __forceinline__ bool skyEscapes(const DeviceScene &sc, const V3 &org, const V3 &dir,
                                           float tBoundary) {
  const float up = sc.skySign * comp(dir, sc.skyUp);
  if (!(up > 0.05f))
    return false;
  const float la = comp(dir, sc.skyA), lb = comp(dir, sc.skyB);
  const float lat = sqrtf(la * la + lb * lb);
  const int G = sc.skyN;
  const int ca = min(G - 1, max(0, (int)floorf((comp(org, sc.skyA) - sc.skyLo[0]) * sc.skyInv[0])));
  const int cb = min(G - 1, max(0, (int)floorf((comp(org, sc.skyB) - sc.skyLo[1]) * sc.skyInv[1])));
  const float2 cell = __ldg(&sc.sky[ca * G + cb]);  // {base, slope}
  const float h0 = sc.skySign * comp(org, sc.skyUp);
  if (!(up > (cell.y * 1.001f + 1e-4f) * lat))
    return false;
  if (!(h0 + VR_TNEAR * up > cell.x + 1e-6f * fmaxf(1.f, fabsf(cell.x))))
    return false;
  const float tTop = (sc.skyTop - h0) / up;
  return tTop * 1.0001f + 1e-5f < tBoundary;
}

__device__ __forceinline__ unsigned long long warpSum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Slab test of one quantised child box.  A node word holds the low and the high plane
// of one axis (16 bits each); sel* is the byte-permute selector that picks the plane the
// ray enters through (0x7610: low half, 0x7632: high half, by the sign of the slope) and
// sel ^ 0x22 the one it leaves through.  The permute also supplies the exponent 0x4B00,
// so the word read as a float is 2^23 + q and the conversion costs no instruction of its
// own: t = (2^23 + q) * slope + (offset - 2^23 * slope).  The folded offset is rounded at
// the magnitude 2^23 * slope, i.e. to half a grid cell; the boxes were widened by a full
// cell at build time, so the test stays conservative.
#ifndef VR_HOIST_SEL
#define VR_HOIST_SEL 0  // 1: keep the far-plane selectors in registers (three more) instead of
                        // deriving them from the near-plane ones at every node
#endif
struct NodeRay {
  float ix, iy, iz, ox, oy, oz;
  uint32_t sx, sy, sz;
#if VR_HOIST_SEL
  uint32_t fx, fy, fz;
#endif
};
__device__ __forceinline__ NodeRay makeNodeRay(const DeviceScene &sc, const V3 &org, const V3 &dir) {
  NodeRay r;
  // reciprocal for the slab tests only; a zero component becomes a huge finite slope so
  // that lo*ix - org*ix keeps the right sign
  // (approximate reciprocal: the slope only steers the conservative box tests)
  r.ix = __fdividef(1.f, fabsf(dir.x) > 1e-20f ? dir.x : copysignf(1e-20f, dir.x));
  r.iy = __fdividef(1.f, fabsf(dir.y) > 1e-20f ? dir.y : copysignf(1e-20f, dir.y));
  r.iz = __fdividef(1.f, fabsf(dir.z) > 1e-20f ? dir.z : copysignf(1e-20f, dir.z));
  // node boxes live on the 16-bit grid: t = q * (scale/d) + (qLo - org)/d
  r.ox = (sc.qLo[0] - org.x) * r.ix;
  r.oy = (sc.qLo[1] - org.y) * r.iy;
  r.oz = (sc.qLo[2] - org.z) * r.iz;
  r.ix *= sc.qScale[0];
  r.iy *= sc.qScale[1];
  r.iz *= sc.qScale[2];
  r.ox = __fmaf_rn(-8388608.f, r.ix, r.ox);
  r.oy = __fmaf_rn(-8388608.f, r.iy, r.oy);
  r.oz = __fmaf_rn(-8388608.f, r.iz, r.oz);
  r.sx = r.ix < 0.f ? 0x7632u : 0x7610u;
  r.sy = r.iy < 0.f ? 0x7632u : 0x7610u;
  r.sz = r.iz < 0.f ? 0x7632u : 0x7610u;
#if VR_HOIST_SEL
  r.fx = r.sx ^ 0x22u;
  r.fy = r.sy ^ 0x22u;
  r.fz = r.sz ^ 0x22u;
#endif
  return r;
}
__device__ __forceinline__ void slabChild(const uint4 c, const NodeRay &r, float tmax, float &n,
                                          float &f) {
  const uint32_t M = 0x4B000000u;
#if VR_HOIST_SEL
  const uint32_t qx = r.fx, qy = r.fy, qz = r.fz;
#else
  const uint32_t qx = r.sx ^ 0x22u, qy = r.sy ^ 0x22u, qz = r.sz ^ 0x22u;
#endif
  const float nx = __fmaf_rn(__uint_as_float(__byte_perm(c.x, M, r.sx)), r.ix, r.ox);
  const float fx = __fmaf_rn(__uint_as_float(__byte_perm(c.x, M, qx)), r.ix, r.ox);
  const float ny = __fmaf_rn(__uint_as_float(__byte_perm(c.y, M, r.sy)), r.iy, r.oy);
  const float fy = __fmaf_rn(__uint_as_float(__byte_perm(c.y, M, qy)), r.iy, r.oy);
  const float nz = __fmaf_rn(__uint_as_float(__byte_perm(c.z, M, r.sz)), r.iz, r.oz);
  const float fz = __fmaf_rn(__uint_as_float(__byte_perm(c.z, M, qz)), r.iz, r.oz);
  n = fmaxf(fmaxf(nx, ny), fmaxf(nz, VR_TNEAR));
  f = fminf(fminf(fx, fy), fminf(fz, tmax));
}

// prefetch variants of the traverse kernel (off unless built with -DVR_PF_*=1)
#ifndef VR_PF_KIDS
#define VR_PF_KIDS 0  // both children of a node as soon as the node has arrived
#endif
#ifndef VR_PF_PUSH
#define VR_PF_PUSH 0  // the deferred child at the time it is pushed
#endif
#ifndef VR_PF_LEAF
#define VR_PF_LEAF 0  // a leaf's disks when a lane leaves the node loop: 1 first, 2 first + last, 4 all
#endif
#ifndef VR_PF_POP
#define VR_PF_POP 0  // the popped entry before the leaf's tests run
#endif
template <int GEO>
__device__ __forceinline__ void prefetchLeaf(const DeviceScene &sc, uint32_t ref) {
  const uint32_t first = (ref & 0x7fffffffu) >> 4;
  constexpr int stride = GEO ? 4 : 2;
  prefetchL1(&sc.prim[(size_t)stride * first]);
#if VR_PF_LEAF == 2
  prefetchL1(&sc.prim[(size_t)stride * (first + (ref & 15u)) - 1]);
#elif VR_PF_LEAF == 4
  for (uint32_t k = 1; k < (ref & 15u); ++k)
    prefetchL1(&sc.prim[(size_t)stride * (first + k)]);
#endif
}
template <int GEO>
__device__ __forceinline__ void prefetchRef(const DeviceScene &sc, uint32_t ref) {
  if (ref < VR_DONE)
    prefetchL1(sc.nodes + ref);
  else
    prefetchLeaf<GEO>(sc, ref);
}

// ---------------------------------------------------------------------------
// closest hit of ONE ray by its own thread (no warp cooperation): the tail
// kernel's traversal.  Same node / primitive tests as traverseKernel; `best`
// comes in as the ray's boundary hit.
// ---------------------------------------------------------------------------
template <int GEO>
__device__ __forceinline__ void traverseOne(const DeviceScene &sc, const V3 &org, const V3 &dir,
                                            Hit &best, unsigned &wNodes, unsigned &wPrims) {
  if (!sc.numPrims)
    return;
  const NodeRay nr = makeNodeRay(sc, org, dir);
  uint32_t stack[VR_STACK];
  int sp = 0;
  uint32_t cur = sc.rootRef;
  while (cur != VR_DONE) {
    if (cur < VR_DONE) {  // inner node
      uint4 c0, c1;
      ldg256(sc.nodes + cur, c0, c1);
      ++wNodes;
      float n0, f0, n1, f1;
      slabChild(c0, nr, best.t, n0, f0);
      slabChild(c1, nr, best.t, n1, f1);
      const bool h0 = n0 * 0.99999f <= f0 * 1.00001f + 1e-6f;
      const bool h1 = n1 * 0.99999f <= f1 * 1.00001f + 1e-6f;
      const uint32_t r0 = c0.w, r1 = c1.w;
      if (h0 && h1) {
        const bool swap = n1 < n0;
        stack[sp++] = swap ? r0 : r1;
        cur = swap ? r1 : r0;
      } else if (h0) {
        cur = r0;
      } else if (h1) {
        cur = r1;
      } else {
        cur = sp ? stack[--sp] : VR_DONE;
      }
    } else {  // leaf
      const uint32_t first = (cur & 0x7fffffffu) >> 4, count = cur & 15u;
      for (uint32_t k = 0; k < count; ++k) {
        const uint32_t i = first + k;
        if (GEO == 0) {
          float4 P, N;
          ldg256(&sc.prim[2 * i], P, N);
          testDisk(P, N, i, org, dir, best);
        } else {
          const float4 a = __ldg(&sc.prim[4 * i]), b = __ldg(&sc.prim[4 * i + 1]),
                       c = __ldg(&sc.prim[4 * i + 2]);
          testTri({a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, 1u, i, __float_as_uint(a.w),
                  org, dir, best, nullptr);
        }
      }
      wPrims += count;
      cur = sp ? stack[--sp] : VR_DONE;
    }
  }
}

// ---------------------------------------------------------------------------
// traverse: closest hit of every live slot
// ---------------------------------------------------------------------------
// one-dimensional bulk copy global -> shared (TMA, cp.async.bulk) completing on an mbarrier
__device__ __forceinline__ void bulkLoadTop(uint4 *dst, const uint4 *src, uint32_t bytes,
                                            unsigned long long *mbar) {
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(mbar);
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], "
                 "%2, [%3];" ::"r"(d),
                 "l"(src), "r"(bytes), "r"(mb)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                 "selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(mb)
                 : "memory");
}

template <int TOP> struct TravShape {
  static constexpr int threads = TOP ? VR_TRAV_THREADS_TOP : VR_TRAV_THREADS;
};
template <int GEO, int WIDE, int COUNT, int TOP>
__global__ void __launch_bounds__(TOP ? VR_TRAV_THREADS_TOP : VR_TRAV_THREADS,
                                  TOP ? VR_TRAV_BLOCKS_TOP
                                      : (WIDE ? VR_TRAV_BLOCKS_WIDE : VR_TRAV_BLOCKS) * 128 /
                                            VR_TRAV_THREADS)
    traverseKernel(const __grid_constant__ TraceParams p) {
  const DeviceScene &sc = p.scene;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned ltMask = (1u << lane) - 1u;
  const uint32_t numSlots = *p.slotCount;
  constexpr int T = TravShape<TOP>::threads;
  extern __shared__ __align__(128) uint4 smTop[];  // TOP: the first sc.topCount entries
#if VR_SMEM_STACK > 0
  __shared__ uint32_t smStack[VR_SMEM_STACK * T];
  uint32_t *const sst = smStack + threadIdx.x;
#endif
  if (TOP) {
    __shared__ unsigned long long mbar;
    bulkLoadTop(smTop, sc.top, sc.topCount * 32u, &mbar);
  }
#if VR_SMEM_STACK > 0
#define VR_PUSH(v)                                                                                 \
  {                                                                                                \
    if (sp < VR_SMEM_STACK)                                                                        \
      sst[sp * T] = (v);                                                                           \
    else                                                                                           \
      stack[sp - VR_SMEM_STACK] = (v);                                                             \
    ++sp;                                                                                          \
  }
#define VR_POP(dst)                                                                                \
  {                                                                                                \
    --sp;                                                                                          \
    dst = sp < VR_SMEM_STACK ? sst[sp * T] : stack[sp - VR_SMEM_STACK];                            \
  }
#else
#define VR_PUSH(v)                                                                                 \
  {                                                                                                \
    stack[sp] = (v);                                                                               \
    ++sp;                                                                                          \
  }
#define VR_POP(dst)                                                                                \
  { dst = stack[--sp]; }
#endif

  uint32_t slot = VR_INVALID_ID;
  V3 org = {0.f, 0.f, 0.f}, dir = {0.f, 0.f, 1.f};
  NodeRay nr = makeNodeRay(sc, org, dir);
  Hit best;
  best.t = 0.f;
  best.geom = best.prim = best.orig = VR_INVALID_ID;
  uint32_t cur = VR_DONE;
  uint32_t stack[VR_STACK];
  int sp = 0;
  bool exhausted = false;
  unsigned wNodes = 0, wPrims = 0;

  for (;;) {
    // ---- replace finished lanes from the slot cursor ------------------------
    const unsigned need = __ballot_sync(0xffffffffu, slot == VR_INVALID_ID);
    if (__popc(need) >= VR_REFILL_MIN && !exhausted) {
      const unsigned nNeed = __popc(need);
      const int leader = __ffs(need) - 1;
      unsigned base = 0;
      if ((int)lane == leader)
        base = atomicAdd(p.slotCursor, nNeed);
      base = __shfl_sync(0xffffffffu, base, leader);
      if (base + nNeed >= numSlots)
        exhausted = true;
      if (slot == VR_INVALID_ID) {
        const unsigned s = base + __popc(need & ltMask);
        if (s < numSlots) {
          const float4 a = __ldcs(&p.pool.od0[s]);
          if (!slotEmpty(a)) {
            const float2 b = __ldcs(&p.pool.od1[s]);
            slot = s;
            org = {a.x, a.y, a.z};
            dir = {a.w, b.x, b.y};
            nr = makeNodeRay(sc, org, dir);
            // the shade / init kernel already intersected the boundary box
            const float4 h0 = __ldcs(&p.pool.hit[s]);
            best.t = h0.x;
            best.prim = best.orig = __float_as_uint(h0.y);
            best.geom = __float_as_uint(h0.z);
            sp = 0;
            cur = sc.numPrims ? (TOP ? VR_TOP_BASE : sc.rootRef) : VR_DONE;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, slot != VR_INVALID_ID)) {
      if (exhausted)
        break;
      continue;
    }

    // ---- inner nodes.  The warp stays in this loop while at least VR_NODE_MIN
    // lanes are at an inner node (or no lane has a leaf to test yet); lanes that
    // reached a leaf wait here, lanes still at nodes wait during the leaf phase.
    for (;;) {
      const bool atNode = cur < VR_DONE;
#if VR_NODE_MIN > 1
      const unsigned nm = __ballot_sync(0xffffffffu, atNode);
      if (nm == 0u)
        break;
      if (__popc(nm) < VR_NODE_MIN && __any_sync(0xffffffffu, cur > VR_DONE))
        break;
#else
      if (!atNode)
        break;
#endif
      if (WIDE && atNode) {
        // 4-wide node: four quantised boxes, hit children sorted near to far
        uint4 c[4];
        ldg256(sc.nodes4 + 4 * (size_t)cur, c[0], c[1]);
        ldg256(sc.nodes4 + 4 * (size_t)cur + 2, c[2], c[3]);
        if (COUNT)
          ++wNodes;
        float key[4];
        uint32_t ref[4];
        int count = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float nn, ff;
          slabChild(c[k], nr, best.t, nn, ff);
          const bool h = nn <= __fmaf_rn(ff, 1.00003f, 2e-6f) && c[k].w != VR_DONE;
          key[k] = h ? nn : 3.402823466e+38f;
          ref[k] = c[k].w;
          count += h ? 1 : 0;
        }
#define VR_CSWAP(a, b)                                                                             \
  {                                                                                                \
    const bool s_ = key[b] < key[a];                                                               \
    const float ka = s_ ? key[b] : key[a], kb = s_ ? key[a] : key[b];                              \
    const uint32_t ra = s_ ? ref[b] : ref[a], rb = s_ ? ref[a] : ref[b];                           \
    key[a] = ka;                                                                                   \
    key[b] = kb;                                                                                   \
    ref[a] = ra;                                                                                   \
    ref[b] = rb;                                                                                   \
  }
        VR_CSWAP(0, 1)
        VR_CSWAP(2, 3)
        VR_CSWAP(0, 2)
        VR_CSWAP(1, 3)
        VR_CSWAP(1, 2)
#undef VR_CSWAP
        if (count == 0) {
          cur = VR_DONE;
          if (sp)
            VR_POP(cur)
        } else {
          cur = ref[0];
          if (count > 3)
            VR_PUSH(ref[3])
          if (count > 2)
            VR_PUSH(ref[2])
          if (count > 1)
            VR_PUSH(ref[1])
        }
      }
      if (!WIDE && atNode) {
        uint4 c0, c1;
        if (TOP && cur >= VR_TOP_BASE) {
          const uint4 *q = smTop + 2u * (cur - VR_TOP_BASE);
          c0 = q[0];
          c1 = q[1];
        } else {
          ldgNode(sc.nodes + cur, c0, c1);
        }
#if VR_PF_KIDS
        // both children's records requested while this node's boxes are tested
        if (c0.w < VR_DONE)
          prefetchL1(sc.nodes + c0.w);
        if (c1.w < VR_DONE)
          prefetchL1(sc.nodes + c1.w);
#endif
        if (COUNT)
          ++wNodes;
        // slab tests with an explicit FMA per plane (slabChild); the boxes were
        // rounded outwards by a full grid cell at build time and the comparison is
        // widened, so rounding here can only add visits
        float n0, f0, n1, f1;
        slabChild(c0, nr, best.t, n0, f0);
        slabChild(c1, nr, best.t, n1, f1);
        const bool h0 = n0 <= __fmaf_rn(f0, 1.00003f, 2e-6f);
        const bool h1 = n1 <= __fmaf_rn(f1, 1.00003f, 2e-6f);
        const uint32_t r0 = c0.w, r1 = c1.w;
        const bool swap = n1 < n0;
        if (h0 && h1) {
          VR_PUSH(swap ? r0 : r1)
#if VR_PF_PUSH
          prefetchRef<GEO>(sc, swap ? r0 : r1);  // in L1 by the time it is popped
#endif
        }
        if (h0 || h1) {
          cur = (h0 && (!h1 || !swap)) ? r0 : r1;
        } else {
          cur = VR_DONE;
          if (sp)
            VR_POP(cur)
        }
#if VR_PF_LEAF
        // a lane that reached a leaf waits for the others: its disks travel meanwhile
        if (cur > VR_DONE)
          prefetchLeaf<GEO>(sc, cur);
#endif
      }
    }

    // ---- one leaf per lane that stands at one -----------------------------------
    if (cur > VR_DONE) {
      const uint32_t first = (cur & 0x7fffffffu) >> 4, count = cur & 15u;
      cur = VR_DONE;
      if (sp)
        VR_POP(cur)
#if VR_PF_POP
      if (cur != VR_DONE)
        prefetchRef<GEO>(sc, cur);  // the next node / leaf travels during this leaf's tests
#endif
      for (uint32_t k = 0; k < count; ++k) {
        const uint32_t i = first + k;
        if (GEO == 0) {
          float4 P, N;
          ldgDisk(&sc.prim[2 * i], P, N);
          testDisk(P, N, i, org, dir, best);
        } else {
          const float4 a = __ldg(&sc.prim[4 * i]), b = __ldg(&sc.prim[4 * i + 1]),
                       c = __ldg(&sc.prim[4 * i + 2]);
          testTri({a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, 1u, i, __float_as_uint(a.w),
                  org, dir, best, nullptr);
        }
      }
      if (COUNT)
        wPrims += count;
    }

    // ---- finished: publish the hit, free the lane ---------------------------------
    if (slot != VR_INVALID_ID && cur == VR_DONE) {
      __stcs(&p.pool.hit[slot], make_float4(best.t, __uint_as_float(best.prim),
                                            __uint_as_float(best.geom), 0.f));
      slot = VR_INVALID_ID;
    }
  }

#undef VR_PUSH
#undef VR_POP
  if (COUNT && p.work) {
    unsigned long long a = warpSum((unsigned long long)wNodes),
                       b = warpSum((unsigned long long)wPrims);
    if (lane == 0) {
      atomicAdd(&p.work[0], a);
      atomicAdd(&p.work[1], b);
    }
  }
}

template <int GEO, int WIDE, int COUNT, int TOP>
static cudaError_t launchTraverseT(const TraceParams &p, int numSMs, cudaStream_t s) {
  constexpr int T = TravShape<TOP>::threads;
  // the table is sized per scene; the launch asks for what this scene's table needs
  const size_t smem = TOP ? (size_t)p.scene.topCount * 32u : 0u;
  // resident blocks per SM of this instantiation (cached; the lanes of a trace and the
  // devices of a multi-device context launch from several host threads)
  static std::mutex mu;
  static int cachedPerSM = 0;
  static size_t smemFor = ~(size_t)0;
  int perSM;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cachedPerSM == 0 || smemFor != smem) {
      cudaError_t e = cudaSuccess;
      int v = 0;
      if (TOP)
        e = cudaFuncSetAttribute(traverseKernel<GEO, WIDE, COUNT, TOP>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(VR_TOP_MAX * 32u));
      if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &v, traverseKernel<GEO, WIDE, COUNT, TOP>, T, smem);
      if (e != cudaSuccess)
        return e;
      cachedPerSM = v < 1 ? 1 : v;
      smemFor = smem;
    }
    perSM = cachedPerSM;
  }
  unsigned want = (p.numSlots + (unsigned)T - 1u) / (unsigned)T;
  unsigned grid = (unsigned)(numSMs * perSM);
  if (want < grid)
    grid = want;
  traverseKernel<GEO, WIDE, COUNT, TOP><<<grid, T, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launchTraverse(const TraceParams &p, int numSMs, cudaStream_t s) {
  if (p.numSlots == 0)
    return cudaSuccess;
  const bool wide = p.scene.nodes4 != nullptr && p.scene.rootRef < VR_DONE;
  // the work counters (VR_COUNT_WORK) are compiled out of the normal kernels
  // the shared-memory top of the tree only pays when there are rays enough to amortise
  // one table copy per block
  const bool top = !wide && p.scene.top != nullptr && p.scene.topCount > 0 &&
                   p.scene.rootRef < VR_DONE && p.numSlots >= 65536u;
  const int which = (p.scene.geoType ? 4 : 0) | (wide ? 2 : 0) | (p.work ? 1 : 0);
  if (top) {
    switch (which) {
    case 0: return launchTraverseT<0, 0, 0, 1>(p, numSMs, s);
    case 1: return launchTraverseT<0, 0, 1, 1>(p, numSMs, s);
    case 4: return launchTraverseT<1, 0, 0, 1>(p, numSMs, s);
    default: return launchTraverseT<1, 0, 1, 1>(p, numSMs, s);
    }
  }
  switch (which) {
  case 0: return launchTraverseT<0, 0, 0, 0>(p, numSMs, s);
  case 1: return launchTraverseT<0, 0, 1, 0>(p, numSMs, s);
  case 2: return launchTraverseT<0, 1, 0, 0>(p, numSMs, s);
  case 3: return launchTraverseT<0, 1, 1, 0>(p, numSMs, s);
  case 4: return launchTraverseT<1, 0, 0, 0>(p, numSMs, s);
  case 5: return launchTraverseT<1, 0, 1, 0>(p, numSMs, s);
  case 6: return launchTraverseT<1, 1, 0, 0>(p, numSMs, s);
  default: return launchTraverseT<1, 1, 1, 0>(p, numSMs, s);
  }
}

// ---------------------------------------------------------------------------
// Boundary::processHit, rayBoundary.hpp:29-127.  Returns `reflect`.
// ---------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ bool boundaryHit(const DeviceScene &sc, V3 &org, V3 &rayDirection,
                                            V3 &dir, uint32_t primID, float t) {
  // unnormalised triangle normal as the intersector reports it
  V3 v0 = {sc.btri[primID][0][0], sc.btri[primID][0][1], sc.btri[primID][0][2]};
  V3 v1 = {sc.btri[primID][1][0], sc.btri[primID][1][1], sc.btri[primID][1][2]};
  V3 v2 = {sc.btri[primID][2][0], sc.btri[primID][2][1], sc.btri[primID][2][2]};
  V3 e1 = {v0.x - v1.x, v0.y - v1.y, v0.z - v1.z};
  V3 e2 = {v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
  V3 ng = cross(e2, e1);
  V3 impact = {org.x + dir.x * t, org.y + dir.y * t, org.z + dir.z * t};
  if (dot(dir, ng) > 0.f) {
    org = impact;
    return true;
  }
  int cond, axis;
  if (D == 2 || primID <= 3) {
    cond = sc.bc[0];
    axis = sc.firstDir;
  } else {
    cond = sc.bc[1];
    axis = sc.secondDir;
  }
  if (cond == VR_BOUNDARY_REFLECTIVE) {
    V3 n = ng;
    normalize(n);
    rayDirection = reflectSpecular(rayDirection, n);
    dir = fillDir<D>(rayDirection);
    org = impact;
    return true;
  }
  if (cond == VR_BOUNDARY_PERIODIC) {
    uint32_t k = primID & 3u;
    setComp(impact, axis, (k <= 1) ? sc.bbox[1][axis] : sc.bbox[0][axis]);
    org = impact;
    return true;
  }
  return false;
}

__device__ __forceinline__ unsigned long long toFixed(float w) {
  return (unsigned long long)(long long)(w * VR_FIXED_SCALE);
}

// fetches the next ray index (warp-aggregated) and samples the source;
// returns false when the job's rays are exhausted.  Must be reached by the
// whole warp.
template <int D>
__device__ __forceinline__ bool regenerate(const TraceParams &p, bool want, uint64_t &idx,
                                           Rng &rng, V3 &org, V3 &rayDirection, V3 &dir) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (!mask)
    return false;
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader)
    base = atomicAdd(p.rayCursor, (unsigned long long)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!want)
    return false;
  const unsigned long long off = base + __popc(mask & ((1u << lane) - 1u));
  if (off >= p.idxEnd - p.idxBegin)
    return false;
  idx = p.idxBegin + off;
  rng.init(p.seed, p.stream, idx);
  if (p.grid)
    sourceSampleGrid<D>(p.src, p.grid, p.gridN, p.eeGrid, idx, rng, org, rayDirection);
  else
    sourceSample<D>(p.src, p.ee, rng, org, rayDirection);
  dir = fillDir<D>(rayDirection);
  return true;
}

template <int D>
__device__ __forceinline__ void storeRay(const RayPool &pool, uint32_t s, const V3 &org,
                                         const V3 &dir, const V3 &rayDirection, float w,
                                         const Rng &rng, uint64_t idx, uint32_t numReflections,
                                         uint32_t boundaryHits, bool hitFromBack) {
  // the pool is streamed (evict-first) so that the scene keeps the L2
  __stcs(&pool.od0[s], make_float4(org.x, org.y, org.z, dir.x));
  __stcs(&pool.od1[s], make_float2(dir.y, dir.z));
  __stcs(&pool.rng[s], rng.save());
  __stcs(&pool.meta[s], make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), numReflections,
                                   boundaryHits | (hitFromBack ? 0x80000000u : 0u)));
  __stcs(&pool.weight[s], w);
  if (D == 2)
    __stcs(&pool.dir3[s], make_float4(rayDirection.x, rayDirection.y, rayDirection.z, 0.f));
}

// ---------------------------------------------------------------------------
// init: fill the pool with the first rays of the shard
// ---------------------------------------------------------------------------
template <int D> __global__ void __launch_bounds__(256) initPoolKernel(const TraceParams p) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool inRange = s < p.numSlots;
  uint64_t idx = 0;
  Rng rng;
  rng.init(0, 0, 0);
  V3 org = {0.f, 0.f, 0.f}, rd = {0.f, 0.f, 0.f}, dir = {0.f, 0.f, 0.f};
  const bool ok = regenerate<D>(p, inRange, idx, rng, org, rd, dir);
  if (!inRange)
    return;
  if (ok) {
    storeRay<D>(p.pool, s, org, dir, rd, 1.f, rng, idx, 0u, 0u, false);
    storeBoundaryHit(p.scene, p.pool, s, org, dir);
  } else {
    p.pool.od0[s] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0x7fc00000u));
  }
}

cudaError_t launchInitPool(const TraceParams &p, cudaStream_t s) {
  if (p.numSlots == 0)
    return cudaSuccess;
  unsigned grid = (p.numSlots + 255u) / 256u;
  if (p.scene.D == 2)
    initPoolKernel<2><<<grid, 256, 0, s>>>(p);
  else
    initPoolKernel<3><<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// shade: rayTraceKernel.hpp:169-333 for the hit of every live slot, then
// regeneration of finished slots
// ---------------------------------------------------------------------------
// The state of one ray while its hit is processed, and the per-thread tallies.
struct RayState {
  V3 org, dir, rayDirection;
  float w;
  uint64_t idx;
  uint32_t numReflections, boundaryHits;
  bool hitFromBack;
  bool rngLoaded;  // rng holds the ray's stream (else it still sits in rs as loaded from the pool)
  bool bhValid;    // bh is the boundary hit of the ray as it stands after the call
  Hit bh;
  Rng rng;
  uint32_t rs;  // the ray's next Philox block as loaded from the pool
};
struct Tally {
  unsigned cTraces, cMiss, cGeo, cBnd, cRefl, cTerm, wNb, wFlux, wSky, cScatter;
};

// Neighbour spread of one geometry hit (rayTraceKernel.hpp:271-280): every neighbour of the
// hit disk that the ray also intersects (checkLocalIntersection) receives the weight.  The
// lists are read from 8-wide rows (sc.nbRow: the first eight neighbours of a disk in one
// 256-bit load, bit 31 of the eighth word = "the CSR holds more"), so the common disk --
// eight neighbours on a flat grid -- needs no offset lookup and no scalar index loads; four
// disks are requested per round so the gathers overlap.  Integer sums: order-free.
// North-star sketch: flux adds aggregated before the global atomics.  Built as an option and
// measured (profiles/r2_experiments.txt): the lanes of a warp almost never hit the same disk
// (16.7M rays in flight over 1M disks, no spatial order), so the match costs more than the
// merged atomics save.
#ifndef VR_FLUX_AGGREGATE
#define VR_FLUX_AGGREGATE 0
#endif
#ifndef VR_PF_ROW
#define VR_PF_ROW 0  // shade kernel: the neighbour row prefetched into L1, 1: beside the normal's load, 2: at the kernel's top
#endif
#ifndef VR_ROW_LD
#define VR_ROW_LD 0  // 1: neighbour rows with L1::no_allocate
#endif
#ifndef VR_N4_LD
#define VR_N4_LD 0  // 1: the hit primitive's normal with L1::no_allocate
#endif
#ifndef VR_SPREAD_LD
#define VR_SPREAD_LD 1  // 1: the neighbour disks of the spread with the traverse kernel's disk
                        // policy (L1::no_allocate: +1.1 % on C4), 0: default policy
#endif
#ifndef VR_NB_ROWS
#define VR_NB_ROWS 1  // 0: the CSR only (A/B switch)
#endif
__device__ __forceinline__ void spreadNeighbors(const TraceParams &p, const uint32_t hprim,
                                                const V3 &org, const V3 &dir,
                                                const unsigned long long wf, unsigned &wNb,
                                                unsigned &wFlux) {
  const DeviceScene &sc = p.scene;
  uint32_t id[4];
  uint32_t k = 0u, k1 = 0u;
#if VR_NB_ROWS
  uint4 ra, rb;
#if VR_ROW_LD
  ldgOnce(sc.nbRow + 2 * (size_t)hprim, ra, rb);
#else
  ldg256(sc.nbRow + 2 * (size_t)hprim, ra, rb);
#endif
  const bool more = rb.w != VR_INVALID_ID && (rb.w >> 31) != 0u;
  if (more)
    rb.w &= 0x7fffffffu;
  id[0] = ra.x, id[1] = ra.y, id[2] = ra.z, id[3] = ra.w;
  int round = 0;
#else
  k = __ldg(&sc.nbOff[hprim]);
  k1 = __ldg(&sc.nbOff[hprim + 1]);
#endif
  // one copy of the round's code for the two halves of the row and the rare CSR rounds
  for (;;) {
#if !VR_NB_ROWS
    if (k >= k1)
      break;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      id[j] = (k + j < k1) ? __ldg(&sc.nbIdx[k + j]) : VR_INVALID_ID;
    k += 4u;
#endif
    float4 P[4], Nn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (id[j] != VR_INVALID_ID) {
#if VR_SPREAD_LD
        ldgDisk(&sc.prim[2 * id[j]], P[j], Nn[j]);
#else
        ldg256(&sc.prim[2 * id[j]], P[j], Nn[j]);
#endif
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (id[j] != VR_INVALID_ID) {
        ++wNb;
        if (checkLocal(P[j], Nn[j], org, dir)) {
          atomicAdd(&p.flux[id[j]], wf);
          ++wFlux;
        }
      }
#if VR_NB_ROWS
    if (++round == 1) {
      id[0] = rb.x, id[1] = rb.y, id[2] = rb.z, id[3] = rb.w;
      continue;
    }
    if (!more)
      break;
    if (round == 2) {
      k = __ldg(&sc.nbOff[hprim]) + 8u;
      k1 = __ldg(&sc.nbOff[hprim + 1]);
    }
    if (k >= k1)
      break;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      id[j] = (k + j < k1) ? __ldg(&sc.nbIdx[k + j]) : VR_INVALID_ID;
    k += 4u;
#endif
  }
}

// What rayTraceKernel.hpp:169-333 does with the hit (t, prim, geom) of a ray: miss,
// boundary handling, back-face rule, neighbour spread, particle functor, roulette, and
// the sky map's shortcut for rays that leave the scene.  Returns true when the ray ended.
// EXT == 1 adds the optional features that are off in the default instantiation:
// mean-free-path scattering (rayTraceKernel.hpp:179-203), the distance-weighted neighbour
// spread of VIENNARAY_USE_WDIST (:258-296) and sticking looked up by the hit primitive's
// materialId (:310-313).
template <int D, int GEO, int EXT, int Q>
__device__ __forceinline__ bool shadeHit(const TraceParams &p, RayState &r, const float ht,
                                         const uint32_t hprim, const uint32_t hgeom, Tally &c) {
  const DeviceScene &sc = p.scene;
  V3 &org = r.org, &dir = r.dir, &rayDirection = r.rayDirection;
  float &w = r.w;
  const uint64_t idx = r.idx;
  uint32_t &numReflections = r.numReflections, &boundaryHits = r.boundaryHits;
  bool &hitFromBack = r.hitFromBack, &rngLoaded = r.rngLoaded, &bhValid = r.bhValid;
  Hit &bh = r.bh;
  Rng &rng = r.rng;
  const uint32_t rs = r.rs;
  unsigned &cTraces = c.cTraces, &cMiss = c.cMiss, &cGeo = c.cGeo, &cBnd = c.cBnd,
           &cRefl = c.cRefl, &cTerm = c.cTerm, &wNb = c.wNb, &wFlux = c.wFlux, &wSky = c.wSky,
           &cScatter = c.cScatter;
  bool finish = false;
  bhValid = false;
  ++cTraces;
  if (rngLoaded)
    rng.discard();  // (tail kernel) this segment's draws start at a fresh block
  bool scattered = false;
  if (EXT && hgeom != VR_INVALID_ID && p.particle.meanFreePath > 0.f) {  // :179-203
    if (!rngLoaded) {
      rng.load(rs, p.seed, p.stream, idx);
      rngLoaded = true;
    }
    const float scatterProbability =
        1.f - exp2det((-ht / p.particle.meanFreePath) * 1.4426950216293335f);
    const float rnd = rng.f();
    if (rnd < scatterProbability) {
      org = {org.x + dir.x * rnd, org.y + dir.y * rnd, org.z + dir.z * rnd};  // sic, :188-190
      float x, y, s2;  // pickRandomPointOnUnitSphere, rayUtil.hpp:266-283
      do {
        x = 2.f * rng.f() - 1.f;
        y = 2.f * rng.f() - 1.f;
        s2 = x * x + y * y;
      } while (s2 >= 1.f);
      const float tmp = 2.f * sqrtf(1.f - s2);
      rayDirection = {x * tmp, y * tmp, 1.f - 2.f * s2};
      dir = fillDir<D>(rayDirection);
      ++cScatter;
      scattered = true;
    }
  }
  if (scattered) {
    // the ray goes on from the scatter point
  } else if (hgeom == VR_INVALID_ID) {  // :172
    ++cMiss;
    finish = true;
  } else if (hgeom == 0u) {  // :206-214
    if (++boundaryHits > p.maxBoundaryHits) {
      ++cTerm;
      finish = true;
    } else if (!boundaryHit<D>(sc, org, rayDirection, dir, hprim, ht)) {
      finish = true;
    }
  } else {
    const V3 hitPoint = {org.x + dir.x * ht, org.y + dir.y * ht, org.z + dir.z * ht};
#if VR_PF_ROW == 1
    if (GEO == 0 && !Q)
      prefetchL1(sc.nbRow + 2 * (size_t)hprim);  // the neighbour row travels during the back-face test
#endif
#if VR_N4_LD
    const float4 N4 = ldgOnce(&sc.prim[GEO == 0 ? 2 * hprim + 1 : 4 * hprim + 3]);
#else
    const float4 N4 = __ldg(&sc.prim[GEO == 0 ? 2 * hprim + 1 : 4 * hprim + 3]);
#endif
    const V3 gn = {N4.x, N4.y, N4.z};
    const bool backface = dot(rayDirection, gn) > 0.f;  // :224
    if (backface) {
      if (GEO == 0 && !hitFromBack) {  // :226-241 let the ray through once
        hitFromBack = true;
        org = hitPoint;
      } else {
        ++cTerm;
        finish = true;
      }
    } else {
      ++cGeo;
      if (EXT && GEO == 0 && (p.flags & VR_FLAG_WDIST)) {
        // :258-296 every hit disk gets w / d_i / sum(1/d) * numDisksHit, d = distance of
        // the impact point to the disk centre (+ 1e-6)
        uint32_t ids[VR_WDIST_CAP];
        float dist[VR_WDIST_CAP];
        uint32_t nh = 1;
        ids[0] = hprim;
        {
          const float4 P0 = __ldg(&sc.prim[2 * hprim]);
          const float qx = hitPoint.x - P0.x, qy = hitPoint.y - P0.y, qz = hitPoint.z - P0.z;
          dist[0] = sqrtf(dot3(qx, qy, qz, qx, qy, qz)) + 1e-6f;
        }
        const uint32_t k0 = __ldg(&sc.nbOff[hprim]), k1 = __ldg(&sc.nbOff[hprim + 1]);
        for (uint32_t k = k0; k < k1; ++k) {
          const uint32_t id = __ldg(&sc.nbIdx[k]);
          float4 P, Nn;
          ldg256(&sc.prim[2 * id], P, Nn);
          float dd;
          ++wNb;
          if (checkLocalDist(P, Nn, org, dir, dd) && nh < VR_WDIST_CAP) {
            ids[nh] = id;
            dist[nh++] = dd + 1e-6f;
          }
        }
        float invSum = 0.f;
        for (uint32_t k = 0; k < nh; ++k)
          invSum += 1.f / dist[k];
        for (uint32_t k = 0; k < nh; ++k) {
          atomicAdd(&p.flux[ids[k]], toFixed(((w / dist[k]) / invSum) * (float)nh));
          ++wFlux;
        }
      } else if (Q) {
        // :271-306 handed to spreadKernel: {org, weight}, {dir, hit disk}
        const unsigned q = atomicAdd(p.spreadCount, 1u);
        p.spreadQ[2 * (size_t)q] = make_float4(org.x, org.y, org.z, w);
        p.spreadQ[2 * (size_t)q + 1] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(hprim));
      } else {
        const unsigned long long wf = toFixed(w);
#if VR_FLUX_AGGREGATE
        {  // warp-aggregated: the lanes that hit the same disk add once (integer sums)
          const unsigned peers = __match_any_sync(__activemask(), hprim);
          unsigned long long sum = 0ull;
          for (unsigned m = peers; m; m &= m - 1u)
            sum += __shfl_sync(peers, wf, __ffs(m) - 1);
          if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1))
            atomicAdd(&p.flux[hprim], sum);
        }
#else
        atomicAdd(&p.flux[hprim], wf);  // :297-306 surfaceCollision
#endif
        ++wFlux;
        if (GEO == 0)  // :271-280 neighbour spread
          spreadNeighbors(p, hprim, org, dir, wf, wNb, wFlux);
      }
      if (!rngLoaded) {
        rng.load(rs, p.seed, p.stream, idx);
        rngLoaded = true;
      }
      // every reflecting lane of the warp generates its block here, side by side (a block
      // that ends up unused is handed back: Rng::save / discard)
      if (rng.left == 0)
        rng.refill();
      const V3 newDir = surfaceReflection<D>(p.particle, rayDirection, gn, rng);  // :310
      float sticking = p.particle.sticking;
      if (EXT && p.matSticking) {  // sticking by the materialId of the hit primitive, :310-313
        const int m = __ldg(&p.matId[hprim]);
        if (m >= 0 && m < p.numMaterials)
          sticking = __ldg(&p.matSticking[m]);
      }
      w -= w * sticking;  // :316
      if (w <= 0.f) {
        finish = true;
      } else if (++numReflections > p.maxReflections) {
        ++cTerm;
        finish = true;
      } else {
        // :435-460 rejectionControl, thresholds 0.1 / 0.3 of the initial weight
        if (w < 0.1f) {
          const float kill = 1.f - w / 0.3f;
          if (rng.f() < kill)
            finish = true;
          else
            w = 0.3f;
        }
        if (!finish) {
          rayDirection = newDir;
          org = hitPoint;
          dir = fillDir<D>(rayDirection);
        }
      }
      if (D == 3 && !finish && sc.sky != nullptr &&
          !(EXT && p.particle.meanFreePath > 0.f)) {  // the boundary walk would need the scatter draws
        // boundary hit of the reflected ray (needed anyway); if the sky map proves
        // that the ray meets no primitive, walk it through the boundary to its
        // end right here: it never needs a traversal
        bh = boundaryTest(sc, org, dir);
        bhValid = true;
        if (skyEscapes(sc, org, dir, bh.t)) {
          ++wSky;
          for (;;) {
            ++cTraces;
            if (bh.geom == VR_INVALID_ID) {
              ++cMiss;
              finish = true;
              break;
            }
            if (++boundaryHits > p.maxBoundaryHits) {
              ++cTerm;
              finish = true;
              break;
            }
            if (!boundaryHit<D>(sc, org, rayDirection, dir, bh.prim, bh.t)) {
              finish = true;
              break;
            }
            bh = boundaryTest(sc, org, dir);
          }
        }
      }
    }
  }
  if (finish) {
    cBnd += boundaryHits;
    cRefl += numReflections;
  }
  return finish;
}

template <int D, int GEO, int EXT, int Q>
__global__ void __launch_bounds__(Q ? VR_SHADE_THREADS_Q : VR_SHADE_THREADS,
                                  Q ? VR_SHADE_BLOCKS_Q * 256 / VR_SHADE_THREADS_Q
                                    : VR_SHADE_BLOCKS * 256 / VR_SHADE_THREADS)
    shadeKernel(const __grid_constant__ TraceParams p) {
  const DeviceScene &sc = p.scene;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t numSlots = *p.slotCount;

  Tally c = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  bool live = false, finish = false;
  float4 a = make_float4(0.f, 0.f, 0.f, __uint_as_float(0x7fc00000u));
  if (s < numSlots)
    a = __ldcs(&p.pool.od0[s]);
  live = !slotEmpty(a);

  RayState r;
  r.org = {a.x, a.y, a.z};
  r.dir = {0.f, 0.f, 0.f};
  r.rayDirection = {0.f, 0.f, 0.f};
  r.w = 0.f;
  r.idx = 0;
  r.numReflections = r.boundaryHits = 0u;
  r.hitFromBack = r.rngLoaded = r.bhValid = false;
  r.bh.t = 0.f;
  r.bh.geom = r.bh.prim = r.bh.orig = VR_INVALID_ID;
  r.rng.init(0, 0, 0);
  r.rs = 0u;
  V3 &org = r.org, &dir = r.dir, &rayDirection = r.rayDirection;
  float &w = r.w;
  uint64_t &idx = r.idx;
  uint32_t &numReflections = r.numReflections, &boundaryHits = r.boundaryHits;
  bool &hitFromBack = r.hitFromBack, &rngLoaded = r.rngLoaded, &bhValid = r.bhValid;
  Hit &bh = r.bh;
  Rng &rng = r.rng;
  uint32_t &rs = r.rs;

  if (live) {
    const float2 b = __ldcs(&p.pool.od1[s]);
    dir = {a.w, b.x, b.y};
    if (D == 2) {
      const float4 d3 = __ldcs(&p.pool.dir3[s]);
      rayDirection = {d3.x, d3.y, d3.z};
    } else {
      rayDirection = dir;
    }
    const float4 hv = __ldcs(&p.pool.hit[s]);
#if VR_PF_ROW == 2
    if (GEO == 0 && !Q && __float_as_uint(hv.z) == 1u) {
      // the hit disk's record and its neighbour row are requested as soon as the hit is known
      prefetchL1(&sc.prim[2 * (size_t)__float_as_uint(hv.y)]);
      prefetchL1(sc.nbRow + 2 * (size_t)__float_as_uint(hv.y));
    }
#endif
    const uint4 meta = __ldcs(&p.pool.meta[s]);
    idx = (uint64_t)meta.x | ((uint64_t)meta.y << 32);
    numReflections = meta.z;
    boundaryHits = meta.w & 0x7fffffffu;
    hitFromBack = (meta.w >> 31) != 0u;
    w = __ldcs(&p.pool.weight[s]);
    rs = __ldcs(&p.pool.rng[s]);  // issued with the other pool loads, not after the neighbour gathers
    finish = shadeHit<D, GEO, EXT, Q>(p, r, hv.x, __float_as_uint(hv.y), __float_as_uint(hv.z), c);
  }

  // ---- regenerate finished slots; write survivors back (in place, or appended to
  // the other pool when compacting the tail) ----------------------------------------
  const bool regen = regenerate<D>(p, live && finish, idx, rng, org, rayDirection, dir);
  const bool survive = live && (!finish || regen);
  if (!p.compact) {
    if (live) {
      if (!finish) {
        __stcs(&p.pool.od0[s], make_float4(org.x, org.y, org.z, dir.x));
        __stcs(&p.pool.od1[s], make_float2(dir.y, dir.z));
        __stcs(&p.pool.meta[s], make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), numReflections,
                                           boundaryHits | (hitFromBack ? 0x80000000u : 0u)));
        __stcs(&p.pool.weight[s], w);
        if (rngLoaded)  // draws were taken from the stream
          __stcs(&p.pool.rng[s], rng.save());
        if (D == 2)
          __stcs(&p.pool.dir3[s],
                 make_float4(rayDirection.x, rayDirection.y, rayDirection.z, 0.f));
      } else if (regen) {
        storeRay<D>(p.pool, s, org, dir, rayDirection, 1.f, rng, idx, 0u, 0u, false);
      } else {
        p.pool.od0[s] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0x7fc00000u));
      }
      if (survive) {
        if (!finish && bhValid)  // the boundary hit of this ray is known already
          __stcs(&p.pool.hit[s], make_float4(bh.t, __uint_as_float(bh.prim),
                                             __uint_as_float(bh.geom), 0.f));
        else
          storeBoundaryHit(sc, p.pool, s, org, dir);
      }
    }
  } else {
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    if (m) {
      const unsigned ln = threadIdx.x & 31u;
      const int leader = __ffs(m) - 1;
      unsigned base = 0;
      if ((int)ln == leader)
        base = atomicAdd(p.liveCount, (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (survive) {
        const uint32_t dst = base + __popc(m & ((1u << ln) - 1u));
        if (!finish) {
          if (!rngLoaded)
            rng.load(rs, p.seed, p.stream, idx);
          storeRay<D>(p.poolOut, dst, org, dir, rayDirection, w, rng, idx, numReflections,
                      boundaryHits, hitFromBack);
        } else {
          storeRay<D>(p.poolOut, dst, org, dir, rayDirection, 1.f, rng, idx, 0u, 0u, false);
        }
        if (!finish && bhValid)
          __stcs(&p.poolOut.hit[dst], make_float4(bh.t, __uint_as_float(bh.prim),
                                                  __uint_as_float(bh.geom), 0.f));
        else
          storeBoundaryHit(sc, p.poolOut, dst, org, dir);
      }
    }
  }
  const bool stillLive = survive && !p.compact;  // compact mode counted them above

  // ---- counters: warp sums (REDUX), then one global atomic per warp and counter on one
  // of VR_COUNTER_COPIES replicas -- no block barrier, so a block's warps retire
  // independently.  Replica word 0 carries the live count of the in-place mode.
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long *cnt =
      p.counters + (size_t)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) %
                            VR_COUNTER_COPIES) * 8;
  // TraceInfo words: 1 traces, 2 misses, 3 geometry hits, 4 particle (scatter) hits,
  // 5 boundary hits, 6 reflections, 7 terminated
  const unsigned vals[8] = {stillLive ? 1u : 0u, c.cTraces, c.cMiss, c.cGeo, c.cBnd, c.cRefl,
                            c.cTerm, c.cScatter};
  const int wordOf[8] = {0, 1, 2, 3, 5, 6, 7, 4};
#pragma unroll
  for (int k = 0; k < (EXT ? 8 : 7); ++k) {
    const unsigned v = __reduce_add_sync(0xffffffffu, vals[k]);
    if (lane == 0 && v)
      atomicAdd(&cnt[wordOf[k]], (unsigned long long)v);
  }
  if (p.work) {
    const unsigned wv[3] = {c.wNb, c.wFlux, c.wSky};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned v = __reduce_add_sync(0xffffffffu, wv[k]);
      if (lane == 0 && v)
        atomicAdd(&p.work[2 + k], (unsigned long long)v);
    }
  }
}

// Neighbour spread as its own pass (rayTraceKernel.hpp:255-306): one thread per queued
// geometry hit, all lanes busy, the same tests and the same fixed-point adds as the inline
// version (integer sums: the order of the adds does not matter).
#ifndef VR_SPREAD_BLOCKS
#define VR_SPREAD_BLOCKS 3  // resident blocks per SM asked of ptxas (80 registers)
#endif
#ifndef VR_SPREAD_GRID
#define VR_SPREAD_GRID 8  // blocks per SM of the grid-stride launch
#endif
__global__ void __launch_bounds__(256, VR_SPREAD_BLOCKS)
    spreadKernel(const __grid_constant__ TraceParams p) {
  const DeviceScene &sc = p.scene;
  const unsigned n = *p.spreadCount;
  unsigned wNb = 0, wFlux = 0;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    float4 a, b;
    ldg256(p.spreadQ + 2 * (size_t)q, a, b);
    const V3 org = {a.x, a.y, a.z}, dir = {b.x, b.y, b.z};
    const uint32_t hprim = __float_as_uint(b.w);
    const unsigned long long wf = toFixed(a.w);
    atomicAdd(&p.flux[hprim], wf);
    ++wFlux;
    spreadNeighbors(p, hprim, org, dir, wf, wNb, wFlux);
  }
  if (p.work) {
    const unsigned a = __reduce_add_sync(0xffffffffu, wNb), b = __reduce_add_sync(0xffffffffu, wFlux);
    if ((threadIdx.x & 31u) == 0) {
      atomicAdd(&p.work[2], (unsigned long long)a);
      atomicAdd(&p.work[3], (unsigned long long)b);
    }
  }
}

cudaError_t launchSpread(const TraceParams &p, int numSMs, cudaStream_t s) {
  if (!p.spreadQ || p.numSlots == 0 || p.scene.geoType != 0 || p.scene.D != 3 ||
      p.particle.meanFreePath > 0.f || (p.flags & VR_FLAG_WDIST) || p.matSticking)
    return cudaSuccess;
  unsigned grid = (p.numSlots + 255u) / 256u;
  const unsigned cap = (unsigned)numSMs * VR_SPREAD_GRID;
  if (grid > cap)
    grid = cap;
  spreadKernel<<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// The thin tail of a trace: once the source is exhausted and only a few
// thousand rays are left, every remaining ray is run to its end by one thread
// of ONE launch -- traverse and shade in a loop -- instead of ~1000 iterations
// of three tiny launches each.  Same arithmetic, same tallies.
template <int D, int GEO, int EXT>
__global__ void __launch_bounds__(128) tailKernel(const __grid_constant__ TraceParams p) {
  const DeviceScene &sc = p.scene;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t numSlots = *p.slotCount;
  Tally c = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  unsigned wNodes = 0, wPrims = 0;
  float4 a = make_float4(0.f, 0.f, 0.f, __uint_as_float(0x7fc00000u));
  if (s < numSlots)
    a = __ldcs(&p.pool.od0[s]);
  if (!slotEmpty(a)) {
    RayState r;
    const float2 b = __ldcs(&p.pool.od1[s]);
    r.org = {a.x, a.y, a.z};
    r.dir = {a.w, b.x, b.y};
    if (D == 2) {
      const float4 d3 = __ldcs(&p.pool.dir3[s]);
      r.rayDirection = {d3.x, d3.y, d3.z};
    } else {
      r.rayDirection = r.dir;
    }
    const float4 hv = __ldcs(&p.pool.hit[s]);
    const uint4 meta = __ldcs(&p.pool.meta[s]);
    r.idx = (uint64_t)meta.x | ((uint64_t)meta.y << 32);
    r.numReflections = meta.z;
    r.boundaryHits = meta.w & 0x7fffffffu;
    r.hitFromBack = (meta.w >> 31) != 0u;
    r.w = __ldcs(&p.pool.weight[s]);
    r.rs = __ldcs(&p.pool.rng[s]);
    r.rngLoaded = false;
    r.rng.init(0, 0, 0);
    r.bhValid = true;  // the pool holds the ray's boundary hit
    r.bh.t = hv.x;
    r.bh.prim = r.bh.orig = __float_as_uint(hv.y);
    r.bh.geom = __float_as_uint(hv.z);
    for (;;) {
      if (!r.bhValid)
        r.bh = boundaryTest(sc, r.org, r.dir);
      Hit best = r.bh;
      traverseOne<GEO>(sc, r.org, r.dir, best, wNodes, wPrims);
      if (shadeHit<D, GEO, EXT, 0>(p, r, best.t, best.prim, best.geom, c))
        break;
    }
  }

  const unsigned lane = threadIdx.x & 31u;
  unsigned long long *cnt =
      p.counters + (size_t)((blockIdx.x * 4u + (threadIdx.x >> 5)) % VR_COUNTER_COPIES) * 8;
  const unsigned vals[7] = {c.cTraces, c.cMiss, c.cGeo, c.cBnd, c.cRefl, c.cTerm, c.cScatter};
  const int wordOf[7] = {1, 2, 3, 5, 6, 7, 4};
#pragma unroll
  for (int k = 0; k < (EXT ? 7 : 6); ++k) {
    const unsigned v = __reduce_add_sync(0xffffffffu, vals[k]);
    if (lane == 0 && v)
      atomicAdd(&cnt[wordOf[k]], (unsigned long long)v);
  }
  if (p.work) {
    const unsigned wv[5] = {wNodes, wPrims, c.wNb, c.wFlux, c.wSky};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const unsigned v = __reduce_add_sync(0xffffffffu, wv[k]);
      if (lane == 0 && v)
        atomicAdd(&p.work[k], (unsigned long long)v);
    }
  }
}

template <int EXT> static void launchTailExt(const TraceParams &p, unsigned grid, cudaStream_t s) {
  if (p.scene.geoType == 0) {
    if (p.scene.D == 2)
      tailKernel<2, 0, EXT><<<grid, 128, 0, s>>>(p);
    else
      tailKernel<3, 0, EXT><<<grid, 128, 0, s>>>(p);
  } else {
    if (p.scene.D == 2)
      tailKernel<2, 1, EXT><<<grid, 128, 0, s>>>(p);
    else
      tailKernel<3, 1, EXT><<<grid, 128, 0, s>>>(p);
  }
}

cudaError_t launchTail(const TraceParams &p, cudaStream_t s) {
  if (p.numSlots == 0)
    return cudaSuccess;
  const unsigned grid = (p.numSlots + 127u) / 128u;
  if (p.particle.meanFreePath > 0.f || (p.flags & VR_FLAG_WDIST) || p.matSticking)
    launchTailExt<1>(p, grid, s);
  else
    launchTailExt<0>(p, grid, s);
  return cudaGetLastError();
}

// ctrl[0] slot cursor, ctrl[1] append cursor of the compacting mode, ctrl[2] slots in
// use, ctrl[3] rays alive after the iteration (read back by the host).  In-place mode:
// the live count is the sum of word 0 of the counter replicas (cleared here).
__global__ void flipKernel(unsigned int *ctrl, unsigned long long *counters, int compact) {
  unsigned long long v = 0;
  for (int c = threadIdx.x; c < VR_COUNTER_COPIES; c += 32) {
    v += counters[(size_t)c * 8];
    counters[(size_t)c * 8] = 0ull;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0) {
    const unsigned int live = compact ? ctrl[1] : (unsigned int)v;
    ctrl[3] = live;
    ctrl[4] = 0u;  // neighbour-spread queue
    ctrl[0] = 0u;
    if (compact)
      ctrl[2] = live;
    ctrl[1] = 0u;
  }
}
cudaError_t launchFlip(unsigned int *ctrl, unsigned long long *counters, int compact,
                       cudaStream_t s) {
  flipKernel<<<1, 32, 0, s>>>(ctrl, counters, compact);
  return cudaGetLastError();
}

template <int EXT> static void launchShadeExt(const TraceParams &p, cudaStream_t s) {
  const unsigned grid = (p.numSlots + VR_SHADE_THREADS - 1u) / VR_SHADE_THREADS;
  const unsigned gridQ = (p.numSlots + VR_SHADE_THREADS_Q - 1u) / VR_SHADE_THREADS_Q;
  if (p.scene.geoType == 0) {
    if (p.scene.D == 2)
      shadeKernel<2, 0, EXT, 0><<<grid, VR_SHADE_THREADS, 0, s>>>(p);
    else if (!EXT && p.spreadQ)  // neighbour spread queued for spreadKernel
      shadeKernel<3, 0, 0, 1><<<gridQ, VR_SHADE_THREADS_Q, 0, s>>>(p);
    else
      shadeKernel<3, 0, EXT, 0><<<grid, VR_SHADE_THREADS, 0, s>>>(p);
  } else {
    if (p.scene.D == 2)
      shadeKernel<2, 1, EXT, 0><<<grid, VR_SHADE_THREADS, 0, s>>>(p);
    else
      shadeKernel<3, 1, EXT, 0><<<grid, VR_SHADE_THREADS, 0, s>>>(p);
  }
}

cudaError_t launchShade(const TraceParams &p, cudaStream_t s) {
  if (p.numSlots == 0)
    return cudaSuccess;
  if (p.particle.meanFreePath > 0.f || (p.flags & VR_FLAG_WDIST) || p.matSticking)
    launchShadeExt<1>(p, s);
  else
    launchShadeExt<0>(p, s);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// primitive bounds (padded so the float slab tests stay conservative)
// ---------------------------------------------------------------------------
__global__ void diskBoundsKernel(const float4 *xyzr, const float4 *nrm, uint32_t n, float4 *lo,
                                 float4 *hi) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float4 P = xyzr[i], N = nrm[i];
  float nn = dot3(N.x, N.y, N.z, N.x, N.y, N.z);
  float c[3] = {P.x, P.y, P.z}, nv[3] = {N.x, N.y, N.z}, l[3], h[3];
  for (int a = 0; a < 3; ++a) {
    float f = nn > 0.f ? 1.f - nv[a] * nv[a] / nn : 1.f;
    float e = P.w * sqrtf(fmaxf(f, 0.f));
    float pad = 2e-4f * P.w + 2e-6f * fabsf(c[a]);
    l[a] = c[a] - e - pad;
    h[a] = c[a] + e + pad;
  }
  lo[i] = make_float4(l[0], l[1], l[2], 0.f);
  hi[i] = make_float4(h[0], h[1], h[2], 0.f);
}

__global__ void triBoundsKernel(const float4 *v0, const float4 *v1, const float4 *v2, uint32_t n,
                                float4 *lo, float4 *hi) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float4 a = v0[i], b = v1[i], c = v2[i];
  float l[3] = {fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)),
                fminf(a.z, fminf(b.z, c.z))};
  float h[3] = {fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)),
                fmaxf(a.z, fmaxf(b.z, c.z))};
  for (int k = 0; k < 3; ++k) {
    float pad = 2e-5f * (h[k] - l[k]) + 2e-6f * fmaxf(fabsf(l[k]), fabsf(h[k])) + 1e-30f;
    l[k] -= pad;
    h[k] += pad;
  }
  lo[i] = make_float4(l[0], l[1], l[2], 0.f);
  hi[i] = make_float4(h[0], h[1], h[2], 0.f);
}

cudaError_t launchDiskBounds(const float4 *xyzr, const float4 *nrm, uint32_t n, float4 *lo,
                             float4 *hi, cudaStream_t s) {
  if (n)
    diskBoundsKernel<<<(n + 255) / 256, 256, 0, s>>>(xyzr, nrm, n, lo, hi);
  return cudaGetLastError();
}
cudaError_t launchTriBounds(const float4 *v0, const float4 *v1, const float4 *v2, uint32_t n,
                            float4 *lo, float4 *hi, cudaStream_t s) {
  if (n)
    triBoundsKernel<<<(n + 255) / 256, 256, 0, s>>>(v0, v1, v2, n, lo, hi);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// parity / debug kernels.  vr_debug_intersect runs the PRODUCTION traverse
// kernel on caller rays loaded into the pool, then converts the hits.
// ---------------------------------------------------------------------------
__global__ void debugLoadRaysKernel(DeviceScene sc, RayPool pool, const float *rays, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  V3 org = {rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]};
  V3 dir = {rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]};
  pool.od0[i] = make_float4(org.x, org.y, org.z, dir.x);
  pool.od1[i] = make_float2(dir.y, dir.z);
  storeBoundaryHit(sc, pool, i, org, dir);
}
cudaError_t launchDebugLoadRays(const DeviceScene &sc, const RayPool &pool, const float *rays,
                                uint32_t m, cudaStream_t s) {
  if (m)
    debugLoadRaysKernel<<<(m + 255) / 256, 256, 0, s>>>(sc, pool, rays, m);
  return cudaGetLastError();
}

__global__ void debugReadHitsKernel(DeviceScene sc, RayPool pool, uint32_t m, uint32_t *geom,
                                    uint32_t *prim, float *t, uint32_t nbCap, uint32_t *nbCount,
                                    uint32_t *nbOut, const uint32_t *sortedToOrig) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  const float4 a = pool.od0[i];
  const float2 b = pool.od1[i];
  const float4 hv = pool.hit[i];
  const uint32_t hprim = __float_as_uint(hv.y), hgeom = __float_as_uint(hv.z);
  geom[i] = hgeom;
  prim[i] = hgeom == VR_INVALID_ID ? VR_INVALID_ID : (hgeom == 0u ? hprim : sortedToOrig[hprim]);
  t[i] = hv.x;
  if (nbCount) {
    V3 org = {a.x, a.y, a.z}, dir = {a.w, b.x, b.y};
    uint32_t cnt = 0;
    if (sc.geoType == 0 && hgeom == 1u) {
      uint32_t k0 = sc.nbOff[hprim], k1 = sc.nbOff[hprim + 1];
      for (uint32_t k = k0; k < k1; ++k) {
        uint32_t id = sc.nbIdx[k];
        if (checkLocal(sc.prim[2 * id], sc.prim[2 * id + 1], org, dir)) {
          if (cnt < nbCap)
            nbOut[(size_t)i * nbCap + cnt] = sortedToOrig[id];
          ++cnt;
        }
      }
    }
    nbCount[i] = cnt;
  }
}
cudaError_t launchDebugReadHits(const DeviceScene &sc, const RayPool &pool, uint32_t m,
                                uint32_t *geom, uint32_t *prim, float *t, uint32_t nbCap,
                                uint32_t *nbCount, uint32_t *nbOut, const uint32_t *sortedToOrig,
                                cudaStream_t s) {
  if (m)
    debugReadHitsKernel<<<(m + 127) / 128, 128, 0, s>>>(sc, pool, m, geom, prim, t, nbCap,
                                                         nbCount, nbOut, sortedToOrig);
  return cudaGetLastError();
}

template <int D>
__global__ void debugSourceKernel(TraceParams p, uint64_t idxBegin, uint32_t m, float *rays) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  Rng rng;
  rng.init(p.seed, p.stream, idxBegin + i);
  V3 org, d = {0.f, 0.f, 0.f};
  if (p.grid)
    sourceSampleGrid<D>(p.src, p.grid, p.gridN, p.eeGrid, idxBegin + i, rng, org, d);
  else
    sourceSample<D>(p.src, p.ee, rng, org, d);
  V3 dir = fillDir<D>(d);
  rays[6 * i] = org.x;
  rays[6 * i + 1] = org.y;
  rays[6 * i + 2] = org.z;
  rays[6 * i + 3] = dir.x;
  rays[6 * i + 4] = dir.y;
  rays[6 * i + 5] = dir.z;
}
cudaError_t launchDebugSourceRays(const TraceParams &p, uint64_t idxBegin, uint32_t m, float *rays,
                                  cudaStream_t s) {
  if (!m)
    return cudaSuccess;
  if (p.scene.D == 2)
    debugSourceKernel<2><<<(m + 127) / 128, 128, 0, s>>>(p, idxBegin, m, rays);
  else
    debugSourceKernel<3><<<(m + 127) / 128, 128, 0, s>>>(p, idxBegin, m, rays);
  return cudaGetLastError();
}

__global__ void debugMathKernel(int which, const float *x, uint32_t m, float param, float *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  if (which == 0) {
    float s, c;
    sincos2pi(x[i], s, c);
    out[i] = s;
    out[m + i] = c;
  } else if (which == 1) {
    out[i] = powdet(x[i], param);
  } else {
    out[i] = acosdet(x[i]);
  }
}
cudaError_t launchDebugMath(int which, const float *x, uint32_t m, float param, float *out,
                            cudaStream_t s) {
  if (m)
    debugMathKernel<<<(m + 255) / 256, 256, 0, s>>>(which, x, m, param, out);
  return cudaGetLastError();
}

__global__ void debugPhiloxKernel(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                  uint32_t c3, uint32_t *out) {
  uint32_t o[4];
  philox4x32(k0, k1, c0, c1, c2, c3, o);
  out[0] = o[0];
  out[1] = o[1];
  out[2] = o[2];
  out[3] = o[3];
}
cudaError_t launchDebugPhilox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                              uint32_t c3, uint32_t *out, cudaStream_t s) {
  debugPhiloxKernel<<<1, 1, 0, s>>>(k0, k1, c0, c1, c2, c3, out);
  return cudaGetLastError();
}

template <int D>
__global__ void debugReflectKernel(vr_particle_desc p, V3 d, V3 n, uint32_t seed, uint64_t idx,
                                   uint32_t m, float *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  Rng rng;
  rng.init(seed, 0u, idx + i);
  V3 o = surfaceReflection<D>(p, d, n, rng);
  out[3 * i] = o.x;
  out[3 * i + 1] = o.y;
  out[3 * i + 2] = o.z;
}
cudaError_t launchDebugReflect(int kind, int D, const float *rayDir, const float *normal,
                               float coneMinAngle, uint32_t seed, uint64_t idx, uint32_t m,
                               float *out, cudaStream_t s) {
  if (!m)
    return cudaSuccess;
  vr_particle_desc p = {kind, 1.f, 1.f, coneMinAngle};
  V3 d = {rayDir[0], rayDir[1], rayDir[2]}, n = {normal[0], normal[1], normal[2]};
  if (D == 2)
    debugReflectKernel<2><<<(m + 127) / 128, 128, 0, s>>>(p, d, n, seed, idx, m, out);
  else
    debugReflectKernel<3><<<(m + 127) / 128, 128, 0, s>>>(p, d, n, seed, idx, m, out);
  return cudaGetLastError();
}

__global__ void unsortFluxKernel(const unsigned long long *src, const uint32_t *sortedToOrig,
                                 uint32_t n, unsigned long long *dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    dst[sortedToOrig[i]] = src[i];
}
cudaError_t launchUnsortFlux(const unsigned long long *src, const uint32_t *sortedToOrig,
                             uint32_t n, unsigned long long *dst, cudaStream_t s) {
  if (n)
    unsortFluxKernel<<<(n + 255) / 256, 256, 0, s>>>(src, sortedToOrig, n, dst);
  return cudaGetLastError();
}

}  // namespace vr
