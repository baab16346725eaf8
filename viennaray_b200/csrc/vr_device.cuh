// Device-side building blocks of the flux kernel: counter-based RNG,
// deterministic elementary functions, primitive tests, boundary handling,
// particle functors and source sampling.
//
// This translation unit is compiled with -fmad=false: every float expression
// is evaluated unfused and in the order written, so that results are
// bit-reproducible against the CPU oracle (IEEE add/mul/div/sqrt only;
// transcendental functions are the fixed polynomials below, not libdevice).
// Reference lines cited are under include/viennaray/ of ViennaRay v4.2.0.
#pragma once
#include "vr_internal.h"

namespace vr {

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by,
                                      float bz) {
  return (ax * bx + ay * by) + az * bz;
}
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ float dot(const V3 &a, const V3 &b) {
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
__device__ __forceinline__ V3 cross(const V3 &a, const V3 &b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ void normalize(V3 &v) {  // Normalize(): v *= 1/|v|
  float inv = 1.0f / sqrtf(dot(v, v));
  v.x *= inv;
  v.y *= inv;
  v.z *= inv;
}
__device__ __forceinline__ float comp(const V3 &v, int a) {
  return a == 0 ? v.x : (a == 1 ? v.y : v.z);
}
__device__ __forceinline__ void setComp(V3 &v, int a, float s) {
  if (a == 0)
    v.x = s;
  else if (a == 1)
    v.y = s;
  else
    v.z = s;
}

// ---- 256-bit loads (sm_100a LDG.E.256): one L1 wavefront per 32-byte record
// instead of two 128-bit requests.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z),
                 "=r"(b.w)
               : "l"(p));
}
// tuning variants of the traverse kernel's loads (measured in DESIGN.md section 4): cache
// policy of the node / disk records in L1 and prefetches into L1
#ifndef VR_NODE_LD
#define VR_NODE_LD 0  // 1: nodes with L1::evict_last
#endif
#ifndef VR_DISK_LD
#define VR_DISK_LD 1  // 0: default policy, 1: disks with L1::no_allocate (+0.7 % on C4: the disk records of a leaf are read once per traversal and would push node lines out), 2: L1::evict_first
#endif
#ifndef VR_PF_L2
#define VR_PF_L2 0  // 1: the prefetch variants target the L2 instead of the L1
#endif
__device__ __forceinline__ void prefetchL1(const void *p) {
#if VR_PF_L2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ void ldgNode(const void *p, uint4 &a, uint4 &b) {
#if VR_NODE_LD == 1
  asm volatile("ld.global.nc.L1::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z),
                 "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void ldgDisk(const float4 *p, float4 &a, float4 &b) {
#if VR_DISK_LD == 1
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif VR_DISK_LD == 2
  asm volatile("ld.global.nc.L1::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z),
                 "=f"(b.w)
               : "l"(p));
}
// read-once records of the shade kernel (neighbour rows, the hit disk's normal)
__device__ __forceinline__ void ldgOnce(const void *p, uint4 &a, uint4 &b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z),
                 "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ float4 ldgOnce(const float4 *p) {
  float4 a;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
               : "l"(p));
  return a;
}
__device__ __forceinline__ void ldg256(const float4 *p, float4 &a, float4 &b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z),
                 "=f"(b.w)
               : "l"(p));
}

// ---- Philox4x32-10 keyed on (seed, stream), counter (idx, block) ----------
__device__ __forceinline__ void philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                           uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t lo0 = 0xD2511F53u * c0, hi0 = __umulhi(0xD2511F53u, c0);
    uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = __umulhi(0xCD9E8D57u, c2);
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// one block of the stream, as a real call (values in registers both ways): the
// generator is used at a dozen call sites of the shade kernel and ten inlined
// Philox rounds per site bloat it beyond the instruction cache
#ifndef VR_PHILOX_INLINE
#define VR_PHILOX_INLINE 0
#endif
#if VR_PHILOX_INLINE
__device__ __forceinline__ uint4 philoxBlock(
#else
__device__ __noinline__ uint4 philoxBlock(
#endif
    uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                          uint32_t blk) {
  uint32_t o[4];
  philox4x32(k0, k1, c0, c1, blk, 0u, o);
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// Per-ray stream; replaces RNG rngState(tea<3>(idx, seed)),
// rayTraceKernel.hpp:120-121.  The stream is consumed block-aligned per ray segment: the
// source sample starts at block 0 and the processing of every traced segment's hit
// (rayTraceKernel.hpp:169-333: scatter test, reflection, roulette) starts at a fresh block,
// unused words of the previous block are dropped (the oracle does the same, rng_discard).
// So the state that travels with a ray between kernels is the block counter alone, and the
// lanes of a warp that reflect generate their block at one common point instead of
// wherever their buffers happen to run dry.  The four-word buffer is shifted instead of
// indexed so it stays in registers.
struct Rng {
  uint32_t k0, k1, c0, c1, blk;
  uint32_t b0, b1, b2, b3;
  int left;
  __device__ __forceinline__ void init(uint32_t seed, uint32_t stream, uint64_t idx) {
    k0 = seed;
    k1 = stream;
    c0 = (uint32_t)idx;
    c1 = (uint32_t)(idx >> 32);
    blk = 0;
    left = 0;
  }
  // state that survives between kernels: the next block (a block generated ahead of its
  // first draw -- refill() at the common point of a warp's reflecting lanes -- and then
  // not touched, left == 4, is handed back: generating ahead never changes the stream)
  __device__ __forceinline__ uint32_t save() const { return left == 4 ? blk - 1u : blk; }
  __device__ __forceinline__ void load(uint32_t nextBlock, uint32_t seed, uint32_t stream,
                                       uint64_t idx) {
    k0 = seed;
    k1 = stream;
    c0 = (uint32_t)idx;
    c1 = (uint32_t)(idx >> 32);
    blk = nextBlock;
    left = 0;
  }
  __device__ __forceinline__ void discard() {
    if (left == 4)
      --blk;
    left = 0;
  }
  __device__ __forceinline__ void refill() {
    const uint4 o = philoxBlock(k0, k1, c0, c1, blk);
    b0 = o.x;
    b1 = o.y;
    b2 = o.z;
    b3 = o.w;
    ++blk;
    left = 4;
  }
  __device__ __forceinline__ uint32_t u32() {
    if (left == 0)
      refill();
    uint32_t r = b0;
    b0 = b1;
    b1 = b2;
    b2 = b3;
    --left;
    return r;
  }
  __device__ __forceinline__ float f() {  // uniform [0,1), top 24 bits
    return (float)(u32() >> 8) * 5.9604644775390625e-8f;
  }
};

// ---- deterministic elementary functions -----------------------------------
__device__ __forceinline__ void sincos2pi(float x, float &s, float &c) {
  int k = (int)(x * 4.0f + 0.5f);
  float r = x - (float)k * 0.25f;
  float a = r * 6.2831854820251465f;
  float a2 = a * a;
  float sp = -1.9841270114e-4f + a2 * 2.7557318840e-6f;
  sp = 8.3333337680e-3f + a2 * sp;
  sp = -1.6666667163e-1f + a2 * sp;
  float sn = a + (a * a2) * sp;
  float cp = -1.3888889225e-3f + a2 * 2.4801587642e-5f;
  cp = 4.1666667908e-2f + a2 * cp;
  cp = -0.5f + a2 * cp;
  float cs = 1.0f + a2 * cp;
  switch (k & 3) {
  case 0:
    s = sn;
    c = cs;
    break;
  case 1:
    s = cs;
    c = -sn;
    break;
  case 2:
    s = -sn;
    c = -cs;
    break;
  default:
    s = -cs;
    c = sn;
    break;
  }
}
__device__ __forceinline__ void sincosRad(float a, float &s, float &c) {
  sincos2pi(a * 0.15915493667125702f, s, c);
}
__device__ __forceinline__ float log2det(float x) {
  uint32_t u = __float_as_uint(x);
  int e = (int)(u >> 23) - 127;
  float m = __uint_as_float((u & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421354f) {
    m = m * 0.5f;
    e += 1;
  }
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float s2 = s * s;
  float p = 0.14285714924f + s2 * 0.11111111194f;
  p = 0.20000000298f + s2 * p;
  p = 0.33333334327f + s2 * p;
  p = 1.0f + s2 * p;
  float ln = (2.0f * s) * p;
  return (float)e + ln * 1.4426950216293335f;
}
__device__ __forceinline__ float exp2det(float y) {
  if (y < -126.0f)
    return 0.0f;
  float kf = floorf(y + 0.5f);
  float f = (y - kf) * 0.69314718246459961f;
  float p = 1.9841270114e-4f + f * 2.4801587642e-5f;
  p = 1.3888889225e-3f + f * p;
  p = 8.3333337680e-3f + f * p;
  p = 4.1666667908e-2f + f * p;
  p = 1.6666667163e-1f + f * p;
  p = 0.5f + f * p;
  p = 1.0f + f * p;
  p = 1.0f + f * p;
  int k = (int)kf + 127;
  if (k <= 0)
    return 0.0f;
  return p * __uint_as_float((uint32_t)k << 23);
}
__device__ __forceinline__ float powdet(float x, float e) {
  if (x <= 0.0f)
    return 0.0f;
  if (e == 0.5f)
    return sqrtf(x);
  float r = exp2det(e * log2det(x));
  return r > 1.0f ? 1.0f : r;
}
__device__ __forceinline__ float acosdet(float x) {  // [0,1], A&S 4.4.46
  float p = 0.0066700901f + x * -0.0012624911f;
  p = -0.0170881256f + x * p;
  p = 0.0308918810f + x * p;
  p = -0.0501743046f + x * p;
  p = 0.0889789874f + x * p;
  p = -0.2145988016f + x * p;
  p = 1.5707963050f + x * p;
  return sqrtf(1.0f - x) * p;
}

// ---- closest-hit rule -------------------------------------------------------
// smallest t; ties -> smaller geomID, then smaller ORIGINAL primID.
struct Hit {
  float t;
  uint32_t geom;
  uint32_t prim;  // internal index for geometry, 0..7 for the boundary
  uint32_t orig;  // original primitive ID (tie-break key)
};
// smallest t, then lower geomID, then lower original primID; written without branches
// (t and b.t are never NaN here) because only a few lanes of a warp ever get this far
__device__ __forceinline__ bool better(float t, uint32_t geom, uint32_t orig, const Hit &b) {
  const bool tie = (geom < b.geom) | ((geom == b.geom) & (orig < b.orig));
  return (t < b.t) | ((t == b.t) & tie);
}

// oriented disc: den = dot(dir,n); t = dot(c-org,n)/den; tnear <= t;
// |org + dir t - c|^2 < r^2 (Embree DiscIntersector1, oriented variant).
#ifndef VR_DISK_FMA
#define VR_DISK_FMA 0
#endif
__device__ __forceinline__ void testDisk(const float4 P, const float4 N, uint32_t prim,
                                         const V3 &org, const V3 &dir, Hit &best) {
#if VR_DISK_FMA
  // fused formulation (timing experiment; the oracle would have to follow)
  float den = __fmaf_rn(dir.z, N.z, __fmaf_rn(dir.y, N.y, dir.x * N.x));
  if (den == 0.f)
    return;
  const float num = __fmaf_rn(P.z - org.z, N.z, __fmaf_rn(P.y - org.y, N.y, (P.x - org.x) * N.x));
  if (num == 0.f)
    return;
  float t = num / den;
  if (!(VR_TNEAR <= t && t <= 3.402823466e+38f))
    return;
  float qx = __fmaf_rn(dir.x, t, org.x) - P.x, qy = __fmaf_rn(dir.y, t, org.y) - P.y,
        qz = __fmaf_rn(dir.z, t, org.z) - P.z;
  if (!(__fmaf_rn(qz, qz, __fmaf_rn(qy, qy, qx * qx)) < P.w * P.w))
    return;
#else
  float den = dot3(dir.x, dir.y, dir.z, N.x, N.y, N.z);
  if (den == 0.f)
    return;
  // a ray that starts in the disk's own plane (every reflection off a flat region meets
  // its coplanar neighbours like this) has num == 0 exactly: t = +-0 fails the tnear test
  // either way, and returning here keeps the warp out of the division's slow path
  const float num = dot3(P.x - org.x, P.y - org.y, P.z - org.z, N.x, N.y, N.z);
  if (num == 0.f)
    return;
  float t = num / den;
  if (!(VR_TNEAR <= t && t <= 3.402823466e+38f))
    return;
  float qx = (org.x + dir.x * t) - P.x, qy = (org.y + dir.y * t) - P.y,
        qz = (org.z + dir.z * t) - P.z;
  if (!(dot3(qx, qy, qz, qx, qy, qz) < P.w * P.w))
    return;
#endif
  uint32_t orig = __float_as_uint(N.w);
  if (better(t, 1u, orig, best)) {
    best.t = t;
    best.geom = 1u;
    best.prim = prim;
    best.orig = orig;
  }
}

// Moeller-Trumbore in Embree's formulation (e1 = v0-v1, e2 = v2-v0,
// Ng = cross(e2,e1)); returns Ng through ng when the hit is taken.
__device__ __forceinline__ bool testTri(const V3 &v0, const V3 &v1, const V3 &v2, uint32_t geom,
                                        uint32_t prim, uint32_t orig, const V3 &org,
                                        const V3 &dir, Hit &best, V3 *ng) {
  V3 e1 = {v0.x - v1.x, v0.y - v1.y, v0.z - v1.z};
  V3 e2 = {v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
  V3 Ng = cross(e2, e1);
  V3 C = {v0.x - org.x, v0.y - org.y, v0.z - org.z};
  V3 R = cross(C, dir);
  float den = dot(Ng, dir);
  if (den == 0.f)
    return false;
  float absDen = fabsf(den);
  float U = dot(R, e2), V = dot(R, e1), T = dot(Ng, C);
  if (den < 0.f) {
    U = -U;
    V = -V;
    T = -T;
  }
  if (!(U >= 0.f && V >= 0.f && U + V <= absDen))
    return false;
  if (!(absDen * VR_TNEAR < T && T <= absDen * 3.402823466e+38f))
    return false;
  float t = T / absDen;
  if (better(t, geom, orig, best)) {
    best.t = t;
    best.geom = geom;
    best.prim = prim;
    best.orig = orig;
    if (ng)
      *ng = Ng;
    return true;
  }
  return false;
}

// rayTraceKernel.hpp:462-507 checkLocalIntersection
__device__ __forceinline__ bool checkLocal(const float4 P, const float4 N, const V3 &org,
                                           const V3 &dir) {
  float prod = dot3(N.x, N.y, N.z, dir.x, dir.y, dir.z);
  if (prod > 0.f)
    return false;
  if (fabsf(prod) < 1e-6f)
    return false;
  float ddneg = dot3(P.x, P.y, P.z, N.x, N.y, N.z);
  float tt = (ddneg - dot3(N.x, N.y, N.z, org.x, org.y, org.z)) / prod;
  if (tt <= 0.f)
    return false;
  float hx = (org.x + dir.x * tt) - P.x, hy = (org.y + dir.y * tt) - P.y,
        hz = (org.z + dir.z * tt) - P.z;
  float distance = sqrtf(dot3(hx, hy, hz, hx, hy, hz));
  return P.w > distance;
}

// checkLocalIntersection with the impact distance handed out (WDIST)
__device__ __forceinline__ bool checkLocalDist(const float4 P, const float4 N, const V3 &org,
                                               const V3 &dir, float &distance) {
  float prod = dot3(N.x, N.y, N.z, dir.x, dir.y, dir.z);
  if (prod > 0.f)
    return false;
  if (fabsf(prod) < 1e-6f)
    return false;
  float ddneg = dot3(P.x, P.y, P.z, N.x, N.y, N.z);
  float tt = (ddneg - dot3(N.x, N.y, N.z, org.x, org.y, org.z)) / prod;
  if (tt <= 0.f)
    return false;
  float hx = (org.x + dir.x * tt) - P.x, hy = (org.y + dir.y * tt) - P.y,
        hz = (org.z + dir.z * tt) - P.z;
  distance = sqrtf(dot3(hx, hy, hz, hx, hy, hz));
  return P.w > distance;
}

// rayUtil.hpp:204-215 fillRayDirection<D>
template <int D> __device__ __forceinline__ V3 fillDir(const V3 &direction) {
  V3 r = direction;
  if (D == 2 && r.z != 0.f) {
    r.z = 0.f;
    normalize(r);
  }
  return r;
}

// rayReflection.hpp:13-29
__device__ __forceinline__ V3 reflectSpecular(const V3 &d, const V3 &n) {
  V3 v = {-d.x, -d.y, -d.z};
  float f = 2.f * dot(n, v);
  return {f * n.x - v.x, f * n.y - v.y, f * n.z - v.z};
}

// rayUtil.hpp:266-283 (Marsaglia) + rayReflection.hpp:32-50
template <int D> __device__ __forceinline__ V3 reflectDiffuse(const V3 &n, Rng &rng) {
  float x, y, s2;
  do {
    x = 2.f * rng.f() - 1.f;
    y = 2.f * rng.f() - 1.f;
    s2 = x * x + y * y;
  } while (s2 >= 1.f);
  float tmp = 2.f * sqrtf(1.f - s2);
  V3 o;
  o.x = x * tmp + n.x;
  o.y = y * tmp + n.y;
  o.z = D == 3 ? (1.f - 2.f * s2) + n.z : 0.f;
  normalize(o);
  return o;
}

// rayReflection.hpp:52-120
template <int D>
__device__ __forceinline__ V3 reflectConedCosine(const V3 &d, const V3 &n, Rng &rng, float cone) {
  if (cone <= 0.f)
    return reflectSpecular(d, n);
  if (cone >= 1.57079637050628662f)
    return reflectDiffuse<D>(n, rng);
  V3 w = reflectSpecular(d, n);
  normalize(w);
  V3 t, b;
  if (w.z < -0.999999f) {
    t = {0.f, -1.f, 0.f};
    b = {-1.f, 0.f, 0.f};
  } else {
    float a = 1.f / (1.f + w.z);
    float bx = -w.x * w.y * a;
    float by = 1.f - w.y * w.y * a;
    t = {1.f - w.x * w.x * a, bx, -w.x};
    b = {bx, by, -w.y};
  }
  float theta, sn, cs;
  for (;;) {
    float u = sqrtf(rng.f());
    float q = 1.f - u;
    float s = sqrtf(q > 0.f ? q : 0.f);
    theta = cone * s;
    float cHalf, sHalf, sTheta, cTheta;
    sincos2pi(0.25f * s, sHalf, cHalf);
    sincosRad(theta, sTheta, cTheta);
    float rhs = cHalf * sTheta;
    if (rng.f() * theta * u <= rhs) {
      sn = sTheta;
      cs = cTheta;
      break;
    }
  }
  float sp, cp;
  sincos2pi(rng.f(), sp, cp);
  V3 o;
  o.x = sn * (cp * t.x + sp * b.x) + cs * w.x;
  o.y = sn * (cp * t.y + sp * b.y) + cs * w.y;
  o.z = sn * (cp * t.z + sp * b.z) + cs * w.z;
  float dp = dot(o, n);
  if (dp <= 0.f) {
    float f = 2.f * dp;
    o.x = o.x - f * n.x;
    o.y = o.y - f * n.y;
    o.z = o.z - f * n.z;
  }
  if (D == 2)
    o.z = 0.f;
  normalize(o);
  return o;
}

// particle functor: surfaceReflection of the built-in particles
// (rayParticle.hpp:137-146,177-186; cone recipe tests/reflection/reflection.cpp:43-46)
template <int D>
__device__ __forceinline__ V3 surfaceReflection(const vr_particle_desc &p, const V3 &d,
                                                const V3 &n, Rng &rng) {
  if (p.kind == VR_PARTICLE_DIFFUSE)
    return reflectDiffuse<D>(n, rng);
  if (p.kind == VR_PARTICLE_SPECULAR)
    return reflectSpecular(d, n);
  float c = -dot(d, n);
  c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
  float inc = acosdet(c);
  float m = inc < p.coneMinAngle ? inc : p.coneMinAngle;
  return reflectConedCosine<D>(d, n, rng, 1.57079637050628662f - m);
}

// raySourceGrid.hpp:23-52
template <int D>
__device__ __forceinline__ void sourceSampleGrid(const vr_source_desc &s, const float *grid,
                                                 uint32_t gridN, float eeGrid, uint64_t idx,
                                                 Rng &rng, V3 &origin, V3 &direction) {
  const float *g = grid + 3 * (size_t)(idx % gridN);
  origin = {g[0], g[1], g[2]};
  const float r1 = rng.f(), r2 = rng.f();
  const float tt = powdet(r2, eeGrid);
  float sinPhi, cosPhi;
  sincos2pi(r1, sinPhi, cosPhi);
  const float st = sqrtf(1.f - tt);
  setComp(direction, s.rayDir, s.posNeg * sqrtf(tt));
  setComp(direction, s.firstDir, cosPhi * st);
  setComp(direction, s.secondDir, D == 2 ? 0.f : sinPhi * st);
  normalize(direction);
}

// raySourceRandom.hpp:50-116
template <int D>
__device__ __forceinline__ void sourceSample(const vr_source_desc &s, float ee, Rng &rng,
                                             V3 &origin, V3 &direction) {
  origin = {0.f, 0.f, 0.f};
  float r1 = rng.f();
  setComp(origin, s.rayDir, s.minMax ? s.bboxMax[s.rayDir] : s.bboxMin[s.rayDir]);
  setComp(origin, s.firstDir,
          s.bboxMin[s.firstDir] + (s.bboxMax[s.firstDir] - s.bboxMin[s.firstDir]) * r1);
  if (D == 2) {
    setComp(origin, s.secondDir, 0.f);
  } else {
    float r2 = rng.f();
    setComp(origin, s.secondDir,
            s.bboxMin[s.secondDir] + (s.bboxMax[s.secondDir] - s.bboxMin[s.secondDir]) * r2);
  }
  for (;;) {
    float q1 = rng.f(), q2 = rng.f();
    float sinPhi, cosPhi;
    sincos2pi(q1, sinPhi, cosPhi);
    float cosTheta = powdet(q2, ee);
    float sinTheta = sqrtf(1.f - cosTheta * cosTheta);
    if (!s.useBasis) {
      setComp(direction, s.rayDir, s.posNeg * cosTheta);
      setComp(direction, s.firstDir, cosPhi * sinTheta);
      setComp(direction, s.secondDir, sinPhi * sinTheta);
      return;
    }
    float r0 = cosTheta, r1b = cosPhi * sinTheta, r2b = sinPhi * sinTheta;
    direction.x = (s.basis[0] * r0 + s.basis[3] * r1b) + s.basis[6] * r2b;
    direction.y = (s.basis[1] * r0 + s.basis[4] * r1b) + s.basis[7] * r2b;
    direction.z = (s.basis[2] * r0 + s.basis[5] * r1b) + s.basis[8] * r2b;
    float dr = comp(direction, s.rayDir);
    if (!((s.posNeg < 0.f && dr > 0.f) || (s.posNeg > 0.f && dr < 0.f)))
      return;
  }
}

}  // namespace vr
