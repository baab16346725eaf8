// The hot path: a persistent kernel that runs the whole Monte Carlo walk of
// rayInternal::TraceKernel::apply (rayTraceKernel.hpp:117-338) per ray --
// source sampling, BVH traversal with disk / triangle tests, boundary
// handling, neighbour spread, particle reflection and Russian roulette -- and
// regenerates finished lanes from a global ray cursor.
#include "vr_device.cuh"

namespace vr {

// ---------------------------------------------------------------------------
// closest hit over geometry (BVH) and the 8 boundary triangles
// ---------------------------------------------------------------------------
#define VR_STACK 96

template <int GEO>
__device__ __forceinline__ void testLeaf(const DeviceScene &sc, uint32_t ref, const V3 &org,
                                         const V3 &dir, Hit &best, unsigned &primTests) {
  uint32_t first = (ref & 0x7fffffffu) >> 4, count = ref & 15u;
  for (uint32_t k = 0; k < count; ++k) {
    uint32_t i = first + k;
    if (GEO == 0) {
      float4 P = __ldg(&sc.primA[i]);
      float4 N = __ldg(&sc.primB[i]);
      testDisk(P, N, i, org, dir, best);
    } else {
      float4 a = __ldg(&sc.primA[i]), b = __ldg(&sc.primB[i]), c = __ldg(&sc.primC[i]);
      testTri({a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, 1u, i, __float_as_uint(a.w), org,
              dir, best, nullptr);
    }
    ++primTests;
  }
}

template <int GEO>
__device__ __forceinline__ void intersectScene(const DeviceScene &sc, const V3 &org, const V3 &dir,
                                               Hit &best, V3 &bng, unsigned &nodeVisits,
                                               unsigned &primTests) {
  best.t = 3.402823466e+38f;
  best.geom = VR_INVALID_ID;
  best.prim = VR_INVALID_ID;
  best.orig = VR_INVALID_ID;
  const float idx = 1.f / dir.x, idy = 1.f / dir.y, idz = 1.f / dir.z;

  uint32_t stack[VR_STACK];
  int sp = 0;
  uint32_t cur = sc.rootRef;
  if (sc.numPrims == 0)
    cur = VR_INVALID_ID;
  while (cur != VR_INVALID_ID) {
    if (cur & VR_LEAF_FLAG) {
      testLeaf<GEO>(sc, cur, org, dir, best, primTests);
      cur = sp ? stack[--sp] : VR_INVALID_ID;
      continue;
    }
    const Node2 *n = sc.nodes + cur;
    const float4 a = __ldg(&n->a), b = __ldg(&n->b), c = __ldg(&n->c), d = __ldg(&n->d);
    ++nodeVisits;
    // conservative slab tests against [tnear, best.t]; NaN (0 * inf) compares
    // false in fminf/fmaxf and leaves the other bound in place
    float t0x = (a.x - org.x) * idx, t1x = (a.w - org.x) * idx;
    float t0y = (a.y - org.y) * idy, t1y = (b.x - org.y) * idy;
    float t0z = (a.z - org.z) * idz, t1z = (b.y - org.z) * idz;
    float n0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), VR_TNEAR));
    float f0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best.t));
    float u0x = (b.z - org.x) * idx, u1x = (c.y - org.x) * idx;
    float u0y = (b.w - org.y) * idy, u1y = (c.z - org.y) * idy;
    float u0z = (c.x - org.z) * idz, u1z = (c.w - org.z) * idz;
    float n1 = fmaxf(fmaxf(fminf(u0x, u1x), fminf(u0y, u1y)), fmaxf(fminf(u0z, u1z), VR_TNEAR));
    float f1 = fminf(fminf(fmaxf(u0x, u1x), fmaxf(u0y, u1y)), fminf(fmaxf(u0z, u1z), best.t));
    // widen: boxes are padded at build time; the factors absorb the rounding
    // of the slab arithmetic itself
    bool h0 = n0 * 0.999999f <= f0 * 1.000001f + 1e-6f;
    bool h1 = n1 * 0.999999f <= f1 * 1.000001f + 1e-6f;
    uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y);
    if (h0 && h1) {
      bool swap = n1 < n0;
      uint32_t nearRef = swap ? r1 : r0, farRef = swap ? r0 : r1;
      if (sp < VR_STACK)
        stack[sp++] = farRef;
      cur = nearRef;
    } else if (h0) {
      cur = r0;
    } else if (h1) {
      cur = r1;
    } else {
      cur = sp ? stack[--sp] : VR_INVALID_ID;
    }
  }
  // boundary box (geomID 0): wins ties against geometry
#pragma unroll 1
  for (uint32_t i = 0; i < 8; ++i) {
    V3 v0 = {sc.btri[i][0][0], sc.btri[i][0][1], sc.btri[i][0][2]};
    V3 v1 = {sc.btri[i][1][0], sc.btri[i][1][1], sc.btri[i][1][2]};
    V3 v2 = {sc.btri[i][2][0], sc.btri[i][2][1], sc.btri[i][2][2]};
    testTri(v0, v1, v2, 0u, i, i, org, dir, best, &bng);
  }
}

// ---------------------------------------------------------------------------
// Boundary::processHit, rayBoundary.hpp:29-127.  Returns `reflect`.
// ---------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ bool boundaryHit(const DeviceScene &sc, V3 &org, V3 &rayDirection,
                                            V3 &dir, const V3 &ng, uint32_t primID, float t) {
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

__device__ __forceinline__ unsigned long long warpSum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// persistent trace kernel
// ---------------------------------------------------------------------------
template <int D, int GEO>
__global__ void __launch_bounds__(128) traceKernel(const __grid_constant__ TraceParams p) {
  const DeviceScene &sc = p.scene;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned ltMask = (1u << lane) - 1u;
  const uint64_t numRays = p.idxEnd - p.idxBegin;

  // per-lane ray state
  bool alive = false;
  bool hitFromBack = false;
  V3 org, dir, rayDirection;
  float w = 0.f;
  unsigned numReflections = 0, boundaryHits = 0;
  Rng rng;
  // TraceInfo counters (rayTraceKernel.hpp:49-55)
  unsigned long long cTraces = 0, cMiss = 0, cGeo = 0, cBnd = 0, cRefl = 0, cTerm = 0;
  unsigned wNodes = 0, wPrims = 0, wNb = 0, wFlux = 0;
  bool exhausted = false;

  for (;;) {
    // ---- regenerate finished lanes (warp-aggregated cursor fetch) ----------
    unsigned dead = __ballot_sync(0xffffffffu, !alive);
    if (dead && !exhausted) {
      unsigned nDead = __popc(dead);
      int leader = __ffs(dead) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader)
        base = atomicAdd(p.rayCursor, (unsigned long long)nDead);
      base = __shfl_sync(0xffffffffu, base, leader);
      if (base + nDead >= numRays)
        exhausted = true;  // warp-uniform
      if (!alive) {
        unsigned long long off = base + __popc(dead & ltMask);
        if (off < numRays) {
          uint64_t idx = p.idxBegin + off;
          rng.init(p.seed, p.stream, idx);
          w = 1.f;  // getInitialRayWeight, raySource.hpp:18
          sourceSample<D>(p.src, p.ee, rng, org, rayDirection);
          dir = fillDir<D>(rayDirection);
          numReflections = 0;
          boundaryHits = 0;
          hitFromBack = false;
          alive = true;
        }
      }
    }
    if (!__any_sync(0xffffffffu, alive))
      break;
    if (alive) {
    // ---- one trace step -----------------------------------------------------
    Hit h;
    V3 bng = {0.f, 0.f, 0.f};
    intersectScene<GEO>(sc, org, dir, h, bng, wNodes, wPrims);
    ++cTraces;
    bool finish = false;
    if (h.geom == VR_INVALID_ID) {  // :172
      ++cMiss;
      finish = true;
    } else if (h.geom == 0u) {  // :206-214
      if (++boundaryHits > p.maxBoundaryHits) {
        ++cTerm;
        finish = true;
      } else if (!boundaryHit<D>(sc, org, rayDirection, dir, bng, h.prim, h.t)) {
        finish = true;
      }
    } else {
      V3 hitPoint = {org.x + dir.x * h.t, org.y + dir.y * h.t, org.z + dir.z * h.t};
      float4 N4 = __ldg(&sc.primN[h.prim]);
      V3 gn = {N4.x, N4.y, N4.z};
      bool backface = dot(rayDirection, gn) > 0.f;  // :224
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
        unsigned long long wf = toFixed(w);
        atomicAdd(&p.flux[h.prim], wf);  // :297-306 surfaceCollision
        ++wFlux;
        if (GEO == 0) {  // :271-280 neighbour spread
          uint32_t k0 = __ldg(&sc.nbOff[h.prim]), k1 = __ldg(&sc.nbOff[h.prim + 1]);
          for (uint32_t k = k0; k < k1; ++k) {
            uint32_t id = __ldg(&sc.nbIdx[k]);
            float4 P = __ldg(&sc.primA[id]);
            float4 Nn = __ldg(&sc.primB[id]);
            ++wNb;
            if (checkLocal(P, Nn, org, dir)) {
              atomicAdd(&p.flux[id], wf);
              ++wFlux;
            }
          }
        }
        V3 newDir = surfaceReflection<D>(p.particle, rayDirection, gn, rng);  // :310
        w -= w * p.particle.sticking;                                         // :316
        if (w <= 0.f) {
          finish = true;
        } else if (++numReflections > p.maxReflections) {
          ++cTerm;
          finish = true;
        } else {
          // :435-460 rejectionControl, thresholds 0.1 / 0.3 of the initial weight
          if (w < 0.1f) {
            float kill = 1.f - w / 0.3f;
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
      }
    }
    if (finish) {
      cBnd += boundaryHits;
      cRefl += numReflections;
      alive = false;
    }
    }  // alive
  }

  // ---- counters: warp reduce, one atomic per warp --------------------------
  cTraces = warpSum(cTraces);
  cMiss = warpSum(cMiss);
  cGeo = warpSum(cGeo);
  cBnd = warpSum(cBnd);
  cRefl = warpSum(cRefl);
  cTerm = warpSum(cTerm);
  if (lane == 0) {
    atomicAdd(&p.counters[1], cTraces);
    atomicAdd(&p.counters[2], cMiss);
    atomicAdd(&p.counters[3], cGeo);
    atomicAdd(&p.counters[5], cBnd);
    atomicAdd(&p.counters[6], cRefl);
    atomicAdd(&p.counters[7], cTerm);
  }
  if (p.work) {
    unsigned long long a = warpSum((unsigned long long)wNodes), b = warpSum((unsigned long long)wPrims),
                       c = warpSum((unsigned long long)wNb), d = warpSum((unsigned long long)wFlux);
    if (lane == 0) {
      atomicAdd(&p.work[0], a);
      atomicAdd(&p.work[1], b);
      atomicAdd(&p.work[2], c);
      atomicAdd(&p.work[3], d);
    }
  }
}

template <int D, int GEO>
static cudaError_t launchTraceT(const TraceParams &p, int numSMs, cudaStream_t s) {
  int perSM = 0;
  cudaError_t e =
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernel<D, GEO>, 128, 0);
  if (e != cudaSuccess)
    return e;
  if (perSM < 1)
    perSM = 1;
  uint64_t numRays = p.idxEnd - p.idxBegin;
  uint64_t want = (numRays + 127) / 128;
  uint64_t grid = (uint64_t)numSMs * perSM;
  if (want < grid)
    grid = want ? want : 1;
  traceKernel<D, GEO><<<(unsigned)grid, 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launchTrace(const TraceParams &p, int numSMs, cudaStream_t s, int *launches) {
  if (launches)
    *launches += 1;
  if (p.scene.geoType == 0)
    return p.scene.D == 2 ? launchTraceT<2, 0>(p, numSMs, s) : launchTraceT<3, 0>(p, numSMs, s);
  return p.scene.D == 2 ? launchTraceT<2, 1>(p, numSMs, s) : launchTraceT<3, 1>(p, numSMs, s);
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
    float pad = 2e-4f * P.w + 1e-6f * fabsf(c[a]);
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
    float pad = 2e-5f * (h[k] - l[k]) + 1e-6f * fmaxf(fabsf(l[k]), fabsf(h[k])) + 1e-30f;
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
// parity / debug kernels
// ---------------------------------------------------------------------------
template <int GEO>
__global__ void debugIntersectKernel(DeviceScene sc, const float *rays, uint32_t m, uint32_t *geom,
                                     uint32_t *prim, float *t, uint32_t nbCap, uint32_t *nbCount,
                                     uint32_t *nbOut, const uint32_t *sortedToOrig) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  V3 org = {rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]};
  V3 dir = {rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]};
  Hit h;
  V3 bng;
  unsigned a = 0, b = 0;
  intersectScene<GEO>(sc, org, dir, h, bng, a, b);
  geom[i] = h.geom;
  prim[i] = h.geom == VR_INVALID_ID ? VR_INVALID_ID : h.orig;
  t[i] = h.t;
  if (nbCount) {
    uint32_t cnt = 0;
    if (GEO == 0 && h.geom == 1u) {
      uint32_t k0 = sc.nbOff[h.prim], k1 = sc.nbOff[h.prim + 1];
      for (uint32_t k = k0; k < k1; ++k) {
        uint32_t id = sc.nbIdx[k];
        if (checkLocal(sc.primA[id], sc.primB[id], org, dir)) {
          if (cnt < nbCap)
            nbOut[(size_t)i * nbCap + cnt] = sortedToOrig[id];
          ++cnt;
        }
      }
    }
    nbCount[i] = cnt;
  }
}

cudaError_t launchDebugIntersect(const DeviceScene &sc, const float *rays, uint32_t m,
                                 uint32_t *geom, uint32_t *prim, float *t, uint32_t nbCap,
                                 uint32_t *nbCount, uint32_t *nbOut, const uint32_t *sortedToOrig,
                                 cudaStream_t s) {
  if (!m)
    return cudaSuccess;
  if (sc.geoType == 0)
    debugIntersectKernel<0><<<(m + 127) / 128, 128, 0, s>>>(sc, rays, m, geom, prim, t, nbCap,
                                                             nbCount, nbOut, sortedToOrig);
  else
    debugIntersectKernel<1><<<(m + 127) / 128, 128, 0, s>>>(sc, rays, m, geom, prim, t, nbCap,
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
  V3 org, d;
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
