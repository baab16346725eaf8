// Internal structures shared by the host runtime and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/viennaray_b200.h"

#define VR_INVALID_ID 0xffffffffu
#define VR_TNEAR 1e-4f            // fillRayPosition default (rayUtil.hpp:218)
#define VR_LEAF_MAX 4u            // primitives per BVH leaf
#define VR_LEAF_FLAG 0x80000000u  // child reference: leaf(first << 4 | count)
#define VR_FIXED_SCALE 1073741824.0f

namespace vr {

// Binary BVH node, 64 bytes: the boxes of both children and their references.
//   a = (lo0.x, lo0.y, lo0.z, hi0.x)   b = (hi0.y, hi0.z, lo1.x, lo1.y)
//   c = (lo1.z, hi1.x, hi1.y, hi1.z)   d = (ref0, ref1, -, -) as bit patterns
struct alignas(16) Node2 {
  float4 a, b, c, d;
};

struct DeviceScene {
  int D;
  int geoType;  // 0 disk, 1 triangle
  uint32_t numPrims;
  // primitives in BVH (Morton) order -- the INTERNAL primitive index
  const float4 *primA;  // disk: x,y,z,r        triangle: v0
  const float4 *primB;  // disk: nx,ny,nz,orig  triangle: v1
  const float4 *primC;  //                      triangle: v2
  const float4 *primN;  // disk: = primB        triangle: nx,ny,nz,orig
  const uint32_t *nbOff;  // neighbour CSR, internal indices
  const uint32_t *nbIdx;
  const Node2 *nodes;
  uint32_t rootRef;
  // boundary (rayBoundary.hpp:164-245)
  float bbox[2][3];
  int firstDir, secondDir;
  int bc[2];
  float btri[8][3][3];  // 8 triangles x 3 vertices
};

struct TraceParams {
  DeviceScene scene;
  vr_source_desc src;
  vr_particle_desc particle;
  float ee;  // 1 / (sourcePower + 1), raySourceRandom.hpp:21
  uint64_t idxBegin, idxEnd;
  uint32_t seed, stream;
  uint32_t maxReflections, maxBoundaryHits;
  unsigned long long *flux;      // numPrims fixed-point sums (internal order)
  unsigned long long *counters;  // 8 TraceInfo counters
  unsigned long long *rayCursor; // next ray offset
  unsigned long long *work;      // optional work counters (4) or null
};

// ---- acceleration structure (vr_bvh.cu) ----------------------------------
struct Bvh {
  Node2 *nodes = nullptr;
  uint32_t numNodes = 0;
  uint32_t rootRef = 0;
  uint32_t *sortedToOrig = nullptr;  // device, numPrims
  float buildMs = 0.f;
  uint32_t numLeaves = 0, maxLeaf = 0;
};
// primLo/primHi: device float4 arrays of per-primitive padded boxes (original
// order); centers from the box centre.  sceneLo/Hi: host.
cudaError_t buildBvh(const float4 *primLo, const float4 *primHi, uint32_t n, const float sceneLo[3],
                     const float sceneHi[3], cudaStream_t stream, Bvh *out);
void freeBvh(Bvh *b);

// ---- kernels (vr_trace.cu) -------------------------------------------------
cudaError_t launchDiskBounds(const float4 *xyzr, const float4 *nrm, uint32_t n, float4 *lo,
                             float4 *hi, cudaStream_t s);
cudaError_t launchTriBounds(const float4 *v0, const float4 *v1, const float4 *v2, uint32_t n,
                            float4 *lo, float4 *hi, cudaStream_t s);
cudaError_t launchTrace(const TraceParams &p, int numSMs, cudaStream_t s, int *launches);
cudaError_t launchDebugIntersect(const DeviceScene &sc, const float *rays, uint32_t m,
                                 uint32_t *geom, uint32_t *prim, float *t, uint32_t nbCap,
                                 uint32_t *nbCount, uint32_t *nbOut, const uint32_t *sortedToOrig,
                                 cudaStream_t s);
cudaError_t launchDebugSourceRays(const TraceParams &p, uint64_t idxBegin, uint32_t m, float *rays,
                                  cudaStream_t s);
cudaError_t launchDebugMath(int which, const float *x, uint32_t m, float param, float *out,
                            cudaStream_t s);
cudaError_t launchDebugPhilox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                              uint32_t c3, uint32_t *out, cudaStream_t s);
cudaError_t launchDebugReflect(int kind, int D, const float *rayDir, const float *normal,
                               float coneMinAngle, uint32_t seed, uint64_t idx, uint32_t m,
                               float *out, cudaStream_t s);
cudaError_t launchUnsortFlux(const unsigned long long *fluxSorted, const uint32_t *sortedToOrig,
                             uint32_t n, unsigned long long *fluxOrig, cudaStream_t s);

}  // namespace vr
