// Internal structures shared by the host runtime and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/viennaray_b200.h"

#define VR_INVALID_ID 0xffffffffu
#define VR_TNEAR 1e-4f            // fillRayPosition default (rayUtil.hpp:218)
// primitives per BVH leaf (<= 15): 8-10 disks measured best once the disk test lost its
// division slow path (round 2, with the disk records bypassing the L1: 6 / 8 / 10 / 12 give
// 284.3 / 279.7 / 278.0 / 289.2 ms on a 2 x 256e6-ray C4 trace; C5 does not care); a
// triangle test costs twice a disk test and 4 stay best there
#ifndef VR_LEAF_MAX
#define VR_LEAF_MAX 10u
#endif
#ifndef VR_LEAF_MAX_TRI
#define VR_LEAF_MAX_TRI 4u
#endif
#define VR_STACK 96          // entries of a ray's traversal stack (see vr_trace.cu)
#define VR_DONE 0x7fffffffu  // traversal finished / unused child (not a valid node index)
#define VR_LEAF_FLAG 0x80000000u  // child reference: leaf(first << 4 | count)
#define VR_FIXED_SCALE 1073741824.0f
#define VR_COUNTER_COPIES 64      // replicated TraceInfo counters (summed on download)
// References >= VR_TOP_BASE (and < VR_DONE) name an entry of the top-of-tree table that the
// traverse kernel stages in shared memory (primitive counts stay below 2^27, so real node
// indices never get here)
#define VR_TOP_BASE 0x7f000000u
#define VR_TOP_MAX 4096u          // capacity of the table (entries of 32 bytes)

namespace vr {

// Compressed binary BVH node, 32 bytes = one L2 sector: per child a box of six
// 16-bit coordinates on the scene's quantisation grid (rounded outwards) and a
// 32-bit reference.
//   c.x = lo.x | hi.x << 16   c.y = lo.y | hi.y << 16   c.z = lo.z | hi.z << 16
//   c.w = reference (inner node index, or VR_LEAF_FLAG | first << 4 | count)
struct alignas(16) Node2 {
  uint4 c0, c1;
};

struct DeviceScene {
  int D;
  int geoType;  // 0 disk, 1 triangle
  uint32_t numPrims;
  // primitives in BVH (Morton) order -- the INTERNAL primitive index.
  // disk i:     prim[2i] = x,y,z,r   prim[2i+1] = nx,ny,nz,original ID  (one sector)
  // triangle i: prim[4i..4i+2] = v0,v1,v2 (w = original ID)  prim[4i+3] = normal
  const float4 *prim;
  const uint32_t *nbOff;  // neighbour CSR, internal indices
  const uint32_t *nbIdx;
  // the first eight neighbours of every disk as one 32-byte row (2 x uint4, padded with
  // VR_INVALID_ID); bit 31 of the eighth word: the CSR row continues at entry 8
  const uint4 *nbRow;
  const Node2 *nodes;
  const uint4 *nodes4;  // optional 4-wide nodes or null
  uint32_t rootRef;
  // top of the tree in breadth-first order (2 x uint4 per entry like a Node2; child
  // references inside the table are VR_TOP_BASE + entry); entry 0 is the root.  topCount 0:
  // no table.
  const uint4 *top;
  uint32_t topCount;
  float qLo[3], qScale[3];  // node box coordinate = qLo + q * qScale
  // sky map (vr_scene.cu buildSky): a G x G grid over the two lateral axes; per
  // cell {base, slope}.  A ray that leaves a surface point of the cell towards
  // the source with a steeper slope than `slope` cannot touch any primitive
  // (slope = +inf: no statement for this cell).  sky == nullptr: disabled.
  const float2 *sky;
  int skyN, skyUp, skyA, skyB;  // grid size, source axis, lateral axes
  float skySign;                // height = skySign * coordinate[skyUp]
  float skyLo[2], skyInv[2];    // cell = (lateral - skyLo) * skyInv
  float skyTop;                 // height of the highest primitive point
  // boundary (rayBoundary.hpp:164-245)
  float bbox[2][3];
  int firstDir, secondDir;
  int bc[2];
  float btri[8][3][3];  // 8 triangles x 3 vertices
  // shortcut of the boundary test (boundaryTest, vr_trace.cu): set when the lateral axes
  // are x and y.  bN[i]: the one non-zero component of triangle i's unnormalised normal
  // cross(v2 - v0, v0 - v1) as the triangle test computes it, bX[i]: v0's coordinate on
  // the plane's axis, bExt: largest box extent (margin scale)
  int bFast;
  float bExt;
  float bN[8], bX[8];
};

// Resident pool of rays in flight (structure of arrays, one slot per ray).
// The traverse kernel reads od0/od1, starts from hit (the ray's boundary hit, found by the
// shade / init kernel) and writes the closest hit back; the shade kernel owns the rest.
// A slot with dir.x = NaN is empty.
struct RayPool {
  uint32_t capacity;
  float4 *od0;   // org.x, org.y, org.z, dir.x
  float2 *od1;   // dir.y, dir.z
  float4 *hit;   // t, prim (bits), geom (bits), -
  uint32_t *rng; // next block of the ray's Philox stream (vr_device.cuh, struct Rng)
  uint4 *meta;   // idx lo, idx hi, numReflections, boundaryHits | hitFromBack << 31
  float *weight;
  float4 *dir3;  // particle-facing direction (differs from the ray's in 2D only)
};

struct TraceParams {
  DeviceScene scene;
  RayPool pool;     // slots read by traverse and shade
  RayPool poolOut;  // shade, compact mode: survivors are appended here
  int compact;      // 0: survivors stay in place; 1: append to poolOut
  vr_source_desc src;
  vr_particle_desc particle;
  // sticking by material (device copies): matId[internal primitive], matSticking[numMaterials];
  // matSticking == nullptr: the particle's constant sticking
  const int *matId;
  const float *matSticking;
  int numMaterials;
  float ee;      // 1 / (sourcePower + 1), raySourceRandom.hpp:21
  float eeGrid;  // 2 / (sourcePower + 1), raySourceGrid.hpp:21
  const float *grid;  // grid source origins (n x 3) or null
  uint32_t gridN;
  uint64_t idxBegin, idxEnd;
  uint32_t seed, stream;
  uint32_t maxReflections, maxBoundaryHits;
  uint32_t flags;  // VR_FLAG_*
  uint32_t numSlots;             // upper bound of the slots in use (sizes the grid)
  const unsigned int *slotCount; // device: slots actually in use
  unsigned long long *flux;      // numPrims fixed-point sums (internal order)
  unsigned long long *counters;  // VR_COUNTER_COPIES x 8 TraceInfo counters
  unsigned long long *rayCursor; // rays handed out so far
  unsigned int *slotCursor;      // traverse kernel: next slot
  unsigned int *liveCount;       // shade kernel: slots still alive afterwards
  unsigned long long *work;      // optional work counters (4) or null
  // optional neighbour-spread queue (disks, default shade instantiation): the shade kernel
  // appends {org, weight}, {dir, hit disk} per geometry hit and spreadKernel distributes the
  // flux afterwards; null: the spread runs inside the shade kernel
  float4 *spreadQ;
  unsigned int *spreadCount;
};

// ---- acceleration structure (vr_bvh.cu) ----------------------------------
struct Bvh {
  Node2 *nodes = nullptr;
  uint4 *nodes4 = nullptr;  // optional 4-wide nodes (4 x uint4 each), same indices as `nodes`
  uint4 *top = nullptr;     // breadth-first top-of-tree table (see DeviceScene::top)
  uint32_t topCount = 0;
  uint32_t numNodes = 0;
  uint32_t rootRef = 0;
  uint32_t *sortedToOrig = nullptr;  // device, numPrims
  float qLo[3] = {0, 0, 0}, qScale[3] = {1, 1, 1};
  float buildMs = 0.f;
  uint32_t numLeaves = 0, maxLeaf = 0;
  uint32_t maxDepth = 0;  // levels of the radix tree (upper bound of the emitted tree's)
  float mortonAlpha = 1.f;  // shape of the Morton cells the kept tree was built with
  float sahInner = 0.f, sahLeaf = 0.f;  // SAH terms: sum of inner-node areas, of leaf areas x
                                        // primitive counts, both over the root area
};
cudaError_t buildBvh(const float4 *primLo, const float4 *primHi, uint32_t n, const float sceneLo[3],
                     const float sceneHi[3], uint32_t leafMax, float alphaHint, cudaStream_t stream,
                     Bvh *out);
void freeBvh(Bvh *b, cudaStream_t stream);

cudaError_t l2ReadBandwidth(size_t bytes, int passes, int numSMs, cudaStream_t s, double *gbps);

// ---- scene preparation (vr_scene.cu) ----------------------------------------
cudaError_t launchPackDiskNormals(const float *nxyz, uint32_t n, float4 *B, cudaStream_t s);
cudaError_t launchPackTriangles(const float *verts, const uint32_t *tris, const float *normals,
                                uint32_t n, float4 *A, float4 *B, float4 *C, float4 *N,
                                cudaStream_t s);
cudaError_t launchGatherPrims(int geoType, const float4 *A, const float4 *B, const float4 *C,
                              const float4 *N, const uint32_t *s2o, uint32_t n, float4 *prim,
                              cudaStream_t s);
cudaError_t postprocessFluxEx(const unsigned long long *fixedOrig, uint32_t n, const float *areas,
                              int mode, double factor, const float *nxyz, const uint32_t *off,
                              const uint32_t *idx, float *tmp, float *out, unsigned int *maxBits,
                              cudaStream_t s);
cudaError_t remapNeighbors(const uint32_t *s2o, const uint32_t *offO, const uint32_t *idxO,
                           uint32_t n, uint32_t *o2s, uint32_t *cnt, uint32_t *off, uint32_t *idx,
                           uint32_t *row, cudaStream_t s);

cudaError_t buildNeighborsDevice(int D, const float *pts, uint32_t n, const float lo[3],
                                 float distance, uint32_t *offOut, uint32_t **idxOut,
                                 size_t *totalOut, cudaStream_t s);

// flux of one particle: fixed point (ORIGINAL order, see vr_flux_device) -> float, optional SOURCE
// normalisation by areas[original id], optional neighbour smoothing, original order
cudaError_t postprocessFlux(const DeviceScene &sc, const unsigned long long *fixedOrig,
                            const uint32_t *s2o, const float *areas, float normFactor, int smooth,
                            float *tmpA, float *tmpB, float *outOrig, cudaStream_t s);

// sky map for rays that travel towards the source (see DeviceScene::sky).
// table: G*G float2, device.  top: host out.
cudaError_t buildSky(const DeviceScene &sc, int G, int upAxis, float upSign, int axisA, int axisB,
                     const float lo[2], const float hi[2], float2 *table, float *topOut,
                     cudaStream_t s);

// ---- kernels (vr_trace.cu) -------------------------------------------------
cudaError_t launchDiskBounds(const float4 *xyzr, const float4 *nrm, uint32_t n, float4 *lo,
                             float4 *hi, cudaStream_t s);
cudaError_t launchTriBounds(const float4 *v0, const float4 *v1, const float4 *v2, uint32_t n,
                            float4 *lo, float4 *hi, cudaStream_t s);
// one wavefront iteration = traverse + shade; init fills the pool
cudaError_t launchInitPool(const TraceParams &p, cudaStream_t s);
cudaError_t launchTraverse(const TraceParams &p, int numSMs, cudaStream_t s);
cudaError_t launchShade(const TraceParams &p, cudaStream_t s);
// the last rays of a trace, each run to its end by one thread (compacted pool in p.pool)
cudaError_t launchSpread(const TraceParams &p, int numSMs, cudaStream_t s);
cudaError_t launchTail(const TraceParams &p, cudaStream_t s);
// between iterations: resets the cursors; compact mode: slotCount = liveCount
cudaError_t launchFlip(unsigned int *ctrl, unsigned long long *counters, int compact,
                       cudaStream_t s);
cudaError_t launchDebugLoadRays(const DeviceScene &sc, const RayPool &pool, const float *rays,
                                uint32_t m, cudaStream_t s);
cudaError_t launchDebugReadHits(const DeviceScene &sc, const RayPool &pool, uint32_t m,
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
