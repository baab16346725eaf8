/* viennaray_b200 -- C ABI of the B200-native Monte Carlo flux tracer.
 *
 * This is the drop-in boundary for ViennaRay's hot path: everything that
 * rayInternal::TraceKernel<T,D,geo>::apply() does
 * (reference include/viennaray/rayTraceKernel.hpp:32-426) happens behind
 * vr_trace(); the scene set-up calls replace what the reference hands to
 * Embree.  Plain pointers and sizes only; all host arrays are owned by the
 * caller and copied by the library; every function returns 0 on success or a
 * VR_ERR_* code, with a message available from vr_last_error().
 *
 * There is no CPU fallback: without a CUDA device (or without the sm_100a
 * kernels in this library) vr_ctx_create fails with VR_ERR_CUDA.
 */
#ifndef VIENNARAY_B200_H
#define VIENNARAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_OK 0
#define VR_ERR_CUDA 1     /* no device / CUDA runtime error                  */
#define VR_ERR_ARGUMENT 2 /* null pointer, bad enum, inconsistent sizes      */
#define VR_ERR_STATE 3    /* call order (e.g. trace before commit)           */
#define VR_ERR_UNSUPPORTED 4

/* particle kinds: the reference's built-in particles as device functors
 * (rayParticle.hpp:124-204, rayReflection.hpp:13-120)                       */
#define VR_PARTICLE_DIFFUSE 0      /* DiffuseParticle                         */
#define VR_PARTICLE_SPECULAR 1     /* SpecularParticle                        */
#define VR_PARTICLE_CONED_COSINE 2 /* ReflectionConedCosine ion               */

/* boundary conditions, values of viennaray::BoundaryCondition
 * (rayBoundary.hpp:10-14)                                                   */
#define VR_BOUNDARY_REFLECTIVE 0
#define VR_BOUNDARY_PERIODIC 1
#define VR_BOUNDARY_IGNORE 2

/* flux is accumulated on the device as unsigned 64-bit fixed point,
 * weight * 2^30, so sums do not depend on accumulation order              */
#define VR_FLUX_FIXED_SCALE 1073741824.0

typedef struct vr_ctx vr_ctx;

/* replaces viennaray::Trace's RTCDevice (rayTrace.hpp:17,29) */
int vr_ctx_create(int cudaDevice, vr_ctx **out);
/* Several GPUs of one node behind one context: the scene calls are
 * replicated on every device, vr_trace* shards the ray-index range over them (one host
 * thread per device) and sums the result words with ONE ncclAllReduce over NVLink before
 * the download; the flux is bit-identical to a single device tracing the whole range.
 * NCCL is bound at run time (libnccl.so.2, or the path in VR_NCCL_LIB).  nDevices == 1 is
 * vr_ctx_create(deviceIds[0]).  Every other entry point takes the returned context as is. */
int vr_ctx_create_multi(int nDevices, const int *deviceIds, vr_ctx **out);
int vr_ctx_num_devices(const vr_ctx *ctx);
void vr_ctx_destroy(vr_ctx *ctx);
/* message of the last failing call on ctx (ctx may be NULL for create) */
const char *vr_last_error(const vr_ctx *ctx);

/* Disk geometry; replaces GeometryDisk::initGeometry's Embree buffers
 * (rayGeometryDisk.hpp:102-193,363-374).  xyzr: N x 4 {x,y,z,radius};
 * nxyz: N x 3 unit normals; nbOffsets (N+1) / nbIndices: neighbour lists in
 * CSR form (rayGeometryDisk.hpp:196-199, rayPointNeighborhood.hpp) -- see
 * vr_build_neighbors.  materialIds may be NULL (all 0). */
int vr_scene_set_disks(vr_ctx *ctx, const float *xyzr, const float *nxyz, uint32_t numDisks,
                       const int32_t *materialIds, const uint32_t *nbOffsets,
                       const uint32_t *nbIndices);

/* Neighbour lists built on the device for the disks set before (pass NULL
 * lists to vr_scene_set_disks): replaces PointNeighborhood::init
 * (rayPointNeighborhood.hpp:43-107,287-298) with the semantics of
 * vr_build_neighbors.  points: N x 3, the coordinates as given by the caller
 * (rayGeometryDisk.hpp:191).  vr_scene_get_neighbors returns malloc'ed copies
 * (release with vr_free). */
int vr_scene_build_neighbors(vr_ctx *ctx, int D, const float *points, float distance);
int vr_scene_get_neighbors(vr_ctx *ctx, uint32_t **offsetsOut, uint32_t **indicesOut);

/* Triangle geometry; replaces GeometryTriangle::initGeometry
 * (rayGeometryTriangle.hpp:15-92,246-254).  normals: N x 3 unit normals. */
int vr_scene_set_triangles(vr_ctx *ctx, const float *vertices, uint32_t numVertices,
                           const uint32_t *indices, uint32_t numTriangles, const float *normals,
                           const int32_t *materialIds);

/* Open box of 8 triangles around the adjusted bounding box; replaces
 * Boundary's constructor (rayBoundary.hpp:20-27,164-245).  condFirst /
 * condSecond: VR_BOUNDARY_* of the two lateral axes; D: 2 or 3. */
int vr_scene_set_boundary(vr_ctx *ctx, const float bboxMin[3], const float bboxMax[3],
                          int firstDir, int secondDir, int condFirst, int condSecond, int D);

/* Builds the acceleration structure on the device; replaces
 * rtcJoinCommitScene (rayTraceKernel.hpp:91). */
int vr_scene_commit(vr_ctx *ctx);

/* SourceRandom's state (raySourceRandom.hpp:14-23,118-129); cosine power is
 * taken from the particle.  basis: rows u,v,w of getOrthonormalBasis
 * (rayUtil.hpp:287-321), used when useBasis != 0. */
typedef struct {
  float bboxMin[3], bboxMax[3];
  int32_t rayDir, firstDir, secondDir, minMax;
  float posNeg;
  int32_t useBasis;
  float basis[9];
  /* != 0: SourceGrid (raySourceGrid.hpp:9-74) over the origins given to
   * vr_source_set_grid; ray idx starts at origin idx % numPoints */
  int32_t useGrid;
} vr_source_desc;

/* origins (n x 3) of the grid source; copied to the device.  n == 0 releases them. */
int vr_source_set_grid(vr_ctx *ctx, const float *points, uint32_t n);

typedef struct {
  int32_t kind;       /* VR_PARTICLE_*                                       */
  float sticking;     /* constant sticking probability                       */
  float sourcePower;  /* getSourceDistributionPower()                        */
  float coneMinAngle; /* coned cosine: cone = pi/2 - min(incAngle, this)     */
  float meanFreePath; /* getMeanFreePath(); <= 0: no scattering
                         (rayTraceKernel.hpp:179-203)                       */
  /* Optional sticking probability per material: the kernel hands the materialId of the hit
   * primitive to surfaceReflection (rayTraceKernel.hpp:310-313, rayParticle.hpp:44-48) and
   * ViennaPS particles switch on it.  stickingByMaterial[m] replaces `sticking` when the hit
   * primitive's materialId m (vr_scene_set_disks / vr_scene_set_triangles) lies in
   * [0, numMaterials); host array, copied by vr_trace*.  NULL: constant sticking. */
  const float *stickingByMaterial;
  int32_t numMaterials;
} vr_particle_desc;

/* KernelConfig (rayUtil.hpp:83-94) + the ray-index shard of this context */
typedef struct {
  uint64_t numRays;      /* total rays of the job (all shards)               */
  uint64_t rayIdxBegin;  /* this context traces idx in [begin, end)          */
  uint64_t rayIdxEnd;
  uint32_t seed;         /* runNumber + rngSeed (rayTraceKernel.hpp:100)     */
  uint32_t maxReflections;
  uint32_t maxBoundaryHits;
  uint32_t flags;        /* VR_FLAG_*                                        */
} vr_config;

/* distance-weighted neighbour spread, the reference's compile-time option
 * VIENNARAY_USE_WDIST (rayTraceKernel.hpp:258-296) */
#define VR_FLAG_WDIST 1u

/* TraceInfo (rayUtil.hpp:65-76) plus rays cut by the hit limits */
typedef struct {
  uint64_t numRays, totalRaysTraced, nonGeometryHits, geometryHits, particleHits, boundaryHits,
      reflections, raysTerminated;
  double time; /* seconds of device time, CUDA events around the kernels */
} vr_trace_info;

/* The hot path.  Traces every particle over the shard and returns
 * fluxOut[p * N + i] = sum of ray weights collected by primitive i (original
 * primitive order, the reference's localData vector 0), infoOut[p].
 * Blocks until the results are on the host. */
int vr_trace(vr_ctx *ctx, const vr_source_desc *source, const vr_particle_desc *particles,
             int numParticles, const vr_config *config, double *fluxOut, vr_trace_info *infoOut);

/* Same launch, results left on the device (multi-GPU: all-reduce the buffer
 * returned by vr_flux_device, then vr_flux_download).  Asynchronous on the
 * context's stream unless `sync` is non-zero. */
int vr_trace_device(vr_ctx *ctx, const vr_source_desc *source, const vr_particle_desc *particles,
                    int numParticles, const vr_config *config, int sync);
/* device pointer to numParticles x N uint64 fixed-point sums in the CALLER's primitive order
 * (so that ranks can sum them whatever BVH each one built) followed by numParticles x 8
 * uint64 counters; vr_flux_download / vr_flux_postprocess read this buffer */
int vr_flux_device(vr_ctx *ctx, void **devicePtr, size_t *numWords);
int vr_flux_download(vr_ctx *ctx, double *fluxOut, vr_trace_info *infoOut);
int vr_flux_download_fixed(vr_ctx *ctx, uint64_t *fluxOut);
/* Post-processing of the last trace on the device, float flux of one particle
 * in the caller's primitive order: SOURCE normalisation
 * flux[i] *= normFactor / areas[i] (rayTraceDisk.hpp:121-138,
 * rayTraceTriangle.hpp:110-126; areas NULL = raw sums) followed, when smooth
 * != 0 and the geometry is disks, by smoothFlux over the geometry's own
 * neighbourhood (rayTraceDisk.hpp:146-193, numNeighbors == 1). */
int vr_flux_postprocess(vr_ctx *ctx, int particle, const float *areas, float normFactor,
                        int smooth, float *fluxOut);
/* The general form: normalizeFlux(SOURCE | MAX) followed by smoothFlux(numNeighbors), all on
 * the device, float flux of one particle in the caller's primitive order.
 *   VR_NORM_SOURCE  flux[i] *= float(normFactor) / areas[i], normFactor = sourceArea / numRays
 *                   (rayTraceDisk.hpp:121-138, rayTraceTriangle.hpp:110-126)
 *   VR_NORM_MAX     disks: flux[i] = flux[i] * ((normFactor / areas[i]) / max(flux)) in double,
 *                   normFactor = radius * radius * pi (rayTraceDisk.hpp:110-118; the argument is
 *                   ignored for triangles: flux[i] /= max(flux) * areas[i], rayTraceTriangle.hpp:
 *                   99-107)
 *   smoothNeighbors 0: none; 1: the geometry's own neighbourhood; k > 1: a neighbourhood of
 *                   k * 2 * diskRadius over the disk centres, built on the device
 *                   (rayTraceDisk.hpp:146-193).  Disks only; ignored for triangles. */
enum { VR_NORM_NONE = 0, VR_NORM_SOURCE = 1, VR_NORM_MAX = 2 };
int vr_flux_postprocess_ex(vr_ctx *ctx, int particle, const float *areas, int normalization,
                           double normFactor, int smoothNeighbors, float diskRadius,
                           float *fluxOut);
/* cudaStream_t of the context (for event timing by the caller) */
void *vr_ctx_stream(vr_ctx *ctx);
int vr_ctx_synchronize(vr_ctx *ctx);
/* device time in ms of the trace kernels of the last vr_trace* call */
float vr_last_kernel_ms(vr_ctx *ctx);
/* kernels launched / wavefront iterations run by the last vr_trace* call */
int vr_last_launch_count(vr_ctx *ctx, int *kernelsOut, int *iterationsOut);

/* ---- host-side geometry helpers (C++ inside the library) ---------------- */
/* Neighbour sets of PointNeighborhood::init (rayPointNeighborhood.hpp:43-107,
 * 287-298): j != i with |dx_a| <= distance on the first D axes and
 * |p_i - p_j|^2 <= distance^2 (float).  points: N x 3.  Rows ascending.
 * Arrays are malloc'ed; release with vr_free. */
int vr_build_neighbors(int D, const float *points, uint32_t numPoints, float distance,
                       uint32_t **offsetsOut, uint32_t **indicesOut);
void vr_free(void *p);

/* ---- parity / debugging entry points ------------------------------------ */
/* closest hit (boundary geomID 0, geometry geomID 1) and, for disks, the
 * neighbour hit set of checkLocalIntersection (rayTraceKernel.hpp:462-507)
 * for caller-given rays (m x 6: origin, direction), tnear = 1e-4.  IDs are
 * original primitive IDs, 0xffffffff on a miss. */
int vr_debug_intersect(vr_ctx *ctx, const float *rays, uint32_t m, uint32_t *geomOut,
                       uint32_t *primOut, float *tOut, uint32_t nbCap, uint32_t *nbCountOut,
                       uint32_t *nbOut);
/* origin + ray direction of rays idxBegin .. idxBegin+m-1 (m x 6) */
int vr_debug_source_rays(vr_ctx *ctx, const vr_source_desc *source,
                         const vr_particle_desc *particle, const vr_config *config,
                         uint64_t idxBegin, uint32_t m, float *raysOut);
/* which: 0 sincos2pi (out = 2m floats: sin then cos), 1 pow(x, param),
 * 2 acos(x) */
int vr_debug_math(vr_ctx *ctx, int which, const float *x, uint32_t m, float param, float *out);
int vr_debug_philox(vr_ctx *ctx, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint32_t *out4);
/* m reflections of (rayDir, normal), ray streams idx .. idx+m-1 */
int vr_debug_reflect(vr_ctx *ctx, int kind, int D, const float *rayDir, const float *normal,
                     float coneMinAngle, uint32_t seed, uint64_t idx, uint32_t m, float *out3);
/* acceleration-structure statistics: out[0] nodes, out[1] leaves, out[2]
 * max leaf size, out[3] node bytes, out[4] build ms (float bits), out[5] / out[6]
 * surface-area-heuristic terms (float bits): sum of the inner-node areas and of
 * the leaf areas x primitive counts, over the root area, out[7] the Morton cell
 * shape of the tree that was kept (float bits; 1 = like the scene box, 0 = cubic) */
int vr_debug_bvh_stats(vr_ctx *ctx, uint64_t *out8);
/* per-phase device time: with timing enabled every kernel launch of vr_trace*
 * is bracketed by CUDA events on the context's stream; vr_debug_phase_ms
 * returns the milliseconds and launch counts accumulated since it was enabled
 * for {traverse kernel, shade kernel, everything else} */
int vr_debug_phase_timing(vr_ctx *ctx, int enable);
int vr_debug_phase_ms(vr_ctx *ctx, double *ms3, int64_t *launches3);
/* per-ray traversal work counters of the last vr_trace* call when the
 * context was created with VR_COUNT_WORK=1 in the environment: out[0] node
 * visits, out[1] primitive tests, out[2] neighbour tests, out[3] flux adds,
 * out[4] rays finished by the sky map without a traversal */
int vr_debug_work_counters(vr_ctx *ctx, uint64_t *out5);
/* measured read bandwidth of an L2-resident buffer on this device (the
 * scene of the 1M-disk trench lives in L2, so this, not HBM, is the level its fetches
 * come from): `bytes` are read `passes` times by a streaming kernel after one warm-up
 * pass; out2[0] = GB/s, out2[1] = cudaDeviceProp::l2CacheSize in bytes */
int vr_debug_l2_read_bandwidth(vr_ctx *ctx, uint64_t bytes, int passes, double *out2);

#ifdef __cplusplus
}
#endif
#endif
