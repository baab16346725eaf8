/* TEST INFRASTRUCTURE (oracle/): CPU restatement, in plain C, of ViennaRay's
 * Monte Carlo flux hot path (rayInternal::TraceKernel::apply,
 * /root/reference/include/viennaray/rayTraceKernel.hpp:32-426) and of the
 * host-side set-up it consumes.  It is the CHECKER for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * it.  The product library never links, includes or calls anything here.
 *
 * Parity status.  Pinned against the reference's own golden vectors
 * (tests/intersectionTest, boundaryHit, boundaryHit2D, createRay,
 * pointNeighborhood*, diskAreas) and, statistically, against the reference's
 * unmodified TraceKernel compiled by oracle/Makefile into oracle/_ref.
 * UNPINNED at the Embree boundary: Embree 4.3.3 is not in the reference tree
 * nor installed, so closest-hit ties between coplanar overlapping disks follow
 * the explicit rule in oracle/mini_rtc.cpp's header, not Embree's traversal
 * order.  The random stream is a counter-based Philox4x32-10 keyed on
 * (seed, ray index), as BASELINE.json's north_star prescribes for the device
 * path, not the reference's mt19937_64 -- per-disk flux parity with the
 * reference is therefore statistical, while parity between this oracle and
 * the CUDA kernels is bit-exact (all float arithmetic unfused, in the order
 * written here; transcendental functions are the polynomial forms below).
 */
#ifndef VR_ORACLE_H
#define VR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vro_scene vro_scene;

enum { VRO_DIFFUSE = 0, VRO_SPECULAR = 1, VRO_CONED_COSINE = 2 };
enum { VRO_REFLECTIVE = 0, VRO_PERIODIC = 1, VRO_IGNORE = 2 };

/* flux is accumulated in unsigned 64-bit fixed point, weight * 2^30, so the
 * sum is independent of accumulation order (bitwise reproducible). */
#define VRO_FLUX_SCALE 1073741824.0

typedef struct {
  int kind;           /* VRO_DIFFUSE | VRO_SPECULAR | VRO_CONED_COSINE */
  float sticking;     /* constant sticking probability */
  float sourcePower;  /* cosine exponent of the source (1 = cosine) */
  float coneMinAngle; /* coned cosine: cone = pi/2 - min(incAngle, this) */
  float meanFreePath; /* getMeanFreePath(); <= 0: no scattering (rayTraceKernel.hpp:179) */
  /* optional sticking per materialId of the hit primitive (the materialId argument of
   * surfaceReflection, rayParticle.hpp:44-48); NULL: constant sticking */
  const float *stickingByMaterial;
  int numMaterials;
} vro_particle;

typedef struct {
  uint64_t numRays;
  uint32_t seed; /* runNumber + rngSeed, rayTraceKernel.hpp:100 */
  uint32_t stream; /* second key word (particle index in a multi-particle run) */
  uint32_t maxReflections;
  uint32_t maxBoundaryHits;
  int usePrimaryDir;
  float primaryDir[3];
  int useWdist; /* VIENNARAY_USE_WDIST: distance-weighted neighbour spread */
} vro_config;

typedef struct {
  uint64_t numRays, totalTraces, nonGeoHits, geoHits, particleHits, boundaryHits, reflections,
      raysTerminated;
} vro_info;

int vro_set_threads(int n); /* n <= 0: query only; returns the OpenMP thread count in effect */
vro_scene *vro_scene_create(int D);
void vro_scene_destroy(vro_scene *s);
int vro_scene_set_disks(vro_scene *s, const float *points, const float *normals, uint32_t n,
                        float radius);
int vro_scene_set_triangles(vro_scene *s, const float *verts, uint32_t nVerts,
                            const uint32_t *tris, uint32_t n);
/* sourceOffset: disk radius (rayTraceDisk.hpp:21-23) or gridDelta
 * (rayTraceTriangle.hpp:21-23) */
int vro_scene_set_material_ids(vro_scene *s, const int *ids /* n, or NULL */);
int vro_scene_setup(vro_scene *s, int sourceDir, const int *bc, float sourceOffset);
/* SourceGrid (raySourceGrid.hpp:9-74): origins points[idx % n]; n == 0: random source */
int vro_scene_set_source_grid(vro_scene *s, const float *points, uint32_t n);
void vro_scene_bbox(const vro_scene *s, float *out6);
uint32_t vro_scene_num_prims(const vro_scene *s);
void vro_scene_neighbors(const vro_scene *s, const uint32_t **offsets, const uint32_t **indices);
const float *vro_scene_normals(const vro_scene *s); /* n x 3 */

int vro_trace(const vro_scene *s, const vro_particle *p, const vro_config *c, uint64_t idxBegin,
              uint64_t idxEnd, uint64_t *fluxFixed, vro_info *info);

int vro_source_rays(const vro_scene *s, const vro_particle *p, const vro_config *c,
                    uint64_t idxBegin, uint32_t m, float *rays6);
int vro_intersect(const vro_scene *s, const float *rays6, uint32_t m, uint32_t *geom,
                  uint32_t *prim, float *t, float *ng3);
int vro_neighbor_hits(const vro_scene *s, const float *rays6, const uint32_t *prim, uint32_t m,
                      uint32_t cap, uint32_t *count, uint32_t *out);
/* Boundary::processHit restated (rayBoundary.hpp:29-127); returns reflect */
int vro_boundary_process_hit(const vro_scene *s, float *org, float *rayDir3, float *dir,
                             const float *ng, uint32_t primID, float t);

void vro_normalize_flux_source(const vro_scene *s, const float *areas, uint64_t numRays,
                               float *flux);
void vro_smooth_flux(const vro_scene *s, float *flux);
/* normalizeFlux(MAX), rayTraceDisk.hpp:110-118 / rayTraceTriangle.hpp:99-107 */
void vro_normalize_flux_max(const vro_scene *s, const float *areas, float *flux);
/* smoothFlux(flux, k), rayTraceDisk.hpp:146-193 (k > 1: neighbourhood of k * 2 * radius) */
void vro_smooth_flux_k(const vro_scene *s, int k, float *flux);

/* building blocks exposed for bit-parity unit tests */
void vro_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t *out4);
void vro_math_sincos2pi(const float *x, uint32_t m, float *s, float *c);
void vro_math_pow(const float *x, float e, uint32_t m, float *out);
void vro_math_acos(const float *x, uint32_t m, float *out);
void vro_reflect(int kind, int D, const float *rayDir, const float *normal, float coneMinAngle,
                 uint32_t seed, uint64_t idx, uint32_t m, float *out3);

#ifdef __cplusplus
}
#endif
#endif
