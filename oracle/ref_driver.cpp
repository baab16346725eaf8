// TEST INFRASTRUCTURE (oracle/) -- never linked into the product library.
//
// ref_driver: a C-callable shim around the reference's OWN, UNMODIFIED headers
// (included from /root/reference/include/viennaray at build time, see
// oracle/Makefile) so that tests and bench.py's cpu_baseline / `--impl
// reference` arm can run the reference TraceKernel
// (rayTraceKernel.hpp:32-426) through its public Trace API
// (rayTraceDisk.hpp:19-57, rayTraceTriangle.hpp:19-61).  The Embree symbols are
// supplied by oracle/mini_rtc.cpp ("reference kernel + substitute
// intersector" -- NOT Embree); ViennaCore by oracle/stubs/.
//
// Nothing here restates reference logic except IonParticle, a user-side
// particle (the reference ships no coned-cosine particle class) that calls the
// reference's ReflectionConedCosine (rayReflection.hpp:52-120) with the cone
// recipe of tests/reflection/reflection.cpp:43-46.

#include <rayParticle.hpp>
#include <rayTraceDisk.hpp>
#include <rayTraceTriangle.hpp>

#include <cstdint>
#include <cstring>
#include <omp.h>

using namespace viennaray;

namespace {

template <typename T, int D> class IonParticle : public Particle<IonParticle<T, D>, T> {
  const T sticking_, power_, minAngle_;
  const std::string label_;

public:
  IonParticle(T sticking, T power, T minAngle, std::string label)
      : sticking_(sticking), power_(power), minAngle_(minAngle), label_(std::move(label)) {}

  std::pair<T, Vec3D<T>> surfaceReflection(T, const Vec3D<T> &rayDir, const Vec3D<T> &geomNormal,
                                           const unsigned int, const int,
                                           const TracingData<T> *, RNG &rng) final {
    T cosTheta = -DotProduct(rayDir, geomNormal);
    cosTheta = std::min(std::max(cosTheta, T(0)), T(1));
    const T incAngle = std::acos(cosTheta);
    const T cone = T(M_PI_2) - std::min(incAngle, minAngle_);
    auto dir = ReflectionConedCosine<T, D>(rayDir, geomNormal, rng, cone);
    return {sticking_, dir};
  }
  void surfaceCollision(T rayWeight, const Vec3D<T> &, const Vec3D<T> &, const unsigned int primID,
                        const int, TracingData<T> &localData, const TracingData<T> *,
                        RNG &) final {
    localData.getVectorData(0)[primID] += rayWeight;
  }
  T getSourceDistributionPower() const final { return power_; }
  std::vector<std::string> getLocalDataLabels() const final { return {label_}; }
};

struct ParticleDesc {
  int kind; // 0 diffuse, 1 specular, 2 coned-cosine ion
  float sticking, sourcePower, coneMinAngle;
};

// Options of the next ref_trace_* calls (ref_set_options): a mean free path > 0 and / or a
// sticking table indexed by the materialId the kernel hands to surfaceReflection
// (rayTraceKernel.hpp:310-313), and the material IDs given to the geometry
// (rayTraceDisk.hpp:96-98).  They select TestParticle below instead of a built-in particle.
struct Options {
  float meanFreePath = -1.f;
  std::vector<float> stickingByMaterial;
  std::vector<int> materialIds;
} g_opt;

// A user-side particle, as ViennaPS writes them: the reference's own reflection functions
// (rayReflection.hpp), a sticking probability looked up by materialId, and a mean free path
// (rayParticle.hpp:72-74).  No reference logic is restated here.
template <typename T, int D> class TestParticle : public Particle<TestParticle<T, D>, T> {
  const int kind_;
  const T sticking_, power_, minAngle_, meanFreePath_;
  const std::vector<float> table_;

public:
  TestParticle(const ParticleDesc &p, const Options &o)
      : kind_(p.kind), sticking_(p.sticking), power_(p.sourcePower), minAngle_(p.coneMinAngle),
        meanFreePath_(o.meanFreePath), table_(o.stickingByMaterial) {}

  std::pair<T, Vec3D<T>> surfaceReflection(T, const Vec3D<T> &rayDir, const Vec3D<T> &geomNormal,
                                           const unsigned int, const int materialId,
                                           const TracingData<T> *, RNG &rng) final {
    Vec3D<T> dir;
    if (kind_ == 0) {
      dir = ReflectionDiffuse<T, D>(geomNormal, rng);
    } else if (kind_ == 1) {
      dir = ReflectionSpecular<T, D>(rayDir, geomNormal);
    } else {
      T cosTheta = -DotProduct(rayDir, geomNormal);
      cosTheta = std::min(std::max(cosTheta, T(0)), T(1));
      const T cone = T(M_PI_2) - std::min(T(std::acos(cosTheta)), minAngle_);
      dir = ReflectionConedCosine<T, D>(rayDir, geomNormal, rng, cone);
    }
    T st = sticking_;
    if (materialId >= 0 && materialId < (int)table_.size())
      st = table_[materialId];
    return {st, dir};
  }
  void surfaceCollision(T rayWeight, const Vec3D<T> &, const Vec3D<T> &, const unsigned int primID,
                        const int, TracingData<T> &localData, const TracingData<T> *,
                        RNG &) final {
    localData.getVectorData(0)[primID] += rayWeight;
  }
  T getSourceDistributionPower() const final { return power_; }
  T getMeanFreePath() const final { return meanFreePath_; }
  std::vector<std::string> getLocalDataLabels() const final { return {"flux"}; }
};

template <int D> std::unique_ptr<AbstractParticle<float>> makeParticle(const ParticleDesc &p) {
  // the reference's DiffuseParticle has a fixed cosine source (rayParticle.hpp:158): a
  // diffuse particle with another source power (config C5) is a user-side particle too
  if (p.kind >= 0 && p.kind <= 2 &&
      (g_opt.meanFreePath > 0.f || !g_opt.stickingByMaterial.empty() ||
       (p.kind == 0 && p.sourcePower != 1.f)))
    return std::make_unique<TestParticle<float, D>>(p, g_opt);
  switch (p.kind) {
  case 0:
    return std::make_unique<DiffuseParticle<float, D>>(p.sticking, "flux");
  case 1:
    return std::make_unique<SpecularParticle<float, D>>(p.sticking, p.sourcePower, "flux");
  case 2:
    return std::make_unique<IonParticle<float, D>>(p.sticking, p.sourcePower, p.coneMinAngle,
                                                   "flux");
  }
  return nullptr;
}

void fillInfo(const TraceInfo &ti, uint64_t *info, double *seconds) {
  if (info) {
    info[0] = ti.numRays;
    info[1] = ti.totalRaysTraced;
    info[2] = ti.nonGeometryHits;
    info[3] = ti.geometryHits;
    info[4] = ti.particleHits;
    info[5] = ti.boundaryHits;
    info[6] = ti.reflections;
    info[7] = (ti.error ? 1 : 0) | (ti.warning ? 2 : 0);
  }
  if (seconds)
    *seconds = ti.time;
}

template <int D>
int traceDisk(const float *points, const float *normals, uint32_t n, float gridDelta,
              const int *bc, int sourceDir, const ParticleDesc &pd, uint64_t raysPerPoint,
              uint64_t raysFixed, unsigned seed, unsigned runs, const float *primaryDir,
              int normalize, int smooth, float *fluxOut, uint64_t *info, double *seconds) {
  std::vector<Vec3D<float>> pts(n), nrm(n);
  for (uint32_t i = 0; i < n; ++i) {
    pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    nrm[i] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
  }
  BoundaryCondition conds[D];
  for (int i = 0; i < D; ++i)
    conds[i] = static_cast<BoundaryCondition>(bc[i]);
  auto particle = makeParticle<D>(pd);
  if (!particle)
    return 2;
  TraceDisk<float, D> tracer;
  tracer.setGeometry(pts, nrm, gridDelta);
  if (g_opt.materialIds.size() == n)
    tracer.setMaterialIds(g_opt.materialIds);
  tracer.setBoundaryConditions(conds);
  tracer.setSourceDirection(static_cast<TraceDirection>(sourceDir));
  tracer.setParticleType(particle);
  if (raysFixed)
    tracer.setNumberOfRaysFixed(raysFixed);
  else
    tracer.setNumberOfRaysPerPoint(raysPerPoint);
  tracer.setRngSeed(seed);
  if (primaryDir)
    tracer.setPrimaryDirection(Vec3D<float>{primaryDir[0], primaryDir[1], primaryDir[2]});
  // `runs` consecutive apply() calls: config_.runNumber increments per call
  // (rayTraceDisk.hpp:54) so each run draws a fresh stream; fluxOut is
  // runs x n (independent repeats for the 3-sigma test).
  double total = 0;
  for (unsigned r = 0; r < runs; ++r) {
    tracer.apply();
    auto flux = tracer.getLocalData().getVectorData(0);
    if (normalize)
      tracer.normalizeFlux(flux, NormalizationType::SOURCE);
    if (smooth > 0)
      tracer.smoothFlux(flux, smooth);
    std::memcpy(fluxOut + size_t(r) * n, flux.data(), sizeof(float) * n);
    auto ti = tracer.getRayTraceInfo();
    total += ti.time;
    fillInfo(ti, info, nullptr);
  }
  if (seconds)
    *seconds = total;
  return 0;
}

// The reference's own post-processing of a GIVEN flux vector: TraceDisk::normalizeFlux(norm)
// and smoothFlux(k) (rayTraceDisk.hpp:103-193) after a short apply() that initialises the
// geometry (disk areas, neighbourhood).  norm: 0 none, 1 SOURCE, 2 MAX.
template <int D>
int postDisk(const float *points, const float *normals, uint32_t n, float gridDelta, const int *bc,
             int sourceDir, uint64_t raysFixed, int norm, int smooth, float *flux) {
  std::vector<Vec3D<float>> pts(n), nrm(n);
  for (uint32_t i = 0; i < n; ++i) {
    pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    nrm[i] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
  }
  BoundaryCondition conds[D];
  for (int i = 0; i < D; ++i)
    conds[i] = static_cast<BoundaryCondition>(bc[i]);
  auto particle = makeParticle<D>(ParticleDesc{0, 1.f, 1.f, 0.f});
  TraceDisk<float, D> tracer;
  tracer.setGeometry(pts, nrm, gridDelta);
  tracer.setBoundaryConditions(conds);
  tracer.setSourceDirection(static_cast<TraceDirection>(sourceDir));
  tracer.setParticleType(particle);
  tracer.setNumberOfRaysFixed(raysFixed);
  tracer.setRngSeed(1);
  tracer.apply();
  std::vector<float> f(flux, flux + n);
  if (norm == 1)
    tracer.normalizeFlux(f, NormalizationType::SOURCE);
  else if (norm == 2)
    tracer.normalizeFlux(f, NormalizationType::MAX);
  if (smooth > 0)
    tracer.smoothFlux(f, smooth);
  std::memcpy(flux, f.data(), sizeof(float) * n);
  return 0;
}

} // namespace

extern "C" {

// flux (n floats) in and out: the reference's normalizeFlux / smoothFlux on it
int ref_post_disk(int D, const float *points, const float *normals, uint32_t n, float gridDelta,
                  const int *bc, int sourceDir, uint64_t raysFixed, int norm, int smooth,
                  float *flux) {
  if (D == 3)
    return postDisk<3>(points, normals, n, gridDelta, bc, sourceDir, raysFixed, norm, smooth, flux);
  if (D == 2)
    return postDisk<2>(points, normals, n, gridDelta, bc, sourceDir, raysFixed, norm, smooth, flux);
  return 1;
}

int ref_max_threads() { return omp_get_max_threads(); }
void ref_set_threads(int n) { omp_set_num_threads(n); }
// 1 when this library was compiled with -DVIENNARAY_USE_WDIST (oracle/Makefile, ref_wdist)
int ref_uses_wdist() {
#ifdef VIENNARAY_USE_WDIST
  return 1;
#else
  return 0;
#endif
}
// Options of the following ref_trace_* calls; all-default arguments switch them off again.
// meanFreePath <= 0: none.  stickingByMaterial: numMaterials floats or NULL.  materialIds:
// numIds ints (must equal the primitive count of the traced geometry) or NULL.
void ref_set_options(float meanFreePath, const float *stickingByMaterial, int numMaterials,
                     const int *materialIds, uint32_t numIds) {
  g_opt.meanFreePath = meanFreePath > 0.f ? meanFreePath : -1.f;
  g_opt.stickingByMaterial.assign(stickingByMaterial ? stickingByMaterial : nullptr,
                                  stickingByMaterial ? stickingByMaterial + numMaterials : nullptr);
  g_opt.materialIds.assign(materialIds ? materialIds : nullptr,
                           materialIds ? materialIds + numIds : nullptr);
}

// kind: see ParticleDesc.  bc: BoundaryCondition per axis (rayBoundary.hpp:10-14).
// sourceDir: TraceDirection (rayUtil.hpp:38-45).  fluxOut: runs x n floats.
// info: 8 x uint64 (numRays, totalRaysTraced, nonGeometryHits, geometryHits,
// particleHits, boundaryHits, reflections, flags) of the LAST run.
int ref_trace_disk(int D, const float *points, const float *normals, uint32_t n, float gridDelta,
                   const int *bc, int sourceDir, int kind, float sticking, float sourcePower,
                   float coneMinAngle, uint64_t raysPerPoint, uint64_t raysFixed, unsigned seed,
                   unsigned runs, const float *primaryDir, int normalize, int smooth,
                   float *fluxOut, uint64_t *info, double *seconds) {
  ParticleDesc pd{kind, sticking, sourcePower, coneMinAngle};
  if (D == 3)
    return traceDisk<3>(points, normals, n, gridDelta, bc, sourceDir, pd, raysPerPoint, raysFixed,
                        seed, runs, primaryDir, normalize, smooth, fluxOut, info, seconds);
  if (D == 2)
    return traceDisk<2>(points, normals, n, gridDelta, bc, sourceDir, pd, raysPerPoint, raysFixed,
                        seed, runs, primaryDir, normalize, smooth, fluxOut, info, seconds);
  return 1;
}

int ref_trace_triangle(const float *verts, uint32_t nVerts, const uint32_t *tris, uint32_t n,
                       float gridDelta, const int *bc, int sourceDir, int kind, float sticking,
                       float sourcePower, float coneMinAngle, uint64_t raysPerPoint,
                       uint64_t raysFixed, unsigned seed, unsigned runs, int normalize,
                       float *fluxOut, uint64_t *info, double *seconds) {
  constexpr int D = 3;
  std::vector<Vec3Df> nodes(nVerts);
  std::vector<Vec3D<unsigned>> elems(n);
  for (uint32_t i = 0; i < nVerts; ++i)
    nodes[i] = {verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]};
  for (uint32_t i = 0; i < n; ++i)
    elems[i] = {tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]};
  TriangleMesh mesh(nodes, elems, gridDelta); // as examples/triangle3D/triangle3D.cpp:24-26
  BoundaryCondition conds[D];
  for (int i = 0; i < D; ++i)
    conds[i] = static_cast<BoundaryCondition>(bc[i]);
  ParticleDesc pd{kind, sticking, sourcePower, coneMinAngle};
  auto particle = makeParticle<D>(pd);
  if (!particle)
    return 2;
  TraceTriangle<float, D> tracer;
  tracer.setGeometry(mesh);
  if (g_opt.materialIds.size() == n)
    tracer.setMaterialIds(g_opt.materialIds);
  tracer.setBoundaryConditions(conds);
  tracer.setSourceDirection(static_cast<TraceDirection>(sourceDir));
  tracer.setParticleType(particle);
  if (raysFixed)
    tracer.setNumberOfRaysFixed(raysFixed);
  else
    tracer.setNumberOfRaysPerPoint(raysPerPoint);
  tracer.setRngSeed(seed);
  double total = 0;
  for (unsigned r = 0; r < runs; ++r) {
    tracer.apply();
    auto flux = tracer.getLocalData().getVectorData(0);
    if (normalize)
      tracer.normalizeFlux(flux, NormalizationType::SOURCE);
    std::memcpy(fluxOut + size_t(r) * n, flux.data(), sizeof(float) * n);
    auto ti = tracer.getRayTraceInfo();
    total += ti.time;
    fillInfo(ti, info, nullptr);
  }
  if (seconds)
    *seconds = total;
  return 0;
}

// Reference PointNeighborhood (rayPointNeighborhood.hpp:43-107) on float
// points; bbox computed as GeometryDisk::initGeometry does
// (rayGeometryDisk.hpp:131-159).  counts: n; indices: row i occupies
// [i*cap, i*cap+counts[i]).  Returns the max row length (may exceed cap, in
// which case rows are truncated).
uint32_t ref_neighbors(int D, const float *points, uint32_t n, float distance, uint32_t cap,
                       uint32_t *counts, uint32_t *indices) {
  std::vector<Vec3D<float>> pts(n);
  Vec3D<float> mn{0, 0, 0}, mx{0, 0, 0};
  for (int a = 0; a < D; ++a) {
    mn[a] = std::numeric_limits<float>::max();
    mx[a] = std::numeric_limits<float>::lowest();
  }
  for (uint32_t i = 0; i < n; ++i) {
    pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    for (int a = 0; a < D; ++a) {
      mn[a] = std::min(mn[a], pts[i][a]);
      mx[a] = std::max(mx[a], pts[i][a]);
    }
  }
  uint32_t maxLen = 0;
  auto fill = [&](auto &nb) {
    for (uint32_t i = 0; i < n; ++i) {
      auto const &row = nb.getNeighborIndices(i);
      counts[i] = (uint32_t)row.size();
      maxLen = std::max(maxLen, counts[i]);
      for (uint32_t k = 0; k < row.size() && k < cap; ++k)
        indices[size_t(i) * cap + k] = row[k];
    }
  };
  if (D == 3) {
    PointNeighborhood<float, 3> nb;
    nb.template init<3>(pts, distance, mn, mx);
    fill(nb);
  } else {
    PointNeighborhood<float, 2> nb;
    nb.template init<3>(pts, distance, mn, mx);
    fill(nb);
  }
  return maxLen;
}

// Reference disk areas (rayGeometryDisk.hpp:266-354) for the given trace setup.
int ref_disk_areas(int D, const float *points, const float *normals, uint32_t n, float gridDelta,
                   const int *bc, int sourceDir, float *areasOut) {
  auto run = [&](auto dimTag) {
    constexpr int DD = decltype(dimTag)::value;
    std::vector<Vec3D<float>> pts(n), nrm(n);
    for (uint32_t i = 0; i < n; ++i) {
      pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
      nrm[i] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
    }
    auto device = rtcNewDevice("");
    float radius = gridDelta * rayInternal::DiskFactor<DD>;
    GeometryDisk<float, DD> geo;
    geo.initGeometry(device, pts, nrm, radius);
    auto bbox = geo.getBoundingBox();
    auto dir = static_cast<TraceDirection>(sourceDir);
    rayInternal::adjustBoundingBox<float, DD>(bbox, dir, radius);
    auto settings = rayInternal::getTraceSettings(dir);
    BoundaryCondition conds[3];
    for (int i = 0; i < 3; ++i)
      conds[i] = static_cast<BoundaryCondition>(bc[i]);
    Boundary<float, DD> boundary(device, bbox, conds, settings);
    geo.computeDiskAreas(boundary);
    for (uint32_t i = 0; i < n; ++i)
      areasOut[i] = geo.getDiskArea(i);
    boundary.releaseGeometry();
    geo.releaseGeometry();
    rtcReleaseDevice(device);
  };
  if (D == 3)
    run(std::integral_constant<int, 3>{});
  else
    run(std::integral_constant<int, 2>{});
  return 0;
}

// Scene exactly as TraceKernel::apply attaches it (boundary = geomID 0,
// geometry = geomID 1; rayTraceKernel.hpp:41-45), then rtcIntersect1 per ray
// with tnear = 1e-4 (rayUtil.hpp:218).  rays: m x 6 (org, dir).  Outputs per
// ray: geomID, primID (0xffffffff on miss), t, Ng[3].
int ref_intersect_disks(const float *points, const float *normals, uint32_t n, float radius,
                        const float *bboxMin, const float *bboxMax, int sourceDir,
                        const float *rays, uint32_t m, uint32_t *geomOut, uint32_t *primOut,
                        float *tOut, float *ngOut) {
  std::vector<Vec3D<float>> pts(n), nrm(n);
  for (uint32_t i = 0; i < n; ++i) {
    pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    nrm[i] = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
  }
  auto device = rtcNewDevice("");
  GeometryDisk<float, 3> geo;
  geo.initGeometry(device, pts, nrm, radius);
  std::array<Vec3D<float>, 2> bbox{Vec3D<float>{bboxMin[0], bboxMin[1], bboxMin[2]},
                                   Vec3D<float>{bboxMax[0], bboxMax[1], bboxMax[2]}};
  auto settings = rayInternal::getTraceSettings(static_cast<TraceDirection>(sourceDir));
  BoundaryCondition conds[3] = {};
  Boundary<float, 3> boundary(device, bbox, conds, settings);
  auto scene = rtcNewScene(device);
  rtcAttachGeometry(scene, boundary.getRTCGeometry());
  rtcAttachGeometry(scene, geo.getRTCGeometry());
  rtcJoinCommitScene(scene);
#pragma omp parallel for
  for (long i = 0; i < (long)m; ++i) {
    alignas(128) RTCRayHit rh{};
    rayInternal::fillRayPosition(rh.ray, Vec3D<float>{rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]});
    rayInternal::fillRayDirection<3>(rh.ray,
                                     Vec3D<float>{rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]});
    rh.ray.tfar = std::numeric_limits<float>::max();
    rh.hit.geomID = RTC_INVALID_GEOMETRY_ID;
    rh.hit.primID = RTC_INVALID_GEOMETRY_ID;
    rtcIntersect1(scene, &rh);
    geomOut[i] = rh.hit.geomID;
    primOut[i] = rh.hit.primID;
    tOut[i] = rh.ray.tfar;
    ngOut[3 * i] = rh.hit.Ng_x;
    ngOut[3 * i + 1] = rh.hit.Ng_y;
    ngOut[3 * i + 2] = rh.hit.Ng_z;
  }
  rtcReleaseScene(scene);
  boundary.releaseGeometry();
  geo.releaseGeometry();
  rtcReleaseDevice(device);
  return 0;
}

// Reference Boundary::processHit (rayBoundary.hpp:29-127) on a caller-built
// hit: in/out org[3], dir[3]; Ng, primID, t as an intersector would report.
// Returns the `reflect` flag.
int ref_boundary_process_hit(int D, const float *bboxMin, const float *bboxMax, const int *bc,
                             int sourceDir, float *org, float *dir, const float *ng,
                             uint32_t primID, float t) {
  auto run = [&](auto dimTag) -> int {
    constexpr int DD = decltype(dimTag)::value;
    auto device = rtcNewDevice("");
    std::array<Vec3D<float>, 2> bbox{Vec3D<float>{bboxMin[0], bboxMin[1], bboxMin[2]},
                                     Vec3D<float>{bboxMax[0], bboxMax[1], bboxMax[2]}};
    auto settings = rayInternal::getTraceSettings(static_cast<TraceDirection>(sourceDir));
    BoundaryCondition conds[3];
    for (int i = 0; i < 3; ++i)
      conds[i] = static_cast<BoundaryCondition>(bc[i]);
    Boundary<float, DD> boundary(device, bbox, conds, settings);
    alignas(128) RTCRayHit rh{};
    rayInternal::fillRayPosition(rh.ray, Vec3D<float>{org[0], org[1], org[2]});
    Vec3D<float> d{dir[0], dir[1], dir[2]};
    rayInternal::fillRayDirection<DD>(rh.ray, d);
    rh.ray.tfar = t;
    rh.hit.Ng_x = ng[0];
    rh.hit.Ng_y = ng[1];
    rh.hit.Ng_z = ng[2];
    rh.hit.primID = primID;
    rh.hit.geomID = 0;
    bool reflect = false;
    boundary.processHit(rh, reflect, d);
    org[0] = rh.ray.org_x;
    org[1] = rh.ray.org_y;
    org[2] = rh.ray.org_z;
    dir[0] = d[0];
    dir[1] = d[1];
    dir[2] = d[2];
    boundary.releaseGeometry();
    rtcReleaseDevice(device);
    return reflect ? 1 : 0;
  };
  if (D == 3)
    return run(std::integral_constant<int, 3>{});
  return run(std::integral_constant<int, 2>{});
}
}
