// Parity tests of the C++ host mirror (include/viennaray_b200/*.hpp), written
// the way the reference writes its own tests (tests/<name>/<name>.cpp of
// ViennaRay v4.2.0): one block per reference test, same set-up, same asserts.
//
//   test_host_api cpu            blocks that need no device
//   test_host_api gpu <outdir>   blocks that trace; flux arrays are dumped to
//                                <outdir>/*.f32 for the Python side to compare
//                                with the C-ABI / oracle path
#include <rayGeometryDisk.hpp>
#include <rayGeometryTriangle.hpp>
#include <rayParticle.hpp>
#include <rayPointNeighborhood.hpp>
#include <rayReflection.hpp>
#include <raySourceGrid.hpp>
#include <raySourceRandom.hpp>
#include <rayTraceDisk.hpp>
#include <rayTraceTriangle.hpp>
#include <rayTracingData.hpp>

#include <cstring>
#include <fstream>
#include <iostream>

using namespace viennaray;

static int failures = 0;
#define VC_TEST_ASSERT(cond)                                                                       \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      std::fprintf(stderr, "ASSERT FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                \
      ++failures;                                                                                  \
    }                                                                                              \
  } while (0)
#define VC_TEST_ASSERT_ISCLOSE(a, b, eps) VC_TEST_ASSERT(std::fabs(double(a) - double(b)) <= (eps))

// rayInternal::createPlaneGrid (rayUtil.hpp:324-351): same point order
template <class T>
static void planeGrid(T gridDelta, T extent, std::array<int, 3> dir, std::vector<Vec3D<T>> &points,
                      std::vector<Vec3D<T>> &normals) {
  Vec3D<T> point = {-extent, -extent, -extent}, normal = {0, 0, 0};
  point[dir[2]] = 0;
  normal[dir[2]] = 1;
  points.clear();
  normals.clear();
  point[dir[0]] = -extent;
  while (point[dir[0]] <= extent) {
    point[dir[1]] = -extent;
    while (point[dir[1]] <= extent) {
      points.push_back(point);
      normals.push_back(normal);
      point[dir[1]] += gridDelta;
    }
    point[dir[0]] += gridDelta;
  }
}

template <class T> static void dump(const std::string &path, const std::vector<T> &v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
}

// tests/tracingData/tracingData.cpp:11-35
static void testTracingData() {
  TracingData<float> defaultData;
  defaultData.setNumberOfVectorData(2);
  defaultData.setNumberOfScalarData(3);
  defaultData.setScalarData(0, 1.f, "zero");
  defaultData.setScalarData(1, 2.f, "one");
  defaultData.setScalarData(2, 3.f);
  defaultData.setVectorData(0, 1000, 0.f, "zero");
  defaultData.setVectorData(1, 1000, 1.f, "one");
  VC_TEST_ASSERT(defaultData.getVectorData(0).size() == 1000);
  VC_TEST_ASSERT(defaultData.getVectorData("one")[999] == 1.f);
  VC_TEST_ASSERT(defaultData.getScalarData("one") == 2.f);
  VC_TEST_ASSERT(defaultData.getVectorDataLabel(1) == "one");
  VC_TEST_ASSERT(defaultData.getScalarDataLabel(2) == "scalarData");
  VC_TEST_ASSERT(defaultData.getVectorDataIndex("one") == 1);
  TracingData<float> moved = std::move(defaultData);
  VC_TEST_ASSERT(moved.getVectorData().size() == 2);
  VC_TEST_ASSERT(moved.getScalarData().size() == 3);
  moved.resizeAllVectorData(10, 2.f);
  VC_TEST_ASSERT(moved.getVectorData(1).size() == 10 && moved.getVectorData(1)[3] == 2.f);
  moved.appendVectorData(0, std::vector<float>{1.f, 2.f});
  VC_TEST_ASSERT(moved.getVectorData(0).size() == 12);
  moved.setVectorMergeType(0, TracingDataMergeEnum::APPEND);
  VC_TEST_ASSERT(moved.getVectorMergeType(0) == TracingDataMergeEnum::APPEND);
  VC_TEST_ASSERT(moved.getScalarMergeType(0) == TracingDataMergeEnum::SUM);
}

// tests/particle/particle.cpp:17-39
static void testParticle() {
  auto particle = std::make_unique<DiffuseParticle<float, 3>>(0.5f, "test");
  VC_TEST_ASSERT(particle->getSourceDistributionPower() == 1.f);
  VC_TEST_ASSERT(particle->getLocalDataLabels().size() == 1);
  VC_TEST_ASSERT(particle->getLocalDataLabels()[0] == "test");
  VC_TEST_ASSERT(particle->getMeanFreePath() == -1.f);
  auto copy = particle->clone();
  VC_TEST_ASSERT(copy->getLocalDataLabels()[0] == "test");
  vr_particle_desc d{};
  VC_TEST_ASSERT(copy->deviceParticle(d) && d.kind == VR_PARTICLE_DIFFUSE && d.sticking == 0.5f);
  auto spec = std::make_unique<SpecularParticle<double, 2>>(1., 100., "s");
  VC_TEST_ASSERT(spec->getSourceDistributionPower() == 100.);
  VC_TEST_ASSERT(spec->deviceParticle(d) && d.kind == VR_PARTICLE_SPECULAR && d.sourcePower == 100.f);
  auto ion = std::make_unique<ConedCosineParticle<float, 3>>(0.5f, 100.f, 1.4835f, "ion");
  VC_TEST_ASSERT(ion->deviceParticle(d) && d.kind == VR_PARTICLE_CONED_COSINE);
}

// a built-in particle with a mean free path (AbstractParticle::getMeanFreePath)
template <class T> class ScatteringParticle : public Particle<ScatteringParticle<T>, T> {
public:
  T getMeanFreePath() const override { return T(3); }
  std::vector<std::string> getLocalDataLabels() const override { return {"flux"}; }
  bool deviceParticle(vr_particle_desc &d) const override {
    d = {VR_PARTICLE_DIFFUSE, 0.2f, 1.f, 0.f, 0.f};
    return true;
  }
};

// a user particle with host-side hooks: must be refused, never run on the CPU
template <class T> class UserParticle : public Particle<UserParticle<T>, T> {
public:
  std::vector<std::string> getLocalDataLabels() const override { return {"user"}; }
};

// tests/smoothing/smoothing.cpp:43,50
static void testSmoothing() {
  std::vector<Vec3D<float>> points = {{0, 0, 0}, {1, 0, 0}, {2, 0, 0}, {0, 1, 0}, {1, 1, 0}, {2, 1, 0}};
  std::vector<Vec3D<float>> normals = {{0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {0, 1, 0}, {0, 1, 0}};
  std::vector<float> flux = {1, 1, 1, 0, 0, 0};
  TraceDisk<float, 3> trace;
  trace.setGeometry(points, normals, 1.0f);
  trace.smoothFlux(flux, 1);
  for (int i = 0; i < 3; ++i)
    VC_TEST_ASSERT_ISCLOSE(flux[i], 1.0, 1e-6);
  for (int i = 3; i < 6; ++i)
    VC_TEST_ASSERT_ISCLOSE(flux[i], 0.0, 1e-6);
  std::vector<float> flux2 = {1, 1, 1, 0, 0, 0};
  trace.smoothFlux(flux2, 2);
  VC_TEST_ASSERT_ISCLOSE(flux2[0], 1.0, 1e-6);
}

// tests/pointNeighborhood/pointNeighborhood.cpp:51 and tests/diskAreas/diskAreas.cpp:76,79,96
static void testNeighborsAndAreas() {
  {
    std::vector<Vec3D<float>> points, normals;
    planeGrid<float>(0.5f, 10.f, {0, 1, 2}, points, normals);
    TraceDisk<float, 3> trace;
    trace.setGeometry(points, normals, 0.5f, 0.5f - 1e-6f);
    auto bd = trace.getBoundingBox();
    for (std::size_t i = 0; i < points.size(); ++i) {
      const bool ex = points[i][0] == bd[0][0] || points[i][0] == bd[1][0];
      const bool ey = points[i][1] == bd[0][1] || points[i][1] == bd[1][1];
      const std::size_t expect = ex && ey ? 3 : (ex || ey ? 5 : 8);
      VC_TEST_ASSERT(trace.getNeighborIndices(i).size() == expect);
    }
  }
  {
    std::vector<Vec3D<float>> points, normals;
    planeGrid<float>(1.f, 2.f, {0, 1, 2}, points, normals);
    TraceDisk<float, 3> trace;
    trace.setGeometry(points, normals, 1.f);
    const double r = 1.f * rayInternal::DiskFactor<3>;
    const double whole = float(r) * float(r) * M_PI;
    const auto bd = trace.getBoundingBox();
    const auto &areas = trace.getDiskAreas(); // default boundary conditions: reflective
    for (std::size_t i = 0; i < points.size(); ++i) {
      const bool ex = points[i][0] == bd[0][0] || points[i][0] == bd[1][0];
      const bool ey = points[i][1] == bd[0][1] || points[i][1] == bd[1][1];
      const double expect = ex && ey ? whole / 4 : (ex || ey ? whole / 2 : whole);
      VC_TEST_ASSERT_ISCLOSE(areas[i], expect, 1e-6);
    }
  }
}

// tests/buildBoundary/buildBoundary.cpp:34-39, tests/createGeometry/createGeometry.cpp:23-33
static void testBoundingBox() {
  std::array<std::array<float, 3>, 2> bb = {{{-1, -1, -1}, {1, 1, 1}}};
  rayInternal::adjustBoundingBox<3>(bb, TraceDirection::POS_Z, 0.5f);
  VC_TEST_ASSERT(bb[1][2] == 2.f && bb[0][2] == -1.f && bb[0][0] == -1.f && bb[1][1] == 1.f);
  bb = {{{-1, -1, 0}, {1, 1, 0}}};
  rayInternal::adjustBoundingBox<2>(bb, TraceDirection::NEG_Y, 0.25f);
  VC_TEST_ASSERT(bb[0][1] == -1.5f && bb[0][2] == -0.25f && bb[1][2] == 0.25f);
  const auto st = rayInternal::getTraceSettings(TraceDirection::POS_Y);
  VC_TEST_ASSERT(st[0] == 1 && st[1] == 0 && st[2] == 2 && st[3] == 1 && st[4] == -1);
  const auto b = rayInternal::getOrthonormalBasis({1.f, 1.f, -1.f});
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double d = 0;
      for (int k = 0; k < 3; ++k)
        d += double(b[i][k]) * b[j][k];
      VC_TEST_ASSERT_ISCLOSE(d, i == j ? 1. : 0., 1e-6);
    }
}

// no device: apply() must fail loudly (error flag), not fall back to host code
static void testNoFallback() {
  std::vector<Vec3D<float>> points, normals;
  planeGrid<float>(0.5f, 1.f, {0, 1, 2}, points, normals);
  TraceDisk<float, 3> trace;
  trace.setGeometry(points, normals, 0.5f);
  auto particle = std::make_unique<DiffuseParticle<float, 3>>(1.f, "flux");
  trace.setParticleType(particle);
  trace.setNumberOfRaysPerPoint(10);
  trace.apply();
  if (trace.getDeviceContext() == nullptr)
    VC_TEST_ASSERT(trace.getRayTraceInfo().error);
  TraceDisk<float, 3> noParticle;
  noParticle.setGeometry(points, normals, 0.5f);
  noParticle.apply();
  VC_TEST_ASSERT(noParticle.getRayTraceInfo().error);
}

// the helper headers a consumer includes by their reference names (rayReflection.hpp,
// raySourceRandom.hpp, rayPointNeighborhood.hpp, rayGeometry*.hpp): tests/reflection,
// tests/pointNeighborhood, tests/createRay of the reference, on the host helpers
static void testHostHelpers() {
  RNG rng(42);
  const Vec3D<float> down{0.f, 0.f, -1.f}, up{0.f, 0.f, 1.f};
  const Vec3D<float> in45{0.70710678f, 0.f, -0.70710678f};
  auto s = ReflectionSpecular<float, 3>(in45, up);
  VC_TEST_ASSERT_ISCLOSE(s[0], 0.70710678f, 1e-6);
  VC_TEST_ASSERT_ISCLOSE(s[2], 0.70710678f, 1e-6);
  double meanCos = 0;
  for (int i = 0; i < 20000; ++i) {
    auto d = ReflectionDiffuse<float, 3>(up, rng);
    VC_TEST_ASSERT(d[2] >= 0.f);
    VC_TEST_ASSERT_ISCLOSE(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1.f, 1e-5);
    meanCos += d[2];
  }
  VC_TEST_ASSERT_ISCLOSE(meanCos / 20000, 2. / 3., 0.01);  // cosine law: E[cos] = 2/3
  VC_TEST_ASSERT((ReflectionDiffuse<float, 2>(Vec3D<float>{0.f, 1.f, 0.f}, rng)[2] == 0.f));
  const float cone = 0.3f;
  for (int i = 0; i < 2000; ++i) {  // stays inside the cone about the specular direction
    auto d = ReflectionConedCosine<float, 3>(in45, up, rng, cone);
    const float c = d[0] * s[0] + d[1] * s[1] + d[2] * s[2];
    VC_TEST_ASSERT(c >= std::cos(cone) - 1e-4f && d[2] >= 0.f);
  }
  auto sp = ReflectionConedCosine<float, 3>(down, up, rng, 0.f);  // cone 0: specular
  VC_TEST_ASSERT_ISCLOSE(sp[2], 1.f, 1e-6);

  // tests/pointNeighborhood/pointNeighborhood.cpp: corner, edge, interior of a plane grid
  std::vector<Vec3D<float>> points, normals;
  planeGrid<float>(1.f, 3.f, {0, 1, 2}, points, normals);
  PointNeighborhood<float, 3> nb;
  nb.init<3>(points, 1.5f, Vec3D<float>{-3.f, -3.f, 0.f}, Vec3D<float>{3.f, 3.f, 0.f});
  VC_TEST_ASSERT(nb.getNumPoints() == 49 && nb.getDistance() == 1.5f);
  VC_TEST_ASSERT(nb.getNeighborIndices(0).size() == 3);
  VC_TEST_ASSERT(nb.getNeighborIndices(1).size() == 5);
  VC_TEST_ASSERT(nb.getNeighborIndices(8).size() == 8);

  GeometryDisk<float, 3> geo;
  geo.setMaterialIds(std::vector<int>(49, 7));
  geo.initGeometry<3>(points, normals, 0.75f);
  VC_TEST_ASSERT(geo.getNumPrimitives() == 49 && !geo.checkGeometryEmpty());
  VC_TEST_ASSERT(geo.getMaterialId(5) == 7 && geo.getNeighborIndices(8).size() == 8);
  VC_TEST_ASSERT(geo.getBoundingBox()[1][0] == 3.f && geo.getPrimNormal(3)[2] == 1.f);
  GeometryTriangle<float, 3> tri;
  tri.initGeometry({{0.f, 0.f, 0.f}, {2.f, 0.f, 0.f}, {0.f, 2.f, 0.f}}, {{0u, 1u, 2u}});
  VC_TEST_ASSERT_ISCLOSE(tri.getPrimArea(0), 2.f, 1e-6);
  VC_TEST_ASSERT_ISCLOSE(tri.getPrimNormal(0)[2], 1.f, 1e-6);

  // tests/createRay/createRay.cpp: origins on the source plane, directions towards the geometry
  std::array<Vec3D<float>, 2> bbox{Vec3D<float>{-1.f, -2.f, 0.f}, Vec3D<float>{1.f, 2.f, 5.f}};
  auto st = rayInternal::getTraceSettings(TraceDirection::POS_Z);
  std::array<Vec3D<float>, 3> basis{};
  SourceRandom<float, 3> src(bbox, 3.f, st, 100, false, basis);
  VC_TEST_ASSERT(src.getNumPoints() == 100 && src.getSourceArea() == 8.f);
  for (int i = 0; i < 1000; ++i) {
    auto od = src.getOriginAndDirection(i, rng);
    VC_TEST_ASSERT(od[0][2] == 5.f && od[1][2] < 0.f);
    VC_TEST_ASSERT(od[0][0] >= -1.f && od[0][0] <= 1.f && od[0][1] >= -2.f && od[0][1] <= 2.f);
    VC_TEST_ASSERT_ISCLOSE(od[1][0] * od[1][0] + od[1][1] * od[1][1] + od[1][2] * od[1][2], 1.f, 1e-5);
  }
  vr_source_desc sd{};
  std::vector<float> origins;
  VC_TEST_ASSERT(src.deviceSource(sd, origins) && sd.rayDir == 2 && sd.useGrid == 0);

  // a particle with material-dependent sticking maps to the descriptor's table
  auto mp = std::make_unique<MaterialStickingParticle<float, 3>>(
      VR_PARTICLE_DIFFUSE, 0.1f, std::vector<float>{0.05f, 0.6f}, 1.f, "flux");
  vr_particle_desc pd{};
  VC_TEST_ASSERT(mp->clone()->deviceParticle(pd) && pd.numMaterials == 2 &&
                 pd.stickingByMaterial[1] == 0.6f && pd.sticking == 0.1f);

  // the data-only particle of the reference's GPU tracer (rayParticle.hpp:208-218)
  gpu::Particle<float> gp;
  gp.name = "neutral";
  gp.sticking = 0.2f;
  gp.cosineExponent = 3.f;
  gp.materialSticking = {{2, 0.9f}, {0, 0.4f}};
  auto dp = gpu::makeParticle<float, 3>(gp);
  vr_particle_desc gd{};
  VC_TEST_ASSERT(dp->deviceParticle(gd) && gd.kind == VR_PARTICLE_DIFFUSE && gd.numMaterials == 3 &&
                 gd.stickingByMaterial[0] == 0.4f && gd.stickingByMaterial[1] == 0.2f &&
                 gd.stickingByMaterial[2] == 0.9f && gd.sourcePower == 3.f);
  VC_TEST_ASSERT(dp->getLocalDataLabels()[0] == "neutral");
}

// ------------------------------------------------------------------ device blocks
// tests/rngSeed/rngSeed.cpp:29,41,48-51
static void testRngSeed(const std::string &out) {
  std::vector<Vec3D<float>> points, normals;
  planeGrid<float>(0.5f, 5.f, {0, 1, 2}, points, normals);
  std::vector<float> flux[2];
  for (int run = 0; run < 2; ++run) {
    TraceDisk<float, 3> tracer;
    auto particle = std::make_unique<DiffuseParticle<float, 3>>(1.0f, "flux");
    tracer.setGeometry(points, normals, 0.5f);
    tracer.setNumberOfRaysPerPoint(10);
    tracer.setParticleType(particle);
    tracer.setRngSeed(12345);
    tracer.apply();
    VC_TEST_ASSERT(!tracer.getRayTraceInfo().error);
    flux[run] = tracer.getLocalData().getVectorData("flux");
  }
  VC_TEST_ASSERT(flux[0].size() == 441);
  VC_TEST_ASSERT(std::memcmp(flux[0].data(), flux[1].data(), sizeof(float) * 441) == 0);
  dump(out + "/rngSeed_flux.f32", flux[0]);
}

// tests/traceInterface/traceInterface.cpp:55-67
static void testTraceInterface(const std::string &out) {
  std::vector<Vec3D<float>> points, normals;
  planeGrid<float>(0.5f, 5.f, {0, 1, 2}, points, normals);
  std::vector<int> matIds(points.size(), 0);
  BoundaryCondition bc[3] = {BoundaryCondition::REFLECTIVE_BOUNDARY, BoundaryCondition::REFLECTIVE_BOUNDARY,
                             BoundaryCondition::REFLECTIVE_BOUNDARY};
  auto particle = std::make_unique<DiffuseParticle<float, 3>>(0.5f, "hitFlux");
  TracingData<float> globalData;
  globalData.setNumberOfVectorData(1);

  TraceDisk<float, 3> rayTracer;
  rayTracer.setParticleType(particle);
  rayTracer.setGeometry(points, normals, 0.5f);
  rayTracer.setBoundaryConditions(bc);
  rayTracer.setMaterialIds(matIds);
  rayTracer.setGlobalData(globalData);
  rayTracer.setSourceDirection(TraceDirection::POS_Z);
  rayTracer.setNumberOfRaysPerPoint(10);
  rayTracer.setUseRandomSeeds(false);
  rayTracer.setMaxBoundaryHits(10);
  rayTracer.apply();
  auto info = rayTracer.getRayTraceInfo();
  VC_TEST_ASSERT(!info.error);
  VC_TEST_ASSERT(info.numRays == 4410);
  VC_TEST_ASSERT(info.totalRaysTraced >= info.numRays);
  VC_TEST_ASSERT(info.geometryHits > 0 && info.nonGeometryHits > 0);
  auto flux = rayTracer.getLocalData().getVectorData("hitFlux");
  dump(out + "/traceInterface_raw.f32", flux);
  {  // device post-processing == host post-processing (normalise + smooth over 1 ring)
    auto hostFlux = flux;
    rayTracer.normalizeFlux(hostFlux);
    rayTracer.smoothFlux(hostFlux, 1);
    auto devFlux = rayTracer.getDeviceFlux(true, true);
    VC_TEST_ASSERT(devFlux.size() == hostFlux.size());
    for (std::size_t i = 0; i < devFlux.size(); ++i)
      VC_TEST_ASSERT_ISCLOSE(devFlux[i], hostFlux[i], 1e-5 * (1 + std::fabs(hostFlux[i])));
  }
  {  // the general device call == host normalizeFlux(MAX) + smoothFlux(2) (wider neighbourhood)
    auto hostFlux = flux;
    rayTracer.normalizeFlux(hostFlux, NormalizationType::MAX);
    rayTracer.smoothFlux(hostFlux, 2);
    auto devFlux = rayTracer.getDeviceFlux(NormalizationType::MAX, 2);
    VC_TEST_ASSERT(devFlux.size() == hostFlux.size());
    for (std::size_t i = 0; i < devFlux.size(); ++i)
      VC_TEST_ASSERT_ISCLOSE(devFlux[i], hostFlux[i], 1e-5 * (1 + std::fabs(hostFlux[i])));
  }
  rayTracer.normalizeFlux(flux);
  rayTracer.smoothFlux(flux, 2);
  double mean = 0;
  for (auto f : flux)
    mean += f;
  mean /= flux.size();
  VC_TEST_ASSERT(mean > 0.5 && mean < 1.5); // open plane under a cosine source: flux ~ 1
  dump(out + "/traceInterface_norm.f32", flux);

  // run number advances: a second apply() draws a different stream
  rayTracer.apply();
  auto flux2 = rayTracer.getLocalData().getVectorData("hitFlux");
  VC_TEST_ASSERT(std::memcmp(flux2.data(), rayTracer.getLocalData().getVectorData(0).data(), 4 * 441) == 0);
  dump(out + "/traceInterface_run2.f32", flux2);

  // mean-free-path scattering shows up in TraceInfo::particleHits (rayTraceKernel.hpp:179-203)
  {
    TraceDisk<float, 3> mfpTracer;
    auto sp = std::make_unique<ScatteringParticle<float>>();
    mfpTracer.setParticleType(sp);
    mfpTracer.setGeometry(points, normals, 0.5f);
    mfpTracer.setNumberOfRaysPerPoint(10);
    mfpTracer.setRngSeed(1);
    mfpTracer.apply();
    VC_TEST_ASSERT(!mfpTracer.getRayTraceInfo().error);
    VC_TEST_ASSERT(mfpTracer.getRayTraceInfo().particleHits > 0);
  }

  // host-side hooks cannot run on the device: explicit error, no fallback
  TraceDisk<float, 3> userTracer;
  auto user = std::make_unique<UserParticle<float>>();
  userTracer.setParticleType(user);
  userTracer.setGeometry(points, normals, 0.5f);
  userTracer.apply();
  VC_TEST_ASSERT(userTracer.getRayTraceInfo().error);
}

// tests/createSourceGrid/createSourceGrid.cpp:43-46 and a trace through setSource()
static void testSourceGrid(const std::string &out) {
  std::vector<Vec3D<float>> points, normals;
  planeGrid<float>(0.5f, 5.f, {0, 1, 2}, points, normals);
  TraceDisk<float, 3> tracer;
  tracer.setGeometry(points, normals, 0.5f);
  auto bb = tracer.getBoundingBox();
  rayInternal::adjustBoundingBox<3>(bb, TraceDirection::POS_Z, float(0.5 * rayInternal::DiskFactor<3>));
  std::array<Vec3D<float>, 2> bdBox = {Vec3D<float>{bb[0][0], bb[0][1], bb[0][2]},
                                       Vec3D<float>{bb[1][0], bb[1][1], bb[1][2]}};
  const auto settings = rayInternal::getTraceSettings(TraceDirection::POS_Z);
  auto grid = rayInternal::createSourceGrid<float, 3>(bdBox, points.size(), 0.5f, settings);
  VC_TEST_ASSERT(grid.size() > 100 && grid.size() <= points.size());
  auto source = std::make_shared<SourceGrid<float, 3>>(bdBox, grid, 1.f, settings);
  RNG rng(0);
  for (std::size_t i = 0; i < grid.size(); ++i) {
    auto od = source->getOriginAndDirection(i, rng);
    VC_TEST_ASSERT(od[1][2] < 0.f);
    VC_TEST_ASSERT_ISCLOSE(od[0][2], bb[1][2], 1e-6);
    VC_TEST_ASSERT_ISCLOSE(od[0][0], grid[i][0], 1e-6);
    VC_TEST_ASSERT_ISCLOSE(od[0][1], grid[i][1], 1e-6);
  }
  if (out.empty())
    return;
  auto particle = std::make_unique<DiffuseParticle<float, 3>>(1.0f, "flux");
  tracer.setParticleType(particle);
  tracer.setSource(source);
  tracer.setNumberOfRaysPerPoint(20);
  tracer.setRngSeed(5);
  tracer.apply();
  auto info = tracer.getRayTraceInfo();
  VC_TEST_ASSERT(!info.error);
  VC_TEST_ASSERT(info.numRays == 20 * grid.size());
  auto flux = tracer.getLocalData().getVectorData(0);
  double sum = 0;
  for (auto f : flux)
    sum += f;
  VC_TEST_ASSERT(sum >= double(info.numRays)); // every ray lands on the plane at least once
  tracer.resetSource();
  tracer.apply();
  VC_TEST_ASSERT(tracer.getRayTraceInfo().numRays == 20 * points.size());
  dump(out + "/sourceGrid_flux.f32", flux);
}

// examples/triangle3D-style run on a two-triangle floor + tests/trace2D
static void testTriangleAnd2D(const std::string &out) {
  {
    std::vector<Vec3D<float>> nodes = {{-5, -5, 0}, {5, -5, 0}, {5, 5, 0}, {-5, 5, 0}};
    std::vector<Vec3D<unsigned>> tris = {{0, 1, 2}, {0, 2, 3}};
    TraceTriangle<float, 3> tracer;
    auto particle = std::make_unique<DiffuseParticle<float, 3>>(1.0f, "flux");
    tracer.setGeometry(nodes, tris, 1.0f);
    tracer.setParticleType(particle);
    tracer.setNumberOfRaysFixed(20000);
    tracer.setRngSeed(7);
    tracer.apply();
    auto info = tracer.getRayTraceInfo();
    VC_TEST_ASSERT(!info.error && info.numRays == 20000);
    auto flux = tracer.getLocalData().getVectorData(0);
    VC_TEST_ASSERT_ISCLOSE(flux[0] + flux[1], 20000., 0.5); // every ray lands on the floor once
    tracer.normalizeFlux(flux);
    VC_TEST_ASSERT_ISCLOSE(flux[0], 1.0, 0.05);
    VC_TEST_ASSERT_ISCLOSE(flux[1], 1.0, 0.05);
    dump(out + "/triangle_norm.f32", flux);
  }
  {
    std::vector<Vec2D<float>> points, normals;
    for (float x = -10.f; x <= 10.f; x += 0.5f) {
      points.push_back({x, 0.f});
      normals.push_back({0.f, 1.f});
    }
    BoundaryCondition bc[2] = {BoundaryCondition::PERIODIC_BOUNDARY, BoundaryCondition::PERIODIC_BOUNDARY};
    TraceDisk<float, 2> tracer;
    auto particle = std::make_unique<DiffuseParticle<float, 2>>(0.1f, "flux");
    tracer.setGeometry(points, normals, 0.5f);
    tracer.setBoundaryConditions(bc);
    tracer.setParticleType(particle);
    tracer.setNumberOfRaysPerPoint(2000);
    tracer.setRngSeed(3);
    tracer.apply();
    VC_TEST_ASSERT(!tracer.getRayTraceInfo().error);
    auto flux = tracer.getLocalData().getVectorData(0);
    tracer.normalizeFlux(flux);
    tracer.smoothFlux(flux);
    for (std::size_t i = 2; i + 2 < flux.size(); ++i)
      VC_TEST_ASSERT_ISCLOSE(flux[i], 1.0, 0.08);
    dump(out + "/disk2D_norm.f32", flux);
  }
}

// areas <in.bin> <out.f64>: disk areas of a point cloud under periodic boundaries
// (input: uint32 n, float gridDelta, n x 3 points, n x 3 normals), for the
// comparison with the reference's GeometryDisk::computeDiskAreas
static int areasTool(const char *in, const char *out) {
  std::ifstream f(in, std::ios::binary);
  std::uint32_t n = 0;
  float gd = 0;
  f.read(reinterpret_cast<char *>(&n), 4);
  f.read(reinterpret_cast<char *>(&gd), 4);
  std::vector<Vec3D<float>> pts(n), nrm(n);
  f.read(reinterpret_cast<char *>(pts.data()), 12 * n);
  f.read(reinterpret_cast<char *>(nrm.data()), 12 * n);
  BoundaryCondition bc[3] = {BoundaryCondition::PERIODIC_BOUNDARY, BoundaryCondition::PERIODIC_BOUNDARY,
                             BoundaryCondition::PERIODIC_BOUNDARY};
  TraceDisk<float, 3> trace;
  trace.setGeometry(pts, nrm, gd);
  trace.setBoundaryConditions(bc);
  dump(out, trace.getDiskAreas());
  return 0;
}

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  if (mode == "areas" && argc > 3)
    return areasTool(argv[2], argv[3]);
  testTracingData();
  testParticle();
  testSmoothing();
  testNeighborsAndAreas();
  testBoundingBox();
  testNoFallback();
  testHostHelpers();
  if (mode != "gpu")
    testSourceGrid("");
  if (mode == "gpu") {
    const std::string out = argc > 2 ? argv[2] : ".";
    testRngSeed(out);
    testTraceInterface(out);
    testTriangleAnd2D(out);
    testSourceGrid(out);
  }
  if (failures) {
    std::fprintf(stderr, "%d assertion(s) failed\n", failures);
    return 1;
  }
  std::printf("test_host_api %s: ok\n", mode.c_str());
  return 0;
}
