// viennaray_b200 -- host-side C++ mirror of ViennaRay's Trace / Particle /
// TracingData interface (reference: include/viennaray/rayTrace.hpp,
// rayTraceDisk.hpp, rayTraceTriangle.hpp, rayParticle.hpp, rayTracingData.hpp,
// rayUtil.hpp of ViennaRay v4.2.0) on top of the C ABI in
// ../viennaray_b200.h.  Same class names, template parameters, setters,
// defaults and result containers, so code written against the reference's
// CPU tracer compiles against this header; apply() runs the sm_100a kernels.
//
// What differs from the reference, by construction:
//   * particles run as device functors.  The built-in particles
//     (DiffuseParticle, SpecularParticle) and ConedCosineParticle describe
//     themselves through AbstractParticle::deviceParticle(); any other
//     subclass makes apply() set TraceInfo::error (there is no CPU fallback);
//   * the same holds for user-defined Source subclasses (setSource);
//   * the per-ray random stream is a counter-based Philox keyed on
//     (seed, ray index) instead of mt19937_64(tea<3>(idx, seed))
//     (rayTraceKernel.hpp:120-121), so equal seeds give bitwise-equal
//     results run to run (tests/rngSeed) but not the reference's stream.
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../viennaray_b200.h"

#if __has_include(<vcVectorType.hpp>) && __has_include(<vcRNG.hpp>) && __has_include(<vcLogger.hpp>)
// inside a ViennaTools build: ViennaCore supplies the vector types, RNG and log macros
#include <vcLogger.hpp>
#include <vcRNG.hpp>
#include <vcVectorType.hpp>
#else
namespace viennacore {
template <class T, std::size_t D> using VectorType = std::array<T, D>;
template <class T> using Vec2D = std::array<T, 2>;
template <class T> using Vec3D = std::array<T, 3>;
using Vec3Df = Vec3D<float>;
using RNG = std::mt19937_64;

template <class T, std::size_t D>
T DotProduct(const std::array<T, D> &a, const std::array<T, D> &b) {
  T s = 0;
  for (std::size_t i = 0; i < D; ++i)
    s += a[i] * b[i];
  return s;
}
template <class T> Vec3D<T> CrossProduct(const Vec3D<T> &a, const Vec3D<T> &b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
template <class T, std::size_t D> T Norm(const std::array<T, D> &v) {
  return std::sqrt(DotProduct(v, v));
}
template <class T, std::size_t D> void Normalize(std::array<T, D> &v) {
  const T n = Norm(v);
  if (n > 0)
    for (auto &x : v)
      x /= n;
}
template <class T, std::size_t D> std::array<T, D> Normalize(const std::array<T, D> &v) {
  auto r = v;
  Normalize(r);
  return r;
}
template <class T, std::size_t D>
std::array<T, D> operator-(const std::array<T, D> &a, const std::array<T, D> &b) {
  std::array<T, D> r;
  for (std::size_t i = 0; i < D; ++i)
    r[i] = a[i] - b[i];
  return r;
}
template <class T, std::size_t D>
std::array<T, D> operator+(const std::array<T, D> &a, const std::array<T, D> &b) {
  std::array<T, D> r;
  for (std::size_t i = 0; i < D; ++i)
    r[i] = a[i] + b[i];
  return r;
}
} // namespace viennacore
#define VIENNACORE_LOG_ERROR(msg) std::fprintf(stderr, "[viennaray_b200] ERROR: %s\n", std::string(msg).c_str())
#define VIENNACORE_LOG_WARNING(msg) std::fprintf(stderr, "[viennaray_b200] WARNING: %s\n", std::string(msg).c_str())
#define VIENNACORE_LOG_DEBUG(msg) ((void)0)
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace viennaray {
using namespace viennacore;

// ---- enums and plain structs (rayUtil.hpp:36-76, rayBoundary.hpp:10-14) ------
enum class NormalizationType : unsigned { SOURCE = 0, MAX = 1 };
enum class TraceDirection : unsigned { POS_X = 0, NEG_X = 1, POS_Y = 2, NEG_Y = 3, POS_Z = 4, NEG_Z = 5 };
enum class BoundaryCondition : unsigned {
  REFLECTIVE_BOUNDARY = VR_BOUNDARY_REFLECTIVE,
  PERIODIC_BOUNDARY = VR_BOUNDARY_PERIODIC,
  IGNORE_BOUNDARY = VR_BOUNDARY_IGNORE
};
enum class GeometryType : unsigned { TRIANGLE = 0, DISK = 1, UNDEFINED = 2 };

template <class NumericType> struct DataLog {
  std::vector<std::vector<NumericType>> data;
  void merge(DataLog<NumericType> &other) {
    assert(other.data.size() == data.size());
    for (std::size_t i = 0; i < data.size(); ++i)
      for (std::size_t j = 0; j < data[i].size(); ++j)
        data[i][j] += other.data[i][j];
  }
};

struct TraceInfo {
  std::size_t numRays = 0, totalRaysTraced = 0, nonGeometryHits = 0, geometryHits = 0,
              particleHits = 0, boundaryHits = 0, reflections = 0;
  double time = 0.0; // seconds of device time (CUDA events around the kernels)
  bool warning = false, error = false;
};

// ---- TracingData (rayTracingData.hpp:10-219) --------------------------------
enum class TracingDataMergeEnum : unsigned { SUM = 0, APPEND = 1, AVERAGE = 2 };

template <typename NumericType> class TracingData {
  using scalarDataType = NumericType;
  using vectorDataType = std::vector<NumericType>;
  using mergeType = std::vector<TracingDataMergeEnum>;

  std::vector<scalarDataType> scalars_;
  std::vector<vectorDataType> vectors_;
  std::vector<std::string> scalarLabels_, vectorLabels_;
  mergeType scalarMerge_, vectorMerge_;

  template <class Vec> static bool inRange(int i, const Vec &v) {
    return i >= 0 && static_cast<std::size_t>(i) < v.size();
  }

public:
  TracingData() = default;
  TracingData(const TracingData &) = default;
  TracingData(TracingData &&) noexcept = default;
  TracingData &operator=(const TracingData &) = default;
  TracingData &operator=(TracingData &&) noexcept = default;

  void appendVectorData(int num, const std::vector<NumericType> &vec) {
    vectors_[num].insert(vectors_[num].end(), vec.begin(), vec.end());
  }
  void setNumberOfVectorData(int size) {
    vectors_.assign(size, {});
    vectorLabels_.assign(size, "vectorData");
    vectorMerge_.assign(size, TracingDataMergeEnum::SUM);
  }
  void setNumberOfScalarData(int size) {
    scalars_.assign(size, NumericType(0));
    scalarLabels_.assign(size, "scalarData");
    scalarMerge_.assign(size, TracingDataMergeEnum::SUM);
  }
  void setScalarData(int num, NumericType value, std::string label = "scalarData") {
    if (!inRange(num, scalars_)) {
      VIENNACORE_LOG_ERROR("Setting scalar data in TracingData out of range.");
      return;
    }
    scalars_[num] = value;
    scalarLabels_[num] = std::move(label);
  }
  void setVectorData(int num, std::vector<NumericType> &vector, std::string label = "vectorData") {
    if (!inRange(num, vectors_)) {
      VIENNACORE_LOG_ERROR("Setting vector data in TracingData out of range.");
      return;
    }
    vectors_[num] = vector;
    vectorLabels_[num] = std::move(label);
  }
  void setVectorData(int num, std::vector<NumericType> &&vector, std::string label = "vectorData") {
    if (!inRange(num, vectors_)) {
      VIENNACORE_LOG_ERROR("Setting vector data in TracingData out of range.");
      return;
    }
    vectors_[num] = std::move(vector);
    vectorLabels_[num] = std::move(label);
  }
  void setVectorData(int num, std::size_t size, NumericType value, std::string label = "vectorData") {
    if (!inRange(num, vectors_)) {
      VIENNACORE_LOG_ERROR("Setting vector data in TracingData out of range.");
      return;
    }
    vectors_[num].assign(size, value);
    vectorLabels_[num] = std::move(label);
  }
  void setVectorData(int num, NumericType value, std::string label = "vectorData") {
    if (!inRange(num, vectors_)) {
      VIENNACORE_LOG_ERROR("Setting vector data in TracingData out of range.");
      return;
    }
    std::fill(vectors_[num].begin(), vectors_[num].end(), value);
    vectorLabels_[num] = std::move(label);
  }
  void resizeAllVectorData(std::size_t size, NumericType val = 0) {
    for (auto &v : vectors_)
      v.assign(size, val);
  }
  void setVectorMergeType(const mergeType &m) { vectorMerge_ = m; }
  void setVectorMergeType(int num, TracingDataMergeEnum m) { vectorMerge_[num] = m; }
  void setScalarMergeType(const mergeType &m) { scalarMerge_ = m; }
  void setScalarMergeType(int num, TracingDataMergeEnum m) { scalarMerge_[num] = m; }

  [[nodiscard]] vectorDataType &getVectorData(int i) { return vectors_[i]; }
  [[nodiscard]] const vectorDataType &getVectorData(int i) const { return vectors_[i]; }
  [[nodiscard]] vectorDataType &getVectorData(const std::string &label) {
    return vectors_[getVectorDataIndex(label)];
  }
  [[nodiscard]] std::vector<vectorDataType> &getVectorData() { return vectors_; }
  [[nodiscard]] const std::vector<vectorDataType> &getVectorData() const { return vectors_; }
  [[nodiscard]] scalarDataType &getScalarData(int i) { return scalars_[i]; }
  [[nodiscard]] const scalarDataType &getScalarData(int i) const { return scalars_[i]; }
  [[nodiscard]] scalarDataType &getScalarData(const std::string &label) {
    return scalars_[getScalarDataIndex(label)];
  }
  [[nodiscard]] std::vector<scalarDataType> &getScalarData() { return scalars_; }
  [[nodiscard]] const std::vector<scalarDataType> &getScalarData() const { return scalars_; }
  [[nodiscard]] std::string getVectorDataLabel(int i) const {
    if (!inRange(i, vectorLabels_)) {
      VIENNACORE_LOG_ERROR("Getting vector data label in TracingData out of range.");
      return "";
    }
    return vectorLabels_[i];
  }
  [[nodiscard]] std::string getScalarDataLabel(int i) const {
    if (!inRange(i, scalarLabels_)) {
      VIENNACORE_LOG_ERROR("Getting scalar data label in TracingData out of range.");
      return "";
    }
    return scalarLabels_[i];
  }
  [[nodiscard]] int getVectorDataIndex(const std::string &label) const {
    for (std::size_t i = 0; i < vectorLabels_.size(); ++i)
      if (vectorLabels_[i] == label)
        return static_cast<int>(i);
    VIENNACORE_LOG_ERROR("Can not find vector data label in TracingData.");
    return -1;
  }
  [[nodiscard]] int getScalarDataIndex(const std::string &label) const {
    for (std::size_t i = 0; i < scalarLabels_.size(); ++i)
      if (scalarLabels_[i] == label)
        return static_cast<int>(i);
    VIENNACORE_LOG_ERROR("Can not find scalar data label in TracingData.");
    return -1;
  }
  [[nodiscard]] TracingDataMergeEnum getVectorMergeType(int num) const { return vectorMerge_[num]; }
  [[nodiscard]] TracingDataMergeEnum getScalarMergeType(int num) const { return scalarMerge_[num]; }
};

// ---- particles (rayParticle.hpp:21-204) ------------------------------------
#define VIENNARAY_PARTICLE_STOP                                                                    \
  std::pair<NumericType, Vec3D<NumericType>> { NumericType(1), Vec3D<NumericType>{} }

template <typename NumericType> class AbstractParticle {
public:
  virtual ~AbstractParticle() = default;
  virtual std::unique_ptr<AbstractParticle> clone() const = 0;
  virtual void initNew(RNG &rngState) = 0;
  virtual Vec3D<NumericType> initNewWithDirection(RNG &rngState) = 0;
  virtual std::pair<NumericType, Vec3D<NumericType>>
  surfaceReflection(NumericType rayWeight, const Vec3D<NumericType> &rayDir,
                    const Vec3D<NumericType> &geomNormal, const unsigned int primId,
                    const int materialId, const TracingData<NumericType> *globalData,
                    RNG &rngState) = 0;
  virtual void surfaceCollision(NumericType rayWeight, const Vec3D<NumericType> &rayDir,
                                const Vec3D<NumericType> &geomNormal, const unsigned int primID,
                                const int materialId, TracingData<NumericType> &localData,
                                const TracingData<NumericType> *globalData, RNG &rngState) = 0;
  virtual NumericType getSourceDistributionPower() const = 0;
  virtual NumericType getMeanFreePath() const = 0;
  [[nodiscard]] virtual std::vector<std::string> getLocalDataLabels() const = 0;
  virtual void logData(DataLog<NumericType> &log) = 0;

  /// B200 path: the device functor this particle maps to.  Built-in particles
  /// override it; a particle that keeps the default cannot be traced (its
  /// virtual hooks are host code) and apply() reports an error.
  virtual bool deviceParticle(vr_particle_desc &) const { return false; }
};

template <typename Derived, typename NumericType>
class Particle : public AbstractParticle<NumericType> {
public:
  std::unique_ptr<AbstractParticle<NumericType>> clone() const final {
    return std::make_unique<Derived>(static_cast<Derived const &>(*this));
  }
  void initNew(RNG &) override {}
  Vec3D<NumericType> initNewWithDirection(RNG &) override { return Vec3D<NumericType>{0, 0, 0}; }
  std::pair<NumericType, Vec3D<NumericType>>
  surfaceReflection(NumericType, const Vec3D<NumericType> &, const Vec3D<NumericType> &,
                    const unsigned int, const int, const TracingData<NumericType> *,
                    RNG &) override {
    return VIENNARAY_PARTICLE_STOP;
  }
  void surfaceCollision(NumericType, const Vec3D<NumericType> &, const Vec3D<NumericType> &,
                        const unsigned int, const int, TracingData<NumericType> &,
                        const TracingData<NumericType> *, RNG &) override {}
  NumericType getSourceDistributionPower() const override { return 1.; }
  NumericType getMeanFreePath() const override { return -1.; }
  [[nodiscard]] std::vector<std::string> getLocalDataLabels() const override { return {}; }
  void logData(DataLog<NumericType> &) override {}

protected:
  Particle() = default;
  Particle(const Particle &) = default;
  Particle(Particle &&) = default;
};

/// Constant sticking, diffuse (cosine) re-emission -- rayParticle.hpp:124-161.
template <typename NumericType, int D>
class DiffuseParticle : public Particle<DiffuseParticle<NumericType, D>, NumericType> {
  const NumericType sticking_;
  const std::string label_;

public:
  DiffuseParticle(NumericType stickingProbability, std::string dataLabel)
      : sticking_(stickingProbability), label_(std::move(dataLabel)) {}
  NumericType getSourceDistributionPower() const final { return 1.; }
  [[nodiscard]] std::vector<std::string> getLocalDataLabels() const final { return {label_}; }
  bool deviceParticle(vr_particle_desc &d) const final {
    d = {VR_PARTICLE_DIFFUSE, static_cast<float>(sticking_), 1.f, 0.f};
    return true;
  }
};

/// Constant sticking, specular reflection, power-cosine source -- rayParticle.hpp:163-204.
template <typename NumericType, int D>
class SpecularParticle : public Particle<SpecularParticle<NumericType, D>, NumericType> {
  const NumericType sticking_, sourcePower_;
  const std::string label_;

public:
  SpecularParticle(NumericType stickingProbability, NumericType sourcePower, std::string dataLabel)
      : sticking_(stickingProbability), sourcePower_(sourcePower), label_(std::move(dataLabel)) {}
  NumericType getSourceDistributionPower() const final { return sourcePower_; }
  [[nodiscard]] std::vector<std::string> getLocalDataLabels() const final { return {label_}; }
  bool deviceParticle(vr_particle_desc &d) const final {
    d = {VR_PARTICLE_SPECULAR, static_cast<float>(sticking_), static_cast<float>(sourcePower_), 0.f};
    return true;
  }
};

/// Ion-like particle: constant sticking, power-cosine source, re-emission by
/// ReflectionConedCosine (rayReflection.hpp:52-120) with the cone
/// pi/2 - min(incidence angle, minAngle) -- the recipe of
/// tests/reflection/reflection.cpp:43-46.
template <typename NumericType, int D>
class ConedCosineParticle : public Particle<ConedCosineParticle<NumericType, D>, NumericType> {
  const NumericType sticking_, sourcePower_, minAngle_;
  const std::string label_;

public:
  ConedCosineParticle(NumericType stickingProbability, NumericType sourcePower,
                      NumericType minAngle, std::string dataLabel)
      : sticking_(stickingProbability), sourcePower_(sourcePower), minAngle_(minAngle),
        label_(std::move(dataLabel)) {}
  NumericType getSourceDistributionPower() const final { return sourcePower_; }
  [[nodiscard]] std::vector<std::string> getLocalDataLabels() const final { return {label_}; }
  bool deviceParticle(vr_particle_desc &d) const final {
    d = {VR_PARTICLE_CONED_COSINE, static_cast<float>(sticking_),
         static_cast<float>(sourcePower_), static_cast<float>(minAngle_)};
    return true;
  }
};

/// A particle whose sticking probability depends on the material of the hit primitive --
/// the pattern of ViennaPS's particles, which switch on the materialId that
/// surfaceReflection receives (rayParticle.hpp:44-48, rayTraceKernel.hpp:310-313).
/// `reflection`: VR_PARTICLE_DIFFUSE / _SPECULAR / _CONED_COSINE.  stickingByMaterial[m] is
/// used for materialId m; IDs outside the table fall back to `stickingProbability`.
/// Any user particle reaches the device the same way: by overriding deviceParticle().
template <typename NumericType, int D>
class MaterialStickingParticle
    : public Particle<MaterialStickingParticle<NumericType, D>, NumericType> {
  const int reflection_;
  const NumericType sticking_, sourcePower_, minAngle_, meanFreePath_;
  const std::vector<float> table_;
  const std::string label_;

public:
  MaterialStickingParticle(int reflection, NumericType stickingProbability,
                           std::vector<float> stickingByMaterial, NumericType sourcePower,
                           std::string dataLabel, NumericType minAngle = 0,
                           NumericType meanFreePath = -1)
      : reflection_(reflection), sticking_(stickingProbability), sourcePower_(sourcePower),
        minAngle_(minAngle), meanFreePath_(meanFreePath), table_(std::move(stickingByMaterial)),
        label_(std::move(dataLabel)) {}
  NumericType getSourceDistributionPower() const final { return sourcePower_; }
  NumericType getMeanFreePath() const final { return meanFreePath_; }
  [[nodiscard]] std::vector<std::string> getLocalDataLabels() const final { return {label_}; }
  bool deviceParticle(vr_particle_desc &d) const final {
    d = {reflection_, static_cast<float>(sticking_), static_cast<float>(sourcePower_),
         static_cast<float>(minAngle_)};
    d.stickingByMaterial = table_.empty() ? nullptr : table_.data();  // lives as long as *this
    d.numMaterials = static_cast<std::int32_t>(table_.size());
    return true;
  }
};

// ---- the data-only particle of the reference's own GPU tracer ----------------------------
// rayParticle.hpp:208-218 (viennaray::gpu::Particle, behind VIENNARAY_USE_GPU there): sticking,
// sticking by material ID, cosine exponent of the source, optional direction.  A caller that
// describes its particles this way gets the device particle from makeParticle(); the direction
// goes to Trace::setPrimaryDirection as in the reference's CPU interface.
namespace gpu {
template <typename T> struct Particle {
  std::string name;
  std::vector<std::string> dataLabels;

  T sticking = 1.;
  std::unordered_map<int, T> materialSticking;
  T cosineExponent = 1.;

  bool useCustomDirection = false;
  Vec3D<T> direction = {0., 0., -1.0};
};

/// Diffuse device particle of a gpu::Particle.  Materials missing from the map keep the
/// particle's own sticking (raygTrace.hpp:174-184 builds the same table); material IDs are
/// the table's indices, so they must be non-negative.
template <typename NumericType, int D>
std::unique_ptr<AbstractParticle<NumericType>> makeParticle(const Particle<NumericType> &p) {
  int maxId = -1;
  for (auto const &kv : p.materialSticking)
    maxId = std::max(maxId, kv.first);
  std::vector<float> table(static_cast<std::size_t>(maxId + 1), static_cast<float>(p.sticking));
  for (auto const &kv : p.materialSticking)
    if (kv.first >= 0)
      table[static_cast<std::size_t>(kv.first)] = static_cast<float>(kv.second);
  return std::make_unique<MaterialStickingParticle<NumericType, D>>(
      VR_PARTICLE_DIFFUSE, p.sticking, std::move(table), p.cosineExponent,
      p.dataLabels.empty() ? (p.name.empty() ? std::string("flux") : p.name) : p.dataLabels[0]);
}
} // namespace gpu

// ---- sources (raySource.hpp:10-19) -------------------------------------------
template <typename NumericType> class Source {
public:
  virtual ~Source() = default;
  virtual std::array<Vec3D<NumericType>, 2> getOriginAndDirection(std::size_t idx,
                                                                  RNG &rngState) const = 0;
  [[nodiscard]] virtual std::size_t getNumPoints() const = 0;
  virtual NumericType getSourceArea() const = 0;
  /// B200 path: device description of the source (and, for a grid source, its
  /// origins as n x 3 floats); false = host-only source.
  virtual bool deviceSource(vr_source_desc &, std::vector<float> &) const { return false; }
};

/// Rays start on a regular grid of origins, ray idx at origin idx % numPoints, with a
/// cos^n direction distribution -- raySourceGrid.hpp:9-74.
template <typename NumericType, int D> class SourceGrid : public Source<NumericType> {
  using boundingBoxType = std::array<Vec3D<NumericType>, 2>;
  const boundingBoxType bdBox_;
  const std::vector<Vec3D<NumericType>> &sourceGrid_;
  const std::array<int, 5> settings_;
  const NumericType cosinePower_;

public:
  SourceGrid(const boundingBoxType &boundingBox, std::vector<Vec3D<NumericType>> &sourceGrid,
             NumericType cosinePower, const std::array<int, 5> &traceSettings)
      : bdBox_(boundingBox), sourceGrid_(sourceGrid), settings_(traceSettings),
        cosinePower_(cosinePower) {}
  std::array<Vec3D<NumericType>, 2> getOriginAndDirection(std::size_t idx, RNG &rngState) const override {
    std::uniform_real_distribution<NumericType> uni;
    const NumericType r1 = uni(rngState), r2 = uni(rngState);
    const NumericType tt = std::pow(r2, NumericType(2) / (cosinePower_ + 1));
    Vec3D<NumericType> d{0, 0, 0};
    d[settings_[0]] = settings_[4] * std::sqrt(tt);
    d[settings_[1]] = std::cos(NumericType(2 * M_PI) * r1) * std::sqrt(1 - tt);
    d[settings_[2]] = D == 2 ? NumericType(0) : std::sin(NumericType(2 * M_PI) * r1) * std::sqrt(1 - tt);
    Normalize(d);
    return {sourceGrid_[idx % sourceGrid_.size()], d};
  }
  [[nodiscard]] std::size_t getNumPoints() const override { return sourceGrid_.size(); }
  NumericType getSourceArea() const override {
    const NumericType a = bdBox_[1][settings_[1]] - bdBox_[0][settings_[1]];
    return D == 2 ? a : a * (bdBox_[1][settings_[2]] - bdBox_[0][settings_[2]]);
  }
  [[nodiscard]] NumericType getCosinePower() const { return cosinePower_; }
  bool deviceSource(vr_source_desc &d, std::vector<float> &origins) const override {
    for (int a = 0; a < 3; ++a) {
      d.bboxMin[a] = static_cast<float>(bdBox_[0][a]);
      d.bboxMax[a] = static_cast<float>(bdBox_[1][a]);
    }
    d.rayDir = settings_[0];
    d.firstDir = settings_[1];
    d.secondDir = settings_[2];
    d.minMax = settings_[3];
    d.posNeg = static_cast<float>(settings_[4]);
    d.useBasis = 0;
    d.useGrid = 1;
    origins.resize(3 * sourceGrid_.size());
    for (std::size_t i = 0; i < sourceGrid_.size(); ++i)
      for (int a = 0; a < 3; ++a)
        origins[3 * i + a] = static_cast<float>(sourceGrid_[i][a]);
    return !origins.empty();
  }
};

// ---- meshes (rayMesh.hpp:88-145), plain data ---------------------------------
struct TriangleMesh {
  std::vector<Vec3Df> nodes;
  std::vector<Vec3D<unsigned>> triangles;
  std::vector<Vec3Df> normals;
  Vec3Df minimumExtent{}, maximumExtent{};
  float gridDelta = 0.f;
  TriangleMesh() = default;
  TriangleMesh(std::vector<Vec3Df> const &pts, std::vector<Vec3D<unsigned>> const &tris, float delta)
      : nodes(pts), triangles(tris), gridDelta(delta) {}
};
struct DiskMesh {
  std::vector<Vec3Df> nodes, normals;
  std::vector<float> radii;
  Vec3Df minimumExtent{}, maximumExtent{};
  float radius = 0.f, gridDelta = 0.f;
  DiskMesh() = default;
  DiskMesh(const std::vector<Vec3Df> &pts, const std::vector<Vec3Df> &nms, float delta)
      : nodes(pts), normals(nms), gridDelta(delta) {}
};
} // namespace viennaray

// ---- internals (rayUtil.hpp:83-202,287-321) ----------------------------------
namespace rayInternal {
using namespace viennaray;

struct KernelConfig {
  std::size_t numRaysPerPoint = 1000;
  std::size_t numRaysFixed = 0;
  unsigned maxReflections = std::numeric_limits<unsigned>::max();
  unsigned maxBoundaryHits = 1000;
  unsigned rngSeed = 0;
  bool useRandomSeed = true;
  bool printProgress = false;
  unsigned runNumber = 1;
};

using rtcNumericType = float;

template <int D>
constexpr double DiskFactor = 0.5 * (D == 3 ? 1.7320508 : 1.41421356237) * (1 + 1e-5);

// {sourceDir, boundaryDir1, boundaryDir2, minMax, posNeg} -- rayUtil.hpp:145-202
inline std::array<int, 5> getTraceSettings(TraceDirection dir) {
  switch (dir) {
  case TraceDirection::POS_X: return {0, 1, 2, 1, -1};
  case TraceDirection::NEG_X: return {0, 1, 2, 0, 1};
  case TraceDirection::POS_Y: return {1, 0, 2, 1, -1};
  case TraceDirection::NEG_Y: return {1, 0, 2, 0, 1};
  case TraceDirection::POS_Z: return {2, 0, 1, 1, -1};
  default: return {2, 0, 1, 0, 1};
  }
}

// rayUtil.hpp:104-143: in 2D the unused z extent becomes +-offset; the source
// plane sits 2 * offset outside the geometry along the tracing axis
template <int D>
void adjustBoundingBox(std::array<std::array<float, 3>, 2> &bbox, TraceDirection dir, float offset) {
  if (D == 2) {
    bbox[0][2] -= offset;
    bbox[1][2] += offset;
  }
  const auto st = getTraceSettings(dir);
  if (st[3])
    bbox[1][st[0]] += 2 * offset;
  else
    bbox[0][st[0]] -= 2 * offset;
}

// regular grid of about numPoints origins on the source plane, 1e-4 inside the lateral
// box -- rayUtil.hpp:566-611
template <class NumericType, int D>
std::vector<Vec3D<NumericType>> createSourceGrid(const std::array<Vec3D<NumericType>, 2> &bdBox,
                                                 const std::size_t numPoints, const NumericType gridDelta,
                                                 const std::array<int, 5> &traceSettings) {
  std::vector<Vec3D<NumericType>> grid;
  const double eps = 1e-4;
  const int rayDir = traceSettings[0], first = traceSettings[1], second = traceSettings[2],
            minMax = traceSettings[3];
  const auto len1 = bdBox[1][first] - bdBox[0][first], len2 = bdBox[1][second] - bdBox[0][second];
  auto n1 = static_cast<std::size_t>(std::round(len1 / gridDelta));
  auto n2 = static_cast<std::size_t>(std::round(len2 / gridDelta));
  const unsigned long ratio = n1 / n2;
  n1 = static_cast<std::size_t>(std::sqrt(numPoints * ratio));
  n2 = static_cast<std::size_t>(std::sqrt(numPoints / ratio));
  const auto d1 = (len1 - 2 * eps) / static_cast<NumericType>(n1 - 1);
  const auto d2 = (len2 - 2 * eps) / static_cast<NumericType>(n2 - 1);
  Vec3D<NumericType> point{};
  point[rayDir] = bdBox[minMax][rayDir];
  for (auto uu = bdBox[0][second] + eps; uu <= bdBox[1][second] - eps; uu += d2) {
    point[second] = D == 2 ? NumericType(0) : static_cast<NumericType>(uu);
    for (auto vv = bdBox[0][first] + eps; vv <= bdBox[1][first] - eps; vv += d1) {
      point[first] = static_cast<NumericType>(vv);
      grid.push_back(point);
    }
  }
  return grid;
}

// rows u, v, w of the orthonormal basis whose first vector is `vec` -- rayUtil.hpp:287-321
inline std::array<std::array<float, 3>, 3> getOrthonormalBasis(std::array<float, 3> u) {
  auto scale = [](std::array<float, 3> &v) {
    const float inv = 1.0f / std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    for (auto &x : v)
      x *= inv;
  };
  scale(u);
  std::array<float, 3> h = std::fabs(u[0]) > std::fabs(u[2]) ? std::array<float, 3>{-u[1], u[0], 0.f}
                                                            : std::array<float, 3>{0.f, -u[2], u[1]};
  scale(h);
  std::array<float, 3> w = {u[1] * h[2] - u[2] * h[1], u[2] * h[0] - u[0] * h[2],
                            u[0] * h[1] - u[1] * h[0]};
  return {u, h, w};
}

// Area of the unit-normal disk (centre c, radius r) inside the prism
// [lo0,hi0] x [lo1,hi1] spanned along the axes a0, a1 -- what
// DiskBoundingBoxXYIntersector::areaInside computes
// (rayDiskBoundingBoxIntersector.hpp:39-76).  Exact: the prism cuts the disk's
// plane in up to four half planes; the circle is intersected with that convex
// polygon edge by edge.
inline double diskAreaInsidePrism(const float c[3], const float nrm[3], double r, int a0, int a1,
                                  double lo0, double hi0, double lo1, double hi1) {
  const double kPi = 3.14159265358979323846;
  if (lo0 <= c[a0] - r && c[a0] + r <= hi0 && lo1 <= c[a1] - r && c[a1] + r <= hi1)
    return r * r * kPi;
  double n[3] = {nrm[0], nrm[1], nrm[2]};
  const double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  for (auto &x : n)
    x /= nl;
  // in-plane frame (u, v)
  double t[3] = {0, 0, 0};
  t[std::fabs(n[0]) < std::fabs(n[1]) ? (std::fabs(n[0]) < std::fabs(n[2]) ? 0 : 2)
                                      : (std::fabs(n[1]) < std::fabs(n[2]) ? 1 : 2)] = 1;
  double u[3] = {n[1] * t[2] - n[2] * t[1], n[2] * t[0] - n[0] * t[2], n[0] * t[1] - n[1] * t[0]};
  const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  for (auto &x : u)
    x /= ul;
  const double v[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};
  // polygon: a square around the circle, clipped by g . (a,b) >= h for each face
  std::vector<std::array<double, 2>> poly = {{-2 * r, -2 * r}, {2 * r, -2 * r}, {2 * r, 2 * r}, {-2 * r, 2 * r}};
  auto clip = [&](double ga, double gb, double h) {
    if (std::fabs(ga) + std::fabs(gb) < 1e-14) { // face parallel to the disk plane
      if (0 < h)
        poly.clear();
      return;
    }
    std::vector<std::array<double, 2>> out;
    for (std::size_t i = 0; i < poly.size(); ++i) {
      const auto &p = poly[i], &q = poly[(i + 1) % poly.size()];
      const double dp = ga * p[0] + gb * p[1] - h, dq = ga * q[0] + gb * q[1] - h;
      if (dp >= 0)
        out.push_back(p);
      if ((dp >= 0) != (dq >= 0)) {
        const double s = dp / (dp - dq);
        out.push_back({p[0] + s * (q[0] - p[0]), p[1] + s * (q[1] - p[1])});
      }
    }
    poly.swap(out);
  };
  clip(u[a0], v[a0], lo0 - c[a0]);
  clip(-u[a0], -v[a0], c[a0] - hi0);
  clip(u[a1], v[a1], lo1 - c[a1]);
  clip(-u[a1], -v[a1], c[a1] - hi1);
  if (poly.size() < 3)
    return 0.;
  auto cross = [](const std::array<double, 2> &p, const std::array<double, 2> &q) {
    return p[0] * q[1] - p[1] * q[0];
  };
  auto sector = [&](const std::array<double, 2> &p, const std::array<double, 2> &q) {
    return 0.5 * r * r * std::atan2(cross(p, q), p[0] * q[0] + p[1] * q[1]);
  };
  double area = 0.;
  for (std::size_t i = 0; i < poly.size(); ++i) {
    const auto &p = poly[i], &q = poly[(i + 1) % poly.size()];
    const double dx = q[0] - p[0], dy = q[1] - p[1];
    const double A = dx * dx + dy * dy, B = 2 * (p[0] * dx + p[1] * dy),
                 C = p[0] * p[0] + p[1] * p[1] - r * r;
    const double disc = B * B - 4 * A * C;
    if (A < 1e-30 || disc <= 0) {
      area += sector(p, q);
      continue;
    }
    const double sq = std::sqrt(disc);
    const double t1 = std::max((-B - sq) / (2 * A), 0.), t2 = std::min((-B + sq) / (2 * A), 1.);
    if (t1 >= t2) {
      area += sector(p, q);
      continue;
    }
    const std::array<double, 2> e1 = {p[0] + t1 * dx, p[1] + t1 * dy}, e2 = {p[0] + t2 * dx, p[1] + t2 * dy};
    area += sector(p, e1) + 0.5 * cross(e1, e2) + sector(e2, q);
  }
  return std::max(area, 0.);
}
} // namespace rayInternal

namespace viennaray {

// ---- Trace (rayTrace.hpp:15-182) ---------------------------------------------
template <class NumericType, int D> class Trace {
public:
  Trace() {
    // VIENNARAY_B200_DEVICES="0,1,2,3" (or "all"): several GPUs of the node behind this one
    // object -- rays sharded over them, one NCCL all-reduce of the flux, no change for the
    // caller.  Otherwise one device: VIENNARAY_B200_DEVICE (default 0).
    std::vector<int> ids;
    if (const char *list = std::getenv("VIENNARAY_B200_DEVICES")) {
      const std::string l(list);
      if (l == "all") {
        for (int i = 0; i < 64; ++i) {  // probe: a plain context per ordinal until one fails
          vr_ctx *probe = nullptr;
          if (vr_ctx_create(i, &probe) != VR_OK)
            break;
          vr_ctx_destroy(probe);
          ids.push_back(i);
        }
      } else {
        std::size_t pos = 0;
        while (pos < l.size()) {
          const std::size_t end = l.find(',', pos);
          const std::string tok = l.substr(pos, end == std::string::npos ? end : end - pos);
          if (!tok.empty())
            ids.push_back(std::atoi(tok.c_str()));
          if (end == std::string::npos)
            break;
          pos = end + 1;
        }
      }
    }
    if (ids.empty()) {
      const char *dev = std::getenv("VIENNARAY_B200_DEVICE");
      ids.push_back(dev ? std::atoi(dev) : 0);
    }
    if (vr_ctx_create_multi(static_cast<int>(ids.size()), ids.data(), &ctx_) != VR_OK) {
      createError_ = vr_last_error(nullptr);
      ctx_ = nullptr;
    }
  }
  Trace(const Trace &) = delete;
  Trace &operator=(const Trace &) = delete;
  Trace(Trace &&) = delete;
  Trace &operator=(Trace &&) = delete;
  virtual ~Trace() { vr_ctx_destroy(ctx_); }

  virtual void apply() {}

  template <typename ParticleType,
            std::enable_if_t<std::is_base_of_v<AbstractParticle<NumericType>, ParticleType>, bool> = true>
  void setParticleType(std::unique_ptr<ParticleType> const &particle) {
    pParticle_ = particle->clone();
  }
  void setBoundaryConditions(BoundaryCondition boundaryConditions[D]) {
    for (int i = 0; i < D; ++i)
      boundaryConditions_[i] = boundaryConditions[i];
    setupChanged();
  }
  void setSource(std::shared_ptr<Source<NumericType>> source) {
    pSource_ = std::move(source);
    useCustomSource = true;
  }
  void resetSource() {
    pSource_.reset();
    useCustomSource = false;
  }
  void enableProgressBar() { config_.printProgress = true; }
  void disableProgressBar() { config_.printProgress = false; }
  void setNumberOfRaysPerPoint(const std::size_t numRaysPerPoint) {
    config_.numRaysPerPoint = numRaysPerPoint;
    config_.numRaysFixed = 0;
  }
  void setNumberOfRaysFixed(const std::size_t numRaysFixed) {
    config_.numRaysFixed = numRaysFixed;
    config_.numRaysPerPoint = 0;
  }
  void setMaxReflections(const unsigned maxReflections) { config_.maxReflections = maxReflections; }
  void setMaxBoundaryHits(const unsigned maxBoundaryHits) { config_.maxBoundaryHits = maxBoundaryHits; }
  void setSourceDirection(const TraceDirection direction) {
    sourceDirection_ = direction;
    setupChanged();
  }
  void setPrimaryDirection(const Vec3D<NumericType> primaryDirection) {
    primaryDirection_ = primaryDirection;
    usePrimaryDirection_ = true;
  }
  void setUseRandomSeeds(const bool useRand) { config_.useRandomSeed = useRand; }
  void setRngSeed(const unsigned int seed) {
    config_.rngSeed = seed;
    config_.useRandomSeed = false;
  }

  virtual void normalizeFlux(std::vector<NumericType> &flux,
                             NormalizationType norm = NormalizationType::SOURCE) = 0;
  virtual void smoothFlux(std::vector<NumericType> &flux, int numNeighbors = 1) = 0;

  [[nodiscard]] TracingData<NumericType> &getLocalData() { return localData_; }
  [[nodiscard]] TracingData<NumericType> *getGlobalData() { return pGlobalData_; }
  void setGlobalData(TracingData<NumericType> &data) { pGlobalData_ = &data; }
  [[nodiscard]] TraceInfo getRayTraceInfo() const { return RTInfo_; }
  [[nodiscard]] DataLog<NumericType> &getDataLog() { return dataLog_; }

  /// B200 additions: shard of the ray-index range traced by this object
  /// (multi-GPU: rank g of G sets [g*N/G, (g+1)*N/G) and all-reduces the flux).
  void setRayIndexShard(std::uint64_t begin, std::uint64_t end) {
    shardBegin_ = begin;
    shardEnd_ = end;
  }
  [[nodiscard]] vr_ctx *getDeviceContext() { return ctx_; }

protected:
  // boundary conditions / source direction changed: derived data that depends on them (the
  // disk areas, which the reference recomputes in every apply(), rayTraceDisk.hpp:28)
  virtual void setupChanged() {}

  // everything apply() needs besides the geometry: boundary, source, particle, config.
  // Returns false (with RTInfo_.error set) when the trace cannot run.
  bool traceCommitted(std::array<std::array<float, 3>, 2> bbox, float sourceOffset,
                      std::size_t numPoints, bool sceneDirty) {
    if (!ctx_) {
      RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("No B200 device context: " + createError_);
      return false;
    }
    if (!pParticle_)  // reported by checkSettings(); nothing can be traced
      return false;
    vr_particle_desc pd{};
    if (!pParticle_->deviceParticle(pd)) {
      RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("This particle type has no device functor; only the built-in particles "
                           "can be traced on the GPU. Aborting.");
      return false;
    }
    pd.meanFreePath = static_cast<float>(pParticle_->getMeanFreePath()); // <= 0: no scattering
    rayInternal::adjustBoundingBox<D>(bbox, sourceDirection_, sourceOffset);
    const auto st = rayInternal::getTraceSettings(sourceDirection_);
    const int condFirst = static_cast<int>(boundaryConditions_[st[1]]);
    const int condSecond = D == 3 ? static_cast<int>(boundaryConditions_[st[2]]) : VR_BOUNDARY_IGNORE;
    if (sceneDirty || bbox != lastBBox_ || condFirst != lastCond_[0] || condSecond != lastCond_[1]) {
      if (vr_scene_set_boundary(ctx_, bbox[0].data(), bbox[1].data(), st[1], st[2], condFirst,
                                condSecond, D) != VR_OK ||
          vr_scene_commit(ctx_) != VR_OK) {
        RTInfo_.error = true;
        VIENNACORE_LOG_ERROR(std::string("Scene commit failed: ") + vr_last_error(ctx_));
        return false;
      }
      lastBBox_ = bbox;
      lastCond_[0] = condFirst;
      lastCond_[1] = condSecond;
    }
    vr_source_desc src{};
    std::size_t sourcePoints = numPoints;  // SourceRandom: one "point" per primitive
    if (useCustomSource) {
      std::vector<float> origins;
      if (!pSource_ || !pSource_->deviceSource(src, origins) ||
          (src.useGrid && vr_source_set_grid(ctx_, origins.data(),
                                             static_cast<std::uint32_t>(origins.size() / 3)) != VR_OK)) {
        RTInfo_.error = true;
        VIENNACORE_LOG_ERROR("This custom source is host code and cannot be traced on the GPU "
                             "(only SourceGrid has a device form). Aborting.");
        return false;
      }
      sourcePoints = pSource_->getNumPoints();
    } else {
      for (int a = 0; a < 3; ++a) {
        src.bboxMin[a] = bbox[0][a];
        src.bboxMax[a] = bbox[1][a];
      }
      src.rayDir = st[0];
      src.firstDir = st[1];
      src.secondDir = st[2];
      src.minMax = st[3];
      src.posNeg = static_cast<float>(st[4]);
      src.useBasis = usePrimaryDirection_ ? 1 : 0;
      if (usePrimaryDirection_) {
        const auto b = rayInternal::getOrthonormalBasis({static_cast<float>(primaryDirection_[0]),
                                                         static_cast<float>(primaryDirection_[1]),
                                                         static_cast<float>(primaryDirection_[2])});
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)
            src.basis[3 * r + c] = b[r][c];
      }
    }
    sourceArea_ = bbox[1][st[1]] - bbox[0][st[1]];
    if (D == 3)
      sourceArea_ *= bbox[1][st[2]] - bbox[0][st[2]];
    sourcePoints_ = sourcePoints;

    const auto labels = pParticle_->getLocalDataLabels();
    if (!labels.empty()) {
      localData_.setNumberOfVectorData(static_cast<int>(labels.size()));
      for (std::size_t i = 0; i < labels.size(); ++i)
        localData_.setVectorData(static_cast<int>(i), numPoints, NumericType(0), labels[i]);
    }

    vr_config cfg{};
    cfg.numRays = config_.numRaysFixed == 0 ? sourcePoints * config_.numRaysPerPoint : config_.numRaysFixed;
    cfg.rayIdxBegin = std::min<std::uint64_t>(shardBegin_, cfg.numRays);
    cfg.rayIdxEnd = std::min<std::uint64_t>(shardEnd_, cfg.numRays);
    // rayTraceKernel.hpp:100-104
    cfg.seed = config_.useRandomSeed ? std::random_device{}() : config_.runNumber + config_.rngSeed;
    cfg.maxReflections = config_.maxReflections;
    cfg.maxBoundaryHits = config_.maxBoundaryHits;
#ifdef VIENNARAY_USE_WDIST // the reference's compile-time option (CMakeLists.txt:70-73)
    cfg.flags |= VR_FLAG_WDIST;
#endif
    std::vector<double> flux(numPoints);
    vr_trace_info info{};
    if (vr_trace(ctx_, &src, &pd, 1, &cfg, flux.data(), &info) != VR_OK) {
      RTInfo_.error = true;
      VIENNACORE_LOG_ERROR(std::string("Trace failed: ") + vr_last_error(ctx_));
      return false;
    }
    if (!labels.empty()) {
      auto &out = localData_.getVectorData(0);
      for (std::size_t i = 0; i < numPoints; ++i)
        out[i] = static_cast<NumericType>(flux[i]);
    }
    RTInfo_.numRays = cfg.numRays;
    RTInfo_.totalRaysTraced = info.totalRaysTraced;
    RTInfo_.nonGeometryHits = info.nonGeometryHits;
    RTInfo_.geometryHits = info.geometryHits;
    RTInfo_.particleHits = info.particleHits;
    RTInfo_.boundaryHits = info.boundaryHits;
    RTInfo_.reflections = info.reflections;
    RTInfo_.time = info.time;
    ++config_.runNumber;
    return true;
  }

  [[nodiscard]] std::size_t totalRays() const {
    return config_.numRaysFixed == 0 ? sourcePoints_ * config_.numRaysPerPoint : config_.numRaysFixed;
  }

  vr_ctx *ctx_ = nullptr;
  std::string createError_;
  std::shared_ptr<Source<NumericType>> pSource_ = nullptr;
  std::unique_ptr<AbstractParticle<NumericType>> pParticle_ = nullptr;
  NumericType gridDelta_ = 0;
  BoundaryCondition boundaryConditions_[D] = {};
  TraceDirection sourceDirection_ = D == 2 ? TraceDirection::POS_Y : TraceDirection::POS_Z;
  Vec3D<NumericType> primaryDirection_{NumericType(0), NumericType(0), NumericType(0)};
  bool usePrimaryDirection_ = false;
  bool useCustomSource = false;
  rayInternal::KernelConfig config_;
  TracingData<NumericType> localData_;
  TracingData<NumericType> *pGlobalData_ = nullptr;
  TraceInfo RTInfo_;
  DataLog<NumericType> dataLog_;

  std::uint64_t shardBegin_ = 0, shardEnd_ = std::numeric_limits<std::uint64_t>::max();
  std::array<std::array<float, 3>, 2> lastBBox_{};
  int lastCond_[2] = {-1, -1};
  bool haveSource_ = false;
  double sourceArea_ = 0;
  std::size_t sourcePoints_ = 0;
};

// ---- TraceDisk (rayTraceDisk.hpp:12-222, rayGeometryDisk.hpp:102-354) -----------
template <class NumericType, int D> class TraceDisk final : public Trace<NumericType, D> {
public:
  TraceDisk() = default;
  ~TraceDisk() override = default;

  void apply() override {
    // like rayTraceDisk.hpp:19-20: every problem is reported and flagged, then the call goes
    // on; what cannot run (no particle, no geometry, a source direction the geometry does not
    // have) is refused by the stage that needs it, with RTInfo_.error set
    checkSettings();
    const bool dirty = sceneDirty_;
    bool uploaded = !xyzr_.empty();
    if (uploaded && dirty && !(geometryOnDevice_ && materialIds_.empty())) {
      // (with material IDs the geometry goes up again: they arrive after setGeometry)
      const std::uint32_t n = static_cast<std::uint32_t>(numPoints());
      if (!this->ctx_ || vr_scene_set_disks(this->ctx_, xyzr_.data(), normals_.data(), n,
                                            materialIds_.empty() ? nullptr : materialIds_.data(),
                                            nbOff_.data(), nbIdx_.empty() ? &zero_ : nbIdx_.data()) != VR_OK) {
        this->RTInfo_.error = true;
        VIENNACORE_LOG_ERROR(std::string("Geometry upload failed: ") +
                             (this->ctx_ ? vr_last_error(this->ctx_) : this->createError_.c_str()));
        uploaded = false;
      }
    }
    if (uploaded && this->traceCommitted(bbox_, static_cast<float>(diskRadius_), numPoints(), dirty)) {
      sceneDirty_ = false;
      this->haveSource_ = true;
    } else {
      ++this->config_.runNumber;  // rayTraceDisk.hpp:54 counts every apply()
    }
  }

  template <std::size_t Dim>
  void setGeometry(std::vector<VectorType<NumericType, Dim>> const &points,
                   std::vector<VectorType<NumericType, Dim>> const &normals,
                   const NumericType gridDelta) {
    static_assert(!(D == 3 && Dim == 2), "Setting 2D geometry in 3D trace object");
    setGeometry(points, normals, gridDelta,
                static_cast<NumericType>(gridDelta * rayInternal::DiskFactor<D>));
  }

  template <std::size_t Dim>
  void setGeometry(std::vector<VectorType<NumericType, Dim>> const &points,
                   std::vector<VectorType<NumericType, Dim>> const &normals,
                   const NumericType gridDelta, const NumericType diskRadii) {
    static_assert(!(D == 3 && Dim == 2), "Setting 2D geometry in 3D trace object");
    assert(points.size() == normals.size());
    this->gridDelta_ = gridDelta;
    diskRadius_ = diskRadii;
    const std::size_t n = points.size();
    std::vector<float> pts3(3 * n, 0.f); // the neighbourhood sees the points as given
    xyzr_.assign(4 * n, 0.f);
    normals_.assign(3 * n, 0.f);
    for (int a = 0; a < 3; ++a) {
      bbox_[0][a] = a < D ? std::numeric_limits<float>::max() : 0.f;
      bbox_[1][a] = a < D ? std::numeric_limits<float>::lowest() : 0.f;
    }
    for (std::size_t i = 0; i < n; ++i) {
      for (std::size_t a = 0; a < Dim; ++a) {
        const float v = static_cast<float>(points[i][a]);
        pts3[3 * i + a] = v;
        if (static_cast<int>(a) < D) { // rayGeometryDisk.hpp:148-151,171-175: z is 0 in 2D
          xyzr_[4 * i + a] = v;
          normals_[3 * i + a] = static_cast<float>(normals[i][a]);
          bbox_[0][a] = std::min(bbox_[0][a], v);
          bbox_[1][a] = std::max(bbox_[1][a], v);
        }
      }
      xyzr_[4 * i + 3] = static_cast<float>(diskRadius_);
    }
    diskAreas_.clear();
    materialIds_.clear();
    sceneDirty_ = true;
    // with a device: upload now and build the neighbourhood there (uniform-grid
    // kernels, ~150x the host build at 1M points); the lists come back for the
    // host-side smoothFlux.  Without one: host build (vr_build_neighbors).
    if (this->ctx_ && n > 0 &&
        vr_scene_set_disks(this->ctx_, xyzr_.data(), normals_.data(), static_cast<std::uint32_t>(n),
                           nullptr, nullptr, nullptr) == VR_OK &&
        vr_scene_build_neighbors(this->ctx_, D, pts3.data(), 2 * static_cast<float>(diskRadius_)) == VR_OK) {
      std::uint32_t *o = nullptr, *x = nullptr;
      if (vr_scene_get_neighbors(this->ctx_, &o, &x) == VR_OK) {
        nbOff_.assign(o, o + n + 1);
        nbIdx_.assign(x, x + nbOff_[n]);
        vr_free(o);
        vr_free(x);
        geometryOnDevice_ = true;
        return;
      }
    }
    geometryOnDevice_ = false;
    buildNeighbors(pts3, 2 * static_cast<float>(diskRadius_), nbOff_, nbIdx_);
  }

  void setGeometry(const DiskMesh &mesh) {
    setGeometry(mesh.nodes, mesh.normals, static_cast<NumericType>(mesh.gridDelta));
  }

  /// rayTraceDisk.hpp:96-98.  The IDs reach the device with the next apply(); a particle
  /// whose deviceParticle() gives a sticking table (vr_particle_desc::stickingByMaterial) is
  /// handed the ID of every disk it hits, as rayTraceKernel.hpp:290,310-313 does.
  template <typename T> void setMaterialIds(std::vector<T> const &materialIds) {
    materialIds_.assign(materialIds.begin(), materialIds.end());
    sceneDirty_ = true;
  }

  // rayTraceDisk.hpp:103-142
  void normalizeFlux(std::vector<NumericType> &flux,
                     NormalizationType norm = NormalizationType::SOURCE) override {
    assert(flux.size() == numPoints() && "Unequal number of points in normalizeFlux");
    computeDiskAreas();
    if (norm == NormalizationType::MAX) {
      const double full = double(diskRadius_) * double(diskRadius_) * M_PI;
      const NumericType maxv = *std::max_element(flux.begin(), flux.end());
      for (std::size_t i = 0; i < flux.size(); ++i)
        flux[i] *= static_cast<NumericType>((full / diskAreas_[i]) / maxv);
    } else if (norm == NormalizationType::SOURCE) {
      if (!this->haveSource_) {
        VIENNACORE_LOG_WARNING("No source was specified in rayTrace for the normalization.");
        return;
      }
      const NumericType normFactor =
          static_cast<NumericType>(this->sourceArea_) / static_cast<NumericType>(this->totalRays());
      for (std::size_t i = 0; i < flux.size(); ++i)
        flux[i] *= normFactor / static_cast<NumericType>(diskAreas_[i]);
    }
  }

  // rayTraceDisk.hpp:146-193
  void smoothFlux(std::vector<NumericType> &flux, int numNeighbors = 1) override {
    assert(flux.size() == numPoints() && "Unequal number of points in smoothFlux");
    if (numNeighbors < 1)
      return;
    const std::vector<std::uint32_t> *off = &nbOff_, *idx = &nbIdx_;
    std::vector<std::uint32_t> wideOff, wideIdx;
    if (numNeighbors > 1) {
      std::vector<float> pts(3 * numPoints());
      for (std::size_t i = 0; i < numPoints(); ++i)
        for (int a = 0; a < 3; ++a)
          pts[3 * i + a] = xyzr_[4 * i + a];
      buildNeighbors(pts, numNeighbors * 2 * static_cast<float>(diskRadius_), wideOff, wideIdx);
      off = &wideOff;
      idx = &wideIdx;
    }
    const std::vector<NumericType> old = flux;
    for (std::size_t i = 0; i < numPoints(); ++i) {
      NumericType vv = old[i], sum = 1;
      const float *ni = &normals_[3 * i];
      for (std::uint32_t k = (*off)[i]; k < (*off)[i + 1]; ++k) {
        const std::uint32_t j = (*idx)[k];
        const float *nj = &normals_[3 * j];
        const NumericType w = static_cast<NumericType>((ni[0] * nj[0] + ni[1] * nj[1]) + ni[2] * nj[2]);
        if (w > 0) {
          vv += old[j] * w;
          sum += w;
        }
      }
      flux[i] = vv / sum;
    }
  }

  /// B200 addition: normalizeFlux(SOURCE) and smoothFlux(1) of the last apply()
  /// computed on the device (vr_flux_postprocess); returns an empty vector on error.
  [[nodiscard]] std::vector<NumericType> getDeviceFlux(bool normalize = true, bool smooth = true) {
    std::vector<float> out(numPoints()), areas;
    if (!this->ctx_ || !this->haveSource_)
      return {};
    if (normalize) {
      computeDiskAreas();
      areas.assign(diskAreas_.begin(), diskAreas_.end());
    }
    const float normFactor = static_cast<float>(this->sourceArea_) / static_cast<float>(this->totalRays());
    if (vr_flux_postprocess(this->ctx_, 0, normalize ? areas.data() : nullptr, normFactor, smooth ? 1 : 0,
                            out.data()) != VR_OK)
      return {};
    return std::vector<NumericType>(out.begin(), out.end());
  }

  /// The general form: normalizeFlux(norm) followed by smoothFlux(numNeighbors) of particle
  /// `particle` of the last apply(), on the device (vr_flux_postprocess_ex); numNeighbors = 0:
  /// no smoothing, > 1: a wider neighbourhood built on the device.  Empty vector on error.
  [[nodiscard]] std::vector<NumericType> getDeviceFlux(NormalizationType norm, int numNeighbors,
                                                       int particle = 0) {
    std::vector<float> out(numPoints()), areas;
    if (!this->ctx_ || (norm == NormalizationType::SOURCE && !this->haveSource_))
      return {};
    computeDiskAreas();
    areas.assign(diskAreas_.begin(), diskAreas_.end());
    const bool max = norm == NormalizationType::MAX;
    // the factors as the reference forms them (rayTraceDisk.hpp:111,129-133)
    const double factor =
        max ? static_cast<double>(static_cast<NumericType>(diskRadius_ * diskRadius_)) * M_PI
            : static_cast<double>(static_cast<float>(this->sourceArea_) /
                                  static_cast<float>(this->totalRays()));
    if (vr_flux_postprocess_ex(this->ctx_, particle, areas.data(), max ? VR_NORM_MAX : VR_NORM_SOURCE,
                               factor, numNeighbors, static_cast<float>(diskRadius_),
                               out.data()) != VR_OK)
      return {};
    return std::vector<NumericType>(out.begin(), out.end());
  }

  // introspection used by the parity tests (GeometryDisk getters)
  [[nodiscard]] std::size_t numPoints() const { return xyzr_.size() / 4; }
  [[nodiscard]] const std::vector<double> &getDiskAreas() {
    computeDiskAreas();
    return diskAreas_;
  }
  [[nodiscard]] std::vector<std::uint32_t> getNeighborIndices(std::size_t i) const {
    return {nbIdx_.begin() + nbOff_[i], nbIdx_.begin() + nbOff_[i + 1]};
  }
  [[nodiscard]] std::array<std::array<float, 3>, 2> getBoundingBox() const { return bbox_; }

private:
  static void buildNeighbors(const std::vector<float> &pts3, float distance,
                             std::vector<std::uint32_t> &off, std::vector<std::uint32_t> &idx) {
    std::uint32_t *o = nullptr, *x = nullptr;
    const std::uint32_t n = static_cast<std::uint32_t>(pts3.size() / 3);
    vr_build_neighbors(D, pts3.data(), n, distance, &o, &x);
    off.assign(o, o + n + 1);
    idx.assign(x, x + off[n]);
    vr_free(o);
    vr_free(x);
  }

  // rayGeometryDisk.hpp:266-354 (areas against the UNadjusted geometry box)
  void computeDiskAreas() {
    if (diskAreas_.size() == numPoints())
      return;
    const auto st = rayInternal::getTraceSettings(this->sourceDirection_);
    const int d0 = st[1], d1 = st[2];
    const auto c0 = this->boundaryConditions_[d0];
    const auto c1 = D == 3 ? this->boundaryConditions_[d1] : BoundaryCondition::IGNORE_BOUNDARY;
    const double r = diskRadius_;
    diskAreas_.assign(numPoints(), 0.);
    for (std::size_t i = 0; i < numPoints(); ++i) {
      const float *p = &xyzr_[4 * i], *nv = &normals_[3 * i];
      if (D == 3) {
        double area = r * r * M_PI;
        if (c0 == BoundaryCondition::IGNORE_BOUNDARY && c1 == BoundaryCondition::IGNORE_BOUNDARY) {
          // open domain: whole disk
        } else if (d0 != 2 && d1 != 2) {
          area = rayInternal::diskAreaInsidePrism(p, nv, r, 0, 1, bbox_[0][0], bbox_[1][0], bbox_[0][1],
                                                  bbox_[1][1]);
        } else {
          const double eps = 1e-3;
          if (std::fabs(p[d0] - bbox_[0][d0]) < eps || std::fabs(p[d0] - bbox_[1][d0]) < eps)
            area /= 2;
          if (std::fabs(p[d1] - bbox_[0][d1]) < eps || std::fabs(p[d1] - bbox_[1][d1]) < eps)
            area /= 2;
        }
        diskAreas_[i] = area;
      } else {
        double len = 2 * r;
        if (c0 != BoundaryCondition::IGNORE_BOUNDARY) {
          for (int side = 0; side < 2; ++side) {
            const double dist = std::fabs(double(p[d0]) - bbox_[side][d0]);
            const double s2 = 1. - double(nv[d0]) * nv[d0];
            if (dist < r && s2 > 1e-4) {
              const double inside = dist / std::sqrt(s2);
              if (inside < r)
                len -= r - inside;
            }
          }
        }
        diskAreas_[i] = len;
      }
    }
  }

  // rayTraceDisk.hpp:196-217: reports and flags, does not stop the call
  void checkSettings() {
    if (this->pParticle_ == nullptr) {
      this->RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("No particle was specified in rayTrace. Aborting.");
    }
    if (xyzr_.empty()) {
      this->RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("No geometry was passed to rayTrace. Aborting.");
    }
    if (D == 2 && (this->sourceDirection_ == TraceDirection::POS_Z ||
                   this->sourceDirection_ == TraceDirection::NEG_Z)) {
      this->RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("Invalid source direction in 2D geometry. Aborting.");
    }
    if (diskRadius_ > this->gridDelta_) {
      this->RTInfo_.warning = true;
      VIENNACORE_LOG_WARNING("Disk radius should be smaller than grid delta. Hit count "
                             "normalization not correct.");
    }
  }

  void setupChanged() override { diskAreas_.clear(); }

  std::vector<float> xyzr_, normals_;
  std::vector<std::uint32_t> nbOff_, nbIdx_;
  std::vector<std::int32_t> materialIds_;
  std::vector<double> diskAreas_;
  std::array<std::array<float, 3>, 2> bbox_{};
  NumericType diskRadius_ = 0;
  bool sceneDirty_ = true, geometryOnDevice_ = false;
  std::uint32_t zero_ = 0;
};

// ---- TraceTriangle (rayTraceTriangle.hpp:12-152, rayGeometryTriangle.hpp) -------
template <class NumericType, int D> class TraceTriangle final : public Trace<NumericType, D> {
public:
  TraceTriangle() = default;
  ~TraceTriangle() override = default;

  void apply() override {
    checkSettings();  // rayTraceTriangle.hpp:19-20: reports and flags, the call goes on
    const bool dirty = sceneDirty_;
    bool uploaded = !tris_.empty();
    if (uploaded && dirty) {
      if (!this->ctx_ ||
          vr_scene_set_triangles(this->ctx_, verts_.data(), static_cast<std::uint32_t>(verts_.size() / 3),
                                 tris_.data(), static_cast<std::uint32_t>(tris_.size() / 3),
                                 normals_.data(),
                                 materialIds_.empty() ? nullptr : materialIds_.data()) != VR_OK) {
        this->RTInfo_.error = true;
        VIENNACORE_LOG_ERROR(std::string("Geometry upload failed: ") +
                             (this->ctx_ ? vr_last_error(this->ctx_) : this->createError_.c_str()));
        uploaded = false;
      }
    }
    if (uploaded &&
        this->traceCommitted(bbox_, static_cast<float>(this->gridDelta_), tris_.size() / 3, dirty)) {
      sceneDirty_ = false;
      this->haveSource_ = true;
    } else {
      ++this->config_.runNumber;  // rayTraceTriangle.hpp:58 counts every apply()
    }
  }

  void setGeometry(std::vector<VectorType<NumericType, 3>> const &points,
                   std::vector<VectorType<unsigned, 3>> const &triangles,
                   const NumericType gridDelta) {
    this->gridDelta_ = gridDelta;
    verts_.resize(3 * points.size());
    tris_.resize(3 * triangles.size());
    normals_.resize(3 * triangles.size());
    areas_.resize(triangles.size());
    for (int a = 0; a < 3; ++a) {
      bbox_[0][a] = std::numeric_limits<float>::max();
      bbox_[1][a] = std::numeric_limits<float>::lowest();
    }
    for (std::size_t i = 0; i < points.size(); ++i)
      for (int a = 0; a < 3; ++a) {
        const float v = static_cast<float>(points[i][a]);
        verts_[3 * i + a] = v;
        bbox_[0][a] = std::min(bbox_[0][a], v);
        bbox_[1][a] = std::max(bbox_[1][a], v);
      }
    for (std::size_t i = 0; i < triangles.size(); ++i) {
      for (int k = 0; k < 3; ++k)
        tris_[3 * i + k] = triangles[i][k];
      const float *p0 = &verts_[3 * tris_[3 * i]], *p1 = &verts_[3 * tris_[3 * i + 1]],
                  *p2 = &verts_[3 * tris_[3 * i + 2]];
      const float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
      const float b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
      float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
      const float len = std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
      areas_[i] = 0.5 * len; // rayGeometryTriangle.hpp: area = |cross| / 2
      const float inv = 1.0f / len;
      for (int k = 0; k < 3; ++k)
        normals_[3 * i + k] = c[k] * inv;
    }
    materialIds_.clear();
    sceneDirty_ = true;
  }

  void setGeometry(const TriangleMesh &mesh) {
    std::vector<VectorType<NumericType, 3>> pts(mesh.nodes.size());
    for (std::size_t i = 0; i < pts.size(); ++i)
      pts[i] = {static_cast<NumericType>(mesh.nodes[i][0]), static_cast<NumericType>(mesh.nodes[i][1]),
                static_cast<NumericType>(mesh.nodes[i][2])};
    setGeometry(pts, mesh.triangles, static_cast<NumericType>(mesh.gridDelta));
  }

  template <typename T> void setMaterialIds(std::vector<T> const &materialIds) {
    materialIds_.assign(materialIds.begin(), materialIds.end());
    sceneDirty_ = true;
  }

  // rayTraceTriangle.hpp:92-130
  void normalizeFlux(std::vector<NumericType> &flux,
                     NormalizationType norm = NormalizationType::SOURCE) override {
    assert(flux.size() == areas_.size() && "Unequal number of points in normalizeFlux");
    if (norm == NormalizationType::MAX) {
      const NumericType maxv = *std::max_element(flux.begin(), flux.end());
      for (std::size_t i = 0; i < flux.size(); ++i)
        flux[i] /= maxv * static_cast<NumericType>(areas_[i]);
    } else if (norm == NormalizationType::SOURCE) {
      if (!this->haveSource_) {
        VIENNACORE_LOG_WARNING("No source was specified in rayTrace for the normalization.");
        return;
      }
      const NumericType normFactor =
          static_cast<NumericType>(this->sourceArea_) / static_cast<NumericType>(this->totalRays());
      for (std::size_t i = 0; i < flux.size(); ++i)
        flux[i] *= normFactor / static_cast<NumericType>(areas_[i]);
    }
  }

  void smoothFlux(std::vector<NumericType> &, int) override {} // no smoothing on elements

private:
  // rayTraceTriangle.hpp:129-138
  void checkSettings() {
    if (this->pParticle_ == nullptr) {
      this->RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("No particle was specified in rayTrace. Aborting.");
    }
    if (tris_.empty()) {
      this->RTInfo_.error = true;
      VIENNACORE_LOG_ERROR("No geometry was passed to rayTrace. Aborting.");
    }
  }

  std::vector<float> verts_, normals_;
  std::vector<std::uint32_t> tris_;
  std::vector<std::int32_t> materialIds_;
  std::vector<double> areas_;
  std::array<std::array<float, 3>, 2> bbox_{};
  bool sceneDirty_ = true;
};

} // namespace viennaray
