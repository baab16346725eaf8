// Host-side helpers of the ViennaRay interface that callers include by their reference
// header names: rayReflection.hpp, raySourceRandom.hpp, rayPointNeighborhood.hpp,
// rayGeometry.hpp, rayGeometryDisk.hpp, rayGeometryTriangle.hpp.  They exist so that code
// written against the reference (ViennaPS particle sources include rayReflection.hpp, for
// one) compiles against this drop-in.  None of it is on the device path: the device runs
// its own functors (viennaray_b200/csrc/vr_device.cuh); these are the plain host versions a
// user-written host routine may call.
#pragma once
#include "vr_host.hpp"

namespace rayInternal {
// uniformly distributed point on the unit sphere (Marsaglia 1972) -- rayUtil.hpp:266-283
template <typename NumericType> Vec3D<NumericType> pickRandomPointOnUnitSphere(RNG &rng) {
  std::uniform_real_distribution<NumericType> uni(-1, 1);
  NumericType a, b, q;
  do {
    a = uni(rng);
    b = uni(rng);
    q = a * a + b * b;
  } while (q >= 1);
  const NumericType root = 2 * std::sqrt(1 - q);
  return Vec3D<NumericType>{a * root, b * root, 1 - 2 * q};
}
} // namespace rayInternal

namespace viennaray {

/// Mirror reflection of rayDir at the surface -- rayReflection.hpp:13-29.
template <typename NumericType, int D = 3>
[[nodiscard]] Vec3D<NumericType> ReflectionSpecular(const Vec3D<NumericType> &rayDir,
                                                    const Vec3D<NumericType> &geomNormal) {
  NumericType k = 0;
  for (int i = 0; i < 3; ++i)
    k += geomNormal[i] * rayDir[i];
  Vec3D<NumericType> out;
  for (int i = 0; i < 3; ++i)
    out[i] = rayDir[i] - 2 * k * geomNormal[i];
  return out;
}

/// Cosine-distributed direction about the normal: normal + random unit vector, normalised
/// (z dropped in 2D) -- rayReflection.hpp:32-50.
template <typename NumericType, int D>
[[nodiscard]] Vec3D<NumericType> ReflectionDiffuse(const Vec3D<NumericType> &geomNormal,
                                                   RNG &rngState) {
  auto out = rayInternal::pickRandomPointOnUnitSphere<NumericType>(rngState);
  for (int i = 0; i < 3; ++i)
    out[i] += geomNormal[i];
  if (D == 2)
    out[2] = 0;
  NumericType len = 0;
  for (int i = 0; i < 3; ++i)
    len += out[i] * out[i];
  len = std::sqrt(len);
  for (int i = 0; i < 3; ++i)
    out[i] /= len;
  return out;
}

/// Direction inside a cone of half-angle maxConeAngle about the specular direction, polar
/// angle by accept-reject of a cosine-shaped density, flipped into the upper half space of
/// the normal -- rayReflection.hpp:52-120.  Cone <= 0: specular; >= pi/2: diffuse.
template <typename NumericType, int D>
[[nodiscard]] Vec3D<NumericType>
ReflectionConedCosine(const Vec3D<NumericType> &rayDir, const Vec3D<NumericType> &geomNormal,
                      RNG &rngState, NumericType maxConeAngle) {
  if (maxConeAngle <= 0)
    return ReflectionSpecular<NumericType, D>(rayDir, geomNormal);
  if (maxConeAngle >= NumericType(M_PI / 2))
    return ReflectionDiffuse<NumericType, D>(geomNormal, rngState);
  auto w = ReflectionSpecular<NumericType, D>(rayDir, geomNormal);
  Normalize(w);
  // branch-free orthonormal frame about w (Frisvad 2012)
  Vec3D<NumericType> t, b;
  if (w[2] < NumericType(-0.999999)) {
    t = {0, -1, 0};
    b = {-1, 0, 0};
  } else {
    const NumericType a = 1 / (1 + w[2]), m = -w[0] * w[1] * a;
    t = {1 - w[0] * w[0] * a, m, -w[0]};
    b = {m, 1 - w[1] * w[1] * a, -w[1]};
  }
  std::uniform_real_distribution<NumericType> uni(0, 1);
  NumericType theta;
  for (;;) {
    const NumericType u = std::sqrt(uni(rngState));
    const NumericType s = std::sqrt(std::max(NumericType(0), 1 - u));
    theta = maxConeAngle * s;
    if (uni(rngState) * theta * u <= std::cos(NumericType(M_PI / 2) * s) * std::sin(theta))
      break;
  }
  const NumericType phi = NumericType(2 * M_PI) * uni(rngState);
  const NumericType sn = std::sin(theta), cs = std::cos(theta), cp = std::cos(phi), sp = std::sin(phi);
  Vec3D<NumericType> out;
  NumericType along = 0;
  for (int i = 0; i < 3; ++i) {
    out[i] = sn * (cp * t[i] + sp * b[i]) + cs * w[i];
    along += out[i] * geomNormal[i];
  }
  if (along <= 0)
    for (int i = 0; i < 3; ++i)
      out[i] -= 2 * along * geomNormal[i];
  if (D == 2)
    out[2] = 0;
  Normalize(out);
  return out;
}

/// The default source: uniform origins on the source plane, power-cosine directions about
/// the plane normal or about a tilted primary direction -- raySourceRandom.hpp:10-132.
/// Given to Trace::setSource it maps to the device's built-in source.
template <typename NumericType, int D> class SourceRandom : public Source<NumericType> {
  using boundingBoxType = std::array<Vec3D<NumericType>, 2>;
  const boundingBoxType bdBox_;
  const std::array<int, 5> settings_;  // rayDir, firstDir, secondDir, minMax, posNeg
  const NumericType cosinePower_, ee_;
  const std::size_t numPoints_;
  const bool customDirection_;
  const std::array<Vec3D<NumericType>, 3> basis_;

public:
  SourceRandom(const boundingBoxType &boundingBox, NumericType cosinePower,
               std::array<int, 5> &pTraceSettings, const std::size_t numPoints,
               const bool customDirection,
               const std::array<Vec3D<NumericType>, 3> &orthonormalBasis)
      : bdBox_(boundingBox), settings_(pTraceSettings), cosinePower_(cosinePower),
        ee_(NumericType(1) / (cosinePower + 1)), numPoints_(numPoints),
        customDirection_(customDirection), basis_(orthonormalBasis) {}

  std::array<Vec3D<NumericType>, 2> getOriginAndDirection(const std::size_t,
                                                          RNG &rngState) const override {
    std::uniform_real_distribution<NumericType> uni;
    const int rd = settings_[0], fd = settings_[1], sd = settings_[2];
    Vec3D<NumericType> origin{0, 0, 0};
    origin[rd] = bdBox_[settings_[3]][rd];
    origin[fd] = bdBox_[0][fd] + (bdBox_[1][fd] - bdBox_[0][fd]) * uni(rngState);
    origin[sd] = D == 2 ? NumericType(0)
                        : bdBox_[0][sd] + (bdBox_[1][sd] - bdBox_[0][sd]) * uni(rngState);
    Vec3D<NumericType> dir{0, 0, 0};
    for (;;) {
      const NumericType phi = NumericType(2 * M_PI) * uni(rngState);
      const NumericType ct = std::pow(uni(rngState), ee_), st = std::sqrt(1 - ct * ct);
      if (!customDirection_) {
        dir[rd] = settings_[4] * ct;
        dir[fd] = std::cos(phi) * st;
        dir[sd] = std::sin(phi) * st;
        break;
      }
      const NumericType r[3] = {ct, std::cos(phi) * st, std::sin(phi) * st};
      for (int i = 0; i < 3; ++i)
        dir[i] = basis_[0][i] * r[0] + basis_[1][i] * r[1] + basis_[2][i] * r[2];
      // keep only directions that leave the source plane towards the geometry
      if (!((settings_[4] < 0 && dir[rd] > 0) || (settings_[4] > 0 && dir[rd] < 0)))
        break;
    }
    if (D == 2) {
      dir[2] = 0;
      Normalize(dir);
    }
    return {origin, dir};
  }
  [[nodiscard]] std::size_t getNumPoints() const override { return numPoints_; }
  NumericType getSourceArea() const override {
    const NumericType a = bdBox_[1][settings_[1]] - bdBox_[0][settings_[1]];
    return D == 2 ? a : a * (bdBox_[1][settings_[2]] - bdBox_[0][settings_[2]]);
  }
  bool deviceSource(vr_source_desc &d, std::vector<float> &) const override {
    for (int a = 0; a < 3; ++a) {
      d.bboxMin[a] = static_cast<float>(bdBox_[0][a]);
      d.bboxMax[a] = static_cast<float>(bdBox_[1][a]);
    }
    d.rayDir = settings_[0];
    d.firstDir = settings_[1];
    d.secondDir = settings_[2];
    d.minMax = settings_[3];
    d.posNeg = static_cast<float>(settings_[4]);
    d.useBasis = customDirection_ ? 1 : 0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        d.basis[3 * r + c] = static_cast<float>(basis_[r][c]);
    d.useGrid = 0;
    return true;
  }
};

/// Neighbour lists of a point cloud: j is a neighbour of i when no coordinate differs by
/// more than `distance` and the Euclidean distance is at most `distance` --
/// rayPointNeighborhood.hpp:14-120,287-298.  Built by the library's helper
/// (vr_build_neighbors), rows ascending.
template <typename NumericType, int D> class PointNeighborhood {
  std::vector<std::vector<unsigned int>> rows_;
  NumericType distance_ = 0;

public:
  PointNeighborhood() = default;
  template <std::size_t Dim>
  void init(std::vector<VectorType<NumericType, Dim>> const &points, NumericType distance,
            Vec3D<NumericType> const & /*minCoords*/, Vec3D<NumericType> const & /*maxCoords*/) {
    distance_ = distance;
    const std::size_t n = points.size();
    std::vector<float> p(3 * n, 0.f);
    for (std::size_t i = 0; i < n; ++i)
      for (std::size_t a = 0; a < Dim; ++a)
        p[3 * i + a] = static_cast<float>(points[i][a]);
    std::uint32_t *off = nullptr, *idx = nullptr;
    rows_.assign(n, {});
    if (vr_build_neighbors(D, p.data(), static_cast<std::uint32_t>(n), static_cast<float>(distance),
                           &off, &idx) != VR_OK)
      return;
    for (std::size_t i = 0; i < n; ++i)
      rows_[i].assign(idx + off[i], idx + off[i + 1]);
    vr_free(off);
    vr_free(idx);
  }
  [[nodiscard]] std::vector<unsigned int> const &getNeighborIndices(const unsigned int idx) const {
    return rows_[idx];
  }
  [[nodiscard]] std::size_t getNumPoints() const { return rows_.size(); }
  [[nodiscard]] NumericType getDistance() const { return distance_; }
};

/// What the reference's Geometry base keeps beside the Embree buffers: the primitive count
/// and the material IDs -- rayGeometry.hpp:10-72.
template <typename NumericType, int D> class Geometry {
public:
  virtual ~Geometry() = default;
  template <typename MatIdType> void setMaterialIds(std::vector<MatIdType> const &ids) {
    materialIds_.assign(ids.begin(), ids.end());
  }
  [[nodiscard]] std::size_t getNumPrimitives() const { return numPrimitives_; }
  [[nodiscard]] int getMaterialId(const unsigned int primID) const {
    return primID < materialIds_.size() ? materialIds_[primID] : 0;
  }
  [[nodiscard]] bool checkGeometryEmpty() const { return numPrimitives_ == 0; }

protected:
  std::size_t numPrimitives_ = 0;
  std::vector<int> materialIds_;
};

/// Host container of a disk cloud with the getters of rayGeometryDisk.hpp:196-262 (points,
/// normals, neighbours, bounding box).  The device copy lives in the vr_ctx of a TraceDisk;
/// this class is for host code that wants the same view of the data.
template <typename NumericType, int D> class GeometryDisk : public Geometry<NumericType, D> {
  std::vector<std::array<float, 4>> disks_;
  std::vector<Vec3D<NumericType>> normals_;
  PointNeighborhood<NumericType, D> neighbors_;
  std::array<Vec3D<NumericType>, 2> bbox_{};

public:
  template <std::size_t Dim>
  void initGeometry(std::vector<VectorType<NumericType, Dim>> const &points,
                    std::vector<VectorType<NumericType, Dim>> const &normals,
                    NumericType const discRadii) {
    const std::size_t n = points.size();
    this->numPrimitives_ = n;
    disks_.assign(n, {0.f, 0.f, 0.f, static_cast<float>(discRadii)});
    normals_.assign(n, Vec3D<NumericType>{0, 0, 0});
    for (int a = 0; a < 3; ++a) {
      bbox_[0][a] = a < D ? std::numeric_limits<NumericType>::max() : NumericType(0);
      bbox_[1][a] = a < D ? std::numeric_limits<NumericType>::lowest() : NumericType(0);
    }
    for (std::size_t i = 0; i < n; ++i)
      for (std::size_t a = 0; a < Dim && static_cast<int>(a) < D; ++a) {
        disks_[i][a] = static_cast<float>(points[i][a]);
        normals_[i][a] = normals[i][a];
        bbox_[0][a] = std::min(bbox_[0][a], points[i][a]);
        bbox_[1][a] = std::max(bbox_[1][a], points[i][a]);
      }
    if (this->materialIds_.size() != n)
      this->materialIds_.assign(n, 0);
    neighbors_.template init<Dim>(points, 2 * discRadii, bbox_[0], bbox_[1]);
  }
  [[nodiscard]] std::array<Vec3D<NumericType>, 2> getBoundingBox() const { return bbox_; }
  [[nodiscard]] Vec3D<NumericType> getPoint(const unsigned int primID) const {
    return {NumericType(disks_[primID][0]), NumericType(disks_[primID][1]),
            NumericType(disks_[primID][2])};
  }
  [[nodiscard]] std::array<float, 4> const &getPrimRef(unsigned int primID) const {
    return disks_[primID];
  }
  [[nodiscard]] Vec3D<NumericType> const &getPrimNormal(const unsigned int primID) const {
    return normals_[primID];
  }
  [[nodiscard]] NumericType getDiscRadius() const {
    return disks_.empty() ? NumericType(0) : NumericType(disks_[0][3]);
  }
  [[nodiscard]] std::vector<unsigned int> const &getNeighborIndices(const unsigned int idx) const {
    return neighbors_.getNeighborIndices(idx);
  }
  [[nodiscard]] PointNeighborhood<NumericType, D> const &getPointNeighborhood() const {
    return neighbors_;
  }
};

/// Host container of a triangle mesh with the getters of rayGeometryTriangle.hpp:94-160.
template <typename NumericType, int D> class GeometryTriangle : public Geometry<NumericType, D> {
  std::vector<Vec3D<NumericType>> points_, normals_;
  std::vector<VectorType<unsigned, 3>> triangles_;
  std::vector<NumericType> areas_;

public:
  void initGeometry(std::vector<Vec3D<NumericType>> const &points,
                    std::vector<VectorType<unsigned, 3>> const &triangles) {
    points_ = points;
    triangles_ = triangles;
    this->numPrimitives_ = triangles.size();
    normals_.resize(triangles.size());
    areas_.resize(triangles.size());
    for (std::size_t i = 0; i < triangles.size(); ++i) {
      const auto &a = points[triangles[i][0]], &b = points[triangles[i][1]],
                 &c = points[triangles[i][2]];
      const Vec3D<NumericType> u{b[0] - a[0], b[1] - a[1], b[2] - a[2]},
          v{c[0] - a[0], c[1] - a[1], c[2] - a[2]};
      Vec3D<NumericType> nrm{u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2],
                             u[0] * v[1] - u[1] * v[0]};
      const NumericType len = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      areas_[i] = len / 2;
      for (int k = 0; k < 3; ++k)
        nrm[k] = len > 0 ? nrm[k] / len : NumericType(0);
      normals_[i] = nrm;
    }
    if (this->materialIds_.size() != triangles.size())
      this->materialIds_.assign(triangles.size(), 0);
  }
  [[nodiscard]] Vec3D<NumericType> const &getPrimNormal(const unsigned int primID) const {
    return normals_[primID];
  }
  [[nodiscard]] NumericType getPrimArea(const unsigned int primID) const { return areas_[primID]; }
  [[nodiscard]] VectorType<unsigned, 3> const &getTriangle(const unsigned int primID) const {
    return triangles_[primID];
  }
  [[nodiscard]] Vec3D<NumericType> const &getPoint(const unsigned int idx) const {
    return points_[idx];
  }
};

} // namespace viennaray
