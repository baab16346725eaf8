// TEST INFRASTRUCTURE (oracle/): stand-in for ViennaCore 2.1.2's
// vcVectorType.hpp, which the reference fetches at configure time
// (/root/reference/CMakeLists.txt:104-108) and which is absent here.  Only
// the names the reference headers/tests use are provided; semantics are the
// obvious ones (UNVERIFIED against upstream ViennaCore bit-for-bit).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <iostream>

namespace viennacore {

template <class T, size_t D> using VectorType = std::array<T, D>;
template <class T> using Vec2D = VectorType<T, 2>;
template <class T> using Vec3D = VectorType<T, 3>;
using Vec2Df = Vec2D<float>;
using Vec3Df = Vec3D<float>;

template <class T, size_t D>
VectorType<T, D> operator+(const VectorType<T, D> &a,
                           const VectorType<T, D> &b) {
  VectorType<T, D> r;
  for (size_t i = 0; i < D; ++i)
    r[i] = a[i] + b[i];
  return r;
}
template <class T, size_t D>
VectorType<T, D> operator-(const VectorType<T, D> &a,
                           const VectorType<T, D> &b) {
  VectorType<T, D> r;
  for (size_t i = 0; i < D; ++i)
    r[i] = a[i] - b[i];
  return r;
}
template <class T, size_t D, class S,
          class = std::enable_if_t<std::is_arithmetic_v<S>>>
VectorType<T, D> operator*(const VectorType<T, D> &a, S s) {
  VectorType<T, D> r;
  for (size_t i = 0; i < D; ++i)
    r[i] = a[i] * static_cast<T>(s);
  return r;
}
template <class T, size_t D, class S,
          class = std::enable_if_t<std::is_arithmetic_v<S>>>
VectorType<T, D> operator*(S s, const VectorType<T, D> &a) {
  return a * s;
}
template <class T, size_t D, class S,
          class = std::enable_if_t<std::is_arithmetic_v<S>>>
VectorType<T, D> operator/(const VectorType<T, D> &a, S s) {
  VectorType<T, D> r;
  for (size_t i = 0; i < D; ++i)
    r[i] = a[i] / static_cast<T>(s);
  return r;
}

template <class T, size_t D>
T DotProduct(const VectorType<T, D> &a, const VectorType<T, D> &b) {
  T s = 0;
  for (size_t i = 0; i < D; ++i)
    s += a[i] * b[i];
  return s;
}
template <class T>
Vec3D<T> CrossProduct(const Vec3D<T> &a, const Vec3D<T> &b) {
  return Vec3D<T>{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2],
                  a[0] * b[1] - a[1] * b[0]};
}
template <class T, size_t D> T Norm2(const VectorType<T, D> &a) {
  return DotProduct(a, a);
}
template <class T, size_t D> T Norm(const VectorType<T, D> &a) {
  return std::sqrt(Norm2(a));
}
template <class T, size_t D> void Normalize(VectorType<T, D> &a) {
  T n = Norm(a);
  if (n <= T(0))
    return;
  T inv = T(1) / n;
  for (size_t i = 0; i < D; ++i)
    a[i] *= inv;
}
template <class T, size_t D>
VectorType<T, D> Normalize(const VectorType<T, D> &a) {
  VectorType<T, D> r = a;
  Normalize(r);
  return r;
}
template <class T, size_t D> bool IsNormalized(const VectorType<T, D> &a) {
  return std::fabs(Norm(a) - T(1)) < T(1e-4);
}
template <class T, size_t D> VectorType<T, D> Inv(const VectorType<T, D> &a) {
  VectorType<T, D> r;
  for (size_t i = 0; i < D; ++i)
    r[i] = -a[i];
  return r;
}
// add + mult * fac
template <class T, size_t D>
VectorType<T, D> ScaleAdd(const VectorType<T, D> &mult,
                          const VectorType<T, D> &add, T fac) {
  VectorType<T, D> r;
  for (size_t i = 0; i < D; ++i)
    r[i] = add[i] + mult[i] * fac;
  return r;
}
template <class T, size_t D>
T Distance(const VectorType<T, D> &a, const VectorType<T, D> &b) {
  return Norm(a - b);
}
// Sum(v1, v2, ...) = element-wise sum of the arguments
template <class T, size_t D>
VectorType<T, D> Sum(const VectorType<T, D> &a, const VectorType<T, D> &b) {
  return a + b;
}
template <class T, size_t D, class... Rest>
VectorType<T, D> Sum(const VectorType<T, D> &a, const VectorType<T, D> &b,
                     const Rest &...rest) {
  return Sum(a + b, rest...);
}
template <class T, size_t D>
std::ostream &operator<<(std::ostream &os, const VectorType<T, D> &v) {
  os << "[";
  for (size_t i = 0; i < D; ++i)
    os << v[i] << (i + 1 < D ? ", " : "]");
  return os;
}
// plane normal of three points (not normalised)
template <class T> Vec3D<T> ComputeNormal(const Vec3D<Vec3D<T>> &p) {
  return CrossProduct(p[1] - p[0], p[2] - p[0]);
}
template <class T, size_t D>
void PrintBoundingBox(const std::array<VectorType<T, D>, 2> &bb) {
  for (int k = 0; k < 2; ++k) {
    std::cout << (k ? "max" : "min") << ":";
    for (size_t i = 0; i < D; ++i)
      std::cout << " " << bb[k][i];
    std::cout << "\n";
  }
}
} // namespace viennacore
