// TEST INFRASTRUCTURE (oracle/) -- never linked into the product library.
//
// mini_rtc: a SUBSTITUTE for the Embree 4 entry points the reference headers
// call (declared in oracle/stubs/embree4/rtcore.h).  Embree 4.3.3 is fetched
// by the reference at configure time (/root/reference/CMakeLists.txt:37-41,
// 124-129) and is neither vendored nor installed here, so this file restates
// the *published* per-primitive algorithms of Embree's single-ray kernels and
// fixes an explicit, traversal-order-independent closest-hit rule:
//
//   oriented disc  (Embree kernels/geometry/disc_intersector.h, oriented
//                   variant):  den = dot(dir,n); den == 0 -> miss;
//                   t = dot(c-org,n)/den; tnear <= t <= tfar;
//                   |org + dir*t - c|^2 < r^2.            No backface culling.
//   triangle       (Embree Moeller-Trumbore, kernels/geometry/
//                   triangle_intersector_moeller.h):  e1=v0-v1, e2=v2-v0,
//                   Ng=cross(e2,e1) (= (v1-v0)x(v2-v0)), C=v0-org,
//                   R=cross(C,dir), den=dot(Ng,dir), U=dot(R,e2)^sgn,
//                   V=dot(R,e1)^sgn, U>=0, V>=0, U+V<=|den|,
//                   T=dot(Ng,C)^sgn, |den|*tnear < T <= |den|*tfar, t=T/|den|.
//   closest hit    smallest float t; ties -> smallest geomID, then primID.
//
// All float arithmetic is unfused, left-to-right (compile with
// -ffp-contract=off), dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z.  Real Embree
// uses FMA and a packet-order tie rule, so hit IDs on exactly coplanar
// overlapping discs are "parity unpinned" against real Embree; they are pinned
// against this rule, which oracle/vr_oracle.c and the CUDA kernels restate.
// Golden vectors: tests/intersectionTest/intersectionTest.cpp:91-92,126-127.

#include <embree4/rtcore.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

struct MiniRtcDevice {
  int dummy;
};

struct MiniRtcGeometry {
  RTCGeometryType type;
  std::atomic<int> refs{1};
  void *vertex = nullptr; // float3 / float4
  size_t vertexStride = 0, vertexCount = 0;
  void *index = nullptr; // uint3
  size_t indexCount = 0;
  void *normal = nullptr; // float3
  size_t normalCount = 0;
  ~MiniRtcGeometry() {
    free(vertex);
    free(index);
    free(normal);
  }
};

namespace {
struct Node {
  float lo[3];
  float hi[3];
  uint32_t left;  // inner: index of left child (right = left+1); leaf: first
  uint32_t count; // 0 = inner
};
struct PrimRef {
  float lo[3], hi[3];
  uint32_t geom, prim;
};
} // namespace

struct MiniRtcScene {
  std::vector<MiniRtcGeometry *> geoms;
  std::vector<Node> nodes;
  std::vector<PrimRef> prims; // leaf order
  std::mutex mtx;
  std::atomic<bool> committed{false};
};

namespace {

inline float dot3(const float *a, const float *b) {
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
inline void cross3(const float *a, const float *b, float *r) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
inline float xorSign(float v, uint32_t sgn) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u ^= sgn;
  memcpy(&v, &u, 4);
  return v;
}

void primBounds(const MiniRtcGeometry *g, uint32_t i, float *lo, float *hi) {
  if (g->type == RTC_GEOMETRY_TYPE_ORIENTED_DISC_POINT) {
    const float *c =
        reinterpret_cast<const float *>((const char *)g->vertex + i * g->vertexStride);
    const float *n = reinterpret_cast<const float *>((const char *)g->normal + i * 12);
    const float r = c[3];
    float nn = dot3(n, n);
    for (int a = 0; a < 3; ++a) {
      // half extent of an oriented disc along axis a: r*sqrt(1 - n_a^2/|n|^2)
      float f = nn > 0.f ? 1.f - n[a] * n[a] / nn : 1.f;
      float e = r * std::sqrt(std::max(f, 0.f));
      float pad = 1e-4f * r + 4e-7f * std::fabs(c[a]);
      lo[a] = c[a] - e - pad;
      hi[a] = c[a] + e + pad;
    }
  } else {
    const uint32_t *t = reinterpret_cast<const uint32_t *>((const char *)g->index + i * 12);
    for (int a = 0; a < 3; ++a) {
      lo[a] = FLT_MAX;
      hi[a] = -FLT_MAX;
    }
    for (int k = 0; k < 3; ++k) {
      const float *v =
          reinterpret_cast<const float *>((const char *)g->vertex + t[k] * g->vertexStride);
      for (int a = 0; a < 3; ++a) {
        lo[a] = std::min(lo[a], v[a]);
        hi[a] = std::max(hi[a], v[a]);
      }
    }
    for (int a = 0; a < 3; ++a) {
      float pad = 1e-5f * (hi[a] - lo[a]) + 4e-7f * std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 1e-30f;
      lo[a] -= pad;
      hi[a] += pad;
    }
  }
}

struct Builder {
  std::vector<PrimRef> &prims;
  std::vector<Node> &nodes;
  static constexpr int kBins = 16;
  static constexpr uint32_t kLeaf = 4;

  static float area(const float *lo, const float *hi) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.f * (dx * dy + dy * dz + dz * dx);
  }

  void build(uint32_t nodeIdx, uint32_t first, uint32_t count) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = first; i < first + count; ++i) {
      for (int a = 0; a < 3; ++a) {
        lo[a] = std::min(lo[a], prims[i].lo[a]);
        hi[a] = std::max(hi[a], prims[i].hi[a]);
        float c = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
        clo[a] = std::min(clo[a], c);
        chi[a] = std::max(chi[a], c);
      }
    }
    memcpy(nodes[nodeIdx].lo, lo, 12);
    memcpy(nodes[nodeIdx].hi, hi, 12);
    if (count <= kLeaf) {
      nodes[nodeIdx].left = first;
      nodes[nodeIdx].count = count;
      return;
    }
    // binned SAH over the centroid box
    int bestAxis = -1, bestSplit = -1;
    float bestCost = FLT_MAX;
    for (int a = 0; a < 3; ++a) {
      float ext = chi[a] - clo[a];
      if (!(ext > 0.f))
        continue;
      float scale = kBins / ext;
      uint32_t cnt[kBins] = {0};
      float blo[kBins][3], bhi[kBins][3];
      for (int b = 0; b < kBins; ++b)
        for (int k = 0; k < 3; ++k) {
          blo[b][k] = FLT_MAX;
          bhi[b][k] = -FLT_MAX;
        }
      for (uint32_t i = first; i < first + count; ++i) {
        float c = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
        int b = std::min(kBins - 1, std::max(0, (int)((c - clo[a]) * scale)));
        cnt[b]++;
        for (int k = 0; k < 3; ++k) {
          blo[b][k] = std::min(blo[b][k], prims[i].lo[k]);
          bhi[b][k] = std::max(bhi[b][k], prims[i].hi[k]);
        }
      }
      float rArea[kBins];
      uint32_t rCnt[kBins];
      float alo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ahi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      uint32_t c = 0;
      for (int b = kBins - 1; b > 0; --b) {
        c += cnt[b];
        for (int k = 0; k < 3; ++k) {
          alo[k] = std::min(alo[k], blo[b][k]);
          ahi[k] = std::max(ahi[k], bhi[b][k]);
        }
        rCnt[b] = c;
        rArea[b] = c ? area(alo, ahi) : 0.f;
      }
      for (int k = 0; k < 3; ++k) {
        alo[k] = FLT_MAX;
        ahi[k] = -FLT_MAX;
      }
      c = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        c += cnt[b];
        for (int k = 0; k < 3; ++k) {
          alo[k] = std::min(alo[k], blo[b][k]);
          ahi[k] = std::max(ahi[k], bhi[b][k]);
        }
        if (c == 0 || rCnt[b + 1] == 0)
          continue;
        float cost = c * area(alo, ahi) + rCnt[b + 1] * rArea[b + 1];
        if (cost < bestCost) {
          bestCost = cost;
          bestAxis = a;
          bestSplit = b;
        }
      }
    }
    uint32_t mid;
    if (bestAxis < 0) {
      mid = first + count / 2; // all centroids coincide: split by index
    } else {
      float scale = kBins / (chi[bestAxis] - clo[bestAxis]);
      float cl = clo[bestAxis];
      int a = bestAxis, s = bestSplit;
      auto it = std::partition(prims.begin() + first, prims.begin() + first + count,
                               [&](const PrimRef &p) {
                                 float c = 0.5f * (p.lo[a] + p.hi[a]);
                                 int b = std::min(kBins - 1, std::max(0, (int)((c - cl) * scale)));
                                 return b <= s;
                               });
      mid = (uint32_t)(it - prims.begin());
      if (mid == first || mid == first + count)
        mid = first + count / 2;
    }
    uint32_t left = (uint32_t)nodes.size();
    nodes.push_back(Node{});
    nodes.push_back(Node{});
    nodes[nodeIdx].left = left;
    nodes[nodeIdx].count = 0;
    build(left, first, mid - first);
    build(left + 1, mid, first + count - mid);
  }
};

struct Best {
  float t;
  uint32_t geom, prim;
  float ng[3], u, v;
};

inline bool better(float t, uint32_t geom, uint32_t prim, const Best &b) {
  if (t < b.t)
    return true;
  if (t > b.t)
    return false;
  if (geom != b.geom)
    return geom < b.geom;
  return prim < b.prim;
}

inline void testDisc(const MiniRtcGeometry *g, uint32_t geom, uint32_t prim,
                     const float *org, const float *dir, float tnear, float tfar, Best &best) {
  const float *c =
      reinterpret_cast<const float *>((const char *)g->vertex + prim * g->vertexStride);
  const float *n = reinterpret_cast<const float *>((const char *)g->normal + prim * 12);
  float den = dot3(dir, n);
  if (den == 0.f)
    return;
  float co[3] = {c[0] - org[0], c[1] - org[1], c[2] - org[2]};
  float t = dot3(co, n) / den;
  if (!(tnear <= t && t <= tfar))
    return;
  float q[3] = {(org[0] + dir[0] * t) - c[0], (org[1] + dir[1] * t) - c[1],
                (org[2] + dir[2] * t) - c[2]};
  float d2 = dot3(q, q);
  if (!(d2 < c[3] * c[3]))
    return;
  if (better(t, geom, prim, best)) {
    best.t = t;
    best.geom = geom;
    best.prim = prim;
    best.ng[0] = n[0];
    best.ng[1] = n[1];
    best.ng[2] = n[2];
    best.u = best.v = 0.f;
  }
}

inline void testTriangle(const MiniRtcGeometry *g, uint32_t geom, uint32_t prim,
                         const float *org, const float *dir, float tnear, float tfar,
                         Best &best) {
  const uint32_t *idx = reinterpret_cast<const uint32_t *>((const char *)g->index + prim * 12);
  const float *v0 = reinterpret_cast<const float *>((const char *)g->vertex + idx[0] * g->vertexStride);
  const float *v1 = reinterpret_cast<const float *>((const char *)g->vertex + idx[1] * g->vertexStride);
  const float *v2 = reinterpret_cast<const float *>((const char *)g->vertex + idx[2] * g->vertexStride);
  float e1[3] = {v0[0] - v1[0], v0[1] - v1[1], v0[2] - v1[2]};
  float e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
  float ng[3];
  cross3(e2, e1, ng);
  float C[3] = {v0[0] - org[0], v0[1] - org[1], v0[2] - org[2]};
  float R[3];
  cross3(C, dir, R);
  float den = dot3(ng, dir);
  if (den == 0.f)
    return;
  uint32_t sgn;
  memcpy(&sgn, &den, 4);
  sgn &= 0x80000000u;
  float absDen = std::fabs(den);
  float U = xorSign(dot3(R, e2), sgn);
  float V = xorSign(dot3(R, e1), sgn);
  if (!(U >= 0.f && V >= 0.f && U + V <= absDen))
    return;
  float T = xorSign(dot3(ng, C), sgn);
  if (!(absDen * tnear < T && T <= absDen * tfar))
    return;
  float t = T / absDen;
  if (better(t, geom, prim, best)) {
    best.t = t;
    best.geom = geom;
    best.prim = prim;
    best.ng[0] = ng[0];
    best.ng[1] = ng[1];
    best.ng[2] = ng[2];
    best.u = U / absDen;
    best.v = V / absDen;
  }
}

} // namespace

extern "C" {

RTCDevice rtcNewDevice(const char *) { return new MiniRtcDevice{0}; }
void rtcReleaseDevice(RTCDevice d) { delete d; }
long rtcGetDeviceProperty(RTCDevice, RTCDeviceProperty) { return 40303; }
RTCError rtcGetDeviceError(RTCDevice) { return RTC_ERROR_NONE; }

RTCScene rtcNewScene(RTCDevice) { return new MiniRtcScene(); }
void rtcSetSceneFlags(RTCScene, RTCSceneFlags) {}
void rtcSetSceneBuildQuality(RTCScene, RTCBuildQuality) {}
unsigned int rtcAttachGeometry(RTCScene s, RTCGeometry g) {
  g->refs++;
  s->geoms.push_back(g);
  return (unsigned)s->geoms.size() - 1;
}

void rtcJoinCommitScene(RTCScene s) {
  if (s->committed.load(std::memory_order_acquire))
    return;
  std::lock_guard<std::mutex> lock(s->mtx);
  if (s->committed.load(std::memory_order_relaxed))
    return;
  s->prims.clear();
  s->nodes.clear();
  for (uint32_t gi = 0; gi < s->geoms.size(); ++gi) {
    const MiniRtcGeometry *g = s->geoms[gi];
    size_t n = g->type == RTC_GEOMETRY_TYPE_TRIANGLE ? g->indexCount : g->vertexCount;
    for (uint32_t i = 0; i < n; ++i) {
      PrimRef p;
      primBounds(g, i, p.lo, p.hi);
      p.geom = gi;
      p.prim = i;
      s->prims.push_back(p);
    }
  }
  s->nodes.reserve(s->prims.size() + 16);
  s->nodes.push_back(Node{});
  if (!s->prims.empty()) {
    Builder b{s->prims, s->nodes};
    b.build(0, 0, (uint32_t)s->prims.size());
  } else {
    s->nodes[0].count = 0;
    s->nodes[0].left = 0;
    for (int a = 0; a < 3; ++a) {
      s->nodes[0].lo[a] = FLT_MAX;
      s->nodes[0].hi[a] = -FLT_MAX;
    }
  }
  s->committed.store(true, std::memory_order_release);
}

void rtcReleaseScene(RTCScene s) {
  for (auto *g : s->geoms)
    rtcReleaseGeometry(g);
  delete s;
}

RTCGeometry rtcNewGeometry(RTCDevice, RTCGeometryType type) {
  auto *g = new MiniRtcGeometry();
  g->type = type;
  return g;
}

void *rtcSetNewGeometryBuffer(RTCGeometry g, RTCBufferType type, unsigned int, RTCFormat,
                              size_t byteStride, size_t itemCount) {
  void *p = calloc(itemCount * byteStride + 16, 1);
  switch (type) {
  case RTC_BUFFER_TYPE_VERTEX:
    free(g->vertex);
    g->vertex = p;
    g->vertexStride = byteStride;
    g->vertexCount = itemCount;
    break;
  case RTC_BUFFER_TYPE_INDEX:
    free(g->index);
    g->index = p;
    g->indexCount = itemCount;
    break;
  case RTC_BUFFER_TYPE_NORMAL:
    free(g->normal);
    g->normal = p;
    g->normalCount = itemCount;
    break;
  }
  return p;
}
void rtcSetGeometryMask(RTCGeometry, unsigned int) {}
void rtcCommitGeometry(RTCGeometry) {}
void rtcReleaseGeometry(RTCGeometry g) {
  if (--g->refs == 0)
    delete g;
}

size_t miniRtcSceneNodeCount(RTCScene s) { return s->nodes.size(); }

void rtcIntersect1(RTCScene s, RTCRayHit *rh) {
  const float org[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z};
  const float dir[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
  const float tnear = rh->ray.tnear, tfar = rh->ray.tfar;
  float idir[3];
  for (int a = 0; a < 3; ++a)
    idir[a] = 1.f / dir[a]; // +-inf for zero components is fine for slabs

  Best best;
  best.t = tfar;
  best.geom = RTC_INVALID_GEOMETRY_ID;
  best.prim = RTC_INVALID_GEOMETRY_ID;
  bool found = false;

  uint32_t stack[128];
  int sp = 0;
  stack[sp++] = 0;
  const Node *nodes = s->nodes.data();
  while (sp) {
    const Node &n = nodes[stack[--sp]];
    // conservative slab test against [tnear, best.t]
    float t0 = tnear, t1 = best.t;
    bool miss = false;
    for (int a = 0; a < 3; ++a) {
      float ta = (n.lo[a] - org[a]) * idir[a];
      float tb = (n.hi[a] - org[a]) * idir[a];
      if (ta != ta || tb != tb) {
        // 0 * inf: origin on the slab plane with zero direction -> inside if
        // lo <= org <= hi
        if (org[a] < n.lo[a] || org[a] > n.hi[a])
          miss = true;
        continue;
      }
      float tn = std::min(ta, tb), tf = std::max(ta, tb);
      tf *= 1.0000005f; // robust upper bound (Ize 2013)
      tn = tn > 0.f ? tn * 0.9999995f : tn * 1.0000005f;
      t0 = std::max(t0, tn);
      t1 = std::min(t1, tf);
    }
    if (miss || t0 > t1)
      continue;
    if (n.count == 0) {
      stack[sp++] = n.left;
      stack[sp++] = n.left + 1;
    } else {
      for (uint32_t i = n.left; i < n.left + n.count; ++i) {
        const PrimRef &p = s->prims[i];
        const MiniRtcGeometry *g = s->geoms[p.geom];
        Best before = best;
        if (g->type == RTC_GEOMETRY_TYPE_TRIANGLE)
          testTriangle(g, p.geom, p.prim, org, dir, tnear, tfar, best);
        else
          testDisc(g, p.geom, p.prim, org, dir, tnear, tfar, best);
        if (best.prim != before.prim || best.geom != before.geom)
          found = true;
      }
    }
  }
  if (found) {
    rh->ray.tfar = best.t;
    rh->hit.geomID = best.geom;
    rh->hit.primID = best.prim;
    rh->hit.Ng_x = best.ng[0];
    rh->hit.Ng_y = best.ng[1];
    rh->hit.Ng_z = best.ng[2];
    rh->hit.u = best.u;
    rh->hit.v = best.v;
    rh->hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
  }
}
}
