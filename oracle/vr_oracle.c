/* TEST INFRASTRUCTURE (oracle/) -- see vr_oracle.h for scope and parity status.
 *
 * Plain-C restatement of ViennaRay's Monte Carlo flux loop.  Every function
 * cites the reference lines it follows (paths relative to
 * /root/reference/include/viennaray/).  Compile with -ffp-contract=off: all
 * float expressions below are evaluated unfused, left to right, and the CUDA
 * kernels reproduce them operation for operation.
 */
#include "vr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define INVALID_ID 0xffffffffu
#define TNEAR 1e-4f /* rayUtil.hpp:218 fillRayPosition default tnear */

/* ------------------------------------------------------------------------ */
/* small vector helpers; dot = (x*x' + y*y') + z*z'                          */
static inline float dot3(const float *a, const float *b) {
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
static inline void cross3(const float *a, const float *b, float *r) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
/* Normalize(): v *= 1/|v| */
static inline void normalize3(float *v) {
  float inv = 1.0f / sqrtf(dot3(v, v));
  v[0] *= inv;
  v[1] *= inv;
  v[2] *= inv;
}

/* ------------------------------------------------------------------------ */
/* counter-based RNG: Philox4x32-10 (Salmon et al. 2011), key = (seed,
 * stream), counter = (idx_lo, idx_hi, block, 0).  Replaces the reference's
 * per-ray mt19937_64 seeded with tea<3>(idx, seed) (rayTraceKernel.hpp:120). */
void vro_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t *out) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

typedef struct {
  uint32_t k0, k1, c0, c1, blk;
  uint32_t buf[4];
  int pos;
} rng_t;

static inline void rng_init(rng_t *r, uint32_t seed, uint32_t stream, uint64_t idx) {
  r->k0 = seed;
  r->k1 = stream;
  r->c0 = (uint32_t)idx;
  r->c1 = (uint32_t)(idx >> 32);
  r->blk = 0;
  r->pos = 4;
}
static inline uint32_t rng_u32(rng_t *r) {
  if (r->pos == 4) {
    vro_philox4x32(r->k0, r->k1, r->c0, r->c1, r->blk, 0u, r->buf);
    r->blk++;
    r->pos = 0;
  }
  return r->buf[r->pos++];
}
/* uniform in [0,1): top 24 bits */
/* The stream is consumed block-aligned per ray segment: the processing of every traced
 * segment's hit starts at a fresh Philox block and the unused words of the previous one are
 * dropped, so that the state a ray carries between the CUDA kernels is the block counter
 * alone (vr_device.cuh, struct Rng). */
static inline void rng_discard(rng_t *r) { r->pos = 4; }
static inline float rng_f(rng_t *r) { return (float)(rng_u32(r) >> 8) * 5.9604644775390625e-8f; }

/* ------------------------------------------------------------------------ */
/* deterministic elementary functions (polynomials in unfused float ops)     */

/* sin and cos of 2*pi*x, x >= 0 */
static inline void sincos2pi(float x, float *s, float *c) {
  int k = (int)(x * 4.0f + 0.5f);
  float r = x - (float)k * 0.25f;
  float a = r * 6.2831854820251465f;
  float a2 = a * a;
  float sp = -1.9841270114e-4f + a2 * 2.7557318840e-6f;
  sp = 8.3333337680e-3f + a2 * sp;
  sp = -1.6666667163e-1f + a2 * sp;
  float sn = a + (a * a2) * sp;
  float cp = -1.3888889225e-3f + a2 * 2.4801587642e-5f;
  cp = 4.1666667908e-2f + a2 * cp;
  cp = -0.5f + a2 * cp;
  float cs = 1.0f + a2 * cp;
  switch (k & 3) {
  case 0:
    *s = sn;
    *c = cs;
    break;
  case 1:
    *s = cs;
    *c = -sn;
    break;
  case 2:
    *s = -sn;
    *c = -cs;
    break;
  default:
    *s = -cs;
    *c = sn;
    break;
  }
}
/* sin/cos of an angle in radians, angle >= 0 */
static inline void sincos_rad(float a, float *s, float *c) {
  sincos2pi(a * 0.15915493667125702f, s, c);
}

static inline float log2_(float x) { /* x > 0, normal */
  uint32_t u;
  memcpy(&u, &x, 4);
  int e = (int)(u >> 23) - 127;
  u = (u & 0x007fffffu) | 0x3f800000u;
  float m;
  memcpy(&m, &u, 4);
  if (m > 1.41421354f) {
    m = m * 0.5f;
    e += 1;
  }
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float s2 = s * s;
  float p = 0.14285714924f + s2 * 0.11111111194f;
  p = 0.20000000298f + s2 * p;
  p = 0.33333334327f + s2 * p;
  p = 1.0f + s2 * p;
  float ln = (2.0f * s) * p;
  return (float)e + ln * 1.4426950216293335f;
}
static inline float exp2_(float y) {
  if (y < -126.0f)
    return 0.0f;
  float kf = floorf(y + 0.5f);
  float f = (y - kf) * 0.69314718246459961f;
  float p = 1.9841270114e-4f + f * 2.4801587642e-5f;
  p = 1.3888889225e-3f + f * p;
  p = 8.3333337680e-3f + f * p;
  p = 4.1666667908e-2f + f * p;
  p = 1.6666667163e-1f + f * p;
  p = 0.5f + f * p;
  p = 1.0f + f * p;
  p = 1.0f + f * p;
  int k = (int)kf + 127;
  if (k <= 0)
    return 0.0f;
  uint32_t u = (uint32_t)k << 23;
  float sc;
  memcpy(&sc, &u, 4);
  return p * sc;
}
/* x^e for x in [0,1], e > 0; e == 0.5 is an exact square root */
static inline float pow_(float x, float e) {
  if (x <= 0.0f)
    return 0.0f;
  if (e == 0.5f)
    return sqrtf(x);
  float r = exp2_(e * log2_(x));
  return r > 1.0f ? 1.0f : r;
}
/* acos on [0,1], Abramowitz & Stegun 4.4.46 */
static inline float acos_(float x) {
  float p = 0.0066700901f + x * -0.0012624911f;
  p = -0.0170881256f + x * p;
  p = 0.0308918810f + x * p;
  p = -0.0501743046f + x * p;
  p = 0.0889789874f + x * p;
  p = -0.2145988016f + x * p;
  p = 1.5707963050f + x * p;
  return sqrtf(1.0f - x) * p;
}

void vro_math_sincos2pi(const float *x, uint32_t m, float *s, float *c) {
  for (uint32_t i = 0; i < m; ++i)
    sincos2pi(x[i], &s[i], &c[i]);
}
void vro_math_pow(const float *x, float e, uint32_t m, float *out) {
  for (uint32_t i = 0; i < m; ++i)
    out[i] = pow_(x[i], e);
}
void vro_math_acos(const float *x, uint32_t m, float *out) {
  for (uint32_t i = 0; i < m; ++i)
    out[i] = acos_(x[i]);
}

/* ------------------------------------------------------------------------ */
typedef struct {
  float lo[3], hi[3];
  uint32_t left, count; /* count == 0: inner, children left, left+1 */
} node_t;

struct vro_scene {
  int D;
  int geoType; /* 0 disk, 1 triangle */
  uint32_t n;
  float *disk;   /* n x 4 (x,y,z,r), rayGeometryDisk.hpp:363-365 */
  float *normal; /* n x 3 */
  float radius;
  float *verts; /* nVerts x 3 */
  uint32_t nVerts;
  uint32_t *tris; /* n x 3 */
  uint32_t *nbOff, *nbIdx;
  float geoMin[3], geoMax[3];
  /* trace set-up */
  float bbox[2][3];
  int rayDir, firstDir, secondDir, minMax;
  float posNeg;
  int bc[2];
  float bverts[8][3];
  uint32_t btris[8][3];
  /* optional grid source (raySourceGrid.hpp): origins, n x 3 */
  float *grid;
  uint32_t gridN;
  /* material IDs, n (rayGeometry.hpp:19-26,54-57); NULL: all 0 */
  int *matId;
  /* BVH over geometry primitives */
  node_t *nodes;
  uint32_t nNodes;
  uint32_t *primOrder;
};

/* threads of the following vro_trace calls (n <= 0: leave as is); returns the count in
 * effect.  A launcher may export OMP_NUM_THREADS=1 to its workers (torch.distributed.run
 * does): bench.py asks for the cores it reports. */
int vro_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* Geometry::setMaterialIds, rayGeometry.hpp:19-26 (call after the primitives are set) */
int vro_scene_set_material_ids(vro_scene *s, const int *ids) {
  free(s->matId);
  s->matId = NULL;
  if (ids && s->n) {
    s->matId = (int *)malloc(sizeof(int) * s->n);
    memcpy(s->matId, ids, sizeof(int) * s->n);
  }
  return 0;
}

vro_scene *vro_scene_create(int D) {
  vro_scene *s = (vro_scene *)calloc(1, sizeof(vro_scene));
  s->D = D;
  return s;
}
void vro_scene_destroy(vro_scene *s) {
  if (!s)
    return;
  free(s->disk);
  free(s->normal);
  free(s->verts);
  free(s->tris);
  free(s->nbOff);
  free(s->nbIdx);
  free(s->nodes);
  free(s->primOrder);
  free(s->matId);
  free(s->grid);
  free(s);
}

/* origins of a grid source (raySourceGrid.hpp); n == 0 returns to the random source */
int vro_scene_set_source_grid(vro_scene *s, const float *points, uint32_t n) {
  free(s->grid);
  s->grid = NULL;
  s->gridN = 0;
  if (n) {
    s->grid = (float *)malloc(sizeof(float) * 3 * (size_t)n);
    memcpy(s->grid, points, sizeof(float) * 3 * (size_t)n);
    s->gridN = n;
  }
  return 0;
}
uint32_t vro_scene_num_prims(const vro_scene *s) { return s->n; }
void vro_scene_bbox(const vro_scene *s, float *o) {
  for (int k = 0; k < 2; ++k)
    for (int a = 0; a < 3; ++a)
      o[3 * k + a] = s->bbox[k][a];
}
void vro_scene_neighbors(const vro_scene *s, const uint32_t **off, const uint32_t **idx) {
  *off = s->nbOff;
  *idx = s->nbIdx;
}
const float *vro_scene_normals(const vro_scene *s) { return s->normal; }

/* ---- BVH (test-speed only; results do not depend on it) ---------------- */
static void prim_bounds(const vro_scene *s, uint32_t i, float *lo, float *hi) {
  if (s->geoType == 0) {
    const float *c = s->disk + 4 * i, *n = s->normal + 3 * i;
    float nn = dot3(n, n);
    for (int a = 0; a < 3; ++a) {
      float f = nn > 0.f ? 1.f - n[a] * n[a] / nn : 1.f;
      float e = c[3] * sqrtf(f > 0.f ? f : 0.f);
      float pad = 1e-4f * c[3] + 4e-7f * fabsf(c[a]);
      lo[a] = c[a] - e - pad;
      hi[a] = c[a] + e + pad;
    }
  } else {
    for (int a = 0; a < 3; ++a) {
      lo[a] = FLT_MAX;
      hi[a] = -FLT_MAX;
    }
    for (int k = 0; k < 3; ++k) {
      const float *v = s->verts + 3 * s->tris[3 * i + k];
      for (int a = 0; a < 3; ++a) {
        if (v[a] < lo[a])
          lo[a] = v[a];
        if (v[a] > hi[a])
          hi[a] = v[a];
      }
    }
    for (int a = 0; a < 3; ++a) {
      float m = fabsf(lo[a]) > fabsf(hi[a]) ? fabsf(lo[a]) : fabsf(hi[a]);
      float pad = 1e-5f * (hi[a] - lo[a]) + 4e-7f * m + 1e-30f;
      lo[a] -= pad;
      hi[a] += pad;
    }
  }
}

typedef struct {
  vro_scene *s;
  float *plo, *phi; /* per prim bounds */
  uint32_t cap;
} bvh_build_t;

static void bvh_build(bvh_build_t *b, uint32_t node, uint32_t first, uint32_t count) {
  vro_scene *s = b->s;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i = first; i < first + count; ++i) {
    uint32_t p = s->primOrder[i];
    for (int a = 0; a < 3; ++a) {
      float l = b->plo[3 * p + a], h = b->phi[3 * p + a], c = 0.5f * (l + h);
      if (l < lo[a])
        lo[a] = l;
      if (h > hi[a])
        hi[a] = h;
      if (c < clo[a])
        clo[a] = c;
      if (c > chi[a])
        chi[a] = c;
    }
  }
  memcpy(s->nodes[node].lo, lo, 12);
  memcpy(s->nodes[node].hi, hi, 12);
  if (count <= 4) {
    s->nodes[node].left = first;
    s->nodes[node].count = count;
    return;
  }
  int ax = 0;
  if (chi[1] - clo[1] > chi[ax] - clo[ax])
    ax = 1;
  if (chi[2] - clo[2] > chi[ax] - clo[ax])
    ax = 2;
  float mid = 0.5f * (clo[ax] + chi[ax]);
  uint32_t i = first, j = first + count;
  while (i < j) {
    uint32_t p = s->primOrder[i];
    float c = 0.5f * (b->plo[3 * p + ax] + b->phi[3 * p + ax]);
    if (c <= mid)
      ++i;
    else {
      --j;
      s->primOrder[i] = s->primOrder[j];
      s->primOrder[j] = p;
    }
  }
  if (i == first || i == first + count)
    i = first + count / 2;
  uint32_t left = s->nNodes;
  s->nNodes += 2;
  s->nodes[node].left = left;
  s->nodes[node].count = 0;
  bvh_build(b, left, first, i - first);
  bvh_build(b, left + 1, i, first + count - i);
}

static void build_bvh(vro_scene *s) {
  free(s->nodes);
  free(s->primOrder);
  s->nodes = (node_t *)calloc(2 * (size_t)s->n + 2, sizeof(node_t));
  s->primOrder = (uint32_t *)malloc(sizeof(uint32_t) * (s->n + 1));
  bvh_build_t b;
  b.s = s;
  b.plo = (float *)malloc(sizeof(float) * 3 * (s->n + 1));
  b.phi = (float *)malloc(sizeof(float) * 3 * (s->n + 1));
  for (uint32_t i = 0; i < s->n; ++i) {
    s->primOrder[i] = i;
    prim_bounds(s, i, b.plo + 3 * i, b.phi + 3 * i);
  }
  s->nNodes = 1;
  if (s->n)
    bvh_build(&b, 0, 0, s->n);
  else {
    s->nodes[0].count = 0;
    s->nodes[0].left = 0;
    s->nNodes = 0;
  }
  free(b.plo);
  free(b.phi);
}

/* ---- neighbourhood ------------------------------------------------------ */
/* rayPointNeighborhood.hpp:287-298 checkDistance: per-axis |d| <= dist over D
 * axes, then Norm2 over all stored components <= dist^2 (float). */
static int nb_check(const vro_scene *s, const float *p, const float *q, float dist, float dist2) {
  for (int a = 0; a < s->D; ++a)
    if (fabsf(p[a] - q[a]) > dist)
      return 0;
  float d[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
  return dot3(d, d) <= dist2;
}

typedef struct {
  uint64_t key;
  uint32_t idx;
} cellref_t;
static int cellref_cmp(const void *a, const void *b) {
  const cellref_t *x = (const cellref_t *)a, *y = (const cellref_t *)b;
  if (x->key != y->key)
    return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
static int u32_cmp(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : (x > y);
}

/* Neighbour SETS of PointNeighborhood::init (rayPointNeighborhood.hpp:43-107);
 * rows are stored in ascending index order (the reference's row order depends
 * on its recursion and is irrelevant to the flux). */
static void build_neighbors(vro_scene *s, const float *pts /* n x 3 */, float dist) {
  uint32_t n = s->n;
  float dist2 = dist * dist;
  free(s->nbOff);
  free(s->nbIdx);
  s->nbOff = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
  if (n == 0 || !(dist > 0.f)) {
    s->nbIdx = (uint32_t *)malloc(4);
    return;
  }
  float cell = dist * 1.0001f;
  cellref_t *refs = (cellref_t *)malloc(sizeof(cellref_t) * n);
  for (uint32_t i = 0; i < n; ++i) {
    uint64_t key = 0;
    for (int a = 0; a < 3; ++a) {
      int64_t c = (int64_t)floorf((pts[3 * i + a] - s->geoMin[a]) / cell) + 1;
      if (a >= s->D)
        c = 1;
      key |= ((uint64_t)c & 0x1fffff) << (21 * a);
    }
    refs[i].key = key;
    refs[i].idx = i;
  }
  qsort(refs, n, sizeof(cellref_t), cellref_cmp);
  /* two passes: count, then fill */
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t *fill = NULL;
    if (pass == 1) {
      uint32_t acc = 0;
      for (uint32_t i = 0; i <= n; ++i) {
        uint32_t c = s->nbOff[i];
        s->nbOff[i] = acc;
        acc += (i < n) ? c : 0;
      }
      s->nbIdx = (uint32_t *)malloc(sizeof(uint32_t) * (s->nbOff[n] + 1));
      fill = (uint32_t *)calloc(n, sizeof(uint32_t));
    }
    for (uint32_t i = 0; i < n; ++i) {
      const float *p = pts + 3 * i;
      int64_t c[3];
      for (int a = 0; a < 3; ++a) {
        c[a] = (int64_t)floorf((p[a] - s->geoMin[a]) / cell) + 1;
        if (a >= s->D)
          c[a] = 1;
      }
      int zr = s->D == 3 ? 1 : 0;
      for (int dz = -zr; dz <= zr; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            uint64_t key = ((uint64_t)(c[0] + dx) & 0x1fffff) |
                           (((uint64_t)(c[1] + dy) & 0x1fffff) << 21) |
                           (((uint64_t)(c[2] + dz) & 0x1fffff) << 42);
            /* lower bound */
            uint32_t lo = 0, hi = n;
            while (lo < hi) {
              uint32_t mid = (lo + hi) / 2;
              if (refs[mid].key < key)
                lo = mid + 1;
              else
                hi = mid;
            }
            for (uint32_t k = lo; k < n && refs[k].key == key; ++k) {
              uint32_t j = refs[k].idx;
              if (j == i)
                continue;
              if (nb_check(s, p, pts + 3 * j, dist, dist2)) {
                if (pass == 0)
                  s->nbOff[i]++;
                else
                  s->nbIdx[s->nbOff[i] + fill[i]++] = j;
              }
            }
          }
    }
    free(fill);
  }
  for (uint32_t i = 0; i < n; ++i)
    qsort(s->nbIdx + s->nbOff[i], s->nbOff[i + 1] - s->nbOff[i], 4, u32_cmp);
  free(refs);
}

/* GeometryDisk::initGeometry, rayGeometryDisk.hpp:102-193 */
int vro_scene_set_disks(vro_scene *s, const float *points, const float *normals, uint32_t n,
                        float radius) {
  s->geoType = 0;
  s->n = n;
  s->radius = radius;
  free(s->matId); /* material IDs belong to the primitives just replaced */
  s->matId = NULL;
  free(s->disk);
  free(s->normal);
  s->disk = (float *)malloc(sizeof(float) * 4 * (n + 1));
  s->normal = (float *)malloc(sizeof(float) * 3 * (n + 1));
  for (int a = 0; a < 3; ++a) {
    s->geoMin[a] = a < s->D ? FLT_MAX : 0.f;
    s->geoMax[a] = a < s->D ? -FLT_MAX : 0.f;
  }
  for (uint32_t i = 0; i < n; ++i) {
    for (int a = 0; a < 3; ++a) {
      float v = points[3 * i + a], nv = normals[3 * i + a];
      if (a < s->D) {
        if (v < s->geoMin[a])
          s->geoMin[a] = v;
        if (v > s->geoMax[a])
          s->geoMax[a] = v;
      } else { /* :148-151,171-175: z forced to 0 in 2D */
        v = 0.f;
        nv = 0.f;
      }
      s->disk[4 * i + a] = v;
      s->normal[3 * i + a] = nv;
    }
    s->disk[4 * i + 3] = radius;
  }
  /* :191 neighbourhood of the ORIGINAL points, distance 2r */
  build_neighbors(s, points, 2 * radius);
  build_bvh(s);
  return 0;
}

/* TriangleMesh ctor + GeometryTriangle::initGeometry(mesh),
 * rayMesh.hpp:88-121, rayGeometryTriangle.hpp:15-92 */
int vro_scene_set_triangles(vro_scene *s, const float *verts, uint32_t nVerts,
                            const uint32_t *tris, uint32_t n) {
  s->geoType = 1;
  s->n = n;
  s->nVerts = nVerts;
  free(s->matId);
  s->matId = NULL;
  free(s->verts);
  free(s->tris);
  free(s->normal);
  s->verts = (float *)malloc(sizeof(float) * 3 * (nVerts + 1));
  s->tris = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (n + 1));
  s->normal = (float *)malloc(sizeof(float) * 3 * (n + 1));
  memcpy(s->verts, verts, sizeof(float) * 3 * nVerts);
  memcpy(s->tris, tris, sizeof(uint32_t) * 3 * n);
  for (int a = 0; a < 3; ++a) {
    s->geoMin[a] = FLT_MAX;
    s->geoMax[a] = -FLT_MAX;
  }
  for (uint32_t i = 0; i < nVerts; ++i)
    for (int a = 0; a < 3; ++a) {
      float v = verts[3 * i + a];
      if (v < s->geoMin[a])
        s->geoMin[a] = v;
      if (v > s->geoMax[a])
        s->geoMax[a] = v;
    }
  for (uint32_t i = 0; i < n; ++i) {
    const float *p0 = verts + 3 * tris[3 * i], *p1 = verts + 3 * tris[3 * i + 1],
                *p2 = verts + 3 * tris[3 * i + 2];
    float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    float b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    float nrm[3];
    cross3(a, b, nrm);
    normalize3(nrm);
    memcpy(s->normal + 3 * i, nrm, 12);
  }
  free(s->nbOff);
  free(s->nbIdx);
  s->nbOff = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
  s->nbIdx = (uint32_t *)malloc(4);
  build_bvh(s);
  return 0;
}

/* adjustBoundingBox + getTraceSettings + Boundary::initBoundary,
 * rayUtil.hpp:104-202, rayBoundary.hpp:164-245 */
int vro_scene_setup(vro_scene *s, int sourceDir, const int *bc, float off) {
  for (int a = 0; a < 3; ++a) {
    s->bbox[0][a] = s->geoMin[a];
    s->bbox[1][a] = s->geoMax[a];
  }
  if (s->D == 2) {
    s->bbox[0][2] -= off;
    s->bbox[1][2] += off;
    if (sourceDir >= 4)
      return 1;
  }
  static const int settings[6][5] = {{0, 1, 2, 1, -1}, {0, 1, 2, 0, 1}, {1, 0, 2, 1, -1},
                                     {1, 0, 2, 0, 1},  {2, 0, 1, 1, -1}, {2, 0, 1, 0, 1}};
  if (sourceDir < 0 || sourceDir > 5)
    return 1;
  const int *st = settings[sourceDir];
  s->rayDir = st[0];
  s->firstDir = st[1];
  s->secondDir = st[2];
  s->minMax = st[3];
  s->posNeg = (float)st[4];
  if (s->minMax)
    s->bbox[1][s->rayDir] += 2 * off;
  else
    s->bbox[0][s->rayDir] -= 2 * off;
  s->bc[0] = bc[s->firstDir];
  s->bc[1] = (s->D == 3) ? bc[s->secondDir] : VRO_IGNORE; /* unused in 2D */
  for (int v = 0; v < 8; ++v) {
    /* vertex order of rayBoundary.hpp:182-212 */
    int xi = (v == 1 || v == 2 || v == 5 || v == 6), yi = (v == 2 || v == 3 || v == 6 || v == 7),
        zi = v >= 4;
    s->bverts[v][0] = s->bbox[xi][0];
    s->bverts[v][1] = s->bbox[yi][1];
    s->bverts[v][2] = s->bbox[zi][2];
  }
  static const uint32_t planes[3][4][3] = {{{0, 3, 7}, {0, 7, 4}, {6, 2, 1}, {6, 1, 5}},
                                           {{0, 4, 5}, {0, 5, 1}, {6, 7, 3}, {6, 3, 2}},
                                           {{0, 1, 2}, {0, 2, 3}, {6, 5, 4}, {6, 4, 7}}};
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k) {
      s->btris[i][k] = planes[s->firstDir][i][k];
      s->btris[i + 4][k] = planes[s->secondDir][i][k];
    }
  return 0;
}

/* ---- intersection ------------------------------------------------------- */
typedef struct {
  float t;
  uint32_t geom, prim;
  float ng[3];
} hit_t;

static inline int better(float t, uint32_t geom, uint32_t prim, const hit_t *b) {
  if (t < b->t)
    return 1;
  if (t > b->t)
    return 0;
  if (geom != b->geom)
    return geom < b->geom;
  return prim < b->prim;
}

/* oriented disc, rule in oracle/mini_rtc.cpp header */
static inline void test_disk(const vro_scene *s, uint32_t prim, const float *org,
                             const float *dir, hit_t *best) {
  const float *c = s->disk + 4 * prim, *n = s->normal + 3 * prim;
  float den = dot3(dir, n);
  if (den == 0.f)
    return;
  float co[3] = {c[0] - org[0], c[1] - org[1], c[2] - org[2]};
  float t = dot3(co, n) / den;
  if (!(TNEAR <= t && t <= FLT_MAX))
    return;
  float q[3] = {(org[0] + dir[0] * t) - c[0], (org[1] + dir[1] * t) - c[1],
                (org[2] + dir[2] * t) - c[2]};
  if (!(dot3(q, q) < c[3] * c[3]))
    return;
  if (better(t, 1u, prim, best)) {
    best->t = t;
    best->geom = 1u;
    best->prim = prim;
    memcpy(best->ng, n, 12);
  }
}

static inline void test_tri(const float *v0, const float *v1, const float *v2, uint32_t geom,
                            uint32_t prim, const float *org, const float *dir, hit_t *best) {
  float e1[3] = {v0[0] - v1[0], v0[1] - v1[1], v0[2] - v1[2]};
  float e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
  float ng[3], R[3];
  cross3(e2, e1, ng);
  float C[3] = {v0[0] - org[0], v0[1] - org[1], v0[2] - org[2]};
  cross3(C, dir, R);
  float den = dot3(ng, dir);
  if (den == 0.f)
    return;
  float absDen = fabsf(den);
  float U = dot3(R, e2), V = dot3(R, e1), T = dot3(ng, C);
  if (den < 0.f) {
    U = -U;
    V = -V;
    T = -T;
  }
  if (!(U >= 0.f && V >= 0.f && U + V <= absDen))
    return;
  if (!(absDen * TNEAR < T && T <= absDen * FLT_MAX))
    return;
  float t = T / absDen;
  if (better(t, geom, prim, best)) {
    best->t = t;
    best->geom = geom;
    best->prim = prim;
    memcpy(best->ng, ng, 12);
  }
}

/* closest hit over boundary (geomID 0) and geometry (geomID 1), as the scene
 * of rayTraceKernel.hpp:41-45 */
static void intersect(const vro_scene *s, const float *org, const float *dir, hit_t *best) {
  best->t = FLT_MAX;
  best->geom = INVALID_ID;
  best->prim = INVALID_ID;
  best->ng[0] = best->ng[1] = best->ng[2] = 0.f;
  float idir[3] = {1.f / dir[0], 1.f / dir[1], 1.f / dir[2]};
  uint32_t stack[128];
  int sp = 0;
  if (s->nNodes)
    stack[sp++] = 0;
  while (sp) {
    const node_t *n = &s->nodes[stack[--sp]];
    float t0 = TNEAR, t1 = best->t;
    int miss = 0;
    for (int a = 0; a < 3; ++a) {
      float ta = (n->lo[a] - org[a]) * idir[a], tb = (n->hi[a] - org[a]) * idir[a];
      if (ta != ta || tb != tb) {
        if (org[a] < n->lo[a] || org[a] > n->hi[a])
          miss = 1;
        continue;
      }
      float tn = ta < tb ? ta : tb, tf = ta < tb ? tb : ta;
      tf *= 1.0000005f;
      tn = tn > 0.f ? tn * 0.9999995f : tn * 1.0000005f;
      if (tn > t0)
        t0 = tn;
      if (tf < t1)
        t1 = tf;
    }
    if (miss || t0 > t1)
      continue;
    if (n->count == 0) {
      stack[sp++] = n->left;
      stack[sp++] = n->left + 1;
    } else {
      for (uint32_t i = n->left; i < n->left + n->count; ++i) {
        uint32_t p = s->primOrder[i];
        if (s->geoType == 0)
          test_disk(s, p, org, dir, best);
        else
          test_tri(s->verts + 3 * s->tris[3 * p], s->verts + 3 * s->tris[3 * p + 1],
                   s->verts + 3 * s->tris[3 * p + 2], 1u, p, org, dir, best);
      }
    }
  }
  for (uint32_t i = 0; i < 8; ++i)
    test_tri(s->bverts[s->btris[i][0]], s->bverts[s->btris[i][1]], s->bverts[s->btris[i][2]], 0u,
             i, org, dir, best);
}

/* rayTraceKernel.hpp:462-507 checkLocalIntersection */
static inline int check_local_d(const vro_scene *s, const float *org, const float *dir,
                                uint32_t prim, float *distOut) {
  const float *n = s->normal + 3 * prim, *c = s->disk + 4 * prim;
  float prod = dot3(n, dir);
  if (prod > 0.f)
    return 0;
  if (fabsf(prod) < 1e-6f)
    return 0;
  float ddneg = dot3(c, n);
  float tt = (ddneg - dot3(n, org)) / prod;
  if (tt <= 0.f)
    return 0;
  float hp[3];
  for (int i = 0; i < 3; ++i)
    hp[i] = (org[i] + dir[i] * tt) - c[i];
  float distance = sqrtf(dot3(hp, hp));
  *distOut = distance;
  return c[3] > distance;
}
static inline int check_local(const vro_scene *s, const float *org, const float *dir,
                              uint32_t prim) {
  float d;
  return check_local_d(s, org, dir, prim, &d);
}

/* rayUtil.hpp:204-215 fillRayDirection<D>: ray.dir from the particle-facing
 * direction; D == 2 drops z and renormalises the COPY */
static inline void fill_dir(int D, const float *direction, float *rayDir) {
  rayDir[0] = direction[0];
  rayDir[1] = direction[1];
  rayDir[2] = direction[2];
  if (D == 2 && rayDir[2] != 0.f) {
    rayDir[2] = 0.f;
    normalize3(rayDir);
  }
}

/* rayReflection.hpp:13-29 */
static inline void reflect_specular(const float *d, const float *n, float *out) {
  float v[3] = {-d[0], -d[1], -d[2]};
  float f = 2.f * dot3(n, v);
  out[0] = f * n[0] - v[0];
  out[1] = f * n[1] - v[1];
  out[2] = f * n[2] - v[2];
}

/* rayUtil.hpp:266-283 Marsaglia + rayReflection.hpp:32-50 */
static inline void reflect_diffuse(int D, const float *n, rng_t *rng, float *out) {
  float x, y, s2;
  do {
    x = 2.f * rng_f(rng) - 1.f;
    y = 2.f * rng_f(rng) - 1.f;
    s2 = x * x + y * y;
  } while (s2 >= 1.f);
  float tmp = 2.f * sqrtf(1.f - s2);
  out[0] = x * tmp + n[0];
  out[1] = y * tmp + n[1];
  out[2] = D == 3 ? (1.f - 2.f * s2) + n[2] : 0.f;
  normalize3(out);
}

/* rayReflection.hpp:52-120 (current variant) */
static inline void reflect_coned_cosine(int D, const float *d, const float *n, rng_t *rng,
                                        float cone, float *out) {
  if (cone <= 0.f) {
    reflect_specular(d, n, out);
    return;
  }
  if (cone >= 1.57079637050628662f) {
    reflect_diffuse(D, n, rng, out);
    return;
  }
  float w[3];
  reflect_specular(d, n, w);
  normalize3(w);
  float t[3], b[3];
  if (w[2] < -0.999999f) {
    t[0] = 0.f;
    t[1] = -1.f;
    t[2] = 0.f;
    b[0] = -1.f;
    b[1] = 0.f;
    b[2] = 0.f;
  } else {
    float a = 1.f / (1.f + w[2]);
    float bx = -w[0] * w[1] * a;
    float by = 1.f - w[1] * w[1] * a;
    t[0] = 1.f - w[0] * w[0] * a;
    t[1] = bx;
    t[2] = -w[0];
    b[0] = bx;
    b[1] = by;
    b[2] = -w[1];
  }
  float theta, sn, cs;
  for (;;) {
    float u = sqrtf(rng_f(rng));
    float q = 1.f - u;
    float s = sqrtf(q > 0.f ? q : 0.f);
    theta = cone * s;
    float cHalf, sHalf, sTheta, cTheta;
    sincos2pi(0.25f * s, &sHalf, &cHalf); /* cos(pi/2 * s) */
    sincos_rad(theta, &sTheta, &cTheta);
    float rhs = cHalf * sTheta;
    if (rng_f(rng) * theta * u <= rhs) {
      sn = sTheta;
      cs = cTheta;
      break;
    }
  }
  float sp, cp;
  sincos2pi(rng_f(rng), &sp, &cp);
  for (int i = 0; i < 3; ++i)
    out[i] = sn * (cp * t[i] + sp * b[i]) + cs * w[i];
  float dp = dot3(out, n);
  if (dp <= 0.f) {
    float f = 2.f * dp;
    out[0] = out[0] - f * n[0];
    out[1] = out[1] - f * n[1];
    out[2] = out[2] - f * n[2];
  }
  if (D == 2)
    out[2] = 0.f;
  normalize3(out);
}

/* particle functor: returns sticking, writes the reflected direction
 * (rayParticle.hpp:137-146,177-186; coned-cosine recipe of
 * tests/reflection/reflection.cpp:43-46) */
static inline float surface_reflection(const vro_scene *s, const vro_particle *p, const float *d,
                                       const float *n, rng_t *rng, float *out) {
  switch (p->kind) {
  case VRO_DIFFUSE:
    reflect_diffuse(s->D, n, rng, out);
    break;
  case VRO_SPECULAR:
    reflect_specular(d, n, out);
    break;
  default: {
    float c = -dot3(d, n);
    c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
    float inc = acos_(c);
    float m = inc < p->coneMinAngle ? inc : p->coneMinAngle;
    reflect_coned_cosine(s->D, d, n, rng, 1.57079637050628662f - m, out);
  }
  }
  return p->sticking;
}

/* rayBoundary.hpp:29-127.  org/dir: the RTCRay; rayDir3: the particle-facing
 * direction (T-typed in the reference); returns `reflect`. */
int vro_boundary_process_hit(const vro_scene *s, float *org, float *rayDir3, float *dir,
                             const float *ng, uint32_t primID, float t) {
  float impact[3] = {org[0] + dir[0] * t, org[1] + dir[1] * t, org[2] + dir[2] * t};
  if (dot3(dir, ng) > 0.f) { /* :38-44 hit from outside: pass through */
    memcpy(org, impact, 12);
    return 1;
  }
  int cond, axis;
  if (s->D == 2 || primID <= 3) {
    cond = s->bc[0];
    axis = s->firstDir;
  } else {
    cond = s->bc[1];
    axis = s->secondDir;
  }
  if (cond == VRO_REFLECTIVE) { /* :261-271 reflectRay */
    float n[3] = {ng[0], ng[1], ng[2]};
    normalize3(n);
    float nd[3];
    reflect_specular(rayDir3, n, nd);
    memcpy(rayDir3, nd, 12);
    fill_dir(s->D, rayDir3, dir);
    memcpy(org, impact, 12);
    return 1;
  }
  if (cond == VRO_PERIODIC) {
    uint32_t k = primID & 3u;
    impact[axis] = (k <= 1) ? s->bbox[1][axis] : s->bbox[0][axis];
    memcpy(org, impact, 12);
    return 1;
  }
  return 0;
}

/* raySourceRandom.hpp:50-116, rayUtil.hpp:287-321 */
typedef struct {
  float ee, eeGrid;
  int custom;
  float B[3][3];
} source_t;

static void source_init(const vro_scene *s, const vro_particle *p, const vro_config *c,
                        source_t *src) {
  (void)s;
  src->ee = 1.0f / (p->sourcePower + 1.0f);
  src->eeGrid = 2.0f / (p->sourcePower + 1.0f); /* raySourceGrid.hpp:21 */
  src->custom = c->usePrimaryDir;
  if (src->custom) {
    float u[3] = {c->primaryDir[0], c->primaryDir[1], c->primaryDir[2]};
    normalize3(u);
    float h[3];
    if (fabsf(u[0]) > fabsf(u[2])) {
      h[0] = -u[1];
      h[1] = u[0];
      h[2] = 0.f;
    } else {
      h[0] = 0.f;
      h[1] = -u[2];
      h[2] = u[1];
    }
    normalize3(h);
    float w[3];
    cross3(u, h, w);
    memcpy(src->B[0], u, 12);
    memcpy(src->B[1], h, 12);
    memcpy(src->B[2], w, 12);
  }
}

/* SourceGrid::getOriginAndDirection, raySourceGrid.hpp:23-52 */
static void source_sample_grid(const vro_scene *s, const source_t *src, uint64_t idx, rng_t *rng,
                               float *origin, float *direction) {
  const float *g = s->grid + 3 * (size_t)(idx % s->gridN);
  origin[0] = g[0];
  origin[1] = g[1];
  origin[2] = g[2];
  float r1 = rng_f(rng), r2 = rng_f(rng);
  float tt = pow_(r2, src->eeGrid);
  float sinPhi, cosPhi;
  sincos2pi(r1, &sinPhi, &cosPhi);
  float st = sqrtf(1.f - tt);
  direction[s->rayDir] = s->posNeg * sqrtf(tt);
  direction[s->firstDir] = cosPhi * st;
  direction[s->secondDir] = s->D == 2 ? 0.f : sinPhi * st;
  normalize3(direction);
}

static void source_sample(const vro_scene *s, const source_t *src, uint64_t idx, rng_t *rng,
                          float *origin, float *direction) {
  if (s->gridN) {
    source_sample_grid(s, src, idx, rng, origin, direction);
    return;
  }
  origin[0] = origin[1] = origin[2] = 0.f;
  float r1 = rng_f(rng);
  origin[s->rayDir] = s->bbox[s->minMax][s->rayDir];
  origin[s->firstDir] =
      s->bbox[0][s->firstDir] + (s->bbox[1][s->firstDir] - s->bbox[0][s->firstDir]) * r1;
  if (s->D == 2) {
    origin[s->secondDir] = 0.f;
  } else {
    float r2 = rng_f(rng);
    origin[s->secondDir] =
        s->bbox[0][s->secondDir] + (s->bbox[1][s->secondDir] - s->bbox[0][s->secondDir]) * r2;
  }
  for (;;) {
    float q1 = rng_f(rng), q2 = rng_f(rng);
    float sinPhi, cosPhi;
    sincos2pi(q1, &sinPhi, &cosPhi);
    float cosTheta = pow_(q2, src->ee);
    float sinTheta = sqrtf(1.f - cosTheta * cosTheta);
    if (!src->custom) {
      direction[s->rayDir] = s->posNeg * cosTheta;
      direction[s->firstDir] = cosPhi * sinTheta;
      direction[s->secondDir] = sinPhi * sinTheta;
      return;
    }
    float rnd[3] = {cosTheta, cosPhi * sinTheta, sinPhi * sinTheta};
    for (int j = 0; j < 3; ++j)
      direction[j] = (src->B[0][j] * rnd[0] + src->B[1][j] * rnd[1]) + src->B[2][j] * rnd[2];
    if (!((s->posNeg < 0.f && direction[s->rayDir] > 0.f) ||
          (s->posNeg > 0.f && direction[s->rayDir] < 0.f)))
      return;
  }
}

static inline uint64_t to_fixed(float w) { return (uint64_t)(int64_t)(w * 1073741824.0f); }

/* rayTraceKernel.hpp:117-338, one ray */
static void trace_one(const vro_scene *s, const vro_particle *p, const vro_config *c,
                      const source_t *src, uint64_t idx, uint64_t *flux, vro_info *info) {
  rng_t rng;
  rng_init(&rng, c->seed, c->stream, idx);
  const float initialWeight = 1.f; /* raySource.hpp:18 */
  float w = initialWeight;
  float org[3], rayDirection[3], dir[3];
  source_sample(s, src, idx, &rng, org, rayDirection);
  fill_dir(s->D, rayDirection, dir);
  unsigned numReflections = 0, boundaryHits = 0;
  int hitFromBack = 0;
  for (;;) {
    hit_t h;
    intersect(s, org, dir, &h);
    rng_discard(&rng); /* the draws of this segment start at a fresh block */
    info->totalTraces++;
    if (h.geom == INVALID_ID) { /* :172 */
      info->nonGeoHits++;
      break;
    }
    if (p->meanFreePath > 0.f) { /* :179-203 scattering event */
      float scatterProbability = 1.f - exp2_((-h.t / p->meanFreePath) * 1.4426950216293335f);
      float rnd = rng_f(&rng);
      if (rnd < scatterProbability) {
        for (int i = 0; i < 3; ++i)
          org[i] = org[i] + dir[i] * rnd; /* sic: moved by the uniform draw, :188-190 */
        float x, y, s2; /* pickRandomPointOnUnitSphere, rayUtil.hpp:266-283 */
        do {
          x = 2.f * rng_f(&rng) - 1.f;
          y = 2.f * rng_f(&rng) - 1.f;
          s2 = x * x + y * y;
        } while (s2 >= 1.f);
        float tmp = 2.f * sqrtf(1.f - s2);
        rayDirection[0] = x * tmp;
        rayDirection[1] = y * tmp;
        rayDirection[2] = 1.f - 2.f * s2;
        fill_dir(s->D, rayDirection, dir);
        info->particleHits++;
        continue;
      }
    }
    if (h.geom == 0u) { /* :206-214 */
      if (++boundaryHits > c->maxBoundaryHits) {
        info->raysTerminated++;
        break;
      }
      if (!vro_boundary_process_hit(s, org, rayDirection, dir, h.ng, h.prim, h.t))
        break;
      continue;
    }
    float hitPoint[3] = {org[0] + dir[0] * h.t, org[1] + dir[1] * h.t, org[2] + dir[2] * h.t};
    const float *gn = s->normal + 3 * h.prim;
    int backface = dot3(rayDirection, gn) > 0.f; /* :224 */
    if (s->geoType == 0) {
      if (backface) {
        if (hitFromBack) {
          info->raysTerminated++;
          break;
        }
        hitFromBack = 1;
        memcpy(org, hitPoint, 12);
        continue;
      }
    } else if (backface) {
      info->raysTerminated++;
      break;
    }
    info->geoHits++;
    uint64_t wf = to_fixed(w);
    if (s->geoType == 0 && c->useWdist) { /* :258-296 with VIENNARAY_USE_WDIST */
      uint32_t ids[64];
      float dist[64];
      uint32_t nh = 1;
      ids[0] = h.prim;
      {
        const float *dc = s->disk + 4 * h.prim;
        float q[3] = {hitPoint[0] - dc[0], hitPoint[1] - dc[1], hitPoint[2] - dc[2]};
        dist[0] = sqrtf(dot3(q, q)) + 1e-6f;
      }
      for (uint32_t k = s->nbOff[h.prim]; k < s->nbOff[h.prim + 1]; ++k) {
        uint32_t id = s->nbIdx[k];
        float d;
        if (check_local_d(s, org, dir, id, &d) && nh < 64) {
          ids[nh] = id;
          dist[nh++] = d + 1e-6f;
        }
      }
      float invSum = 0.f;
      for (uint32_t k = 0; k < nh; ++k)
        invSum += 1.f / dist[k];
      for (uint32_t k = 0; k < nh; ++k) {
        uint64_t wk = to_fixed(((w / dist[k]) / invSum) * (float)nh);
#pragma omp atomic
        flux[ids[k]] += wk;
      }
    } else if (s->geoType == 0) { /* :255-300 */
#pragma omp atomic
      flux[h.prim] += wf;
      for (uint32_t k = s->nbOff[h.prim]; k < s->nbOff[h.prim + 1]; ++k) {
        uint32_t id = s->nbIdx[k];
        if (check_local(s, org, dir, id)) {
#pragma omp atomic
          flux[id] += wf;
        }
      }
    } else {
#pragma omp atomic
      flux[h.prim] += wf;
    }
    float newDir[3];
    float sticking = surface_reflection(s, p, rayDirection, gn, &rng, newDir); /* :310 */
    if (p->stickingByMaterial) {
      /* a particle whose sticking depends on the materialId it is handed
       * (rayTraceKernel.hpp:310-313, rayParticle.hpp:44-48); IDs outside the table keep
       * the constant */
      const int m = s->matId ? s->matId[h.prim] : 0;
      if (m >= 0 && m < p->numMaterials)
        sticking = p->stickingByMaterial[m];
    }
    w -= w * sticking;                                                        /* :316 */
    if (w <= 0.f)
      break;
    if (++numReflections > c->maxReflections) {
      info->raysTerminated++;
      break;
    }
    /* :435-460 rejectionControl */
    const float lower = 0.1f * initialWeight, renew = 0.3f * initialWeight;
    if (w < lower) {
      float kill = 1.f - w / renew;
      if (rng_f(&rng) < kill)
        break;
      w = renew;
    }
    memcpy(rayDirection, newDir, 12);
    memcpy(org, hitPoint, 12);
    fill_dir(s->D, rayDirection, dir);
  }
  info->boundaryHits += boundaryHits;
  info->reflections += numReflections;
}

int vro_trace(const vro_scene *s, const vro_particle *p, const vro_config *c, uint64_t idxBegin,
              uint64_t idxEnd, uint64_t *flux, vro_info *info) {
  source_t src;
  source_init(s, p, c, &src);
  vro_info total;
  memset(&total, 0, sizeof(total));
#pragma omp parallel
  {
    vro_info local;
    memset(&local, 0, sizeof(local));
#pragma omp for schedule(dynamic, 256)
    for (long long idx = (long long)idxBegin; idx < (long long)idxEnd; ++idx)
      trace_one(s, p, c, &src, (uint64_t)idx, flux, &local);
#pragma omp critical
    {
      total.totalTraces += local.totalTraces;
      total.nonGeoHits += local.nonGeoHits;
      total.geoHits += local.geoHits;
      total.particleHits += local.particleHits;
      total.boundaryHits += local.boundaryHits;
      total.reflections += local.reflections;
      total.raysTerminated += local.raysTerminated;
    }
  }
  info->numRays += idxEnd - idxBegin;
  info->totalTraces += total.totalTraces;
  info->nonGeoHits += total.nonGeoHits;
  info->geoHits += total.geoHits;
  info->particleHits += total.particleHits;
  info->boundaryHits += total.boundaryHits;
  info->reflections += total.reflections;
  info->raysTerminated += total.raysTerminated;
  return 0;
}

int vro_source_rays(const vro_scene *s, const vro_particle *p, const vro_config *c,
                    uint64_t idxBegin, uint32_t m, float *rays) {
  source_t src;
  source_init(s, p, c, &src);
  for (uint32_t i = 0; i < m; ++i) {
    rng_t rng;
    rng_init(&rng, c->seed, c->stream, idxBegin + i);
    float d[3];
    source_sample(s, &src, idxBegin + i, &rng, rays + 6 * i, d);
    fill_dir(s->D, d, rays + 6 * i + 3);
  }
  return 0;
}

int vro_intersect(const vro_scene *s, const float *rays, uint32_t m, uint32_t *geom,
                  uint32_t *prim, float *t, float *ng) {
#pragma omp parallel for schedule(dynamic, 256)
  for (long long i = 0; i < (long long)m; ++i) {
    hit_t h;
    intersect(s, rays + 6 * i, rays + 6 * i + 3, &h);
    geom[i] = h.geom;
    prim[i] = h.prim;
    t[i] = h.t;
    if (ng)
      memcpy(ng + 3 * i, h.ng, 12);
  }
  return 0;
}

int vro_neighbor_hits(const vro_scene *s, const float *rays, const uint32_t *prim, uint32_t m,
                      uint32_t cap, uint32_t *count, uint32_t *out) {
  for (uint32_t i = 0; i < m; ++i) {
    uint32_t c = 0;
    if (prim[i] != INVALID_ID && s->geoType == 0)
      for (uint32_t k = s->nbOff[prim[i]]; k < s->nbOff[prim[i] + 1]; ++k)
        if (check_local(s, rays + 6 * i, rays + 6 * i + 3, s->nbIdx[k])) {
          if (c < cap)
            out[(size_t)i * cap + c] = s->nbIdx[k];
          ++c;
        }
    count[i] = c;
  }
  return 0;
}

/* rayTraceDisk.hpp:120-139 / rayTraceTriangle.hpp:104-123, SOURCE mode */
void vro_normalize_flux_source(const vro_scene *s, const float *areas, uint64_t numRays,
                               float *flux) {
  float sourceArea = s->bbox[1][s->firstDir] - s->bbox[0][s->firstDir];
  if (s->D == 3)
    sourceArea = sourceArea * (s->bbox[1][s->secondDir] - s->bbox[0][s->secondDir]);
  float normFactor = sourceArea / (float)numRays;
  for (uint32_t i = 0; i < s->n; ++i)
    flux[i] *= normFactor / areas[i];
}

/* rayTraceDisk.hpp:146-193 with numNeighbors == 1 */
void vro_smooth_flux(const vro_scene *s, float *flux) {
  float *old = (float *)malloc(sizeof(float) * (s->n + 1));
  memcpy(old, flux, sizeof(float) * s->n);
  for (uint32_t i = 0; i < s->n; ++i) {
    float vv = old[i], sum = 1.f;
    const float *n = s->normal + 3 * i;
    for (uint32_t k = s->nbOff[i]; k < s->nbOff[i + 1]; ++k) {
      uint32_t j = s->nbIdx[k];
      float wgt = dot3(n, s->normal + 3 * j);
      if (wgt > 0.f) {
        vv += old[j] * wgt;
        sum += wgt;
      }
    }
    flux[i] = vv / sum;
  }
  free(old);
}

/* normalizeFlux(MAX): rayTraceDisk.hpp:110-118 (disks: the factor is formed in double,
 * diskRadius_ * diskRadius_ * M_PI) and rayTraceTriangle.hpp:99-107 (triangles: float) */
void vro_normalize_flux_max(const vro_scene *s, const float *areas, float *flux) {
  float maxv = flux[0];
  for (uint32_t i = 1; i < s->n; ++i)
    if (flux[i] > maxv)
      maxv = flux[i];
  if (s->geoType == 0) {
    const double total = (double)(s->radius * s->radius) * 3.14159265358979323846;
    for (uint32_t i = 0; i < s->n; ++i)
      flux[i] = (float)((double)flux[i] * ((total / (double)areas[i]) / (double)maxv));
  } else {
    for (uint32_t i = 0; i < s->n; ++i)
      flux[i] = flux[i] / (maxv * areas[i]);
  }
}

/* rayTraceDisk.hpp:146-193 with numNeighbors = k > 1: a new neighbourhood of distance
 * k * 2 * radius over the points (init<3>: all three axes), rows in ascending index order */
void vro_smooth_flux_k(const vro_scene *s, int k, float *flux) {
  if (k <= 1) {
    vro_smooth_flux(s, flux);
    return;
  }
  vro_scene t;
  memset(&t, 0, sizeof(t));
  t.D = 3;
  t.n = s->n;
  memcpy(t.geoMin, s->geoMin, sizeof(t.geoMin));
  memcpy(t.geoMax, s->geoMax, sizeof(t.geoMax));
  float *pts = (float *)malloc(sizeof(float) * 3 * (s->n + 1));
  for (uint32_t i = 0; i < s->n; ++i)
    for (int a = 0; a < 3; ++a)
      pts[3 * i + a] = s->disk[4 * i + a];
  build_neighbors(&t, pts, (float)k * 2 * s->radius);
  t.normal = s->normal;
  vro_smooth_flux(&t, flux);
  free(t.nbOff);
  free(t.nbIdx);
  free(pts);
}

/* m reflections of (rayDir, normal) with ray streams idx .. idx+m-1 */
void vro_reflect(int kind, int D, const float *rayDir, const float *normal, float coneMinAngle,
                 uint32_t seed, uint64_t idx, uint32_t m, float *out) {
  vro_scene s;
  memset(&s, 0, sizeof(s));
  s.D = D;
  vro_particle p = {kind, 1.f, 1.f, coneMinAngle};
  for (uint32_t i = 0; i < m; ++i) {
    rng_t rng;
    rng_init(&rng, seed, 0u, idx + i);
    surface_reflection(&s, &p, rayDir, normal, &rng, out + 3 * i);
  }
}
