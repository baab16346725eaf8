// Device-side construction of the acceleration structure; replaces Embree's
// rtcJoinCommitScene (rayTraceKernel.hpp:91).
//
//   1. 63-bit Morton code of every primitive's box centre
//   2. radix sort (CUB) of (code, primitive)
//   3. Karras-2012 binary radix tree over the sorted codes, one thread per
//      internal node; index bits break ties between equal codes
//   4. bottom-up refit of node boxes with one atomic flag per node
//   5. emission of 32-byte nodes holding both child boxes on a 16-bit grid; subtrees of at
//      most VR_LEAF_MAX primitives become leaves (contiguous sorted ranges).  Only about
//      one radix-tree node in seven survives the leaf collapse, so the survivors are
//      renumbered densely by a prefix sum over the radix-tree order: a subtree's nodes stay
//      contiguous (every internal node of a subtree lies inside its key range), the node
//      array shrinks from 32 B x (n-1) to 32 B x numNodes and a 128-byte cache line holds
//      four nodes that are visited together instead of one
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>

#include "vr_internal.h"

// entries of the top-of-tree table the traverse kernel keeps in shared memory (0: none);
// VR_TOP_NODES in the environment overrides it
#ifndef VR_TOP_DEFAULT
#define VR_TOP_DEFAULT 0u
#endif

namespace vr {

namespace {

__device__ __forceinline__ unsigned long long expandBits21(unsigned long long v) {
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

__global__ void mortonKernel(const float4 *lo, const float4 *hi, uint32_t n, float3 sLo,
                             float3 sInv, unsigned long long *keys, uint32_t *vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float4 l = lo[i], h = hi[i];
  float cx = (0.5f * (l.x + h.x) - sLo.x) * sInv.x;
  float cy = (0.5f * (l.y + h.y) - sLo.y) * sInv.y;
  float cz = (0.5f * (l.z + h.z) - sLo.z) * sInv.z;
  const float S = 2097151.f;
  unsigned long long x = (unsigned long long)fminf(fmaxf(cx * S, 0.f), S);
  unsigned long long y = (unsigned long long)fminf(fmaxf(cy * S, 0.f), S);
  unsigned long long z = (unsigned long long)fminf(fmaxf(cz * S, 0.f), S);
  keys[i] = (expandBits21(x) << 2) | (expandBits21(y) << 1) | expandBits21(z);
  vals[i] = i;
}

// length of the common prefix of (key_i, i) and (key_j, j); -1 out of range
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
  if (j < 0 || j >= n)
    return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a != b)
    return __clzll((long long)(a ^ b));
  return 64 + __clz(i ^ j);
}

// Karras 2012: internal node i covers [first,last]; children are split, split+1
__global__ void radixTreeKernel(const unsigned long long *keys, int n, int2 *range, int *split,
                                int *parentOfInternal, int *parentOfLeaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dMin = delta(keys, n, i, i - d);
  int lMax = 2;
  while (delta(keys, n, i, i + lMax * d) > dMin)
    lMax <<= 1;
  int l = 0;
  for (int t = lMax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dMin)
      l += t;
  int j = i + l * d;
  int dNode = delta(keys, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dNode)
      s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int first = min(i, j), last = max(i, j);
  range[i] = make_int2(first, last);
  split[i] = gamma;
  if (first == gamma)
    parentOfLeaf[gamma] = i;
  else
    parentOfInternal[gamma] = i;
  if (last == gamma + 1)
    parentOfLeaf[gamma + 1] = i;
  else
    parentOfInternal[gamma + 1] = i;
  if (i == 0)
    parentOfInternal[0] = -1;
}

__global__ void refitKernel(const float4 *lo, const float4 *hi, const uint32_t *sorted, int n,
                            const int2 *range, const int *split, const int *parentOfInternal,
                            const int *parentOfLeaf, float4 *nodeLo, float4 *nodeHi, int *flags) {
  int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n)
    return;
  int node = parentOfLeaf[leaf];
  while (node >= 0) {
    if (atomicAdd(&flags[node], 1) == 0)
      return;  // the sibling subtree finishes this node
    __threadfence();
    int g = split[node];
    int2 r = range[node];
    float4 l0, h0, l1, h1;
    if (r.x == g) {
      uint32_t p = sorted[g];
      l0 = lo[p];
      h0 = hi[p];
    } else {
      l0 = __ldcg(&nodeLo[g]);  // written by another SM: read through L2
      h0 = __ldcg(&nodeHi[g]);
    }
    if (r.y == g + 1) {
      uint32_t p = sorted[g + 1];
      l1 = lo[p];
      h1 = hi[p];
    } else {
      l1 = __ldcg(&nodeLo[g + 1]);
      h1 = __ldcg(&nodeHi[g + 1]);
    }
    nodeLo[node] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
    nodeHi[node] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
    __threadfence();
    node = parentOfInternal[node];
  }
}

__device__ __forceinline__ uint32_t childRef(int first, int last, int internalIdx,
                                             uint32_t leafMax) {
  uint32_t cnt = (uint32_t)(last - first + 1);
  if (cnt <= leafMax)
    return VR_LEAF_FLAG | ((uint32_t)first << 4) | cnt;
  return (uint32_t)internalIdx;
}

// depth of the radix tree: every primitive climbs to the root and counts the levels above
// it (the emitted tree is never deeper: collapsing subtrees into leaves only removes levels)
__global__ void depthKernel(const int *parentOfInternal, const int *parentOfLeaf, int n,
                            unsigned int *maxDepth) {
  int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned d = 0;
  if (leaf < n)
    for (int node = parentOfLeaf[leaf]; node >= 0; node = parentOfInternal[node])
      ++d;
  d = __reduce_max_sync(0xffffffffu, d);
  if ((threadIdx.x & 31) == 0)
    atomicMax(maxDepth, d);
}

// radix-tree node i survives the leaf collapse (is referenced by the emitted tree)
__global__ void liveKernel(const int2 *range, int n, uint32_t leafMax, uint32_t *live) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;
  const int2 r = range[i];
  live[i] = ((uint32_t)(r.y - r.x + 1) > leafMax || i == 0) ? 1u : 0u;
}
// reference of the radix tree -> reference of the dense node array
__device__ __forceinline__ uint32_t denseRef(uint32_t ref, const uint32_t *dense) {
  return (ref & VR_LEAF_FLAG) ? ref : dense[ref];
}

// box -> six 16-bit grid coordinates, rounded outwards by one extra cell so the
// decode (a fused multiply-add in the traversal) can never shrink the box
__device__ __forceinline__ uint4 quantizeChild(float4 l, float4 h, float3 qLo, float3 qInv,
                                               uint32_t ref) {
  auto qdn = [](float v, float o, float inv) -> uint32_t {
    float q = floorf((v - o) * inv) - 1.f;
    return (uint32_t)fminf(fmaxf(q, 0.f), 65535.f);
  };
  auto qup = [](float v, float o, float inv) -> uint32_t {
    float q = ceilf((v - o) * inv) + 1.f;
    return (uint32_t)fminf(fmaxf(q, 0.f), 65535.f);
  };
  uint32_t lx = qdn(l.x, qLo.x, qInv.x), ly = qdn(l.y, qLo.y, qInv.y), lz = qdn(l.z, qLo.z, qInv.z);
  uint32_t hx = qup(h.x, qLo.x, qInv.x), hy = qup(h.y, qLo.y, qInv.y), hz = qup(h.z, qLo.z, qInv.z);
  return make_uint4(lx | (hx << 16), ly | (hy << 16), lz | (hz << 16), ref);
}

__device__ __forceinline__ void emitNode(int i, const float4 *lo, const float4 *hi,
                                         const uint32_t *sorted, int n, const int2 *range,
                                         const int *split, const float4 *nodeLo,
                                         const float4 *nodeHi, const uint32_t *dense,
                                         Node2 *nodes, float3 qLo,
                                         float3 qInv, double *sah, double &aIn, double &aLeaf,
                                         unsigned &cNodes, unsigned &cLeaves,
                                         unsigned &cMaxLeaf, uint32_t leafMax) {
  if (i >= n - 1)
    return;
  int2 r = range[i];
  if ((uint32_t)(r.y - r.x + 1) <= leafMax && i != 0)
    return;  // inside a collapsed leaf (never referenced)
  int g = split[i];
  float4 l0, h0, l1, h1;
  uint32_t ref0, ref1;
  if (r.x == g) {
    uint32_t p = sorted[g];
    l0 = lo[p];
    h0 = hi[p];
    ref0 = VR_LEAF_FLAG | ((uint32_t)g << 4) | 1u;
  } else {
    l0 = nodeLo[g];
    h0 = nodeHi[g];
    int2 rc = range[g];
    ref0 = childRef(rc.x, rc.y, g, leafMax);
  }
  if (r.y == g + 1) {
    uint32_t p = sorted[g + 1];
    l1 = lo[p];
    h1 = hi[p];
    ref1 = VR_LEAF_FLAG | ((uint32_t)(g + 1) << 4) | 1u;
  } else {
    l1 = nodeLo[g + 1];
    h1 = nodeHi[g + 1];
    int2 rc = range[g + 1];
    ref1 = childRef(rc.x, rc.y, g + 1, leafMax);
  }
  Node2 nd;
  nd.c0 = quantizeChild(l0, h0, qLo, qInv, denseRef(ref0, dense));
  nd.c1 = quantizeChild(l1, h1, qLo, qInv, denseRef(ref1, dense));
  nodes[dense[i]] = nd;
  cNodes = 1u;
  cLeaves = ((ref0 & VR_LEAF_FLAG) ? 1u : 0u) + ((ref1 & VR_LEAF_FLAG) ? 1u : 0u);
  if (ref0 & VR_LEAF_FLAG)
    cMaxLeaf = ref0 & 15u;
  if (ref1 & VR_LEAF_FLAG)
    cMaxLeaf = max(cMaxLeaf, ref1 & 15u);
  // surface-area-heuristic terms: area of this inner node, areas of its leaf children
  // times their primitive counts (the host divides by the root area)
  auto area = [](float4 l, float4 h) {
    const float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
    return 2.f * ((dx * dy + dy * dz) + dz * dx);
  };
  float leafArea = 0.f;
  if (ref0 & VR_LEAF_FLAG)
    leafArea += area(l0, h0) * (float)(ref0 & 15u);
  if (ref1 & VR_LEAF_FLAG)
    leafArea += area(l1, h1) * (float)(ref1 & 15u);
  aIn = (double)area(nodeLo[i], nodeHi[i]);
  aLeaf = (double)leafArea;
  if (i == 0)
    sah[2] = aIn;
}

__global__ void emitKernel(const float4 *lo, const float4 *hi, const uint32_t *sorted, int n,
                           const int2 *range, const int *split, const float4 *nodeLo,
                           const float4 *nodeHi, const uint32_t *dense, Node2 *nodes,
                           unsigned int *stats, float3 qLo, float3 qInv, double *sah,
                           uint32_t leafMax) {
  double aIn = 0., aLeaf = 0.;
  unsigned cNodes = 0u, cLeaves = 0u, cMaxLeaf = 0u;
  emitNode(blockIdx.x * blockDim.x + threadIdx.x, lo, hi, sorted, n, range, split, nodeLo, nodeHi,
           dense, nodes, qLo, qInv, sah, aIn, aLeaf, cNodes, cLeaves, cMaxLeaf, leafMax);
  // statistics and SAH terms summed per warp first: a million atomics on one address
  // serialise
  cNodes = __reduce_add_sync(0xffffffffu, cNodes);
  cLeaves = __reduce_add_sync(0xffffffffu, cLeaves);
  cMaxLeaf = __reduce_max_sync(0xffffffffu, cMaxLeaf);
  for (int o = 16; o > 0; o >>= 1) {
    aIn += __shfl_down_sync(0xffffffffu, aIn, o);
    aLeaf += __shfl_down_sync(0xffffffffu, aLeaf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cNodes) {
      atomicAdd(&stats[0], cNodes);
      atomicAdd(&stats[1], cLeaves);
      atomicMax(&stats[2], cMaxLeaf);
    }
    if (aIn > 0.)
      atomicAdd(&sah[0], aIn);
    if (aLeaf > 0.)
      atomicAdd(&sah[1], aLeaf);
  }
}

// ---- optional 4-wide nodes (VR_BVH_WIDE=1): every live binary node collapses its two
// children's children into one 64-byte node of up to four quantised boxes.  Entry k is
// four words like a Node2 child; unused entries carry the reference VR_DONE.
struct ChildBox {
  float4 lo, hi;
  uint32_t ref;
};

__device__ __forceinline__ ChildBox childOf(int i, int side, const float4 *lo, const float4 *hi,
                                            const uint32_t *sorted, const int2 *range,
                                            const int *split, const float4 *nodeLo,
                                            const float4 *nodeHi, uint32_t leafMax) {
  const int2 r = range[i];
  const int c = split[i] + side;
  const bool single = side == 0 ? (r.x == c) : (r.y == c);
  ChildBox b;
  if (single) {
    const uint32_t p = sorted[c];
    b.lo = lo[p];
    b.hi = hi[p];
    b.ref = VR_LEAF_FLAG | ((uint32_t)c << 4) | 1u;
  } else {
    b.lo = nodeLo[c];
    b.hi = nodeHi[c];
    const int2 rc = range[c];
    b.ref = childRef(rc.x, rc.y, c, leafMax);
  }
  return b;
}

__global__ void emit4Kernel(const float4 *lo, const float4 *hi, const uint32_t *sorted, int n,
                            const int2 *range, const int *split, const float4 *nodeLo,
                            const float4 *nodeHi, const uint32_t *dense, uint4 *nodes4,
                            float3 qLo, float3 qInv, uint32_t leafMax) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;
  const int2 r = range[i];
  if ((uint32_t)(r.y - r.x + 1) <= leafMax && i != 0)
    return;
  uint4 e[4];
  int k = 0;
  for (int side = 0; side < 2; ++side) {
    const ChildBox b = childOf(i, side, lo, hi, sorted, range, split, nodeLo, nodeHi, leafMax);
    if (b.ref & VR_LEAF_FLAG) {
      e[k++] = quantizeChild(b.lo, b.hi, qLo, qInv, b.ref);  // a leaf reference
    } else {
      for (int s2 = 0; s2 < 2; ++s2) {
        const ChildBox g =
            childOf((int)b.ref, s2, lo, hi, sorted, range, split, nodeLo, nodeHi, leafMax);
        e[k++] = quantizeChild(g.lo, g.hi, qLo, qInv, denseRef(g.ref, dense));
      }
    }
  }
  for (; k < 4; ++k)
    e[k] = make_uint4(0u, 0u, 0u, VR_DONE);
  for (k = 0; k < 4; ++k)
    nodes4[4 * (size_t)dense[i] + k] = e[k];
}

// Top of the tree in breadth-first order for the traverse kernel's shared-memory copy.
// One block walks the levels; a child that still fits the table gets the reference
// VR_TOP_BASE + entry, every other reference is kept as it is.
__global__ void topTableKernel(const Node2 *nodes, uint32_t maxTop, uint4 *top,
                               uint32_t *countOut) {
  __shared__ uint32_t ids[VR_TOP_MAX];
  __shared__ uint32_t count;
  if (threadIdx.x == 0) {
    ids[0] = 0u;
    count = 1u;
  }
  __syncthreads();
  uint32_t b = 0, e = 1;
  while (b < e) {
    for (uint32_t k = b + threadIdx.x; k < e; k += blockDim.x) {
      Node2 nd = nodes[ids[k]];
      if (nd.c0.w < VR_DONE) {
        const uint32_t pos = atomicAdd(&count, 1u);
        if (pos < maxTop) {
          ids[pos] = nd.c0.w;
          nd.c0.w = VR_TOP_BASE + pos;
        }
      }
      if (nd.c1.w < VR_DONE) {
        const uint32_t pos = atomicAdd(&count, 1u);
        if (pos < maxTop) {
          ids[pos] = nd.c1.w;
          nd.c1.w = VR_TOP_BASE + pos;
        }
      }
      top[2 * k] = nd.c0;
      top[2 * k + 1] = nd.c1;
    }
    __syncthreads();
    b = e;
    e = min(count, maxTop);
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *countOut = e;
}

}  // namespace

#define VR_CK(x)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      cleanup();                                                                                   \
      return e_;                                                                                   \
    }                                                                                              \
  } while (0)

// one LBVH over Morton codes computed with the per-axis scales sInv
static cudaError_t buildOne(const float4 *primLo, const float4 *primHi, uint32_t n,
                            const float sceneLo[3], const float sceneHi[3], float3 sInv,
                            uint32_t leafMax, cudaStream_t stream, Bvh *out) {
  freeBvh(out, stream);
  bool wide = getenv("VR_BVH_WIDE") && atoi(getenv("VR_BVH_WIDE")) > 0;
  unsigned long long *keys = nullptr, *keysSorted = nullptr;
  uint32_t *vals = nullptr;
  void *tmp = nullptr;
  int2 *range = nullptr;
  int *split = nullptr, *parI = nullptr, *parL = nullptr, *flags = nullptr;
  float4 *nodeLo = nullptr, *nodeHi = nullptr;
  unsigned int *stats = nullptr;
  uint32_t *live = nullptr, *dense = nullptr;
  void *scanTmp = nullptr;
  double *sah = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    cudaFreeAsync(live, stream);
    cudaFreeAsync(dense, stream);
    cudaFreeAsync(scanTmp, stream);
    cudaFreeAsync(keys, stream);
    cudaFreeAsync(keysSorted, stream);
    cudaFreeAsync(vals, stream);
    cudaFreeAsync(tmp, stream);
    cudaFreeAsync(range, stream);
    cudaFreeAsync(split, stream);
    cudaFreeAsync(parI, stream);
    cudaFreeAsync(parL, stream);
    cudaFreeAsync(flags, stream);
    cudaFreeAsync(nodeLo, stream);
    cudaFreeAsync(nodeHi, stream);
    cudaFreeAsync(stats, stream);
    cudaFreeAsync(sah, stream);
    if (e0)
      cudaEventDestroy(e0);
    if (e1)
      cudaEventDestroy(e1);
  };
  if (n == 0) {
    out->rootRef = VR_INVALID_ID;
    return cudaSuccess;
  }
  VR_CK(cudaEventCreate(&e0));
  VR_CK(cudaEventCreate(&e1));
  VR_CK(cudaEventRecord(e0, stream));
  VR_CK(cudaMallocAsync(&keys, sizeof(unsigned long long) * n, stream));
  VR_CK(cudaMallocAsync(&keysSorted, sizeof(unsigned long long) * n, stream));
  VR_CK(cudaMallocAsync(&vals, sizeof(uint32_t) * n, stream));
  VR_CK(cudaMallocAsync(&out->sortedToOrig, sizeof(uint32_t) * n, stream));
  // quantisation grid of the node boxes: 65536 cells over the (padded) scene box
  float3 qLo, qInv;
  {
    float ql[3], qi[3];
    for (int a = 0; a < 3; ++a) {
      float ext = sceneHi[a] - sceneLo[a];
      float margin = 1e-3f * ext + 1e-3f;
      ql[a] = sceneLo[a] - margin;
      float sc = (ext + 2.f * margin) / 65535.f;
      out->qLo[a] = ql[a];
      out->qScale[a] = sc;
      qi[a] = 1.f / sc;
    }
    qLo = make_float3(ql[0], ql[1], ql[2]);
    qInv = make_float3(qi[0], qi[1], qi[2]);
  }
  float3 sLo = make_float3(sceneLo[0], sceneLo[1], sceneLo[2]);
  const int B = 256;
  mortonKernel<<<(n + B - 1) / B, B, 0, stream>>>(primLo, primHi, n, sLo, sInv, keys, vals);
  VR_CK(cudaGetLastError());
  size_t tmpBytes = 0;
  VR_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, vals,
                                        out->sortedToOrig, (int)n, 0, 63, stream));
  VR_CK(cudaMallocAsync(&tmp, tmpBytes ? tmpBytes : 16, stream));
  VR_CK(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, vals, out->sortedToOrig,
                                        (int)n, 0, 63, stream));
  if (n <= leafMax) {
    out->rootRef = VR_LEAF_FLAG | (0u << 4) | n;
    out->numNodes = 0;
    out->numLeaves = 1;
    out->maxLeaf = n;
    VR_CK(cudaMallocAsync(&out->nodes, sizeof(Node2), stream));
  } else {
    VR_CK(cudaMallocAsync(&range, sizeof(int2) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&split, sizeof(int) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&parI, sizeof(int) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&parL, sizeof(int) * n, stream));
    VR_CK(cudaMallocAsync(&flags, sizeof(int) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&nodeLo, sizeof(float4) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&nodeHi, sizeof(float4) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&stats, sizeof(unsigned int) * 8, stream));
    VR_CK(cudaMallocAsync(&live, sizeof(uint32_t) * (n - 1), stream));
    VR_CK(cudaMallocAsync(&dense, sizeof(uint32_t) * (n - 1), stream));
    VR_CK(cudaMemsetAsync(flags, 0, sizeof(int) * (n - 1), stream));
    VR_CK(cudaMemsetAsync(stats, 0, sizeof(unsigned int) * 8, stream));
    VR_CK(cudaMallocAsync(&sah, sizeof(double) * 4, stream));
    VR_CK(cudaMemsetAsync(sah, 0, sizeof(double) * 4, stream));
    radixTreeKernel<<<(n - 1 + B - 1) / B, B, 0, stream>>>(keysSorted, (int)n, range, split, parI,
                                                           parL);
    VR_CK(cudaGetLastError());
    refitKernel<<<(n + B - 1) / B, B, 0, stream>>>(primLo, primHi, out->sortedToOrig, (int)n,
                                                   range, split, parI, parL, nodeLo, nodeHi, flags);
    VR_CK(cudaGetLastError());
    depthKernel<<<(n + B - 1) / B, B, 0, stream>>>(parI, parL, (int)n, stats + 4);
    VR_CK(cudaGetLastError());
    // dense numbering of the nodes that survive the leaf collapse (radix-tree order kept)
    liveKernel<<<(n - 1 + B - 1) / B, B, 0, stream>>>(range, (int)n, leafMax, live);
    VR_CK(cudaGetLastError());
    size_t scanBytes = 0;
    VR_CK(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, live, dense, (int)(n - 1), stream));
    VR_CK(cudaMallocAsync(&scanTmp, scanBytes ? scanBytes : 16, stream));
    VR_CK(cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, live, dense, (int)(n - 1), stream));
    uint32_t lastDense = 0, lastLive = 0, depth = 0;
    VR_CK(cudaMemcpyAsync(&depth, stats + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VR_CK(cudaMemcpyAsync(&lastDense, dense + (n - 2), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          stream));
    VR_CK(cudaMemcpyAsync(&lastLive, live + (n - 2), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          stream));
    VR_CK(cudaStreamSynchronize(stream));
    const size_t numLive = (size_t)lastDense + lastLive;
    out->maxDepth = depth;
    // The traversal stacks are unchecked (VR_STACK entries).  Binary descent: one push per
    // level at most.  4-wide descent: three per node, a node spanning two binary levels.
    if (depth > VR_STACK) {
      cleanup();
      return cudaErrorInvalidConfiguration;  // cannot happen for 63 + 32 key bits
    }
    if (wide && 3u * ((depth + 1u) / 2u) > VR_STACK)
      wide = false;  // too deep for the wide traversal's stack: binary nodes only
    VR_CK(cudaMallocAsync(&out->nodes, sizeof(Node2) * std::max<size_t>(numLive, 1), stream));
    emitKernel<<<(n - 1 + B - 1) / B, B, 0, stream>>>(primLo, primHi, out->sortedToOrig, (int)n,
                                                      range, split, nodeLo, nodeHi, dense,
                                                      out->nodes, stats, qLo, qInv, sah, leafMax);
    VR_CK(cudaGetLastError());
    if (wide) {
      VR_CK(cudaMallocAsync(&out->nodes4, sizeof(uint4) * 4 * std::max<size_t>(numLive, 1), stream));
      emit4Kernel<<<(n - 1 + B - 1) / B, B, 0, stream>>>(primLo, primHi, out->sortedToOrig, (int)n,
                                                         range, split, nodeLo, nodeHi, dense,
                                                         out->nodes4, qLo, qInv, leafMax);
      VR_CK(cudaGetLastError());
    }
    uint32_t topWanted = VR_TOP_DEFAULT;
    if (const char *tn = getenv("VR_TOP_NODES"))
      topWanted = (uint32_t)std::min<long>(std::max<long>(atol(tn), 0), (long)VR_TOP_MAX);
    if (topWanted && !wide) {
      VR_CK(cudaMallocAsync(&out->top, sizeof(uint4) * 2 * topWanted, stream));
      topTableKernel<<<1, 256, 0, stream>>>(out->nodes, topWanted, out->top, stats + 3);
      VR_CK(cudaGetLastError());
    }
    unsigned int hs[8];
    double hsah[4];
    VR_CK(cudaMemcpyAsync(hs, stats, sizeof(hs), cudaMemcpyDeviceToHost, stream));
    VR_CK(cudaMemcpyAsync(hsah, sah, sizeof(hsah), cudaMemcpyDeviceToHost, stream));
    VR_CK(cudaStreamSynchronize(stream));
    out->sahInner = hsah[2] > 0 ? (float)(hsah[0] / hsah[2]) : 0.f;
    out->sahLeaf = hsah[2] > 0 ? (float)(hsah[1] / hsah[2]) : 0.f;
    out->numNodes = hs[0];
    out->numLeaves = hs[1];
    out->maxLeaf = hs[2];
    out->topCount = out->top ? hs[3] : 0u;
    out->rootRef = 0u;
  }
  VR_CK(cudaEventRecord(e1, stream));
  VR_CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&out->buildMs, e0, e1);
  cleanup();
  return cudaSuccess;
}

// The quality of a Morton-ordered tree depends on the shape of the Morton cells: cells
// stretched like the scene box (every axis normalised by its own extent) suit a trench,
// cubic cells a field of deep holes.  So a few cell shapes are built -- a build is a few
// ms -- and the tree with the lowest surface-area-heuristic cost is kept
// (cost = sum of inner-node areas + 0.4 x sum of leaf areas x primitives; 0.4 is the
// measured instruction ratio of a disk test to a node visit).
cudaError_t buildBvh(const float4 *primLo, const float4 *primHi, uint32_t n, const float sceneLo[3],
                     const float sceneHi[3], uint32_t leafMax, float alphaHint, cudaStream_t stream,
                     Bvh *out) {
  float ext[3], maxExt = 0.f;
  for (int a = 0; a < 3; ++a) {
    ext[a] = sceneHi[a] - sceneLo[a];
    maxExt = fmaxf(maxExt, ext[a]);
  }
  float alphas[3] = {1.f, 0.5f, 0.f};  // 1: cells shaped like the scene box ... 0: cubic cells
  int numAlpha = 3;
  if (const char *fa = getenv("VR_MORTON_ALPHA")) {
    alphas[0] = (float)atof(fa);
    numAlpha = 1;
  } else if (alphaHint >= 0.f) {  // the shape the caller's previous, similar scene settled on
    alphas[0] = alphaHint;
    numAlpha = 1;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  Bvh best;
  float bestCost = 3.402823466e+38f;
  cudaError_t err = cudaSuccess;
  for (int k = 0; k < numAlpha; ++k) {
    float inv[3];
    for (int a = 0; a < 3; ++a)
      inv[a] = ext[a] > 0.f ? 1.f / (powf(ext[a], alphas[k]) * powf(maxExt, 1.f - alphas[k])) : 0.f;
    Bvh cand;
    err = buildOne(primLo, primHi, n, sceneLo, sceneHi, make_float3(inv[0], inv[1], inv[2]),
                   leafMax, stream, &cand);
    if (err != cudaSuccess) {
      freeBvh(&cand, stream);
      break;
    }
    // the heuristic is only trusted for clear differences (a few per cent of SAH cost do
    // not predict the measured rate): another shape must beat the first by 5 %
    const float cost = (cand.sahInner + 0.4f * cand.sahLeaf) * (k == 0 ? 0.95f : 1.f);
    if (cost < bestCost || k == 0) {
      freeBvh(&best, stream);
      best = cand;
      bestCost = cost;
      best.mortonAlpha = alphas[k];
    } else {
      freeBvh(&cand, stream);
    }
    if (n <= leafMax)
      break;  // a single leaf: nothing to choose
  }
  cudaEventRecord(e1, stream);
  cudaEventSynchronize(e1);
  if (err == cudaSuccess) {
    freeBvh(out, stream);
    *out = best;
    cudaEventElapsedTime(&out->buildMs, e0, e1);
  } else {
    freeBvh(&best, stream);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return err;
}

void freeBvh(Bvh *b, cudaStream_t stream) {
  cudaFreeAsync(b->nodes, stream);
  cudaFreeAsync(b->nodes4, stream);
  cudaFreeAsync(b->top, stream);
  cudaFreeAsync(b->sortedToOrig, stream);
  *b = Bvh();
}

}  // namespace vr
