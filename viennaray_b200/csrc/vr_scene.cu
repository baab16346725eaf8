// Device-side scene preparation: the caller's arrays are uploaded once in their
// original order; packing into 32-byte primitive records, bounding boxes, the
// permutation into BVH (Morton) order and the remapping of the neighbour lists
// all run as kernels, so vr_scene_commit moves no primitive data through host
// loops.  Replaces what GeometryDisk / GeometryTriangle::initGeometry hand to
// Embree (rayGeometryDisk.hpp:102-193, rayGeometryTriangle.hpp:15-92).
#include <cub/device/device_scan.cuh>

#include "vr_internal.h"

namespace vr {

namespace {

// disk i: A = {x,y,z,r} is the caller's xyzr row; B = {nx,ny,nz,original ID}
__global__ void packDiskNormalsKernel(const float *nxyz, uint32_t n, float4 *B) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    B[i] = make_float4(nxyz[3 * i], nxyz[3 * i + 1], nxyz[3 * i + 2], __uint_as_float(i));
}

// triangle i: v0, v1, v2 (w = original ID) and the unit normal
__global__ void packTrianglesKernel(const float *verts, const uint32_t *tris, const float *normals,
                                    uint32_t n, float4 *A, float4 *B, float4 *C, float4 *N) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float w = __uint_as_float(i);
  const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
  A[i] = make_float4(verts[3 * a], verts[3 * a + 1], verts[3 * a + 2], w);
  B[i] = make_float4(verts[3 * b], verts[3 * b + 1], verts[3 * b + 2], w);
  C[i] = make_float4(verts[3 * c], verts[3 * c + 1], verts[3 * c + 2], w);
  N[i] = make_float4(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2], w);
}

__global__ void gatherDisksKernel(const float4 *A, const float4 *B, const uint32_t *s2o, uint32_t n,
                                  float4 *prim) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t o = s2o[i];
  prim[2 * i] = A[o];
  prim[2 * i + 1] = B[o];
}

__global__ void gatherTrianglesKernel(const float4 *A, const float4 *B, const float4 *C,
                                      const float4 *N, const uint32_t *s2o, uint32_t n,
                                      float4 *prim) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t o = s2o[i];
  prim[4 * i] = A[o];
  prim[4 * i + 1] = B[o];
  prim[4 * i + 2] = C[o];
  prim[4 * i + 3] = N[o];
}

// o2s = inverse of s2o; cnt[i] = neighbour count of the primitive at sorted slot i
__global__ void invertAndCountKernel(const uint32_t *s2o, const uint32_t *offO, uint32_t n,
                                     uint32_t *o2s, uint32_t *cnt) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t o = s2o[i];
  o2s[o] = i;
  cnt[i] = offO ? offO[o + 1] - offO[o] : 0u;
}

__global__ void remapRowsKernel(const uint32_t *s2o, const uint32_t *o2s, const uint32_t *offO,
                                const uint32_t *idxO, const uint32_t *off, uint32_t n,
                                uint32_t *idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t o = s2o[i];
  uint32_t w = off[i];
  for (uint32_t k = offO[o]; k < offO[o + 1]; ++k)
    idx[w++] = o2s[idxO[k]];
}

// off[n] = off[n-1] + cnt[n-1] (the exclusive scan leaves the total out)
__global__ void closeOffsetsKernel(const uint32_t *cnt, uint32_t n, uint32_t *off) {
  off[n] = off[n - 1] + cnt[n - 1];
}

}  // namespace

cudaError_t launchPackDiskNormals(const float *nxyz, uint32_t n, float4 *B, cudaStream_t s) {
  packDiskNormalsKernel<<<(n + 255) / 256, 256, 0, s>>>(nxyz, n, B);
  return cudaGetLastError();
}

cudaError_t launchPackTriangles(const float *verts, const uint32_t *tris, const float *normals,
                                uint32_t n, float4 *A, float4 *B, float4 *C, float4 *N,
                                cudaStream_t s) {
  packTrianglesKernel<<<(n + 255) / 256, 256, 0, s>>>(verts, tris, normals, n, A, B, C, N);
  return cudaGetLastError();
}

cudaError_t launchGatherPrims(int geoType, const float4 *A, const float4 *B, const float4 *C,
                              const float4 *N, const uint32_t *s2o, uint32_t n, float4 *prim,
                              cudaStream_t s) {
  if (geoType == 0)
    gatherDisksKernel<<<(n + 255) / 256, 256, 0, s>>>(A, B, s2o, n, prim);
  else
    gatherTrianglesKernel<<<(n + 255) / 256, 256, 0, s>>>(A, B, C, N, s2o, n, prim);
  return cudaGetLastError();
}

// Neighbour CSR from original to internal (BVH order) indices.  off: n+1 words,
// idx: as many words as idxO.  tmp (n words) and o2s (n words) are scratch.
cudaError_t remapNeighbors(const uint32_t *s2o, const uint32_t *offO, const uint32_t *idxO,
                           uint32_t n, uint32_t *o2s, uint32_t *cnt, uint32_t *off, uint32_t *idx,
                           cudaStream_t s) {
  invertAndCountKernel<<<(n + 255) / 256, 256, 0, s>>>(s2o, offO, n, o2s, cnt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  size_t bytes = 0;
  e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, off, (int)n, s);
  if (e != cudaSuccess)
    return e;
  void *tmp = nullptr;
  e = cudaMallocAsync(&tmp, bytes ? bytes : 16, s);
  if (e != cudaSuccess)
    return e;
  e = cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, off, (int)n, s);
  cudaFreeAsync(tmp, s);
  if (e != cudaSuccess)
    return e;
  closeOffsetsKernel<<<1, 1, 0, s>>>(cnt, n, off);
  if (offO && idxO)
    remapRowsKernel<<<(n + 255) / 256, 256, 0, s>>>(s2o, o2s, offO, idxO, off, n, idx);
  return cudaGetLastError();
}

}  // namespace vr
