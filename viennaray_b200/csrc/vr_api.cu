// C ABI of the library (include/viennaray_b200.h): context, scene upload,
// device BVH build, trace launches and result download.  Host-side C++ only
// prepares buffers; all per-ray work runs in the sm_100a kernels of
// vr_trace.cu.  No CPU fallback exists: every entry point needs a device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include "vr_internal.h"

using namespace vr;

// NCCL, bound at run time (dlopen) so that the library carries no link-time dependency and a
// process that already holds an NCCL -- torch's -- shares it.  Only the five entry points the
// flux all-reduce needs; types as in nccl.h.
typedef struct ncclComm *vrNcclComm;
struct NcclApi {
  void *lib = nullptr;
  int (*commInitAll)(vrNcclComm *, int, const int *) = nullptr;
  int (*commDestroy)(vrNcclComm) = nullptr;
  int (*allReduce)(const void *, void *, size_t, int, int, vrNcclComm, cudaStream_t) = nullptr;
  int (*groupStart)() = nullptr;
  int (*groupEnd)() = nullptr;
  const char *(*getErrorString)(int) = nullptr;
};
static const int VR_NCCL_UINT64 = 5, VR_NCCL_SUM = 0;  // ncclUint64, ncclSum (nccl.h)

#define VR_STAGE_CHUNK ((size_t)8 << 20)  // bytes of one pinned staging chunk (two per context)

struct vr_ctx {
  // A multi-device context (vr_ctx_create_multi) is a parent that owns one ordinary context
  // per device: scene calls are replicated, a trace shards the ray-index range over the
  // children (one host thread each), one NCCL all-reduce sums the result words, and the
  // downloads read the first child.  `children` is empty for an ordinary context.
  std::vector<vr_ctx *> children;
  std::vector<vrNcclComm> comms;
  NcclApi nccl;
  int device = 0;
  int numSMs = 0;
  char *hStage = nullptr;  // pinned staging of the uploads (stagedCopy), allocated on first use
  cudaEvent_t stageEv[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;

  // the scene as set by the caller, resident on the device in original order
  int geoType = -1;
  uint32_t n = 0, numVerts = 0;
  float *dXyzr = nullptr, *dNxyz = nullptr;               // disks: n x 4, n x 3
  float *dVerts = nullptr, *dTriN = nullptr;              // triangles: V x 3, n x 3
  uint32_t *dTris = nullptr;                              // n x 3
  uint32_t *dNbOffO = nullptr, *dNbIdxO = nullptr;        // neighbour CSR, original indices
  size_t nbTotal = 0;
  std::vector<int32_t> materialIds;
  float geoLo[3] = {0, 0, 0}, geoHi[3] = {0, 0, 0};
  bool boundarySet = false;
  int D = 3;

  // device scene (internal = BVH order)
  float4 *dPrim = nullptr;
  uint32_t *dNbOff = nullptr, *dNbIdx = nullptr, *dNbRow = nullptr;
  int *dMatId = nullptr;       // material IDs, internal order
  float *dMatTab = nullptr;    // sticking tables of the running trace (all particles)
  size_t matTabCap = 0;
  std::vector<size_t> matTabOffset;
  Bvh bvh;
  DeviceScene scene{};
  bool committed = false;

  // results of the last trace.  dResult: what the kernels add into, np*n flux words in
  // internal (BVH) order + np*8 counters.  dFluxOrig: the same words in the CALLER's
  // primitive order + the counters, written at the end of every trace; this is the buffer
  // vr_flux_device exposes (a multi-GPU caller all-reduces it in place: its layout does
  // not depend on the BVH a rank happened to build) and every download reads.
  unsigned long long *dResult = nullptr;
  unsigned long long *dFluxOrig = nullptr;
  size_t resultN = 0;       // primitives / particles the two buffers were sized for
  int resultNp = 0;
  bool resultValid = false; // dFluxOrig holds the flux of the committed scene's last trace
  unsigned long long *dCursor = nullptr;   // [0] ray cursor
  unsigned int *dSlotCursor = nullptr;      // [0] slot cursor, [1] live count
  unsigned long long *dCounterCopies = nullptr;  // VR_COUNTER_COPIES x 8
  unsigned int *hLive = nullptr;            // pinned ring of live counts
  cudaEvent_t liveEv[4] = {nullptr, nullptr, nullptr, nullptr};
  RayPool pool{}, pool2{};
  unsigned long long *dWork = nullptr;
  size_t resultWords = 0;
  int numParticles = 0;
  uint64_t lastNumRays = 0;
  float lastMs = 0.f;
  int kernelLaunches = 0;
  int iterations = 0;
  bool countWork = false;
  float *dGrid = nullptr;  // grid source origins
  uint32_t gridN = 0;
  // sky map (escape culling), built lazily for the source axis / side of a trace
  float2 *dSky = nullptr;
  int skyAxis = -1;
  float skySign = 0.f, skyTop = 0.f;
  uint32_t tailRays = 262144;  // VR_TAIL_RAYS: survivors handed to the one-launch tail kernel
  int skyCells = 128;  // VR_SKY_CELLS; 0 disables the map
  bool dumpLaunches = false;  // VR_DUMP_LAUNCHES=1: one line per wavefront iteration (debug)
  bool timeKernels = false;  // VR_TIME_KERNELS=1: CUDA events around every launch
  std::vector<cudaEvent_t> tev;
  std::vector<int> tevKind;
  double phaseMs[3] = {0, 0, 0};  // traverse, shade, other (accumulated)
  long long phaseLaunches[3] = {0, 0, 0};
  uint32_t poolSlots = 1u << 24;
  // neighbour spread as its own kernel (queue of 2 x float4 per pool slot): -1 = when the
  // disks and their neighbour lists do not fit the L2 cache (the gathers then wait on DRAM
  // and gain from full warps: +5 % on the 4M-disk hole array, -1.4 % on the L2-resident
  // 1M-disk trench), 0 / 1 = VR_SPREAD_SPLIT
  float4 *dSpreadQ = nullptr;
  uint32_t spreadCap = 0;
  int spreadMode = -1;
  size_t l2Bytes = 0;
  // Extra wavefront lanes: a trace runs its jobs (one per particle; a lone particle's ray
  // range is cut in two) two at a time, each on its own stream, pool pair and cursors, driven
  // by its own host thread, so that the thin start and end of one wavefront (few rays in
  // flight) and the tail of every persistent traverse launch are covered by the other lane's
  // kernels.  Allocated by the first trace that uses them.  Rays are independent and the sums
  // are integers: the result does not depend on the lanes.
  struct Lane1 {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr, liveEv = nullptr;
    RayPool pool{}, pool2{};
    unsigned long long *dCursor = nullptr;
    unsigned int *dSlotCursor = nullptr;
    unsigned long long *dCounterCopies = nullptr;
    unsigned int *hLive = nullptr;
    float4 *dSpreadQ = nullptr;
    uint32_t spreadCap = 0;
  } xlane[3];
  int lanes = 2;  // VR_LANES (1..4); 1: one particle after the other on one stream
  float alphaLast = 1.f;  // Morton cell shape of the last full search (vr_scene_commit)
  uint32_t alphaN = 0;
  int alphaGeo = -1, alphaAge = 0;
};

static std::string g_createError;

static int fail(vr_ctx *ctx, int code, const std::string &msg) {
  if (ctx)
    ctx->err = msg;
  else
    g_createError = msg;
  return code;
}
static int failCuda(vr_ctx *ctx, cudaError_t e, const char *what) {
  return fail(ctx, VR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return failCuda(ctx, e_, #call);                                                             \
  } while (0)

static void freeDeviceScene(vr_ctx *c) {
  cudaFreeAsync(c->dPrim, c->stream);
  cudaFreeAsync(c->dNbOff, c->stream);
  cudaFreeAsync(c->dNbIdx, c->stream);
  cudaFreeAsync(c->dNbRow, c->stream);
  cudaFreeAsync(c->dMatId, c->stream);
  c->dMatId = nullptr;
  c->dPrim = nullptr;
  c->dNbOff = c->dNbIdx = c->dNbRow = nullptr;
  freeBvh(&c->bvh, c->stream);
  cudaFreeAsync(c->dSky, c->stream);
  c->dSky = nullptr;
  c->skyAxis = -1;
  c->committed = false;
}
static void freeInputs(vr_ctx *c) {
  cudaFreeAsync(c->dXyzr, c->stream);
  cudaFreeAsync(c->dNxyz, c->stream);
  cudaFreeAsync(c->dVerts, c->stream);
  cudaFreeAsync(c->dTriN, c->stream);
  cudaFreeAsync(c->dTris, c->stream);
  cudaFreeAsync(c->dNbOffO, c->stream);
  cudaFreeAsync(c->dNbIdxO, c->stream);
  c->dXyzr = c->dNxyz = c->dVerts = c->dTriN = nullptr;
  c->dTris = c->dNbOffO = c->dNbIdxO = nullptr;
  c->nbTotal = 0;
}
// Host -> device copy of a caller array.  The caller's memory is pageable; large arrays go
// through two pinned staging chunks owned by the context, so that the CPU copy of one chunk
// overlaps the DMA of the other and several contexts (one per GPU) uploading at once do not
// queue behind the driver's own staging buffer.
static cudaError_t stagedCopy(vr_ctx *c, void *dst, const void *src, size_t bytes) {
  const size_t CH = VR_STAGE_CHUNK;
  if (bytes < (1u << 20))
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
  if (!c->hStage) {
    cudaError_t e = cudaMallocHost(&c->hStage, 2 * CH);
    if (e == cudaSuccess)
      e = cudaEventCreateWithFlags(&c->stageEv[0], cudaEventDisableTiming);
    if (e == cudaSuccess)
      e = cudaEventCreateWithFlags(&c->stageEv[1], cudaEventDisableTiming);
    if (e != cudaSuccess)
      return e;
  }
  size_t off = 0;
  int k = 0;
  while (off < bytes) {
    const size_t m = std::min(CH, bytes - off);
    cudaError_t e = cudaEventSynchronize(c->stageEv[k]);  // the chunk's previous DMA is done
    if (e != cudaSuccess)
      return e;
    memcpy(c->hStage + (size_t)k * CH, (const char *)src + off, m);
    e = cudaMemcpyAsync((char *)dst + off, c->hStage + (size_t)k * CH, m, cudaMemcpyHostToDevice,
                        c->stream);
    if (e == cudaSuccess)
      e = cudaEventRecord(c->stageEv[k], c->stream);
    if (e != cudaSuccess)
      return e;
    off += m;
    k ^= 1;
  }
  return cudaSuccess;
}
// stream-ordered upload of a caller array into a fresh device buffer
template <class T>
static cudaError_t uploadArray(vr_ctx *c, T **dst, const T *src, size_t count) {
  cudaError_t e = cudaMallocAsync((void **)dst, sizeof(T) * std::max<size_t>(count, 1), c->stream);
  if (e == cudaSuccess && count)
    e = stagedCopy(c, *dst, src, sizeof(T) * count);
  return e;
}
static void freeResults(vr_ctx *c) {
  cudaFree(c->dResult);
  cudaFree(c->dFluxOrig);
  c->dResult = c->dFluxOrig = nullptr;
  c->resultWords = 0;
  c->resultN = 0;
  c->resultNp = 0;
  c->resultValid = false;
  c->numParticles = 0;
}

static void freeOnePool(RayPool &q) {
  cudaFree(q.od0);
  cudaFree(q.od1);
  cudaFree(q.hit);
  cudaFree(q.rng);
  cudaFree(q.meta);
  cudaFree(q.weight);
  cudaFree(q.dir3);
  q = RayPool{};
}
static void freePool(vr_ctx *c) {
  freeOnePool(c->pool);
  freeOnePool(c->pool2);
  cudaFree(c->dSpreadQ);
  c->dSpreadQ = nullptr;
  c->spreadCap = 0;
  for (auto &l : c->xlane) {
    freeOnePool(l.pool);
    freeOnePool(l.pool2);
    cudaFree(l.dSpreadQ);
    l.dSpreadQ = nullptr;
    l.spreadCap = 0;
  }
}
static cudaError_t allocOnePool(RayPool &q, uint32_t slots) {
  cudaError_t e;
  if ((e = cudaMalloc(&q.od0, sizeof(float4) * (size_t)slots)) != cudaSuccess ||
      (e = cudaMalloc(&q.od1, sizeof(float2) * (size_t)slots)) != cudaSuccess ||
      (e = cudaMalloc(&q.hit, sizeof(float4) * (size_t)slots)) != cudaSuccess ||
      (e = cudaMalloc(&q.rng, sizeof(uint32_t) * (size_t)slots)) != cudaSuccess ||
      (e = cudaMalloc(&q.meta, sizeof(uint4) * (size_t)slots)) != cudaSuccess ||
      (e = cudaMalloc(&q.weight, sizeof(float) * (size_t)slots)) != cudaSuccess ||
      (e = cudaMalloc(&q.dir3, sizeof(float4) * (size_t)slots)) != cudaSuccess)
    return e;
  q.capacity = slots;
  return cudaSuccess;
}
// two pools: the tail of a trace compacts survivors from one into the other
static cudaError_t ensurePool(vr_ctx *c, uint32_t slots) {
  if (c->pool.capacity >= slots && c->pool.od0 && c->pool2.od0)
    return cudaSuccess;
  freePool(c);
  cudaError_t e = allocOnePool(c->pool, slots);
  if (e == cudaSuccess)
    e = allocOnePool(c->pool2, slots);
  if (e != cudaSuccess)
    freePool(c);
  return e;
}
// an extra lane's stream, cursors and pools (first use)
static cudaError_t ensureLane(vr_ctx *c, int which, uint32_t slots, bool spreadSplit) {
  vr_ctx::Lane1 &l = c->xlane[which];
  cudaError_t e = cudaSuccess;
  if (!l.stream) {
    if ((e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&l.liveEv, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaMalloc(&l.dCursor, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&l.dSlotCursor, 8 * sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaMalloc(&l.dCounterCopies,
                        VR_COUNTER_COPIES * 8 * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMallocHost(&l.hLive, 8 * sizeof(unsigned int))) != cudaSuccess)
      return e;
  }
  if (!(l.pool.capacity >= slots && l.pool.od0 && l.pool2.od0)) {
    freeOnePool(l.pool);
    freeOnePool(l.pool2);
    if ((e = allocOnePool(l.pool, slots)) != cudaSuccess ||
        (e = allocOnePool(l.pool2, slots)) != cudaSuccess) {
      freeOnePool(l.pool);
      freeOnePool(l.pool2);
      return e;
    }
  }
  if (spreadSplit && l.spreadCap < slots) {
    cudaFree(l.dSpreadQ);
    l.dSpreadQ = nullptr;
    l.spreadCap = 0;
    if ((e = cudaMalloc(&l.dSpreadQ, sizeof(float4) * 2 * (size_t)slots)) != cudaSuccess)
      return e;
    l.spreadCap = slots;
  }
  return cudaSuccess;
}

__global__ void gatherMaterialsKernel(const int *orig, const uint32_t *sortedToOrig, uint32_t n,
                                      int *out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = orig[sortedToOrig[i]];
}

__global__ void reduceCountersKernel(const unsigned long long *copies, unsigned long long *out) {
  const int k = threadIdx.x;
  if (k < 8) {
    unsigned long long v = 0;
    for (int c = 0; c < VR_COUNTER_COPIES; ++c)
      v += copies[c * 8 + k];
    atomicAdd(&out[k], v);  // (a particle's ray range may be traced as several jobs)
  }
}

// optional per-phase timing: an event after every launch, classified by kind
static void mark(vr_ctx *c, int kind) {
  if (!c->timeKernels)
    return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, c->stream);
  c->tev.push_back(e);
  c->tevKind.push_back(kind);
}
static void collectMarks(vr_ctx *c) {
  if (!c->timeKernels || c->tev.empty())
    return;
  cudaEventSynchronize(c->tev.back());
  for (size_t i = 1; i < c->tev.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->tev[i - 1], c->tev[i]);
    c->phaseMs[c->tevKind[i]] += ms;
    c->phaseLaunches[c->tevKind[i]] += 1;
  }
  for (auto e : c->tev)
    cudaEventDestroy(e);
  c->tev.clear();
  c->tevKind.clear();
}

// ---- multi-device parent: helpers ---------------------------------------------------
static bool loadNccl(NcclApi &a, std::string &why) {
  const char *names[] = {getenv("VR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm)
      continue;
    a.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (a.lib)
      break;
  }
  if (!a.lib) {
    why = "NCCL not found (libnccl.so.2; set VR_NCCL_LIB to its path)";
    return false;
  }
  a.commInitAll = (int (*)(vrNcclComm *, int, const int *))dlsym(a.lib, "ncclCommInitAll");
  a.commDestroy = (int (*)(vrNcclComm))dlsym(a.lib, "ncclCommDestroy");
  a.allReduce = (int (*)(const void *, void *, size_t, int, int, vrNcclComm, cudaStream_t))dlsym(
      a.lib, "ncclAllReduce");
  a.groupStart = (int (*)())dlsym(a.lib, "ncclGroupStart");
  a.groupEnd = (int (*)())dlsym(a.lib, "ncclGroupEnd");
  a.getErrorString = (const char *(*)(int))dlsym(a.lib, "ncclGetErrorString");
  if (!a.commInitAll || !a.commDestroy || !a.allReduce || !a.groupStart || !a.groupEnd) {
    why = "NCCL library lacks an expected symbol";
    return false;
  }
  return true;
}
// first failing child's message becomes the parent's
static int childFail(vr_ctx *parent, vr_ctx *child, int rc) {
  parent->err = "device " + std::to_string(child->device) + ": " + child->err;
  return rc;
}
#define EACH_CHILD(call)                                                                           \
  do {                                                                                             \
    for (vr_ctx * ch : ctx->children) {                                                            \
      int rc_ = (call);                                                                            \
      if (rc_ != VR_OK)                                                                            \
        return childFail(ctx, ch, rc_);                                                            \
    }                                                                                              \
    return VR_OK;                                                                                  \
  } while (0)

extern "C" {

const char *vr_last_error(const vr_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_createError.c_str();
}

int vr_ctx_create(int cudaDevice, vr_ctx **out) {
  if (!out)
    return fail(nullptr, VR_ERR_ARGUMENT, "vr_ctx_create: out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, VR_ERR_CUDA,
                std::string("vr_ctx_create: no CUDA device (") + cudaGetErrorString(e) +
                    "); this library has no CPU path");
  if (cudaDevice < 0 || cudaDevice >= count)
    return fail(nullptr, VR_ERR_ARGUMENT, "vr_ctx_create: device index out of range");
  e = cudaSetDevice(cudaDevice);
  if (e != cudaSuccess)
    return failCuda(nullptr, e, "cudaSetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, cudaDevice);
  if (e != cudaSuccess)
    return failCuda(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major < 10)
    return fail(nullptr, VR_ERR_CUDA,
                "vr_ctx_create: kernels are built for sm_100a only; device is sm_" +
                    std::to_string(prop.major * 10 + prop.minor));
  {
    // keep freed stream-ordered allocations cached: scene commits reuse them
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, cudaDevice) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  vr_ctx *ctx = new vr_ctx();
  ctx->device = cudaDevice;
  ctx->numSMs = prop.multiProcessorCount;
  ctx->l2Bytes = (size_t)prop.l2CacheSize;
  const char *cw = getenv("VR_COUNT_WORK");
  ctx->countWork = cw && cw[0] == '1';
  if (const char *tr = getenv("VR_TAIL_RAYS")) {
    long v = atol(tr);
    if (v >= 0 && v <= (1l << 24))
      ctx->tailRays = (uint32_t)v;
  }
  if (const char *sk = getenv("VR_SKY_CELLS")) {
    long v = atol(sk);
    if (v >= 0 && v <= 1024)
      ctx->skyCells = (int)v;
  }
  if (const char *ss = getenv("VR_SPREAD_SPLIT"))
    ctx->spreadMode = ss[0] == '1' ? 1 : 0;
  const char *tk = getenv("VR_TIME_KERNELS");
  ctx->timeKernels = tk && tk[0] == '1';
  if (const char *dl = getenv("VR_DUMP_LAUNCHES"))
    ctx->dumpLaunches = dl[0] == '1';
  if (const char *ln = getenv("VR_LANES")) {
    const int v = atoi(ln);
    if (v >= 1 && v <= 4)
      ctx->lanes = v;
  }
  if (const char *ps = getenv("VR_POOL_SLOTS")) {
    long v = atol(ps);
    if (v >= 1024 && v <= (1l << 26))
      ctx->poolSlots = (uint32_t)v;
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->dCursor, sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMalloc(&ctx->dSlotCursor, 8 * sizeof(unsigned int))) != cudaSuccess ||
      (e = cudaMalloc(&ctx->dCounterCopies,
                      VR_COUNTER_COPIES * 8 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMallocHost(&ctx->hLive, 8 * sizeof(unsigned int))) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->liveEv[0], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->liveEv[1], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->liveEv[2], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->liveEv[3], cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->dWork, 8 * sizeof(unsigned long long))) != cudaSuccess) {
    failCuda(nullptr, e, "vr_ctx_create");
    vr_ctx_destroy(ctx);
    return VR_ERR_CUDA;
  }
  *out = ctx;
  return VR_OK;
}

int vr_ctx_create_multi(int nDevices, const int *deviceIds, vr_ctx **out) {
  if (!out)
    return fail(nullptr, VR_ERR_ARGUMENT, "vr_ctx_create_multi: out is null");
  *out = nullptr;
  if (nDevices < 1 || !deviceIds)
    return fail(nullptr, VR_ERR_ARGUMENT, "vr_ctx_create_multi: no devices given");
  for (int i = 0; i < nDevices; ++i)
    for (int j = 0; j < i; ++j)
      if (deviceIds[i] == deviceIds[j])
        return fail(nullptr, VR_ERR_ARGUMENT, "vr_ctx_create_multi: a device is listed twice");
  if (nDevices == 1)
    return vr_ctx_create(deviceIds[0], out);
  vr_ctx *parent = new vr_ctx();
  parent->device = deviceIds[0];
  for (int i = 0; i < nDevices; ++i) {
    vr_ctx *ch = nullptr;
    int rc = vr_ctx_create(deviceIds[i], &ch);
    if (rc != VR_OK) {
      vr_ctx_destroy(parent);
      return rc;  // g_createError holds the message
    }
    parent->children.push_back(ch);
  }
  std::string why;
  if (!loadNccl(parent->nccl, why)) {
    vr_ctx_destroy(parent);
    return fail(nullptr, VR_ERR_UNSUPPORTED, "vr_ctx_create_multi: " + why);
  }
  parent->comms.assign((size_t)nDevices, nullptr);
  int nrc = parent->nccl.commInitAll(parent->comms.data(), nDevices, deviceIds);
  if (nrc != 0) {
    std::string msg = parent->nccl.getErrorString ? parent->nccl.getErrorString(nrc) : "error";
    parent->comms.clear();
    vr_ctx_destroy(parent);
    return fail(nullptr, VR_ERR_CUDA, "vr_ctx_create_multi: ncclCommInitAll: " + msg);
  }
  *out = parent;
  return VR_OK;
}

int vr_ctx_num_devices(const vr_ctx *ctx) {
  return !ctx ? 0 : (ctx->children.empty() ? 1 : (int)ctx->children.size());
}

void vr_ctx_destroy(vr_ctx *ctx) {
  if (!ctx)
    return;
  if (!ctx->children.empty() || ctx->nccl.lib) {  // multi-device parent
    for (size_t i = 0; i < ctx->comms.size(); ++i)
      if (ctx->comms[i])
        ctx->nccl.commDestroy(ctx->comms[i]);
    for (vr_ctx *ch : ctx->children)
      vr_ctx_destroy(ch);
    // the NCCL handle stays open: unloading it under a process that shares it is not safe
    delete ctx;
    return;
  }
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    freeDeviceScene(ctx);
    freeInputs(ctx);
    cudaFreeAsync(ctx->dGrid, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  freeResults(ctx);
  cudaFreeHost(ctx->hStage);
  for (auto &e : ctx->stageEv)
    if (e)
      cudaEventDestroy(e);
  cudaFree(ctx->dCursor);
  cudaFree(ctx->dSlotCursor);
  cudaFree(ctx->dCounterCopies);
  cudaFreeHost(ctx->hLive);
  for (auto &e : ctx->liveEv)
    if (e)
      cudaEventDestroy(e);
  freePool(ctx);
  for (auto &l : ctx->xlane) {
    cudaFree(l.dCursor);
    cudaFree(l.dSlotCursor);
    cudaFree(l.dCounterCopies);
    cudaFreeHost(l.hLive);
    if (l.done)
      cudaEventDestroy(l.done);
    if (l.liveEv)
      cudaEventDestroy(l.liveEv);
    if (l.stream)
      cudaStreamDestroy(l.stream);
  }
  cudaFree(ctx->dMatTab);
  cudaFree(ctx->dWork);
  if (ctx->ev0)
    cudaEventDestroy(ctx->ev0);
  if (ctx->ev1)
    cudaEventDestroy(ctx->ev1);
  if (ctx->stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void *vr_ctx_stream(vr_ctx *ctx) {
  if (ctx && !ctx->children.empty())
    return vr_ctx_stream(ctx->children[0]);
  return ctx ? (void *)ctx->stream : nullptr;
}
int vr_ctx_synchronize(vr_ctx *ctx) {
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!ctx->children.empty())
    EACH_CHILD(vr_ctx_synchronize(ch));
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}
float vr_last_kernel_ms(vr_ctx *ctx) {
  if (ctx && !ctx->children.empty())
    return ctx->lastMs;  // slowest device of the last trace
  return ctx ? ctx->lastMs : 0.f;
}
int vr_last_launch_count(vr_ctx *ctx, int *kernelsOut, int *iterationsOut) {
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!ctx->children.empty()) {  // summed over the devices
    int k = 0, it = 0;
    for (vr_ctx *ch : ctx->children) {
      k += ch->kernelLaunches;
      it += ch->iterations;
    }
    if (kernelsOut)
      *kernelsOut = k;
    if (iterationsOut)
      *iterationsOut = it;
    return VR_OK;
  }
  if (kernelsOut)
    *kernelsOut = ctx->kernelLaunches;
  if (iterationsOut)
    *iterationsOut = ctx->iterations;
  return VR_OK;
}

int vr_scene_set_disks(vr_ctx *ctx, const float *xyzr, const float *nxyz, uint32_t n,
                       const int32_t *materialIds, const uint32_t *nbOffsets,
                       const uint32_t *nbIndices) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_scene_set_disks(ch, xyzr, nxyz, n, materialIds, nbOffsets, nbIndices));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!xyzr || !nxyz || n == 0)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_disks: no geometry was passed");
  if (n >= (1u << 27))
    return fail(ctx, VR_ERR_UNSUPPORTED, "vr_scene_set_disks: more than 2^27 primitives");
  if ((nbOffsets == nullptr) != (nbIndices == nullptr) && nbOffsets && nbOffsets[n] != 0)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_disks: nbOffsets without nbIndices");
  size_t total = 0;
  if (nbOffsets) {
    if (nbOffsets[0] != 0)
      return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_disks: nbOffsets[0] must be 0");
    for (uint32_t i = 0; i < n; ++i)
      if (nbOffsets[i + 1] < nbOffsets[i])
        return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_disks: nbOffsets not monotone");
    total = nbOffsets[n];
    for (size_t k = 0; k < total; ++k)
      if (nbIndices[k] >= n)
        return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_disks: neighbour index out of range");
  }
  CK(cudaSetDevice(ctx->device));
  ctx->committed = false;
  ctx->resultValid = false;  // results belong to the scene that was traced
  freeInputs(ctx);
  ctx->geoType = 0;
  ctx->n = n;
  ctx->numVerts = 0;
  for (int a = 0; a < 3; ++a) {
    ctx->geoLo[a] = INFINITY;
    ctx->geoHi[a] = -INFINITY;
  }
  for (uint32_t i = 0; i < n; ++i) {
    const float r = xyzr[4 * i + 3];
    for (int a = 0; a < 3; ++a) {
      const float v = xyzr[4 * i + a];
      ctx->geoLo[a] = std::min(ctx->geoLo[a], v - r);
      ctx->geoHi[a] = std::max(ctx->geoHi[a], v + r);
    }
  }
  CK(uploadArray(ctx, &ctx->dXyzr, xyzr, (size_t)4 * n));
  CK(uploadArray(ctx, &ctx->dNxyz, nxyz, (size_t)3 * n));
  if (nbOffsets) {
    CK(uploadArray(ctx, &ctx->dNbOffO, nbOffsets, (size_t)n + 1));
    CK(uploadArray(ctx, &ctx->dNbIdxO, nbIndices, total));
    ctx->nbTotal = total;
  }
  if (materialIds)
    ctx->materialIds.assign(materialIds, materialIds + n);
  else
    ctx->materialIds.assign(n, 0);
  // the caller's arrays may be reused as soon as this returns
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

int vr_scene_set_triangles(vr_ctx *ctx, const float *verts, uint32_t nVerts, const uint32_t *idx,
                           uint32_t n, const float *normals, const int32_t *materialIds) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_scene_set_triangles(ch, verts, nVerts, idx, n, normals, materialIds));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!verts || !idx || !normals || n == 0 || nVerts == 0)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_triangles: no geometry was passed");
  if (n >= (1u << 27))
    return fail(ctx, VR_ERR_UNSUPPORTED, "vr_scene_set_triangles: more than 2^27 primitives");
  for (size_t k = 0; k < (size_t)3 * n; ++k)
    if (idx[k] >= nVerts)
      return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_triangles: vertex index out of range");
  CK(cudaSetDevice(ctx->device));
  ctx->committed = false;
  ctx->resultValid = false;  // results belong to the scene that was traced
  freeInputs(ctx);
  ctx->geoType = 1;
  ctx->n = n;
  ctx->numVerts = nVerts;
  for (int a = 0; a < 3; ++a) {
    ctx->geoLo[a] = INFINITY;
    ctx->geoHi[a] = -INFINITY;
  }
  for (size_t k = 0; k < (size_t)3 * n; ++k) {  // bounds of the referenced vertices
    const float *p = verts + 3 * (size_t)idx[k];
    for (int a = 0; a < 3; ++a) {
      ctx->geoLo[a] = std::min(ctx->geoLo[a], p[a]);
      ctx->geoHi[a] = std::max(ctx->geoHi[a], p[a]);
    }
  }
  CK(uploadArray(ctx, &ctx->dVerts, verts, (size_t)3 * nVerts));
  CK(uploadArray(ctx, &ctx->dTris, idx, (size_t)3 * n));
  CK(uploadArray(ctx, &ctx->dTriN, normals, (size_t)3 * n));
  if (materialIds)
    ctx->materialIds.assign(materialIds, materialIds + n);
  else
    ctx->materialIds.assign(n, 0);
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

int vr_source_set_grid(vr_ctx *ctx, const float *points, uint32_t n) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_source_set_grid(ch, points, n));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (n && !points)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_source_set_grid: null points");
  CK(cudaSetDevice(ctx->device));
  cudaFreeAsync(ctx->dGrid, ctx->stream);
  ctx->dGrid = nullptr;
  ctx->gridN = 0;
  if (n) {
    CK(uploadArray(ctx, &ctx->dGrid, points, (size_t)3 * n));
    ctx->gridN = n;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

int vr_scene_build_neighbors(vr_ctx *ctx, int D, const float *points, float distance) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_scene_build_neighbors(ch, D, points, distance));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!points || (D != 2 && D != 3) || !(distance > 0.f))
    return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_build_neighbors: invalid argument");
  if (ctx->geoType != 0 || ctx->n == 0)
    return fail(ctx, VR_ERR_STATE, "vr_scene_build_neighbors: set the disks first");
  CK(cudaSetDevice(ctx->device));
  const uint32_t n = ctx->n;
  float lo[3] = {INFINITY, INFINITY, INFINITY};
  for (uint32_t i = 0; i < n; ++i)
    for (int a = 0; a < D; ++a)
      lo[a] = std::min(lo[a], points[3 * (size_t)i + a]);
  ctx->committed = false;
  ctx->resultValid = false;  // results belong to the scene that was traced
  cudaFreeAsync(ctx->dNbOffO, ctx->stream);
  cudaFreeAsync(ctx->dNbIdxO, ctx->stream);
  ctx->dNbOffO = ctx->dNbIdxO = nullptr;
  ctx->nbTotal = 0;
  float *dPts = nullptr;
  CK(uploadArray(ctx, &dPts, points, (size_t)3 * n));
  cudaError_t e = cudaMallocAsync(&ctx->dNbOffO, sizeof(uint32_t) * ((size_t)n + 1), ctx->stream);
  if (e == cudaSuccess)
    e = buildNeighborsDevice(D, dPts, n, lo, distance, ctx->dNbOffO, &ctx->dNbIdxO, &ctx->nbTotal,
                             ctx->stream);
  cudaFreeAsync(dPts, ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_scene_build_neighbors");
  return VR_OK;
}

int vr_scene_get_neighbors(vr_ctx *ctx, uint32_t **offsetsOut, uint32_t **indicesOut) {
  if (ctx && !ctx->children.empty())
    return vr_scene_get_neighbors(ctx->children[0], offsetsOut, indicesOut);
  if (!ctx || !offsetsOut || !indicesOut)
    return VR_ERR_ARGUMENT;
  if (ctx->geoType != 0 || !ctx->dNbOffO)
    return fail(ctx, VR_ERR_STATE, "vr_scene_get_neighbors: no neighbour lists on the device");
  CK(cudaSetDevice(ctx->device));
  const size_t n = ctx->n;
  uint32_t *off = (uint32_t *)malloc(sizeof(uint32_t) * (n + 1));
  uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * std::max<size_t>(ctx->nbTotal, 1));
  cudaError_t e = cudaMemcpyAsync(off, ctx->dNbOffO, sizeof(uint32_t) * (n + 1),
                                  cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && ctx->nbTotal)
    e = cudaMemcpyAsync(idx, ctx->dNbIdxO, sizeof(uint32_t) * ctx->nbTotal, cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    free(off);
    free(idx);
    return failCuda(ctx, e, "vr_scene_get_neighbors");
  }
  *offsetsOut = off;
  *indicesOut = idx;
  return VR_OK;
}

int vr_scene_set_boundary(vr_ctx *ctx, const float bboxMin[3], const float bboxMax[3],
                          int firstDir, int secondDir, int condFirst, int condSecond, int D) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_scene_set_boundary(ch, bboxMin, bboxMax, firstDir, secondDir, condFirst, condSecond, D));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!bboxMin || !bboxMax || (D != 2 && D != 3) || firstDir < 0 || firstDir > 2 ||
      secondDir < 0 || secondDir > 2 || firstDir == secondDir || condFirst < 0 || condFirst > 2 ||
      condSecond < 0 || condSecond > 2)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_scene_set_boundary: invalid argument");
  DeviceScene &s = ctx->scene;
  ctx->D = D;
  s.D = D;
  for (int a = 0; a < 3; ++a) {
    s.bbox[0][a] = bboxMin[a];
    s.bbox[1][a] = bboxMax[a];
  }
  s.firstDir = firstDir;
  s.secondDir = secondDir;
  s.bc[0] = condFirst;
  s.bc[1] = condSecond;
  // vertex and triangle tables of Boundary::initBoundary, rayBoundary.hpp:182-233
  float v[8][3];
  for (int k = 0; k < 8; ++k) {
    int xi = (k == 1 || k == 2 || k == 5 || k == 6), yi = (k == 2 || k == 3 || k == 6 || k == 7),
        zi = k >= 4;
    v[k][0] = s.bbox[xi][0];
    v[k][1] = s.bbox[yi][1];
    v[k][2] = s.bbox[zi][2];
  }
  static const int planes[3][4][3] = {{{0, 3, 7}, {0, 7, 4}, {6, 2, 1}, {6, 1, 5}},
                                      {{0, 4, 5}, {0, 5, 1}, {6, 7, 3}, {6, 3, 2}},
                                      {{0, 1, 2}, {0, 2, 3}, {6, 5, 4}, {6, 4, 7}}};
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) {
        s.btri[i][k][a] = v[planes[firstDir][i][k]][a];
        s.btri[i + 4][k][a] = v[planes[secondDir][i][k]][a];
      }
  // constants of the boundary test's shortcut.  The normal component is formed with the
  // float operations of testTri (vr_device.cuh): e1 = v0 - v1, e2 = v2 - v0,
  // Ng = cross(e2, e1); volatile keeps every intermediate a rounded float
  s.bFast = (D == 3 && firstDir == 0 && secondDir == 1 && !getenv("VR_BOUNDARY_GENERIC")) ? 1 : 0;
  s.bExt = std::max(std::max(s.bbox[1][0] - s.bbox[0][0], s.bbox[1][1] - s.bbox[0][1]),
                    s.bbox[1][2] - s.bbox[0][2]);
  for (int i = 0; i < 8; ++i) {
    const int a = i < 4 ? firstDir : secondDir, b = (a + 1) % 3, c = (a + 2) % 3;
    volatile float e1b = s.btri[i][0][b] - s.btri[i][1][b], e1c = s.btri[i][0][c] - s.btri[i][1][c];
    volatile float e2b = s.btri[i][2][b] - s.btri[i][0][b], e2c = s.btri[i][2][c] - s.btri[i][0][c];
    volatile float p1 = e2b * e1c, p2 = e2c * e1b;
    volatile float na = p1 - p2;  // component a of cross(e2, e1)
    s.bN[i] = na;
    s.bX[i] = s.btri[i][0][a];
  }
  ctx->boundarySet = true;
  return VR_OK;
}

// Everything below the upload runs on the device: 32-byte primitive records,
// padded boxes, Morton sort + LBVH, the permutation of the primitives into BVH
// order and the remapping of the neighbour lists into that index space.
int vr_scene_commit(vr_ctx *ctx) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_scene_commit(ch));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (ctx->geoType < 0)
    return fail(ctx, VR_ERR_STATE, "vr_scene_commit: no geometry was passed");
  if (!ctx->boundarySet)
    return fail(ctx, VR_ERR_STATE, "vr_scene_commit: no boundary was set");
  CK(cudaSetDevice(ctx->device));
  freeDeviceScene(ctx);
  const uint32_t n = ctx->n;
  const bool tri = ctx->geoType == 1;
  cudaStream_t st = ctx->stream;
  float4 *A = nullptr, *B = nullptr, *C = nullptr, *N = nullptr, *lo = nullptr, *hi = nullptr;
  uint32_t *o2s = nullptr, *cnt = nullptr;
  auto tmpFree = [&]() {
    if (tri)
      cudaFreeAsync(A, st);  // disks: A aliases the uploaded xyzr rows
    cudaFreeAsync(B, st);
    cudaFreeAsync(C, st);
    cudaFreeAsync(N, st);
    cudaFreeAsync(lo, st);
    cudaFreeAsync(hi, st);
    cudaFreeAsync(o2s, st);
    cudaFreeAsync(cnt, st);
  };
#define CKT(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      tmpFree();                                                                                   \
      return failCuda(ctx, e_, #call);                                                             \
    }                                                                                              \
  } while (0)
  const size_t bytes = sizeof(float4) * n;
  CKT(cudaMallocAsync(&B, bytes, st));
  CKT(cudaMallocAsync(&lo, bytes, st));
  CKT(cudaMallocAsync(&hi, bytes, st));
  if (tri) {
    CKT(cudaMallocAsync(&A, bytes, st));
    CKT(cudaMallocAsync(&C, bytes, st));
    CKT(cudaMallocAsync(&N, bytes, st));
    CKT(launchPackTriangles(ctx->dVerts, ctx->dTris, ctx->dTriN, n, A, B, C, N, st));
    CKT(launchTriBounds(A, B, C, n, lo, hi, st));
  } else {
    A = reinterpret_cast<float4 *>(ctx->dXyzr);
    CKT(launchPackDiskNormals(ctx->dNxyz, n, B, st));
    CKT(launchDiskBounds(A, B, n, lo, hi, st));
  }
  // A time-stepping caller commits a slowly changing scene over and over: the Morton cell
  // shape the full three-way search picked is reused while the primitive count stays within
  // 1/8 of that scene's, and searched again every 16th commit
  float alphaHint = -1.f;
  if (ctx->alphaN && ctx->alphaGeo == ctx->geoType && ctx->alphaAge < 16 &&
      (n > ctx->alphaN ? n - ctx->alphaN : ctx->alphaN - n) <= ctx->alphaN / 8)
    alphaHint = ctx->alphaLast;
  CKT(buildBvh(lo, hi, n, ctx->geoLo, ctx->geoHi, tri ? VR_LEAF_MAX_TRI : VR_LEAF_MAX, alphaHint,
               st, &ctx->bvh));
  if (alphaHint < 0.f) {
    ctx->alphaLast = ctx->bvh.mortonAlpha;
    ctx->alphaN = n;
    ctx->alphaGeo = ctx->geoType;
    ctx->alphaAge = 0;
  } else {
    ++ctx->alphaAge;
  }
  const size_t per = tri ? 4 : 2;  // float4 records per primitive
  CKT(cudaMallocAsync(&ctx->dPrim, sizeof(float4) * per * n, st));
  CKT(launchGatherPrims(ctx->geoType, A, B, C, N, ctx->bvh.sortedToOrig, n, ctx->dPrim, st));
  CKT(cudaMallocAsync(&ctx->dNbOff, sizeof(uint32_t) * ((size_t)n + 1), st));
  CKT(cudaMallocAsync(&ctx->dNbIdx, sizeof(uint32_t) * std::max<size_t>(ctx->nbTotal, 1), st));
  CKT(cudaMallocAsync(&o2s, sizeof(uint32_t) * n, st));
  CKT(cudaMallocAsync(&cnt, sizeof(uint32_t) * n, st));
  CKT(cudaMallocAsync(&ctx->dNbRow, sizeof(uint32_t) * 8 * (size_t)n, st));
  CKT(remapNeighbors(ctx->bvh.sortedToOrig, ctx->dNbOffO, ctx->dNbIdxO, n, o2s, cnt, ctx->dNbOff,
                     ctx->dNbIdx, ctx->dNbRow, st));
  {  // material IDs into the internal (BVH) order
    int *matOrig = nullptr;
    CKT(uploadArray(ctx, &matOrig, ctx->materialIds.data(), (size_t)n));
    cudaError_t em = cudaMallocAsync(&ctx->dMatId, sizeof(int) * (size_t)n, st);
    if (em == cudaSuccess) {
      gatherMaterialsKernel<<<(n + 255) / 256, 256, 0, st>>>(matOrig, ctx->bvh.sortedToOrig, n,
                                                             ctx->dMatId);
      em = cudaGetLastError();
    }
    cudaFreeAsync(matOrig, st);
    CKT(em);
  }
  tmpFree();
#undef CKT
  CK(cudaStreamSynchronize(st));
  DeviceScene &s = ctx->scene;
  s.geoType = ctx->geoType;
  s.numPrims = n;
  s.prim = ctx->dPrim;
  for (int a = 0; a < 3; ++a) {
    s.qLo[a] = ctx->bvh.qLo[a];
    s.qScale[a] = ctx->bvh.qScale[a];
  }
  s.nbOff = ctx->dNbOff;
  s.nbIdx = ctx->dNbIdx;
  s.nbRow = reinterpret_cast<const uint4 *>(ctx->dNbRow);
  s.nodes = ctx->bvh.nodes;
  s.nodes4 = ctx->bvh.nodes4;
  s.rootRef = ctx->bvh.rootRef;
  s.top = ctx->bvh.top;
  s.topCount = ctx->bvh.topCount;
  s.sky = nullptr;  // rebuilt by the next trace
  ctx->committed = true;
  return VR_OK;
}

static int fillParams(vr_ctx *ctx, const vr_source_desc *src, const vr_particle_desc *part,
                      const vr_config *cfg, int particleIndex, TraceParams &p) {
  if (!src || !part || !cfg)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: null descriptor");
  if (part->kind < 0 || part->kind > VR_PARTICLE_CONED_COSINE)
    return fail(ctx, VR_ERR_UNSUPPORTED,
                "trace: unknown particle kind; only the built-in particles run on the device");
  if (src->rayDir < 0 || src->rayDir > 2 || src->firstDir < 0 || src->firstDir > 2 ||
      src->secondDir < 0 || src->secondDir > 2)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: invalid source axes");
  if (ctx->D == 2 && src->rayDir == 2)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: invalid source direction in 2D geometry");
  if (cfg->rayIdxEnd < cfg->rayIdxBegin || cfg->rayIdxEnd > cfg->numRays)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: ray index shard out of range");
  p.scene = ctx->scene;
  p.src = *src;
  p.particle = *part;
  p.particle.stickingByMaterial = nullptr;  // host pointer: the device copy is p.matSticking
  p.matId = ctx->dMatId;
  p.matSticking = nullptr;
  p.numMaterials = 0;
  if (part->stickingByMaterial && part->numMaterials > 0 && ctx->dMatTab &&
      (size_t)particleIndex < ctx->matTabOffset.size()) {
    p.matSticking = ctx->dMatTab + ctx->matTabOffset[particleIndex];
    p.numMaterials = part->numMaterials;
  }
  p.ee = 1.0f / (part->sourcePower + 1.0f);
  p.eeGrid = 2.0f / (part->sourcePower + 1.0f);
  p.grid = nullptr;
  p.gridN = 0;
  if (src->useGrid) {
    if (!ctx->dGrid || ctx->gridN == 0)
      return fail(ctx, VR_ERR_STATE, "trace: grid source without vr_source_set_grid");
    p.grid = ctx->dGrid;
    p.gridN = ctx->gridN;
  }
  p.idxBegin = cfg->rayIdxBegin;
  p.idxEnd = cfg->rayIdxEnd;
  p.seed = cfg->seed;
  p.stream = (uint32_t)particleIndex;
  p.maxReflections = cfg->maxReflections;
  p.maxBoundaryHits = cfg->maxBoundaryHits;
  p.flags = cfg->flags;
  p.pool = ctx->pool;
  p.poolOut = ctx->pool2;
  p.compact = 0;
  p.numSlots = 0;
  p.rayCursor = ctx->dCursor;
  p.slotCursor = ctx->dSlotCursor;
  p.liveCount = ctx->dSlotCursor + 1;
  p.slotCount = ctx->dSlotCursor + 2;
  p.work = ctx->countWork ? ctx->dWork : nullptr;
  p.spreadQ = nullptr;  // set per trace (vr_trace_device)
  p.spreadCount = ctx->dSlotCursor + 4;
  return VR_OK;
}

// Multi-device trace: device g of G traces the g-th contiguous slice of [rayIdxBegin,
// rayIdxEnd) (a ray's walk depends only on (seed, particle, index), so the union of the slices
// is the single-device job); one host thread per device drives its wavefront, then ONE
// ncclAllReduce(sum) over the uint64 result words (flux in the caller's primitive order +
// counters), queued on every device's stream.  Integer sums: the result is bit-identical to
// one device tracing the whole range.
static int traceMulti(vr_ctx *ctx, const vr_source_desc *src, const vr_particle_desc *particles,
                      int np, const vr_config *cfg, int sync) {
  if (!cfg)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: null descriptor");
  if (cfg->rayIdxEnd < cfg->rayIdxBegin || cfg->rayIdxEnd > cfg->numRays)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: ray index shard out of range");
  const size_t G = ctx->children.size();
  const uint64_t total = cfg->rayIdxEnd - cfg->rayIdxBegin;
  std::vector<int> rcs(G, VR_OK);
  std::vector<std::thread> threads;
  for (size_t g = 0; g < G; ++g) {
    vr_config c = *cfg;
    const uint64_t base = total / G, rem = total % G;
    c.rayIdxBegin = cfg->rayIdxBegin + g * base + std::min<uint64_t>(g, rem);
    c.rayIdxEnd = c.rayIdxBegin + base + (g < rem ? 1 : 0);
    threads.emplace_back([=, &rcs]() {
      rcs[g] = vr_trace_device(ctx->children[g], src, particles, np, &c, 0);
    });
  }
  for (auto &t : threads)
    t.join();
  for (size_t g = 0; g < G; ++g)
    if (rcs[g] != VR_OK)
      return childFail(ctx, ctx->children[g], rcs[g]);
  const size_t words = ctx->children[0]->resultWords;
  int nrc = ctx->nccl.groupStart();
  for (size_t g = 0; g < G && nrc == 0; ++g) {
    vr_ctx *ch = ctx->children[g];
    nrc = ctx->nccl.allReduce(ch->dFluxOrig, ch->dFluxOrig, words, VR_NCCL_UINT64, VR_NCCL_SUM,
                              ctx->comms[g], ch->stream);
  }
  const int erc = ctx->nccl.groupEnd();
  if (nrc == 0)
    nrc = erc;
  if (nrc != 0)
    return fail(ctx, VR_ERR_CUDA, std::string("trace: ncclAllReduce: ") +
                                      (ctx->nccl.getErrorString ? ctx->nccl.getErrorString(nrc) : "error"));
  ctx->lastMs = 0.f;
  for (vr_ctx *ch : ctx->children) {
    ch->lastNumRays = total;  // TraceInfo.numRays of the whole job
    if (sync) {
      CK(cudaSetDevice(ch->device));
      CK(cudaStreamSynchronize(ch->stream));
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ch->ev0, ch->ev1) == cudaSuccess)
        ctx->lastMs = std::max(ctx->lastMs, ms);
    }
  }
  if (sync)
    for (vr_ctx *ch : ctx->children)
      ch->lastMs = ctx->lastMs;
  return VR_OK;
}

// The buffers one particle's wavefront loop runs on (the context's own, or its second lane's).
struct Lane {
  cudaStream_t stream = nullptr;
  RayPool pool{}, pool2{};
  unsigned long long *dCursor = nullptr;
  unsigned int *dSlotCursor = nullptr;
  unsigned long long *dCounterCopies = nullptr;
  unsigned int *hLive = nullptr;
  cudaEvent_t liveEv = nullptr;
  float4 *dSpreadQ = nullptr;
  bool instrumented = false;  // the per-launch timing marks belong to the context's stream
  int kernelLaunches = 0, iterations = 0;
  std::string err;
};
static void markLane(vr_ctx *c, const Lane &L, int kind) {
  if (L.instrumented)
    mark(c, kind);
}
#define LCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      L.err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
      return VR_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

// One particle's wavefront on one lane: pool init, the traverse / shade iterations, the
// compacting end and the tail kernel, then the particle's counters.  p comes filled
// (fillParams); the lane's buffers are bound here.
static int traceParticle(vr_ctx *ctx, Lane &L, TraceParams p, uint32_t poolSlots, int k, int np,
                         size_t n) {
  if (p.idxEnd == p.idxBegin)
    return VR_OK;
  const uint64_t shardRays = p.idxEnd - p.idxBegin;
  const uint32_t slots = (uint32_t)std::min<uint64_t>(poolSlots, shardRays);
  p.pool = L.pool;
  p.poolOut = L.pool2;
  p.rayCursor = L.dCursor;
  p.slotCursor = L.dSlotCursor;
  p.liveCount = L.dSlotCursor + 1;
  p.slotCount = L.dSlotCursor + 2;
  p.spreadCount = L.dSlotCursor + 4;
  p.spreadQ = L.dSpreadQ;
  p.counters = L.dCounterCopies;
  p.numSlots = slots;
  {
    LCK(cudaMemsetAsync(L.dCursor, 0, sizeof(unsigned long long), L.stream));
    {
      const unsigned int ctrl[8] = {0u, 0u, slots, 0u, 0u, 0u, 0u, 0u};
      LCK(cudaMemcpyAsync(L.dSlotCursor, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice,
                         L.stream));
    }
    LCK(cudaMemsetAsync(L.dCounterCopies, 0, VR_COUNTER_COPIES * 8 * sizeof(unsigned long long),
                       L.stream));
    LCK(launchInitPool(p, L.stream));
    LCK(launchFlip(L.dSlotCursor, L.dCounterCopies, 0, L.stream));
    L.kernelLaunches += 2;
    // Wavefront iterations, launched in batches with one read-back (survivor
    // count, ray cursor) per batch.  While the source still has rays every slot
    // stays alive, and an iteration hands out at most `slots` new rays, so
    // remaining / slots iterations can be queued blind.  Once a survivor count
    // below the pool size shows that the source is exhausted, the shade kernel
    // compacts the survivors into the other pool and the launches shrink.
    bool compact = false, compacted = false;
    uint32_t bound = slots;  // upper bound of the survivors (from the last read-back)
    uint64_t handed = std::min<uint64_t>(slots, shardRays);
    int cur = 0;
    for (;;) {
      int batch = 1;
      if (!compact)
        batch = (int)std::min<uint64_t>(std::max<uint64_t>((shardRays - handed) / slots, 1), 256);
      else if (compacted)
        batch = bound > 262144u ? 2 : (bound > 8192u ? 16 : 64);  // the long thin tail
      cudaEvent_t dumpEv[3] = {nullptr, nullptr, nullptr};
      if (ctx->dumpLaunches) {
        batch = 1;
        for (auto &e : dumpEv)
          cudaEventCreate(&e);
        cudaEventRecord(dumpEv[0], L.stream);
      }
      for (int b = 0; b < batch; ++b) {
        p.pool = cur ? L.pool2 : L.pool;
        p.poolOut = cur ? L.pool : L.pool2;
        p.compact = compact ? 1 : 0;
        p.numSlots = compacted ? bound : slots;  // the first compacting pass reads every slot
        markLane(ctx, L, 2);
        LCK(launchTraverse(p, ctx->numSMs, L.stream));
        markLane(ctx, L, 0);
        if (dumpEv[1])
          cudaEventRecord(dumpEv[1], L.stream);
        LCK(launchShade(p, L.stream));
        LCK(launchSpread(p, ctx->numSMs, L.stream));
        markLane(ctx, L, 1);
        LCK(launchFlip(L.dSlotCursor, L.dCounterCopies, p.compact, L.stream));
        if (b == batch - 1) {
          LCK(cudaMemcpyAsync(&L.hLive[0], L.dSlotCursor + 3, sizeof(unsigned int),
                             cudaMemcpyDeviceToHost, L.stream));
          LCK(cudaMemcpyAsync(&L.hLive[2], L.dCursor, sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, L.stream));
        }
        L.kernelLaunches += 3;
        ++L.iterations;
        if (compact) {
          cur ^= 1;
          compacted = true;
        }
      }
      if (dumpEv[2])
        cudaEventRecord(dumpEv[2], L.stream);
      LCK(cudaEventRecord(L.liveEv, L.stream));
      LCK(cudaEventSynchronize(L.liveEv));
      const uint32_t live = L.hLive[0];
      unsigned long long cursor;
      memcpy(&cursor, &L.hLive[2], sizeof(cursor));
      if (dumpEv[2]) {
        float tms = 0.f, sms = 0.f;
        cudaEventElapsedTime(&tms, dumpEv[0], dumpEv[1]);
        cudaEventElapsedTime(&sms, dumpEv[1], dumpEv[2]);
        fprintf(stderr, "[vr] particle %d iter %d slots %u compact %d live %u handed %llu traverse %.4f ms shade+flip %.4f ms\n",
                k, L.iterations, compacted ? bound : slots, (int)compact, live, cursor, tms, sms);
        for (auto &e : dumpEv)
          cudaEventDestroy(e);
      }
      handed = std::min<uint64_t>(cursor, shardRays);
      if (live == 0u)
        break;
      if (live < slots) {
        compact = true;
        bound = std::min(bound, live);
      }
      if (compacted && live <= ctx->tailRays) {
        // the thin tail: every remaining ray is finished by one thread of one launch
        p.pool = cur ? L.pool2 : L.pool;
        p.numSlots = bound;
        markLane(ctx, L, 2);
        p.spreadQ = nullptr;  // the tail kernel spreads inline
        cudaEvent_t tl[2] = {nullptr, nullptr};
        if (ctx->dumpLaunches) {
          cudaEventCreate(&tl[0]);
          cudaEventCreate(&tl[1]);
          cudaEventRecord(tl[0], L.stream);
        }
        LCK(launchTail(p, L.stream));
        if (tl[0]) {
          float ms = 0.f;
          cudaEventRecord(tl[1], L.stream);
          cudaEventSynchronize(tl[1]);
          cudaEventElapsedTime(&ms, tl[0], tl[1]);
          fprintf(stderr, "[vr] particle %d tail kernel: %u rays, %.4f ms\n", k, live, ms);
          cudaEventDestroy(tl[0]);
          cudaEventDestroy(tl[1]);
        }
        markLane(ctx, L, 1);
        L.kernelLaunches += 1;
        ++L.iterations;
        break;
      }
    }
    reduceCountersKernel<<<1, 32, 0, L.stream>>>(L.dCounterCopies,
                                                    ctx->dResult + (size_t)np * n + (size_t)k * 8);
    LCK(cudaGetLastError());
  }
  return VR_OK;
}

int vr_trace_device(vr_ctx *ctx, const vr_source_desc *src, const vr_particle_desc *particles,
                    int np, const vr_config *cfg, int sync) {
  if (!ctx)
    return VR_ERR_ARGUMENT;
  if (!ctx->children.empty())
    return traceMulti(ctx, src, particles, np, cfg, sync);
  if (!ctx->committed)
    return fail(ctx, VR_ERR_STATE, "trace: scene not committed");
  if (np < 1 || !particles)
    return fail(ctx, VR_ERR_ARGUMENT, "trace: no particle was specified");
  CK(cudaSetDevice(ctx->device));
  const size_t n = ctx->n;
  const size_t words = (size_t)np * n + (size_t)np * 8;
  if (n != ctx->resultN || np != ctx->resultNp) {  // (np, n), not their product: (2, 46) and
    freeResults(ctx);                              // (1, 100) have the same word count
    CK(cudaMalloc(&ctx->dResult, sizeof(unsigned long long) * words));
    CK(cudaMalloc(&ctx->dFluxOrig, sizeof(unsigned long long) * words));
    ctx->resultWords = words;
    ctx->resultN = n;
    ctx->resultNp = np;
  }
  ctx->resultValid = false;
  ctx->numParticles = np;
  ctx->lastNumRays = cfg ? cfg->rayIdxEnd - cfg->rayIdxBegin : 0;
  {  // sticking tables by material, one after the other
    std::vector<float> tab;
    ctx->matTabOffset.assign((size_t)np, 0);
    for (int k = 0; k < np; ++k) {
      ctx->matTabOffset[k] = tab.size();
      if (particles[k].stickingByMaterial && particles[k].numMaterials > 0)
        tab.insert(tab.end(), particles[k].stickingByMaterial,
                   particles[k].stickingByMaterial + particles[k].numMaterials);
    }
    if (!tab.empty()) {
      if (tab.size() > ctx->matTabCap) {
        cudaFree(ctx->dMatTab);
        ctx->dMatTab = nullptr;
        ctx->matTabCap = 0;
        CK(cudaMalloc(&ctx->dMatTab, sizeof(float) * tab.size()));
        ctx->matTabCap = tab.size();
      }
      // pageable source: the copy has read `tab` when the call returns
      CK(cudaMemcpyAsync(ctx->dMatTab, tab.data(), sizeof(float) * tab.size(),
                         cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
  }
  CK(cudaMemsetAsync(ctx->dResult, 0, sizeof(unsigned long long) * words, ctx->stream));
  if (ctx->countWork)
    CK(cudaMemsetAsync(ctx->dWork, 0, 8 * sizeof(unsigned long long), ctx->stream));
  ctx->kernelLaunches = 0;
  ctx->iterations = 0;
  const uint64_t shardRays = cfg ? cfg->rayIdxEnd - cfg->rayIdxBegin : 0;
  const uint32_t slots = (uint32_t)std::min<uint64_t>(ctx->poolSlots, std::max<uint64_t>(shardRays, 1));
  CK(ensurePool(ctx, slots));
  // neighbour spread in its own kernel?
  const size_t sceneBytes = (size_t)n * 32 + (size_t)n * 32;  // disk records + neighbour rows
  const bool spreadSplit = ctx->geoType == 0 && ctx->D == 3 &&
                           (ctx->spreadMode == 1 || (ctx->spreadMode < 0 && sceneBytes > ctx->l2Bytes));
  if (spreadSplit && ctx->spreadCap < slots) {
    cudaFree(ctx->dSpreadQ);
    ctx->dSpreadQ = nullptr;
    ctx->spreadCap = 0;
    CK(cudaMalloc(&ctx->dSpreadQ, sizeof(float4) * 2 * (size_t)slots));
    ctx->spreadCap = slots;
  }
  // sky map of this source side (3D only): rays that provably meet no primitive
  // are finished inside the shade kernel instead of being traversed
  if (ctx->D == 3 && ctx->skyCells > 0 && src && src->rayDir >= 0 && src->rayDir <= 2) {
    const float sign = src->posNeg < 0.f ? 1.f : -1.f;  // towards the source plane
    if (ctx->skyAxis != src->rayDir || ctx->skySign != sign || !ctx->dSky) {
      const int G = ctx->skyCells;
      if (!ctx->dSky)
        CK(cudaMallocAsync(&ctx->dSky, sizeof(float2) * G * G, ctx->stream));
      DeviceScene &s = ctx->scene;
      s.skyA = ctx->scene.firstDir;
      s.skyB = ctx->scene.secondDir;
      const float lo[2] = {ctx->geoLo[s.skyA], ctx->geoLo[s.skyB]};
      const float hi[2] = {ctx->geoHi[s.skyA], ctx->geoHi[s.skyB]};
      CK(buildSky(s, G, src->rayDir, sign, s.skyA, s.skyB, lo, hi, ctx->dSky, &ctx->skyTop,
                  ctx->stream));
      ctx->skyAxis = src->rayDir;
      ctx->skySign = sign;
      s.sky = ctx->dSky;
      s.skyN = G;
      s.skyUp = src->rayDir;
      s.skySign = sign;
      for (int k = 0; k < 2; ++k) {
        s.skyLo[k] = lo[k];
        s.skyInv[k] = hi[k] > lo[k] ? (float)G / (hi[k] - lo[k]) : 0.f;
      }
      s.skyTop = ctx->skyTop;
    }
  } else {
    ctx->scene.sky = nullptr;
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  // The jobs: one per particle; with fewer particles than lanes a particle's ray range is cut
  // into pieces (uneven, so that the pieces do not end together).  One after the other on the
  // context's stream, or ctx->lanes at a time on the extra lanes (vr_ctx::xlane).
  const bool serial = ctx->lanes <= 1 || ctx->timeKernels || ctx->dumpLaunches;
  std::vector<TraceParams> params;
  std::vector<int> jobParticle;
  for (int k = 0; k < np; ++k) {
    TraceParams p;
    int rc = fillParams(ctx, src, &particles[k], cfg, k, p);
    if (rc)
      return rc;
    p.flux = ctx->dResult + (size_t)k * n;
    if (p.idxEnd == p.idxBegin)
      continue;
    const uint64_t R = p.idxEnd - p.idxBegin;
    int pieces = serial ? 1 : std::max(1, (ctx->lanes + np - 1) / np);
    // short jobs stay whole, every piece pays its own start and end: measured on one B200,
    // 4e8 rays of the 4M-disk hole array in two pieces +3 %, 1e8 rays of the trench -6 %
    while (pieces > 1 && R / (uint64_t)pieces < 8ull * ctx->poolSlots)
      --pieces;
    uint64_t b0 = p.idxBegin;
    for (int j = 0; j < pieces; ++j) {
      // shares 1 + 0.2 (pieces - 1 - j) of the mean: the first piece is the longest
      const double w = (1.0 + 0.2 * (pieces - 1 - 2 * j) / 2.0) / pieces;
      uint64_t e0 = j == pieces - 1 ? p.idxEnd : std::min<uint64_t>(p.idxEnd, b0 + (uint64_t)(w * (double)R));
      TraceParams q = p;
      q.idxBegin = b0;
      q.idxEnd = e0;
      b0 = e0;
      if (q.idxEnd > q.idxBegin) {
        params.push_back(q);
        jobParticle.push_back(k);
      }
    }
  }
  const int jobs = (int)params.size();
  Lane lane0;
  lane0.stream = ctx->stream;
  lane0.pool = ctx->pool;
  lane0.pool2 = ctx->pool2;
  lane0.dCursor = ctx->dCursor;
  lane0.dSlotCursor = ctx->dSlotCursor;
  lane0.dCounterCopies = ctx->dCounterCopies;
  lane0.hLive = ctx->hLive;
  lane0.liveEv = ctx->liveEv[0];
  lane0.dSpreadQ = spreadSplit ? ctx->dSpreadQ : nullptr;
  lane0.instrumented = true;
  const int numLanes = serial ? 1 : std::min(ctx->lanes, jobs);
  if (numLanes <= 1) {
    for (int j = 0; j < jobs; ++j) {
      int rc = traceParticle(ctx, lane0, params[j], ctx->poolSlots, jobParticle[j], np, n);
      if (rc)
        return fail(ctx, rc, lane0.err);
    }
  } else {
    std::vector<Lane> lanes((size_t)numLanes);
    lanes[0] = lane0;
    for (int li = 1; li < numLanes; ++li) {
      CK(ensureLane(ctx, li - 1, slots, spreadSplit));
      vr_ctx::Lane1 &x = ctx->xlane[li - 1];
      Lane &l = lanes[li];
      l.stream = x.stream;
      l.pool = x.pool;
      l.pool2 = x.pool2;
      l.dCursor = x.dCursor;
      l.dSlotCursor = x.dSlotCursor;
      l.dCounterCopies = x.dCounterCopies;
      l.hLive = x.hLive;
      l.liveEv = x.liveEv;
      l.dSpreadQ = spreadSplit ? x.dSpreadQ : nullptr;
      CK(cudaStreamWaitEvent(l.stream, ctx->ev0, 0));  // the cleared result words
    }
    std::atomic<int> next{0};
    std::vector<int> rcs((size_t)numLanes, VR_OK);
    auto worker = [&](int li) {
      cudaSetDevice(ctx->device);
      for (;;) {
        const int j = next.fetch_add(1);
        if (j >= jobs)
          break;
        rcs[li] = traceParticle(ctx, lanes[li], params[j], ctx->poolSlots, jobParticle[j], np, n);
        if (rcs[li] != VR_OK)
          break;
      }
    };
    std::vector<std::thread> threads;
    for (int li = 1; li < numLanes; ++li)
      threads.emplace_back(worker, li);
    worker(0);
    for (auto &t : threads)
      t.join();
    for (int li = 0; li < numLanes; ++li)
      if (rcs[li] != VR_OK)
        return fail(ctx, rcs[li], lanes[li].err);
    for (int li = 1; li < numLanes; ++li) {
      CK(cudaEventRecord(ctx->xlane[li - 1].done, lanes[li].stream));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->xlane[li - 1].done, 0));
      lanes[0].kernelLaunches += lanes[li].kernelLaunches;
      lanes[0].iterations += lanes[li].iterations;
    }
    lane0.kernelLaunches = lanes[0].kernelLaunches;
    lane0.iterations = lanes[0].iterations;
  }
  ctx->kernelLaunches = lane0.kernelLaunches;
  ctx->iterations = lane0.iterations;
  // the result in the caller's primitive order (+ the counters behind it)
  for (int k = 0; k < np; ++k)
    CK(launchUnsortFlux(ctx->dResult + (size_t)k * n, ctx->bvh.sortedToOrig, (uint32_t)n,
                        ctx->dFluxOrig + (size_t)k * n, ctx->stream));
  CK(cudaMemcpyAsync(ctx->dFluxOrig + (size_t)np * n, ctx->dResult + (size_t)np * n,
                     sizeof(unsigned long long) * (size_t)np * 8, cudaMemcpyDeviceToDevice,
                     ctx->stream));
  ctx->resultValid = true;
  mark(ctx, 2);
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (sync) {
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(&ctx->lastMs, ctx->ev0, ctx->ev1));
  }
  if (ctx->timeKernels && sync) {
    collectMarks(ctx);
    if (getenv("VR_TIME_KERNELS"))
      fprintf(stderr, "[vr] phases (accumulated): traverse %.3f ms, shade %.3f ms, other %.3f ms\n",
              ctx->phaseMs[0], ctx->phaseMs[1], ctx->phaseMs[2]);
  }
  return VR_OK;
}

int vr_flux_device(vr_ctx *ctx, void **devicePtr, size_t *numWords) {
  if (ctx && !ctx->children.empty())
    return vr_flux_device(ctx->children[0], devicePtr, numWords);
  if (!ctx || !devicePtr || !numWords)
    return VR_ERR_ARGUMENT;
  if (!ctx->dFluxOrig || !ctx->resultValid)
    return fail(ctx, VR_ERR_STATE, "vr_flux_device: no trace has run on the current scene");
  *devicePtr = ctx->dFluxOrig;
  *numWords = ctx->resultWords;
  return VR_OK;
}

static int downloadFixed(vr_ctx *ctx, std::vector<unsigned long long> &flux,
                         std::vector<unsigned long long> &counters) {
  if (!ctx->dFluxOrig || !ctx->resultValid)
    return fail(ctx, VR_ERR_STATE, "download: no trace has run on the current scene");
  CK(cudaSetDevice(ctx->device));
  const size_t n = ctx->resultN, np = ctx->resultNp;
  flux.resize(np * n);
  counters.resize(np * 8);
  CK(cudaMemcpyAsync(flux.data(), ctx->dFluxOrig, sizeof(unsigned long long) * np * n,
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(counters.data(), ctx->dFluxOrig + np * n,
                     sizeof(unsigned long long) * np * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (cudaEventQuery(ctx->ev1) == cudaSuccess)
    cudaEventElapsedTime(&ctx->lastMs, ctx->ev0, ctx->ev1);
  return VR_OK;
}

int vr_flux_download(vr_ctx *ctx, double *fluxOut, vr_trace_info *infoOut) {
  if (ctx && !ctx->children.empty())
    {
      int rc_ = vr_flux_download(ctx->children[0], fluxOut, infoOut);
      return rc_ ? childFail(ctx, ctx->children[0], rc_) : VR_OK;
    };
  if (!ctx)
    return VR_ERR_ARGUMENT;
  std::vector<unsigned long long> flux, counters;
  int rc = downloadFixed(ctx, flux, counters);
  if (rc)
    return rc;
  if (fluxOut)
    for (size_t i = 0; i < flux.size(); ++i)
      fluxOut[i] = (double)flux[i] * (1.0 / VR_FLUX_FIXED_SCALE);
  if (infoOut)
    for (int k = 0; k < ctx->numParticles; ++k) {
      const unsigned long long *c = &counters[(size_t)k * 8];
      vr_trace_info &ti = infoOut[k];
      ti.numRays = ctx->lastNumRays;
      ti.totalRaysTraced = c[1];
      ti.nonGeometryHits = c[2];
      ti.geometryHits = c[3];
      ti.particleHits = c[4];
      ti.boundaryHits = c[5];
      ti.reflections = c[6];
      ti.raysTerminated = c[7];
      ti.time = (double)ctx->lastMs * 1e-3;
    }
  return VR_OK;
}

int vr_flux_postprocess(vr_ctx *ctx, int particle, const float *areas, float normFactor,
                        int smooth, float *fluxOut) {
  if (ctx && !ctx->children.empty())
    {
      int rc_ = vr_flux_postprocess(ctx->children[0], particle, areas, normFactor, smooth, fluxOut);
      return rc_ ? childFail(ctx, ctx->children[0], rc_) : VR_OK;
    };
  if (!ctx || !fluxOut)
    return VR_ERR_ARGUMENT;
  if (!ctx->dFluxOrig || !ctx->resultValid || !ctx->committed)
    return fail(ctx, VR_ERR_STATE, "vr_flux_postprocess: no trace has run on the current scene");
  if (particle < 0 || particle >= ctx->numParticles)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_flux_postprocess: particle index out of range");
  CK(cudaSetDevice(ctx->device));
  const size_t n = ctx->n;
  float *buf = nullptr;  // areas | tmpA | tmpB | out
  CK(cudaMallocAsync(&buf, sizeof(float) * 4 * n, ctx->stream));
  cudaError_t e = cudaSuccess;
  if (areas)
    e = stagedCopy(ctx, buf, areas, sizeof(float) * n);
  if (e == cudaSuccess)
    e = postprocessFlux(ctx->scene, ctx->dFluxOrig + (size_t)particle * n, ctx->bvh.sortedToOrig,
                        areas ? buf : nullptr, normFactor, smooth, buf + n, buf + 2 * n,
                        buf + 3 * n, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(fluxOut, buf + 3 * n, sizeof(float) * n, cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  cudaFreeAsync(buf, ctx->stream);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_flux_postprocess");
  return VR_OK;
}

int vr_flux_postprocess_ex(vr_ctx *ctx, int particle, const float *areas, int normalization,
                           double normFactor, int smoothNeighbors, float diskRadius,
                           float *fluxOut) {
  if (ctx && !ctx->children.empty()) {
    int rc_ = vr_flux_postprocess_ex(ctx->children[0], particle, areas, normalization, normFactor,
                                     smoothNeighbors, diskRadius, fluxOut);
    return rc_ ? childFail(ctx, ctx->children[0], rc_) : VR_OK;
  }
  if (!ctx || !fluxOut)
    return VR_ERR_ARGUMENT;
  if (!ctx->dFluxOrig || !ctx->resultValid || !ctx->committed)
    return fail(ctx, VR_ERR_STATE, "vr_flux_postprocess_ex: no trace has run on the current scene");
  if (particle < 0 || particle >= ctx->numParticles)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_flux_postprocess_ex: particle index out of range");
  if (normalization < VR_NORM_NONE || normalization > VR_NORM_MAX ||
      (normalization != VR_NORM_NONE && !areas))
    return fail(ctx, VR_ERR_ARGUMENT, "vr_flux_postprocess_ex: normalisation needs the areas");
  if (smoothNeighbors < 0)
    return fail(ctx, VR_ERR_ARGUMENT, "vr_flux_postprocess_ex: negative smoothing width");
  const bool smooth = smoothNeighbors > 0 && ctx->geoType == 0;  // triangles: no smoothing
  if (smooth && smoothNeighbors == 1 && !ctx->dNbOffO)
    return fail(ctx, VR_ERR_STATE, "vr_flux_postprocess_ex: the scene has no neighbour lists");
  if (smooth && smoothNeighbors > 1 && !(diskRadius > 0.f))
    return fail(ctx, VR_ERR_ARGUMENT, "vr_flux_postprocess_ex: wider smoothing needs the disk radius");
  CK(cudaSetDevice(ctx->device));
  const uint32_t n = ctx->n;
  float *buf = nullptr, *dPts = nullptr;  // buf: areas | tmp | out | maximum
  uint32_t *dOff = nullptr, *dIdx = nullptr;
  cudaError_t e = cudaMallocAsync(&buf, sizeof(float) * (3 * (size_t)n + 4), ctx->stream);
  const uint32_t *off = nullptr, *idx = nullptr;
  if (e == cudaSuccess && areas)
    e = stagedCopy(ctx, buf, areas, sizeof(float) * n);
  if (e == cudaSuccess && smooth) {
    if (smoothNeighbors == 1) {
      off = ctx->dNbOffO;
      idx = ctx->dNbIdxO;
    } else {
      // a new neighbourhood of numNeighbors * 2 * radius over the points (rayTraceDisk.hpp:
      // 160-168, init<3>), built on the device
      size_t total = 0;
      e = cudaMallocAsync(&dPts, sizeof(float) * 3 * (size_t)n, ctx->stream);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(dPts, 3 * sizeof(float), ctx->dXyzr, 4 * sizeof(float),
                              3 * sizeof(float), n, cudaMemcpyDeviceToDevice, ctx->stream);
      if (e == cudaSuccess)
        e = cudaMallocAsync(&dOff, sizeof(uint32_t) * ((size_t)n + 1), ctx->stream);
      if (e == cudaSuccess)
        e = buildNeighborsDevice(3, dPts, n, ctx->geoLo, (float)smoothNeighbors * 2 * diskRadius,
                                 dOff, &dIdx, &total, ctx->stream);
      off = dOff;
      idx = dIdx;
    }
  }
  // MAX on triangles divides by max * area (rayTraceTriangle.hpp:99-107)
  const int mode = normalization == VR_NORM_MAX ? (ctx->geoType == 0 ? 2 : 3) : normalization;
  if (e == cudaSuccess)
    e = postprocessFluxEx(ctx->dFluxOrig + (size_t)particle * n, n, areas ? buf : nullptr, mode,
                          normFactor, ctx->dNxyz, off, idx, buf + n, buf + 2 * (size_t)n,
                          reinterpret_cast<unsigned int *>(buf + 3 * (size_t)n), ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(fluxOut, buf + 2 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  cudaFreeAsync(buf, ctx->stream);
  cudaFreeAsync(dPts, ctx->stream);
  cudaFreeAsync(dOff, ctx->stream);
  cudaFreeAsync(dIdx, ctx->stream);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_flux_postprocess_ex");
  return VR_OK;
}

int vr_flux_download_fixed(vr_ctx *ctx, uint64_t *fluxOut) {
  if (ctx && !ctx->children.empty())
    {
      int rc_ = vr_flux_download_fixed(ctx->children[0], fluxOut);
      return rc_ ? childFail(ctx, ctx->children[0], rc_) : VR_OK;
    };
  if (!ctx || !fluxOut)
    return VR_ERR_ARGUMENT;
  std::vector<unsigned long long> flux, counters;
  int rc = downloadFixed(ctx, flux, counters);
  if (rc)
    return rc;
  memcpy(fluxOut, flux.data(), sizeof(uint64_t) * flux.size());
  return VR_OK;
}

int vr_trace(vr_ctx *ctx, const vr_source_desc *src, const vr_particle_desc *particles, int np,
             const vr_config *cfg, double *fluxOut, vr_trace_info *infoOut) {
  int rc = vr_trace_device(ctx, src, particles, np, cfg, 1);
  if (rc)
    return rc;
  return vr_flux_download(ctx, fluxOut, infoOut);
}

// ---- host-side geometry helper ---------------------------------------------
int vr_build_neighbors(int D, const float *pts, uint32_t n, float dist, uint32_t **offOut,
                       uint32_t **idxOut) {
  if (!pts || !offOut || !idxOut || (D != 2 && D != 3))
    return VR_ERR_ARGUMENT;
  uint32_t *off = (uint32_t *)calloc((size_t)n + 1, sizeof(uint32_t));
  *offOut = off;
  *idxOut = nullptr;
  if (n == 0 || !(dist > 0.f)) {
    *idxOut = (uint32_t *)malloc(sizeof(uint32_t));
    return VR_OK;
  }
  // uniform grid of cell size >= dist over the first D axes, points sorted by cell
  float lo[3] = {INFINITY, INFINITY, INFINITY};
  for (uint32_t i = 0; i < n; ++i)
    for (int a = 0; a < D; ++a)
      lo[a] = std::min(lo[a], pts[3 * i + a]);
  const float cell = dist * 1.0001f;
  auto cellOf = [&](const float *p, int a) -> int64_t {
    return a < D ? (int64_t)std::floor((p[a] - lo[a]) / cell) + 1 : 1;
  };
  auto keyOf = [](int64_t x, int64_t y, int64_t z) -> uint64_t {
    return ((uint64_t)x & 0x1fffff) | (((uint64_t)y & 0x1fffff) << 21) |
           (((uint64_t)z & 0x1fffff) << 42);
  };
  std::vector<std::pair<uint64_t, uint32_t>> refs(n);
  for (uint32_t i = 0; i < n; ++i) {
    const float *p = pts + 3 * i;
    refs[i] = {keyOf(cellOf(p, 0), cellOf(p, 1), cellOf(p, 2)), i};
  }
  std::sort(refs.begin(), refs.end());
  const float dist2 = dist * dist;
  auto isNb = [&](const float *p, const float *q) {
    for (int a = 0; a < D; ++a)
      if (std::fabs(p[a] - q[a]) > dist)
        return false;
    float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return (dx * dx + dy * dy) + dz * dz <= dist2;
  };
  std::vector<std::vector<uint32_t>> rows(n);
  const int zr = D == 3 ? 1 : 0;
  for (uint32_t i = 0; i < n; ++i) {
    const float *p = pts + 3 * i;
    int64_t cx = cellOf(p, 0), cy = cellOf(p, 1), cz = cellOf(p, 2);
    for (int dz = -zr; dz <= zr; ++dz)
      for (int dy = -1; dy <= 1; ++dy) {
        // the three x-cells are consecutive keys: one range scan
        uint64_t k0 = keyOf(cx - 1, cy + dy, cz + dz), k1 = keyOf(cx + 1, cy + dy, cz + dz);
        auto it = std::lower_bound(refs.begin(), refs.end(), std::make_pair(k0, 0u));
        for (; it != refs.end() && it->first <= k1; ++it) {
          uint32_t j = it->second;
          if (j != i && isNb(p, pts + 3 * j))
            rows[i].push_back(j);
        }
      }
    std::sort(rows[i].begin(), rows[i].end());
  }
  for (uint32_t i = 0; i < n; ++i)
    off[i + 1] = off[i] + (uint32_t)rows[i].size();
  uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * std::max<size_t>(off[n], 1));
  for (uint32_t i = 0; i < n; ++i)
    std::copy(rows[i].begin(), rows[i].end(), idx + off[i]);
  *idxOut = idx;
  return VR_OK;
}

void vr_free(void *p) { free(p); }

// ---- parity / debugging ------------------------------------------------------
int vr_debug_intersect(vr_ctx *ctx, const float *rays, uint32_t m, uint32_t *geomOut,
                       uint32_t *primOut, float *tOut, uint32_t nbCap, uint32_t *nbCountOut,
                       uint32_t *nbOut) {
  if (ctx && !ctx->children.empty())
    return vr_debug_intersect(ctx->children[0], rays, m, geomOut, primOut, tOut, nbCap, nbCountOut, nbOut);
  if (!ctx || !rays || !geomOut || !primOut || !tOut)
    return VR_ERR_ARGUMENT;
  if (!ctx->committed)
    return fail(ctx, VR_ERR_STATE, "vr_debug_intersect: scene not committed");
  CK(cudaSetDevice(ctx->device));
  float *dRays = nullptr, *dT = nullptr;
  uint32_t *dGeom = nullptr, *dPrim = nullptr, *dCnt = nullptr, *dNb = nullptr;
  auto cleanup = [&]() {
    cudaFree(dRays);
    cudaFree(dT);
    cudaFree(dGeom);
    cudaFree(dPrim);
    cudaFree(dCnt);
    cudaFree(dNb);
  };
#define CKD(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      cleanup();                                                                                   \
      return failCuda(ctx, e_, #call);                                                             \
    }                                                                                              \
  } while (0)
  const bool wantNb = nbCountOut && nbOut && nbCap;
  CKD(cudaMalloc(&dRays, sizeof(float) * 6 * (size_t)m + 16));
  CKD(cudaMalloc(&dT, sizeof(float) * m + 16));
  CKD(cudaMalloc(&dGeom, sizeof(uint32_t) * m + 16));
  CKD(cudaMalloc(&dPrim, sizeof(uint32_t) * m + 16));
  if (wantNb) {
    CKD(cudaMalloc(&dCnt, sizeof(uint32_t) * m + 16));
    CKD(cudaMalloc(&dNb, sizeof(uint32_t) * (size_t)m * nbCap + 16));
    CKD(cudaMemsetAsync(dNb, 0xff, sizeof(uint32_t) * (size_t)m * nbCap, ctx->stream));
  }
  CKD(cudaMemcpyAsync(dRays, rays, sizeof(float) * 6 * (size_t)m, cudaMemcpyHostToDevice,
                      ctx->stream));
  CKD(ensurePool(ctx, std::max<uint32_t>(m, 1u)));
  {
    TraceParams p{};
    p.scene = ctx->scene;
    p.pool = ctx->pool;
    p.numSlots = m;
    p.slotCursor = ctx->dSlotCursor;
    p.liveCount = ctx->dSlotCursor + 1;
    p.slotCount = ctx->dSlotCursor + 2;
    p.work = nullptr;
    p.spreadQ = nullptr;
    {
      const unsigned int ctrl[4] = {0u, 0u, m, 0u};
      CKD(cudaMemcpyAsync(ctx->dSlotCursor, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice,
                          ctx->stream));
    }
    CKD(launchDebugLoadRays(ctx->scene, ctx->pool, dRays, m, ctx->stream));
    CKD(cudaEventRecord(ctx->ev0, ctx->stream));
    CKD(launchTraverse(p, ctx->numSMs, ctx->stream));
    CKD(cudaEventRecord(ctx->ev1, ctx->stream));
    CKD(launchDebugReadHits(ctx->scene, ctx->pool, m, dGeom, dPrim, dT, wantNb ? nbCap : 0u,
                            wantNb ? dCnt : nullptr, dNb, ctx->bvh.sortedToOrig, ctx->stream));
  }
  CKD(cudaMemcpyAsync(geomOut, dGeom, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
  CKD(cudaMemcpyAsync(primOut, dPrim, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
  CKD(cudaMemcpyAsync(tOut, dT, sizeof(float) * m, cudaMemcpyDeviceToHost, ctx->stream));
  if (wantNb) {
    CKD(cudaMemcpyAsync(nbCountOut, dCnt, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost,
                        ctx->stream));
    CKD(cudaMemcpyAsync(nbOut, dNb, sizeof(uint32_t) * (size_t)m * nbCap, cudaMemcpyDeviceToHost,
                        ctx->stream));
  }
  CKD(cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->lastMs, ctx->ev0, ctx->ev1);  // traverse kernel alone
  cleanup();
  return VR_OK;
}

int vr_debug_source_rays(vr_ctx *ctx, const vr_source_desc *src, const vr_particle_desc *part,
                         const vr_config *cfg, uint64_t idxBegin, uint32_t m, float *raysOut) {
  if (ctx && !ctx->children.empty())
    return vr_debug_source_rays(ctx->children[0], src, part, cfg, idxBegin, m, raysOut);
  if (!ctx || !raysOut)
    return VR_ERR_ARGUMENT;
  if (!ctx->boundarySet)
    return fail(ctx, VR_ERR_STATE, "vr_debug_source_rays: no boundary was set");
  CK(cudaSetDevice(ctx->device));
  TraceParams p;
  int rc = fillParams(ctx, src, part, cfg, 0, p);
  if (rc)
    return rc;
  float *d = nullptr;
  CK(cudaMalloc(&d, sizeof(float) * 6 * (size_t)m + 16));
  cudaError_t e = launchDebugSourceRays(p, idxBegin, m, d, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(raysOut, d, sizeof(float) * 6 * (size_t)m, cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_debug_source_rays");
  return VR_OK;
}

int vr_debug_math(vr_ctx *ctx, int which, const float *x, uint32_t m, float param, float *out) {
  if (ctx && !ctx->children.empty())
    return vr_debug_math(ctx->children[0], which, x, m, param, out);
  if (!ctx || !x || !out || which < 0 || which > 2)
    return VR_ERR_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  const size_t outN = which == 0 ? 2 * (size_t)m : m;
  float *dx = nullptr, *dout = nullptr;
  CK(cudaMalloc(&dx, sizeof(float) * m + 16));
  cudaError_t e = cudaMalloc(&dout, sizeof(float) * outN + 16);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(dx, x, sizeof(float) * m, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = launchDebugMath(which, dx, m, param, dout, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out, dout, sizeof(float) * outN, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dx);
  cudaFree(dout);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_debug_math");
  return VR_OK;
}

int vr_debug_philox(vr_ctx *ctx, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint32_t *out4) {
  if (ctx && !ctx->children.empty())
    return vr_debug_philox(ctx->children[0], k0, k1, c0, c1, c2, c3, out4);
  if (!ctx || !out4)
    return VR_ERR_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  uint32_t *d = nullptr;
  CK(cudaMalloc(&d, 16));
  cudaError_t e = launchDebugPhilox(k0, k1, c0, c1, c2, c3, d, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out4, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_debug_philox");
  return VR_OK;
}

int vr_debug_reflect(vr_ctx *ctx, int kind, int D, const float *rayDir, const float *normal,
                     float coneMinAngle, uint32_t seed, uint64_t idx, uint32_t m, float *out3) {
  if (ctx && !ctx->children.empty())
    return vr_debug_reflect(ctx->children[0], kind, D, rayDir, normal, coneMinAngle, seed, idx, m, out3);
  if (!ctx || !rayDir || !normal || !out3 || (D != 2 && D != 3) || kind < 0 || kind > 2)
    return VR_ERR_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  float *d = nullptr;
  CK(cudaMalloc(&d, sizeof(float) * 3 * (size_t)m + 16));
  cudaError_t e =
      launchDebugReflect(kind, D, rayDir, normal, coneMinAngle, seed, idx, m, d, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out3, d, sizeof(float) * 3 * (size_t)m, cudaMemcpyDeviceToHost,
                        ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess)
    return failCuda(ctx, e, "vr_debug_reflect");
  return VR_OK;
}

int vr_debug_bvh_stats(vr_ctx *ctx, uint64_t *out5) {
  if (ctx && !ctx->children.empty())
    return vr_debug_bvh_stats(ctx->children[0], out5);
  if (!ctx || !out5)
    return VR_ERR_ARGUMENT;
  if (!ctx->committed)
    return fail(ctx, VR_ERR_STATE, "vr_debug_bvh_stats: scene not committed");
  out5[0] = ctx->bvh.numNodes;
  out5[1] = ctx->bvh.numLeaves;
  out5[2] = ctx->bvh.maxLeaf;
  out5[3] = sizeof(Node2);
  uint32_t bits;
  memcpy(&bits, &ctx->bvh.buildMs, 4);
  out5[4] = bits;
  memcpy(&bits, &ctx->bvh.sahInner, 4);
  out5[5] = bits;
  memcpy(&bits, &ctx->bvh.sahLeaf, 4);
  out5[6] = bits;
  memcpy(&bits, &ctx->bvh.mortonAlpha, 4);
  out5[7] = bits;
  return VR_OK;
}

int vr_debug_phase_timing(vr_ctx *ctx, int enable) {
  if (ctx && !ctx->children.empty())
    EACH_CHILD(vr_debug_phase_timing(ch, enable));
  if (!ctx)
    return VR_ERR_ARGUMENT;
  ctx->timeKernels = enable != 0;
  for (int k = 0; k < 3; ++k) {
    ctx->phaseMs[k] = 0;
    ctx->phaseLaunches[k] = 0;
  }
  return VR_OK;
}

int vr_debug_phase_ms(vr_ctx *ctx, double *ms3, int64_t *launches3) {
  if (ctx && !ctx->children.empty())
    return vr_debug_phase_ms(ctx->children[0], ms3, launches3);
  if (!ctx || !ms3)
    return VR_ERR_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  collectMarks(ctx);
  for (int k = 0; k < 3; ++k) {
    ms3[k] = ctx->phaseMs[k];
    if (launches3)
      launches3[k] = ctx->phaseLaunches[k];
  }
  return VR_OK;
}

int vr_debug_work_counters(vr_ctx *ctx, uint64_t *out5) {
  if (ctx && !ctx->children.empty())
    return vr_debug_work_counters(ctx->children[0], out5);
  if (!ctx || !out5)
    return VR_ERR_ARGUMENT;
  if (!ctx->countWork)
    return fail(ctx, VR_ERR_STATE, "vr_debug_work_counters: set VR_COUNT_WORK=1 before create");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out5, ctx->dWork, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return VR_OK;
}

int vr_debug_l2_read_bandwidth(vr_ctx *ctx, uint64_t bytes, int passes, double *out2) {
  if (ctx && !ctx->children.empty())
    return vr_debug_l2_read_bandwidth(ctx->children[0], bytes, passes, out2);
  if (!ctx || !out2 || bytes < 16 || passes < 1)
    return VR_ERR_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  double gbps = 0.;
  CK(vr::l2ReadBandwidth((size_t)bytes, passes, ctx->numSMs, ctx->stream, &gbps));
  out2[0] = gbps;
  out2[1] = (double)prop.l2CacheSize;
  return VR_OK;
}

}  // extern "C"
