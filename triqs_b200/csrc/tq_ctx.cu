// Context, error plumbing, staging buffers, memory helpers and the NCCL communicator of libtqgpu.
#include "tq_common.cuh"
#include "tq_fft.cuh"

#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

tq_status tq_fail(tq_ctx *ctx, tq_status st, const std::string &msg) {
  if (ctx) ctx->err = msg;
  return st;
}

extern "C" const char *tq_version(void) { return "tqgpu 0.1 sm_100a"; }

extern "C" tq_status tq_create(int device, tq_ctx **out) {
  if (!out) return TQ_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) return TQ_ERR_NO_DEVICE; // never compute on the CPU
  if (device < 0 || device >= n) return TQ_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return TQ_ERR_CUDA;
  auto *ctx = new tq_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return TQ_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  *out = ctx;
  return TQ_OK;
}

static void free_buf(DeviceBuffer &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

extern "C" tq_status tq_comm_destroy_internal(tq_ctx *ctx);

extern "C" void tq_destroy(tq_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (tq_ctx *c : ctx->lanes) tq_destroy(c);
  ctx->lanes.clear();
  tq_comm_destroy_internal(ctx);
  ctx->fft_plans.clear();
  ctx->matsubara_plans.clear();
  ctx->tail_plans.clear();
  free_buf(ctx->stage_in);
  free_buf(ctx->stage_in2);
  free_buf(ctx->stage_in3);
  free_buf(ctx->stage_out);
  free_buf(ctx->scratch);
  free_buf(ctx->small);
  free_buf(ctx->pipe_w);
  free_buf(ctx->leg_tmp);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < 4; ++i) {
    if (ctx->bounce[i]) cudaFreeHost(ctx->bounce[i]);
    if (ctx->bounce_ev[i]) cudaEventDestroy(ctx->bounce_ev[i]);
  }
  if (ctx->order_event) cudaEventDestroy(ctx->order_event);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *tq_last_error(const tq_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void *tq_stream(tq_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" uint64_t tq_launch_count(const tq_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" tq_status tq_profile_enable(tq_ctx *ctx, int on) {
  if (!ctx) return TQ_ERR_INVALID;
  ctx->profiling = on != 0;
  return TQ_OK;
}

extern "C" tq_status tq_profile_report(tq_ctx *ctx, char *buf, size_t len) {
  if (!ctx || !buf || len < 3) return TQ_ERR_INVALID;
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::map<std::string, std::pair<long, double>> agg;
  for (auto &r : ctx->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto &a = agg[r.tag];
    a.first += 1;
    a.second += ms;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->prof.clear();
  std::string out = "{";
  bool first = true;
  for (auto &kv : agg) {
    char line[256];
    snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "}";
  strncpy(buf, out.c_str(), len - 1);
  buf[len - 1] = 0;
  return TQ_OK;
}

extern "C" tq_status tq_set_option(tq_ctx *ctx, const char *name, long value) {
  if (!ctx || !name) return TQ_ERR_INVALID;
  const std::string n(name);
  if (n == "no_pipe")
    ctx->opt_no_pipe = value;
  else if (n == "pipe_lag0")
    ctx->opt_pipe_lag0 = value;
  else if (n == "force_planned")
    ctx->opt_force_planned = value;
  else if (n == "no_col")
    ctx->opt_no_col = value;
  else if (n == "no_tr")
    ctx->opt_no_tr = value;
  else if (n == "no_two_for_one")
    ctx->opt_no_two_for_one = value;
  else if (n == "no_bluestein")
    ctx->opt_no_bluestein = value;
  else if (n == "force_bluestein")
    ctx->opt_force_bluestein = value;
  else if (n == "blue_chunk_mb")
    ctx->opt_blue_chunk_mb = std::max<long>(1, value);
  else if (n == "scratch_chunk_mb")
    ctx->opt_scratch_chunk_mb = value;
  else if (n == "bounce")
    ctx->opt_bounce = value;
  else if (n == "host_chunk_kb")
    ctx->opt_host_chunk_kb = std::max<long>(1, value);
  else if (n == "stage_nt")
    ctx->opt_stage_nt = value;
  else if (n == "stage_threads")
    ctx->opt_stage_threads = std::max<long>(0, value);
  else if (n == "host_lanes")
    ctx->opt_host_lanes = std::max<long>(1, std::min<long>(value, 8));
  else
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_set_option: unknown option '" + n + "'");
  return TQ_OK;
}

extern "C" tq_status tq_synchronize(tq_ctx *ctx) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TQ_OK;
}

// cross-stream ordering without a host synchronisation (include/tqgpu.h "Stream-ordering contract")
static tq_status order_streams(tq_ctx *ctx, cudaStream_t first, cudaStream_t then) {
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->order_event) TQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->order_event, cudaEventDisableTiming));
  TQ_CUDA(ctx, cudaEventRecord(ctx->order_event, first));
  TQ_CUDA(ctx, cudaStreamWaitEvent(then, ctx->order_event, 0));
  return TQ_OK;
}
extern "C" tq_status tq_stream_wait_external(tq_ctx *ctx, void *other) {
  if (!ctx) return TQ_ERR_INVALID;
  return order_streams(ctx, (cudaStream_t)other, ctx->stream);
}
extern "C" tq_status tq_stream_signal_external(tq_ctx *ctx, void *other) {
  if (!ctx) return TQ_ERR_INVALID;
  return order_streams(ctx, ctx->stream, (cudaStream_t)other);
}

tq_status tq_reserve(tq_ctx *ctx, DeviceBuffer &b, size_t bytes) {
  if (bytes <= b.cap) return TQ_OK;
  if (b.p) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TQ_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t cap = bytes + bytes / 8 + 256;
  TQ_CUDA(ctx, cudaMalloc(&b.p, cap));
  b.cap = cap;
  return TQ_OK;
}

tq_status tq_reserve_pinned(tq_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap) return TQ_OK;
  if (ctx->pinned) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TQ_CUDA(ctx, cudaFreeHost(ctx->pinned));
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
  }
  TQ_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes + 256));
  ctx->pinned_cap = bytes + 256;
  return TQ_OK;
}

bool tq_is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Pageable user memory (a TRIQS nda array, a NumPy array): a cudaMemcpy from / to it is staged by the driver through ONE internal
// buffer per process at ~7 GB/s, whatever the number of threads (profiles/r02_e2e_lanes_probe_before_bounce.jsonl).  The library
// stages such buffers itself: the calling thread copies 8 MiB pieces into its own pinned slots and issues asynchronous DMA copies from
// there, so that the CPU copy of piece i + 1 overlaps the DMA of piece i, and different lanes (threads) proceed independently.
constexpr size_t TQ_BOUNCE_BYTES = 8u << 20;
static bool host_is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
static tq_status bounce_init(tq_ctx *ctx) {
  for (int i = 0; i < 4; ++i) {
    if (!ctx->bounce[i]) TQ_CUDA(ctx, cudaMallocHost(&ctx->bounce[i], TQ_BOUNCE_BYTES));
    if (!ctx->bounce_ev[i]) TQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bounce_ev[i], cudaEventDisableTiming));
  }
  return TQ_OK;
}

// Large pageable buffers of a call that is NOT cut into lanes (one BlockGf, a lattice gf: too few chunks to overlap anything): the copy itself
// is spread over a few helper threads, each staging its share of the pieces through the pinned slots of a helper context (the contexts the
// lanes use, created on first use) on that context's stream.  One thread moves ~5-8 GB/s through its slots; eight reach the PCIe rate.
// Measured for ONE BlockGf of configs[1] in NumPy arrays (80 MB in, 16 MB out): 5.0 -> ~2 ms per forward call.
// Copy with non-temporal stores (x86-64), used on the way IN: a staging piece is written once and read next by the DMA engine, so pulling the
// destination lines into the cache first (read-for-ownership) only costs memory bandwidth (measured: 3-4 % on the forward call; on the way
// OUT, into the caller's buffer, the same stores are 15 % slower than memcpy and are not used).
static void stage_copy(void *dst, const void *src, size_t n, bool nt) {
#if defined(__x86_64__)
  if (nt && n >= 4096 && (((uintptr_t)dst) & 15) == 0) {
    char *d = (char *)dst;
    const char *s = (const char *)src;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
      const __m128i a = _mm_loadu_si128((const __m128i *)(s + 64 * i)), b = _mm_loadu_si128((const __m128i *)(s + 64 * i + 16)),
                    c = _mm_loadu_si128((const __m128i *)(s + 64 * i + 32)), e = _mm_loadu_si128((const __m128i *)(s + 64 * i + 48));
      _mm_stream_si128((__m128i *)(d + 64 * i), a);
      _mm_stream_si128((__m128i *)(d + 64 * i + 16), b);
      _mm_stream_si128((__m128i *)(d + 64 * i + 32), c);
      _mm_stream_si128((__m128i *)(d + 64 * i + 48), e);
    }
    _mm_sfence();
    if (n > blocks * 64) memcpy(d + blocks * 64, s + blocks * 64, n - blocks * 64);
    return;
  }
#endif
  (void)nt;
  memcpy(dst, src, n);
}
constexpr size_t TQ_PAR_STAGE_MIN = 12u << 20; // (below: starting the helper threads costs more than the driver's pageable copy)
static int stage_helpers(const tq_ctx *ctx) {
  if (ctx->opt_stage_threads > 0) return (int)std::min<long>(ctx->opt_stage_threads, 32);
  return (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
}
static tq_status stage_helper_contexts(tq_ctx *ctx, int n) {
  while ((int)ctx->lanes.size() < n) {
    tq_ctx *c = nullptr;
    tq_status st = tq_create(ctx->device, &c);
    if (st != TQ_OK) return tq_fail(ctx, st, "cannot create an internal lane context");
    c->opt_host_lanes = 1;
    ctx->lanes.push_back(c);
  }
  for (int l = 0; l < n; ++l) TQ_TRY(bounce_init(ctx->lanes[l]));
  return TQ_OK;
}
// dir = 0: user -> device, 1: device -> user
static tq_status stage_parallel(tq_ctx *ctx, void *user, void *dev, size_t bytes, int dir) {
  int T = stage_helpers(ctx);
  // pieces of 1-2 MiB: the CPU copy of one piece overlaps the DMA of the previous one, several pieces per helper
  const size_t piece = std::min<size_t>(2u << 20, std::max<size_t>(1u << 20, (bytes / (size_t)(4 * T) + 4095) & ~size_t(4095)));
  const size_t np = (bytes + piece - 1) / piece;
  T = (int)std::min<size_t>((size_t)T, np);
  TQ_TRY(stage_helper_contexts(ctx, T));
  std::vector<cudaError_t> errs((size_t)T, cudaSuccess);
  auto work = [&](int t) {
    cudaSetDevice(ctx->device);
    tq_ctx *h = ctx->lanes[t];
    cudaError_t e = cudaSuccess;
    const int base = dir ? 2 : 0; // slots 0, 1: host -> device; 2, 3: device -> host
    size_t k = 0;                 // pieces handled by this helper
    size_t prev_off = 0, prev_n = 0;
    int prev_slot = -1;
    for (size_t i = (size_t)t; i < np && e == cudaSuccess; i += (size_t)T, ++k) {
      const int slot = base + (int)(k & 1);
      const size_t off = i * piece, n = std::min(piece, bytes - off);
      if (dir == 0) {
        if (k >= 2) e = cudaEventSynchronize(h->bounce_ev[slot]); // the DMA out of this slot has finished
        if (e != cudaSuccess) break;
        stage_copy(h->bounce[slot], (const char *)user + off, n, ctx->opt_stage_nt != 0);
        e = cudaMemcpyAsync((char *)dev + off, h->bounce[slot], n, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->bounce_ev[slot], h->stream);
      } else {
        e = cudaMemcpyAsync(h->bounce[slot], (const char *)dev + off, n, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->bounce_ev[slot], h->stream);
        if (prev_slot >= 0 && e == cudaSuccess) { // the CPU copy of the previous piece overlaps the DMA of this one
          e = cudaEventSynchronize(h->bounce_ev[prev_slot]);
          if (e == cudaSuccess) memcpy((char *)user + prev_off, h->bounce[prev_slot], prev_n); // (non-temporal stores into the caller's buffer: measured slower)
        }
        prev_slot = slot;
        prev_off = off;
        prev_n = n;
      }
    }
    if (e == cudaSuccess && dir == 1 && prev_slot >= 0) {
      e = cudaEventSynchronize(h->bounce_ev[prev_slot]);
      if (e == cudaSuccess) memcpy((char *)user + prev_off, h->bounce[prev_slot], prev_n);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream); // every piece of this helper has arrived / every slot is free again
    errs[(size_t)t] = e;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  for (int t = 0; t < T; ++t)
    if (errs[(size_t)t] != cudaSuccess) return tq_fail(ctx, TQ_ERR_CUDA, std::string("parallel staging: ") + cudaGetErrorString(errs[(size_t)t]));
  return TQ_OK;
}

tq_status tq_stage_in(tq_ctx *ctx, const void *user, size_t bytes, DeviceBuffer &stage, const void **dev) {
  if (!user) {
    *dev = nullptr;
    return TQ_OK;
  }
  if (tq_is_device_ptr(user)) {
    *dev = user;
    return TQ_OK;
  }
  TQ_TRY(tq_reserve(ctx, stage, bytes));
  *dev = stage.p;
  if (bytes >= TQ_PAR_STAGE_MIN && ctx->opt_bounce == 1 && ctx->opt_host_lanes > 1 && host_is_pageable(user)) {
    // (the device buffer may still be read by kernels of the previous call on this context's stream)
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return stage_parallel(ctx, const_cast<void *>(user), stage.p, bytes, 0);
  }
  // (only inside a multi-lane call: a single thread copying through its own slots is CPU bound at ~5 GB/s, below the driver's 7 GB/s)
  if (bytes >= (1u << 20) && ctx->opt_bounce == 2 && host_is_pageable(user)) {
    TQ_TRY(bounce_init(ctx));
    size_t off = 0;
    for (int i = 0; off < bytes; ++i, off += TQ_BOUNCE_BYTES) {
      const int slot = i & 1;
      const size_t n = std::min(TQ_BOUNCE_BYTES, bytes - off);
      if (i >= 2) TQ_CUDA(ctx, cudaEventSynchronize(ctx->bounce_ev[slot])); // the DMA out of this slot has finished
      memcpy(ctx->bounce[slot], (const char *)user + off, n);
      TQ_CUDA(ctx, cudaMemcpyAsync((char *)stage.p + off, ctx->bounce[slot], n, cudaMemcpyHostToDevice, ctx->stream));
      TQ_CUDA(ctx, cudaEventRecord(ctx->bounce_ev[slot], ctx->stream));
    }
    // the slots are reused by the next staging call of this context: drain them
    TQ_CUDA(ctx, cudaEventSynchronize(ctx->bounce_ev[0]));
    TQ_CUDA(ctx, cudaEventSynchronize(ctx->bounce_ev[1]));
    return TQ_OK;
  }
  TQ_CUDA(ctx, cudaMemcpyAsync(stage.p, user, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return TQ_OK;
}

tq_status tq_stage_out_begin(tq_ctx *ctx, void *user, size_t bytes, DeviceBuffer &stage, void **dev, bool *is_host) {
  if (tq_is_device_ptr(user)) {
    *dev = user;
    *is_host = false;
    return TQ_OK;
  }
  TQ_TRY(tq_reserve(ctx, stage, bytes));
  *dev = stage.p;
  *is_host = true;
  return TQ_OK;
}

tq_status tq_stage_out_finish(tq_ctx *ctx, void *user, size_t bytes, void *dev, bool is_host) {
  if (!is_host) return TQ_OK;
  if (bytes >= TQ_PAR_STAGE_MIN && ctx->opt_bounce == 1 && ctx->opt_host_lanes > 1 && host_is_pageable(user)) {
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the kernels that produce `dev` have finished
    return stage_parallel(ctx, user, dev, bytes, 1);
  }
  if (bytes >= (1u << 20) && ctx->opt_bounce == 2 && host_is_pageable(user)) {
    TQ_TRY(bounce_init(ctx));
    // DMA of piece i into slot 2 + (i & 1) overlaps the CPU copy of piece i - 1 out of the other slot
    const size_t np = (bytes + TQ_BOUNCE_BYTES - 1) / TQ_BOUNCE_BYTES;
    for (size_t i = 0; i <= np; ++i) {
      if (i < np) {
        const int slot = 2 + (int)(i & 1);
        const size_t off = i * TQ_BOUNCE_BYTES, n = std::min(TQ_BOUNCE_BYTES, bytes - off);
        TQ_CUDA(ctx, cudaMemcpyAsync(ctx->bounce[slot], (const char *)dev + off, n, cudaMemcpyDeviceToHost, ctx->stream));
        TQ_CUDA(ctx, cudaEventRecord(ctx->bounce_ev[slot], ctx->stream));
      }
      if (i > 0) {
        const int slot = 2 + (int)((i - 1) & 1);
        const size_t off = (i - 1) * TQ_BOUNCE_BYTES, n = std::min(TQ_BOUNCE_BYTES, bytes - off);
        TQ_CUDA(ctx, cudaEventSynchronize(ctx->bounce_ev[slot]));
        memcpy((char *)user + off, ctx->bounce[slot], n);
      }
    }
    TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TQ_OK;
  }
  TQ_CUDA(ctx, cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TQ_OK;
}

extern "C" tq_status tq_device_alloc(tq_ctx *ctx, size_t bytes, void **dptr) {
  if (!ctx || !dptr) return TQ_ERR_INVALID;
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  TQ_CUDA(ctx, cudaMalloc(dptr, bytes));
  return TQ_OK;
}
extern "C" tq_status tq_device_free(tq_ctx *ctx, void *dptr) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  TQ_CUDA(ctx, cudaFree(dptr));
  return TQ_OK;
}
extern "C" tq_status tq_memcpy_h2d(tq_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TQ_OK;
}
extern "C" tq_status tq_memcpy_d2h(tq_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (!ctx) return TQ_ERR_INVALID;
  TQ_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TQ_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL, loaded at run time (the torch-bundled libnccl.so.2 or the system one): the library has no
// link-time dependency on it, single-GPU users never load it.
// ---------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_t_;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t *);
typedef int (*fn_CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(ncclComm_t_);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);
struct NcclApi {
  void *handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
  std::string load_error;
};
NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char *env = ::getenv("TQ_NCCL_LIB");
  const char *cands[] = {env, "libnccl.so.2", "libnccl.so"};
  std::string why;
  for (const char *c : cands) {
    if (!c) continue;
    api.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    const char *e = dlerror(); // (one call: dlerror() clears the message it returns)
    why += std::string(why.empty() ? "" : "; ") + (e ? e : "?");
    if (c == env) break; // an explicit TQ_NCCL_LIB is not silently replaced by another library
  }
  if (!api.handle) {
    api.load_error = "cannot dlopen libnccl.so.2: " + why;
    return api;
  }
  api.GetUniqueId = (fn_GetUniqueId)dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (fn_CommInitRank)dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (fn_CommDestroy)dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (fn_AllReduce)dlsym(api.handle, "ncclAllReduce");
  api.GetErrorString = (fn_GetErrorString)dlsym(api.handle, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
    api.load_error = "libnccl is missing symbols";
    api.handle = nullptr;
  }
  return api;
}
} // namespace

extern "C" tq_status tq_comm_unique_id(unsigned char id[TQ_COMM_ID_BYTES]) {
  NcclApi &a = nccl();
  if (!a.handle) return TQ_ERR_COMM;
  ncclUniqueId_t u;
  if (a.GetUniqueId(&u) != 0) return TQ_ERR_COMM;
  static_assert(sizeof(u) == TQ_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, TQ_COMM_ID_BYTES);
  return TQ_OK;
}

extern "C" tq_status tq_comm_init(tq_ctx *ctx, int n_ranks, int rank, const unsigned char id[TQ_COMM_ID_BYTES]) {
  if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks) return tq_fail(ctx, TQ_ERR_INVALID, "tq_comm_init: bad rank / n_ranks");
  NcclApi &a = nccl();
  if (!a.handle) return tq_fail(ctx, TQ_ERR_COMM, a.load_error);
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  tq_comm_destroy_internal(ctx);
  ncclUniqueId_t u;
  memcpy(&u, id, TQ_COMM_ID_BYTES);
  ncclComm_t_ comm = nullptr;
  int r = a.CommInitRank(&comm, n_ranks, u, rank);
  if (r != 0) return tq_fail(ctx, TQ_ERR_COMM, std::string("ncclCommInitRank: ") + (a.GetErrorString ? a.GetErrorString(r) : "error"));
  ctx->nccl_comm = comm;
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  return TQ_OK;
}

extern "C" tq_status tq_comm_destroy_internal(tq_ctx *ctx) {
  if (ctx && ctx->nccl_comm) {
    NcclApi &a = nccl();
    if (a.handle) a.CommDestroy((ncclComm_t_)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
  }
  return TQ_OK;
}

extern "C" tq_status tq_all_reduce_sum(tq_ctx *ctx, double *data, size_t count) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!ctx->nccl_comm) {
    if (ctx->n_ranks == 1) return TQ_OK; // single rank: the sum over ranks is the identity
    return tq_fail(ctx, TQ_ERR_COMM, "tq_all_reduce_sum: communicator not initialised (call tq_comm_init)");
  }
  NcclApi &a = nccl();
  // ncclDouble = 8, ncclSum = 0 (nccl.h)
  int r = a.AllReduce(data, data, count, 8, 0, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
  if (r != 0) return tq_fail(ctx, TQ_ERR_COMM, std::string("ncclAllReduce: ") + (a.GetErrorString ? a.GetErrorString(r) : "error"));
  return TQ_OK;
}
