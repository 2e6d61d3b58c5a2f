// Shared declarations of libtqgpu: complex helpers, the context, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/tqgpu.h"

#define TQ_HD __host__ __device__ __forceinline__
#define TQ_D __device__ __forceinline__

typedef double2 cd;

// Developer / tuning switches (TQ_PHASE_DEBUG, TQ_PIPE_*, TQ_COL_*, TQ_SCRATCH_ALIAS, TQ_KNOCK ...: some of them change results
// by construction) are read from the environment ONLY in -DTQ_DEVELOPER builds.  The release library never looks at them --
// the names are not even in the binary (tests/test_cabi_symbols.py greps for them); the two switches tests and callers
// legitimately need (engine selection, scratch chunk size) are tq_set_option() of the C ABI.
#ifdef TQ_DEVELOPER
#include <cstdlib>
#define TQ_DEV_ENV(name) ::getenv(name)
#define TQ_DEV_ENV_NAME(name) ::getenv(name)
#else
#define TQ_DEV_ENV(name) ((const char *)nullptr)
#define TQ_DEV_ENV_NAME(name) ((const char *)nullptr)
#endif

TQ_HD cd cmk(double x, double y) { return make_double2(x, y); }
TQ_HD cd cadd(cd a, cd b) { return cmk(a.x + b.x, a.y + b.y); }
TQ_HD cd csub(cd a, cd b) { return cmk(a.x - b.x, a.y - b.y); }
TQ_HD cd cneg(cd a) { return cmk(-a.x, -a.y); }
TQ_HD cd cconj(cd a) { return cmk(a.x, -a.y); }
TQ_HD cd cscale(double s, cd a) { return cmk(s * a.x, s * a.y); }
TQ_HD cd cmul(cd a, cd b) { return cmk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
// a + b*c
TQ_HD cd cfma(cd b, cd c, cd a) { return cmk(fma(b.x, c.x, fma(-b.y, c.y, a.x)), fma(b.x, c.y, fma(b.y, c.x, a.y))); }
// a + s*b (s real)
TQ_HD cd caxpy(double s, cd b, cd a) { return cmk(fma(s, b.x, a.x), fma(s, b.y, a.y)); }
// multiply by +i / -i
TQ_HD cd cmuli(cd a) { return cmk(-a.y, a.x); }
TQ_HD cd cmulmi(cd a) { return cmk(a.y, -a.x); }
TQ_HD double cabs2(cd a) { return fma(a.x, a.x, a.y * a.y); }

// ---------------------------------------------------------------------------------------------

struct DeviceBuffer {
  void *p = nullptr;
  size_t cap = 0;
};

struct MatsubaraPlan;  // fourier_matsubara.cu
struct FftPlanHost;    // tq_fft_plan.cu
struct TailFitterPlan; // tail_fit.cu
struct PipePlanCache;  // fourier_pipe.cu
struct LatTwiddles;    // lattice_pipe.cu
struct ColPlanCache;   // fourier_col.cu
struct RealPlanCache;  // fourier_real.cu
struct LegendrePlanCache; // legendre.cu
struct DirectTwiddles;    // lattice.cu
struct BluePlanCache;     // fourier_matsubara.cu (Bluestein plans)
struct BlueAxisCache;     // ... for plain axis transforms

struct tq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t order_event = nullptr; // tq_stream_wait_external / tq_stream_signal_external
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  // grow-only scratch / staging
  DeviceBuffer stage_in, stage_in2, stage_in3, stage_out, scratch, small, pipe_w, leg_tmp, promote;
  void *pinned = nullptr;
  size_t pinned_cap = 0;
  // pinned bounce buffers for PAGEABLE user memory (tq_stage_in / tq_stage_out_finish): two slots each way, events per slot
  void *bounce[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t bounce_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // caches (mesh-only data, like the reference's tail_fitter::_lss and FFTW plans)
  std::map<std::tuple<long, int>, std::shared_ptr<FftPlanHost>> fft_plans;                        // (N, dummy)
  std::map<std::tuple<double, int, long, long>, std::shared_ptr<MatsubaraPlan>> matsubara_plans;  // beta, stat, n_tau, n_iw
  std::map<std::tuple<double, int, long, double, int, int, int, int, double>, std::shared_ptr<TailFitterPlan>> tail_plans;
  std::shared_ptr<PipePlanCache> pipe_cache;
  std::shared_ptr<LatTwiddles> lat_tw;
  std::shared_ptr<ColPlanCache> col_cache;
  std::shared_ptr<RealPlanCache> real_cache;
  std::shared_ptr<LegendrePlanCache> leg_cache;
  std::shared_ptr<DirectTwiddles> direct_tw;
  std::shared_ptr<BluePlanCache> blue_cache;
  std::shared_ptr<BlueAxisCache> blue_axis_cache;
  // per-kernel event timing (tq_profile_enable)
  bool profiling = false;
  struct ProfRec {
    const char *tag;
    cudaEvent_t e0, e1;
  };
  std::vector<ProfRec> prof;
  // tq_set_option
  long opt_force_planned = 0;    // 1: run the L = 10^5 mesh through the planned (run-time radix) kernels too (A/B measurements)
  long opt_pipe_lag0 = 0;        // packet order of the planned kernels (fourier_pipe_kernel.cuh)
  long opt_no_pipe = 0;          // 1: skip the TMA-pipelined / column specialisations (generic tile kernels only)
  long opt_no_col = 0;           // 1: skip the single-column kernel of fourier_col.cu (A/B measurements)
  long opt_blue_chunk_mb = 512;  // convolution buffer of the Bluestein path per group of gfs (MiB)
  long opt_force_bluestein = 0;  // 1: every imtime <-> imfreq transform with L >= 512 goes through the Bluestein convolution (A/B measurements, tests)
  long opt_no_bluestein = 0;     // 1: lengths without a plan always go through the direct O(N^2) window DFT (A/B measurements, tests)
  long opt_no_two_for_one = 0;   // 1: real-valued G(tau) is promoted column by column instead of packed two columns per complex one (A/B, tests)
  long opt_no_tr = 0;            // 1: narrow targets never run as "gfs as columns" through the planned kernels (A/B measurements)
  long opt_scratch_chunk_mb = 0; // > 0: four-step scratch budget per launch pair in MiB
  long opt_host_lanes = 3;       // internal lanes that overlap H2D / kernels / D2H of host-buffer calls (1: off)
  long opt_stage_nt = 1;         // non-temporal stores in the staging copies of pageable buffers (x86-64)
  long opt_stage_threads = 0;    // helper threads that stage a large pageable buffer of an un-chunked call (0: min(8, cores / 2))
  long opt_bounce = 1;           // stage pageable host buffers through the library's own pinned slots (tq_stage_in)
  long opt_host_chunk_kb = 48L << 10; // input bytes per chunk of a host-buffer call (KiB): one C2 block (40 MB) per chunk
  std::vector<tq_ctx *> lanes;   // child contexts of the lanes (created on first use)
  long call_batch = 0;           // lane contexts: batch of the whole host-buffer call (path decisions must not depend on the chunking)
  // NCCL
  void *nccl_comm = nullptr;
  int n_ranks = 1, rank = 0;
};

tq_status tq_fail(tq_ctx *ctx, tq_status st, const std::string &msg);

#define TQ_CUDA(ctx, call)                                                                                                       \
  do {                                                                                                                           \
    cudaError_t e__ = (call);                                                                                                    \
    if (e__ != cudaSuccess)                                                                                                      \
      return tq_fail(ctx, TQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " at " + __FILE__ + ":" +          \
                                           std::to_string(__LINE__));                                                           \
  } while (0)

#define TQ_TRY(expr)                                                                                                             \
  do {                                                                                                                           \
    tq_status s__ = (expr);                                                                                                      \
    if (s__ != TQ_OK) return s__;                                                                                                \
  } while (0)

tq_status tq_reserve(tq_ctx *ctx, DeviceBuffer &b, size_t bytes);
tq_status tq_reserve_pinned(tq_ctx *ctx, size_t bytes);
bool tq_is_device_ptr(const void *p);

// Resolves a user pointer that may live on the host: returns a device pointer, staging through `stage` if needed.
struct StagedIn {
  const void *dev = nullptr;
};
tq_status tq_stage_in(tq_ctx *ctx, const void *user, size_t bytes, DeviceBuffer &stage, const void **dev);
// Output: returns device pointer to write to; call tq_stage_out_finish afterwards to copy back when user ptr is host.
tq_status tq_stage_out_begin(tq_ctx *ctx, void *user, size_t bytes, DeviceBuffer &stage, void **dev, bool *is_host);
tq_status tq_stage_out_finish(tq_ctx *ctx, void *user, size_t bytes, void *dev, bool is_host);

inline void tq_count_launch(tq_ctx *ctx, int n = 1) { ctx->launches += n; }

// brackets one kernel launch with events when profiling is on:  { KernelScope ks(ctx, "tag"); kernel<<<...>>>(...); }
struct KernelScope {
  tq_ctx *ctx;
  cudaEvent_t e1 = nullptr;
  KernelScope(tq_ctx *c, const char *tag) : ctx(c) {
    if (!ctx->profiling) return;
    cudaEvent_t e0;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    ctx->prof.push_back({tag, e0, e1});
  }
  ~KernelScope() {
    if (e1) cudaEventRecord(e1, ctx->stream);
  }
};
