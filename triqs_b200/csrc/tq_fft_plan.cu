// Host-side FFT planning: factorisation into the radices the tile engine implements and the master
// twiddle table.  This is the analogue of fftw_plan_many_dft in c++/triqs/gfs/transform/fourier_common.cpp:30,
// except that plans are cached per length in the context instead of re-created on every call.
#include "tq_fft.cuh"

#include <cmath>

FftPlanHost::~FftPlanHost() {
  if (plan.tw) cudaFree((void *)plan.tw);
  if (plan.pos) cudaFree((void *)plan.pos);
}

bool tq_fft_factorize(int N, FftPlan &P) {
  P.N = N;
  P.npass = 0;
  if (N < 1) return false;
  int n = N;
  int c2 = 0, c3 = 0, c5 = 0;
  while (n % 2 == 0) { n /= 2; ++c2; }
  while (n % 3 == 0) { n /= 3; ++c3; }
  while (n % 5 == 0) { n /= 5; ++c5; }
  std::vector<int> rad;
  // other primes (7, 11, 13, 41, 101, ...) go through the generic O(R^2) butterfly; do them first while
  // the sub-blocks are long (more parallel butterflies per pass)
  for (int p = 7; (long)p * p <= n; p += 2)
    while (n % p == 0) { rad.push_back(p); n /= p; }
  if (n > 1) rad.push_back(n);
  for (int r : rad)
    if (r > TQ_MAX_GENERIC_RADIX) return false;
  while (c2 > 0 && c5 > 0) { rad.push_back(10); --c2; --c5; }
  while (c2 >= 4) { rad.push_back(16); c2 -= 4; }
  if (c2 == 3) { rad.push_back(8); c2 = 0; }
  if (c2 == 2) { rad.push_back(4); c2 = 0; }
  if (c2 == 1) { rad.push_back(2); c2 = 0; }
  while (c5 > 0) { rad.push_back(5); --c5; }
  while (c3 > 0) { rad.push_back(3); --c3; }
  if (N == 1) rad.clear();
  if ((int)rad.size() > TQ_MAX_PASSES) return false;
  int M = N;
  for (size_t i = 0; i < rad.size(); ++i) {
    P.radix[i] = rad[i];
    P.M[i] = M;
    M /= rad[i];
  }
  P.npass = (int)rad.size();
  return true;
}

void tq_fft_twiddles_host(int N, std::vector<cd> &tw) {
  tw.resize(N);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int j = 0; j < N; ++j) {
    // reduce to the first octant so that symmetric entries are exact mirror images
    long double x = (long double)j / (long double)N; // fraction of a turn in [0,1)
    int oct = (int)(x * 8.0L);
    long double c, s;
    switch (oct) {
      case 0: c = cosl(two_pi * x); s = sinl(two_pi * x); break;
      case 1: { long double y = 0.25L - x; c = sinl(two_pi * y); s = cosl(two_pi * y); } break;
      case 2: { long double y = x - 0.25L; c = -sinl(two_pi * y); s = cosl(two_pi * y); } break;
      case 3: { long double y = 0.5L - x; c = -cosl(two_pi * y); s = sinl(two_pi * y); } break;
      case 4: { long double y = x - 0.5L; c = -cosl(two_pi * y); s = -sinl(two_pi * y); } break;
      case 5: { long double y = 0.75L - x; c = -sinl(two_pi * y); s = -cosl(two_pi * y); } break;
      case 6: { long double y = x - 0.75L; c = sinl(two_pi * y); s = -cosl(two_pi * y); } break;
      default: { long double y = 1.0L - x; c = cosl(two_pi * y); s = -sinl(two_pi * y); } break;
    }
    tw[j] = cmk((double)c, (double)s);
  }
}

tq_status tq_get_fft_plan(tq_ctx *ctx, long N, std::shared_ptr<FftPlanHost> *out) {
  auto key = std::make_tuple(N, 0);
  auto it = ctx->fft_plans.find(key);
  if (it != ctx->fft_plans.end()) {
    *out = it->second;
    return TQ_OK;
  }
  if (N < 1 || N > (1L << 30)) return tq_fail(ctx, TQ_ERR_INVALID, "FFT length out of range");
  auto ph = std::make_shared<FftPlanHost>();
  ph->supported = tq_fft_factorize((int)N, ph->plan);
  if (!ph->supported)
    return tq_fail(ctx, TQ_ERR_UNSUPPORTED,
                   "FFT length " + std::to_string(N) + " has a prime factor > " + std::to_string(TQ_MAX_GENERIC_RADIX));
  tq_fft_twiddles_host((int)N, ph->tw_host);
  cd *d = nullptr;
  TQ_CUDA(ctx, cudaMalloc(&d, sizeof(cd) * N));
  TQ_CUDA(ctx, cudaMemcpyAsync(d, ph->tw_host.data(), sizeof(cd) * N, cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ph->plan.tw = d;
  ph->pos_host.resize(N);
  for (int k = 0; k < (int)N; ++k) ph->pos_host[k] = fft_pos(ph->plan, k);
  int *dp = nullptr;
  TQ_CUDA(ctx, cudaMalloc(&dp, sizeof(int) * N));
  TQ_CUDA(ctx, cudaMemcpyAsync(dp, ph->pos_host.data(), sizeof(int) * N, cudaMemcpyHostToDevice, ctx->stream));
  TQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ph->plan.pos = dp;
  ctx->fft_plans[key] = ph;
  *out = ph;
  return TQ_OK;
}
