// eps_k on a Brillouin-zone mesh from a tight-binding Hamiltonian (the input builder of the Dyson k-sum).
//
// Restates tight_binding::fourier(mesh::brzone)  c++/triqs/lattice/tight_binding.hpp:88-123:
//   h_k = sum_j m_j exp(2 pi i k . r_j),  k in units of the reciprocal lattice vectors = (i0/d0, i1/d1, i2/d2) on the mesh
//   (brzone.hpp:285-288 with the default bz), data index i0 d1 d2 + i1 d2 + i2 (brzone.hpp:187-190).
// The phase is reduced in integer arithmetic ((i_l r_l) mod d_l) before sincospi, so it is exact up to the final rounding.
#include "tq_common.cuh"

__global__ void k_tb_fourier(const int *__restrict__ displ, const cd *__restrict__ hop, int n_r, int nn, long d0, long d1, long d2, long total,
                             cd *__restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long k = i / nn;
  const int e = (int)(i - k * nn);
  const long i2 = k % d2, i1 = (k / d2) % d1, i0 = k / (d1 * d2);
  cd acc = cmk(0.0, 0.0);
  for (int j = 0; j < n_r; ++j) {
    const long r0 = displ[3 * j], r1 = displ[3 * j + 1], r2 = displ[3 * j + 2];
    long m0 = (i0 * r0) % d0, m1 = (i1 * r1) % d1, m2 = (i2 * r2) % d2;
    if (m0 < 0) m0 += d0;
    if (m1 < 0) m1 += d1;
    if (m2 < 0) m2 += d2;
    double x = (double)m0 / (double)d0 + (double)m1 / (double)d1 + (double)m2 / (double)d2; // turns
    x -= floor(x);
    double s, c;
    sincospi(2.0 * x, &s, &c);
    acc = cfma(__ldg(&hop[(size_t)j * nn + e]), cmk(c, s), acc);
  }
  out[i] = acc;
}

extern "C" tq_status tq_tight_binding_fourier(tq_ctx *ctx, int n_orb, int n_r, const int *displ, const tq_complex *hop, const long dims[3],
                                              tq_complex *out) {
  if (!ctx) return TQ_ERR_INVALID;
  if (!displ || !hop || !dims || !out || n_orb < 1 || n_r < 1 || dims[0] < 1 || dims[1] < 1 || dims[2] < 1)
    return tq_fail(ctx, TQ_ERR_INVALID, "tq_tight_binding_fourier: invalid argument");
  TQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const long nk = dims[0] * dims[1] * dims[2];
  const int nn = n_orb * n_orb;
  const long total = nk * nn;
  const void *d_displ = nullptr, *d_hop = nullptr;
  void *d_out = nullptr;
  bool out_host = false;
  TQ_TRY(tq_stage_in(ctx, displ, sizeof(int) * 3 * (size_t)n_r, ctx->stage_in, &d_displ));
  TQ_TRY(tq_stage_in(ctx, hop, sizeof(cd) * (size_t)n_r * nn, ctx->stage_in2, &d_hop));
  TQ_TRY(tq_stage_out_begin(ctx, out, sizeof(cd) * (size_t)total, ctx->stage_out, &d_out, &out_host));
  {
    KernelScope ks(ctx, "tight_binding_fourier");
    k_tb_fourier<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const int *)d_displ, (const cd *)d_hop, n_r, nn, dims[0], dims[1], dims[2], total,
                                                                          (cd *)d_out);
  }
  tq_count_launch(ctx);
  TQ_CUDA(ctx, cudaGetLastError());
  TQ_TRY(tq_stage_out_finish(ctx, out, sizeof(cd) * (size_t)total, d_out, out_host));
  return TQ_OK;
}
