// Planned instantiation of the pipelined Matsubara kernel with the EXTENDED radix set (tq_fft2.cuh TQ_RADIX_SWITCH_X: + 32 and the
// Rader primes 11, 13, 17, 41), mode 2.  Kept apart from fourier_pipe_planned2.cu so that the register pressure of these
// butterflies does not touch the kernels every smooth mesh length uses.
#include "fourier_pipe_kernel.cuh"

tq_status tq_launch_pipe_heavy_m2(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const CUtensorMap &tmap2, const char *tag) {
  return launch_pipe<2, 0, 0, 0, TQ_NW_HEAVY, 1>(ctx, A, tmap, tmap2, tag);
}
