// Planned instantiation of the pipelined Matsubara kernel with the EXTENDED radix set (tq_fft2.cuh TQ_RADIX_SWITCH_X: + 32 and the
// Rader primes 11, 13, 17, 41), mode 0.  Kept apart from fourier_pipe_planned0.cu so that the register pressure of these
// butterflies does not touch the kernels every smooth mesh length uses.
#include "fourier_pipe_kernel.cuh"

// F1 with a Rader butterfly in pass B spills 648 bytes at 168 registers (12 warps) and 40 at 255 (8 warps): measured with 7 workers
// tau2iw_pass1 1.87 -> 1.21 ms at L = 61500, 0.92 -> 0.82 at 7800, 0.80 -> 0.78 at 6600 (the other three modes are faster with 11 warps for
// at least one of the radices 11 / 13 / 17 / 41 and keep them; profiles/r02_heavy_workers_ab.log)
#ifndef TQ_NW_HEAVY_F1
#define TQ_NW_HEAVY_F1 7
#endif

tq_status tq_launch_pipe_heavy_m0(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const CUtensorMap &tmap2, const char *tag) {
  return launch_pipe<0, 0, 0, 0, TQ_NW_HEAVY_F1, 1>(ctx, A, tmap, tmap2, tag);
}
