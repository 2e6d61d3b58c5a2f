// Planned (run-time radix pair) instantiation of the pipelined Matsubara kernel, mode 3 (I2: scratch -> tau).
// One translation unit per mode: each kernel carries one unrolled butterfly per register radix (tq_fft2.cuh TQ_RADIX_SWITCH).
#include "fourier_pipe_kernel.cuh"

tq_status tq_launch_pipe_planned_m3(tq_ctx *ctx, const PipeArgs &A, const CUtensorMap &tmap, const CUtensorMap &tmap2, const char *tag) {
  return launch_pipe<3, 0, 0, 0, PIPE_NW>(ctx, A, tmap, tmap2, tag);
}
