// Shared plumbing of the TMA-pipelined kernels (fourier_pipe.cu, lattice_pipe.cu): PTX wrappers for mbarriers and bulk
// asynchronous copies, and the driver entry point that encodes tensor maps.
#pragma once
#include "tq_fft2.cuh"

#include <cuda.h>

// -------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk async copies (TMA engine; SASS: UBLKCP / UTMALDG) + proxy fence
// -------------------------------------------------------------------------------------------------
TQ_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
TQ_D void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
TQ_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TQ_D void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TQ_D void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
TQ_D void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef TQ_MBAR_HINT
  // (developer A/B: try_wait with an explicit suspend-time hint, the form CUTLASS uses)
  asm volatile(
     "{\n"
     ".reg .pred P1;\n"
     "LAB_WAIT:\n"
     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
     "@P1 bra DONE;\n"
     "bra LAB_WAIT;\n"
     "DONE:\n"
     "}\n" ::"r"(smem_u32(bar)),
     "r"(parity), "r"((uint32_t)TQ_MBAR_HINT)
     : "memory");
#else
  asm volatile(
     "{\n"
     ".reg .pred P1;\n"
     "LAB_WAIT:\n"
     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
     "@P1 bra DONE;\n"
     "bra LAB_WAIT;\n"
     "DONE:\n"
     "}\n" ::"r"(smem_u32(bar)),
     "r"(parity)
     : "memory");
#endif
}
TQ_D void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
TQ_D void tma_load_4d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                  smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}
TQ_D void tma_load_3d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                  smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of a future tile (no shared-memory destination, no barrier): shortens the flight time of the real copy
TQ_D void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
TQ_D void tma_prefetch_4d(const CUtensorMap *tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
TQ_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// sum_i a_i / (i w - b_i) with b3 = -b2:   -(b1 r1) a1 - i (w r1) a1 - (b2 r2)(a2 - a3) - i (w r2)(a2 + a3)
TQ_D cd tail_iw4(cd w01, cd w23, cd a1, cd D23, cd A23) {
  const double p1 = w01.x, u = w01.y, p2 = w23.x, v = w23.y;
  double re = fma(u, a1.y, -p1 * a1.x);
  re = fma(-p2, D23.x, re);
  re = fma(v, A23.y, re);
  double im = fma(-u, a1.x, -p1 * a1.y);
  im = fma(-p2, D23.y, im);
  im = fma(-v, A23.x, im);
  return cmk(re, im);
}


// -------------------------------------------------------------------------------------------------
// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}


// -------------------------------------------------------------------------------------------------
// host: tile shapes of the planned two-pass engine (tq_fft2.cuh), shared by fourier_pipe.cu and lattice_pipe.cu
// -------------------------------------------------------------------------------------------------
#include <cmath>
#include <cstdlib>
// Tile shape of one pass: length N = RA * RB; RA always has a register butterfly, RB either has one or is handled by direct
// summation (`generic`: a prime factor 11 ... 127 of the mesh length, e.g. 41 | 6150, 101 | 9999).
struct TileShape {
  int N = 0, RA = 0, RB = 0;
  bool generic = false;
  bool heavy = false; // uses a radix of the extended set (tq_fft2.cuh TQ_RADIX_SWITCH_X): separate kernel instantiation
  double cost = 0; // ~ FP64 work per element relative to a register radix pass
};
constexpr int PIPE_MAX_GENERIC = 127;   // largest radix of the direct-sum pass B
static const int kRegRadix[] = {41, 32, 25, 20, 17, 16, 15, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1};

static inline bool tile_shape(int N, TileShape *out, bool allow_heavy = true) {
  bool found = false;
  TileShape best;
  for (int ra : kRegRadix) {
    if (N % ra || (tq_radix_is_heavy(ra) && !allow_heavy)) continue;
    if (!tq_radix_in_registers(ra, true)) continue; // (a radix dropped from the register set by -DTQ_NO_RADER)
    const int rb = N / ra;
    TileShape t;
    t.N = N;
    t.RA = ra;
    t.RB = rb;
    if (tq_radix_in_registers(rb, allow_heavy)) {
      t.generic = false;
      // per-element work of a register pass grows slowly with the radix; two balanced passes beat a long and a trivial one
      // (a pass has RB x columns (A) / RA x columns (B) butterflies to spread over the warps: a lopsided pair starves one of them)
      t.cost = (ra == 1 ? 0.35 : 1.0) + (rb == 1 ? 0.35 : 1.0) + 0.05 * std::min(std::abs(ra - rb), 15);
      // a Rader butterfly is two transforms of length P - 1 plus a pointwise product, and its kernel spills; next to a unit radix
      // a tile has only `columns` such butterflies to spread over the warps
      t.heavy = tq_radix_is_heavy(ra) || tq_radix_is_heavy(rb);
      if (tq_radix_is_heavy(ra)) t.cost += ra == 32 ? 0.3 : 1.0;
      if (tq_radix_is_heavy(rb)) t.cost += rb == 32 ? 0.3 : 1.0;
      if (t.heavy && (ra == 1 || rb == 1)) t.cost += 1.5;
    } else if (rb <= PIPE_MAX_GENERIC) {
      t.generic = true;
      t.heavy = tq_radix_is_heavy(ra);
      if (t.heavy) continue; // (a heavy RA with a direct-sum RB: no such instantiation is worth having)
      t.cost = (ra == 1 ? 0.35 : 1.0) + 0.5 + rb / 9.0; // 4 rb FP64 instructions per element against ~40 for a register pass
    } else {
      continue;
    }
    if (!found || t.cost < best.cost) {
      best = t;
      found = true;
    }
  }
  if (found) *out = best;
  return found;
}

