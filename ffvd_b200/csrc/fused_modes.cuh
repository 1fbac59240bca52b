// Mode constants and shared-memory sizing of the fused tile kernel (host-visible part of fused.cuh).
#pragma once
#include "ffvd_common.cuh"

namespace ffvd {

#ifndef FFVD_ABLATE
#define FFVD_ABLATE 0      // timing experiments only: 1 = no S REDs, 2 = no SYRK, 3 = no L^{-1} operand loads,
                           // 4 = no K scratch stores, 5 = no exp, 6 = no r^2 loop
#endif

#ifndef FFVD_SYRK_PIPE
#define FFVD_SYRK_PIPE 1   // 1: software-pipelined SYRK flush (syrk_units_pipe), two staging buffers per warp
#endif

#ifndef FFVD_SCR_TMA
#define FFVD_SCR_TMA 1      // 1: the K tile's L2 scratch copy is made by TMA bulk copies shared -> global (0: 256-bit stores inside the K-tile phase)
#endif
#ifndef FFVD_G2_EARLY
#define FFVD_G2_EARLY 0    // 1: operand prologue of Kbar = A L^{-1} requested before the SYRK (costs the SYRK registers: slower)
#endif

enum { MODE_UNCOLLAPSED = 0, MODE_COLLAPSED_P1 = 1, MODE_COLLAPSED_P2 = 2, MODE_FORWARD = 3, MODE_COND = 4 };

template <int RB>
__host__ __device__ constexpr int bt_of() { return 8 * RB; }

// xsc (SE: scaled x rows, used while the K tile is formed), stage (per-warp 8 x 40 transposition buffers of the SYRK
// flush) and part (128 x 32 partial products of W [Z,1], used by the last phase) are live in disjoint phases separated
// by CTA barriers and share one region
__host__ __device__ inline size_t fused_xsc_part_doubles(int BT, int NW, int nbm) {
  const size_t a = ((size_t)BT * FFVD_XLD + 1) & ~(size_t)1, b = (size_t)128 * 8 * nbm, c = (size_t)NW * 8 * 40 * (FFVD_SYRK_PIPE ? 2 : 1);
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// nbm = ceil((Din + 1) / 8): 8-column blocks of the [Xc,1] / [Z,1] operands
__host__ __device__ inline size_t fused_smem_bytes(int RB, int Mp, int NW, int nbm) {
  const int BT = 8 * RB;
  // every sub-array is rounded up to an even number of doubles so that all of them stay 16-byte aligned
  size_t n = (size_t)BT * (Mp + 4) + (((size_t)(BT + 1) * FFVD_XLD + 1) & ~(size_t)1) +
             fused_xsc_part_doubles(BT, NW, nbm) + 2 * Mp + 64 + 3 * 8 * BT + 64 + 8 + 40 + 64 /* exp table */;
  return n * sizeof(double);
}

}  // namespace ffvd
