// Fused per-tile GPSSM kernels (sm_100a): K tile on the fly -> L^{-1} contraction on the FP64
// tensor pipe (DMMA m8n8k4) -> mean / variance / Gaussian log-density -> hand-derived backward.
// Replaces, per (sample s, tile of BT time steps, output dim d), the TF sub-graph built by
//   conditionals_multi_output.py:73-120 + dgp_model.py:337-359 (uncollapsed),
//   conditionals_multi_output.py:230-257 (collapsed), and its tf.gradients (base_model.py:148).
// Nothing of size T x M ever goes to HBM.
#pragma once
#include "ffvd_common.cuh"

namespace ffvd {

// Optional per-phase cycle accounting (build with -DFFVD_PHASE_TIMING; tools/phase_timing.py).  Thread 0 of
// every CTA adds the clock64() delta since its previous mark to g_phase_clocks[i]; phases end at CTA
// barriers, so thread 0's timeline is the CTA's.
#ifdef FFVD_PHASE_TIMING
__device__ unsigned long long g_phase_clocks[16];
#define FFVD_MARK(i)                                                                 \
  do {                                                                               \
    if (tid == 0) {                                                                  \
      const long long _now = clock64();                                              \
      atomicAdd(&g_phase_clocks[i], (unsigned long long)(_now - _phase_last));       \
      _phase_last = _now;                                                            \
    }                                                                                \
  } while (0)
#else
#define FFVD_MARK(i) do { } while (0)
#endif

} // namespace ffvd
#include "fused_modes.cuh"
namespace ffvd {

// Four consecutive doubles of a tile row as two 16-byte shared-memory accesses.  A quarter-warp (the unit of a 16-byte
// access) is lanes (g, g+1) x q = 0..3, and with lda = 4 mod 16 rows g and g+1 start 4 doubles (two 16-byte banks) apart:
// if every lane touched columns {0,1} first, lane (g, q) and lane (g+1, q-1) would hit the same bank (ncu: 11% of the
// kernel's shared wavefronts were such 2-way conflicts).  Odd rows therefore go {2,3} first, {0,1} second.
__device__ __forceinline__ void st_tile4(double* p, int g, double a, double b, double c, double d) {
  const bool odd = (g & 1) != 0;
  *reinterpret_cast<double2*>(p + (odd ? 2 : 0)) = odd ? make_double2(c, d) : make_double2(a, b);
  *reinterpret_cast<double2*>(p + (odd ? 0 : 2)) = odd ? make_double2(a, b) : make_double2(c, d);
}
__device__ __forceinline__ void ld_tile4(const double* p, int g, double& a, double& b, double& c, double& d) {
  const bool odd = (g & 1) != 0;
  const double2 u = *reinterpret_cast<const double2*>(p + (odd ? 2 : 0));
  const double2 v = *reinterpret_cast<const double2*>(p + (odd ? 0 : 2));
  a = odd ? v.x : u.x; b = odd ? v.y : u.y; c = odd ? u.x : v.x; d = odd ? u.y : v.y;
}

#define FFVD_TILE_ROWS_FACTOR 2      /* compute_k_tile sees RB = rows per warp / 8; a tile has at most twice that (NW = 16) */

struct Smem {
  double* tile;     // BT x lda
  double* xs;       // (BT+1) x XLD : X~ rows t0..t0+BT  = [x_t, ctrl_t, 1, 0..]
  double* xsc;      // BT x XLD     : SE: augmented scaled rows [x/l, -1/2 |x/l|^2, 1, 0..] (A operand of the K-tile product); unused for Linear
  double* us;       // Mp           : u_d (uncollapsed) / w'_d (collapsed pass 2)
  double* ws;       // Mp           : uncollapsed: w_d = L^{-T} u_d
  double* es;       // 64           : e_t (uncollapsed) / delta_t (collapsed)
  double* rowpart;  // 3 x 8 x BT
  double* stage;    // NW warps x 8 x 40
  double* part;     // 128 x 8 NBM: partial products of W [Z,1] per k slice (overlays xsc / stage)
  double* small;    // 64: invl2[32], sil[32]
  double* sc;       // 8: per-item scalars v, Q, 1/Q, log Q
  double* red;      // 40: block-level reductions (smem atomics)
  double* exptab;   // 64: v_d 2^(j/64), the table of exp_nonpos_n pre-scaled by the current output dim's kernel variance
};

// ---------------------------------------------------------------------------------------------
// K tile: k(x_r, z_j) for this warp's rows [row0, row0 + 8 RB) and column groups, written to the shared tile (and,
// when SCR, to the CTA's L2 scratch copy that W = Kbar o K reads back).  Rows >= nvalid are NOT masked here (the
// caller zeroes them in the rare partial tile).  Columns >= M: Linear gets exact zeros (zero columns of Z~^T); SE gets
// v * exp(-745) ~ 1e-308 (hyper_kernel puts -1e4 into their -|z~|^2/2 slot), which every consumer multiplies by an exact
// zero (the padded rows / columns of L^{-1}, L^{-T}, N, u, w) -- no per-element mask in the tile loop.
// SE follows the reference's expansion (kernels_multi_output.py:163-182): -r^2/2 = x~.z~ - |x~|^2/2 - |z~|^2/2 with
// x~ = x/l and z~ = z/l, formed as ONE product of augmented operands on the tensor pipe,
//     [x~, -|x~|^2/2, 1, 0..] (sm.xsc, BT x 4 KS)  times  [z~; 1; -|z~|^2/2; 0..] (ZTd, 4 KS x Mp, written by hyper_kernel),
// KS = ceil((Din + 2) / 4) DMMA k-steps.  As scalar FMAs (one per element and input dim, the first version) the phase was
// issue bound: 3.8k instructions per warp and tile at Din = 9, 2.4x the FP64-pipe time; a DMMA does the work of 8 DFMA
// warp-instructions in the same pipe time.  The two DMMAs of a 16-column group take the even / the odd columns as their
// B fragments (one 16-byte load per lane and k-step), so lane (g, q) ends up with row g, columns 4 q .. 4 q + 3 -- the
// layout of every other tile phase (see gemm_segment).  exp through the branch-free, lock-step exp_nonpos_n with the table pre-scaled by v_d.
// Linear: ZTd = Z~^T unscaled, KS = ceil(Din / 4), A columns >= Din masked (column Din of xs holds the ones).
// KS > 0: k-steps unrolled and the B fragments of the NEXT column group requested before the work of the current one
// (4 KS registers); KS == 0: any Din, k loop rolled, fragments loaded where they are used.
template <int KIND, int RB, int NGW, bool SCR, int NCW, int KS>
__device__ __forceinline__ void compute_k_tile(const Smem& sm, const double* __restrict__ ZTd, int Mp, int Din, double v, int nmin,
                                               int ks_rt, int wc, int g, int q, int M, int row0, int lda,
                                               double* __restrict__ kscr, const double2 (&kb0)[FFVD_KB0]) {
  // The tile of a warp (8 RB rows x 16 NGW columns) is formed in passes of RBB row blocks x one 16-column group: 4 RBB
  // accumulators per lane.  The pass loop is NOT unrolled (instruction-cache footprint of the item loop).
  constexpr int RBB = RB >= 4 ? 4 : RB;
  constexpr int NPASS = RB / RBB;
  constexpr int KSA = KS > 0 ? KS : 1;
  const double* zq = ZTd + (size_t)q * Mp + 2 * g;
  const double* xsrc = ((KIND == 0) ? sm.xsc : sm.xs) + (row0 + g) * FFVD_XLD + q;
  double bc[KSA][2], bn[KSA][2];
  auto load_b = [&](double (&b)[KSA][2], int jg) {
#pragma unroll
    for (int ks = 0; ks < KSA; ++ks) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(zq + (size_t)(4 * ks) * Mp + jg));
      b[ks][0] = t.x; b[ks][1] = t.y;
    }
  };
  if (KS > 0) {
    // group 0: the first FFVD_KB0 k-steps were requested by the caller at the top of the d iteration (kb0)
#pragma unroll
    for (int ks = 0; ks < KSA; ++ks) {
      if (ks < FFVD_KB0) { bc[ks][0] = kb0[ks].x; bc[ks][1] = kb0[ks].y; }
      else {
        const double2 t = __ldg(reinterpret_cast<const double2*>(zq + (size_t)(4 * ks) * Mp + 16 * group_index<NCW>(wc, 0)));
        bc[ks][0] = t.x; bc[ks][1] = t.y;
      }
    }
  }
#pragma unroll 1
  for (int ng = 0; ng < NGW; ++ng) {
    const int jg = 16 * group_index<NCW>(wc, ng);
    const int jb = jg + 4 * q;
    if (KS > 0 && ng + 1 < NGW) load_b(bn, 16 * group_index<NCW>(wc, ng + 1));
#pragma unroll 1
    for (int ps = 0; ps < NPASS; ++ps) {
      const int rbase = ps * RBB;
      const double* xa = xsrc + 8 * rbase * FFVD_XLD;
      double kv[RBB * 4];
#pragma unroll
      for (int i = 0; i < RBB * 4; ++i) kv[i] = 0.0;
      if (FFVD_ABLATE != 6) {
        if (KS > 0) {
#pragma unroll
          for (int ks = 0; ks < KSA; ++ks) {
#pragma unroll
            for (int rb = 0; rb < RBB; ++rb) {
              double a = xa[8 * rb * FFVD_XLD + 4 * ks];
              if (KIND == 1 && 4 * ks + 3 >= Din) a = (4 * ks + q < Din) ? a : 0.0;
              dmma884(kv[4 * rb], kv[4 * rb + 2], a, bc[ks][0]);
              dmma884(kv[4 * rb + 1], kv[4 * rb + 3], a, bc[ks][1]);
            }
          }
        } else {
#pragma unroll 1
          for (int ks = 0; ks < ks_rt; ++ks) {
            const double2 b2 = __ldg(reinterpret_cast<const double2*>(zq + (size_t)(4 * ks) * Mp + jg));
#pragma unroll
            for (int rb = 0; rb < RBB; ++rb) {
              double a = xa[8 * rb * FFVD_XLD + 4 * ks];
              if (KIND == 1) a = (4 * ks + q < Din) ? a : 0.0;
              dmma884(kv[4 * rb], kv[4 * rb + 2], a, b2.x);
              dmma884(kv[4 * rb + 1], kv[4 * rb + 3], a, b2.y);
            }
          }
        }
      }
#if FFVD_ABLATE != 5
      if (KIND == 0) exp_nonpos_n<RBB * 4>(kv, sm.exptab, nmin);
#endif
      if (KIND == 1) {
#pragma unroll
        for (int i = 0; i < RBB * 4; ++i) kv[i] *= v;
      }
#pragma unroll
      for (int i = 0; i < RBB; ++i) {
        const int row = row0 + 8 * (rbase + i) + g;
        double* p = sm.tile + row * lda + jb;
        FFVD_ASSERT(row >= 0 && row < 8 * RB * (FFVD_TILE_ROWS_FACTOR) && jb >= 0 && jb + 4 <= Mp && jb + 4 <= lda);
        st_tile4(p, g, kv[4 * i], kv[4 * i + 1], kv[4 * i + 2], kv[4 * i + 3]);
        if (SCR && FFVD_ABLATE != 4) stg256(kscr + (size_t)row * Mp + jb, kv[4 * i], kv[4 * i + 1], kv[4 * i + 2], kv[4 * i + 3]);
      }
    }
    if (KS > 0) {
#pragma unroll
      for (int ks = 0; ks < KSA; ++ks) { bc[ks][0] = bn[ks][0]; bc[ks][1] = bn[ks][1]; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// acc (BT x Mp, distributed) = Aop(tile) (BT x Mp) * B (Mp x Mp, row-major in global/L2).
// TRI = +1: B upper triangular (B[k][n] != 0 only for k <= n);  -1: lower;  0: dense.
// Lane (g,q) of the warp ends up with rows 8*rb+g, columns 16*group+4q+{0,1,2,3}.
//
// The k range is cut into segments on which the set of column groups with non-zero B rows is fixed at compile
// time (a suffix of the warp's groups for TRI > 0, a prefix for TRI < 0: the groups are in ascending column order),
// so no DMMA is ever issued predicated-off -- on sm_100a a predicated-off DMMA still occupies the tensor pipe.
// B fragments stream from L2 through a 4-slot register ring, 3 k-steps ahead; the k loop is unrolled by 4 (segment
// lengths are multiples of 16 rows) so ring slots are static and no register moves are needed.
template <int RB, int NG, int LO, int HI, int TRI, class AOp>
__device__ __forceinline__ void gemm_segment(double (&acc)[NG][RB][4], double2 (&ring)[4][NG], const double* aq, int lda,
                                             const double* bq, int Mp, const int (&joff)[NG], int kbeg, int kend, AOp aop,
                                             int q) {
  // loads issued in this segment: groups [LO,HI) unconditionally (the k-step they are for is inside their non-zero
  // range), plus one boundary group under a predicate:
  //   TRI > 0: group LO only while the step is still below its end  (kl < joff[LO] + 16)
  //   TRI < 0: group HI (if any) once the step has reached its start (kl >= joff[HI])
  //   loads past the matrix (kl >= Mp) are suppressed
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kc = k0 + 4 * u;              // k-step computed now (ring slot u)
      const int kl = kc + 12;                 // k-step loaded now  (ring slot (u + 3) & 3)
      const double* bl = bq + (size_t)kl * Mp;
#pragma unroll
      for (int ng = 0; ng < NG; ++ng) {
        bool doit = (ng >= LO && ng < HI);
        if (TRI > 0 && ng == LO) doit = kl < joff[LO] + 16;
        if (TRI < 0 && ng == HI) doit = kl >= joff[ng];
        if (TRI <= 0) doit = doit && (kl < Mp);
        if ((ng >= LO && ng < HI) || (TRI < 0 && ng == HI)) {
          if (doit) {
            FFVD_ASSERT(kl >= 0 && kl + q < Mp && joff[ng] >= 0 && joff[ng] + 16 <= Mp);
            ring[(u + 3) & 3][ng] = ldg_stream2(bl + joff[ng]);
          }
        }
      }
      double a[RB];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) a[rb] = aop(aq[rb * 8 * lda + kc], rb, kc + q);
#pragma unroll
      for (int ng = LO; ng < HI; ++ng) {
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          dmma884(acc[ng][rb][0], acc[ng][rb][2], a[rb], ring[u][ng].x);
          dmma884(acc[ng][rb][1], acc[ng][rb][3], a[rb], ring[u][ng].y);
        }
      }
    }
  }
}

template <int RB, int NG, int A, int TRI, class AOp>
__device__ __forceinline__ void gemm_segments(double (&acc)[NG][RB][4], double2 (&ring)[4][NG], const double* aq, int lda,
                                              const double* bq, int Mp, const int (&joff)[NG], AOp aop, int q) {
  if constexpr (A < NG) {
    if constexpr (TRI > 0) {
      const int kbeg = (A == 0) ? 0 : joff[A - 1 < 0 ? 0 : A - 1] + 16;
      gemm_segment<RB, NG, A, NG, TRI>(acc, ring, aq, lda, bq, Mp, joff, kbeg, joff[A] + 16, aop, q);
    } else {
      const int kend = (A + 1 < NG) ? joff[A + 1 < NG ? A + 1 : A] : Mp;
      gemm_segment<RB, NG, 0, A + 1, TRI>(acc, ring, aq, lda, bq, Mp, joff, joff[A], kend, aop, q);
    }
    gemm_segments<RB, NG, A + 1, TRI>(acc, ring, aq, lda, bq, Mp, joff, aop, q);
  }
}

// Prologue of a contraction: the B fragments of the first three k-steps of the first segment.  Split from tile_gemm_chunk so
// that the caller can issue it EARLY (before the CTA barrier / the SYRK in front of the contraction): the operands come from
// L2 (~800 clk), and with all warps starting a contraction together nothing else hides that round trip.  volatile: the
// requests must stay where they are written (a plain asm load may be sunk to its first use).
template <int NGW, int TRI, int NCW>
__device__ __forceinline__ void tile_gemm_prologue(double2 (&ring)[4][NGW], const double* __restrict__ B, int Mp, int warp, int ng0,
                                                   int g, int q) {
  int joff[NGW];
#pragma unroll
  for (int ng = 0; ng < NGW; ++ng) joff[ng] = 16 * group_index<NCW>(warp, ng0 + ng);
  const double* bq = B + (size_t)q * Mp + 2 * g;
  const int kfirst = (TRI < 0) ? joff[0] : 0;
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int ng = 0; ng < NGW; ++ng) {
      ring[u][ng] = make_double2(0.0, 0.0);
      const bool active = (TRI < 0) ? (ng == 0) : true;    // TRI < 0 starts with group 0 alone; else every group is live at k = 0
      if (active) ring[u][ng] = ldg_stream2_v(bq + (size_t)(kfirst + 4 * u) * Mp + joff[ng]);
    }
#pragma unroll
  for (int ng = 0; ng < NGW; ++ng) ring[3][ng] = make_double2(0.0, 0.0);
}

template <int RB, int NGW, int TRI, int NCW, class AOp>
__device__ __forceinline__ void tile_gemm_chunk(double (&acc)[NGW][RB][4], double2 (&ring)[4][NGW], const double* tile, int lda,
                                                const double* __restrict__ B, int Mp, int warp, int ng0, int g, int q, AOp aop) {
  int joff[NGW];
#pragma unroll
  for (int ng = 0; ng < NGW; ++ng) {
    joff[ng] = 16 * group_index<NCW>(warp, ng0 + ng);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[ng][rb][c] = 0.0;
  }
  const double* bq = B + (size_t)q * Mp + 2 * g;
  const double* aq = tile + g * lda + q;
  if constexpr (TRI == 0) {
    gemm_segment<RB, NGW, 0, NGW, 0>(acc, ring, aq, lda, bq, Mp, joff, 0, Mp, aop, q);
  } else {
    gemm_segments<RB, NGW, 0, TRI>(acc, ring, aq, lda, bq, Mp, joff, aop, q);
  }
}

// Column groups are processed at most 4 at a time so that the B-fragment prefetch ring and the running
// pointers stay in registers for large M (NGW up to 16); the A fragments are re-read from shared memory per chunk.
// ring0: the prologue of the FIRST chunk, issued by the caller (tile_gemm_prologue<tile_gemm_ch0<NGW>(), TRI, NCW>(ring0, B, Mp, warp, 0, g, q)).
template <int NGW>
__host__ __device__ constexpr int tile_gemm_ch0() { return NGW <= 4 ? NGW : ((NGW % 4 == 0) ? 4 : 2); }

template <int RB, int NGW, int TRI, int NCW, class AOp>
__device__ __forceinline__ void tile_gemm(double (&acc)[NGW][RB][4], double2 (&ring0)[4][tile_gemm_ch0<NGW>()], const double* tile, int lda,
                                          const double* __restrict__ B, int Mp, int warp, int g, int q, AOp aop) {
  if constexpr (NGW <= 4) {
    tile_gemm_chunk<RB, NGW, TRI, NCW>(acc, ring0, tile, lda, B, Mp, warp, 0, g, q, aop);
  } else {
    static_assert(NGW % 2 == 0, "large NGW must be even");
    constexpr int CH = tile_gemm_ch0<NGW>();
    tile_gemm_chunk<RB, CH, TRI, NCW>(reinterpret_cast<double(&)[CH][RB][4]>(acc[0]), ring0, tile, lda, B, Mp, warp, 0, g, q, aop);
#pragma unroll
    for (int c0 = CH; c0 < NGW; c0 += CH) {
      double2 ring[4][CH];
      tile_gemm_prologue<CH, TRI, NCW>(ring, B, Mp, warp, c0, g, q);
      tile_gemm_chunk<RB, CH, TRI, NCW>(reinterpret_cast<double(&)[CH][RB][4]>(acc[c0]), ring, tile, lda, B, Mp, warp, c0, g, q, aop);
    }
  }
}

// store the distributed (BT x Mp) fragment set into the shared tile
template <int RB, int NGW, int NCW>
__device__ __forceinline__ void store_tile(const double (&acc)[NGW][RB][4], double* tile, int lda, int warp, int g,
                                           int q) {
#pragma unroll
  for (int ng = 0; ng < NGW; ++ng) {
    const int jb = 16 * group_index<NCW>(warp, ng) + 4 * q;
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      double* p = tile + (8 * rb + g) * lda + jb;
      FFVD_ASSERT(jb >= 0 && jb + 4 <= lda - 4);
      st_tile4(p, g, acc[ng][rb][0], acc[ng][rb][1], acc[ng][rb][2], acc[ng][rb][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// S (lower triangle, in 8x8 blocks) += tile^T tile  over the BT rows, flushed with coalesced RED.add.f64.
// Work units: square tiles of NB x NB blocks (NB = 4: 32 x 32; NB = 2: 16 x 16 for Mp <= 128, where 32 x 32 tiles are too
// few to go round -- 10 units for 16 warps at Mp = 128 left six warps idle and the phase at 48% of its DMMA floor): the
// nt(nt-1)/2 strictly-lower tiles followed by the nt diagonal tiles (only the blocks on or below the diagonal are formed),
// dealt to the warps in snake order so the per-warp block counts differ by at most a few percent (Mp = 256: 64 vs 68).
template <int RB, bool DIAG, int NB>
__device__ __forceinline__ void syrk_tile(const double* tile, int lda, int Mp, double* __restrict__ S, double* stage_w,
                                          int m0, int n0, int lane) {
  const int g = lane >> 2, q = lane & 3;
  double c[NB][NB][2];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) c[i][j][0] = c[i][j][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < 8 * RB; k0 += 4) {
    double a[NB], b[NB];
    const double* row = tile + (k0 + q) * lda + g;
#pragma unroll
    for (int i = 0; i < NB; ++i) a[i] = row[m0 + 8 * i];
#pragma unroll
    for (int j = 0; j < NB; ++j) b[j] = DIAG ? a[j] : row[n0 + 8 * j];
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (!DIAG || j <= i) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
  }
  // transpose 8 rows at a time through the per-warp staging buffer (8 x 40 doubles), flush with coalesced REDs
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NB; ++j)
      *reinterpret_cast<double2*>(stage_w + g * 40 + 8 * j + 2 * q) = make_double2(c[i][j][0], c[i][j][1]);
    __syncwarp();
    const int ncol = DIAG ? 8 * (i + 1) : 8 * NB;  // diagonal tiles: nothing right of block column i in block row i
    if (lane < ncol) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#if FFVD_ABLATE != 1
        FFVD_ASSERT(m0 + 8 * i + r < Mp && n0 + lane < Mp && n0 <= m0);
        red_add(S + (size_t)(m0 + 8 * i + r) * Mp + n0 + lane, stage_w[r * 40 + lane]);
#else
        if (stage_w[r * 40 + lane] == 1.2345e300) red_add(S, 1.0);
#endif
      }
    }
  }
}

template <int RB, int NW, int NB>
__device__ __forceinline__ void syrk_units(const double* tile, int lda, int Mp, double* __restrict__ S, double* stage_w,
                                           int warp, int lane) {
  const int nt = Mp / (8 * NB);
  const int nfull = nt * (nt - 1) / 2, nunits = nfull + nt;
  // units of this warp: i = 0 .. nu-1 with u(i) = i NW + (i odd ? NW-1-warp : warp) < nunits
  int nu = 0;
  while (nu * NW + ((nu & 1) ? (NW - 1 - warp) : warp) < nunits) ++nu;
  // The two warps that share an SM sub-partition (w and w+4) walk their lists in opposite directions: one starts with
  // full tiles, the other with its (shorter) diagonal tiles, so their flushes -- during which a warp issues no DMMA --
  // do not coincide and the partner keeps the tensor pipe busy.
  const bool rev = NW >= 8 && (warp & 4) != 0;
  for (int ii = 0; ii < nu; ++ii) {
    const int i = rev ? nu - 1 - ii : ii;
    const int u = i * NW + ((i & 1) ? (NW - 1 - warp) : warp);
    if (u < nfull) {
      // decode (ti > tj) from the linear strictly-lower index u = ti(ti-1)/2 + tj
      int ti = (int)((sqrtf(8.0f * (float)u + 1.0f) + 1.0f) * 0.5f);
      while (ti * (ti - 1) / 2 > u) --ti;
      while ((ti + 1) * ti / 2 <= u) ++ti;
      const int tj = u - ti * (ti - 1) / 2;
      syrk_tile<RB, false, NB>(tile, lda, Mp, S, stage_w, 8 * NB * ti, 8 * NB * tj, lane);
    } else {
      const int t = u - nfull;
      syrk_tile<RB, true, NB>(tile, lda, Mp, S, stage_w, 8 * NB * t, 8 * NB * t, lane);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Software-pipelined SYRK (FFVD_SYRK_PIPE): the flush of unit u (transposition through the staging buffer + REDs) is issued
// in eight slots inside the k loop of unit u + 1, so the warp keeps issuing DMMAs while the results of the previous unit
// drain.  During a flush executed on its own a warp issues no DMMA and its sub-partition partner alone reaches only ~85% of
// the issue rate (ncu: the flush regions are 3% of the kernel's warp samples).  Two staging buffers per warp (block rows
// alternate), so a block row's stores and its loads two k-steps later are separated by exactly one __syncwarp.
template <int NB>
struct SyrkPending {
  double c[NB][NB][2];
  int m0, n0;
  bool diag, valid;
};

// one flush slot: slot = 2 i     -> store block row i of the pending unit into staging buffer (i & 1)
//                 slot = 2 i + 1 -> load it back transposed and RED it into S
template <int NB>
__device__ __forceinline__ void syrk_flush_slot(const SyrkPending<NB>& pd, int slot, int Mp, double* __restrict__ S, double* stage_w,
                                                int lane) {
  const int g = lane >> 2, q = lane & 3;
  const int i = slot >> 1;
  double* st = stage_w + (i & 1) * (8 * 40);
  if ((slot & 1) == 0) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      *reinterpret_cast<double2*>(st + g * 40 + 8 * j + 2 * q) = make_double2(pd.c[i][j][0], pd.c[i][j][1]);
    __syncwarp();
  } else {
    const int ncol = pd.diag ? 8 * (i + 1) : 8 * NB;
    if (lane < ncol) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        FFVD_ASSERT(pd.m0 + 8 * i + r < Mp && pd.n0 + lane < Mp && pd.n0 <= pd.m0);
        red_add(S + (size_t)(pd.m0 + 8 * i + r) * Mp + pd.n0 + lane, st[r * 40 + lane]);
      }
    }
    __syncwarp();
  }
}

template <int RB, bool DIAG, int NB>
__device__ __forceinline__ void syrk_tile_pipe(const double* tile, int lda, int Mp, double* __restrict__ S, double* stage_w,
                                               int m0, int n0, int lane, SyrkPending<NB>& pd) {
  const int g = lane >> 2, q = lane & 3;
  double c[NB][NB][2];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  constexpr int KSTEPS = 2 * RB;                       // k-steps of 4 rows
  constexpr int NSLOT = 2 * NB;                        // flush slots of the pending unit
  constexpr int EVERY = KSTEPS >= NSLOT ? KSTEPS / NSLOT : 1;
  // operand fragments double buffered in registers: the loads of k-step ks + 1 are issued before the DMMAs of k-step ks
  double a[NB], b[NB], an[NB], bn[NB];
  auto load_frag = [&](double (&fa)[NB], double (&fb)[NB], int ks) {
    const double* row = tile + (4 * ks + q) * lda + g;
#pragma unroll
    for (int i = 0; i < NB; ++i) fa[i] = row[m0 + 8 * i];
#pragma unroll
    for (int j = 0; j < NB; ++j) fb[j] = DIAG ? fa[j] : row[n0 + 8 * j];
  };
  load_frag(a, b, 0);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    if (ks + 1 < KSTEPS) load_frag(an, bn, ks + 1);
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (!DIAG || j <= i) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
#pragma unroll
    for (int i = 0; i < NB; ++i) { a[i] = an[i]; b[i] = bn[i]; }
    if (KSTEPS >= NSLOT) {
      if (ks % EVERY == EVERY - 1 && ks / EVERY < NSLOT && pd.valid) syrk_flush_slot<NB>(pd, ks / EVERY, Mp, S, stage_w, lane);
    }
  }
  if (KSTEPS < NSLOT && pd.valid) {
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) syrk_flush_slot<NB>(pd, sl, Mp, S, stage_w, lane);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) { pd.c[i][j][0] = c[i][j][0]; pd.c[i][j][1] = c[i][j][1]; }
  pd.m0 = m0; pd.n0 = n0; pd.diag = DIAG; pd.valid = true;
}

template <int RB, int NW, int NB, class F>
__device__ __forceinline__ void syrk_units_pipe(const double* tile, int lda, int Mp, double* __restrict__ S, double* stage_w,
                                                int warp, int lane, F before_drain) {
  const int nt = Mp / (8 * NB);
  const int nfull = nt * (nt - 1) / 2, nunits = nfull + nt;
  // units of this warp: i = 0 .. nu-1 with u(i) = i NW + (i odd ? NW-1-warp : warp) < nunits  (snake order, see syrk_units)
  const int full = nunits / NW, rem = nunits % NW;
  const int nu = full + ((((full & 1) ? (NW - 1 - warp) : warp) < rem) ? 1 : 0);
  const bool rev = NW >= 8 && (warp & 4) != 0;
  SyrkPending<NB> pd;
  pd.valid = false; pd.diag = false; pd.m0 = pd.n0 = 0;
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) pd.c[i][j][0] = pd.c[i][j][1] = 0.0;
#pragma unroll 1
  for (int ii = 0; ii < nu; ++ii) {
    const int i = rev ? nu - 1 - ii : ii;
    const int u = i * NW + ((i & 1) ? (NW - 1 - warp) : warp);
    if (u < nfull) {
      int ti = (int)((sqrtf(8.0f * (float)u + 1.0f) + 1.0f) * 0.5f);
      ti -= (ti * (ti - 1) / 2 > u) ? 1 : 0;
      ti += ((ti + 1) * ti / 2 <= u) ? 1 : 0;
      const int tj = u - ti * (ti - 1) / 2;
      syrk_tile_pipe<RB, false, NB>(tile, lda, Mp, S, stage_w, 8 * NB * ti, 8 * NB * tj, lane, pd);
    } else {
      const int t = u - nfull;
      syrk_tile_pipe<RB, true, NB>(tile, lda, Mp, S, stage_w, 8 * NB * t, 8 * NB * t, lane, pd);
    }
  }
  // the last unit drains on its own; `before_drain` lets the caller put requests in flight under it (the operand prologue of
  // the next contraction)
  before_drain();
  if (pd.valid) {
#pragma unroll
    for (int sl = 0; sl < 2 * NB; ++sl) syrk_flush_slot<NB>(pd, sl, Mp, S, stage_w, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// Back-propagate W (BT x Mp, in the shared tile) through K(Xc, Z):
//   WtX~ = W^T [Xc,1]   (Mp x (Din+1))   ->  dJ/dZ rows
//   WZ~  = W [Z,1]      (BT x (Din+1))   ->  dJ/dXc rows
// SE: W = kbar*k;  Linear: W = kbar (the factor v is applied here).
// Both products run on the tensor pipe with every global operand (Z~ fragments, the Z values of the epilogue)
// requested ahead of its use, and all reductions end in fire-and-forget global REDs.
// NBM = number of 8-column blocks of [Xc,1] / [Z,1] = ceil((Din+1)/8), exact: a predicated-off DMMA for an unused block
// would still occupy the tensor pipe.
template <int KIND, int RB, int NGW, int NW, int NBM>
__device__ __forceinline__ void contract_W(const Smem& sm, int lda, const DevProblem& P, int d, double v, double* gXs,
                                           long long doff1, long long doff2, int t0, int nvalid, int warp, int lane, int tid
#ifdef FFVD_PHASE_TIMING
                                           , long long& _phase_last
#endif
                                           ) {
  const int g = lane >> 2, q = lane & 3;
  const int Din = P.Din, M = P.M, D = P.D;
  constexpr int Mp = 16 * NGW * (NW < 8 ? NW : 8);      // compile-time width (see fused_kernel)
  constexpr int nbx = NBM;                      // n-blocks covering Din+1 columns
  constexpr int BT = 8 * RB;
  const double* tile = sm.tile;
  // ---- W^T X~ : 8 warps own m-blocks w, w+8, ...  (NW == 16: warps 8..15, concurrently with W Z~ on warps 0..7)
  constexpr int NCW = NW < 8 ? NW : 8;         // column warps of the tile (Mp = 16 NGW NCW)
  constexpr int NA = NCW;                      // warps that share each of the two products
  if (NW <= 8 || warp >= 8) {
    const int wA = warp % NA;
    constexpr int NMB = 2 * NGW;                 // m-blocks per warp = (Mp/8)/NA
    constexpr int MCH = NMB < 4 ? NMB : 4;       // processed MCH at a time
    // The products are flushed RAW: column jd < Din of row m is sum_t W[t][m] x[t][jd], column Din (the ones) the column sum of
    // W.  Scaling by 1/l^2, the -colsum * z term and the d/dlogl row part are applied once per evaluation by zbar_post_kernel
    // (prep_post.cuh) on the accumulated [D][Mp][8 NBM] array -- in the tile loop that epilogue was 16 loads of Z, a shuffle, 32 FMAs
    // and a second reduction per lane and d.
    double* gzd = det_at(P.gZd + (size_t)d * Mp * (8 * NBM), doff1);
#pragma unroll 1
    for (int i0 = 0; i0 < NMB; i0 += MCH) {
      double c[MCH][NBM][2];
#pragma unroll
      for (int u = 0; u < MCH; ++u)
#pragma unroll
        for (int nb = 0; nb < NBM; ++nb) c[u][nb][0] = c[u][nb][1] = 0.0;
      const double* arow = tile + q * lda + 8 * (wA + NA * i0) + g;
      const double* brow = sm.xs + q * FFVD_XLD + g;
#pragma unroll 2
      for (int k0 = 0; k0 < BT; k0 += 4) {
        double b[NBM];
#pragma unroll
        for (int nb = 0; nb < NBM; ++nb) b[nb] = (nb < nbx) ? brow[k0 * FFVD_XLD + 8 * nb] : 0.0;
#pragma unroll
        for (int u = 0; u < MCH; ++u) {
          const double a = arow[k0 * lda + 8 * NA * u];
#pragma unroll
          for (int nb = 0; nb < NBM; ++nb)
            if (nb < nbx) dmma884(c[u][nb][0], c[u][nb][1], a, b[nb]);
        }
      }
#pragma unroll
      for (int u = 0; u < MCH; ++u) {
        // layout [nb][e][Mp][4] (zbar_index): for a fixed (nb, e) the 32 lanes of a warp (row m = 8 mb + g, column pair q) add to 32
        // consecutive doubles -- 8 full sectors per RED instruction instead of 16 half-used ones in a row-major [Mp][8 NBM] array
        const int m = 8 * (wA + NA * (i0 + u)) + g;
#pragma unroll
        for (int nb = 0; nb < NBM; ++nb)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            red_add_if(gzd + zbar_index(Mp, m, 8 * nb + 2 * q + e), c[u][nb][e], m < M && 8 * nb + 2 * q + e <= Din);
      }
    }
  }
  FFVD_MARK(8);
  // ---- W Z~ : warps 0..7 -> (group of RPW row blocks, k slice).  B fragments come coalesced from the fragment-ordered
  //      Zf, requested PF trips (of 4 k-steps) ahead; each fragment feeds RPW DMMAs.
  constexpr int RPW = RB >= 2 ? 2 : 1;
  constexpr int RG = RB / RPW;                 // row groups
  constexpr int KS = NA / RG;                  // k slices (8 warps: RB = 8: 2, 4: 4, 2: 8, 1: 8)
  constexpr int KLEN = 16 * NCW * NGW / KS;    // columns of W per slice
  constexpr int TRIPS = KLEN / 16;
  constexpr int PF = 2;
  constexpr int PW = 8 * NBM;                  // row stride of part
  static_assert(KS * BT <= 128, "part holds 128 rows");
  if (warp < NA) {
    const int rg = warp % RG, ks = warp / RG;
    double c[RPW][NBM][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int nb = 0; nb < NBM; ++nb) c[r][nb][0] = c[r][nb][1] = 0.0;
    const double* zf = P.Zf + (size_t)(ks * (KLEN / 4)) * 128 + lane;
    const double* ap = tile + (8 * RPW * rg + g) * lda + ks * KLEN + q;
    double b[TRIPS][4][NBM];                   // fully unrolled below: lives in registers by liveness
#pragma unroll
    for (int t = 0; t < PF && t < TRIPS; ++t)
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int nb = 0; nb < NBM; ++nb) b[t][u][nb] = (nb < nbx) ? __ldg(zf + ((t * 4 + u) * 4 + nb) * 32) : 0.0;
#pragma unroll
    for (int t = 0; t < TRIPS; ++t) {
      if (t + PF < TRIPS) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int nb = 0; nb < NBM; ++nb)
            b[t + PF][u][nb] = (nb < nbx) ? __ldg(zf + (((t + PF) * 4 + u) * 4 + nb) * 32) : 0.0;
      }
      double a[RPW][4];
#pragma unroll
      for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u) a[r][u] = ap[r * 8 * lda + 16 * t + 4 * u];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
          for (int nb = 0; nb < NBM; ++nb)
            if (nb < nbx) dmma884(c[r][nb][0], c[r][nb][1], a[r][u], b[t][u][nb]);
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int nb = 0; nb < NBM; ++nb)
        if (nb < nbx) {
          FFVD_ASSERT((ks * BT + 8 * (RPW * rg + r) + g) < 128 && 8 * nb + 2 * q + 2 <= PW);
          *reinterpret_cast<double2*>(sm.part + (size_t)(ks * BT + 8 * (RPW * rg + r) + g) * PW + 8 * nb + 2 * q) =
              make_double2(c[r][nb][0], c[r][nb][1]);
        }
  }
  FFVD_MARK(9);
  __syncthreads();
  FFVD_MARK(10);
  // ---- rows of dJ/dXc: warp -> rows warp, warp+NW, ...; lane -> input column
  {
    double lsum = 0.0, vacc = 0.0;
    const double il2 = (KIND == 0) ? sm.small[lane] : 0.0;
    // all shared-memory reads of the warp's rows first (one latency, not one per row), then the arithmetic and the REDs
    constexpr int NR = BT / NW;
    double wzv[NR], rsv[NR], xv[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = warp + NW * i;
      wzv[i] = rsv[i] = 0.0;
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        if (lane < PW) wzv[i] += sm.part[(size_t)(k * BT + r) * PW + lane];
        rsv[i] += sm.part[(size_t)(k * BT + r) * PW + Din];
      }
      xv[i] = sm.xs[r * FFVD_XLD + lane];
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = warp + NW * i;
      {
        // branch free: rows >= nvalid of W and of the x tile are zero, so their terms vanish by themselves
        const bool in = lane < Din;
        const double x = in ? xv[i] : 0.0, wz = in ? wzv[i] : 0.0, rs = rsv[i];
        double xb;
        if (KIND == 0) {
          xb = -il2 * (x * rs - wz);
          lsum = fma(-x, xb, lsum);                 // d/dlogl row part
          vacc += (lane == 0) ? rs : 0.0;           // d/dlogv = sum W
        } else {
          xb = v * wz;
          vacc = fma(x, xb, vacc);                  // sum kbar*k
        }
        red_add_if(gXs + (size_t)(t0 + r) * D + lane, xb, lane < D && r < nvalid);
      }
    }
    if (KIND == 0 && lane < Din) red_add(det_at(P.gl + (size_t)d * Din + lane, doff2), lsum);
    vacc = warp_sum(vacc);
    if (lane == 0) red_add(det_at(P.gv + d, doff2), vacc);
  }
}

// ---------------------------------------------------------------------------------------------
// NW warps per CTA (8 or 16).  Warp w owns the column groups of "column warp" wc = w & 7 and the
// row blocks [wr*RBW, (wr+1)*RBW) with wr = w >> 3, RBW = RB / (NW/8): 16 warps share one tile and
// double the latency-hiding capacity of the SM without shrinking the tile.
// MINB = CTAs per SM the kernel is built for (register cap 65536 / (32 NW MINB)): two co-resident CTAs let one CTA's
// tensor-pipe phases run under the other's K-tile / statistics / flush phases.
template <int KIND, int RB, int NGW, int MODE, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
fused_kernel(const DevProblem* __restrict__ probs, int nprob, long long total_items, double* __restrict__ kscr_base) {
  extern __shared__ __align__(16) double smem_raw[];
  constexpr int BT = 8 * RB;
  constexpr int NTH = 32 * NW;
  constexpr int NCW = NW < 8 ? NW : 8;        // column warps; NW = 4: a half-width CTA built to run two per SM (MINB = 2)
  constexpr int RBW = RB / (NW / NCW);
  static_assert(RBW >= 1 && RBW * (NW / NCW) == RB, "RB must be divisible by the row split");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  // Column warp of this warp.  With two row halves (NW = 16, NGW = 1) the second half owns the column groups in REVERSE
  // order: a warp's triangular work is proportional to (group + 1), and the four warps of one SM sub-partition (w, w+4,
  // w+8, w+12) then hold groups p, p+4, 7-p, 3-p -- 18 units on every sub-partition instead of 12 ... 24.
  const int wr = warp / NCW;
  const int wc = (wr & 1) ? (NCW - 1 - warp % NCW) : (warp % NCW);
  const int row0 = wr * 8 * RBW;
  double* kscr = kscr_base + (size_t)blockIdx.x * BT * probs[0].Mp;     // per-CTA K-tile scratch (BT x Mp)

  __shared__ DevProblem sP;
  int cur_pi = -1;
#ifdef FFVD_PHASE_TIMING
  long long _phase_last = clock64();
#endif
  for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
    // ---- decode item -> (problem, s, tile, d)
    int pi = 0;
    {
      int lo = 0, hi = nprob - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (probs[mid].item_begin <= item) lo = mid; else hi = mid - 1;
      }
      pi = lo;
    }
    // The problem descriptor lives in shared memory: its ~40 fields are read all over the item, and as global loads
    // each first use per phase was a dependent L2 round trip in front of the phase's own loads.
    if (pi != cur_pi) {
      __syncthreads();          // nobody still reads the previous descriptor
      const int* src = reinterpret_cast<const int*>(probs + pi);
      int* dst = reinterpret_cast<int*>(&sP);
      for (int i = tid; i < (int)(sizeof(DevProblem) / sizeof(int)); i += NTH) dst[i] = src[i];
      cur_pi = pi;
      __syncthreads();
    }
    const DevProblem& P = sP;
    const long long li = item - P.item_begin;
    // Work-item order inside a problem: blocks of `dblk` output dims (slowest), then (sample, tile), then d inside the
    // block.  Two forms (chosen per problem by the host, DevProblem::dl):
    //   dl == 1: one work item per (sample, tile, d), d fastest: the dblk items of one (sample, tile) run on neighbouring
    //            CTAs at the same time, so the x tile is fetched from HBM once per block and its x-bar contributions meet
    //            in L2 (small problems: as many items as possible for the persistent grid);
    //   dl  > 1: one work item per (sample, tile, block of dblk dims): the CTA stages the x tile ONCE and loops over the
    //            dims of the block (large problems: the staging latency -- an HBM round trip plus two barriers, 5% of an
    //            item at C3 -- is paid once per dblk dims).
    // Either way the L^{-1} / L^{-T} operands in flight at any time are those of one block (dblk * 16 Mp^2 bytes, chosen
    // by the host to stay a small part of L2).
    int d0, nd, tile_i, s;
    {
      const int DB = P.dblk;
      if (P.dl > 1) {
        const long long per = (long long)P.S * P.ntiles;
        const long long blk = li / per;
        const long long st = li - blk * per;
        d0 = (int)blk * DB;
        nd = min(DB, P.D - d0);
        tile_i = (int)(st % P.ntiles);
        s = (int)(st / P.ntiles);
      } else {
        const long long per_blk = (long long)P.S * P.ntiles * DB;
        const int nfull = P.D / DB;
        const long long blk = li / per_blk;
        long long rem = li - blk * per_blk;
        int bs = DB, db0 = (int)blk * DB;
        if (blk >= nfull) { rem = li - (long long)nfull * per_blk; bs = P.D - nfull * DB; db0 = nfull * DB; }
        d0 = db0 + (int)(rem % bs);
        nd = 1;
        const long long st = rem / bs;
        tile_i = (int)(st % P.ntiles);
        s = (int)(st / P.ntiles);
      }
    }
    const int T = P.T, D = P.D, Dx = P.Dx, Din = P.Din, M = P.M, nc = P.nc;
    // The padded width is a property of the instantiation (the host launches the template that matches P.Mp): as compile-time
    // constants Mp and the tile stride fold into the immediate offsets of the shared / global accesses -- with a run-time
    // stride every fragment address of the contraction loops was an IMAD of its own (ncu: 0.7 IMAD per DMMA).
    constexpr int Mp = 16 * NGW * NCW;
    constexpr int lda = Mp + 4;
    FFVD_ASSERT(li >= 0 && li < P.nitems && d0 >= 0 && nd >= 1 && d0 + nd <= P.D && tile_i >= 0 && tile_i < P.ntiles && s >= 0 && s < P.S);
    FFVD_ASSERT(P.Mp == Mp);
    const int t0 = tile_i * BT;
    const int nvalid = min(BT, T - t0);
    const double* Xs = P.X + (size_t)s * P.xrows * Dx;
    double* gXs = (MODE == MODE_COND || MODE == MODE_FORWARD) ? nullptr : P.gX + (size_t)s * (T + 1) * D;
    // deterministic mode: the three other classes of x-bar contributions go to their own planes, so that every element of
    // every plane has at most two contributing threads (a two-term sum is order independent); finalize adds the planes
    double* gXb = (gXs && P.gXp) ? P.gXp + (size_t)s * (T + 1) * D : gXs;                        // +-e (uncollapsed) / +-gx (collapsed pass 2)
    double* gXe = (gXs && P.gXp) ? P.gXp + P.gXp_stride + (size_t)s * (T + 1) * D : gXs;         // emission
    double* gXl = (gXs && P.gXp) ? P.gXp + 2 * P.gXp_stride + (size_t)s * (T + 1) * D : gXs;     // LinearK: -v/Q x_t
    const long long doff1 = det_off1(P), doff2 = det_off2(P);       // deterministic mode: this CTA's / warp's private accumulator copies

    Smem sm;
    {
      double* p = smem_raw;
      sm.tile = p; p += (size_t)BT * lda;
      sm.xs = p; p += ((BT + 1) * FFVD_XLD + 1) & ~1;
      sm.xsc = p; sm.part = p; sm.stage = p; p += fused_xsc_part_doubles(BT, NW, (Din + 1 + 7) >> 3);
      sm.us = p; p += Mp;
      sm.ws = p; p += Mp;
      sm.es = p; p += 64;
      sm.rowpart = p; p += 3 * 8 * BT;
      sm.small = p; p += 64;
      sm.sc = p; p += 8;
      sm.red = p; p += 40;
      sm.exptab = p;
    }

    __syncthreads();   // previous item fully done with shared memory
    // ---- P0a: stage the raw x tile, once per work item.  All global loads of the tile are issued before the first one
    //      is consumed (a load -> store loop pays one L2 / HBM latency per trip: ~5k clk per work item at C3).
    {
      constexpr int NIT = ((BT + 1) * FFVD_XCOLS + NTH - 1) / NTH;
      double xv[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int idx = tid + it * NTH;
        const int r = idx >> 5, c = idx & 31;          // c == lane: NTH is a multiple of 32
        const int t = t0 + r;
        const double* src = nullptr;
        if (r <= nvalid && r <= BT && t < P.xrows) {   // rows above nvalid stay zero (row nvalid is x_{t+1} of the last valid row)
          if (c < Dx) src = Xs + (size_t)t * Dx + c;
          else if (c < Din && t < T) src = P.ctrl + (size_t)t * nc + (c - Dx);
        }
        xv[it] = src ? __ldg(src) : 0.0;
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int idx = tid + it * NTH;
        const int r = idx >> 5, c = idx & 31;
        if (r <= BT) sm.xs[r * FFVD_XLD + c] = (c == Din) ? ((r < nvalid) ? 1.0 : 0.0) : xv[it];
      }
    }
  for (int di = 0; di < nd; ++di) {
    const int d = d0 + di;
    const int dh = d * P.hs;
    // ---- P0b: per-d vectors.  The per-evaluation scalars (1/l, 1/l^2, v, Q, ...) and U^T come precomputed from
    //      hyper_kernel; the scaled rows x~ = x/l and -1/2 |x~|^2 are formed from the staged tile.
    const double* hyp = P.hyp + (size_t)dh * 72;
    // requests for this d's vectors go out before the barrier that retires the previous d's use of shared memory
    double sil4[4];                    // 1 / l of the four input columns this thread scales (columns (tid & 7) + 8 c4)
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) sil4[c4] = (KIND == 0 && (tid & 7) + 8 * c4 < Din) ? __ldg(hyp + 32 + (tid & 7) + 8 * c4) : 0.0;
    double hv0 = 0.0, hv1 = 0.0, scv = 0.0;
    if (tid < 32) { hv0 = __ldg(hyp + tid); hv1 = __ldg(hyp + 32 + tid); }
    if (tid == 40) scv = __ldg(hyp + 64);
    if (tid >= 41 && tid < 44) scv = __ldg(P.hq + (size_t)d * 4 + (tid - 41));
    double2 kb0[FFVD_KB0];
    {
      // first operand fragments of the K-tile product (group 0 of this warp): pulled into L1 while the barriers and the
      // per-d staging below run (compute_k_tile requests the later groups one group ahead itself)
      const double* zp = ((KIND == 0) ? P.ZTs + (size_t)dh * FFVD_ZTS_ROWS * Mp : P.ZT) + (size_t)q * Mp + 2 * g + 16 * group_index<NCW>(wc, 0);
      const int ksn = (KIND == 0) ? (Din + 2 + 3) >> 2 : (Din + 3) >> 2;
      for (int ks = FFVD_KB0; ks < ksn; ++ks) prefetch_l1(zp + (size_t)(4 * ks) * Mp);
      // ... and the first FFVD_KB0 k-steps of them as register loads (ncu: with the L1 prefetch alone the phase still began with
      // an exposed L2 round trip -- the loads of group 0 and, behind them on the same scoreboard, the requests for group 1)
#pragma unroll
      for (int ks = 0; ks < FFVD_KB0; ++ks) kb0[ks] = (ks < ksn) ? ldg_nc2_v(zp + (size_t)(4 * ks) * Mp) : make_double2(0.0, 0.0);
    }
    // u_d / w_d (Mp entries each): requested here, stored after the barrier
    constexpr int NUS = (Mp + NTH - 1) / NTH;
    double usv[NUS], wsv[NUS];
#pragma unroll
    for (int i = 0; i < NUS; ++i) {
      const int j = tid + i * NTH;
      usv[i] = wsv[i] = 0.0;
      if (j < Mp) {
        if (MODE == MODE_UNCOLLAPSED || MODE == MODE_FORWARD || MODE == MODE_COND) usv[i] = __ldg(P.UT + (size_t)d * Mp + j);
        else if (MODE == MODE_COLLAPSED_P2) usv[i] = (j < M) ? __ldg(P.wvec + ((size_t)s * D + d) * Mp + j) : 0.0;
        if (MODE == MODE_UNCOLLAPSED) wsv[i] = __ldg(P.wvec + (size_t)d * Mp + j);
        if (MODE == MODE_COND) wsv[i] = (P.qmode == 2 && j < M) ? __ldg(P.qmat + (size_t)j * D + d) : 0.0;
      }
    }
    // SE: the exp table pre-scaled by v_d (threads 64..127).  Only the two REQUESTS go out here: a consumer of a loaded value in
    // front of the barrier below makes the whole CTA wait there for the L2 round trip (the scaling product sat here at first).
    double tab_v = 0.0, tab_e = 0.0;
    if (KIND == 0 && tid >= 64 && tid < 128) { tab_v = __ldg(hyp + 64); tab_e = g_exp2_tab[tid - 64]; }
    __syncthreads();   // x tile staged (di == 0) / previous d fully done with shared memory (di > 0)
    FFVD_MARK(11);
    if (tid < 64) {
      if (tid < 32) { sm.small[tid] = hv0; sm.small[32 + tid] = hv1; }
      if (tid < 40) sm.red[tid] = 0.0;
      if (tid == 40) sm.sc[0] = scv;
      if (tid >= 41 && tid < 44) sm.sc[tid - 40] = scv;
    }
    if (KIND == 0 && tid >= 64 && tid < 128) sm.exptab[tid - 64] = (exp_nmin(tab_v) == 0 && tab_v < 1.0) ? 0.0 : tab_v * tab_e;
    if (KIND == 0) {
      // x~ = x / l and -1/2 |x~_r|^2 in ONE pass: a quarter-warp per row (8 lanes x 4 columns each, the 1/l of the four
      // columns in registers since the top of the iteration), all rows of a pass loaded before the first is used.  Columns
      // Din, Din + 1 of xsc hold the augmentation [-1/2 |x~|^2, 1] of the K-tile product, the columns after them zeros.
      constexpr int RPP = NTH / 8;                       // rows per pass
      constexpr int NPS = (BT + RPP - 1) / RPP;
      const int l8 = tid & 7;
      double xr[NPS][4];
#pragma unroll
      for (int ps = 0; ps < NPS; ++ps) {
        const int r = (tid >> 3) + ps * RPP;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) xr[ps][c4] = (r < BT) ? sm.xs[r * FFVD_XLD + l8 + 8 * c4] : 0.0;
      }
#pragma unroll
      for (int ps = 0; ps < NPS; ++ps) {
        const int r = (tid >> 3) + ps * RPP;
        double a = 0.0;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int c = l8 + 8 * c4;
          const double sv = xr[ps][c4] * sil4[c4];
          a = fma(sv, sv, a);
          if (r < BT && c != Din && c != Din + 1) sm.xsc[r * FFVD_XLD + c] = sv;
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (l8 == 0 && r < BT) {
          sm.xsc[r * FFVD_XLD + Din] = -0.5 * a;
          sm.xsc[r * FFVD_XLD + Din + 1] = 1.0;
        }
      }
      if (Din + 2 > FFVD_XCOLS)
        for (int idx = tid; idx < BT * 4; idx += NTH)
          if (32 + (idx & 3) > Din + 1) sm.xsc[(idx >> 2) * FFVD_XLD + 32 + (idx & 3)] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NUS; ++i) {
      const int j = tid + i * NTH;
      if (j < Mp) {
        sm.us[j] = usv[i];
        if (MODE == MODE_UNCOLLAPSED || MODE == MODE_COND) sm.ws[j] = wsv[i];
      }
    }
    __syncthreads();
    const double v = sm.sc[0], invQ = sm.sc[2], logQd = sm.sc[3];
    FFVD_MARK(0);

    // ---- P1: K tile -> shared (SE uncollapsed: also to this CTA's L2-resident scratch, needed again for W = Kbar o K)
    {
      const double* ZTd = (KIND == 0) ? P.ZTs + (size_t)dh * FFVD_ZTS_ROWS * Mp : P.ZT;
      const int ksn = (KIND == 0) ? (Din + 2 + 3) >> 2 : (Din + 3) >> 2;      // DMMA k-steps of the (augmented) product
      const int nmin = (KIND == 0) ? exp_nmin(v) : 0;
      // SE uncollapsed keeps a copy of the K tile in this CTA's L2 scratch (W = Kbar o K needs it again after the tile has been
      // overwritten by A).  FFVD_SCR_TMA: the copy is made by bulk shared -> global copies after the phase (below) instead of
      // 256-bit stores from the lanes' registers inside it (ncu: the next pass waited for those stores to leave the LSU queue).
      constexpr bool SCR = (KIND == 0 && MODE == MODE_UNCOLLAPSED) && !FFVD_SCR_TMA;
      switch (ksn) {
        case 2: compute_k_tile<KIND, RBW, NGW, SCR, NCW, 2>(sm, ZTd, Mp, Din, v, nmin, ksn, wc, g, q, M, row0, lda, kscr, kb0); break;
        case 3: compute_k_tile<KIND, RBW, NGW, SCR, NCW, 3>(sm, ZTd, Mp, Din, v, nmin, ksn, wc, g, q, M, row0, lda, kscr, kb0); break;
        case 5: compute_k_tile<KIND, RBW, NGW, SCR, NCW, 5>(sm, ZTd, Mp, Din, v, nmin, ksn, wc, g, q, M, row0, lda, kscr, kb0); break;
        default: compute_k_tile<KIND, RBW, NGW, SCR, NCW, 0>(sm, ZTd, Mp, Din, v, nmin, ksn, wc, g, q, M, row0, lda, kscr, kb0); break;
      }
    }
    if (nvalid < BT) {
      // partial last tile of a sample: rows >= nvalid must be inert (exact zeros)
      __syncthreads();
      for (int idx = tid; idx < (BT - nvalid) * Mp; idx += NTH) sm.tile[(nvalid + idx / Mp) * lda + idx % Mp] = 0.0;
    }
    const double* LinvT = P.LinvT + (size_t)dh * Mp * Mp;
    const double* Linv = P.Linv + (size_t)dh * Mp * Mp;
    constexpr int CH0 = tile_gemm_ch0<NGW>();
    double2 ring0[4][CH0];            // first operand fragments of the next contraction, requested ahead of the barrier / SYRK in front of it
    if (MODE != MODE_COLLAPSED_P2) tile_gemm_prologue<CH0, +1, NCW>(ring0, LinvT, Mp, wc, 0, g, q);
    else tile_gemm_prologue<CH0, 0, NCW>(ring0, P.Nmat + ((size_t)s * D + d) * Mp * Mp, Mp, wc, 0, g, q);
    // every thread orders its own K-tile stores (generic proxy) before the async proxy that will read the tile, THEN the barrier
    if (FFVD_SCR_TMA && KIND == 0 && MODE == MODE_UNCOLLAPSED) bulk_fence_shared();
    __syncthreads();
    FFVD_MARK(1);
    if (FFVD_SCR_TMA && KIND == 0 && MODE == MODE_UNCOLLAPSED && warp == 0) {
      // K tile -> L2 scratch, one bulk copy per row (the rows are lda apart in shared memory, Mp apart in the scratch); in
      // flight during the whole first contraction, awaited by the issuing lanes in front of the barrier that ends it
      for (int r = lane; r < BT; r += 32) bulk_copy_s2g(kscr + (size_t)r * Mp, sm.tile + (size_t)r * lda, (unsigned)(Mp * sizeof(double)));
      bulk_commit();
    }

    double acc[NGW][RBW][4];
    double* wtile = sm.tile + (size_t)row0 * lda;     // this warp's rows of the shared tile

    if (MODE != MODE_COLLAPSED_P2) {
      // ---- P2: A = K L^{-T}   (rows a_t = L^{-1} k_t)
      tile_gemm<RBW, NGW, +1, NCW>(acc, ring0, wtile, lda, LinvT, Mp, wc, g, q,
                              [](double x, int, int) { return x; });
      // row partial sums: a.u and a.a
      {
        double su[RBW], sa[RBW], sq[RBW];
#pragma unroll
        for (int rb = 0; rb < RBW; ++rb) su[rb] = sa[rb] = sq[rb] = 0.0;
#pragma unroll
        for (int ng = 0; ng < NGW; ++ng) {
          const int jb = 16 * group_index<NCW>(wc, ng) + 4 * q;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double u = sm.us[jb + c];
            const double qs = (MODE == MODE_COND) ? sm.ws[jb + c] : 0.0;     // q_sqrt given as (M,R) scales
#pragma unroll
            for (int rb = 0; rb < RBW; ++rb) {
              su[rb] = fma(acc[ng][rb][c], u, su[rb]);
              sa[rb] = fma(acc[ng][rb][c], acc[ng][rb][c], sa[rb]);
              if (MODE == MODE_COND) {
                const double t = acc[ng][rb][c] * qs;
                sq[rb] = fma(t, t, sq[rb]);
              }
            }
          }
        }
#pragma unroll
        for (int rb = 0; rb < RBW; ++rb) {
          su[rb] += __shfl_xor_sync(0xffffffffu, su[rb], 1);
          su[rb] += __shfl_xor_sync(0xffffffffu, su[rb], 2);
          sa[rb] += __shfl_xor_sync(0xffffffffu, sa[rb], 1);
          sa[rb] += __shfl_xor_sync(0xffffffffu, sa[rb], 2);
          if (MODE == MODE_COND) {
            sq[rb] += __shfl_xor_sync(0xffffffffu, sq[rb], 1);
            sq[rb] += __shfl_xor_sync(0xffffffffu, sq[rb], 2);
          }
          if (q == 0) {
            sm.rowpart[(0 * 8 + wc) * BT + row0 + 8 * rb + g] = su[rb];
            sm.rowpart[(1 * 8 + wc) * BT + row0 + 8 * rb + g] = sa[rb];
            if (MODE == MODE_COND) sm.rowpart[(2 * 8 + wc) * BT + row0 + 8 * rb + g] = sq[rb];
          }
        }
      }
      if (FFVD_SCR_TMA && KIND == 0 && MODE == MODE_UNCOLLAPSED && warp == 0) bulk_wait_all();     // scratch copy complete (long since)
      __syncthreads();          // everyone is done reading K from the tile
      FFVD_MARK(2);
      if (MODE != MODE_FORWARD && MODE != MODE_COND) store_tile<RBW, NGW, NCW>(acc, wtile, lda, wc, g, q);
      FFVD_MARK(12);
      // ---- per-row statistics (threads 0..BT-1)
      if (tid < 64) {
        double jxq = 0.0, jtr = 0.0, gq = 0.0, gvd = 0.0;
        if (tid < BT) {
          const int r = tid;
          double e = 0.0;
          if (r < nvalid) {
            double su = 0.0, sa = 0.0;
#pragma unroll
            for (int w = 0; w < NCW; ++w) {
              su += sm.rowpart[(0 * 8 + w) * BT + r];
              sa += sm.rowpart[(1 * 8 + w) * BT + r];
            }
            const double xd = sm.xs[r * FFVD_XLD + d], xn = sm.xs[(r + 1) * FFVD_XLD + d];
            double kdiag = v;
            if (KIND == 1) {
              double ss = 0.0;
              for (int c = 0; c < Din; ++c) ss = fma(sm.xs[r * FFVD_XLD + c], sm.xs[r * FFVD_XLD + c], ss);
              kdiag = v * ss;
            }
            const double sig2 = kdiag - sa;
            if (MODE == MODE_COND) {
              // conditionals_multi_output.py:41,48 : fvar = Knn - sum A^2 ; fmean = A^T f ; :50-52 q_sqrt (M,R) term
              double sq = 0.0;
#pragma unroll
              for (int w = 0; w < NCW; ++w) sq += sm.rowpart[(2 * 8 + w) * BT + r];
              // REUSE_KZZ content guard (hyper_kernel): factors that do not belong to this Z poison the result
              const bool stale = P.guard && P.guard[3] != 0;
              P.cond_mean[(size_t)(t0 + r) * D + d] = stale ? __longlong_as_double(0x7ff8000000000000ll) : su;
              P.cond_var[(size_t)(t0 + r) * D + d] = sig2 + sq;
            } else if (MODE == MODE_COLLAPSED_P1) {
              const double delta = xn - xd;
              e = delta;                                   // b += F^T delta
              jxq = -0.5 * delta * delta * invQ - 0.5 * logQd;
              jtr = -0.5 * sig2 * invQ;
              gq = 0.5 * sig2 * invQ + 0.5 * delta * delta * invQ - 0.5;   // explicit part of dJ/dlogQ
            } else {
              const double res = xn - (xd + su);
              e = res * invQ;
              jxq = -0.5 * res * res * invQ - 0.5 * logQd;
              jtr = -0.5 * sig2 * invQ;
              gq = 0.5 * res * res * invQ - 0.5 + 0.5 * sig2 * invQ;
              if (MODE == MODE_UNCOLLAPSED) {
                FFVD_ASSERT(t0 + r + 1 <= T && d < D);
                red_add(gXb + (size_t)(t0 + r) * D + d, e);              // (element [t][d] gets exactly two terms: +e_t, -e_{t-1})
                red_add(gXb + (size_t)(t0 + r + 1) * D + d, -e);
              }
            }
            gvd = -0.5 * kdiag * invQ;
            if (KIND == 1 && MODE == MODE_UNCOLLAPSED) {
              for (int c = 0; c < D; ++c) red_add(gXl + (size_t)(t0 + r) * D + c, -v * invQ * sm.xs[r * FFVD_XLD + c]);
            }
          }
          sm.es[r] = e;
        }
        jxq = warp_sum(jxq); jtr = warp_sum(jtr); gq = warp_sum(gq); gvd = warp_sum(gvd);
        if (lane == 0) {
          double* rw = sm.red + 4 * warp;      // warps 0 and 1 only: private slots, summed at the flush
          rw[0] = jxq; rw[1] = jtr; rw[2] = gq; rw[3] = gvd;
        }
      }
      FFVD_MARK(15);
      // ---- emission term, once per (s, tile): dgp_model.py:248-250,264
      if (d == 0 && MODE != MODE_COLLAPSED_P2 && MODE != MODE_COND && warp >= 2 && warp < 2 + (BT + 31) / 32) {
        const int r = (warp - 2) * 32 + lane;
        const int Dy = P.Dy;
        double ll = 0.0;
        for (int y = 0; y < Dy; ++y) {
          // all requests first: column y of C rides in the lanes (lane c holds C[c][y], D <= 31) and is broadcast by shuffles --
          // as loads inside the c loops this block was ~7k clk once per work item, with every other warp of the CTA waiting
          const bool valid = r < nvalid;
          const double lR = __ldg(P.logR + y), dv = __ldg(P.dvec + y);
          const double Cl = (lane < D) ? __ldg(P.C + (size_t)lane * Dy + y) : 0.0;
          const double Yv = valid ? __ldg(P.Y + (size_t)(t0 + r) * Dy + y) : 0.0;
          const double Ry = exp(lR);
          double yhat = dv;
          for (int c = 0; c < D; ++c) yhat = fma(sm.xs[(r + 1) * FFVD_XLD + c], __shfl_sync(0xffffffffu, Cl, c), yhat);
          const double res = valid ? (Yv - yhat) / Ry : 0.0;
          if (valid) ll += -0.5 * res * res - lR;
          const double dy = res / Ry, rr = valid ? res * res - 1.0 : 0.0;
          if (MODE != MODE_FORWARD) {
            for (int c = 0; c < D; ++c) red_add_if(gXe + (size_t)(t0 + r + 1) * D + c, dy * __shfl_sync(0xffffffffu, Cl, c), valid);
            // dJ/dC column: four warp sums at a time (interleaved shuffle rounds)
            for (int c0 = 0; c0 < D; c0 += 4) {
              double t[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) t[k] = (valid && c0 + k < D) ? dy * sm.xs[(r + 1) * FFVD_XLD + c0 + k] : 0.0;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int k = 0; k < 4; ++k) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
#pragma unroll
              for (int k = 0; k < 4; ++k) red_add_if(det_at(P.gC + (size_t)(c0 + k) * Dy + y, doff2), t[k], lane == 0 && c0 + k < D);
            }
            const double sd = warp_sum(dy), sr = warp_sum(rr);
            if (lane == 0) { red_add(det_at(P.gd + y, doff2), sd); red_add(det_at(P.gR + y, doff2), sr); }
          }
        }
        ll = warp_sum(ll);
        if (lane == 0) red_add(det_at(P.terms_raw + (size_t)s * FFVD_NTERMS_RAW + FFVD_RAW_EMIS, doff2), ll);
      }
      __syncthreads();          // A tile + es[] visible
      FFVD_MARK(3);
    }

    if (MODE == MODE_COND && P.qmode == 3) {
      // conditionals_multi_output.py:53-62 (q_sqrt R x M x M): fvar += sum_m (Q^T a_t)_m^2, a second contraction of the A
      // tile (still in registers) with the dense zero-padded factor
      tile_gemm_prologue<CH0, 0, NCW>(ring0, P.qmat + (size_t)(P.nq == 1 ? 0 : d) * Mp * Mp, Mp, wc, 0, g, q);
      store_tile<RBW, NGW, NCW>(acc, wtile, lda, wc, g, q);
      __syncthreads();
      tile_gemm<RBW, NGW, 0, NCW>(acc, ring0, wtile, lda, P.qmat + (size_t)(P.nq == 1 ? 0 : d) * Mp * Mp, Mp, wc, g, q,
                             [](double x, int, int) { return x; });
      double sq[RBW];
#pragma unroll
      for (int rb = 0; rb < RBW; ++rb) sq[rb] = 0.0;
#pragma unroll
      for (int ng = 0; ng < NGW; ++ng)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int rb = 0; rb < RBW; ++rb) sq[rb] = fma(acc[ng][rb][c], acc[ng][rb][c], sq[rb]);
#pragma unroll
      for (int rb = 0; rb < RBW; ++rb) {
        sq[rb] += __shfl_xor_sync(0xffffffffu, sq[rb], 1);
        sq[rb] += __shfl_xor_sync(0xffffffffu, sq[rb], 2);
        if (q == 0) sm.rowpart[(2 * 8 + wc) * BT + row0 + 8 * rb + g] = sq[rb];
      }
      __syncthreads();
      if (tid < nvalid) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NCW; ++w) t += sm.rowpart[(2 * 8 + w) * BT + tid];
        P.cond_var[(size_t)(t0 + tid) * D + d] += t;      // written by this same thread above
      }
    }
    if (MODE == MODE_FORWARD || MODE == MODE_COND) {
      // forward only: flush the scalar sums and move on
      if (MODE == MODE_FORWARD && tid == 0) {
        red_add(det_at(P.terms_raw + (size_t)s * FFVD_NTERMS_RAW + FFVD_RAW_XQ, doff2), sm.red[0] + sm.red[4]);
        red_add(det_at(P.terms_raw + (size_t)s * FFVD_NTERMS_RAW + FFVD_RAW_TRACE, doff2), sm.red[1] + sm.red[5]);
      }
      continue;
    }

    if (MODE == MODE_UNCOLLAPSED || MODE == MODE_COLLAPSED_P1) {
      // ---- ubar_j = sum_r e_r a_rj  (collapsed pass 1: b_j = sum_r delta_r f_rj), from registers
      {
        double er[RBW];
#pragma unroll
        for (int rb = 0; rb < RBW; ++rb) er[rb] = sm.es[row0 + 8 * rb + g];
        // four columns at a time: the partial sums first, then the three shuffle rounds with the four chains interleaved (one
        // column after the other is a dependent shuffle -> add chain of ~100 clk per column with nothing to overlap it)
        double* ub = det_at(P.ubar + ((size_t)(MODE == MODE_COLLAPSED_P1 ? s * D + d : d)) * Mp, doff2);
#pragma unroll
        for (int ng = 0; ng < NGW; ++ng) {
          const int jb = 16 * group_index<NCW>(wc, ng) + 4 * q;
          double t[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            t[c] = 0.0;
#pragma unroll
            for (int rb = 0; rb < RBW; ++rb) t[c] = fma(er[rb], acc[ng][rb][c], t[c]);
          }
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
            for (int c = 0; c < 4; ++c) t[c] += __shfl_xor_sync(0xffffffffu, t[c], o);
          // lanes g = 0..3 of each q flush one column each (all four hold the full sums)
          red_add_if(ub + jb + g, g == 0 ? t[0] : (g == 1 ? t[1] : (g == 2 ? t[2] : t[3])), g < 4 && jb + g < M);
        }
      }
      // operand prologue of the next contraction (Kbar = A L^{-1}): in flight during the SYRK
#if FFVD_G2_EARLY
      if (MODE == MODE_UNCOLLAPSED) tile_gemm_prologue<CH0, -1, NCW>(ring0, Linv, Mp, wc, 0, g, q);
#endif
      // ---- S += A^T A
      double* Sd = det_at(P.Sacc + ((size_t)(MODE == MODE_COLLAPSED_P1 ? s * D + d : d)) * Mp * Mp, doff1);
#if FFVD_ABLATE != 2
      // unit size fixed at compile time (Mp = 16 NGW NCW): one SYRK variant per instantiation keeps the item loop's code small
#if FFVD_SYRK_PIPE
      syrk_units_pipe<RB, NW, (16 * NGW * NCW <= 128) ? 2 : 4>(sm.tile, lda, Mp, Sd, sm.stage + warp * 2 * 8 * 40, warp, lane, [&]() {
        // (only where the prologue is small: at NGW = 4 its 64 registers on top of the draining unit cost more than they hide)
        if (MODE == MODE_UNCOLLAPSED && !FFVD_G2_EARLY && NGW <= 2) tile_gemm_prologue<CH0, -1, NCW>(ring0, Linv, Mp, wc, 0, g, q);
      });
#else
      syrk_units<RB, NW, (16 * NGW * NCW <= 128) ? 2 : 4>(sm.tile, lda, Mp, Sd, sm.stage + warp * 8 * 40, warp, lane);
#endif
#endif
      FFVD_MARK(4);             // warp 0's own time: no barrier between the SYRK and the next contraction
    }

    if (MODE == MODE_UNCOLLAPSED) {
      // ---- P4: Kbar = Abar L^{-1} with abar_r = e_r u + a_r / Q.  By linearity Kbar = (A L^{-1})/Q + e w^T with
      //      w = L^{-T} u (ltu_kernel, once per evaluation), so the contraction reads A unmodified.
      if (!FFVD_G2_EARLY && (!FFVD_SYRK_PIPE || FFVD_ABLATE == 2 || NGW > 2)) tile_gemm_prologue<CH0, -1, NCW>(ring0, Linv, Mp, wc, 0, g, q);
      tile_gemm<RBW, NGW, -1, NCW>(acc, ring0, wtile, lda, Linv, Mp, wc, g, q,
                              [](double x, int, int) { return x; });
      double er[RBW];
#pragma unroll
      for (int rb = 0; rb < RBW; ++rb) er[rb] = sm.es[row0 + 8 * rb + g];
#pragma unroll
      for (int ng = 0; ng < NGW; ++ng) {
        const int jb = 16 * group_index<NCW>(wc, ng) + 4 * q;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double w = sm.ws[jb + c];
#pragma unroll
          for (int rb = 0; rb < RBW; ++rb) acc[ng][rb][c] = fma(acc[ng][rb][c], invQ, er[rb] * w);
        }
      }
    } else if (MODE == MODE_COLLAPSED_P2) {
      // ---- Kbar = K N + delta w'^T ; also dbar_r = k_r . w'
      tile_gemm<RBW, NGW, 0, NCW>(acc, ring0, wtile, lda, P.Nmat + ((size_t)s * D + d) * Mp * Mp, Mp, wc, g, q,
                             [](double x, int, int) { return x; });
    }

    FFVD_MARK(5);
    if (MODE == MODE_UNCOLLAPSED || MODE == MODE_COLLAPSED_P2) {
      if (MODE == MODE_COLLAPSED_P2) {
        // delta_r and dbar_r = sum_j k_rj w'_j (from the K tile still in shared memory)
        for (int r = warp; r < BT; r += NW) {
          double t = 0.0;
          for (int j = lane; j < Mp; j += 32) t = fma(sm.tile[r * lda + j], sm.us[j], t);
          t = warp_sum(t);
          if (lane == 0) {
            double delta = 0.0;
            if (r < nvalid) {
              delta = sm.xs[(r + 1) * FFVD_XLD + d] - sm.xs[r * FFVD_XLD + d];
              const double gx = t - delta * invQ;
              red_add(gXb + (size_t)(t0 + r + 1) * D + d, gx);
              red_add(gXb + (size_t)(t0 + r) * D + d, -gx);
              if (KIND == 1)
                for (int c = 0; c < D; ++c) red_add(gXl + (size_t)(t0 + r) * D + c, -v * invQ * sm.xs[r * FFVD_XLD + c]);
            }
            sm.es[r] = delta;
          }
        }
        if (tid < 64) {
          // Kdiag part of d/dlogv
          double gvd = 0.0;
          if (tid < nvalid) {
            double kdiag = v;
            if (KIND == 1) {
              double ss = 0.0;
              for (int c = 0; c < Din; ++c) ss = fma(sm.xs[tid * FFVD_XLD + c], sm.xs[tid * FFVD_XLD + c], ss);
              kdiag = v * ss;
            }
            gvd = -0.5 * kdiag * invQ;
          }
          gvd = warp_sum(gvd);
          if (lane == 0) sm.red[8 + warp] = gvd;
        }
        __syncthreads();
      }
      // ---- P6: W = Kbar o K (SE) or Kbar (Linear) -> shared tile
      {
#pragma unroll
        for (int ng = 0; ng < NGW; ++ng) {
          const int jb = 16 * group_index<NCW>(wc, ng) + 4 * q;
          double kv[RBW][4];
          if (KIND == 0) {
            if (MODE == MODE_COLLAPSED_P2) {
#pragma unroll
              for (int rb = 0; rb < RBW; ++rb) {
                const double* p = wtile + (8 * rb + g) * lda + jb;
                ld_tile4(p, g, kv[rb][0], kv[rb][1], kv[rb][2], kv[rb][3]);
              }
            } else {
#pragma unroll
              for (int rb = 0; rb < RBW; ++rb) {
                FFVD_ASSERT(row0 + 8 * rb + g < BT && jb + 4 <= Mp);
                ldg256_cg(kscr + (size_t)(row0 + 8 * rb + g) * Mp + jb, kv[rb][0], kv[rb][1], kv[rb][2], kv[rb][3]);
              }
            }
          }
#pragma unroll
          for (int rb = 0; rb < RBW; ++rb)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              double kb = acc[ng][rb][c];
              if (MODE == MODE_COLLAPSED_P2) kb = fma(sm.es[row0 + 8 * rb + g], sm.us[jb + c], kb);
              acc[ng][rb][c] = (KIND == 0) ? kb * kv[rb][c] : ((jb + c < M && row0 + 8 * rb + g < nvalid) ? kb : 0.0);
            }
        }
        FFVD_MARK(13);
        __syncthreads();        // all warps done reading A (GEMM2 / SYRK) or K (pass 2)
        FFVD_MARK(14);
        store_tile<RBW, NGW, NCW>(acc, wtile, lda, wc, g, q);
      }
      __syncthreads();
      FFVD_MARK(6);
#ifdef FFVD_PHASE_TIMING
#define FFVD_CW_TAIL , _phase_last
#else
#define FFVD_CW_TAIL
#endif
      switch ((Din + 1 + 7) >> 3) {
        case 1: contract_W<KIND, RB, NGW, NW, 1>(sm, lda, P, d, v, gXs, doff1, doff2, t0, nvalid, warp, lane, tid FFVD_CW_TAIL); break;
        case 2: contract_W<KIND, RB, NGW, NW, 2>(sm, lda, P, d, v, gXs, doff1, doff2, t0, nvalid, warp, lane, tid FFVD_CW_TAIL); break;
        case 3: contract_W<KIND, RB, NGW, NW, 3>(sm, lda, P, d, v, gXs, doff1, doff2, t0, nvalid, warp, lane, tid FFVD_CW_TAIL); break;
        default: contract_W<KIND, RB, NGW, NW, 4>(sm, lda, P, d, v, gXs, doff1, doff2, t0, nvalid, warp, lane, tid FFVD_CW_TAIL); break;
      }
#undef FFVD_CW_TAIL
    }

    // ---- flush block-level scalars
    __syncthreads();
    FFVD_MARK(7);
    if (tid < 32) {
      if (tid == 0) {
        if (MODE != MODE_COLLAPSED_P2) {
          red_add(det_at(P.terms_raw + (size_t)s * FFVD_NTERMS_RAW + FFVD_RAW_XQ, doff2), sm.red[0] + sm.red[4]);
          red_add(det_at(P.terms_raw + (size_t)s * FFVD_NTERMS_RAW + FFVD_RAW_TRACE, doff2), sm.red[1] + sm.red[5]);
          red_add(det_at(P.gQ + d, doff2), sm.red[2] + sm.red[6]);
        }
        if (MODE != MODE_COLLAPSED_P1) red_add(det_at(P.gv + d, doff2), (sm.red[3] + sm.red[7]) + (sm.red[8] + sm.red[9]));
      }
    }
  }   // d loop
  }   // item loop
}

}  // namespace ffvd
