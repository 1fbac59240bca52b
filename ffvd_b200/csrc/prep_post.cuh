// Once-per-evaluation kernels around the fused tile kernel (sm_100a):
//   kzz_prep      K(Z,Z)+jitter I -> Cholesky -> L^{-1}, L^{-T}   (conditionals_multi_output.py:124-169)
//   bgemm_nn      batched FP64 DMMA GEMM for the O(M^3) products of the Cholesky backward
//   symmetrize / wz / kzz_bwd / finalize   hand-derived backward through chol(Kzz) and the priors
//   collapsed_solve  H = F^T F / Q + I, logdet, c = H^{-1} b   (conditionals_multi_output.py:246-254)
#pragma once
#include "ffvd_common.cuh"

namespace ffvd {

#ifdef FFVD_PREP_DEBUG
__device__ long long g_prep_dbg[256];
#define FFVD_DBG(i) do { if (threadIdx.x == 0) g_prep_dbg[i] = clock64(); } while (0)
#else
#define FFVD_DBG(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// device helpers operating on a generic-address matrix (shared or global)
// In-place lower Cholesky of the leading M x M block of A (ld = lda); colbuf: M doubles of smem.
// Returns 0 or the 1-based index of the first non-positive pivot (uniform across the block).
__device__ int chol_inplace(double* A, int lda, int M, double* colbuf, int* flag) {
  const int tid = threadIdx.x, nth = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nth >> 5;
  if (tid == 0) *flag = 0;
  for (int j = 0; j < M; ++j) {
    __syncthreads();
    const double piv = A[(size_t)j * lda + j];
    if (!(piv > 0.0)) {
      if (tid == 0) *flag = j + 1;
      break;
    }
    const double ljj = sqrt(piv);
    const double inv = 1.0 / ljj;
    for (int i = j + 1 + tid; i < M; i += nth) {
      const double x = A[(size_t)i * lda + j] * inv;
      A[(size_t)i * lda + j] = x;
      colbuf[i] = x;
    }
    __syncthreads();
    if (tid == 0) A[(size_t)j * lda + j] = ljj;
    for (int i = j + 1 + warp; i < M; i += nw) {
      const double aij = colbuf[i];
      double* row = A + (size_t)i * lda;
      for (int k = j + 1 + lane; k <= i; k += 32) row[k] = fma(-aij, colbuf[k], row[k]);
    }
  }
  __syncthreads();
  return *flag;
}

// X = L^{-1} (lower) by forward substitution, one column per thread; also XT = X^T.
// L is M x M lower (ld = ldl); X, XT are Mp x Mp row-major global buffers whose padding is
// already zero.  rowbuf: M doubles of shared memory.
__device__ void tri_inverse(const double* L, int ldl, int M, double* X, double* XT, int Mp, double* rowbuf) {
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int i = 0; i < M; ++i) {
    __syncthreads();
    for (int k = tid; k <= i; k += nth) rowbuf[k] = L[(size_t)i * ldl + k];
    __syncthreads();
    const double dinv = 1.0 / rowbuf[i];
    for (int j = i + 1 + tid; j < M; j += nth) X[(size_t)i * Mp + j] = 0.0;   // the buffer may hold a stale dense matrix
    for (int j = tid; j <= i; j += nth) {
      double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = j;
      for (; k + 3 < i; k += 4) {
        s0 = fma(-rowbuf[k], X[(size_t)k * Mp + j], s0);
        s1 = fma(-rowbuf[k + 1], X[(size_t)(k + 1) * Mp + j], s1);
        s2 = fma(-rowbuf[k + 2], X[(size_t)(k + 2) * Mp + j], s2);
        s3 = fma(-rowbuf[k + 3], X[(size_t)(k + 3) * Mp + j], s3);
      }
      for (; k < i; ++k) s0 = fma(-rowbuf[k], X[(size_t)k * Mp + j], s0);
      X[(size_t)i * Mp + j] = ((s0 + s1) + (s2 + s3)) * dinv;
    }
  }
  __syncthreads();
  // transpose (XT may alias the buffer that held L: every entry of the M x M block is rewritten)
  for (int idx = tid; idx < M * M; idx += nth) {
    const int r = idx / M, c = idx % M;
    XT[(size_t)r * Mp + c] = (c >= r) ? X[(size_t)c * Mp + r] : 0.0;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Fast path for matrices small enough that L AND its inverse fit in shared memory (M <= FFVD_FAST_MAXM).
//
// Why not the generic routines above: at M ~ 100 they are pure latency.  Every trip of their inner loops is a dependent
// load -> FMA (-> store) of ~70 clocks that the compiler cannot pipeline (it must assume the stores alias the loads), and
// lanes with different trip counts wait for the longest one, so one column / row step costs 1.2-1.6k clocks whatever its
// arithmetic (measured on B200: Cholesky 163k, inverse 197k clocks at M = 100).  Both routines below keep the values
// that are updated in REGISTERS with compile-time indices; shared memory only carries the one column (Cholesky) or the
// read-only factor (inverse) of the current step, and every load of a step is independent of the step's arithmetic.
#define FFVD_FAST_MAXM 119        // 4x4 register tiles of the lower triangle <= 512 threads; 8 columns x 16 warps; offsets r < 120

// Column j = 4J + JJ of chol_owner_smem (below).  Returns false when the CTA must stop: past the last column, or a
// non-positive pivot (flag set).  CTA uniform.
template <int JJ>
__device__ __forceinline__ bool chol_col_step(double (&a)[4][4], double* A, int lda, int M, double* cb, int cbs, int* flag, int J, int tk, int i0,
                                              int k0) {
  const int j = 4 * J + JJ;
  if (j >= M) return false;
  FFVD_DBG(128 + j);
  const double* cur = cb + (JJ & 1) * cbs;
  double* nxt = cb + ((JJ + 1) & 1) * cbs;
  const double piv = cur[j];
  if (!(piv > 0.0)) {                          // same value in every thread: the whole CTA leaves together
    if (threadIdx.x == 0) *flag = j + 1;
    return false;
  }
  // the tile still has a column > j  <=>  tk > J, or tk == J and JJ < 3
  if (tk > J || (tk == J && JJ < 3)) {
    const double inv2 = 1.0 / piv;
    const double2 ci01 = *reinterpret_cast<const double2*>(cur + i0), ci23 = *reinterpret_cast<const double2*>(cur + i0 + 2);
    const double2 ck01 = *reinterpret_cast<const double2*>(cur + k0), ck23 = *reinterpret_cast<const double2*>(cur + k0 + 2);
    const double ci[4] = {-ci01.x * inv2, -ci01.y * inv2, -ci23.x * inv2, -ci23.y * inv2};
    const double ck[4] = {ck01.x, ck01.y, ck23.x, ck23.y};
    // columns <= j of the tile are finished (their final values were parked in A when they were published) and are
    // updated along unconditionally -- garbage from here on, never read again
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r][c] = fma(ci[r], ck[c], a[r][c]);
    constexpr int cn = (JJ + 1) & 3;           // tile column that holds matrix column j + 1 ...
    const bool mine = tk == J + (JJ == 3 ? 1 : 0);   // ... in the tiles of this tile column
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (mine && i0 + r < M && i0 + r >= k0 + cn) {   // lower part only (a diagonal tile carries unused upper entries)
        nxt[i0 + r] = a[r][cn];
        A[(size_t)(i0 + r) * lda + k0 + cn] = a[r][cn];
      }
  }
  __syncthreads();
  return true;
}

// In-place Cholesky of the lower triangle of A (M x M, row stride lda, SHARED memory), owner computes: thread t keeps
// one 4x4 tile of the lower triangle (tiles numbered column-major, so finished tile columns are a prefix of the thread
// order and whole warps retire) in registers for the whole factorisation.
// Step j: every active thread reads the rows / columns of column j of the trailing matrix that its tile needs (published
// by their owners at the end of step j-1, double buffered in cb[2][cbs]), forms 1/pivot itself and applies the rank-1
// update a_ik -= a_ij a_kj / a_jj; the owners of column j+1 publish theirs and park its final (unscaled) values in A.
// One CTA barrier per column, 8 shared loads per 16 multiply-adds.  The column loop is unrolled by the tile width so
// that "which of my columns is the next pivot column" is a compile-time register index (the body is a single-warp
// latency chain: every select / compare removed from it is time).  The scaling by 1/sqrt(a_jj) is applied once at the
// end (L_ik = a_ik / sqrt(a_kk), a the value when column k was reached).
// Returns 0 or the 1-based failing pivot (like chol_inplace).
// blockDim.x == 512, M <= FFVD_FAST_MAXM, cbs a multiple of 4 >= M, cb 16-byte aligned.
__device__ int chol_owner_smem(double* A, int lda, int M, double* cb, int cbs, int* flag) {
  const int tid = threadIdx.x;
  const int nt = (M + 3) >> 2, ntile = nt * (nt + 1) / 2;
  if (tid == 0) *flag = 0;
  // column-major tile index -> (tk, ti): tile columns 0..tk-1 hold tk*nt - tk(tk-1)/2 tiles
  const float fm = (float)nt + 0.5f;
  int tk = (int)(fm - sqrtf(fmaxf(fm * fm - 2.0f * (float)tid, 0.0f)));
  tk = max(0, min(tk, nt - 1));
  while (tk > 0 && tk * nt - tk * (tk - 1) / 2 > tid) --tk;
  while (tk < nt - 1 && (tk + 1) * nt - (tk + 1) * tk / 2 <= tid) ++tk;
  const bool valid = tid < ntile;
  const int i0 = valid ? 4 * (tk + (tid - (tk * nt - tk * (tk - 1) / 2))) : 0;
  if (!valid) tk = -2;                         // never active, never publishes
  const int k0 = 4 * tk;
  double a[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + r, k = k0 + c;
      a[r][c] = (valid && i < M && k <= i) ? A[(size_t)i * lda + k] : 0.0;
    }
  if (tk == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (i0 + r < M) cb[i0 + r] = a[r][0];
  }
  __syncthreads();
  for (int J = 0; J < nt; ++J) {
    if (!chol_col_step<0>(a, A, lda, M, cb, cbs, flag, J, tk, i0, k0)) break;
    if (!chol_col_step<1>(a, A, lda, M, cb, cbs, flag, J, tk, i0, k0)) break;
    if (!chol_col_step<2>(a, A, lda, M, cb, cbs, flag, J, tk, i0, k0)) break;
    if (!chol_col_step<3>(a, A, lda, M, cb, cbs, flag, J, tk, i0, k0)) break;
  }
  __syncthreads();                            // the early exit above skips the barrier of its iteration
  const int st = *flag;
  if (st == 0) {
    // A holds the unscaled columns (column k as it was when it became the pivot column; its diagonal entry is the pivot)
    double sc[4], v[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int k = min(max(k0 + c, 0), M - 1);
      sc[c] = rsqrt(A[(size_t)k * lda + k]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = min(i0 + r, M - 1);
        v[r][c] = A[(size_t)i * lda + k];
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + r, k = k0 + c;
        if (valid && i < M && k < M && k <= i) A[(size_t)i * lda + k] = v[r][c] * sc[c];
      }
  }
  __syncthreads();
  return st;
}

// Panel form of chol_owner_smem: four columns per step, two CTA barriers per step.
//   phase A  the owners of tile column J (their registers ARE the panel, column block J of the trailing matrix) read
//            the raw 4x4 diagonal block, factorise it (each of them, redundantly: ~M/4 threads), turn their own rows
//            into rows of L by a 4-step forward substitution (L_i = P_i L_DD^{-T}), store them into A (final, no
//            deferred scaling) and publish them in lb;
//   phase B  every tile right of the panel reads its L_i and L_k rows from lb and applies the rank-4 update
//            a -= L_i L_k^T (16 shared loads per 64 multiply-adds); the next diagonal tile publishes its raw block.
// ceil(M/4) steps instead of M: the per-column chain (barrier, pivot load, reciprocal, publish) is paid once per four
// columns, and nothing but the tiny diagonal factorisation is redundant (a first version in which every thread solved
// for the L rows it needed was FP64-throughput bound and no faster than the column form).
// Padded rows / columns (M not a multiple of 4) carry an identity diagonal.  lb holds the L rows per 4-row group, 16
// values plus 2 doubles of padding: neighbouring lanes own neighbouring row groups, and a group stride of 9 x 16 bytes
// (odd) spreads the 16-byte accesses of eight lanes over all banks.  scratch: FFVD_PTILE * (ceil(M/4) + 1) doubles,
// 16-byte aligned.  Returns 0 or the 1-based failing pivot.  blockDim.x == 512, M <= FFVD_FAST_MAXM.
#define FFVD_PTILE 18
__device__ int chol_panel_smem(double* A, int lda, int M, double* scratch, int* flag) {
  const int tid = threadIdx.x;
  const int nt = (M + 3) >> 2, ntile = nt * (nt + 1) / 2;
  double* db = scratch;                        // raw diagonal block of the current step (row r at db + 4 r)
  double* lb = scratch + FFVD_PTILE;           // L rows of the current panel, group g at lb + FFVD_PTILE * g
  if (tid == 0) *flag = 0;
  const float fm = (float)nt + 0.5f;
  int tk = (int)(fm - sqrtf(fmaxf(fm * fm - 2.0f * (float)tid, 0.0f)));
  tk = max(0, min(tk, nt - 1));
  while (tk > 0 && tk * nt - tk * (tk - 1) / 2 > tid) --tk;
  while (tk < nt - 1 && (tk + 1) * nt - (tk + 1) * tk / 2 <= tid) ++tk;
  const bool valid = tid < ntile;
  const int ti = valid ? tk + (tid - (tk * nt - tk * (tk - 1) / 2)) : 0;
  const int i0 = 4 * ti;
  if (!valid) tk = -2;                         // never active
  const int k0 = 4 * tk;
  double a[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + r, k = k0 + c;
      a[r][c] = (valid && i < M && k <= i) ? A[(size_t)i * lda + k] : ((valid && i == k) ? 1.0 : 0.0);
    }
  if (tk == 0 && ti == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      *reinterpret_cast<double2*>(db + 4 * r) = make_double2(a[r][0], a[r][1]);
      *reinterpret_cast<double2*>(db + 4 * r + 2) = make_double2(a[r][2], a[r][3]);
    }
  }
  __syncthreads();
  for (int J = 0; J < nt; ++J) {
    if (tk == J) {
      // 4x4 diagonal block (lower part) -> its Cholesky factor l and the reciprocal diagonal r
      const double d00 = db[0];
      const double2 d1 = *reinterpret_cast<const double2*>(db + 4);
      const double2 d2a = *reinterpret_cast<const double2*>(db + 8);
      const double d22 = db[10];
      const double2 d3a = *reinterpret_cast<const double2*>(db + 12), d3b = *reinterpret_cast<const double2*>(db + 14);
      int bad = 0;
      if (!(d00 > 0.0)) bad = 1;
      const double r0 = rsqrt(d00);
      const double l10 = d1.x * r0, l20 = d2a.x * r0, l30 = d3a.x * r0;
      const double e11 = fma(-l10, l10, d1.y);
      if (!bad && !(e11 > 0.0)) bad = 2;
      const double r1 = rsqrt(e11);
      const double l21 = fma(-l20, l10, d2a.y) * r1, l31 = fma(-l30, l10, d3a.y) * r1;
      const double e22 = fma(-l21, l21, fma(-l20, l20, d22));
      if (!bad && !(e22 > 0.0)) bad = 3;
      const double r2 = rsqrt(e22);
      const double l32 = fma(-l31, l21, fma(-l30, l20, d3b.x)) * r2;
      const double e33 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, d3b.y)));
      if (!bad && !(e33 > 0.0)) bad = 4;
      const double r3 = rsqrt(e33);
      if (bad) {
        if (4 * J + bad <= M) *flag = 4 * J + bad;   // every panel owner writes the same value (padding cannot fail)
      } else {
        double* lg = lb + FFVD_PTILE * ti;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double x0 = a[r][0] * r0;
          const double x1 = fma(-x0, l10, a[r][1]) * r1;
          const double x2 = fma(-x1, l21, fma(-x0, l20, a[r][2])) * r2;
          const double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, a[r][3]))) * r3;
          *reinterpret_cast<double2*>(lg + 4 * r) = make_double2(x0, x1);
          *reinterpret_cast<double2*>(lg + 4 * r + 2) = make_double2(x2, x3);
          const int i = i0 + r;
          if (i < M) {                           // rows of L, final (a diagonal tile carries unused upper entries)
            double* Ar = A + (size_t)i * lda + k0;
            Ar[0] = x0;
            if (k0 + 1 <= i) Ar[1] = x1;
            if (k0 + 2 <= i) Ar[2] = x2;
            if (k0 + 3 <= i) Ar[3] = x3;
          }
        }
      }
    }
    __syncthreads();
    if (*flag) break;                          // CTA uniform
    if (tk > J) {
      const double* li = lb + FFVD_PTILE * ti;
      const double* lk = lb + FFVD_PTILE * tk;
      double Li[4][4], Lk[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double2 p01 = *reinterpret_cast<const double2*>(li + 4 * r), p23 = *reinterpret_cast<const double2*>(li + 4 * r + 2);
        const double2 q01 = *reinterpret_cast<const double2*>(lk + 4 * r), q23 = *reinterpret_cast<const double2*>(lk + 4 * r + 2);
        Li[r][0] = p01.x; Li[r][1] = p01.y; Li[r][2] = p23.x; Li[r][3] = p23.y;
        Lk[r][0] = q01.x; Lk[r][1] = q01.y; Lk[r][2] = q23.x; Lk[r][3] = q23.y;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int m = 0; m < 4; ++m) a[r][c] = fma(-Li[r][m], Lk[c][m], a[r][c]);
      if (tk == J + 1 && ti == tk) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          *reinterpret_cast<double2*>(db + 4 * r) = make_double2(a[r][0], a[r][1]);
          *reinterpret_cast<double2*>(db + 4 * r + 2) = make_double2(a[r][2], a[r][3]);
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  return *flag;
}

// Cholesky of the fast path: the panel form when its panel buffer fits the scratch region (the still unused X staging
// matrix, M * ldx doubles), else the column form.
__device__ __forceinline__ int chol_fast_smem(double* A, int lda, int M, double* cb, int cbs, double* scratch, int* flag) {
#ifndef FFVD_CHOL_COLUMN_FORM
  if (M >= 8) return chol_panel_smem(A, lda, M, scratch, flag);     // 18 (ceil(M/4) + 1) <= M * ldx
#endif
  return chol_owner_smem(A, lda, M, cb, cbs, flag);
}

// X = L^{-1} (lower) by forward substitution, L x_j = e_j; L (M x ldl) and the scratch dinv (M doubles) in SHARED memory.
// Four lanes per column, eight columns per warp (j0 .. j0+7, j0 a multiple of 8).  Rows advance in blocks of four
// ABSOLUTE rows i = 4B .. 4B+3, starting at the block that holds row j0 (entries above the diagonal come out as the
// zeros they are); lane `sub` of a column keeps X[4B' + sub][j] of the finished blocks in registers.  The registers
// are indexed relative to the current block (y[s] = the entry produced s blocks ago, shifted once per block), which makes
// every register index a compile-time constant inside a rolled block loop.  All eight columns of a warp work on the same
// rows and columns of L, so every load of L is a 4-address broadcast.  Per block the four row sums over the finished
// registers are independent of the block's own results and are formed first (two terms per exit test, eight loads in
// flight); then two shuffles per row, and the block's own 4x4 triangular solve is done redundantly by the four lanes so
// that no shuffle sits on the dependent chain.
// The result is staged in the shared matrix Xs (M x ldx) and written to the zero-padded global buffers X and XT = X^T.
__device__ void tri_inverse_smem(const double* __restrict__ L, int ldl, int M, double* __restrict__ Xs, int ldx, double* __restrict__ dinv,
                                 double* __restrict__ X, double* __restrict__ XT, int Mp) {
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nth >> 5;
  for (int i = tid; i < M; i += nth) dinv[i] = 1.0 / L[(size_t)i * ldl + i];
  __syncthreads();
  const int sub = lane & 3, cw = lane >> 2;
  constexpr int NS = (FFVD_FAST_MAXM + 3) / 4;      // register slots per lane
  const int nB = (M + 3) >> 2;
  for (int j0 = 8 * warp; j0 < M; j0 += 8 * nw) {
    const int j = j0 + cw;                       // may be >= M: the lane group solves for a zero right-hand side, stores nothing
    double y[NS + 1];
#pragma unroll
    for (int t = 0; t <= NS; ++t) y[t] = 0.0;
    const int Bs = j0 >> 2;
    for (int B = Bs; B < nB; ++B) {
      FFVD_DBG(64 + B);
      const int b = B - Bs;                      // finished blocks (warp uniform)
      // rows of the block (clamped into the matrix) and their operand pointers at column 4B + sub
      const double* Lr[4];
      int row[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        row[p] = min(4 * B + p, M - 1);
        Lr[p] = L + (size_t)row[p] * ldl + 4 * B + sub;
      }
      double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int s = 1; s < NS; s += 2) {
        if (s > b) break;
        const bool two = s + 1 <= b && s + 1 < NS;
        double l0[4], l1[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          l0[p] = Lr[p][-4 * s];
          l1[p] = two ? Lr[p][-4 * (s + 1)] : 0.0;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) v[p] = fma(l1[p], y[(s + 1 < NS) ? s + 1 : s], fma(l0[p], y[s], v[p]));
      }
      // (kept below the sums on purpose: loaded earlier these values push the sums out of registers and the compiler
      // serialises every load of the loop above behind its multiply-add)
      asm volatile("" ::: "memory");
      // the 4x4 diagonal block of L that couples the block's own four results, and 1/L_ii
      double dv[4], lb[4][3];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        dv[p] = dinv[row[p]];
#pragma unroll
        for (int c = 0; c < 3; ++c) lb[p][c] = (c < p) ? Lr[p][c - sub] : 0.0;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) v[p] += __shfl_xor_sync(0xffffffffu, v[p], 1);
#pragma unroll
      for (int p = 0; p < 4; ++p) v[p] += __shfl_xor_sync(0xffffffffu, v[p], 2);
      // x_p = (delta(4B+p, j) - v_p - sum_{c<p} l_pc x_c) / L_pp
      const int jr = j - 4 * B;                  // row of the block that carries the unit right-hand side (if 0..3)
      const double x0 = ((jr == 0 ? 1.0 : 0.0) - v[0]) * dv[0];
      const double x1 = ((jr == 1 ? 1.0 : 0.0) - fma(lb[1][0], x0, v[1])) * dv[1];
      const double x2 = ((jr == 2 ? 1.0 : 0.0) - fma(lb[2][1], x1, fma(lb[2][0], x0, v[2]))) * dv[2];
      const double x3 = ((jr == 3 ? 1.0 : 0.0) - fma(lb[3][2], x2, fma(lb[3][1], x1, fma(lb[3][0], x0, v[3])))) * dv[3];
      y[0] = (sub == 0) ? x0 : (sub == 1) ? x1 : (sub == 2) ? x2 : x3;
      if (j < M && 4 * B + sub < M) Xs[(size_t)(4 * B + sub) * ldx + j] = y[0];
#pragma unroll
      for (int t = NS; t > 0; --t) y[t] = y[t - 1];
    }
  }
  FFVD_DBG(100);
  __syncthreads();
  FFVD_DBG(101);
  // rows above each column's first block were never staged: the upper triangle is written as zeros from the index test
  for (int r = warp; r < M; r += nw)
    for (int c = lane; c < M; c += 32) {
      X[(size_t)r * Mp + c] = (c <= r) ? Xs[(size_t)r * ldx + c] : 0.0;
      if (XT) XT[(size_t)r * Mp + c] = (c >= r) ? Xs[(size_t)c * ldx + r] : 0.0;
    }
  __syncthreads();
}

// shared-memory doubles of the fast path (0 = does not fit): L (M x (M+1)), X (M x ldx), three Mp vectors
__host__ __device__ inline size_t chol_fast_smem_doubles(int M, int Mp) {
  const int ldx = (M + 1) & ~1;
  return (size_t)M * (M + 1) + (size_t)M * ldx + (size_t)3 * Mp;
}
__host__ inline bool chol_fast_fits(int M, int Mp, size_t max_smem) {
  return M <= FFVD_FAST_MAXM && chol_fast_smem_doubles(M, Mp) * 8 <= max_smem;
}

// ---------------------------------------------------------------------------------------------
// grid (D, nprob); block 512.  Dynamic smem: 2*Mp doubles (+ M*(M+1) when use_smem; chol_fast_smem_doubles when use_smem == 2).
template <int KIND>
__global__ void __launch_bounds__(512) kzz_prep_kernel(const DevProblem* __restrict__ probs, double jitter, int use_smem, int do_ltu) {
  extern __shared__ __align__(16) double sh[];
  __shared__ int flag;
  const DevProblem& P = probs[blockIdx.y];
  const int d = blockIdx.x, M = P.M, Mp = P.Mp, Din = P.Din;   // d indexes kernels (grid.x = D*hs or 1)
  const int tid = threadIdx.x, nth = blockDim.x;
  double* colbuf = sh;
  double* rowbuf = sh + Mp;
  double* Asm = sh + 2 * Mp;
  if (d == 0) {
    // Z~^T, 32 rows: Z^T (zero padded), then a row of ones over the valid inducing points, then zeros
    for (int idx = tid; idx < 32 * Mp; idx += nth) {
      const int jd = idx / Mp, m = idx % Mp;
      double zv = 0.0;
      if (m < M) zv = (jd < Din) ? P.Z[(size_t)m * Din + jd] : (jd == Din ? 1.0 : 0.0);
      P.ZT[idx] = zv;
      P.Zf[((size_t)(m >> 2) * 4 + (jd >> 3)) * 32 + (jd & 7) * 4 + (m & 3)] = zv;
    }
  }
  double* Lt = P.LinvT + (size_t)d * Mp * Mp;   // global scratch for L when it does not fit in smem
  double* A = use_smem ? Asm : Lt;
  const int lda = use_smem ? (M + 1) : Mp;
  const double v = exp(P.logv[d]);
  // K(Z,Z) + jitter I, lower triangle.  kernels_multi_output.py:202-214 / kernels.py:270-276
  const int ldx2 = (M + 1) & ~1;
  if (use_smem == 2 && Din <= ldx2) {
    // fast path: z~ = z / l (SE) or z (linear) staged once in the (still unused) X region, a warp per row of K
    double* zs = Asm + (size_t)M * (M + 1);
    const double* hy = P.hyp + (size_t)d * 72 + 32;          // 1 / l, written by hyper_kernel
    for (int idx = tid; idx < M * Din; idx += nth) zs[idx] = (KIND == 0) ? P.Z[idx] * hy[idx % Din] : P.Z[idx];
    __syncthreads();
    const int lane = tid & 31, nw = nth >> 5;
    for (int m = tid >> 5; m < M; m += nw)
      for (int n = lane; n <= m; n += 32) {
        double s = 0.0;
        for (int jd = 0; jd < Din; ++jd) {
          const double a = zs[m * Din + jd], b = zs[n * Din + jd];
          if (KIND == 0) {
            const double t = a - b;
            s = fma(t, t, s);
          } else {
            s = fma(a, b, s);
          }
        }
        double k = (KIND == 0) ? v * exp(-0.5 * s) : v * s;
        if (m == n) k += jitter;
        Asm[(size_t)m * (M + 1) + n] = k;
      }
  } else
  for (int idx = tid; idx < M * M; idx += nth) {
    const int m = idx / M, n = idx % M;
    if (n > m) continue;
    double s = 0.0;
    for (int jd = 0; jd < Din; ++jd) {
      const double a = P.Z[(size_t)m * Din + jd], b = P.Z[(size_t)n * Din + jd];
      if (KIND == 0) {
        const double il = exp(-P.logl[(size_t)d * Din + jd]);
        const double t = a * il - b * il;
        s = fma(t, t, s);
      } else {
        s = fma(a, b, s);
      }
    }
    double k = (KIND == 0) ? v * exp(-0.5 * s) : v * s;
    if (m == n) k += jitter;
    A[(size_t)m * lda + n] = k;
  }
  __syncthreads();
  const int st = (use_smem == 2) ? chol_fast_smem(Asm, M + 1, M, sh, Mp, Asm + (size_t)M * (M + 1), &flag) : chol_inplace(A, lda, M, colbuf, &flag);
  if (tid == 0) P.status[d] = st;
  if (st != 0) return;
  if (use_smem == 2) {
    double* Xs = Asm + (size_t)M * (M + 1);
    const int ldx = (M + 1) & ~1;
    tri_inverse_smem(Asm, M + 1, M, Xs, ldx, Xs + (size_t)M * ldx, P.Linv + (size_t)d * Mp * Mp, Lt, Mp);
    if (do_ltu && P.U) {
      // the work of ltu_kernel, w = L^{-T} u, from the inverse still in shared memory (one launch and an L2 round trip of
      // L^{-T} less): thread m forms sum_{n >= m} X[n][m] u[n] over consecutive-m (conflict-free) rows of X
      double* us = Xs + (size_t)M * ldx;                     // the inverse's 1/L_ii scratch is free again
      for (int dd = P.hs ? d : 0; dd < (P.hs ? d + 1 : P.D); ++dd) {
        __syncthreads();
        for (int n = tid; n < M; n += nth) us[n] = P.UT[(size_t)dd * Mp + n];
        __syncthreads();
        double* w = P.wvec + (size_t)dd * Mp;
        for (int m = tid; m < Mp; m += nth) {
          double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
          if (m < M) {
            int n = m;
            for (; n + 3 < M; n += 4) {
              t0 = fma(Xs[(size_t)n * ldx + m], us[n], t0);
              t1 = fma(Xs[(size_t)(n + 1) * ldx + m], us[n + 1], t1);
              t2 = fma(Xs[(size_t)(n + 2) * ldx + m], us[n + 2], t2);
              t3 = fma(Xs[(size_t)(n + 3) * ldx + m], us[n + 3], t3);
            }
            for (; n < M; ++n) t0 = fma(Xs[(size_t)n * ldx + m], us[n], t0);
          }
          w[m] = (t0 + t1) + (t2 + t3);
        }
      }
    }
  } else {
    tri_inverse(A, lda, M, P.Linv + (size_t)d * Mp * Mp, Lt, Mp, rowbuf);
  }
}

// Blocked path (M too large for shared memory): fill K(Z,Z) + jitter I (lower triangle, identity on the
// padded diagonal) into Lfac[b], b = problem * nk + d, and write Z~^T.   grid (ceil(Mp*Mp/256), nk, nprob)
template <int KIND>
__global__ void __launch_bounds__(256) kzz_fill_kernel(const DevProblem* __restrict__ probs, double* __restrict__ Lfac, double jitter) {
  const DevProblem& P = probs[blockIdx.z];
  const int d = blockIdx.y, M = P.M, Mp = P.Mp, Din = P.Din, nk = gridDim.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d == 0 && idx < 32LL * Mp) {
    const int jd = (int)(idx / Mp), m = (int)(idx % Mp);
    double zv = 0.0;
    if (m < M) zv = (jd < Din) ? P.Z[(size_t)m * Din + jd] : (jd == Din ? 1.0 : 0.0);
    P.ZT[idx] = zv;
    P.Zf[((size_t)(m >> 2) * 4 + (jd >> 3)) * 32 + (jd & 7) * 4 + (m & 3)] = zv;
  }
  if (idx >= (long long)Mp * Mp) return;
  const int m = (int)(idx / Mp), n = (int)(idx % Mp);
  double k = 0.0;
  if (n <= m) {
    if (m < M) {
      double s = 0.0;
      for (int jd = 0; jd < Din; ++jd) {
        const double a = P.Z[(size_t)m * Din + jd], b = P.Z[(size_t)n * Din + jd];
        if (KIND == 0) {
          const double il = exp(-P.logl[(size_t)d * Din + jd]);
          const double t = a * il - b * il;
          s = fma(t, t, s);
        } else {
          s = fma(a, b, s);
        }
      }
      const double v = exp(P.logv[d]);
      k = (KIND == 0) ? v * exp(-0.5 * s) : v * s;
      if (m == n) k += jitter;
    } else if (m == n) {
      k = 1.0;
    }
  }
  Lfac[((size_t)blockIdx.z * nk + d) * Mp * Mp + idx] = k;
}

// Collapsed bound, blocked path: H = S/Q + I into Wk[b] (identity on the padding).  grid (ceil(Mp*Mp/256), nb, nprob)
__global__ void __launch_bounds__(256) collapsed_fill_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.z];
  const int b = blockIdx.y, d = b % P.D, M = P.M, Mp = P.Mp;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Mp * Mp) return;
  const int m = (int)(idx / Mp), n = (int)(idx % Mp);
  double h = 0.0;
  if (m < M && n < M) h = P.Sacc[(size_t)b * Mp * Mp + idx] * exp(-P.logQ[d]) + (m == n ? 1.0 : 0.0);
  else if (m == n) h = 1.0;
  P.Wk[(size_t)b * Mp * Mp + idx] = h;
}

// -1/2 logdet H = -sum_i log L_ii from the blocked factor in Wk[b].  grid (nb, nprob); block 256
__global__ void __launch_bounds__(256) collapsed_logdet_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.y];
  const int b = blockIdx.x, s = b / P.D, M = P.M, Mp = P.Mp;
  const double* L = P.Wk + (size_t)b * Mp * Mp;
  double t = 0.0;
  for (int i = threadIdx.x; i < M; i += 256) t += log(L[(size_t)i * Mp + i]);
  t = warp_sum(t);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    P.collb[(size_t)b * 4 + 0] = -tot;          // per (s,d), summed over d in a fixed order by finalize_kernel (no RED: repeatable)
    (void)s;
  }
}

// ---------------------------------------------------------------------------------------------
// C[b] = alpha * A[b] * B[b]  (all n x n row-major, ld = n, n multiple of 64), FP64 DMMA.
// grid (n/64, n/64, batch); block 256.  Batch strides in elements (0 = shared operand).
struct BatchMap { int div, mul, mod; };   // matrix index = (z / div) * mul + (z % mod)
__device__ __forceinline__ size_t bmap(const BatchMap& m, int z) { return (size_t)(z / m.div) * m.mul + (z % m.mod); }

// TM = 64 or 32 rows of C per CTA (32: twice the CTAs, for launches that would leave most SMs idle -- a 128^3 product
// per matrix of a 4-matrix batch is 16 CTAs of 64 x 64, each bound by its own DMMA throughput).
template <int TM>
__global__ void __launch_bounds__(256) bgemm_nn_kernel(double* __restrict__ C, const double* __restrict__ A,
                                                        const double* __restrict__ B, int n, double alpha,
                                                        BatchMap mC, BatchMap mA, BatchMap mB, int tri) {
  // tri: bit 0 = A is upper triangular (A[m][k] = 0 for k < m: L^{-T}), bit 1 = B is lower triangular (B[k][n] = 0 for
  // k < n: L^{-1}).  The k loop then starts at the first slab that can hold a non-zero product (half the work per
  // triangular operand: at M = 2048 the two products of the Cholesky backward were 9 of 100 ms per evaluation).
  // TM x 64 tile of C per CTA, k in steps of 32.  The next step's A / B slabs are fetched into registers (16-byte
  // loads, all in flight together) while the current one is multiplied out of shared memory: at n = 128 the kernel is
  // four L2 round trips long instead of the 64 serialised ones of a load -> store loop.
  constexpr int BK = 32, NI = TM / 16, NA = TM / 16;   // NI: 8-row blocks per warp, NA: A loads per thread
  __shared__ __align__(16) double As[TM][BK + 4];     // stride 36 = 4 mod 16: fragment reads conflict free per half-warp
  __shared__ __align__(16) double Bs[BK][68];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps: each (TM/2) x 16
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * 64;
  A += bmap(mA, blockIdx.z) * n * n; B += bmap(mB, blockIdx.z) * n * n; C += bmap(mC, blockIdx.z) * n * n;
  double c[NI][2][2];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  double2 ra[NA], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + 256 * i;
      ra[i] = *reinterpret_cast<const double2*>(A + (size_t)(m0 + (idx >> 4)) * n + k0 + 2 * (idx & 15));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      rb[i] = *reinterpret_cast<const double2*>(B + (size_t)(k0 + (idx >> 5)) * n + n0 + 2 * (idx & 31));
    }
  };
  const int kbeg = max((tri & 1) ? m0 : 0, (tri & 2) ? n0 : 0) & ~(BK - 1);
  fetch(kbeg);
  for (int k0 = kbeg; k0 < n; k0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + 256 * i;
      *reinterpret_cast<double2*>(&As[idx >> 4][2 * (idx & 15)]) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      *reinterpret_cast<double2*>(&Bs[idx >> 5][2 * (idx & 31)]) = rb[i];
    }
    __syncthreads();
    if (k0 + BK < n) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[NI], b[2];
#pragma unroll
      for (int i = 0; i < NI; ++i) a[i] = As[wm * (TM / 2) + 8 * i + g][kk + q];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bs[kk + q][wn * 16 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double* p = C + (size_t)(m0 + wm * (TM / 2) + 8 * i + g) * n + n0 + wn * 16 + 8 * j + 2 * q;
      *reinterpret_cast<double2*>(p) = make_double2(alpha * c[i][j][0], alpha * c[i][j][1]);
    }
}

// Sacc[b] <- symmetric matrix built from the lower triangle of
//   mode 0 (uncollapsed G):   Sacc[b]/Q_d + U[:,d] (x) ubar[d]
//   mode 1 (collapsed):       Sacc[b]                         (S itself)
//   mode 2 (collapsed G):     HxT[b] + c[b] (x) b[b]/Q_d       (HxT holds Mat' S)
// grid (ceil(M*M/256), nb, nprob).
__global__ void symmetrize_lower_kernel(const DevProblem* __restrict__ probs, int mode) {
  const DevProblem& P = probs[blockIdx.z];
  const int b = blockIdx.y, d = b % P.D, M = P.M, Mp = P.Mp;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * M) return;
  const int m = (int)(idx / M), n = (int)(idx % M);
  if (n > m) return;
  const size_t off = (size_t)b * Mp * Mp;
  double val;
  if (mode == 0) {
    val = P.Sacc[off + (size_t)m * Mp + n] * exp(-P.logQ[d]) + P.U[(size_t)m * P.D + d] * P.ubar[(size_t)b * Mp + n];
  } else if (mode == 1) {
    val = P.Sacc[off + (size_t)m * Mp + n];
  } else {
    val = P.HxT[off + (size_t)m * Mp + n] + P.cvec[(size_t)b * Mp + m] * P.ubar[(size_t)b * Mp + n] * exp(-P.logQ[d]);
  }
  P.Sacc[off + (size_t)m * Mp + n] = val;
  P.Sacc[off + (size_t)n * Mp + m] = val;
}

// Wz = Kbar_zz o Kzz (SE, in place) and row sums rs[b][m]; Linear: rs unused, Wz = Kbar_zz.
// grid (ceil(M/8), batch); block 256 (warp per row).
template <int KIND>
__global__ void __launch_bounds__(256) wz_kernel(const DevProblem* __restrict__ probs) {
  // blockIdx.z = problem ; blockIdx.y = local batch (d or s*D+d)
  const DevProblem& P = probs[blockIdx.z];
  const int b = blockIdx.y, d = b % P.D;
  const int M = P.M, Mp = P.Mp, Din = P.Din;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  double* Kb = P.Sacc + (size_t)b * Mp * Mp + (size_t)m * Mp;
  double rs = 0.0;
  if (KIND == 0) {
    const double v = exp(P.logv[d]);
    for (int n = lane; n < M; n += 32) {
      double s = 0.0;
      for (int jd = 0; jd < Din; ++jd) {
        const double il = exp(-P.logl[(size_t)d * Din + jd]);
        const double t = P.Z[(size_t)m * Din + jd] * il - P.Z[(size_t)n * Din + jd] * il;
        s = fma(t, t, s);
      }
      const double wz = Kb[n] * v * exp(-0.5 * s);
      Kb[n] = wz;
      rs += wz;
    }
    rs = warp_sum(rs);
  }
  if (lane == 0) P.rs[(size_t)b * Mp + m] = rs;
}

// wz_kernel + kzz_bwd_kernel in one pass (the small-M path: two launches and an L2 round trip of Wz less, and the loop
// over n is spread over the lanes so that a row costs ~3 L2 round trips instead of ~M/2 partially overlapped ones).
// A warp per row m of Kbar_zz (in Sacc, left untouched): lanes take n = lane, lane + 32, ...; per n the row Z[n][:] is
// loaded once and used both for the kernel value and for the weighted sums  wzz[jd] = sum_n Wz[m][n] Z[n][jd].
// grid (ceil(M/8), batch, nprob); block 256.  hyp (1/l, 1/l^2, v) comes from hyper_kernel.
// Deterministic mode (P.kzzpart != null): the three kinds of contributions are STORED per batch entry -- Z-bar rows [M][Din], the
// per-block dJ/dlogl [blocks][Din] and dJ/dlogv [blocks] partials -- and kzz_bwd_reduce_kernel adds them in a fixed order.
template <int KIND>
__global__ void __launch_bounds__(256) kzz_bwd_fused_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.z];
  const int b = blockIdx.y, d = b % P.D;
  const int nblk_ = gridDim.x;
  double* part = P.kzzpart ? P.kzzpart + (size_t)b * ((size_t)P.M * P.Din + (size_t)nblk_ * (P.Din + 1)) : nullptr;
  const int M = P.M, Mp = P.Mp, Din = P.Din;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 8 + warp;
  __shared__ double zm[8][32], lsm[8][32];
  const double* hy = P.hyp + (size_t)d * P.hs * 72;
  const double il = (KIND == 0 && lane < Din) ? hy[32 + lane] : 0.0;      // 1 / l_jd in lane jd
  const bool rowok = m < M;
  const int mc = rowok ? m : M - 1;
  const double zmine = lane < Din ? P.Z[(size_t)mc * Din + lane] : 0.0;
  zm[warp][lane] = (KIND == 0) ? zmine * il : zmine;                        // z~_m (SE) for the distance
  lsm[warp][lane] = il;
  __syncwarp();
  const double* Kb = P.Sacc + (size_t)b * Mp * Mp + (size_t)mc * Mp;
  const double v = hy[64];
  double rs = 0.0, lpart = 0.0, vpart = 0.0, zb = 0.0;
  // chunks of eight input dimensions (registers); the kernel values are recomputed per chunk when Din > 8
  for (int j0 = 0; j0 < Din; j0 += 8) {
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double rsum = 0.0;
    for (int n = lane; n < M; n += 32) {
      const double* zn = P.Z + (size_t)n * Din;
      double wz = Kb[n];
      if (KIND == 0) {
        double s = 0.0;
        for (int jd = 0; jd < Din; ++jd) {
          const double t = zm[warp][jd] - zn[jd] * lsm[warp][jd];
          s = fma(t, t, s);
        }
        wz *= v * exp(-0.5 * s);
      }
      rsum += wz;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (j0 + q < Din) acc[q] = fma(wz, zn[j0 + q], acc[q]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
    rs = rsum;
    double wzz = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) wzz = (lane == j0 + q) ? acc[q] : wzz;
    if (rowok && lane >= j0 && lane < j0 + 8 && lane < Din) {
      if (KIND == 0) {
        zb = -2.0 * hy[lane] * (zmine * rs - wzz);
        lpart = -zmine * zb;
      } else {
        zb = 2.0 * v * wzz;
        vpart = 0.5 * zmine * zb;
      }
      if (part) part[(size_t)m * Din + lane] = zb;
      else red_add(P.gZ + (size_t)m * Din + lane, zb);
    }
  }
  if (KIND == 0) {
    if (rowok && lane == 0) vpart = rs;
    // reduce lpart over the 8 rows of the block through shared memory, one atomic per jd
    __syncthreads();
    lsm[warp][lane] = lpart;
    __syncthreads();
    if (warp == 0 && lane < Din) {
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += lsm[r][lane];
      if (part) part[(size_t)M * Din + (size_t)blockIdx.x * Din + lane] = t;
      else red_add(P.gl + (size_t)d * Din + lane, t);
    }
  }
  vpart = warp_sum(vpart);
  if (part) {
    // the eight row values of the block, summed in warp order
    __shared__ double vsm[8];
    if (lane == 0) vsm[warp] = vpart;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int r = 0; r < 8; ++r) t += vsm[r];
      part[(size_t)M * Din + (size_t)nblk_ * Din + blockIdx.x] = t;
    }
  } else if (lane == 0 && vpart != 0.0) {
    red_add(P.gv + d, vpart);
  }
}

// Deterministic mode: gZ / gl / gv += the stored contributions of kzz_bwd_fused_kernel, batch entries and blocks in index order.
// grid (nprob); block 256.
__global__ void __launch_bounds__(256) kzz_bwd_reduce_kernel(const DevProblem* __restrict__ probs, int nb, int nblk) {
  const DevProblem& P = probs[blockIdx.x];
  const int M = P.M, Din = P.Din, D = P.D;
  const size_t per = (size_t)M * Din + (size_t)nblk * (Din + 1);
  for (int i = threadIdx.x; i < M * Din; i += blockDim.x) {
    double t = P.gZ[i];
    for (int b = 0; b < nb; ++b) t += P.kzzpart[(size_t)b * per + i];
    P.gZ[i] = t;
  }
  for (int i = threadIdx.x; i < D * Din; i += blockDim.x) {
    const int d = i / Din, jd = i % Din;
    double t = P.gl[i];
    for (int b = d; b < nb; b += D)
      for (int k = 0; k < nblk; ++k) t += P.kzzpart[(size_t)b * per + (size_t)M * Din + (size_t)k * Din + jd];
    P.gl[i] = t;
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double t = P.gv[d];
    for (int b = d; b < nb; b += D)
      for (int k = 0; k < nblk; ++k) t += P.kzzpart[(size_t)b * per + (size_t)M * Din + (size_t)nblk * Din + k];
    P.gv[d] = t;
  }
}

// Deterministic mode: dst[i] += sum over the private copies c = 0 .. ncopies-1 of priv[c][i], in index order.
__global__ void det_reduce_kernel(double* __restrict__ dst, const char* __restrict__ priv, size_t stride_bytes, size_t n, int ncopies) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double t = dst[i];
    for (int c = 0; c < ncopies; ++c) t += reinterpret_cast<const double*>(priv + (size_t)c * stride_bytes)[i];
    dst[i] = t;
  }
}

// Fold the raw W^T [Xc,1] sums of the fused kernel (P.gZd: per d Mp x PW values in the zbar_index layout, column jd < Din = sum_t W[t][m] x[t][jd], column Din = the
// column sum of W) into dJ/dZ and the Z part of dJ/dlogl:
//   SE:      zb = (c - colsum * z[m][jd]) / l_{d,jd}^2 ;  gZ[m][jd] += sum_d zb ;  gl[d][jd] -= sum_m z[m][jd] zb
//   Linear:  zb = v_d c                               ;  gZ[m][jd] += sum_d zb
// grid (D + 1, nprob), block 256 (lane = input column, a warp per row m).  Blocks 0..D-1 form gl[d] (their own row of gl: no
// other writer), block D forms gZ (the only writer of gZ at this point of the stream, d in index order): plain stores in a
// fixed order, so the step is repeatable as it stands (FFVD_FLAG_DETERMINISTIC needs nothing extra here).
template <int KIND>
__global__ void __launch_bounds__(256) zbar_post_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.y];
  const int D = P.D, Din = P.Din, M = P.M, Mp = P.Mp;
  const int nbm = (Din + 1 + 7) >> 3, PW = 8 * (nbm < 4 ? nbm : 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool in = lane < Din;
  if ((int)blockIdx.x < D) {
    if (KIND != 0) return;
    const int d = blockIdx.x;
    const double* cz = P.gZd + (size_t)d * Mp * PW;
    const double il2 = in ? P.hyp[(size_t)d * P.hs * 72 + lane] : 0.0;
    double ls = 0.0;
    for (int m = warp; m < M; m += 8) {
      const double c = in ? cz[zbar_index(Mp, m, lane)] : 0.0, cs = cz[zbar_index(Mp, m, Din)];
      const double z = in ? P.Z[(size_t)m * Din + lane] : 0.0;
      const double zb = il2 * (c - cs * z);
      ls = fma(-z, zb, ls);
    }
    __shared__ double part[8][32];
    part[warp][lane] = ls;
    __syncthreads();
    if (warp == 0 && in) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane];
      P.gl[(size_t)d * Din + lane] += t;
    }
  } else {
    for (int m = warp; m < M; m += 8) {
      if (!in) continue;
      const double z = P.Z[(size_t)m * Din + lane];
      double t = 0.0;
      for (int d = 0; d < D; ++d) {
        const double* cz = P.gZd + (size_t)d * Mp * PW;
        if (KIND == 0) t += P.hyp[(size_t)d * P.hs * 72 + lane] * (cz[zbar_index(Mp, m, lane)] - cz[zbar_index(Mp, m, Din)] * z);
        else t += P.hyp[(size_t)d * P.hs * 72 + 64] * cz[zbar_index(Mp, m, lane)];
      }
      P.gZ[(size_t)m * Din + lane] += t;
    }
  }
}

// dJ/dZ, dJ/dlogl, dJ/dlogv contributions of Kbar_zz.  grid (ceil(M/8), batch, nprob); block (32, 8).
template <int KIND>
__global__ void __launch_bounds__(256) kzz_bwd_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.z];
  const int b = blockIdx.y, d = b % P.D;
  const int M = P.M, Mp = P.Mp, Din = P.Din;
  const int jd = threadIdx.x & 31, mi = threadIdx.x >> 5;
  const int m = blockIdx.x * 8 + mi;
  double lpart = 0.0, vpart = 0.0;
  if (m < M && jd < Din) {
    const double* Wz = P.Sacc + (size_t)b * Mp * Mp + (size_t)m * Mp;
    double s0 = 0.0, s1 = 0.0;
    int n = 0;
    for (; n + 1 < M; n += 2) {
      s0 = fma(Wz[n], P.Z[(size_t)n * Din + jd], s0);
      s1 = fma(Wz[n + 1], P.Z[(size_t)(n + 1) * Din + jd], s1);
    }
    if (n < M) s0 = fma(Wz[n], P.Z[(size_t)n * Din + jd], s0);
    const double wzz = s0 + s1;
    const double z = P.Z[(size_t)m * Din + jd];
    double zb;
    if (KIND == 0) {
      const double il2 = exp(-2.0 * P.logl[(size_t)d * Din + jd]);
      const double rs = P.rs[(size_t)b * Mp + m];
      zb = -2.0 * il2 * (z * rs - wzz);
      lpart = -z * zb;
      if (jd == 0) vpart = rs;
    } else {
      zb = 2.0 * exp(P.logv[d]) * wzz;
      vpart = 0.5 * z * zb;
    }
    red_add(P.gZ + (size_t)m * Din + jd, zb);
  }
  if (KIND == 0) {
    // reduce lpart over the 8 rows of the block through shared memory, one atomic per jd
    __shared__ double lsm[8][32];
    lsm[mi][jd] = lpart;
    __syncthreads();
    if (mi == 0 && jd < Din) {
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += lsm[r][jd];
      red_add(P.gl + (size_t)d * Din + jd, t);
    }
  }
  vpart = warp_sum(vpart);
  if (jd == 0 && vpart != 0.0) red_add(P.gv + d, vpart);
}

// ---------------------------------------------------------------------------------------------
// Collapsed bound, per (problem, s, d): H = S/Q + I  -> chol -> logdet, H^{-1}, c = H^{-1} b/Q,
// quad = 1/2 b^T c /Q..., w' = L^{-T} c / Q, implicit dJ/dlogQ.   conditionals_multi_output.py:246-254
// grid (S*D, nprob); block 512; dynamic smem 2*Mp doubles.
// Buffers: Sacc[b] holds the full symmetric S (already symmetrized); Wk[b] <- H then L_H;
//          Hx[b] <- L_H^{-1}; HxT[b] <- L_H^{-T}.
__global__ void __launch_bounds__(512) collapsed_chol_kernel(const DevProblem* __restrict__ probs, int fast) {
  extern __shared__ __align__(16) double sh[];
  __shared__ int flag;
  const DevProblem& P = probs[blockIdx.y];
  const int b = blockIdx.x, d = b % P.D, s = b / P.D;
  const int M = P.M, Mp = P.Mp, tid = threadIdx.x, nth = blockDim.x;
  const double iq = exp(-P.logQ[d]);
  const double* S = P.Sacc + (size_t)b * Mp * Mp;
  // fast: H, its factor and the inverse live in shared memory (sh: 2 Mp vectors, then H, then X, then dinv)
  double* H = fast ? sh + 2 * Mp : P.Wk + (size_t)b * Mp * Mp;
  const int ldh = fast ? M + 1 : Mp;
  {
    // a warp per row, four loads in flight per lane (a flat load -> store loop pays one L2 round trip per element)
    const int lane = tid & 31, nw = nth >> 5;
    for (int m = tid >> 5; m < M; m += nw) {
      const double* __restrict__ Srow = S + (size_t)m * Mp;
      double* __restrict__ Hrow = H + (size_t)m * ldh;
#pragma unroll 4
      for (int n = lane; n < M; n += 32) Hrow[n] = Srow[n] * iq + (m == n ? 1.0 : 0.0);
    }
  }
  __syncthreads();
  const int st = fast ? chol_fast_smem(sh + 2 * Mp, M + 1, M, sh, Mp, sh + 2 * Mp + (size_t)M * (M + 1), &flag) : chol_inplace(H, ldh, M, sh, &flag);
  if (st != 0) {
    if (tid == 0) P.status[d] = st;
    return;
  }
  // -1/2 logdet H = -sum log L_ii
  if (tid < 32) {
    double t = 0.0;
    for (int i = tid; i < M; i += 32) t += log(H[(size_t)i * ldh + i]);
    t = warp_sum(t);
    if (tid == 0) P.collb[(size_t)b * 4 + 0] = -t;      // per (s,d); finalize_kernel sums over d in a fixed order
    (void)s;
  }
  double* X = P.Hx + (size_t)b * Mp * Mp;
  double* XT = P.HxT + (size_t)b * Mp * Mp;
  if (fast) {
    double* Hs = sh + 2 * Mp;
    double* Xs = Hs + (size_t)M * (M + 1);
    tri_inverse_smem(Hs, M + 1, M, Xs, (M + 1) & ~1, Xs + (size_t)M * ((M + 1) & ~1), X, XT, Mp);
  } else {
    tri_inverse(H, Mp, M, X, XT, Mp, sh + Mp);
  }
}

// Per-evaluation scalars the tile kernel would otherwise re-derive (with FP64 exp calls) in every work item:
// hyp[k] = {1/l_j^2, 1/l_j, v} per kernel, hq[d] = {Q, 1/Q, log Q} per output dim, UT = U^T zero padded.
// grid (max(nk, D), nprob); block 128.
// zs != 0 (SE only): also the per-kernel scaled inducing inputs z~ = z / l_d (transposed, zero padded), a row of ones and, in
// row Din+1, -1/2 |z~_m|^2 -- the column half of the reference's expansion of the scaled squared distance
// (kernels_multi_output.py:163-182).  Skipped when the factors of the previous call are reused (Z, l unchanged).
// guard != 0: content guard of FFVD_FLAG_REUSE_KZZ.  Every block adds the (order-independent) hash of its slice of Z, logv,
// logl to P.guard[0]; the last block to finish either records it as the hash the factors were built from (guard == 1) or
// compares it with that record (guard == 2, the factors are being reused) and raises the sticky stale flag P.guard[3] on a
// mismatch -- the evaluation then returns NaN (finalize_kernel / the COND statistics) and the next synchronising call
// FFVD_E_STALE, instead of silently using the factors of another Z.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
__global__ void hyper_kernel(const DevProblem* __restrict__ probs, int kind, int nk, int zs, int guard) {
  const DevProblem& P = probs[blockIdx.y];
  const int k = blockIdx.x, t = threadIdx.x, Din = P.Din, D = P.D, M = P.M, Mp = P.Mp;
  __shared__ double sils[32];
  if (k < nk) {
    double* h = P.hyp + (size_t)k * 72;
    if (t < 32) {
      double il2 = 0.0, sil = 0.0;
      if (kind == 0 && t < Din) {
        const double ll = P.logl[(size_t)k * Din + t];
        il2 = exp(-2.0 * ll);
        sil = exp(-ll);
      }
      h[t] = il2; h[32 + t] = sil;
      sils[t] = sil;
    }
    if (t == 32) h[64] = exp(P.logv[k]);
  }
  __syncthreads();
  if (zs && k < nk) {
    double* out = P.ZTs + (size_t)k * FFVD_ZTS_ROWS * Mp;
    // rows 0..Din-1: z~; row Din: ones; row Din+1: -1/2 |z~|^2; zero rows up to the next multiple of four and beyond.  The
    // tile kernel forms -r^2/2 = [x~, -1/2|x~|^2, 1] . [z~; 1; -1/2|z~|^2] as ONE tensor-pipe product over these rows.
    for (int m = t; m < Mp; m += blockDim.x) {
      double a = 0.0;
      for (int jd = 0; jd < Din; ++jd) {
        const double z = (m < M) ? P.Z[(size_t)m * Din + jd] * sils[jd] : 0.0;
        out[(size_t)jd * Mp + m] = z;
        a = fma(z, z, a);
      }
      out[(size_t)Din * Mp + m] = 1.0;
      out[(size_t)(Din + 1) * Mp + m] = (m < M) ? -0.5 * a : -1.0e4;     // padded columns: exp saturates at ~1e-308 (see compute_k_tile)
      for (int jd = Din + 2; jd < FFVD_ZTS_ROWS; ++jd) out[(size_t)jd * Mp + m] = 0.0;
    }
  }
  if (k < D) {
    if (t == 33) {
      const double lq = P.logQ ? P.logQ[k] : 0.0;
      const double Qv = exp(lq);
      double* q = P.hq + (size_t)k * 4;
      q[0] = Qv; q[1] = 1.0 / Qv; q[2] = lq; q[3] = 0.0;
    }
    if (P.U)
      for (int j = t; j < Mp; j += blockDim.x) P.UT[(size_t)k * Mp + j] = (j < M) ? P.U[(size_t)j * D + k] : 0.0;
  }
  if (guard && P.guard) {
    const int nz = M * Din, nl = (kind == 0) ? nk * Din : 0, ntot = nz + nk + nl;
    unsigned long long h = 0;
    for (int i = k * blockDim.x + t; i < ntot; i += gridDim.x * blockDim.x) {
      const double v = (i < nz) ? P.Z[i] : (i < nz + nk ? P.logv[i - nz] : P.logl[i - nz - nk]);
      h += mix64((unsigned long long)__double_as_longlong(v) ^ mix64((unsigned long long)i + 1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((t & 31) == 0 && h) atomicAdd(P.guard + 0, h);
    __syncthreads();
    if (t == 0) {
      __threadfence();
      const unsigned long long done = atomicAdd(P.guard + 2, 1ull);
      if (done == (unsigned long long)gridDim.x - 1) {
        __threadfence();
        const unsigned long long cur = atomicAdd(P.guard + 0, 0ull) | 1ull;      // never 0: 0 marks "no valid factors"
        if (guard == 1) { P.guard[1] = cur; P.guard[3] = 0; }
        else if (P.guard[1] != cur) P.guard[3] = 1;
        P.guard[0] = 0; P.guard[2] = 0;
      }
    }
  }
}

// x-bar <- 0 for every problem of a batch.  grid (blocks, nprob); block 256.
__global__ void zero_gx_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.y];
  if (!P.gX) return;
  const size_t n = (size_t)P.S * (P.T + 1) * P.D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) P.gX[i] = 0.0;
}

// Four rows of a matrix-vector product per warp: t[r] = sum_{n < M} A[(m0 + r) * ld + n] * x[n] on every lane.  The
// four loads of a trip are independent and two trips are unrolled, so eight L2 round trips are in flight per lane (a
// row-at-a-time loop pays them one after the other: ~1 us per row).  Rows >= M repeat row M-1 (discard the result).
__device__ __forceinline__ void warp_rows4_dot(const double* __restrict__ A, int ld, const double* __restrict__ x, int m0, int M, int lane,
                                               double (&t)[4]) {
  const double* a[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    t[r] = 0.0;
    a[r] = A + (size_t)min(m0 + r, M - 1) * ld;
  }
#pragma unroll 2
  for (int n = lane; n < M; n += 32) {
    const double xv = x[n];
#pragma unroll
    for (int r = 0; r < 4; ++r) t[r] = fma(a[r][n], xv, t[r]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < 4; ++r) t[r] += __shfl_xor_sync(0xffffffffu, t[r], o);
}

// Uncollapsed: w_d = L_d^{-T} u_d, so that the fused kernel can form Kbar = (A L^{-1})/Q + e w^T without touching the
// A operand of its second contraction.  grid (D, nprob, row chunks); block 256 (four rows of L^{-T} per warp; the zero lower part
// of the rows is multiplied through).  u_d is read from the transposed copy written by hyper_kernel.
__global__ void __launch_bounds__(256) ltu_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.y];
  const int d = blockIdx.x, M = P.M, Mp = P.Mp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* LT = P.LinvT + (size_t)d * P.hs * Mp * Mp;
  const double* u = P.UT + (size_t)d * Mp;
  double* w = P.wvec + (size_t)d * Mp;
  // rows dealt over gridDim.z CTAs in chunks of 32 (M = 2048: one CTA per d read its 32 MB of L^{-T} alone -- 3 ms)
  for (int m0 = 4 * warp + 32 * blockIdx.z; m0 < M; m0 += 32 * gridDim.z) {
    double t[4];
    warp_rows4_dot(LT, Mp, u, m0, M, lane, t);
    if (lane < 4 && m0 + lane < M) w[m0 + lane] = (lane == 0) ? t[0] : (lane == 1) ? t[1] : (lane == 2) ? t[2] : t[3];
  }
  if (blockIdx.z == 0)
    for (int m = M + threadIdx.x; m < Mp; m += 256) w[m] = 0.0;
}

// After Hinv = L_H^{-T} L_H^{-1} is in Wk[b]:  c = Hinv b/Q ; w' = L^{-T} c / Q ; quad; dJ/dlogQ;
// then Wk[b] <- Mat' = (I - Hinv - c c^T)/Q.     grid (S*D, nprob); block 256.
__global__ void __launch_bounds__(1024) collapsed_vec_kernel(const DevProblem* __restrict__ probs) {
  const DevProblem& P = probs[blockIdx.y];
  const int b = blockIdx.x, d = b % P.D, s = b / P.D;
  const int M = P.M, Mp = P.Mp, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nth = blockDim.x, nw = nth >> 5;
  const double iq = exp(-P.logQ[d]);
  double* Hinv = P.Wk + (size_t)b * Mp * Mp;
  const double* bv = P.ubar + (size_t)b * Mp;      // unscaled F^T delta
  double* c = P.cvec + (size_t)b * Mp;
  double* w = P.wvec + (size_t)b * Mp;
  const double* S = P.Sacc + (size_t)b * Mp * Mp;
  const double* LT = P.LinvT + (size_t)d * Mp * Mp;
  __shared__ double red[4];
  __shared__ double bs[2048], cs[2048];            // b/Q and c (Mp <= 2048)
  if (tid < 4) red[tid] = 0.0;
  for (int n = tid; n < Mp; n += nth) {
    bs[n] = (n < M) ? bv[n] * iq : 0.0;
    cs[n] = 0.0;
  }
  __syncthreads();
  // c = Hinv (b/Q)
  for (int m0 = 4 * warp; m0 < M; m0 += 4 * nw) {
    double t[4];
    warp_rows4_dot(Hinv, Mp, bs, m0, M, lane, t);
    if (lane < 4 && m0 + lane < M) cs[m0 + lane] = (lane == 0) ? t[0] : (lane == 1) ? t[1] : (lane == 2) ? t[2] : t[3];
  }
  __syncthreads();
  for (int n = tid; n < Mp; n += nth) c[n] = cs[n];
  // quad = 1/2 b_s^T c ; c^T b_s ; trace Hinv ; c^T (H - I) c = c^T S c / Q
  double qd = 0.0, tr = 0.0, csc = 0.0;
  for (int m = tid; m < M; m += nth) {
    qd = fma(bs[m], cs[m], qd);
    tr += Hinv[(size_t)m * Mp + m];
  }
  for (int m0 = 4 * warp; m0 < M; m0 += 4 * nw) {
    double t[4], u[4];
    warp_rows4_dot(S, Mp, cs, m0, M, lane, t);
    // w' = L^{-T} c / Q   (LinvT is upper; its zero lower part is multiplied through)
    warp_rows4_dot(LT, Mp, cs, m0, M, lane, u);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (m0 + r < M) csc = fma(cs[m0 + r], t[r], csc);
    }
    if (lane < 4 && m0 + lane < M) w[m0 + lane] = ((lane == 0) ? u[0] : (lane == 1) ? u[1] : (lane == 2) ? u[2] : u[3]) * iq;
  }
  for (int m = M + tid; m < Mp; m += nth) w[m] = 0.0;
  qd = warp_sum(qd); tr = warp_sum(tr); csc = warp_sum(csc);
  __shared__ double wred[32][3];
  if (lane == 0) { wred[warp][0] = qd; wred[warp][1] = tr; wred[warp][2] = csc; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 0; w < nw; ++w) { red[0] += wred[w][0]; red[1] += wred[w][1]; red[2] += wred[w][2]; }     // fixed order
    // per (s,d) values, summed over d (terms) / over s (dJ/dlogQ) in a fixed order by finalize_kernel
    P.collb[(size_t)b * 4 + 1] = 0.5 * red[0];
    P.collb[(size_t)b * 4 + 2] = 0.5 * ((double)M - red[1]) - red[0] + 0.5 * red[2] * iq;
  }
  (void)s;
  // Mat' in place: four rows per warp in flight (loads of all four first, then the stores)
  for (int m0 = 4 * warp; m0 < M; m0 += 4 * nw)
    for (int n = lane; n < M; n += 32) {
      double h[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) h[r] = Hinv[(size_t)min(m0 + r, M - 1) * Mp + n];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (m0 + r < M) Hinv[(size_t)(m0 + r) * Mp + n] = ((m0 + r == n ? 1.0 : 0.0) - h[r] - cs[m0 + r] * cs[n]) * iq;
    }
}

// ---------------------------------------------------------------------------------------------
struct OutPtrs {
  double *nll, *terms, *g_Z, *g_U, *g_logv, *g_logl, *g_logQ, *g_C, *g_d, *g_logR;
};

// priors (dgp_model.py:105-143,252,326-334), -1/T scaling, term assembly (dgp_model.py:286-297).
// grid (1 + gx_blocks, nprob); block 256: block 0 of a problem assembles the terms and the small gradients, the others
// scale x-bar (gx_blocks = 0: no x-bar, e.g. FFVD_FLAG_NO_GRADS).
template <int KIND>
__global__ void __launch_bounds__(256) finalize_kernel(const DevProblem* __restrict__ probs, const OutPtrs* __restrict__ outs,
                                                       int collapsed, int flags, int gx_blocks, int nstatus,
                                                       const int* __restrict__ status2, int nstatus2) {
  const DevProblem& P = probs[blockIdx.y];
  if (blockIdx.x > 0) {
    // blocks 1 .. gx_blocks: the work of scale_gx_kernel (one launch less), g_X = -(raw - [t==0] X_0)/T in place
    const size_t per = (size_t)(P.T + 1) * P.D, n = per * P.S;
    const double scx = -1.0 / (double)P.T;
    for (size_t i = (size_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x; i < n; i += (size_t)gx_blocks * blockDim.x) {
      const size_t r = i % per;
      double v = P.gX[i];
      if (P.gXp) v = ((v + P.gXp[i]) + P.gXp[P.gXp_stride + i]) + P.gXp[2 * P.gXp_stride + i];      // deterministic mode: the x-bar planes
      if (r < (size_t)P.D && !(flags & 32)) v -= P.X[i];
      P.gX[i] = scx * v;
    }
    return;
  }
  const OutPtrs& O = outs[blockIdx.y];
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nth >> 5;
  const int M = P.M, Mp = P.Mp, D = P.D, Din = P.Din, Dy = P.Dy, S = P.S, T = P.T;
  const double sc = -1.0 / (double)T;
  const bool shared_priors = (flags & 16) == 0, x0_prior = (flags & 32) == 0;
  const bool replicated = (flags & 512) == 0;      // FFVD_FLAG_NO_REPLICATED: the H-dependent terms are counted by another rank
  const double npri = !shared_priors ? 0.0 : ((flags & 2) ? 1.0 : (double)S);
  const bool zprior = (flags & 1) != 0;
  const double log005 = log(0.05);
  // The eight prior sums of squares in ONE pass: the loads of a trip are independent (one L2 round trip per trip instead
  // of one block reduction, three barriers and a round trip per array), then a single block reduction of the 8-vector.
  //   0 Z   1 logv - log 0.05   2 logl   3 U   4 logQ   5 C   6 d   7 logR
  const double* arr[8] = {zprior ? P.Z : nullptr, P.logv, KIND == 0 ? P.logl : nullptr, collapsed ? nullptr : P.U, P.logQ, P.C, P.dvec, P.logR};
  const int cnt[8] = {M * Din, D, D * Din, M * D, D, D * Dy, Dy, Dy * Dy};
  int maxn = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) maxn = max(maxn, arr[k] ? cnt[k] : 0);
  double part[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int base = 0; base < maxn; base += nth) {
    const int i = base + tid;
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (arr[k] && i < cnt[k]) ? arr[k][i] - (k == 1 ? log005 : 0.0) : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) part[k] = fma(v[k], v[k], part[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 8; ++k) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
  __shared__ double wred[32][8];
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) wred[warp][k] = part[k];
  }
  __syncthreads();
  double tot[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    tot[k] = 0.0;
    for (int w = 0; w < nw; ++w) tot[k] += wred[w][k];        // fixed order: the same value in every thread, run to run
  }
  double pz = -0.5 * tot[0], ph = -0.5 * (tot[1] + tot[2]), pu = -0.5 * tot[3], hyp = -0.5 * (((tot[4] + tot[5]) + tot[6]) + tot[7]);
  if (!shared_priors) pz = ph = pu = hyp = 0.0;
  const bool stale = P.guard && P.guard[3] != 0;
  if (tid == 0 && P.guard) {
    // a failed factorisation must not be reusable (FFVD_FLAG_ASYNC callers never read the status): drop the record
    bool bad = false;
    for (int i = 0; i < nstatus; ++i) bad = bad || P.status[i] != 0;
    for (int i = 0; i < nstatus2; ++i) bad = bad || status2[i] != 0;        // blocked path: any matrix of the batch
    if (bad) P.guard[1] = 0;
  }
  // per-sample terms: a warp per sample (D <= 31 values of x_0)
  for (int s = warp; s < S; s += nw) {
    double x0 = (x0_prior && lane < D) ? P.X[(size_t)s * (T + 1) * D + lane] : 0.0;
    x0 = warp_sum(x0 * x0);
    const double px0 = -0.5 * x0;
    if (lane == 0) {
      const double* r = P.terms_raw + (size_t)s * FFVD_NTERMS_RAW;
      double t[6];
      t[0] = sc * (pu + ph + pz + px0 + hyp);
      t[1] = sc * r[FFVD_RAW_EMIS];
      t[2] = sc * r[FFVD_RAW_XQ];
      t[3] = sc * r[FFVD_RAW_TRACE];
      double ld = 0.0, qd = 0.0;
      if (collapsed && replicated)
        for (int dd = 0; dd < D; ++dd) { ld += P.collb[((size_t)s * D + dd) * 4 + 0]; qd += P.collb[((size_t)s * D + dd) * 4 + 1]; }
      t[4] = sc * ld;
      t[5] = sc * qd;
      if (stale) {                  // FFVD_FLAG_REUSE_KZZ with a Z / hyper-parameter content that the cached factors were not built from
#pragma unroll
        for (int k = 0; k < 6; ++k) t[k] = __longlong_as_double(0x7ff8000000000000ll);
      }
      if (O.terms) for (int k = 0; k < 6; ++k) O.terms[(size_t)s * 6 + k] = t[k];
      if (O.nll) O.nll[s] = ((t[0] + t[1]) + (t[2] + t[3])) + (t[4] + t[5]);
    }
  }
  if (flags & 4) return;
  if (O.g_Z) for (int i = tid; i < M * Din; i += nth) O.g_Z[i] = sc * (P.gZ[i] - (zprior ? npri * P.Z[i] : 0.0));
  if (O.g_U) for (int i = tid; i < M * D; i += nth) {
    const int m = i / D, d = i % D;
    O.g_U[i] = collapsed ? 0.0 : sc * (P.ubar[(size_t)d * Mp + m] - npri * P.U[i]);
  }
  if (O.g_logv) for (int i = tid; i < D; i += nth) O.g_logv[i] = sc * (P.gv[i] - npri * (P.logv[i] - log005));
  if (KIND == 0 && O.g_logl) for (int i = tid; i < D * Din; i += nth) O.g_logl[i] = sc * (P.gl[i] - npri * P.logl[i]);
  if (O.g_logQ) for (int i = tid; i < D; i += nth) {
    double rep = 0.0;
    if (collapsed && replicated)
      for (int ss = 0; ss < S; ++ss) rep += P.collb[((size_t)ss * D + i) * 4 + 2];        // H-dependent part, samples in index order
    O.g_logQ[i] = sc * ((P.gQ[i] + rep) - npri * P.logQ[i]);
  }
  if (O.g_C) for (int i = tid; i < D * Dy; i += nth) O.g_C[i] = sc * (P.gC[i] - npri * P.C[i]);
  if (O.g_d) for (int i = tid; i < Dy; i += nth) O.g_d[i] = sc * (P.gd[i] - npri * P.dvec[i]);
  if (O.g_logR) for (int i = tid; i < Dy * Dy; i += nth) O.g_logR[i] = sc * ((i < Dy ? P.gR[i] : 0.0) - npri * P.logR[i]);
}

}  // namespace ffvd
