// Blocked (64 x 64) Cholesky + triangular inverse on a pool of Mp x Mp matrices (Mp multiple of 64),
// for M too large for the single-CTA shared-memory path of kzz_prep / collapsed_chol.
// Right-looking factorisation: per block column one potrf (diagonal block, shared memory), one TRSM
// panel launch and one SYRK trailing-update launch, all FP64 DMMA.  The inverse X = L^{-1} is
// computed by independent CTAs per block column (block forward substitution).
// Replaces tf.linalg.cholesky + triangular_solve(L^T, I) of conditionals_multi_output.py:162-166.
#pragma once
#include "ffvd_common.cuh"

namespace ffvd {

// c (64x64 distributed over 8 warps: warp (wm,wn) owns rows wm*32.., cols wn*16..) += A[64xK] * op(B)
// A: global, row-major (lda).  TRANS_B: B is [64 x K] row-major (ldb) and enters transposed; else B is [K x 64].
// As/Bs: shared staging  As[64][20], Bs[16][68].
template <bool TRANS_B>
__device__ __forceinline__ void block_gemm_acc(double (&c)[4][2][2], const double* __restrict__ A, int lda,
                                               const double* __restrict__ B, int ldb, int K, double (*As)[20],
                                               double (*Bs)[68]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  // the next 16-deep slab travels (registers) while the current one is multiplied out of shared memory
  double ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      ra[i] = A[(size_t)(idx >> 4) * lda + k0 + (idx & 15)];
      rb[i] = TRANS_B ? B[(size_t)(idx >> 4) * ldb + k0 + (idx & 15)] : B[(size_t)(k0 + (idx >> 6)) * ldb + (idx & 63)];
    }
  };
  if (K > 0) fetch(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + 256 * i;
      As[idx >> 4][idx & 15] = ra[i];
      if (TRANS_B) Bs[idx & 15][idx >> 4] = rb[i];
      else Bs[idx >> 6][idx & 63] = rb[i];
    }
    __syncthreads();
    if (k0 + 16 < K) fetch(k0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[wm * 32 + 8 * i + g][kk + q];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bs[kk + q][wn * 16 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
}

__device__ __forceinline__ void block_zero(double (&c)[4][2][2]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
}

// visit the (row, col) pairs this lane owns in the 64x64 block
template <class F>
__device__ __forceinline__ void block_foreach(double (&c)[4][2][2], F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) f(wm * 32 + 8 * i + g, wn * 16 + 8 * j + 2 * q + e, c[i][j][e]);
}

// Factor the diagonal block kb of every matrix, store L_kk (lower) back and inv(L_kk) into Dinv[b][kb].
// grid (nbatch); block 256.  status[b] = first failing pivot (1-based, global index) if it is < M.
// The 64 x 64 block goes through the register-resident routines of the small-M path (prep_post.cuh: panel Cholesky,
// row-block forward substitution): this kernel is launched once per block column, strictly one after the other, and
// with a textbook column loop (three barriers per column, one thread per column of the inverse) it cost ~100 us a time.
// dynamic smem: L 64 x 65, X staging 64 x 64 (+ 64 spare: the panel scratch lives there during the factorisation), 64 x 1/L_ii
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, double* __restrict__ Dinv, int* __restrict__ status,
                                                          int kb, int M, int Mp) {
  extern __shared__ __align__(16) double dsm[];
  double* L = dsm;                                       // 64 x 65
  double* Xs = dsm + 64 * 65;                            // 64 x 64
  double* dinv = dsm + 2 * 64 * 65;                      // 64
  __shared__ int flag;
  const int b = blockIdx.x, tid = threadIdx.x, nblk = Mp / 64;
  double* Ab = A + (size_t)b * Mp * Mp + (size_t)(kb * 64) * Mp + kb * 64;
  for (int idx = tid; idx < 64 * 64; idx += 256) L[(idx >> 6) * 65 + (idx & 63)] = Ab[(size_t)(idx >> 6) * Mp + (idx & 63)];
  __syncthreads();
  const int st = chol_panel_smem(L, 65, 64, Xs, &flag);
  if (st != 0) {
    if (tid == 0 && kb * 64 + st <= M && status[b] == 0) status[b] = kb * 64 + st;
    return;
  }
  for (int idx = tid; idx < 64 * 64; idx += 256) {
    const int i = idx >> 6, k = idx & 63;
    Ab[(size_t)i * Mp + k] = (k <= i) ? L[i * 65 + k] : 0.0;
  }
  tri_inverse_smem(L, 65, 64, Xs, 64, dinv, Dinv + ((size_t)b * nblk + kb) * 64 * 64, nullptr, 64);
}

// Panel: A[i,kb] <- A[i,kb] * inv(L_kk)^T for block rows i > kb.  grid (nblk-kb-1, nbatch); block 256.
__global__ void __launch_bounds__(256) trsm_panel_kernel(double* __restrict__ A, const double* __restrict__ Dinv, int kb, int Mp) {
  __shared__ __align__(16) double As[64][20];
  __shared__ __align__(16) double Bs[16][68];
  const int b = blockIdx.y, i = kb + 1 + blockIdx.x, nblk = Mp / 64;
  double* Aik = A + (size_t)b * Mp * Mp + (size_t)(i * 64) * Mp + kb * 64;
  const double* Db = Dinv + ((size_t)b * nblk + kb) * 64 * 64;
  double c[4][2][2];
  block_zero(c);
  block_gemm_acc<true>(c, Aik, Mp, Db, 64, 64, As, Bs);
  __syncthreads();
  block_foreach(c, [&](int r, int cc, double v) { Aik[(size_t)r * Mp + cc] = v; });
}

// Trailing update: A[i,j] -= A[i,kb] A[j,kb]^T for kb < j <= i.  grid (npairs, nbatch); block 256.
__global__ void __launch_bounds__(256) syrk_trailing_kernel(double* __restrict__ A, int kb, int Mp) {
  __shared__ __align__(16) double As[64][20];
  __shared__ __align__(16) double Bs[16][68];
  const int b = blockIdx.y;
  int ti = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= (int)blockIdx.x) ++ti;
  while (ti * (ti + 1) / 2 > (int)blockIdx.x) --ti;
  const int tj = blockIdx.x - ti * (ti + 1) / 2;
  const int i = kb + 1 + ti, j = kb + 1 + tj;
  double* Ab = A + (size_t)b * Mp * Mp;
  double c[4][2][2];
  block_zero(c);
  block_gemm_acc<true>(c, Ab + (size_t)(i * 64) * Mp + kb * 64, Mp, Ab + (size_t)(j * 64) * Mp + kb * 64, Mp, 64, As, Bs);
  double* Aij = Ab + (size_t)(i * 64) * Mp + j * 64;
  block_foreach(c, [&](int r, int cc, double v) { Aij[(size_t)r * Mp + cc] -= v; });
}

// X = L^{-1} by block forward substitution, one CTA per (block column j, matrix).
//   X[j,j] = inv(L_jj);  X[i,j] = -inv(L_ii) * sum_{k=j}^{i-1} L[i,k] X[k,j]
// grid (nblk, nbatch); block 256.
__global__ void __launch_bounds__(256) trtri_column_kernel(const double* __restrict__ Lf, const double* __restrict__ Dinv,
                                                            double* __restrict__ X, int Mp) {
  __shared__ __align__(16) double As[64][20];
  __shared__ __align__(16) double Bs[16][68];
  extern __shared__ __align__(16) double dsm[];          // 64 x 68 doubles
  double (*Ts)[68] = reinterpret_cast<double (*)[68]>(dsm);
  const int b = blockIdx.y, j = blockIdx.x, nblk = Mp / 64, tid = threadIdx.x;
  const double* Lb = Lf + (size_t)b * Mp * Mp;
  double* Xb = X + (size_t)b * Mp * Mp;
  const double* Db = Dinv + (size_t)b * nblk * 64 * 64;
  for (int idx = tid; idx < 64 * 64; idx += 256)
    Xb[(size_t)(j * 64 + (idx >> 6)) * Mp + j * 64 + (idx & 63)] = Db[(size_t)j * 64 * 64 + idx];
  for (int i = j + 1; i < nblk; ++i) {
    __syncthreads();           // X[k,j] of earlier block rows are visible to the whole CTA
    double c[4][2][2];
    block_zero(c);
    for (int k = j; k < i; ++k)
      block_gemm_acc<false>(c, Lb + (size_t)(i * 64) * Mp + k * 64, Mp, Xb + (size_t)(k * 64) * Mp + j * 64, Mp, 64, As, Bs);
    __syncthreads();
    block_foreach(c, [&](int r, int cc, double v) { Ts[r][cc] = v; });
    __syncthreads();
    // out = -inv(L_ii) * T  : A operand from Dinv (global), B operand from the shared T
    double c2[4][2][2];
    block_zero(c2);
    {
      const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3, wm = warp >> 2, wn = warp & 3;
      const double* Di = Db + (size_t)i * 64 * 64;
      for (int k0 = 0; k0 < 64; k0 += 16) {
        __syncthreads();
        for (int idx = tid; idx < 64 * 16; idx += 256) As[idx >> 4][idx & 15] = Di[(size_t)(idx >> 4) * 64 + k0 + (idx & 15)];
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
          double a[4], bb[2];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) a[ii] = As[wm * 32 + 8 * ii + g][kk + q];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) bb[jj] = Ts[k0 + kk + q][wn * 16 + 8 * jj + g];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) dmma884(c2[ii][jj][0], c2[ii][jj][1], a[ii], bb[jj]);
        }
      }
    }
    double* Xij = Xb + (size_t)(i * 64) * Mp + j * 64;
    block_foreach(c2, [&](int r, int cc, double v) { Xij[(size_t)r * Mp + cc] = -v; });
  }
}

// XT = X^T, and clear everything that belongs to the padding (rows / cols >= M) in both.
// grid (Mp/32, Mp/32, nbatch); block (32, 8).
__global__ void transpose_pad_kernel(double* __restrict__ X, double* __restrict__ XT, int M, int Mp) {
  __shared__ double t[32][33];
  const size_t off = (size_t)blockIdx.z * Mp * Mp;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int row = r0 + r, col = c0 + threadIdx.x;
    double v = X[off + (size_t)row * Mp + col];
    if (row >= M || col >= M || col > row) v = 0.0;
    X[off + (size_t)row * Mp + col] = v;
    t[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) XT[off + (size_t)(c0 + r) * Mp + r0 + threadIdx.x] = t[threadIdx.x][r];
}

}  // namespace ffvd
