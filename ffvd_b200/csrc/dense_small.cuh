// Small dense FP64 helpers for the API-surface branches of the conditional (full_cov=True, q_sqrt with white=False,
// return_Lm: conditionals.py:27-66 / conditionals_multi_output.py:44-69).  These branches are prediction-time conveniences
// on a handful of points -- never the hot path -- so they are plain tiled kernels on explicit matrices, following the
// reference op for op (A = L^{-1} Kmn; fvar = Knn - A^T A; second substitution when not white; LTA = q^T A ...).
#pragma once
#include "ffvd_common.cuh"

namespace ffvd {

// C (m x n, row-major, ldc) = alpha * op(A) * op(B) + beta * C;  op(A) is m x k (ta: A is stored k x m), op(B) is k x n.
// grid (ceil(n/32), ceil(m/32)); block (32, 8): each thread four rows of a 32 x 32 tile.
__global__ void __launch_bounds__(256) dgemm_small_kernel(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda, int ta,
                                                          const double* __restrict__ B, int ldb, int tb, int m, int n, int k,
                                                          double alpha, double beta) {
  __shared__ double As[32][33], Bs[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < k; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      const int ar = row0 + r, ac = k0 + tx;            // op(A)[ar][ac]
      As[r][tx] = (ar < m && ac < k) ? (ta ? A[(size_t)ac * lda + ar] : A[(size_t)ar * lda + ac]) : 0.0;
      const int br = k0 + r, bc = col0 + tx;            // op(B)[br][bc]
      Bs[r][tx] = (br < k && bc < n) ? (tb ? B[(size_t)bc * ldb + br] : B[(size_t)br * ldb + bc]) : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double b = Bs[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fma(As[ty + 8 * i][kk], b, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty + 8 * i, cc = col0 + tx;
    if (r < m && cc < n) {
      double* p = C + (size_t)r * ldc + cc;
      *p = alpha * acc[i] + (beta != 0.0 ? beta * *p : 0.0);
    }
  }
}

// out[n * ldo] = (add ? out : base[n]) + sign * sum_m (A[m][n] * (scale ? scale[m * lds] : 1))^2      (A is M x N)
__global__ void colsumsq_kernel(double* __restrict__ out, int ldo, const double* __restrict__ base, const double* __restrict__ A, int M, int N,
                                const double* __restrict__ scale, int lds, double sign, int add) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int m = 0; m < M; ++m) {
      const double a = A[(size_t)m * N + n] * (scale ? scale[(size_t)m * lds] : 1.0);
      s = fma(a, a, s);
    }
    double* p = out + (size_t)n * ldo;
    *p = (add ? *p : base[n]) + sign * s;
  }
}

// dst[m][n] = src[m][n] * scale[m * lds]
__global__ void scale_rows_kernel(double* __restrict__ dst, const double* __restrict__ src, int M, int N, const double* __restrict__ scale, int lds) {
  const size_t tot = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * scale[(i / N) * lds];
}

// dst (M x M, packed) = lower triangle of src (ld = Mp), zeros above the diagonal
__global__ void tril_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, int M, int Mp) {
  const size_t tot = (size_t)M * M;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / M), c = (int)(i % M);
    dst[i] = (c <= r) ? src[(size_t)r * Mp + c] : 0.0;
  }
}

}  // namespace ffvd
