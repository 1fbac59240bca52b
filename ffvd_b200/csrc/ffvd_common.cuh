// Shared device-side definitions for libffvd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FFVD_NTHREADS 256
#define FFVD_NWARPS 8
#define FFVD_XLD 33            // row stride of the x-tile arrays in shared memory (doubles)
#define FFVD_XCOLS 32          // padded column count of X~ = [x, ctrl, 1, 0...]
#define FFVD_MAX_DIN 31
#define FFVD_NTERMS_RAW 8      // raw per-sample sums, see below

// raw per-sample term slots (sums of J pieces, before the -1/T scaling)
#define FFVD_RAW_XQ 0          // sum_t,d [-1/2 r^2/Q - 1/2 logQ]
#define FFVD_RAW_TRACE 1       // sum_t,d [-1/2 sigma^2/Q]
#define FFVD_RAW_EMIS 2        // sum_t loglik_t
#define FFVD_RAW_LOGDET 3      // collapsed: sum_d -1/2 logdet H_d
#define FFVD_RAW_QUAD 4        // collapsed: sum_d  1/2 b^T H^{-1} b

// One GPSSM problem as the kernels see it.  All pointers are device pointers.
struct DevProblem {
  // inputs
  const double *X, *Z, *U, *logv, *logl, *logQ, *C, *dvec, *logR, *Y, *ctrl;
  // derived inputs (written by kzz_prep)
  double *ZT;          // [Din][Mp]   transposed, zero padded inducing inputs
  double *Linv;        // [D][Mp][Mp] L^{-1}  (lower), zero padded
  double *LinvT;       // [D][Mp][Mp] L^{-T}  (upper)
  // accumulators (zeroed before each evaluation)
  double *Sacc;        // [D][Mp][Mp] sum_t a a^T (lower 32x32 tiles), later Kbar_zz
  double *Wk;          // [D][Mp][Mp] scratch for the M^3 products
  double *ubar;        // [D][Mp]     sum_t e_t a_t   (collapsed: b = F^T delta, unscaled)
  double *gZ;          // [M][Din]    raw dJ/dZ
  double *gl;          // [D][Din]    raw dJ/dlogl
  double *gv;          // [D]
  double *gQ;          // [D]
  double *gC;          // [D][Dy]
  double *gd;          // [Dy]
  double *gR;          // [Dy]   (row 0 of logR)
  double *terms_raw;   // [S][FFVD_NTERMS_RAW]
  double *gX;          // [S][T+1][D] raw dJ/dX  (the caller's g_X tensor, scaled in place at the end)
  // collapsed-only small vectors / matrices
  double *cvec;        // [D][Mp]  c = H^{-1} b
  double *wvec;        // [D][Mp]  w' = L^{-T} c / Q
  double *Nmat;        // [D][Mp][Mp]  N = L^{-T} Mat' L^{-1}
  double *Hx, *HxT;    // [S*D][Mp][Mp] collapsed scratch: L_H^{-1}, L_H^{-T} / Mat' S
  double *rs;          // [nb][Mp] row sums of Wz
  int *status;         // [D] 0 ok, else 1-based failing pivot
  double *cond_mean, *cond_var;   // conditional(): (N,R) outputs
  int S, T, D, Din, nc, Dy, M, Mp;
  int Dx;              // columns of X (== D for the nll paths, == Din for conditional())
  int xrows;           // valid rows of X per sample (T+1 for nll, N for conditional())
  int hs;              // hyper-parameter / Linv stride per output dim: 1 = list of D kernels, 0 = one shared kernel
  int ntiles;          // tiles per sample
  long long item_begin;   // first work item of this problem (prefix sum)
  long long nitems;       // D * S * ntiles
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// column-group (16 columns) owned by `warp` as its i-th group: zig-zag so that the
// triangular work (proportional to the group index) is balanced across the 8 warps.
__device__ __forceinline__ int group_index(int warp, int i) {
  return (i & 1) ? (i * FFVD_NWARPS + (FFVD_NWARPS - 1 - warp)) : (i * FFVD_NWARPS + warp);
}
