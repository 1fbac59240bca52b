// Shared device-side definitions for libffvd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FFVD_NTHREADS 256
#define FFVD_NWARPS 8
#define FFVD_XLD 36            // row stride of the x-tile arrays in shared memory (doubles); 36 = 4 mod 16 keeps the
                               // DMMA fragment reads (row q, column g) and the per-row reads conflict free per half-warp
#define FFVD_XCOLS 32          // padded column count of X~ = [x, ctrl, 1, 0...]
#define FFVD_MAX_DIN 31
#define FFVD_ZTS_ROWS 41         // rows of the scaled Z~^T copy: 0..Din-1 z/l, zeros up to 39 (branch-free padded steps), row 40 = -|z~|^2/2
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
  double *ZT;          // [32][Mp]    Z~^T: transposed, zero padded inducing inputs, then a row of ones (m < M), zeros
  double *Zf;          // [Mp/4][4][32] the same matrix in DMMA B-fragment order: entry ((m>>2)*4 + (jd>>3))*32 + (jd&7)*4 + (m&3)
  double *ZTs;         // [nk][FFVD_ZTS_ROWS][Mp] SE: per-kernel scaled copy z~ = z/l (rows 0..Din-1), zero rows, row 40 = -1/2 |z~_m|^2
  double *hyp;         // [nk][72]   per kernel: 1/l^2 [0..31], 1/l [32..63], v [64]           (hyper_kernel)
  double *hq;          // [D][4]     per output dim: Q, 1/Q, log Q                              (hyper_kernel)
  double *UT;          // [D][Mp]    U transposed, zero padded (coalesced staging of u_d)       (hyper_kernel)
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
  unsigned long long *guard;   // [4] content guard of the cached K(Z,Z) factors (FFVD_FLAG_REUSE_KZZ): running hash of
                               //     Z / logv / logl, hash at factorisation time, block counter, stale flag (hyper_kernel)
  double *cond_mean, *cond_var;   // conditional(): (N,R) outputs
  const double *qmat;  // conditional() with q_sqrt: qmode 3 -> [nq][Mp][Mp] zero padded factors (var += |Q^T a|^2);
                       //                            qmode 2 -> [M][R] per-point scales      (var += sum (q_m a_m)^2)
  int qmode, nq;       // nq = 1: one factor shared by all outputs (the reference's [:, :, 0] indexing), else R
  int S, T, D, Din, nc, Dy, M, Mp;
  int Dx;              // columns of X (== D for the nll paths, == Din for conditional())
  int xrows;           // valid rows of X per sample (T+1 for nll, N for conditional())
  int hs;              // hyper-parameter / Linv stride per output dim: 1 = list of D kernels, 0 = one shared kernel
  int ntiles;          // tiles per sample
  int dblk;            // output dims per block of the work-item order (see fused_kernel)
  long long item_begin;   // first work item of this problem (prefix sum)
  long long nitems;       // D * S * ntiles
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// exp(x) for x <= ~0 (the SE kernel argument -r^2/2), branch free, N independent evaluations advanced in lock step
// so that the FP64 pipe always has N independent FMAs to issue (the pipe is narrow: two warps per scheduler cannot
// hide its latency with one dependent Horner chain each).
// x = n ln2 + r, degree-13 Taylor polynomial on |r| <= ln2/2 (truncation 4e-18), 2^n applied to the exponent bits.
// <= 1 ulp from the correctly rounded result on [-700, 1e-9] (checked on the host against libm).  The argument is
// clamped at -745 first (one DMNMX): the low word of t must hold n, and without the clamp an argument below ~-1.5e9
// (|x/l - z/l| > 5e4, e.g. a diverging chain on logl) would wrap n and return a huge / Inf / NaN kernel value instead of
// 0.  Below -708 the result saturates near 1e-308, i.e. 0 at the scale of every quantity the kernels form.
template <int N>
__device__ __forceinline__ void exp_nonpos_n(double (&x)[N]) {
  double r[N], p[N];
  int n[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x[i] = fmax(x[i], -745.0);
    const double t = fma(x[i], 1.4426950408889634, 6755399441055744.0);
    n[i] = max(__double2loint(t), -1022);      // x < -708: the scale saturates at 2^-1022 (result ~ 1e-308 ~ 0)
    const double nd = t - 6755399441055744.0;
    r[i] = fma(nd, -6.93147180369123816490e-01, x[i]);
    r[i] = fma(nd, -1.90821492927058770002e-10, r[i]);
    p[i] = fma(1.0 / 6227020800.0, r[i], 1.0 / 479001600.0);
  }
  const double cf[12] = {1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
                         1.0 / 120.0,      1.0 / 24.0,      1.0 / 6.0,      0.5,           1.0,          1.0};
#pragma unroll
  for (int k = 0; k < 12; ++k)
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], cf[k]);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = __hiloint2double(__double2hiint(p[i]) + (n[i] << 20), __double2loint(p[i]));
}

__device__ __forceinline__ double exp_nonpos(double x) {
  double a[1] = {x};
  exp_nonpos_n<1>(a);
  return a[0];
}

// Fire-and-forget FP64 add to GLOBAL memory.  atomicAdd() on a pointer whose address space the compiler cannot prove
// (every pointer loaded from DevProblem) becomes ATOM.E.ADD + a predicate wait + shared/local CAS fall-backs, i.e. a
// synchronous L2 round trip per call; the explicit red.global is one REDG with no return value.
__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v));
}

// 16-byte read-only load for the L^{-1} / L^{-T} operand streams: no L1 allocation (each element is used once per
// tile) and an L2 evict-last policy -- the D matrices (8 MB at C3) are re-read by every tile for the whole kernel while
// X / x-bar (hundreds of MB) stream through the same L2; without the hint the streams kept evicting them and the
// kernel re-fetched them from HBM ~700 times (ncu: 6.4 GB of DRAM reads against 0.8 GB of X + x-bar).
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  double2 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}

// 32-byte (256-bit) global accesses, sm_100a LDG/STG.E.256: one lane moves four consecutive doubles, so the four lanes
// that own a 16-column group of a tile row write / read one full 128-byte line (two 16-byte accesses per lane leave every
// 32-byte sector half written per instruction).  Addresses must be 32-byte aligned.
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(__cvta_generic_to_global(p)), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ldg256_nc(const double* p, double2& lo, double2& hi) {
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(lo.x), "=d"(lo.y), "=d"(hi.x), "=d"(hi.y) : "l"(__cvta_generic_to_global(p)));
}
__device__ __forceinline__ void ldg256_cg(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(__cvta_generic_to_global(p)));
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
