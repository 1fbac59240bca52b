// Shared device-side definitions for libffvd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Index checks of the fused kernels (make check: -DFFVD_BOUNDS_CHECK, tools/bounds_check.py).  compute-sanitizer is not
// available on the development pool, so the checks the sanitizer would make on the hand-computed shared / global indices are
// asserted in a diagnostic build instead: a violation prints the condition and traps (the call returns FFVD_E_CUDA).
#ifdef FFVD_BOUNDS_CHECK
#include <cstdio>
#define FFVD_ASSERT(cond)                                                                                              \
  do {                                                                                                                 \
    if (!(cond)) {                                                                                                     \
      printf("FFVD_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                                                        \
    }                                                                                                                  \
  } while (0)
#else
#define FFVD_ASSERT(cond) do { } while (0)
#endif

#define FFVD_NTHREADS 256
#define FFVD_NWARPS 8
#define FFVD_XLD 36            // row stride of the x-tile arrays in shared memory (doubles); 36 = 4 mod 16 keeps the
                               // DMMA fragment reads (row q, column g) and the per-row reads conflict free per half-warp
#define FFVD_XCOLS 32          // padded column count of X~ = [x, ctrl, 1, 0...]
#define FFVD_MAX_DIN 31
#define FFVD_ZTS_ROWS 41         // rows of the scaled, augmented Z~^T copy: 0..Din-1 z/l, row Din = 1, row Din+1 = -|z~|^2/2, zeros (k-steps of 4)
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
  double *ZTs;         // [nk][FFVD_ZTS_ROWS][Mp] SE: per-kernel scaled copy z~ = z/l (rows 0..Din-1), row Din = 1, row Din+1 = -1/2 |z~_m|^2, zero rows
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
  double *gZd;         // [D][Mp][8 NBM] raw W^T [Xc,1] sums per output dim (NBM = ceil((Din+1)/8)): the K(X,Z) part of dJ/dZ, dJ/dlogl before
                       //             scaling; folded into gZ / gl by zbar_post_kernel.  Lives in the S region (per-CTA copies in deterministic mode)
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
  // FFVD_FLAG_DETERMINISTIC (all null / zero otherwise): private accumulator copies and x-bar planes, see det_ptr1 / det_ptr2
  char *det1, *det_base1;      // per-CTA copies of the S region [det_base1, +det_stride1): element p lives at det1 + cta*stride + (p - base)
  char *det2, *det_base2;      // per-(CTA, warp) copies of the small-accumulator region (u-bar, small gradients, raw terms)
  long long det_stride1, det_stride2;
  double *gXp;                 // [3][S][T+1][D] x-bar planes: +-e / +-gx terms, emission, LinearK diagonal term (the W back-propagation
  long long gXp_stride;        //   goes to g_X itself); summed in a fixed order by finalize_kernel
  double *kzzpart;             // kzz_bwd_fused partials (deterministic mode): per batch entry [M][Din] | [blocks][Din] | [blocks]
  double *collb;               // [nb][4] collapsed: per (s,d) -1/2 logdet H, quadratic term, H-part of dJ/dlogQ (plain stores, summed by finalize)
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
  int dl;              // 1: one work item per (sample, tile, d); > 1 (= dblk): one per (sample, tile, block of dims)
  long long item_begin;   // first work item of this problem (prefix sum)
  long long nitems;       // work items: D * S * ntiles (dl == 1) or ceil(D / dblk) * S * ntiles (dl > 1)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// exp(x) for x <= ~0 (the SE kernel argument -r^2/2), branch free, N independent evaluations advanced in lock step
// so that the FP64 pipe always has N independent FMAs to issue (the pipe is narrow: two warps per scheduler cannot
// hide its latency with one dependent Horner chain each).
// Table method: x = (64 n + j) ln2/64 + r with |r| <= ln2/128, exp(x) = 2^n * T[j] * exp(r), T[j] = 2^(j/64) correctly
// rounded (64 doubles, staged in shared memory by the caller), exp(r) - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120)
// (truncation r^6/720 <= 3.5e-17 relative).  11 FP64 operations per value against 19 for the degree-13 polynomial this
// replaces (the K tile is FP64-pipe work that competes with the DMMAs); <= 1.5 ulp on [-745, 0] (tests: ffvd_debug_exp).
// The argument is clamped at -745 (high word: -745.99) first: the low word of t must hold the integer, and without the clamp an argument
// below ~-2e7 (|x/l - z/l| > 6e3, e.g. a diverging chain on logl) would wrap it and return a huge / Inf / NaN kernel
// value instead of 0.  Below -708 the scale saturates at 2^-1022 (result ~ 1e-308, i.e. 0 at the scale of every quantity
// the kernels form).
static __device__ const double g_exp2_tab[64] = {
  0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
  0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
  0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
  0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
  0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
  0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
  0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
  0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
  0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
  0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
  0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
  0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
  0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
  0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
  0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
  0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// tab may be pre-scaled by a positive factor v (the fused kernels stage v_d * 2^(j/64) per output dim, which saves the
// multiplication by the kernel variance); nmin then bounds the exponent shift from below so that the exponent field of
// v * 2^(j/64) * (1 + ...) cannot wrap: nmin = -1022 - min(ilogb(v), 0)  (exp_nmin).
__device__ __forceinline__ int exp_nmin(double v) {
  const int e = ((__double2hiint(v) >> 20) & 0x7ff) - 1023;
  return e < -500 ? 0 : -1022 - min(e, 0);      // v < 2^-500: the caller stages a zero table, and n = 0 keeps the zeros
}
template <int N>
__device__ __forceinline__ void exp_nonpos_n(double (&x)[N], const double* __restrict__ tab /* shared-memory copy of g_exp2_tab */,
                                             int nmin = -1022) {
  double r[N], q[N], tj[N];
  int n[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    // x = max(x, -745.99...) on the high word: for negative doubles a larger magnitude is a larger unsigned high word (one
    // integer min instead of an FP64 compare and two selects; NaN saturates like a very negative argument, as fmax did)
    x[i] = __hiloint2double((int)min((unsigned)__double2hiint(x[i]), 0xC0874800u), __double2loint(x[i]));
    const double t = fma(x[i], 92.33248261689366, 6755399441055744.0);     // 64 / ln 2, 1.5 * 2^52
    const int k = __double2loint(t);                                        // round(x * 64 / ln 2), two's complement
    const double kd = t - 6755399441055744.0;
    r[i] = fma(kd, -0x1.62e42fec00000p-7, x[i]);                            // ln2/64 head (30 bits: kd * head is exact)
    r[i] = fma(kd, -0x1.d1cf79abc9e3bp-38, r[i]);                           // ln2/64 tail
    tj[i] = tab[k & 63];
    n[i] = max(k >> 6, nmin);
    q[i] = fma(r[i], 1.0 / 120.0, 1.0 / 24.0);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 1.0 / 6.0);
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double r2 = r[i] * r[i];
    const double em1 = fma(r2, q[i], r[i]);
    const double p = fma(tj[i], em1, tj[i]);
    x[i] = __hiloint2double(__double2hiint(p) + (n[i] << 20), __double2loint(p));
  }
}

// Fire-and-forget FP64 add to GLOBAL memory.  atomicAdd() on a pointer whose address space the compiler cannot prove
// (every pointer loaded from DevProblem) becomes ATOM.E.ADD + a predicate wait + shared/local CAS fall-backs, i.e. a
// synchronous L2 round trip per call; the explicit red.global is one REDG with no return value.
__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v));
}

// The same under a predicate, as ONE predicated instruction: `if (c) red_add(...)` compiles to a divergence region
// (BSSY / BRA / BSYNC) around every call, which in the epilogues with a dozen conditional REDs per lane is most of the code.
// The address must be valid to FORM, not to access, when c is false.
__device__ __forceinline__ void red_add_if(double* p, double v, bool c) {
  asm volatile("{\n\t.reg .pred pr;\n\tsetp.ne.s32 pr, %2, 0;\n\t@pr red.global.add.f64 [%0], %1;\n\t}" ::"l"(__cvta_generic_to_global(p)), "d"(v), "r"((int)c));
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

__device__ __forceinline__ double2 ldg_stream2_v(const double* p) {     // same, pinned in program order (early prologue requests)
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
// Bulk asynchronous copy shared -> global through the TMA engine (SASS UBLKCP): the K tile's scratch copy leaves the SM without
// passing through registers or the LSU store queue.  16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_copy_s2g(double* gdst, const double* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_fence_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#define FFVD_KB0 3      // k-steps of the K-tile product whose group-0 operand fragments are requested as register loads ahead of the phase
__device__ __forceinline__ double2 ldg_nc2_v(const double* p) {            // 16-byte read-only load, pinned in program order
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(__cvta_generic_to_global(p)));
  return v;
}
__device__ __forceinline__ void prefetch_l1(const double* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p)));
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

// FFVD_FLAG_DETERMINISTIC: where an accumulator element lives for this CTA / warp.
// FP64 REDs from different CTAs (and from different warps of one CTA) land in an order that changes from run to run, so the
// shared sums differ in their last bits.  In deterministic mode every CTA adds into its OWN copy of the S region and every
// (CTA, warp) into its own copy of the small accumulators -- each element then has a single writer thread, whose REDs to one
// address apply in program order -- and det_reduce_kernel sums the copies in a fixed order.  Off: the shared accumulator itself.
#define FFVD_DET_MAX_WARPS 16
// The byte offsets are formed ONCE per work item (det_off1 / det_off2: zero when the mode is off) and added to the
// accumulator addresses (det_at): looking the mode up per RED cost two shared-memory loads and a divergence region per call.
__device__ __forceinline__ long long det_off1(const DevProblem& P) {
  return P.det1 ? (long long)((P.det1 + (size_t)blockIdx.x * P.det_stride1) - P.det_base1) : 0ll;
}
__device__ __forceinline__ long long det_off2(const DevProblem& P) {
  return P.det2 ? (long long)((P.det2 + ((size_t)blockIdx.x * FFVD_DET_MAX_WARPS + (threadIdx.x >> 5)) * P.det_stride2) - P.det_base2) : 0ll;
}
__device__ __forceinline__ double* det_at(double* p, long long off) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(p) + off);
}

// Element (row m, column jd) of one output dim's block of DevProblem::gZd (Mp x 8 NBM values, see contract_W / zbar_post_kernel)
__host__ __device__ __forceinline__ size_t zbar_index(int Mp, int m, int jd) {
  return ((size_t)(((jd >> 3) << 1) | (jd & 1)) * Mp + m) * 4 + ((jd & 7) >> 1);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// column-group (16 columns) owned by `warp` as its i-th group: zig-zag so that the
// triangular work (proportional to the group index) is balanced across the 8 warps.
// NCW = number of column warps of the tile (8; 4 for the two-CTAs-per-SM shapes).
template <int NCW>
__device__ __forceinline__ int group_index(int warp, int i) {
  return (i & 1) ? (i * NCW + (NCW - 1 - warp)) : (i * NCW + warp);
}
