// Isolates the triangular contraction of the fused kernel (tile_gemm<8,2,TRI>) and its SYRK (syrk_tile) in a loop, to see
// how far from the DMMA issue peak the loops themselves run, away from every other phase of a work item.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/gemm_probe tools/probe/gemm_probe.cu
#include <cstdio>
#include "../../ffvd_b200/csrc/fused.cuh"
using namespace ffvd;

template <int RB, int NGW, int TRI, int VARIANT>
__global__ void __launch_bounds__(256, 1) k_gemm(const double* __restrict__ B, int Mp, int iters, double* out) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int lda = Mp + 4;
  for (int i = tid; i < 8 * RB * lda; i += 256) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  double acc[NGW][RB][4];
  double s = 0.0;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 0) {
      tile_gemm<RB, NGW, TRI>(acc, sm, lda, B, Mp, warp, g, q, [](double x, int, int) { return x; });
#pragma unroll
      for (int ng = 0; ng < NGW; ++ng)
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) s += acc[ng][rb][0] + acc[ng][rb][3];
    } else if (VARIANT == 4) {
      // 64x32 units (8x4 blocks, 12 fragment loads per 32 DMMAs), all blocks formed (loop-efficiency probe only)
      const int ns = Mp >> 6, ncol = Mp >> 5;
      int total = 0;
      for (int i = 0; i < ns; ++i) total += 2 * i + 2;
      for (int it2 = 0;; ++it2) {
        const int u = it2 * 8 + ((it2 & 1) ? (7 - warp) : warp);
        if (u >= total) break;
        int si = 0, acc_u = 0;
        while (acc_u + 2 * si + 2 <= u) { acc_u += 2 * si + 2; ++si; }
        const int m0 = 64 * si, n0 = 32 * (u - acc_u);
        (void)ncol;
        double c[8][4][2];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;
#pragma unroll 2
        for (int k0 = 0; k0 < 8 * RB; k0 += 4) {
          double av[8], bv[4];
          const double* row = sm + (k0 + q) * lda + g;
#pragma unroll
          for (int a = 0; a < 8; ++a) av[a] = row[m0 + 8 * a];
#pragma unroll
          for (int b = 0; b < 4; ++b) bv[b] = row[n0 + 8 * b];
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(c[a][b][0], c[a][b][1], av[a], bv[b]);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) s += c[a][b][0];
      }
    } else if (VARIANT == 3) {
      // tile-row strips: warp handles rows ti, and column tiles tj in pairs (two 32x32 accumulators at once)
      const int nt = Mp >> 5;
      // strips (ti, pair p): pairs cover tj = 2p, 2p+1 <= ti ; count per ti = ti/2 + 1
      int total = 0;
      for (int ti = 0; ti < nt; ++ti) total += ti / 2 + 1;
      for (int i = 0;; ++i) {
        const int u = i * 8 + ((i & 1) ? (7 - warp) : warp);
        if (u >= total) break;
        int ti = 0, acc_u = 0;
        while (acc_u + ti / 2 + 1 <= u) { acc_u += ti / 2 + 1; ++ti; }
        const int p = u - acc_u;
        const int m0 = 32 * ti, n0 = 64 * p;
        const bool two = (2 * p + 1 <= ti);
        double c[4][8][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) c[a][b][0] = c[a][b][1] = 0.0;
#pragma unroll 4
        for (int k0 = 0; k0 < 8 * RB; k0 += 4) {
          double av[4], bv[8];
          const double* row = sm + (k0 + q) * lda + g;
#pragma unroll
          for (int a = 0; a < 4; ++a) av[a] = row[m0 + 8 * a];
#pragma unroll
          for (int b = 0; b < 8; ++b) bv[b] = (b < 4 || two) ? row[n0 + 8 * b] : 0.0;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(c[a][b][0], c[a][b][1], av[a], bv[b]);
          if (two) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 4; b < 8; ++b) dmma884(c[a][b][0], c[a][b][1], av[a], bv[b]);
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) s += c[a][b][0];
      }
    } else {
      // SYRK units of this warp, without the flush (VARIANT 1) -- see syrk_flush for the unit list
      const int nt = Mp >> 5, nfull = nt * (nt - 1) / 2, nunits = nfull + nt;
      for (int i = 0;; ++i) {
        const int u = i * 8 + ((i & 1) ? (7 - warp) : warp);
        if (u >= nunits) break;
        int m0, n0; bool diag = u >= nfull;
        if (!diag) {
          int ti = (int)((sqrtf(8.0f * (float)u + 1.0f) + 1.0f) * 0.5f);
          while (ti * (ti - 1) / 2 > u) --ti;
          while ((ti + 1) * ti / 2 <= u) ++ti;
          m0 = 32 * ti; n0 = 32 * (u - ti * (ti - 1) / 2);
        } else { m0 = n0 = 32 * (u - nfull); }
        double c[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 8 * RB; k0 += 4) {
          if (VARIANT == 1 && (k0 & 4)) { asm volatile("" ::: "memory"); }
          double av[4], bv[4];
          const double* row = sm + (k0 + q) * lda + g;
#pragma unroll
          for (int a = 0; a < 4; ++a) av[a] = row[m0 + 8 * a];
#pragma unroll
          for (int b = 0; b < 4; ++b) bv[b] = row[n0 + 8 * b];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (!diag || b <= a) dmma884(c[a][b][0], c[a][b][1], av[a], bv[b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) s += c[a][b][0];
      }
    }
  }
  out[blockIdx.x * 256 + tid] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  const int Mp = 256, RB = 8, NGW = 2, iters = 2000, nsm = 148;
  double *B, *out;
  cudaMalloc(&B, sizeof(double) * Mp * Mp * 8); cudaMemset(B, 0, sizeof(double) * Mp * Mp * 8);
  cudaMalloc(&out, sizeof(double) * nsm * 256);
  const size_t smem = (size_t)8 * RB * (Mp + 4) * 8;
  auto run = [&](auto kern, const char* name, double dmma_per_iter) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float ms = timeit([&] { kern<<<nsm, 256, smem>>>(B, Mp, iters, out); });
    const double tf = dmma_per_iter * iters * nsm * 512.0 / ms * 1e-9;
    printf("%-46s %8.3f ms  %6.2f TFLOP/s executed (%.1f%% of 37.15)\n", name, ms, tf, 100 * tf / 37.15);
  };
  const double G = Mp / 16.0;
  const double tri_dmma = 128.0 * G * (G + 1) * (8 * RB) / 256.0;          // MACs per row x rows / 256
  const int nt = Mp / 32;
  const double syrk_dmma = (16.0 * nt * (nt - 1) / 2 + 10.0 * nt) * (8 * RB / 4);
  run(k_gemm<RB, NGW, +1, 0>, "A = K L^-T   (upper-triangular operand)", tri_dmma);
  run(k_gemm<RB, NGW, -1, 0>, "Kbar = A L^-1 (lower-triangular operand)", tri_dmma);
  run(k_gemm<RB, NGW, 0, 0>, "dense operand", (double)Mp * Mp * (8 * RB) / 256.0);
  run(k_gemm<RB, NGW, 0, 1>, "SYRK units without flush (k loop fully unrolled)", syrk_dmma);
  {
    double strips = 0;                    // DMMAs of the strip variant: full 16-block tiles everywhere (diagonal not trimmed)
    for (int ti = 0; ti < nt; ++ti) strips += (ti + 1) * 16.0;
    run(k_gemm<RB, NGW, 0, 3>, "SYRK 32x64 strips (diagonal tiles untrimmed)", strips * (8 * RB / 4));
  }
  {
    double units = 0;
    for (int i = 0; i < Mp / 64; ++i) units += 2 * i + 2;
    run(k_gemm<RB, NGW, 0, 4>, "SYRK 64x32 units (untrimmed)", units * 32.0 * (8 * RB / 4));
  }
  return 0;
}
