// Hardware probe for the FFVD B200 design decisions (not product code).
// Measures, on one B200: FP64 DMMA (mma.sync m8n8k4) rate, DFMA rate, the two
// interleaved, red.global.add.f64 throughput, and cp.reduce.async.bulk add.f64
// throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
  double c0[NACC], c1[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
  double c[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NACC DMMAs + NF DFMAs per iteration, interleaved.
template <int NACC, int NF>
__global__ void k_mixed(double* out, int iters) {
  double c0[NACC], c1[NACC], f[NF];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      dmma884(c0[i], c1[i], a, b);
      if (i < NF) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory: warp tile 32 x 32, K loop over a smem panel.
// A tile: 32 x KP (ld = KP+4), B tile: 32(n) x KP (ld=KP+4) per warp-shared.
template <int KP>
__global__ void k_dmma_smem(double* out, int iters) {
  extern __shared__ double sm[];
  const int ld = KP + 4;
  double* As = sm;                 // 32 x ld
  double* Bs = sm + 32 * ld;       // 64 x ld  (8 warps -> 2 groups of n? keep 32 per warp pair)
  for (int i = threadIdx.x; i < (32 + 64) * ld; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  double c0[16], c1[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c0[i] = 0; c1[i] = 0; }
  const double* Bw = Bs + (warp & 1) * 32 * ld;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int k = 0; k < KP; k += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(i * 8 + g) * ld + k + q];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bw[(j * 8 + g) * ld + k + q];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(c0[i * 4 + j], c1[i * 4 + j], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every CTA sweeps the whole buffer of n doubles with RED adds (coalesced)
__global__ void k_red(double* buf, size_t n, int sweeps) {
  for (int s = 0; s < sweeps; ++s) {
    // stagger the start so CTAs do not all hit the same lines at once
    size_t start = ((size_t)blockIdx.x * 7919u * 256u) % n;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
      size_t j = start + i; if (j >= n) j -= n;
      atomicAdd(&buf[j], 1.0);
    }
  }
}

// bulk async reduce from shared memory, chunk bytes each op
__global__ void k_bulkred(double* buf, size_t n, int sweeps, int chunk_doubles) {
  extern __shared__ __align__(128) double sm[];
  for (int i = threadIdx.x; i < chunk_doubles; i += blockDim.x) sm[i] = 1.0;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sm);
    size_t nch = n / chunk_doubles;
    for (int s = 0; s < sweeps; ++s) {
      size_t start = ((size_t)blockIdx.x * 7919u) % nch;
      for (size_t c = 0; c < nch; ++c) {
        size_t j = start + c; if (j >= nch) j -= nch;
        double* g = buf + j * chunk_doubles;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                     :: "l"(g), "r"(saddr), "r"(chunk_doubles * 8) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if ((c & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s sms %d smem/blk optin %zu clock %d kHz\n", p.name, nsm, p.sharedMemPerBlockOptin, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int cps : {1, 2}) {
      if (threads * cps > 2048) continue;
      int grid = nsm * cps;
      float ms = timeit([&] { k_dmma<8><<<grid, threads>>>(out, iters); });
      double fl = (double)grid * (threads / 32) * iters * 8 * 512.0;
      printf("dmma884 acc8  thr %4d cta/sm %d : %8.3f ms  %7.2f TFLOP/s\n", threads, cps, ms, fl / ms * 1e-9);
    }
  }
  {
    int grid = nsm, threads = 256;
    float ms = timeit([&] { k_dmma<2><<<grid, threads>>>(out, iters); });
    printf("dmma884 acc2  thr 256 : %8.3f ms %7.2f TFLOP/s\n", ms, (double)grid * 8 * iters * 2 * 512.0 / ms * 1e-9);
    ms = timeit([&] { k_dmma<4><<<grid, threads>>>(out, iters); });
    printf("dmma884 acc4  thr 256 : %8.3f ms %7.2f TFLOP/s\n", ms, (double)grid * 8 * iters * 4 * 512.0 / ms * 1e-9);
    ms = timeit([&] { k_dmma<1><<<grid, 32>>>(out, iters); });
    printf("dmma884 acc1 1 warp (latency): %8.3f ms -> %7.1f ns per dependent DMMA\n", ms, ms * 1e6 / iters);
  }
  for (int threads : {256, 512, 1024}) {
    int grid = nsm * (threads == 1024 ? 2 : 2);
    float ms = timeit([&] { k_dfma<8><<<grid, threads>>>(out, iters); });
    double fl = (double)grid * threads * iters * 8 * 2.0;
    printf("dfma acc8 thr %4d grid %d : %8.3f ms  %7.2f TFLOP/s\n", threads, grid, ms, fl / ms * 1e-9);
  }
  {
    int grid = nsm * 2, threads = 256;
    float ms = timeit([&] { k_mixed<8, 2><<<grid, threads>>>(out, iters); });
    double fm = (double)grid * 8 * iters * 8 * 512.0, ff = (double)grid * threads * iters * 2 * 2.0;
    printf("mixed 8 dmma + 2 dfma: %8.3f ms  dmma %7.2f TF + dfma %7.2f TF\n", ms, fm / ms * 1e-9, ff / ms * 1e-9);
    ms = timeit([&] { k_mixed<8, 8><<<grid, threads>>>(out, iters); });
    ff = (double)grid * threads * iters * 8 * 2.0;
    printf("mixed 8 dmma + 8 dfma: %8.3f ms  dmma %7.2f TF + dfma %7.2f TF\n", ms, fm / ms * 1e-9, ff / ms * 1e-9);
  }
  {
    const int KP = 64; int ld = KP + 4; size_t sh = (32 + 64) * ld * sizeof(double);
    CK(cudaFuncSetAttribute(k_dmma_smem<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    for (int cps : {1, 2}) {
      int grid = nsm * cps, threads = 256, it2 = 2000;
      float ms = timeit([&] { k_dmma_smem<KP><<<grid, threads, sh>>>(out, it2); });
      double fl = (double)grid * 8 * it2 * (KP / 4) * 16 * 512.0;
      printf("dmma smem-fed warp 32x32 cta/sm %d: %8.3f ms %7.2f TFLOP/s\n", cps, ms, fl / ms * 1e-9);
    }
  }
  // RED throughput
  for (size_t n : {(size_t)131072, (size_t)4 << 20}) {
    double* buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
    int sweeps = n > 1000000 ? 1 : 8;
    float ms = timeit([&] { k_red<<<nsm, 256>>>(buf, n, sweeps); }, 2);
    double adds = (double)nsm * n * sweeps;
    printf("red.add.f64 buffer %zu doubles, %d CTAs sweep all: %8.3f ms  %7.2f Gadd/s  (%6.1f GB/s)\n", n, nsm, ms, adds / ms * 1e-6, adds * 8 / ms * 1e-6);
    for (int chunk : {256, 1024, 4096}) {
      CK(cudaFuncSetAttribute(k_bulkred, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk * 8));
      float ms2 = timeit([&] { k_bulkred<<<nsm, 128, chunk * 8>>>(buf, n, sweeps, chunk); }, 2);
      printf("bulk reduce add.f64 chunk %5d B buffer %zu: %8.3f ms  %7.2f Gadd/s  (%6.1f GB/s)\n", chunk * 8, n, ms2, adds / ms2 * 1e-6, adds * 8 / ms2 * 1e-6);
    }
    // verify the sum on a few entries
    double h[4]; CK(cudaMemcpy(h, buf, 32, cudaMemcpyDeviceToHost));
    printf("  check buf[0]=%g\n", h[0]);
    CK(cudaFree(buf));
  }
  printf("done\n");
  return 0;
}
