// FP64 scalar-pipe probe: DFMA throughput as a function of warps per SM and independent chains per thread, and the
// dependent-DFMA latency.  One CTA per SM (a large dynamic shared-memory request keeps a second CTA away).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/probe2 tools/probe/probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k_dfma(double* out, int iters) {
  extern __shared__ double dummy[];
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
  if (s == 1.2345) dummy[0] = s;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int NACC>
void run(double* out, int nsm, int threads, int iters) {
  const int smem = 120 * 1024;
  cudaFuncSetAttribute(k_dfma<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float ms = timeit([&] { k_dfma<NACC><<<nsm, threads, smem>>>(out, iters); });
  double fl = (double)nsm * threads * iters * NACC * 2.0;
  printf("dfma  warps/SM %2d  chains/thread %2d : %8.3f ms  %6.2f TFLOP/s\n", threads / 32, NACC, ms, fl / ms * 1e-9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * nsm * 1024);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    run<2>(out, nsm, threads, iters); run<4>(out, nsm, threads, iters); run<8>(out, nsm, threads, iters);
    run<16>(out, nsm, threads, iters); run<32>(out, nsm, threads, iters);
  }
  {
    cudaFuncSetAttribute(k_dfma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    float ms = timeit([&] { k_dfma<1><<<1, 32, 120 * 1024>>>(out, 2000000); });
    printf("dependent DFMA latency: %.2f ns = %.1f clk at %.0f MHz\n", ms * 1e6 / 2000000, ms * 1e6 / 2000000 * p.clockRate * 1e-6, p.clockRate * 1e-3);
  }
  return 0;
}
