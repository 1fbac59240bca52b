// Latency probe for the small-M preparation kernels: dependent DFMA, LDS, generic LD to shared, shuffle, DADD,
// CTA barrier at 512 threads, divergent-trip-count loop followed by a shuffle.   nvcc -arch=sm_100a lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) probe(double* out, long long* clk, const double* gl, int useg, int n) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 4096; i += blockDim.x) sh[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const double* gp = useg ? gl : sh;   // runtime select -> generic loads
  double s = 1.0, a = 1.0000001;
  long long t[10];
  t[0] = clock64();
  for (int i = 0; i < n; ++i) s = fma(s, a, 1e-9);
  t[1] = clock64();
  int idx = lane;
  for (int i = 0; i < n; ++i) idx = (int)sh[idx & 4095] + (idx & 31);          // dependent LDS (+cvt)
  t[2] = clock64();
  for (int i = 0; i < n; ++i) idx = (int)gp[idx & 4095] + (idx & 31);          // dependent generic LD
  t[3] = clock64();
  for (int i = 0; i < n; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1);         // shuffle + DADD
  t[4] = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  t[5] = clock64();
  // divergent trip counts, then shuffle
  for (int i = 0; i < n; ++i) {
    double q = 0.0;
    for (int k = lane & 3; k < (i & 15) + (lane >> 2); k += 4) q = fma(sh[k * 33 + lane], gp[k + i], q);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    if ((lane & 3) == (i & 3)) sh[1024 + lane + (i & 63) * 32] = q;
    s += q;
  }
  t[6] = clock64();
  for (int i = 0; i < n; ++i) { sh[2048 + tid] = s; s = sh[2048 + tid] + 1.0; }   // STS -> LDS same address
  t[7] = clock64();
  out[blockIdx.x * blockDim.x + tid] = s + idx;
  if (tid == 0 && blockIdx.x == 0) for (int i = 0; i < 7; ++i) clk[i] = t[i + 1] - t[i];
}
int main() {
  double *out, *gl; long long* clk;
  cudaMalloc(&out, 512 * 8 * 148); cudaMalloc(&gl, 4096 * 8); cudaMallocManaged(&clk, 64);
  cudaMemset(gl, 0, 4096 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[7] = {"dep DFMA", "dep LDS+cvt", "dep generic LD(smem)+cvt", "shfl+DADD", "syncthreads", "divergent step", "STS->LDS+DADD"};
  for (int nthr : {32, 512}) for (int rep = 0; rep < 2; ++rep) {
    const int n = 1000;
    probe<<<1, nthr, 48 * 1024>>>(out, clk, gl, 0, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    if (rep) { printf("threads %d:", nthr); for (int i = 0; i < 7; ++i) printf("  %s %.1f", names[i], (double)clk[i] / n); printf(" clk\n"); }
  }
  return 0;
}
