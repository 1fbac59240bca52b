// Second latency probe: pointer-chase LDS (32/64 bit), 64-bit shuffle chains with per-lane data, reciprocal / rsqrt chains,
// convergence-barrier cost of a tiny divergent branch.   nvcc -arch=sm_100a lat_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) probe(double* out, long long* clk, int n, int zero) {
  __shared__ int chase[1024];
  __shared__ double dsh[1024];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 1024; i += blockDim.x) { chase[i] = (i + 33) & 1023; dsh[i] = 1.0 + 1e-3 * i; }
  __syncthreads();
  long long t[12];
  int idx = lane;
  t[0] = clock64();
  for (int i = 0; i < n; ++i) idx = chase[idx];                                  // LDS.32 pointer chase
  t[1] = clock64();
  double s = 1.0 + 1e-3 * lane;
  for (int i = 0; i < n; ++i) s = dsh[(__double2loint(s) & 1023)] + 1e-9;       // LDS.64 -> DADD -> address
  t[2] = clock64();
  for (int i = 0; i < n; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1 + (i & 1)); // SHFL64 + DADD
  t[3] = clock64();
  for (int i = 0; i < n; ++i) s = 1.0 / (s + 1.5);                               // DADD + reciprocal
  t[4] = clock64();
  for (int i = 0; i < n; ++i) s = rsqrt(s + 1.5);                                // DADD + rsqrt
  t[5] = clock64();
  for (int i = 0; i < n; ++i) { if ((lane & 3) == (i & 3)) s = s * 1.0000001; s += 1e-12; }   // tiny divergent branch / select
  t[6] = clock64();
  for (int i = 0; i < n; ++i) {                                                  // the inverse's block tail
    double v = s + zero;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    const double x0 = -v * 0.99, x1 = -fma(0.5, x0, v) * 0.98, x2 = -fma(0.4, x1, fma(0.3, x0, v)) * 0.97;
    const double x3 = -fma(0.1, x2, fma(0.2, x1, fma(0.6, x0, v))) * 0.96;
    s = ((lane & 3) == 0 ? x0 : (lane & 3) == 1 ? x1 : (lane & 3) == 2 ? x2 : x3) * 1e-3 + 1.0;
  }
  t[7] = clock64();
  for (int i = 0; i < n; ++i) { s = sqrt(s + 1.5); }                             // DADD + sqrt
  t[8] = clock64();
  out[blockIdx.x * blockDim.x + tid] = s + idx;
  if (tid == 0 && blockIdx.x == 0) for (int i = 0; i < 8; ++i) clk[i] = t[i + 1] - t[i];
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 512 * 8 * 148); cudaMallocManaged(&clk, 128);
  const char* names[8] = {"LDS32 chase", "LDS64+DADD+addr", "SHFL64+DADD", "DADD+rcp", "DADD+rsqrt", "divergent select", "block tail", "DADD+sqrt"};
  for (int nthr : {32, 512}) for (int rep = 0; rep < 2; ++rep) {
    const int n = 1000;
    probe<<<1, nthr>>>(out, clk, n, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    if (rep) { printf("threads %d:", nthr); for (int i = 0; i < 8; ++i) printf("  %s %.1f", names[i], (double)clk[i] / n); printf(" clk\n"); }
  }
  return 0;
}
