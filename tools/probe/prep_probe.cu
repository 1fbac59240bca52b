// Stand-alone timing + correctness harness for kzz_prep_kernel (small-M Cholesky + triangular inverse), no Python.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../ffvd_b200/csrc -o prep_probe prep_probe.cu
//   ./prep_probe [M] [Din] [D]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define FFVD_PREP_DEBUG 1
#include "prep_post.cuh"
using namespace ffvd;


// phase clocks of the fast path, same sequence as kzz_prep_kernel mode 2 minus the fill
__global__ void __launch_bounds__(512) phase_kernel(const double* Ain, int M, int Mp, double* X, double* XT, long long* clk) {
  extern __shared__ __align__(16) double sh[];
  __shared__ int flag;
  double* Asm = sh + 2 * Mp;
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) Asm[(idx / M) * (M + 1) + idx % M] = Ain[idx];
  __syncthreads();
  const long long t0 = clock64();
  chol_fast_smem(Asm, M + 1, M, sh, Mp, Asm + (size_t)M * (M + 1), &flag);
  const long long t1 = clock64();
  double* Xs = Asm + (size_t)M * (M + 1);
  tri_inverse_smem(Asm, M + 1, M, Xs, (M + 1) & ~1, Xs + (size_t)M * ((M + 1) & ~1), X, XT, Mp);
  const long long t2 = clock64();
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 100, Din = argc > 2 ? atoi(argv[2]) : 4, D = argc > 3 ? atoi(argv[3]) : 4;
  const int Mp = (M + 31) / 32 * 32;
  const double jitter = 1e-6;
  std::vector<double> Z((size_t)M * Din), logl((size_t)D * Din), logv(D);
  srand(7);
  for (auto& z : Z) z = 4.0 * rand() / RAND_MAX - 2.0;
  for (auto& l : logl) l = 0.3 * rand() / RAND_MAX - 0.1;
  for (auto& v : logv) v = 0.2 * rand() / RAND_MAX;
  DevProblem P{};
  double *dZ, *dl, *dv;
  CK(cudaMalloc(&dZ, Z.size() * 8)); CK(cudaMalloc(&dl, logl.size() * 8)); CK(cudaMalloc(&dv, D * 8));
  CK(cudaMemcpy(dZ, Z.data(), Z.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dl, logl.data(), logl.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, logv.data(), D * 8, cudaMemcpyHostToDevice));
  P.Z = dZ; P.logl = dl; P.logv = dv; P.U = nullptr; P.logQ = nullptr;
  CK(cudaMalloc(&P.ZT, 32 * Mp * 8)); CK(cudaMalloc(&P.Zf, 32 * Mp * 8)); CK(cudaMalloc(&P.hyp, D * 72 * 8)); CK(cudaMalloc(&P.hq, D * 4 * 8));
  CK(cudaMalloc(&P.UT, (size_t)D * Mp * 8));
  CK(cudaMalloc(&P.Linv, (size_t)D * Mp * Mp * 8)); CK(cudaMalloc(&P.LinvT, (size_t)D * Mp * Mp * 8)); CK(cudaMalloc(&P.status, D * 4));
  CK(cudaMemset(P.Linv, 0, (size_t)D * Mp * Mp * 8)); CK(cudaMemset(P.LinvT, 0, (size_t)D * Mp * Mp * 8));
  P.M = M; P.Mp = Mp; P.Din = Din; P.D = D; P.hs = 1;
  DevProblem* dP;
  CK(cudaMalloc(&dP, sizeof P)); CK(cudaMemcpy(dP, &P, sizeof P, cudaMemcpyHostToDevice));
  int max_smem = 0;
  CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
  // host reference (long double)
  std::vector<long double> Lh((size_t)D * M * M, 0.0L), Xh((size_t)D * M * M, 0.0L);
  for (int d = 0; d < D; ++d) {
    long double* A = &Lh[(size_t)d * M * M];
    for (int m = 0; m < M; ++m)
      for (int n = 0; n <= m; ++n) {
        long double s = 0;
        for (int jd = 0; jd < Din; ++jd) { const long double il = expl(-(long double)logl[d * Din + jd]); const long double t = Z[m * Din + jd] * il - Z[n * Din + jd] * il; s += t * t; }
        A[m * M + n] = expl(logv[d]) * expl(-0.5L * s) + (m == n ? jitter : 0.0);
      }
    for (int j = 0; j < M; ++j) {
      A[j * M + j] = sqrtl(A[j * M + j]);
      for (int i = j + 1; i < M; ++i) A[i * M + j] /= A[j * M + j];
      for (int i = j + 1; i < M; ++i) for (int k = j + 1; k <= i; ++k) A[i * M + k] -= A[i * M + j] * A[k * M + j];
    }
    long double* X = &Xh[(size_t)d * M * M];
    for (int j = 0; j < M; ++j) {
      X[j * M + j] = 1.0L / A[j * M + j];
      for (int i = j + 1; i < M; ++i) { long double s = 0; for (int k = j; k < i; ++k) s += A[i * M + k] * X[k * M + j]; X[i * M + j] = -s / A[i * M + i]; }
    }
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<double> X((size_t)D * Mp * Mp), XT((size_t)D * Mp * Mp);
  for (int mode = 1; mode <= 2; ++mode) {
    size_t smem = (size_t)2 * Mp * 8 + (size_t)M * (M + 1) * 8;
    if (mode == 2) { if (!chol_fast_fits(M, Mp, (size_t)max_smem)) { printf("mode 2 does not fit\n"); continue; } smem = chol_fast_smem_doubles(M, Mp) * 8; }
    CK(cudaFuncSetAttribute(kzz_prep_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hyper_kernel<<<dim3(D, 1), 128>>>(dP, 0, D, 0);
    for (int it = 0; it < 3; ++it) kzz_prep_kernel<0><<<dim3(D, 1), 512, smem>>>(dP, jitter, mode, 0);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 20;
    for (int it = 0; it < reps; ++it) kzz_prep_kernel<0><<<dim3(D, 1), 512, smem>>>(dP, jitter, mode, 0);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(X.data(), P.Linv, X.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(XT.data(), P.LinvT, XT.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0, errT = 0, ref = 0, pad = 0;
    for (int d = 0; d < D; ++d)
      for (int r = 0; r < Mp; ++r)
        for (int c = 0; c < Mp; ++c) {
          const double x = X[((size_t)d * Mp + r) * Mp + c], xt = XT[((size_t)d * Mp + c) * Mp + r];
          if (r < M && c < M) {
            const double h = (double)Xh[((size_t)d * M + r) * M + c];
            err = fmax(err, fabs(x - h)); errT = fmax(errT, fabs(xt - h)); ref = fmax(ref, fabs(h));
          } else pad = fmax(pad, fmax(fabs(x), fabs(xt)));
        }
    printf("M %d Din %d D %d mode %d: %.2f us per launch   max|Linv - ref| %.3e  (transpose %.3e)  / max|ref| %.3e = %.2e   padding %.1e\n", M, Din, D, mode,
           1e3 * ms / reps, err, errT, ref, err / ref, pad);
  }
  if (chol_fast_fits(M, Mp, (size_t)max_smem)) {
    std::vector<double> Ah((size_t)M * M, 0.0);
    for (int m = 0; m < M; ++m) for (int n = 0; n <= m; ++n) {
      long double s = 0;
      for (int jd = 0; jd < Din; ++jd) { const long double il = expl(-(long double)logl[jd]); const long double t = Z[m * Din + jd] * il - Z[n * Din + jd] * il; s += t * t; }
      Ah[m * M + n] = (double)(expl(logv[0]) * expl(-0.5L * s) + (m == n ? jitter : 0.0));
    }
    double* dA; long long* dclk; long long hclk[2];
    CK(cudaMalloc(&dA, Ah.size() * 8)); CK(cudaMalloc(&dclk, 16));
    CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 8, cudaMemcpyHostToDevice));
    const size_t smem = chol_fast_smem_doubles(M, Mp) * 8;
    CK(cudaFuncSetAttribute(phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int it = 0; it < 3; ++it) phase_kernel<<<1, 512, smem>>>(dA, M, Mp, P.Linv, P.LinvT, dclk);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hclk, dclk, 16, cudaMemcpyDeviceToHost));
    printf("phase clocks M %d: chol %lld  inverse+writeout %lld\n", M, hclk[0], hclk[1]);
    long long dbg[256];
    CK(cudaMemcpyFromSymbol(dbg, g_prep_dbg, sizeof dbg));
    printf("  inverse blocks (warp 0):");
    const int nblk = (M + 3) / 4;
    for (int b = 0; b + 1 < nblk; ++b) printf(" %lld", dbg[64 + b + 1] - dbg[64 + b]);
    printf("\n  last block+exit %lld  barrier wait %lld\n", dbg[100] - dbg[64 + nblk - 1], dbg[101] - dbg[100]);
    printf("  chol columns:");
    for (int j = 0; j + 1 < M; ++j) printf(" %lld", dbg[128 + j + 1] - dbg[128 + j]);
    printf("\n");
  }
  return 0;
}
