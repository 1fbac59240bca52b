// Elementwise / small standalone operators of the hot path (sm_100a).
#pragma once
#include "ffvd_common.cuh"

namespace ffvd {

// kernels_multi_output.py:202-214,246-247 (SE) / kernels.py:270-276 (Linear): out[i][j] = k(X_i, X2_j)
template <int KIND>
__global__ void kernel_K_kernel(const double* __restrict__ X, const double* __restrict__ X2, int N, int N2, int Din,
                                const double* __restrict__ logv, const double* __restrict__ logl,
                                double* __restrict__ out) {
  const size_t n = (size_t)N * N2;
  const double v = exp(logv[0]);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t i = idx / N2, j = idx % N2;
    double s = 0.0;
    for (int c = 0; c < Din; ++c) {
      const double a = X[i * Din + c], b = X2[j * Din + c];
      if (KIND == 0) {
        const double il = exp(-logl[c]);
        const double t = a * il - b * il;
        s = fma(t, t, s);
      } else {
        s = fma(a, b, s);
      }
    }
    out[idx] = (KIND == 0) ? v * exp(-0.5 * s) : v * s;
  }
}

// kernels_multi_output.py:199-200 / kernels.py:278-281
template <int KIND>
__global__ void kernel_Kdiag_kernel(const double* __restrict__ X, int N, int Din, const double* __restrict__ logv,
                                    double* __restrict__ out) {
  const double v = exp(logv[0]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)N; i += (size_t)gridDim.x * blockDim.x) {
    if (KIND == 0) {
      out[i] = v;
    } else {
      double s = 0.0;
      for (int c = 0; c < Din; ++c) s = fma(X[i * Din + c], X[i * Din + c], s);
      out[i] = v * s;
    }
  }
}

// likelihoods.py:89-93 (vec=0 -> (N,Dy)) and :96-111 (vec=1 -> (N)); no -0.5 log 2 pi
__global__ void logdensity_diag_kernel(const double* __restrict__ y, const double* __restrict__ ymean,
                                       const double* __restrict__ R, int N, int Dy, int vec, double* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)N; i += (size_t)gridDim.x * blockDim.x) {
    double acc = 0.0, lr = 0.0;
    for (int c = 0; c < Dy; ++c) {
      const double r = (y[i * Dy + c] - ymean[i * Dy + c]) / R[c];
      const double e = -0.5 * r * r;
      if (vec) { acc += e; lr += log(R[c]); }
      else out[i * Dy + c] = e - log(R[c]);
    }
    if (vec) out[i] = acc - lr;
  }
}

// likelihoods.py:114-127 logdensity_norm with a FULL lower-triangular noise factor L = Rchols (the particle-Gibbs weights,
// base_model.py:62-66): alpha = L^{-1} (y - ymean)^T by forward substitution (tf.linalg.triangular_solve(lower=True) reads
// only the lower triangle), out[n] = -1/2 |alpha_n|^2 - sum_i log L_ii.  y has N rows or ONE row broadcast against the N
// rows of ymean (the reference subtracts self.Y[tt] from a (P,Dy) matrix).  One thread per row; Dy <= 64.
#define FFVD_MAX_DY 64
__global__ void __launch_bounds__(128) logdensity_full_kernel(const double* __restrict__ y, int y_rows, const double* __restrict__ ymean,
                                                              const double* __restrict__ L, int N, int Dy, double* __restrict__ out) {
  __shared__ double Ls[FFVD_MAX_DY * FFVD_MAX_DY];
  __shared__ double logdet;
  for (int i = threadIdx.x; i < Dy * Dy; i += blockDim.x) Ls[i] = L[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < Dy; ++i) a += log(Ls[i * Dy + i]);
    logdet = a;
  }
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    double alpha[FFVD_MAX_DY];
    const double* yr = y + (size_t)(y_rows == 1 ? 0 : n) * Dy;
    double ss = 0.0;
    for (int i = 0; i < Dy; ++i) {
      double t = yr[i] - ymean[(size_t)n * Dy + i];
      for (int j = 0; j < i; ++j) t = fma(-Ls[i * Dy + j], alpha[j], t);
      t /= Ls[i * Dy + i];
      alpha[i] = t;
      ss = fma(t, t, ss);
    }
    out[n] = -0.5 * ss - logdet;
  }
}

// base_model.py:150-179: adaptive SG-HMC, Jacobi semantics (all right-hand sides read pre-step state).
// 12 streams/element in burn-in (7 reads + 5 writes = 96 B), 7 in sampling (5 R + 2 W = 56 B).
__device__ __forceinline__ void sghmc_one(double& th, double gr, double nz, double& xio, double& go, double& g2o, double& po,
                                          int burn_in, double eps2, double mdecay, double ns) {
  const double Minv = 1.0 / (sqrt(g2o + 1e-16) + 1e-16);                 // :160 (old g2)
  const double sigma = sqrt(fmax(ns * Minv, 1e-16));                      // :169-170
  const double pt = po - eps2 * Minv * gr - mdecay * po + nz * sigma;     // :172
  if (burn_in) {
    const double r = 1.0 / (xio + 1.0);                                   // :156
    const double gn = (1.0 - r) * go + r * gr;                            // :157
    const double g2n = (1.0 - r) * g2o + r * gr * gr;                     // :158
    xio = 1.0 + xio * (1.0 - go * go / (g2o + 1e-16));                    // :159
    go = gn; g2o = g2n;
  }
  po = pt;
  th = th + pt;                                                           // :173
}

// 16-byte accesses, two elements per thread and trip (VEC = 1; all pointers 16-byte aligned), scalar tail / fallback (VEC = 0).
template <int BURN_IN, int VEC>
__global__ void __launch_bounds__(256) sghmc_kernel(double* __restrict__ theta, const double* __restrict__ grad,
                                                    const double* __restrict__ noise, double* __restrict__ xi,
                                                    double* __restrict__ g, double* __restrict__ g2,
                                                    double* __restrict__ p, size_t n, double eps, double mdecay,
                                                    double eps_scaled) {
  const double eps2 = eps * eps;
  const double ns = 2.0 * eps_scaled * eps_scaled * mdecay;
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const size_t n2 = n / 2;
    for (size_t i = i0; i < n2; i += stride) {
      double2 th = reinterpret_cast<double2*>(theta)[i], po = reinterpret_cast<double2*>(p)[i], g2o = reinterpret_cast<double2*>(g2)[i];
      const double2 gr = reinterpret_cast<const double2*>(grad)[i], nz = reinterpret_cast<const double2*>(noise)[i];
      double2 xio = make_double2(0.0, 0.0), go = make_double2(0.0, 0.0);
      if (BURN_IN) { xio = reinterpret_cast<double2*>(xi)[i]; go = reinterpret_cast<double2*>(g)[i]; }
      sghmc_one(th.x, gr.x, nz.x, xio.x, go.x, g2o.x, po.x, BURN_IN, eps2, mdecay, ns);
      sghmc_one(th.y, gr.y, nz.y, xio.y, go.y, g2o.y, po.y, BURN_IN, eps2, mdecay, ns);
      if (BURN_IN) {
        reinterpret_cast<double2*>(g)[i] = go; reinterpret_cast<double2*>(g2)[i] = g2o; reinterpret_cast<double2*>(xi)[i] = xio;
      }
      reinterpret_cast<double2*>(p)[i] = po;
      reinterpret_cast<double2*>(theta)[i] = th;
    }
  }
  for (size_t i = (VEC ? 2 * (n / 2) : 0) + i0; i < n; i += stride) {
    double th = theta[i], po = p[i], g2o = g2[i], xio = BURN_IN ? xi[i] : 0.0, go = BURN_IN ? g[i] : 0.0;
    sghmc_one(th, grad[i], noise[i], xio, go, g2o, po, BURN_IN, eps2, mdecay, ns);
    if (BURN_IN) { g[i] = go; g2[i] = g2o; xi[i] = xio; }
    p[i] = po;
    theta[i] = th;
  }
}

// TF1 AdamOptimizer apply (dgp_model.py:303-305): epsilon-hat form.  HBM-bound: 4 reads + 3 writes = 56 B / element;
// VEC = 1 uses 16-byte accesses (all pointers 16-byte aligned), VEC = 0 is the scalar fallback / tail.
__device__ __forceinline__ void adam_one(double& th, double gr, double& m, double& v, double lr_t, double b1, double b2, double eps) {
  m = b1 * m + (1.0 - b1) * gr;
  v = b2 * v + (1.0 - b2) * gr * gr;
  th = th - lr_t * m / (sqrt(v) + eps);
}

template <int VEC>
__global__ void __launch_bounds__(256) adam_kernel(double* __restrict__ theta, const double* __restrict__ grad,
                                                   double* __restrict__ m, double* __restrict__ v, size_t n,
                                                   double lr_t, double b1, double b2, double eps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    for (size_t i = i0; i < n / 2; i += stride) {
      double2 th = reinterpret_cast<double2*>(theta)[i], mm = reinterpret_cast<double2*>(m)[i], vv = reinterpret_cast<double2*>(v)[i];
      const double2 gr = reinterpret_cast<const double2*>(grad)[i];
      adam_one(th.x, gr.x, mm.x, vv.x, lr_t, b1, b2, eps);
      adam_one(th.y, gr.y, mm.y, vv.y, lr_t, b1, b2, eps);
      reinterpret_cast<double2*>(m)[i] = mm; reinterpret_cast<double2*>(v)[i] = vv; reinterpret_cast<double2*>(theta)[i] = th;
    }
  }
  for (size_t i = (VEC ? 2 * (n / 2) : 0) + i0; i < n; i += stride) {
    double th = theta[i], mm = m[i], vv = v[i];
    adam_one(th, grad[i], mm, vv, lr_t, b1, b2, eps);
    m[i] = mm; v[i] = vv; theta[i] = th;
  }
}

}  // namespace ffvd
