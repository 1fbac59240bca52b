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

// base_model.py:150-179: adaptive SG-HMC, Jacobi semantics (all right-hand sides read pre-step state).
// 12 streams/element in burn-in (7 reads + 5 writes = 96 B), 7 in sampling (5 R + 2 W = 56 B).
template <int BURN_IN>
__global__ void __launch_bounds__(256) sghmc_kernel(double* __restrict__ theta, const double* __restrict__ grad,
                                                    const double* __restrict__ noise, double* __restrict__ xi,
                                                    double* __restrict__ g, double* __restrict__ g2,
                                                    double* __restrict__ p, size_t n, double eps, double mdecay,
                                                    double eps_scaled) {
  const double eps2 = eps * eps;
  const double ns = 2.0 * eps_scaled * eps_scaled * mdecay;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double gr = grad[i], g2o = g2[i], po = p[i], th = theta[i], nz = noise[i];
    const double Minv = 1.0 / (sqrt(g2o + 1e-16) + 1e-16);                 // :160 (old g2)
    const double sigma = sqrt(fmax(ns * Minv, 1e-16));                      // :169-170
    const double pt = po - eps2 * Minv * gr - mdecay * po + nz * sigma;     // :172
    if (BURN_IN) {
      const double xio = xi[i], go = g[i];
      const double r = 1.0 / (xio + 1.0);                                   // :156
      g[i] = (1.0 - r) * go + r * gr;                                       // :157
      g2[i] = (1.0 - r) * g2o + r * gr * gr;                                // :158
      xi[i] = 1.0 + xio * (1.0 - go * go / (g2o + 1e-16));                  // :159
    }
    p[i] = pt;
    theta[i] = th + pt;                                                     // :173
  }
}

// TF1 AdamOptimizer apply (dgp_model.py:303-305): epsilon-hat form
__global__ void __launch_bounds__(256) adam_kernel(double* __restrict__ theta, const double* __restrict__ grad,
                                                   double* __restrict__ m, double* __restrict__ v, size_t n,
                                                   double lr_t, double b1, double b2, double eps) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double gr = grad[i];
    const double mt = b1 * m[i] + (1.0 - b1) * gr;
    const double vt = b2 * v[i] + (1.0 - b2) * gr * gr;
    m[i] = mt; v[i] = vt;
    theta[i] = theta[i] - lr_t * mt / (sqrt(vt) + eps);
  }
}

}  // namespace ffvd
