"""CPU float64 oracle for the FFVD GPSSM hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product (``ffvd_b200``)
never does.

It restates, op for op, the TensorFlow graph of xuhuifan/FFVD for

  * ``vfegpssm/kernels_multi_output.py:163-182,199-214,246-247`` (ARD SE kernel),
  * ``vfegpssm/kernels.py:250-281`` (LinearK),
  * ``vfegpssm/conditionals_multi_output.py:6-70,73-120,124-169,206-257`` and
    ``vfegpssm/conditionals.py:6-107`` (conditionals, collapsed bound),
  * ``vfegpssm/likelihoods.py:76-79,89-127`` (Gaussian emission + log-densities),
  * ``vfegpssm/dgp_model.py:105-143,248-297,326-359`` (priors, nll assembly),
  * ``vfegpssm/base_model.py:143-179`` (adaptive SG-HMC update) and
    ``dgp_model.py:303-305`` (TF1 Adam),

using torch CPU float64 tensors (LAPACK/BLAS, like TF's CPU kernels) with
reverse-mode autograd standing in for ``tf.gradients``.

Pinning status: TensorFlow is not installed in the build container, the
reference ships no tests or golden vectors, so this oracle cannot be compared
with a real TF run.  It IS pinned against the reference's own Python source
executed on a torch-backed ``tensorflow`` API shim
(``tests/golden/make_reference_golden.py`` -> ``tests/golden/reference_shim_golden.npz``);
see DESIGN.md "Parity pinning".
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

DT = torch.float64


def _t(x, requires_grad: bool = False) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x.detach().to(DT).clone()
    else:
        t = torch.tensor(np.asarray(x, dtype=np.float64), dtype=DT)
    t.requires_grad_(requires_grad)
    return t


# --------------------------------------------------------------------------
# kernels  (kernels_multi_output.py / kernels.py)
# --------------------------------------------------------------------------
class SquaredExponential:
    """ARD squared-exponential kernel, `kernels_multi_output.py:130-247`.

    Holds ``logvariance`` () and ``loglengthscales`` (input_dim,) like the
    reference's tf.Variables (`:156,160`); ``variance``/``lengthscales`` are
    their exponentials (`:157,161`).
    """

    kind = 0

    def __init__(self, input_dim, variance=0.1, lengthscales=1.0, active_dims=None, ARD=None,
                 name=None, kernel_optimization=False):
        self.input_dim = int(input_dim)
        lengthscales = np.asarray(lengthscales, dtype=np.float64) * np.ones(self.input_dim)
        self.logvariance = _t(np.log(variance))
        self.loglengthscales = _t(np.log(lengthscales))

    @property
    def variance(self):
        return torch.exp(self.logvariance)

    @property
    def lengthscales(self):
        return torch.exp(self.loglengthscales)

    def _scaled_square_dist(self, X, X2):
        # kernels_multi_output.py:163-182 -- the |x|^2 + |z|^2 - 2 x.z expansion
        X = X / self.lengthscales
        Xs = torch.sum(torch.square(X), dim=-1, keepdim=True)
        if X2 is None:
            dist = -2 * (X @ X.T)
            dist = dist + (Xs + Xs.T)
            return dist
        X2 = X2 / self.lengthscales
        X2s = torch.sum(torch.square(X2), dim=-1, keepdim=True)
        dist = -2 * (X @ X2.T)
        dist = dist + (Xs + X2s.T)
        return dist

    def K(self, X, X2=None, presliced=False):
        # kernels_multi_output.py:202-214 + K_r2 :246-247 ; _slice :84-106
        if not presliced:
            X = X[..., : self.input_dim]
            if X2 is not None:
                X2 = X2[..., : self.input_dim]
        return self.variance * torch.exp(-self._scaled_square_dist(X, X2) / 2.0)

    def Kdiag(self, X, presliced=False):
        # kernels_multi_output.py:199-200
        return torch.ones(X.shape[:-1], dtype=DT) * self.variance


class LinearK:
    """Linear kernel, `kernels.py:250-281` (scalar variance, ARD=False)."""

    kind = 1

    def __init__(self, input_dim, variance=1.0, active_dims=None, ARD=None, name=None):
        self.input_dim = int(input_dim)
        self.logvariance = _t(np.log(variance))

    @property
    def variance(self):
        return torch.exp(self.logvariance)

    def K(self, X, X2=None, presliced=False):
        if not presliced:
            X = X[..., : self.input_dim]
            if X2 is not None:
                X2 = X2[..., : self.input_dim]
        if X2 is None:
            return (X * self.variance) @ X.T
        return (X * self.variance) @ X2.T

    def Kdiag(self, X, presliced=False):
        if not presliced:
            X = X[..., : self.input_dim]
        return torch.sum(torch.square(X) * self.variance, 1)


# --------------------------------------------------------------------------
# conditionals
# --------------------------------------------------------------------------
def _tri_solve(L, B, lower=True):
    return torch.linalg.solve_triangular(L, B, upper=not lower)


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False, return_Lm=False):
    """`conditionals_multi_output.py:6-70` == `conditionals.py:6-66`."""
    num_func = f.shape[1]
    Lm = torch.linalg.cholesky(Kmm)
    A = _tri_solve(Lm, Kmn, lower=True)
    if full_cov:
        fvar = Knn - A.T @ A
        fvar = fvar[None, :, :].repeat(num_func, 1, 1)
    else:
        fvar = Knn - torch.sum(torch.square(A), 0)
        fvar = fvar[None, :].repeat(num_func, 1)
    if not white:
        A = _tri_solve(Lm.T, A, lower=False)
    fmean = A.T @ f
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A * q_sqrt.T.unsqueeze(2)
        elif q_sqrt.dim() == 3:
            A_tiled = A.unsqueeze(0).repeat(num_func, 1, 1)
            LTA = q_sqrt.transpose(-1, -2) @ A_tiled
        else:
            raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.dim()))
        if full_cov:
            fvar = fvar + LTA.transpose(-1, -2) @ LTA
        else:
            fvar = fvar + torch.sum(torch.square(LTA), 1)
    if not full_cov:
        fvar = fvar.T
    if return_Lm:
        return fmean, fvar, Lm
    return fmean, fvar


def conditional(Xnew, X, kern, f, *, full_cov=False, q_sqrt=None, white=False, return_Lm=False):
    """Single-kernel conditional, `conditionals.py:69-107` (jitter 1e-7)."""
    num_data = X.shape[0]
    Kmm = kern.K(X) + torch.eye(num_data, dtype=DT) * 1e-7
    Kmn = kern.K(X, Xnew)
    Knn = kern.K(Xnew) if full_cov else kern.Kdiag(Xnew)
    return base_conditional(Kmn, Kmm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white, return_Lm=return_Lm)


def conditional_multi_output(Xnew, X, kern: Sequence, f, *, full_cov=False, q_sqrt=None, white=False):
    """List-of-kernels conditional, `conditionals_multi_output.py:73-120` (jitter 1e-5)."""
    num_data = X.shape[0]
    f_mu, f_var = [], []
    for kk in range(len(kern)):
        Kmm = kern[kk].K(X) + torch.eye(num_data, dtype=DT) * 1e-5
        Kmn = kern[kk].K(X, Xnew)
        Knn = kern[kk].K(Xnew) if full_cov else kern[kk].Kdiag(Xnew)
        mu, var = base_conditional(Kmn, Kmm, Knn, f[:, kk][:, None], full_cov=full_cov, q_sqrt=q_sqrt, white=white)
        f_mu.append(mu)
        f_var.append(var)
    return torch.stack(f_mu)[:, :, 0].T, torch.stack(f_var)[:, :, 0].T


def kernel_pre_cal(X, kern: Sequence):
    """`conditionals_multi_output.py:124-169`: per kernel, L^{-T} of chol(K(X)+1e-5 I)."""
    num_data = X.shape[0]
    out = []
    for kk in range(len(kern)):
        Kmm = kern[kk].K(X) + torch.eye(num_data, dtype=DT) * 1e-5
        Lm = torch.linalg.cholesky(Kmm)
        out.append(_tri_solve(Lm.T, torch.eye(num_data, dtype=DT), lower=False))
    return out


def collapse_u_mean_after_kernel_precalculation(Lm_inverse_seq, X_combine, X, Z, kern, Q):
    """`conditionals_multi_output.py:206-227`: optimal q(u) mean and H^{-1/2} factors."""
    U_mean, Linv_dd = [], []
    M = Z.shape[0]
    for dd in range(len(kern)):
        Knm = kern[dd].K(X_combine, Z)
        tilde_F = Knm @ Lm_inverse_seq[dd]
        H = tilde_F.T @ tilde_F / Q[dd] + torch.eye(M, dtype=DT)
        X_transpose = (X[1:, dd] - X[:-1, dd])[None, :]
        b = X_transpose @ tilde_F / Q[dd]
        U_mean.append(torch.linalg.solve(H, b.T))
        Lm_dd = torch.linalg.cholesky(H)
        Linv_dd.append(_tri_solve(Lm_dd.T, torch.eye(M, dtype=DT), lower=False))
    U_mean = torch.stack(U_mean)           # D x M x 1
    return U_mean[:, :, 0].T, torch.stack(Linv_dd)


def base_conditional_after_kernel_precalculation(Kmn, Lm_inverse_seq_kk, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """`conditionals_multi_output.py:324-387`, op for op (including the doubled L^{-1} of its non-white branch and TF's
    batch broadcasting of a (R',M,M) q_sqrt against the (num_func,M,N) tiled A)."""
    num_func = f.shape[1]
    A = Lm_inverse_seq_kk.T @ Kmn
    if full_cov:
        fvar = Knn - A.T @ A
        fvar = fvar[None, :, :].repeat(num_func, 1, 1)
    else:
        fvar = Knn - torch.sum(torch.square(A), 0)
        fvar = fvar[None, :].repeat(num_func, 1)
    if not white:
        A = Lm_inverse_seq_kk.T @ A
    fmean = A.T @ f
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A * q_sqrt.T.unsqueeze(2)
        elif q_sqrt.dim() == 3:
            A_tiled = A.unsqueeze(0).repeat(num_func, 1, 1)
            LTA = q_sqrt.transpose(-1, -2) @ A_tiled                 # broadcasts (R',M,M) x (1,M,N) -> (R',M,N)
        else:
            raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.dim()))
        if full_cov:
            fvar = fvar + LTA.transpose(-1, -2) @ LTA
        else:
            fvar = fvar + torch.sum(torch.square(LTA), 1)
    if not full_cov:
        fvar = fvar.T
    return fmean, fvar


def conditional_after_kernel_precalculation(Lm_inverse_seq, Xnew, Z, kern, f, *, full_cov=False, q_sqrt=None, white=False):
    """`conditionals_multi_output.py:306-322`; the final `[:, :, 0]` keeps output 0's q_sqrt term for every kernel
    (SURVEY Q9)."""
    f_mu, f_var = [], []
    for kk in range(len(kern)):
        Kmn = kern[kk].K(Z, Xnew)
        Knn = kern[kk].K(Xnew) if full_cov else kern[kk].Kdiag(Xnew)
        mu, var = base_conditional_after_kernel_precalculation(Kmn, Lm_inverse_seq[kk], Knn, f[:, kk][:, None], full_cov=full_cov,
                                                               q_sqrt=q_sqrt, white=white)
        f_mu.append(mu)
        f_var.append(var)
    return torch.stack(f_mu)[:, :, 0].T, torch.stack(f_var)[:, :, 0].T


def rollout(Lm_inverse_seq, x_last, ctrl_future, Z, kern, U_val, q_sqrt, Q, noise):
    """The prediction roll-out of `base_model.py:283-310` for one posterior sample: x_{t+1} = x_t + f(x_t, c_t) +
    eps_t sqrt(var_t + Q), eps_t = noise[t] (the reference draws tf.random.normal; injected here).  Returns the stacked
    states (L,D) and their predictive variances var_t + Q (L,D)."""
    x_t = x_last[None, :]
    xs, vs = [], []
    for t in range(noise.shape[0]):
        xc = torch.cat((x_t, ctrl_future[t][None, :]), dim=1) if ctrl_future.shape[1] > 0 else x_t
        mu, var = conditional_after_kernel_precalculation(Lm_inverse_seq, xc, Z, kern, U_val, white=True, full_cov=False, q_sqrt=q_sqrt)
        mu = mu + x_t
        x_next = mu + noise[t][None, :] * torch.sqrt(var + Q)
        xs.append(x_next[0])
        vs.append((var + Q)[0])
        x_t = x_next
    return torch.stack(xs), torch.stack(vs)


def categorical_from_uniform(logits, u):
    """Categorical(logits).sample with the randomness injected: inverse CDF of softmax(logits) at u in [0,1)."""
    p = torch.softmax(logits, dim=0)
    cdf = torch.cumsum(p, dim=0)
    return torch.clamp(torch.searchsorted(cdf, u.reshape(-1)), max=logits.shape[0] - 1)


def pg_for_x(X, Y, ctrl, Z, kern, U, Q, C, d, Rchols, PG_particles, normals, eps, uniforms):
    """The conditional-SMC (particle Gibbs) sweep of `base_model.py:29-75` (PG_for_X), literally: particles are WHOLE
    trajectories, gathered on resampling; the last particle is the current trajectory X.  Randomness injected:
    normals (P-1,D) initial states, eps (T,P-1,D) transition noise, uniforms (T,P-1) resampling draws (row T-1, entry 0
    picks the returned trajectory).  Returns the new trajectory (T+1,D)."""
    P1 = PG_particles - 1
    T = X.shape[0] - 1
    Linv = kernel_pre_cal(Z, kern)
    particles = [normals[i][None, :] for i in range(P1)]
    particles.append(X[0][None, :])
    for tt in range(T):
        x_t = torch.stack([p[-1] for p in particles[:-1]])
        xc = torch.cat((x_t, ctrl[tt] * torch.ones((P1, 1), dtype=DT)), dim=1) if ctrl.shape[1] > 0 else x_t
        mu, var = conditional_after_kernel_precalculation(Linv, xc, Z, kern, U, white=True, full_cov=False)
        mu = mu + x_t
        x_next = mu + eps[tt] * torch.sqrt(var + Q)
        for i in range(P1):
            particles[i] = torch.cat((particles[i], x_next[i][None, :]), dim=0)
        w = list(torch.unbind(logdensity_norm(Y[tt], x_next @ C + d, Rchols)))
        w.append(logdensity_norm(Y[tt], X[tt + 1][None, :] @ C + d, Rchols)[0])
        particles[-1] = X[:tt + 2]
        logits = torch.stack(w)
        if tt < T - 1:
            idx = categorical_from_uniform(logits, uniforms[tt])
            particles = [particles[int(i)] for i in idx]
            particles.append(X[:tt + 2])
        else:
            idx = categorical_from_uniform(logits, uniforms[tt][:1])
            return particles[int(idx[0])]


def collapse_after_kernel_precalculation(Lm_inverse_seq, X_combine, X, Z, kern, Q, batch_size, Y_N):
    """`conditionals_multi_output.py:230-257`: the three collapsed-bound terms."""
    term1 = 0.0
    term2 = 0.0
    trace_Q_inverse_B = 0.0
    M = Z.shape[0]
    for dd in range(len(kern)):
        Knm = kern[dd].K(X_combine, Z)
        tilde_F = Knm @ Lm_inverse_seq[dd]
        Knn_diag = kern[dd].Kdiag(X_combine)
        H = tilde_F.T @ tilde_F / (batch_size * Q[dd]) * Y_N + torch.eye(M, dtype=DT)
        X_transpose = (X[1:, dd] - X[:-1, dd])[None, :]
        b = X_transpose @ tilde_F / (batch_size * Q[dd]) * Y_N
        term1 = term1 + (-0.5 * torch.linalg.slogdet(H)[1])
        term2 = term2 + 0.5 * (b @ torch.linalg.solve(H, b.T))[0, 0]
        trace_Q_inverse_B = trace_Q_inverse_B + (-0.5 * torch.sum((Knn_diag - torch.sum(tilde_F ** 2, dim=1)) / Q[dd]))
    return -term1 / Y_N, -term2 / Y_N, -trace_Q_inverse_B / Y_N


# --------------------------------------------------------------------------
# likelihoods.py
# --------------------------------------------------------------------------
def logdensity_norm_diag_nonvec(y, ymean, Rchols):
    """`likelihoods.py:89-93` (no -0.5 log 2 pi)."""
    return -0.5 * (((y - ymean) / Rchols[None, :]) ** 2) + (-torch.log(Rchols)[None, :])


def logdensity_norm_diag(y, ymean, Rchols):
    """`likelihoods.py:96-111`."""
    return -0.5 * torch.sum(((y - ymean) / Rchols[None, :]) ** 2, dim=1) + (-torch.sum(torch.log(Rchols)))


def logdensity_norm(y, ymean, Rchols):
    """`likelihoods.py:114-127` (full Cholesky factor)."""
    alphav = _tri_solve(Rchols, (y - ymean).T, lower=True)
    return -0.5 * torch.sum(torch.square(alphav), dim=0) + (-torch.sum(torch.log(torch.diagonal(Rchols))))


# --------------------------------------------------------------------------
# model state + nll assembly (dgp_model.py)
# --------------------------------------------------------------------------
PARAM_NAMES = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")


@dataclass
class Problem:
    """One GPSSM problem in the layout the C-ABI uses (all float64, C-order).

    X (T+1,D) [or (S,T+1,D)], Z (M,Din), U (M,D), logv (D,), logl (D,Din) [SE only],
    logQ (D,), C (D,Dy), d (Dy,), logR (Dy,Dy) [only row 0 is used, dgp_model.py:250],
    Y (T,Dy), ctrl (T,nc).
    """
    X: np.ndarray
    Z: np.ndarray
    U: np.ndarray
    logv: np.ndarray
    logl: Optional[np.ndarray]
    logQ: np.ndarray
    C: np.ndarray
    d: np.ndarray
    logR: np.ndarray
    Y: np.ndarray
    ctrl: np.ndarray
    kind: int = 0                      # 0 SE, 1 LinearK
    prior_type: str = "normal"         # dgp_model.py:105-121 ("normal" is the CLI default)
    name: str = ""

    def params(self) -> Dict[str, np.ndarray]:
        out = {k: getattr(self, k) for k in PARAM_NAMES if getattr(self, k) is not None}
        return out


def _make_kernels(logv, logl, kind, Din):
    kerns = []
    D = logv.shape[0]
    for k in range(D):
        if kind == 0:
            kk = SquaredExponential(Din, 1.0, np.ones(Din))
            kk.logvariance = logv[k]
            kk.loglengthscales = logl[k]
        else:
            kk = LinearK(Din, 1.0)
            kk.logvariance = logv[k]
        kerns.append(kk)
    return kerns


def _priors(p, kind, prior_type, include_U):
    # dgp_model.py:105-143 (prior_Z / prior_hyper / prior_U) and :326-334 (hyper prior)
    if prior_type == "normal":
        prior_Z = -torch.sum(torch.square(p["Z"])) / 2.0
    elif prior_type == "uniform":
        prior_Z = torch.zeros((), dtype=DT)
    else:
        raise ValueError("oracle supports prior_type normal|uniform")
    log005 = math.log(0.05)
    if kind == 0:
        prior_hyper = torch.zeros((), dtype=DT)
        for k in range(p["logv"].shape[0]):
            prior_hyper = prior_hyper + (-torch.sum(torch.square(p["logl"][k])) / 2.0
                                         - torch.sum(torch.square(p["logv"][k] - log005)) / 2.0)
    else:
        # dgp_model.py:129-130 is written for a single LinearK object; with a list of D
        # kernels we sum the same expression over the list.
        prior_hyper = -torch.sum(torch.square(p["logv"] - log005)) / 2.0
    prior_U = -0.5 * torch.sum(torch.square(p["U"])) if include_U else torch.zeros((), dtype=DT)
    hyp = (-torch.sum(torch.square(p["logQ"])) / 2.0 - torch.sum(torch.square(p["C"])) / 2.0
           - torch.sum(torch.square(p["d"])) / 2.0 - torch.sum(torch.square(p["logR"])) / 2.0)
    return prior_Z, prior_hyper, prior_U, hyp


def nll_terms(p: Dict[str, torch.Tensor], Y, ctrl, *, collapsed: bool, kind: int = 0,
              prior_type: str = "normal", shared_priors: bool = True, x0_prior: bool = True):
    """Scalar nll and its named terms for ONE trajectory, `dgp_model.py:248-297`.

    Full batch: batch_placeholder = [0, X_N] (base_model.py:194) so batch_size == Y_N == T.
    """
    X = p["X"]
    T = X.shape[0] - 1
    Tf = float(T)
    D = X.shape[1]
    Din = D + (ctrl.shape[1] if ctrl is not None and ctrl.numel() > 0 else 0)
    kerns = _make_kernels(p["logv"], p.get("logl"), kind, Din)
    Q = torch.exp(p["logQ"])
    Rchols = torch.exp(p["logR"])
    y_mean = X[1:T + 1] @ p["C"] + p["d"]                                  # likelihoods.py:76-79
    log_lik = logdensity_norm_diag(Y[0:T], y_mean, Rchols[0])              # dgp_model.py:250
    prior_x_0 = -torch.sum(torch.square(X[0])) / 2.0                        # :252
    prior_Z, prior_hyper, prior_U, hyp = _priors(p, kind, prior_type, include_U=not collapsed)
    if not shared_priors:          # time-sharded evaluation: blocks after the first (FFVD_FLAG_NO_SHARED_PRIORS)
        prior_Z = prior_hyper = prior_U = hyp = torch.zeros((), dtype=DT)
    if not x0_prior:               # blocks that do not start at t = 0 (FFVD_FLAG_NO_X0_PRIOR)
        prior_x_0 = torch.zeros((), dtype=DT)
    nll_log_likelihood = -torch.sum(log_lik) / Tf                           # :264
    Xc = torch.cat((X[0:T], ctrl[0:T]), dim=1) if Din > D else X[0:T]       # :269 / :340
    out = {}
    if collapsed:
        Linv = kernel_pre_cal(p["Z"], kerns)                                # :273
        t1, t2, tr = collapse_after_kernel_precalculation(Linv, Xc, X[0:T + 1], p["Z"], kerns, Q, Tf, Tf)
        x_t_prior_Q = -torch.sum(logdensity_norm_diag_nonvec(X[1:T + 1], X[0:T], Q ** 0.5)) / Tf   # :283
        nll_part_prior = -(prior_hyper + prior_Z + prior_x_0 + hyp) / Tf    # :286
        nll = nll_part_prior + nll_log_likelihood + x_t_prior_Q + tr + t1 + t2   # :288
        out.update(term1=t1, term2=t2)
    else:
        mean_reg, var_reg = conditional_multi_output(Xc, p["Z"], kerns, p["U"], white=True)   # :344
        mean_reg = mean_reg + X[:-1]                                        # :347
        reg_trace = -0.5 * torch.sum((Q[None, :] ** (-1)) * var_reg, dim=1) # :349
        reg_x_prior = logdensity_norm_diag(X[1:], mean_reg, Q ** 0.5)       # :352
        tr = -torch.sum(reg_trace) / Tf                                     # :292
        x_t_prior_Q = -torch.sum(reg_x_prior) / Tf                          # :294
        nll_part_prior = -(prior_U + prior_hyper + prior_Z + prior_x_0 + hyp) / Tf   # :296
        nll = nll_part_prior + nll_log_likelihood + x_t_prior_Q + tr        # :297
        z = torch.zeros((), dtype=DT)
        out.update(term1=z, term2=z)
    out.update(nll=nll, prior=nll_part_prior, loglik=nll_log_likelihood, xq=x_t_prior_Q, trace=tr)
    return out


TERM_NAMES = ("prior", "loglik", "xq", "trace", "term1", "term2")


def nll_and_grads(prob: Problem, *, collapsed: bool, shared_priors: bool = True, x0_prior: bool = True) -> Dict[str, np.ndarray]:
    """nll, the six terms and d nll / d{X,Z,U,logv,logl,logQ,C,d,logR} by autograd
    (the stand-in for `tf.gradients`, base_model.py:148 / `adam.minimize`, dgp_model.py:305).

    For X of shape (S,T+1,D): per-sample nll/terms of length S, per-sample X-grad, and
    shared-parameter grads summed over samples (SURVEY section 8 batch-axis contract).
    """
    X = np.asarray(prob.X, dtype=np.float64)
    single = X.ndim == 2
    Xs = X[None] if single else X
    S = Xs.shape[0]
    Y = _t(prob.Y)
    ctrl = _t(prob.ctrl)
    acc: Dict[str, np.ndarray] = {}
    nlls = np.zeros(S)
    terms = np.zeros((S, len(TERM_NAMES)))
    gX = np.zeros_like(Xs)
    for s in range(S):
        p = {"X": _t(Xs[s], True)}
        for k in PARAM_NAMES[1:]:
            v = getattr(prob, k)
            if v is not None:
                p[k] = _t(v, True)
        out = nll_terms(p, Y, ctrl, collapsed=collapsed, kind=prob.kind, prior_type=prob.prior_type, shared_priors=shared_priors,
                        x0_prior=x0_prior)
        names = list(p.keys())
        grads = torch.autograd.grad(out["nll"], [p[k] for k in names], allow_unused=True)
        nlls[s] = out["nll"].item()
        terms[s] = [out[k].item() for k in TERM_NAMES]
        for k, g in zip(names, grads):
            gnp = np.zeros(tuple(p[k].shape)) if g is None else g.numpy().copy()
            if k == "X":
                gX[s] = gnp
            else:
                acc[k] = acc.get(k, 0.0) + gnp
    res = {"nll": nlls, "terms": terms, "g_X": gX}
    for k, v in acc.items():
        res["g_" + k] = v
    if single:
        res["nll"] = res["nll"][0]
        res["terms"] = res["terms"][0]
        res["g_X"] = res["g_X"][0]
    return res


# --------------------------------------------------------------------------
# SG-HMC + Adam updates (base_model.py:143-179, dgp_model.py:303-305)
# --------------------------------------------------------------------------
def sghmc_update(theta, grad, noise, xi, g, g2, p, *, epsilon=0.01, mdecay=0.05, X_N, burn_in: bool):
    """One adaptive SG-HMC update with Jacobi semantics (every right-hand side reads the
    pre-step state).  ``noise`` is the standard-normal draw that the reference
    takes from `tf.random.normal` (`base_model.py:171`).  Returns new (theta, xi, g, g2, p).
    """
    theta, grad, noise, xi, g, g2, p = (np.asarray(a, dtype=np.float64) for a in (theta, grad, noise, xi, g, g2, p))
    r_t = 1.0 / (xi + 1.0)
    g_t = (1.0 - r_t) * g + r_t * grad
    g2_t = (1.0 - r_t) * g2 + r_t * grad ** 2
    xi_t = 1.0 + xi * (1.0 - g * g / (g2 + 1e-16))
    Minv = 1.0 / (np.sqrt(g2 + 1e-16) + 1e-16)
    epsilon_scaled = epsilon / np.sqrt(float(X_N))
    noise_scale = 2.0 * epsilon_scaled ** 2 * mdecay * Minv
    sigma = np.sqrt(np.maximum(noise_scale, 1e-16))
    sample_t = noise * sigma
    p_t = p - epsilon ** 2 * Minv * grad - mdecay * p + sample_t
    theta_t = theta + p_t
    if burn_in:
        return theta_t, xi_t, g_t, g2_t, p_t
    return theta_t, xi.copy(), g.copy(), g2.copy(), p_t


def adam_update(theta, grad, m, v, *, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF1 `AdamOptimizer` apply (epsilon-hat form): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps).  ``step`` is the 1-based step count."""
    theta, grad, m, v = (np.asarray(a, dtype=np.float64) for a in (theta, grad, m, v))
    lr_t = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    m_t = beta1 * m + (1.0 - beta1) * grad
    v_t = beta2 * v + (1.0 - beta2) * grad * grad
    theta_t = theta - lr_t * m_t / (np.sqrt(v_t) + eps)
    return theta_t, m_t, v_t


def adam_learning_rate(global_step: int = 1) -> float:
    """`base_model.py:190`."""
    return 0.003 * (0.95 ** (global_step / 1000))
