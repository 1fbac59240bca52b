"""Mirror of `vfegpssm/likelihoods.py:10-127`: Gaussian linear emission and the log-densities
(which omit the -1/2 log 2 pi constant)."""
from __future__ import annotations

import numpy as np

from . import _capi
from ._tensor import as_f64, context_for, empty_like_lib, is_torch, to_lib


class Gaussian(object):
    """y = x C + d with noise Cholesky exp(log_Rchols), `likelihoods.py:10-83`."""

    def __init__(self, Y_dim, X_output_dim, CC=None, DD=None, RR_chol=None, hyperparameter_sampling=False,
                 likelihood_traning=True):
        self.trainable = bool(likelihood_traning) and not hyperparameter_sampling
        self.CC = np.ones((X_output_dim, Y_dim)) if CC is None else as_f64(CC)              # :12-24
        self.DD = np.zeros((Y_dim,)) if DD is None else as_f64(DD)
        self._fixed_R = None
        if Y_dim != 1:
            # likelihoods.py:56-61: a NON-trainable factor, ones below the diagonal and exp(log 1.5) on it; RR_chol is ignored
            # and there is no log_Rchols attribute (so, as in the reference, a DGPSSM cannot be built on top of it:
            # dgp_model.py:250,333 read likelihood.Rchols[0] / likelihood.log_Rchols)
            self.Rchols_initial = np.ones((Y_dim, Y_dim))
            self.Rchols_diag = np.ones(Y_dim) * np.log(1.5)
            self._Rchols_full = self.Rchols_initial * np.tril(np.ones((Y_dim, Y_dim)), -1) + np.diag(np.exp(self.Rchols_diag))
            return
        if RR_chol is None:
            self.log_Rchols = np.ones((Y_dim, Y_dim)) * np.log(0.1)                              # :50-52
        else:
            R = as_f64(RR_chol)
            self.log_Rchols = R.log() if is_torch(R) else np.log(R)                              # :54
        self._fixed_R = RR_chol if (hyperparameter_sampling and RR_chol is not None) else None

    @property
    def Rchols(self):
        if getattr(self, "_Rchols_full", None) is not None:
            return self._Rchols_full
        L = self.log_Rchols
        return L.exp() if is_torch(L) else np.exp(L)

    def conditional_mean(self, F):
        return F

    def predict_mean(self, X_end):
        """`likelihoods.py:76-79`: X_end C + d (evaluated as a LinearK Gram matrix on the device)."""
        X_end = as_f64(X_end)
        Ct = to_lib(X_end, self.CC).T
        Ct = Ct.contiguous() if is_torch(Ct) else np.ascontiguousarray(Ct)
        out = empty_like_lib(X_end, (X_end.shape[0], Ct.shape[0]))
        context_for(X_end).kernel_K(_capi.KERNEL_LINEAR, X_end, Ct, to_lib(X_end, np.zeros(1)), None, out)
        return out + to_lib(X_end, self.DD)

    def predict_density(self, ymean, Rchols, Y):
        return logdensity_norm(Y, ymean, Rchols)


def _ld(y, ymean, Rchols, vec):
    y = as_f64(y)
    ymean = to_lib(y, as_f64(ymean))
    R = to_lib(y, as_f64(Rchols)).reshape(-1)
    out = empty_like_lib(y, (y.shape[0],) if vec else tuple(y.shape))
    return context_for(y).logdensity_norm_diag(y, ymean, R, vec, out)


def logdensity_norm_diag_nonvec(y, ymean, Rchols):
    """`likelihoods.py:89-93` -> (N,Dy)."""
    return _ld(y, ymean, Rchols, False)


def logdensity_norm_diag(y, ymean, Rchols):
    """`likelihoods.py:96-111` -> (N,)."""
    return _ld(y, ymean, Rchols, True)


def logdensity_norm(y, ymean, Rchols):
    """`likelihoods.py:114-127` with a full lower-triangular factor (forward substitution on the device; the strictly
    upper triangle is ignored, as by `tf.linalg.triangular_solve(lower=True)`).  `y` may be a single row broadcast against
    the rows of `ymean` (the reference's `self.Y[tt] - y_t_mu`, base_model.py:62-66).  Returns (N,)."""
    ymean = as_f64(ymean)
    if ymean.ndim == 1:
        ymean = ymean[None, :]
    y = to_lib(ymean, as_f64(y))
    R = to_lib(ymean, as_f64(Rchols))
    Dy = ymean.shape[1]
    if R.ndim != 2:
        R = R.reshape(Dy, Dy)
    out = empty_like_lib(ymean, (ymean.shape[0],))
    return context_for(ymean).logdensity_norm(y, ymean, R, out)
