"""Mirror of `vfegpssm/kernels_multi_output.py` (the SE kernel the driver uses, models.py:7,58).

Same class names, constructor arguments, attributes (`input_dim`, `logvariance`,
`loglengthscales`, `variance`, `lengthscales`, `ARD`) and methods (`K`, `Kdiag`); evaluation
is eager on DLPack tensors through `ffvd_kernel_K` / `ffvd_kernel_Kdiag`.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from ._tensor import as_f64, context_for, empty_like_lib, is_torch, to_lib


class Kernel(object):
    """`kernels_multi_output.py:6-106`: input_dim / active_dims handling and `_slice`."""

    kind = None

    def __init__(self, input_dim, active_dims=None, name=None):
        self.input_dim = int(input_dim)
        if active_dims is None:
            self.active_dims = slice(input_dim)
        elif isinstance(active_dims, slice):
            self.active_dims = active_dims
            if active_dims.start is not None and active_dims.stop is not None and active_dims.step is not None:
                assert len(range(active_dims.start, active_dims.stop, active_dims.step)) == input_dim
        else:
            self.active_dims = np.array(active_dims, dtype=np.int32)
            assert len(active_dims) == input_dim
        self.name = name

    def _validate_ard_shape(self, name, value, ARD=None):
        # kernels_multi_output.py:35-59 -- same error behaviour
        if ARD is None:
            ARD = np.asarray(value).squeeze().shape != ()
        if ARD:
            value = value * np.ones(self.input_dim, dtype=float)
        correct_shape = () if (self.input_dim == 1 or not ARD) else (self.input_dim,)
        if np.asarray(value).squeeze().shape != correct_shape:
            raise ValueError("shape of {} does not match input_dim".format(name))
        return value, ARD

    def _slice(self, X, X2):
        # kernels_multi_output.py:84-106: first input_dim columns (or active_dims), width asserted
        def sl(A):
            if A is None:
                return None
            A = A[..., self.active_dims] if isinstance(self.active_dims, slice) else A[..., list(self.active_dims)]
            if A.shape[-1] != self.input_dim:
                raise ValueError("input has %d columns, kernel expects input_dim=%d" % (A.shape[-1], self.input_dim))
            return as_f64(A)
        return sl(X), sl(X2)

    def compute_K(self, X, Z):
        return self.K(X, Z)

    def compute_K_symm(self, X):
        return self.K(X)

    def compute_Kdiag(self, X):
        return self.Kdiag(X)

    # hyper-parameters as float64 arrays in the library of `ref`
    def _hyper(self, ref):
        logv = to_lib(ref, self.logvariance).reshape(1)
        logl = None
        if self.kind == _capi.KERNEL_SE:
            ll = to_lib(ref, self.loglengthscales).reshape(-1)
            if ll.shape[0] == 1 and self.input_dim > 1:      # isotropic (ARD=False): one lengthscale for all dims
                ll = ll.repeat(self.input_dim) if is_torch(ll) else np.repeat(ll, self.input_dim)
            logl = ll.contiguous() if is_torch(ll) else np.ascontiguousarray(ll)
        return logv, logl

    def K(self, X, X2=None, presliced=False):
        if not presliced:
            X, X2 = self._slice(X, X2)
        else:
            X, X2 = as_f64(X), as_f64(X2)
        logv, logl = self._hyper(X)
        n2 = X.shape[0] if X2 is None else X2.shape[0]
        out = empty_like_lib(X, (X.shape[0], n2))
        return context_for(X).kernel_K(self.kind, X, X2, logv, logl, out)

    def Kdiag(self, X, presliced=False):
        if not presliced:
            X, _ = self._slice(X, None)
        else:
            X = as_f64(X)
        logv, logl = self._hyper(X)
        out = empty_like_lib(X, (X.shape[0],))
        return context_for(X).kernel_Kdiag(self.kind, X, logv, logl, out)


class Stationary(Kernel):
    """`kernels_multi_output.py:130-161`."""

    _opt_kw = "kernel_optimization"

    def __init__(self, input_dim, variance=0.1, lengthscales=1.0, active_dims=None, ARD=None, name=None,
                 kernel_optimization=False):
        super().__init__(input_dim, active_dims, name=name)
        self._v = variance
        self._l = lengthscales
        self.trainable = bool(kernel_optimization)
        self.logvariance = np.asarray(np.log(variance), dtype=np.float64)
        lengthscales, self.ARD = self._validate_ard_shape("lengthscales", lengthscales, ARD)
        self.loglengthscales = np.asarray(np.log(lengthscales), dtype=np.float64)

    @property
    def variance(self):
        return np.exp(self.logvariance) if not is_torch(self.logvariance) else self.logvariance.exp()

    @property
    def lengthscales(self):
        return np.exp(self.loglengthscales) if not is_torch(self.loglengthscales) else self.loglengthscales.exp()

    def __str__(self):
        return "\n".join(["======= Kernel: RBF", " Variance = %.3f" % self._v,
                          " Lengthscales = %s (ARD = %s)" % (np.array2string(np.array(self._l), precision=3), self.ARD)])


class SquaredExponential(Stationary):
    """`kernels_multi_output.py:240-247`: k(x,x') = v exp(-1/2 sum_j ((x_j-x'_j)/l_j)^2)."""

    kind = _capi.KERNEL_SE
