"""Mirror of `vfegpssm/kernels.py`: the same SE kernel with the constructor keyword spelled
`U_kernel_optimization` (kernels.py:141,156,160) and the live `LinearK` (kernels.py:250-281)."""
from __future__ import annotations

import numpy as np

from . import _capi
from .kernels_multi_output import Kernel, Stationary as _Stationary
from ._tensor import is_torch


class Stationary(_Stationary):
    def __init__(self, input_dim, variance=0.1, lengthscales=1.0, active_dims=None, ARD=None, name=None,
                 U_kernel_optimization=False):
        super().__init__(input_dim, variance, lengthscales, active_dims, ARD, name, kernel_optimization=U_kernel_optimization)


class SquaredExponential(Stationary):
    kind = _capi.KERNEL_SE


class LinearK(Kernel):
    """k(x,x') = v x.x' with scalar v (ARD=False), `kernels.py:250-281`."""

    kind = _capi.KERNEL_LINEAR

    def __init__(self, input_dim, variance=1.0, active_dims=None, ARD=None, name=None):
        super().__init__(input_dim, active_dims, name=name)
        variance, self.ARD = self._validate_ard_shape("variance", variance, ARD)
        if self.ARD:
            raise NotImplementedError("LinearK with per-dimension variances is not on the FFVD hot path (SURVEY Q1)")
        self.logvariance = np.asarray(np.log(variance), dtype=np.float64)
        self.trainable = False

    @property
    def variance(self):
        return np.exp(self.logvariance) if not is_torch(self.logvariance) else self.logvariance.exp()
