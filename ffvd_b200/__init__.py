"""ffvd_b200 -- B200-native (sm_100a) implementation of FFVD's GPSSM log-joint + gradient
evaluation and SG-HMC update, behind the reference's `vfegpssm` operator names.

Sub-modules mirror the reference package: kernels, kernels_multi_output, conditionals,
conditionals_multi_output, likelihoods, dgp_model, base_model.  All arithmetic runs in
hand-written CUDA through the C ABI in include/ffvd_b200.h (ffvd_b200/_capi.py); there is no
CPU fallback.
"""
from ._capi import (Context, FFVDError, NotPositiveDefinite, StaleFactorsError, KERNEL_SE, KERNEL_LINEAR, FLAG_PRIOR_Z_NORMAL,
                    FLAG_PRIOR_ONCE, FLAG_NO_GRADS, FLAG_ASYNC, FLAG_NO_SHARED_PRIORS, FLAG_NO_X0_PRIOR, FLAG_REUSE_KZZ, FLAG_COLLAPSED_P1_ONLY,
                    FLAG_COLLAPSED_RESUME, FLAG_NO_REPLICATED, FLAG_DETERMINISTIC, LIB_PATH, load_library)

__all__ = ["Context", "FFVDError", "NotPositiveDefinite", "StaleFactorsError", "KERNEL_SE", "KERNEL_LINEAR", "FLAG_PRIOR_Z_NORMAL",
           "FLAG_PRIOR_ONCE", "FLAG_NO_GRADS", "FLAG_ASYNC", "FLAG_NO_SHARED_PRIORS", "FLAG_NO_X0_PRIOR", "FLAG_REUSE_KZZ", "FLAG_COLLAPSED_P1_ONLY", "FLAG_COLLAPSED_RESUME",
           "FLAG_NO_REPLICATED", "FLAG_DETERMINISTIC", "LIB_PATH", "load_library"]
