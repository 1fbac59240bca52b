"""Mirror of `vfegpssm/dgp_model.py`: `Layer` (variables X/U/Z and priors, :45-143) and `DGPSSM`
(SG-HMC variable selection :213-244, nll assembly :248-297, regularizer :337-359), eager on the
device.  One call to `DGPSSM.evaluate()` is one evaluation of the reference's `nll` tensor plus
`tf.gradients` w.r.t. every parameter, executed by the fused CUDA path."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import _capi
from .base_model import BaseModel
from . import conditionals_multi_output
from ._tensor import context_for

TERM_NAMES = ("nll_part_prior", "nll_log_likelihood", "x_t_prior_Q", "nll_reg_trace_inverse_Q_B", "later_term1", "later_term2")


def sghmc_variable_names(kind, case_val, kernel_optimization, kernel_train_flag, U_optimization, U_collapse,
                         Z_optimization, hyperparameter_sampling=False):
    """Which parameters SG-HMC samples, `dgp_model.py:213-244`.  Pure host logic."""
    variables = []
    if case_val == 7:
        variables += ["U", "X"]
    else:
        if not kernel_optimization and kernel_train_flag:
            variables += ["logv", "logl"] if kind == _capi.KERNEL_SE else ["logv"]
        if not U_optimization and not U_collapse:
            variables += ["U"]
        if not Z_optimization:
            variables += ["Z"]
    if hyperparameter_sampling:
        variables += ["logQ", "C", "d", "logR"]
    return variables


class Layer(object):
    """`dgp_model.py:45-143`."""

    def __init__(self, ZZ, U_ini, X_0_ini, X_train_ini, kern, outputs, n_inducing, fixed_mean, x_dims_l, Y_len, full_cov,
                 prior_type="uniform", kernel_type='SquaredExponential', U_optimization=False, U_collapse=False,
                 Z_optimization=False, X_PG=False, case_val=1, device=0):
        import torch
        if not isinstance(kern, (list, tuple)):
            raise TypeError("'%s' object is not subscriptable" % type(kern).__name__)      # dgp_model.py:49
        self.inputs, self.outputs, self.kernel, self.kernel_type = kern[0].input_dim, outputs, kern, kernel_type
        self.M, self.fixed_mean, self.full_cov, self.prior_type = n_inducing, fixed_mean, full_cov, prior_type
        dev = torch.device("cuda", device)
        X_ini_val = np.zeros((Y_len + 1, x_dims_l))
        X_ini_val[0] = X_0_ini                                       # :56-58
        X_ini_val[1:] = X_train_ini
        self.X = torch.as_tensor(X_ini_val, dtype=torch.float64, device=dev)
        self.X_trainable = not (X_PG or case_val == 7)               # :62-66
        self.U = torch.as_tensor(np.asarray(U_ini, dtype=np.float64), device=dev).contiguous()
        self.Z = torch.as_tensor(np.asarray(ZZ, dtype=np.float64), device=dev).contiguous()
        self.U_trainable, self.Z_trainable = bool(U_optimization), bool(Z_optimization)
        self.Lm = None

    def conditional(self, X):
        # dgp_model.py:97-103 is broken in the reference (return_Lm=True): same error here
        return conditionals_multi_output.conditional(X, self.Z, self.kernel, self.U, white=True, full_cov=self.full_cov, return_Lm=True)

    def prior_Z(self):
        if self.prior_type == "uniform":
            return 0.0
        if self.prior_type == "normal":
            return -(self.Z * self.Z).sum() / 2.0                    # :108-109
        raise NotImplementedError("prior_type %r is not on the default path (SURVEY 2.1)" % self.prior_type)

    def prior_U(self):
        return -0.5 * (self.U * self.U).sum()                        # :132-135


class DGPSSM(BaseModel):
    """`dgp_model.py:159-359`.  Extra keyword arguments (all optional): `device` (GPU index) and
    `X_samples` (S,T+1,D) to evaluate S trajectories that share every other parameter."""

    def __init__(self, Y, x_dims, n_inducing, kernels, likelihood, minibatch_size, window_size, output_dim=None,
                 prior_type="uniform", full_cov=False, epsilon=0.01, mdecay=0.05, QQ_chol=None, ZZ=None, variance=None,
                 lengthscales=None, control_inputs=None, kernel_type='SquaredExponential', kernel_train_flag=True,
                 U_ini=None, X_0_ini=None, X_train_ini=None, X_PG=False, PG_particles=100, hyperparameter_sampling=False,
                 kernel_optimization=False, U_optimization=False, U_collapse=False, Z_optimization=False, case_val=1,
                 device=0, X_samples=None):
        import torch
        self.x_dims, self.n_inducing, self.kernels, self.likelihood = x_dims, n_inducing, kernels, likelihood
        self.window_size = window_size
        self.PG_particles = PG_particles
        self.output_dim = output_dim or x_dims[-1]
        self.U_collapse = bool(U_collapse)
        self.prior_type = prior_type
        self.device = torch.device("cuda", device)
        if len(kernels) != 1:
            raise NotImplementedError("the reference driver only ever builds one layer (FFVD_Main.py:373)")
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        T = Y.shape[0]
        self.layers = [Layer(ZZ, U_ini, X_0_ini, X_train_ini, kernels[0], self.output_dim, n_inducing, False, x_dims[0], T,
                             False, prior_type=prior_type, kernel_type=kernel_type, U_optimization=U_optimization,
                             U_collapse=U_collapse, Z_optimization=Z_optimization, X_PG=X_PG, case_val=case_val, device=device)]
        layer = self.layers[-1]
        self.X_N = T + 1                                              # :205
        kern = kernels[0]
        dev = self.device
        t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev).contiguous()
        self.kind = kern[0].kind
        D = self.output_dim
        logv = t64(np.array([float(np.asarray(k.logvariance)) for k in kern]))
        logl = t64(np.stack([np.asarray(k.loglengthscales, dtype=np.float64) * np.ones(k.input_dim) for k in kern])) \
            if self.kind == _capi.KERNEL_SE else None
        for i, k in enumerate(kern):                                  # kernel objects now view the model's tensors
            k.logvariance = logv[i]
            if logl is not None:
                k.loglengthscales = logl[i]
        logQ = t64(np.ones(D) * np.log(0.1)) if QQ_chol is None else t64(2.0 * np.log(np.asarray(QQ_chol, dtype=np.float64)))   # :176-185
        ctrl = None
        if control_inputs is not None and np.asarray(control_inputs).shape[0] > 0:
            ctrl = t64(np.asarray(control_inputs, dtype=np.float64)[:T])   # :254-255 (full batch)
        X = layer.X if X_samples is None else t64(X_samples)
        self.params: Dict[str, object] = dict(X=X, Z=layer.Z, U=layer.U, logv=logv, logQ=logQ, C=t64(likelihood.CC),
                                              d=t64(likelihood.DD), logR=t64(likelihood.log_Rchols))
        if logl is not None:
            self.params["logl"] = logl
        likelihood.CC, likelihood.DD, likelihood.log_Rchols = self.params["C"], self.params["d"], self.params["logR"]
        self.log_Q = logQ
        self.data = dict(Y=t64(Y), ctrl=ctrl)
        variables = sghmc_variable_names(self.kind, case_val, kernel_optimization, kernel_train_flag, U_optimization,
                                         U_collapse, Z_optimization, hyperparameter_sampling)
        # ---- Adam-trainable set (tf `trainable=` flags: :62-69,141-156 kernels, :182-184, likelihoods.py:17-24,50-54)
        tr = []
        if layer.X_trainable: tr.append("X")
        if U_optimization: tr.append("U")
        if Z_optimization: tr.append("Z")
        if kernel_optimization and self.kind == _capi.KERNEL_SE: tr += ["logv", "logl"]
        if not hyperparameter_sampling and case_val != 7: tr.append("logQ")
        if getattr(likelihood, "trainable", True): tr += ["C", "d", "logR"]
        self.trainable = [n for n in tr if n not in variables]
        self.ctx = context_for(X)
        self.flags = _capi.FLAG_PRIOR_Z_NORMAL if prior_type == "normal" else 0
        if prior_type not in ("normal", "uniform"):
            raise NotImplementedError("prior_type %r is not on the default path (SURVEY 2.1)" % prior_type)
        super().__init__(Y, variables, minibatch_size, window_size)
        self._out = None
        self.generate_update_step(None, epsilon, mdecay)               # :302

    # ------------------------------------------------------------------------------------------
    def _alloc_outputs(self):
        import torch
        X = self.params["X"]
        S = 1 if X.dim() == 2 else X.shape[0]
        out = dict(nll=torch.empty(S, dtype=torch.float64, device=self.device),
                   terms=torch.empty(S, 6, dtype=torch.float64, device=self.device))
        for k, v in self.params.items():
            out["g_" + k] = torch.empty_like(v)
        return out

    def evaluate(self, flags_extra: int = 0):
        """One evaluation of nll (dgp_model.py:248-297) and all its gradients (base_model.py:148)."""
        if self._out is None:
            self._out = self._alloc_outputs()
            self._prepared = {}
        call = self._prepared.get(flags_extra)
        if call is None:
            # tensors are bound once (they are only ever updated in place); 21 evaluations per sghmc_step re-use the binding
            prob = dict(self.params)
            prob.update(self.data)
            prob.setdefault("logl", None)
            call = self.ctx.prepare_nll_grads(self.kind, self.U_collapse, prob, self._out, flags=self.flags | flags_extra, jitter=1e-5)
            self._prepared[flags_extra] = call
        call.run()
        return self._out

    @property
    def nll(self):
        out = self.evaluate()
        return out["nll"][0] if self.params["X"].dim() == 2 else out["nll"]

    def nll_terms(self) -> Dict[str, object]:
        """The reference's per-term tensors (`print_sample_performance`, base_model.py:952-989)."""
        out = self.evaluate()
        return {n: out["terms"][:, i] for i, n in enumerate(TERM_NAMES)}

    def regularizer(self, X_batch, control_inputs_batch):
        """`dgp_model.py:337-359`: (reg_trace_inverse_Q_B (T,), reg_x_prior (T,)) of the uncollapsed form."""
        import torch
        from .likelihoods import logdensity_norm_diag
        Xb = X_batch if torch.is_tensor(X_batch) else torch.as_tensor(np.asarray(X_batch), dtype=torch.float64, device=self.device)
        if control_inputs_batch is not None and len(control_inputs_batch) > 0:
            cb = control_inputs_batch if torch.is_tensor(control_inputs_batch) else torch.as_tensor(
                np.asarray(control_inputs_batch), dtype=torch.float64, device=self.device)
            xc = torch.cat((Xb[:-1], cb), dim=1).contiguous()
        else:
            xc = Xb[:-1].contiguous()
        mean_reg, var_reg = conditionals_multi_output.conditional(xc, self.params["Z"], self.kernels[-1], self.params["U"], white=True)
        mean_reg = mean_reg + Xb[:-1]
        Q = self.log_Q.exp()
        reg_trace = -0.5 * ((Q[None, :] ** (-1)) * var_reg).sum(dim=1)
        reg_x_prior = logdensity_norm_diag(Xb[1:].contiguous(), mean_reg.contiguous(), Q ** 0.5)
        return reg_trace, reg_x_prior
