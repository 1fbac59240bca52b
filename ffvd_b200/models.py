"""Mirror of `vfegpssm/models.py`: `Model` / `RegressionModel` -- argument holder, model construction (:47-74) and the
outer training loop (:142-168: per iteration one `sghmc_step` (21 nll+gradient evaluations) and one Adam
`train_hypers` step, plus a particle-Gibbs sweep over X in case 6).  TensorBoard summaries and timing hooks of the
reference loop are out of scope."""
from __future__ import annotations

import numpy as np

from .dgp_model import DGPSSM
from .kernels import LinearK as LinearKernel
from .kernels_multi_output import SquaredExponential as BgpSE
from .likelihoods import Gaussian

PRIORS = ['uniform', 'normal', 'determinantal', 'strauss']


class Model(object):
    def __init__(self, prior_type, output_dim=None):
        class ARGS:                                   # models.py:21-32
            num_inducing = 100
            iterations = 1000
            minibatch_size = 100
            window_size = 64
            num_posterior_samples = 20
            posterior_sample_spacing = 50
            full_cov = True
            n_layers = 1
            prior_type = None
            logdir = '/tmp/'
            x_dims = [1]

        self.ARGS = ARGS
        self.model = None
        self.output_dim = output_dim
        self.global_step = 0
        if prior_type in PRIORS:
            self.ARGS.prior_type = prior_type
        else:
            raise Exception("Invalid prior type")

    def _fit(self, Y_train, tensorboard_savepath, dataname, fileid, lik, X_train, Ystd, kernel_type, kernel_train_flag,
             Y_test=None, data_uu=None, progress=None, sghmc_noise=None, window_index=None, **kwargs):
        """`sghmc_noise(it, k)` -> {name: N(0,1) array} for SG-HMC evaluation k (0..20) of outer iteration `it`, and
        `window_index(it)` -> the window entry `train_hypers` feeds, inject the randomness the reference draws unseeded
        (tf.random.normal, np.random.randint) so that a run can be compared with the reference's (tests)."""
        Y_train = np.asarray(Y_train, dtype=np.float64)
        if len(Y_train.shape) == 1:
            Y_train = Y_train[:, None]
        A = self.ARGS
        if not self.model:
            kerns = []
            control = np.asarray(A.control_inputs, dtype=np.float64)
            for i in range(A.n_layers):
                Z_dim = control.shape[1] + A.x_dims[-1]
                if kernel_type == 'SquaredExponential':
                    kerns.append([BgpSE(Z_dim, ARD=True, variance=A.variance[kk], lengthscales=A.lengthscales[kk],
                                        kernel_optimization=A.kernel_optimization) for kk in range(A.x_dims[-1])])
                elif kernel_type == 'LinearK':
                    setting_variance = 1.0 if A.variance is None else A.variance
                    kerns.append(LinearKernel(Z_dim, ARD=False, variance=setting_variance))      # models.py:61-62
            mb_size = A.minibatch_size if Y_train.shape[0] > A.minibatch_size else Y_train.shape[0]
            self.model = DGPSSM(Y_train, A.x_dims, A.num_inducing, kerns, lik, minibatch_size=mb_size, window_size=A.window_size,
                                full_cov=A.full_cov, prior_type=A.prior_type, output_dim=self.output_dim, QQ_chol=A.QQ_chol,
                                ZZ=A.ZZ, variance=A.variance, lengthscales=A.lengthscales, control_inputs=control,
                                kernel_type=kernel_type, kernel_train_flag=kernel_train_flag, U_ini=A.UU_ini, X_0_ini=A.XX_0_ini,
                                X_train_ini=A.x_initialization, X_PG=A.X_PG, PG_particles=getattr(A, 'PG_particles', 100),
                                hyperparameter_sampling=A.hyperparameter_sampling, kernel_optimization=A.kernel_optimization,
                                U_optimization=A.U_optimization, U_collapse=A.U_collapse, Z_optimization=A.Z_optimization,
                                case_val=A.case_val, **kwargs)
        self.nll_seq, self.rmse_seq, self.ll_seq, self.running_time_seq = [], [], [], []
        it = 0
        while it < 2 * A.iterations:                  # models.py:142-168
            it += 1
            self.global_step += 1
            self.model.global_step = self.global_step
            self.model.sghmc_step(noise_fn=(lambda k, _it=it - 1: sghmc_noise(_it, k)) if sghmc_noise is not None else None)
            if A.X_PG:                                # case 6: models.py:155-158
                self.model.gp_x_sampling()
            if self.model.trainable:                  # hasattr(self.model, 'hyper_train_op')
                self.model.train_hypers(window_index(it - 1) if window_index is not None else None)
            if progress is not None and it % 100 == 0:
                progress(it)
        return self.model


class RegressionModel(Model):
    def __init__(self, prior_type, output_dim=None):
        super().__init__(prior_type, output_dim)

    def fit(self, Y_train, Y_test=None, tensorboard_savepath='', dataname='', fileid='', kernel_type='SquaredExponential',
            kernel_train_flag=True, likelihood_traning=True, X_train=None, X_test=None, Ystd=None, data_uu=None, sghmc_noise=None,
            window_index=None, **kwargs):
        Y_train = np.asarray(Y_train, dtype=np.float64)
        lik = Gaussian(Y_train.shape[1], self.ARGS.x_dims[-1], CC=self.ARGS.CC, DD=self.ARGS.DD, RR_chol=self.ARGS.RR_chol,
                       hyperparameter_sampling=self.ARGS.hyperparameter_sampling, likelihood_traning=likelihood_traning)
        return self._fit(Y_train, tensorboard_savepath, dataname, fileid, lik, X_train, Ystd, kernel_type, kernel_train_flag,
                         Y_test=Y_test, data_uu=data_uu, sghmc_noise=sghmc_noise, window_index=window_index, **kwargs)


def configure(model: Model, arguments: dict, control_inputs, case_val, Y_train_std=1.0, **overrides):
    """`FFVD_Main.py:231-324`: copy the warm-start fields (`datasets.arguments_from_factnonlin`) and the case flags
    onto `model.ARGS`."""
    from .datasets import CASE_TABLE
    A = model.ARGS
    for k, v in arguments.items():
        setattr(A, k, v)
    A.control_inputs = np.asarray(control_inputs, dtype=np.float64)
    A.Y_train_std = Y_train_std
    A.full_cov = False                                # FFVD_Main.py:267
    A.x_dims = [int(np.asarray(arguments['XX_0_ini']).shape[0])]
    A.num_inducing = int(np.asarray(arguments['ZZ']).shape[0])
    A.case_val = case_val
    A.hyperparameter_sampling = False
    A.kernel_optimization, A.U_optimization, A.Z_optimization, A.U_collapse, A.X_PG = CASE_TABLE[case_val]
    for k, v in overrides.items():
        setattr(A, k, v)
    return model
