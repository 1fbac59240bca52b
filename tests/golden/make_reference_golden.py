"""Golden vectors from the REFERENCE'S OWN SOURCE, executed here over a torch-backed
`tensorflow` shim (tests/golden/tf_shim) because TensorFlow itself is not installable.

What runs unmodified from /root/reference: FFVD_Main.create_dataset (data standardisation),
vfegpssm.kernels{,_multi_output}, conditionals{,_multi_output}, likelihoods.Gaussian,
dgp_model.{Layer,DGPSSM} (nll graph assembly, SG-HMC variable selection) and
base_model.BaseModel.generate_update_step (SG-HMC update expressions, tf.gradients).
Only the wiring of `Model._fit` (models.py:47-74) is restated below, and the out-of-scope
particle-Gibbs graph builder (`PG_for_X_speedup`, needs tensorflow_probability) is stubbed.

Run in the build container:  python tests/golden/make_reference_golden.py
Writes tests/golden/reference_shim_golden.npz.
"""
import collections
import collections.abc
import glob
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, REF)
collections.Iterable = collections.abc.Iterable       # quadrature.py:16 (Python < 3.10 idiom)

import tensorflow as tf  # noqa: E402  (the shim)

os.chdir(REF)                                          # create_dataset uses the relative path 'data/'
import FFVD_Main  # noqa: E402
from vfegpssm import base_model, dgp_model  # noqa: E402
from vfegpssm.kernels import LinearK  # noqa: E402
from vfegpssm.kernels import SquaredExponential as SE_alias  # noqa: E402,F401
from vfegpssm.kernels_multi_output import SquaredExponential as BgpSE  # noqa: E402
from vfegpssm import conditionals, conditionals_multi_output  # noqa: E402
from vfegpssm.likelihoods import Gaussian, logdensity_norm, logdensity_norm_diag, logdensity_norm_diag_nonvec  # noqa: E402

base_model.BaseModel.PG_for_X_speedup = lambda self, *a, **k: None     # out of scope

CASES = {  # FFVD_Main.py:273-324 : (kernel_optimization, U_optimization, Z_optimization, U_collapse)
    1: (True, True, True, False), 2: (False, False, True, False), 3: (False, False, False, False),
    4: (True, False, True, True), 5: (False, False, True, True), 7: (False, False, False, False),
}


def build_model_keep(ds, ini_file, case_val):
    """build_model without resetting the shim (make_reference_golden_next.py re-runs the constructor on a kept variable store)."""
    return build_model(ds, ini_file, case_val, seed=None)


def build_model(ds, ini_file, case_val, seed):
    if seed is not None:
        tf.reset_shim(seed)
    Y_train, Y_test, control_inputs, Y_std, Y_mean, u_mean, u_std = FFVD_Main.create_dataset(ds + "/")
    f = np.load(ini_file, allow_pickle=True)
    T = Y_train.shape[0]
    ko, uo, zo, uc = CASES[case_val]
    # FFVD_Main.py:245-254, 340
    CC = tf.convert_to_tensor(f["C_val"].T, dtype=tf.float64)
    DD = tf.convert_to_tensor(f["d_val"], dtype=tf.float64)
    RR = tf.convert_to_tensor(f["R_chol_val"], dtype=tf.float64)
    ZZ = tf.convert_to_tensor(f["Z_val"], dtype=tf.float64)
    x_init = np.mean(f["x_samples_training"], axis=1)
    lik = Gaussian(Y_train.shape[1], 4, CC=CC, DD=DD, RR_chol=RR)           # models.py:320
    Z_dim = control_inputs.shape[1] + 4
    kerns = [[BgpSE(Z_dim, ARD=True, variance=f["kernel_variance"][kk], lengthscales=f["kernel_lengthscales"][kk],
                    kernel_optimization=ko) for kk in range(4)]]            # models.py:54-59
    tf.PLACEHOLDER_VALUES[tf.int64] = [0, T + 1]                             # base_model.py:194 full batch
    tf.PLACEHOLDER_VALUES[tf.float64] = torch.tensor(0.003, dtype=torch.float64)
    model = dgp_model.DGPSSM(Y_train, [4], 100, kerns, lik, minibatch_size=min(1000, T), window_size=64,
                             full_cov=False, prior_type="normal", output_dim=None, QQ_chol=f["Q_sqrt_ini"], ZZ=ZZ,
                             variance=f["kernel_variance"], lengthscales=f["kernel_lengthscales"],
                             control_inputs=control_inputs, kernel_type="SquaredExponential", kernel_train_flag=True,
                             U_ini=f["Umu_ini"].T, X_0_ini=f["qx1_mu_ini"], X_train_ini=x_init, X_PG=False,
                             PG_particles=100, hyperparameter_sampling=False, kernel_optimization=ko,
                             U_optimization=uo, U_collapse=uc, Z_optimization=zo, case_val=case_val)   # models.py:66-74
    return model, lik, kerns[0], (Y_train, control_inputs)


def n(x):
    return x.detach().numpy().copy() if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)


def all_grads(model, lik, kerns):
    layer = model.layers[-1]
    xs = [layer.X, layer.Z, layer.U, model.log_Q, lik.CC, lik.DD, lik.log_Rchols]
    xs += [k.logvariance for k in kerns] + [k.loglengthscales for k in kerns]
    gs = tf.gradients(model.nll, xs)
    out = {"g_X": n(gs[0]), "g_Z": n(gs[1]), "g_U": n(gs[2]), "g_logQ": n(gs[3]), "g_C": n(gs[4]), "g_d": n(gs[5]),
           "g_logR": n(gs[6]), "g_logv": np.array([float(g) for g in gs[7:11]]), "g_logl": np.stack([n(g) for g in gs[11:15]])}
    return out


def main():
    out = {}
    # ---- nll + gradients, collapsed (case 4, CLI default) and uncollapsed (case 2), first two inits per dataset
    for ds in ("dryer", "drive", "gas_furnace", "actuator", "flutter", "ballbeam"):
        files = sorted(glob.glob(os.path.join(REF, "Factnonlin_ini", "factnonlin_initialized_10000_" + ds + "*.npz")))
        for idx in (0, 1):
            for case_val, tag in ((4, "collapsed"), (2, "uncollapsed")):
                model, lik, kerns, _ = build_model(ds, files[idx], case_val, seed=idx)
                key = "%s/%d/%s" % (ds, idx, tag)
                out[key + "/nll"] = float(model.nll)
                terms = [model.nll_part_prior, model.nll_log_likelihood, model.x_t_prior_Q, model.nll_reg_trace_inverse_Q_B]
                terms += [model.later_term1, model.later_term2] if case_val == 4 else [0.0, 0.0]
                out[key + "/terms"] = np.array([float(t) for t in terms])
                for k, v in all_grads(model, lik, kerns).items():
                    out[key + "/" + k] = v
                print(key, "nll = %.12f" % out[key + "/nll"], flush=True)
    # ---- SG-HMC update expressions (base_model.py:143-179) on the reference's own variable set: case 2
    # (kernel hypers + U sampled) and case 7 (U and X sampled).  Second step is run on the state after the first.
    ds = "actuator"
    files = sorted(glob.glob(os.path.join(REF, "Factnonlin_ini", "factnonlin_initialized_10000_" + ds + "*.npz")))
    for case_val in (2, 7):
        model, lik, kerns, _ = build_model(ds, files[0], case_val, seed=123)
        key = "sghmc/case%d" % case_val
        nv = len(model.vars)
        out[key + "/nvars"] = nv
        out[key + "/X_N"] = float(model.X_N)
        # burn_in_op = burn_in_updates (xi,g,g2 per var) + sample_updates (theta,p per var)  base_model.py:162-179
        bi = model.burn_in_op
        for i in range(nv):
            theta = model.vars[i]
            (xi, xi_t), (g, g_t), (g2, g2_t) = bi[3 * i], bi[3 * i + 1], bi[3 * i + 2]
            (th, th_t), (p, p_t) = bi[3 * nv + 2 * i], bi[3 * nv + 2 * i + 1]
            assert th is theta
            grad = tf.gradients(model.nll, [theta])[0]
            pre = "%s/var%d/" % (key, i)
            out[pre + "theta"] = n(theta); out[pre + "grad"] = n(grad); out[pre + "noise"] = n(tf.NOISE_LOG[i])
            out[pre + "xi"] = n(xi); out[pre + "g"] = n(g); out[pre + "g2"] = n(g2); out[pre + "p"] = n(p)
            out[pre + "xi_t"] = n(xi_t); out[pre + "g_t"] = n(g_t); out[pre + "g2_t"] = n(g2_t)
            out[pre + "theta_t"] = n(th_t); out[pre + "p_t"] = n(p_t)
        print(key, "vars", nv, flush=True)
    # ---- operator level: LinearK through both conditionals, SE K/Kdiag, log-densities
    tf.reset_shim(7)
    rng = np.random.default_rng(7)
    Zs = rng.standard_normal((20, 3)); Xn = rng.standard_normal((33, 3)); fm = rng.standard_normal((20, 2))
    lk = LinearK(3, variance=1.0)
    mu, var = conditionals.conditional(tf.constant(Xn), tf.constant(Zs), lk, tf.constant(fm), white=True)
    out["op/linear_single/Z"] = Zs; out["op/linear_single/Xnew"] = Xn; out["op/linear_single/f"] = fm
    out["op/linear_single/mean"] = n(mu); out["op/linear_single/var"] = n(var)
    mu, var = conditionals.conditional(tf.constant(Xn), tf.constant(Zs), lk, tf.constant(fm), white=False)
    out["op/linear_single/mean_nonwhite"] = n(mu); out["op/linear_single/var_nonwhite"] = n(var)
    mu, var = conditionals_multi_output.conditional(tf.constant(Xn), tf.constant(Zs), [lk, lk], tf.constant(fm), white=True)
    out["op/linear_multi/mean"] = n(mu); out["op/linear_multi/var"] = n(var)
    out["op/linear/K"] = n(lk.K(tf.constant(Xn), tf.constant(Zs))); out["op/linear/Kdiag"] = n(lk.Kdiag(tf.constant(Xn)))
    se = BgpSE(3, variance=0.37, lengthscales=np.array([0.9, 1.7, 2.6]), ARD=True)
    se2 = BgpSE(3, variance=0.11, lengthscales=np.array([1.3, 0.8, 3.1]), ARD=True)
    out["op/se/K"] = n(se.K(tf.constant(Xn), tf.constant(Zs))); out["op/se/Kzz"] = n(se.K(tf.constant(Zs)))
    out["op/se/Kdiag"] = n(se.Kdiag(tf.constant(Xn)))
    mu, var = conditionals_multi_output.conditional(tf.constant(Xn), tf.constant(Zs), [se, se2], tf.constant(fm), white=True)
    out["op/se_multi/mean"] = n(mu); out["op/se_multi/var"] = n(var)
    Li = conditionals_multi_output.kernel_pre_cal(tf.constant(Zs), [se, se2])
    out["op/se_multi/LinvT"] = np.stack([n(a) for a in Li])
    y = rng.standard_normal((33, 2)); ym = rng.standard_normal((33, 2)); R = np.array([0.4, 1.3])
    out["op/ld/y"] = y; out["op/ld/ymean"] = ym; out["op/ld/R"] = R
    out["op/ld/diag"] = n(logdensity_norm_diag(tf.constant(y), tf.constant(ym), tf.constant(R)))
    out["op/ld/diag_nonvec"] = n(logdensity_norm_diag_nonvec(tf.constant(y), tf.constant(ym), tf.constant(R)))
    Rc = np.array([[0.4, 0.0], [0.3, 1.3]])
    out["op/ld/Rfull"] = Rc
    out["op/ld/full"] = n(logdensity_norm(tf.constant(y), tf.constant(ym), tf.constant(Rc)))
    path = os.path.join(HERE, "reference_shim_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
