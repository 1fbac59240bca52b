"""Golden vectors for the "next" rows of SURVEY section 8(f), again from the REFERENCE'S OWN SOURCE executed over the
torch-backed `tensorflow` shim (see make_reference_golden.py for the approach and its caveat).

What runs unmodified from /root/reference here:
  f2  conditionals_multi_output.collapse_u_mean_after_kernel_precalculation (:206-227),
      conditional_after_kernel_precalculation with q_sqrt (:306-387) and BaseModel.collect_samples_formal
      (base_model.py:197-522: the posterior roll-out, predict_y / predict_y_var / fit_y / RMSE and the _results.npz file);
  f4  BaseModel.PG_for_X (base_model.py:29-75), with the tensorflow_probability draws served by the shim and logged;
  f3  the outer training loop: every `session.run(burn_in_op | sample_op | hyper_train_op)` of `sghmc_step`
      (base_model.py:915-933) and `train_hypers` (:944-950) is emulated by RE-RUNNING the reference's model constructor
      on the current variable values (tf.VARIABLE_OVERRIDES) and reading the update expressions it builds -- the TF1
      meaning of re-evaluating the graph.  Restated here: the schedule of those two methods (the loop glue), the window
      feed, and TF1's AdamOptimizer apply rule (TensorFlow's, not the reference's, code).

Run in the build container:  python tests/golden/make_reference_golden_next.py
Writes tests/golden/reference_shim_golden_next.npz.
"""
import glob
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G  # noqa: E402  (sets up the shim, chdir to the reference, imports vfegpssm)

tf = G.tf
REF = G.REF
from vfegpssm import base_model, conditionals_multi_output as cmo  # noqa: E402

n = G.n


def first_init(ds, idx=0):
    return sorted(glob.glob(os.path.join(REF, "Factnonlin_ini", "factnonlin_initialized_10000_" + ds + "*.npz")))[idx]


def golden_f2(out):
    ds = "actuator"
    model, lik, kerns, (Y_train, ctrl_full) = G.build_model(ds, first_init(ds), 4, seed=11)
    layer = model.layers[-1]
    T = Y_train.shape[0]
    Linv = cmo.kernel_pre_cal(layer.Z, kerns)
    xc = tf.concat((layer.X[:T], tf.constant(ctrl_full[:T])), axis=1)
    U_val, Lseq = cmo.collapse_u_mean_after_kernel_precalculation(Linv, xc, layer.X, layer.Z, kerns, model.Q)
    out["f2/U_mean"] = n(U_val)              # (1, M, D)
    out["f2/LHinvT"] = n(Lseq)               # (D, M, M)
    Xnew = xc[:37]
    out["f2/Xnew"] = n(Xnew)
    for tag, q in (("q3", Lseq), ("qnone", None)):
        mu, var = cmo.conditional_after_kernel_precalculation(Linv, Xnew, layer.Z, kerns, U_val[0], white=True, full_cov=False, q_sqrt=q)
        out["f2/cond_%s/mean" % tag] = n(mu); out["f2/cond_%s/var" % tag] = n(var)
    # ---- the reference's own roll-out + results file, collapsed (case 4) and uncollapsed (case 2)
    Y_test = FFVD_test(ds)
    for case_val, tag in ((4, "collapsed"), (2, "uncollapsed")):
        model, lik, kerns, (Y_train, ctrl_full) = G.build_model(ds, first_init(ds), case_val, seed=21)
        tf.NOISE_LOG.clear()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "run")
            model.collect_samples_formal(3, 1, ctrl_full, 30, sghmc_var_len=0, U_collapse=(case_val == 4), Y_test=Y_test,
                                         Y_train_std=1.7, save_path_file=path, Y_train=Y_train, case="C%d" % case_val)
            res = np.load(path + "_results.npz", allow_pickle=True)
            for k in ("y_train_vfe", "y_test_vfe", "v_test_vfe_var", "X_val", "U_val", "Z_val", "log_QQ"):
                out["f2/rollout_%s/file/%s" % (tag, k)] = np.asarray(res[k], dtype=np.float64)
            out["f2/rollout_%s/file_keys" % tag] = np.array(sorted(res.files))
        out["f2/rollout_%s/noise" % tag] = np.stack([n(z) for z in tf.NOISE_LOG]).reshape(3, 30, -1)   # (num, L, D) in draw order
        out["f2/rollout_%s/predict_y" % tag] = np.asarray(model.predict_y)
        out["f2/rollout_%s/predict_y_var" % tag] = np.asarray(model.predict_y_var)
        out["f2/rollout_%s/fit_y" % tag] = np.asarray(model.fit_y)
        out["f2/rollout_%s/RMSE" % tag] = float(model.RMSE_val)
        print("f2 rollout", tag, "RMSE", model.RMSE_val, flush=True)
    out["f2/Y_test"] = Y_test
    out["f2/ctrl_future"] = np.asarray(ctrl_full[Y_train.shape[0]:Y_train.shape[0] + 30], dtype=np.float64)   # control_inputs[test_i + T]


def FFVD_test(ds):
    return G.FFVD_Main.create_dataset(ds + "/")[1]


def golden_f4(out):
    ds = "gas_furnace"                      # the shortest series (T = 148): whole-trajectory particles stay cheap
    P = 6
    model, lik, kerns, (Y_train, ctrl_full) = G.build_model(ds, first_init(ds), 6 if 6 in G.CASES else 1, seed=31)
    tf.NOISE_LOG.clear(); tf.UNIFORM_LOG.clear()
    T = Y_train.shape[0]
    X_before = n(model.layers[-1].X)
    base_model.BaseModel.PG_for_X(model, tf.constant(ctrl_full[:T]), P)
    draws = [n(z) for z in tf.NOISE_LOG]
    out["f4/P"] = P
    out["f4/X_before"] = X_before
    out["f4/normals"] = draws[0].reshape(P - 1, -1)                     # Normal(...).sample(): (P-1,1,D)
    out["f4/eps"] = np.stack(draws[1:])                                 # (T, P-1, D): one tf.random.normal per step
    u = [n(z) for z in tf.UNIFORM_LOG]
    U = np.zeros((T, P - 1))
    for t, v in enumerate(u):
        U[t, :v.shape[0]] = v
    out["f4/uniforms"] = U
    out["f4/X_after"] = n(model.layers[-1].X)
    print("f4 PG sweep: moved %d of %d rows" % (int(np.sum(np.any(out["f4/X_after"] != X_before, axis=1))), X_before.shape[0]), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class GraphRerun:
    """TF1 semantics on the eager shim: the variable store lives here (creation index -> value); every `run(op)` rebuilds
    the reference model on the current values and applies the (variable, new value) pairs of the requested op."""

    def __init__(self, ds, ini, case_val):
        self.ds, self.ini, self.case_val = ds, ini, case_val
        self.store = {}
        self.evals = 0
        self.adam_m, self.adam_v, self.adam_t = {}, {}, 0
        self.noise = []          # per SG-HMC evaluation: list (per sampled variable) of the tf.random.normal draws

    def build(self, feed=None):
        tf.reset_shim(1000 + self.evals, keep_overrides=True)
        tf.VARIABLE_OVERRIDES.clear()
        tf.VARIABLE_OVERRIDES.update(self.store)
        if feed:
            tf.VARIABLE_OVERRIDES.update(feed)
        self.evals += 1
        model, lik, kerns, _ = G.build_model_keep(self.ds, self.ini, self.case_val)
        return model, lik, kerns

    def run_sghmc(self, burn_in):
        model, _, _ = self.build()
        if not model.vars:
            return model
        self.noise.append([n(z) for z in tf.NOISE_LOG[:len(model.vars)]])
        pairs = model.burn_in_op if burn_in else model.sample_op
        for var, val in pairs:                       # all right-hand sides were formed from the pre-step values (Jacobi)
            self.store[var._tf_index] = val.detach().clone()
        return model

    def snapshot(self, model):
        return {v._tf_index: self.store.get(v._tf_index, v.detach().clone()) for v in model.vars}

    def run_adam(self, feed, lr):
        model, _, _ = self.build(feed)
        if not hasattr(model, "hyper_train_op"):
            return model
        self.adam_t += 1
        t = self.adam_t
        lr_t = lr * np.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)          # tf.compat.v1.train.AdamOptimizer, defaults
        for g, v in model.hyper_train_op:
            i = v._tf_index
            m = self.adam_m.get(i, torch.zeros_like(g)); vv = self.adam_v.get(i, torch.zeros_like(g))
            m = 0.9 * m + 0.1 * g
            vv = 0.999 * vv + 0.001 * g * g
            self.adam_m[i], self.adam_v[i] = m, vv
            cur = self.store.get(i, v.detach())      # the fed window value is temporary: the update applies to the stored value
            self.store[i] = (cur - lr_t * m / (torch.sqrt(vv) + 1e-8)).detach().clone()
        return model


def golden_f3(out):
    ds = "actuator"
    for case_val, iters in ((2, 2), (4, 4)):      # even: Model._fit runs 2 * ARGS.iterations outer iterations (models.py:142)
        gr = GraphRerun(ds, first_init(ds), case_val)
        window, widx = [], []
        model = None
        for it in range(iters):
            # sghmc_step, base_model.py:915-933
            model = gr.run_sghmc(True)
            for _ in range(10):
                gr.run_sghmc(True)
                model = gr.run_sghmc(False)
            window.append(gr.snapshot(model))
            # train_hypers, base_model.py:944-950: feed a window entry (index injected: the reference draws np.random.randint)
            i = (7 * it + 3) % len(window)
            widx.append(i)
            lr = 0.003 * (0.95 ** (1 / 1000))                              # get_minibatch() default global_step = 1
            model = gr.run_adam(window[i] if model.vars else None, lr)
        # final values of every variable, by the name of the parameter it is
        model, lik, kerns = gr.build()
        layer = model.layers[-1]
        key = "f3/case%d" % case_val
        out[key + "/iters"] = iters
        out[key + "/window_index"] = np.array(widx)
        out[key + "/X"] = n(layer.X); out[key + "/Z"] = n(layer.Z); out[key + "/U"] = n(layer.U)
        out[key + "/logQ"] = n(model.log_Q); out[key + "/C"] = n(lik.CC); out[key + "/d"] = n(lik.DD); out[key + "/logR"] = n(lik.log_Rchols)
        out[key + "/logv"] = np.array([float(k.logvariance) for k in kerns]); out[key + "/logl"] = np.stack([n(k.loglengthscales) for k in kerns])
        out[key + "/nll_final"] = float(model.nll)
        out[key + "/nvars"] = len(model.vars)
        out[key + "/trainable"] = np.array([getattr(v, "_tf_name", None) or "?" for v in tf.compat.v1.trainable_variables()])
        if gr.noise:
            for vi in range(len(gr.noise[0])):
                out[key + "/noise_var%d" % vi] = np.stack([e[vi] for e in gr.noise])      # (evaluations, *shape)
        print(key, "evals", gr.evals, "nll_final %.12f" % out[key + "/nll_final"], "sampled vars", len(model.vars), flush=True)


def golden_conditional_options(out):
    """base_conditional's other branches (conditionals.py:27-66): full_cov, q_sqrt 2-d / 3-d whitened or not, return_Lm --
    the single-kernel conditional on a small SE problem, and the multi-output one with full_cov (its (N,1,D) quirk)."""
    from vfegpssm import conditionals
    tf.reset_shim(9)
    rng = np.random.default_rng(9)
    M, R, N, Din = 23, 3, 11, 4
    Zs = rng.standard_normal((M, Din)) * 1.5; Xn = rng.standard_normal((N, Din)); fm = rng.standard_normal((M, R))
    q3 = np.tril(rng.standard_normal((R, M, M))) * 0.2; q2 = rng.uniform(0.1, 1.0, (M, R))
    ls = rng.uniform(1.0, 3.0, Din)
    se = G.BgpSE(Din, variance=0.7, lengthscales=ls, ARD=True)
    out["cond/Z"] = Zs; out["cond/Xnew"] = Xn; out["cond/f"] = fm; out["cond/q3"] = q3; out["cond/q2"] = q2; out["cond/ls"] = ls
    c = lambda a: tf.constant(a)
    for tag, kw in (("full", dict(full_cov=True, white=True)), ("full_q3", dict(full_cov=True, white=True, q_sqrt=c(q3))),
                    ("full_q2", dict(full_cov=True, white=True, q_sqrt=c(q2))), ("nonwhite_q3", dict(white=False, q_sqrt=c(q3))),
                    ("nonwhite_q2", dict(white=False, q_sqrt=c(q2))), ("full_nonwhite_q3", dict(full_cov=True, white=False, q_sqrt=c(q3)))):
        mu, var = conditionals.conditional(c(Xn), c(Zs), se, c(fm), **kw)
        out["cond/%s/mean" % tag] = n(mu); out["cond/%s/var" % tag] = n(var)
    mu, var, Lm = conditionals.conditional(c(Xn), c(Zs), se, c(fm), white=True, return_Lm=True)
    out["cond/return_Lm/mean"] = n(mu); out["cond/return_Lm/var"] = n(var); out["cond/return_Lm/Lm"] = n(Lm)
    se2 = G.BgpSE(Din, variance=0.3, lengthscales=ls[::-1].copy(), ARD=True)
    mu, var = cmo.conditional(c(Xn), c(Zs), [se, se2], c(fm[:, :2]), white=True, full_cov=True)
    out["cond/multi_full/mean"] = n(mu); out["cond/multi_full/var"] = n(var)
    print("conditional options: multi full_cov var shape", out["cond/multi_full/var"].shape, flush=True)


def main():
    out = {}
    golden_conditional_options(out)
    golden_f2(out)
    golden_f4(out)
    golden_f3(out)
    path = os.path.join(HERE, "reference_shim_golden_next.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
