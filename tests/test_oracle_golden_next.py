"""CPU tests: the oracle's restatements of the SURVEY 8(f) "next" rows against golden vectors produced by the REFERENCE'S
OWN SOURCE (tests/golden/make_reference_golden_next.py -> reference_shim_golden_next.npz): the collapsed q(u) and the
prediction-time conditional (cmo:206-227, 306-387), the posterior roll-out and its summary statistics
(base_model.py:197-522), the particle-Gibbs sweep (base_model.py:29-75) and the outer training loop (models.py:142-168 /
base_model.py:915-950).  These pin the oracle functions that the GPU tests of the same rows compare the CUDA path with."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, assert_close

from oracle import ffvd_oracle as O
from oracle import fixtures, loop

tt = torch.as_tensor


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "reference_shim_golden_next.npz"), allow_pickle=False)


@pytest.fixture(scope="module")
def packed():
    p = fixtures.load_packed()
    return {q.name: q for q in p["problems"]}


def _kern(prob):
    return O._make_kernels(tt(prob.logv), tt(prob.logl), 0, prob.Z.shape[1])


def test_collapsed_qu_and_predictive_conditional_vs_reference_source(gold, packed):
    prob = packed["actuator/0"]
    T = prob.Y.shape[0]
    kern = _kern(prob)
    Linv = O.kernel_pre_cal(tt(prob.Z), kern)
    Xc = np.concatenate([prob.X[:T], prob.ctrl], axis=1)
    U, Ls = O.collapse_u_mean_after_kernel_precalculation(Linv, tt(Xc), tt(prob.X), tt(prob.Z), kern, tt(np.exp(prob.logQ)))
    assert_close(gold["f2/U_mean"][0], U.numpy(), 1e-10, "U_mean")
    assert_close(gold["f2/LHinvT"], Ls.numpy(), 1e-10, "LHinvT")
    np.testing.assert_allclose(gold["f2/Xnew"], Xc[:37], rtol=0, atol=1e-15)
    for tag, q in (("q3", tt(gold["f2/LHinvT"])), ("qnone", None)):
        mu, var = O.conditional_after_kernel_precalculation(Linv, tt(Xc[:37]), tt(prob.Z), kern, tt(gold["f2/U_mean"][0]), white=True, q_sqrt=q)
        assert_close(gold["f2/cond_%s/mean" % tag], mu.numpy(), 1e-10, tag + " mean")
        assert np.max(np.abs(gold["f2/cond_%s/var" % tag] - var.numpy())) <= 1e-10 * float(np.max(np.exp(prob.logv)))


@pytest.mark.parametrize("tag", ("collapsed", "uncollapsed"))
def test_rollout_statistics_vs_reference_source(gold, packed, tag):
    """`collect_samples_formal` executed from the reference source vs the oracle's roll-out + the summary formulas
    (base_model.py:330-347) with the reference's logged noise."""
    prob = packed["actuator/0"]
    T, D = prob.Y.shape[0], prob.X.shape[1]
    kern = _kern(prob)
    Linv = O.kernel_pre_cal(tt(prob.Z), kern)
    Q = tt(np.exp(prob.logQ))
    ctrl_full = fixtures_ctrl_full("actuator", gold, prob)
    if tag == "collapsed":
        Xc = np.concatenate([prob.X[:T], prob.ctrl], axis=1)
        U, Ls = O.collapse_u_mean_after_kernel_precalculation(Linv, tt(Xc), tt(prob.X), tt(prob.Z), kern, Q)
    else:
        U, Ls = tt(prob.U), None
    noise = gold["f2/rollout_%s/noise" % tag]
    xs, vs = [], []
    for i in range(noise.shape[0]):
        x, v = O.rollout(Linv, tt(prob.X[-1]), tt(ctrl_full), tt(prob.Z), kern, U, Ls, Q, tt(noise[i]))
        xs.append(x.numpy()); vs.append(v.numpy())
    xs, vs = np.stack(xs), np.stack(vs)
    predict_y = (np.mean(np.einsum("ijk,kl->ijl", xs, prob.C), axis=0) + prob.d[None, :]).reshape(-1)
    predict_y_var = (np.mean(np.einsum("ijk,kl->ijl", vs, prob.C ** 2), axis=0)).reshape(-1) + np.exp(2 * prob.logR).reshape(-1)
    assert_close(gold["f2/rollout_%s/predict_y" % tag], predict_y, 1e-9, "predict_y")
    assert_close(gold["f2/rollout_%s/predict_y_var" % tag].reshape(-1), predict_y_var, 1e-9, "predict_y_var")
    rmse = np.sqrt(np.mean((gold["f2/Y_test"][:30].reshape(-1) - predict_y[:30]) ** 2)) * 1.7
    assert abs(rmse - float(gold["f2/rollout_%s/RMSE" % tag])) <= 1e-9 * rmse
    fit_y = (prob.X[1:] @ prob.C + prob.d).reshape(-1)
    assert_close(gold["f2/rollout_%s/fit_y" % tag], fit_y, 1e-12, "fit_y")
    want = {"y_train_vfe", "y_test_vfe", "v_test_vfe_var", "Y_test_data", "Y_train_data", "Y_train_std", "CC_val", "DD_val", "log_R_cholesky",
            "log_QQ", "Z_val", "U_val", "X_val", "k_lengthscales", "k_log_variances", "case", "ll_seq", "running_time_seq", "PG_num",
            "mc_posterior_samples"}
    assert set(str(k) for k in gold["f2/rollout_%s/file_keys" % tag]) == want


def fixtures_ctrl_full(ds, gold, prob):
    """The 30 control inputs after the training half, as `collect_samples_formal` reads them (control_inputs[test_i + T]):
    the packed fixtures keep only the training half, the golden file carries the rest."""
    return gold["f2/ctrl_future"]


def test_particle_gibbs_vs_reference_source(gold, packed):
    prob = packed["gas_furnace/0"]
    np.testing.assert_allclose(gold["f4/X_before"], prob.X, rtol=0, atol=0)
    P = int(gold["f4/P"])
    ref = O.pg_for_x(tt(prob.X), tt(prob.Y), tt(prob.ctrl), tt(prob.Z), _kern(prob), tt(prob.U), tt(np.exp(prob.logQ)), tt(prob.C), tt(prob.d),
                     tt(np.exp(prob.logR)), P, tt(gold["f4/normals"]), tt(gold["f4/eps"]), tt(gold["f4/uniforms"])).numpy()
    assert_close(gold["f4/X_after"], ref, 1e-9, "PG trajectory")
    assert np.any(gold["f4/X_after"] != prob.X)          # the sweep did move the trajectory


@pytest.mark.parametrize("case_val", (2, 4))
def test_outer_loop_vs_reference_source(gold, packed, case_val):
    """The oracle-driven outer loop vs the reference's own update expressions re-executed per session.run."""
    prob = packed["actuator/0"]
    key = "f3/case%d" % case_val
    iters = int(gold[key + "/iters"])
    assert int(gold[key + "/nvars"]) == (2 * prob.X.shape[1] + 1 if case_val == 2 else 0)
    nf = loop.reference_noise_fn(gold, key, prob.X.shape[1]) if case_val == 2 else None
    res = loop.outer_loop(prob, case_val, iters, noise_fn=nf, window_index=[int(i) for i in gold[key + "/window_index"]])
    for k, v in res["params"].items():
        assert_close(gold["%s/%s" % (key, k)], v, 2e-9, "%s %s" % (key, k))
    assert abs(res["nll"] - float(gold[key + "/nll_final"])) <= 1e-9 * abs(res["nll"])
    # the loop moved every trainable parameter (this is a trajectory, not a fixed point)
    for k in loop.trainable_names(case_val):
        assert np.max(np.abs(res["params"][k] - getattr(prob, k))) > 0


def test_conditional_option_branches_vs_reference_source(gold):
    """base_conditional's full_cov / q_sqrt (2-d, 3-d; whitened or not) / return_Lm branches (conditionals.py:27-66)."""
    g = gold
    Z, Xn, f, q3, q2 = (tt(g["cond/" + k]) for k in ("Z", "Xnew", "f", "q3", "q2"))
    se = O.SquaredExponential(Z.shape[1], variance=0.7, lengthscales=g["cond/ls"], ARD=True)
    for tag, kw in (("full", dict(full_cov=True, white=True)), ("full_q3", dict(full_cov=True, white=True, q_sqrt=q3)),
                    ("full_q2", dict(full_cov=True, white=True, q_sqrt=q2)), ("nonwhite_q3", dict(white=False, q_sqrt=q3)),
                    ("nonwhite_q2", dict(white=False, q_sqrt=q2)), ("full_nonwhite_q3", dict(full_cov=True, white=False, q_sqrt=q3))):
        mu, var = O.conditional(Xn, Z, se, f, **kw)
        assert_close(g["cond/%s/mean" % tag], mu.numpy(), 1e-10, tag)
        assert_close(g["cond/%s/var" % tag], var.numpy(), 1e-10, tag)
    mu, var, Lm = O.conditional(Xn, Z, se, f, white=True, return_Lm=True)
    assert_close(g["cond/return_Lm/Lm"], Lm.numpy(), 1e-12, "Lm")
    se2 = O.SquaredExponential(Z.shape[1], variance=0.3, lengthscales=g["cond/ls"][::-1].copy(), ARD=True)
    mu, var = O.conditional_multi_output(Xn, Z, [se, se2], f[:, :2], white=True, full_cov=True)
    assert tuple(var.shape) == tuple(g["cond/multi_full/var"].shape) == (Xn.shape[0], 1, 2)
    assert_close(g["cond/multi_full/var"], var.numpy(), 1e-10, "multi full_cov")
