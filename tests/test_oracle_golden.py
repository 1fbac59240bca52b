"""CPU tests: the oracle against the golden vectors produced by the reference's own source
(tests/golden/make_reference_golden.py), plus cross-checks that do not depend on either."""
import copy

import numpy as np
import pytest

from oracle import ffvd_oracle as O
from oracle import fixtures, handgrad
from util import assert_close, load_golden

DATASETS = ("dryer", "drive", "gas_furnace", "actuator", "flutter", "ballbeam")
GRADS = ("g_X", "g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR")


@pytest.fixture(scope="module")
def packed():
    p = fixtures.load_packed()
    return {q.name: q for q in p["problems"]}, p["extra"]


@pytest.fixture(scope="module")
def golden():
    return load_golden()


def test_fixture_inventory(packed):
    byname, extra = packed
    assert len(byname) == 95                      # 16 per dataset, 15 for drive
    assert sum(n.startswith("drive/") for n in byname) == 15
    T = {ds: byname[ds + "/0"].Y.shape[0] for ds in DATASETS}
    assert T == dict(dryer=500, drive=250, gas_furnace=148, actuator=512, flutter=512, ballbeam=500)
    for p in byname.values():
        assert p.Z.shape == (100, 5) and p.U.shape == (100, 4) and p.X.shape[1] == 4


@pytest.mark.parametrize("ds", DATASETS)
@pytest.mark.parametrize("idx", (0, 1))
@pytest.mark.parametrize("mode", ("collapsed", "uncollapsed"))
def test_oracle_matches_reference_source(packed, golden, ds, idx, mode):
    """nll, its six terms and every gradient: oracle (restatement) == reference source on the TF shim."""
    prob = packed[0]["%s/%d" % (ds, idx)]
    res = O.nll_and_grads(prob, collapsed=(mode == "collapsed"))
    key = "%s/%d/%s/" % (ds, idx, mode)
    assert abs(res["nll"] - float(golden[key + "nll"])) <= 1e-12 * abs(float(golden[key + "nll"]))
    assert_close(golden[key + "terms"], res["terms"], 1e-12, key + "terms")
    for g in GRADS:
        ref = golden[key + g]
        if mode == "collapsed" and g == "g_U":
            assert np.all(ref == 0)               # U is not in the collapsed graph
        assert_close(ref.reshape(res[g].shape), res[g], 1e-11, key + g)


@pytest.mark.parametrize("case", (2, 7))
def test_sghmc_oracle_matches_reference_source(golden, case):
    key = "sghmc/case%d" % case
    for i in range(int(golden[key + "/nvars"])):
        pre = "%s/var%d/" % (key, i)
        args = [golden[pre + n] for n in ("theta", "grad", "noise", "xi", "g", "g2", "p")]
        th, xi, g, g2, p = O.sghmc_update(*args, epsilon=0.01, mdecay=0.05, X_N=float(golden[key + "/X_N"]), burn_in=True)
        for name, val in (("theta_t", th), ("xi_t", xi), ("g_t", g), ("g2_t", g2), ("p_t", p)):
            assert_close(golden[pre + name], val, 1e-14, pre + name)
        th2, xi2, g2b, g22, p2 = O.sghmc_update(*args, epsilon=0.01, mdecay=0.05, X_N=float(golden[key + "/X_N"]), burn_in=False)
        assert_close(golden[pre + "theta_t"], th2, 1e-14)      # sample_op moves theta and p only
        assert np.array_equal(xi2, golden[pre + "xi"]) and np.array_equal(g22, golden[pre + "g2"])


def test_operator_goldens(golden):
    import torch
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)
    Z, Xn, f = (golden["op/linear_single/" + k] for k in ("Z", "Xnew", "f"))
    lk = O.LinearK(3, 1.0)
    mu, var = O.conditional(t(Xn), t(Z), lk, t(f), white=True)
    assert_close(golden["op/linear_single/mean"], mu.numpy(), 1e-12)
    assert np.max(np.abs(golden["op/linear_single/var"] - var.numpy())) <= 1e-12 * np.max(np.abs(golden["op/linear/Kdiag"]))
    mu, var = O.conditional(t(Xn), t(Z), lk, t(f), white=False)
    assert_close(golden["op/linear_single/mean_nonwhite"], mu.numpy(), 1e-9)
    mu, var = O.conditional_multi_output(t(Xn), t(Z), [lk, lk], t(f), white=True)
    assert_close(golden["op/linear_multi/mean"], mu.numpy(), 1e-12)
    assert_close(golden["op/linear/K"], lk.K(t(Xn), t(Z)).numpy(), 1e-14)
    assert_close(golden["op/linear/Kdiag"], lk.Kdiag(t(Xn)).numpy(), 1e-14)
    se = O.SquaredExponential(3, 0.37, np.array([0.9, 1.7, 2.6]))
    se2 = O.SquaredExponential(3, 0.11, np.array([1.3, 0.8, 3.1]))
    assert_close(golden["op/se/K"], se.K(t(Xn), t(Z)).numpy(), 1e-14)
    assert_close(golden["op/se/Kzz"], se.K(t(Z)).numpy(), 1e-14)
    assert_close(golden["op/se/Kdiag"], se.Kdiag(t(Xn)).numpy(), 1e-15)
    mu, var = O.conditional_multi_output(t(Xn), t(Z), [se, se2], t(f), white=True)
    assert_close(golden["op/se_multi/mean"], mu.numpy(), 1e-12)
    assert_close(golden["op/se_multi/var"], var.numpy(), 1e-12)
    Li = O.kernel_pre_cal(t(Z), [se, se2])
    assert_close(golden["op/se_multi/LinvT"], np.stack([a.numpy() for a in Li]), 1e-12)
    y, ym, R = golden["op/ld/y"], golden["op/ld/ymean"], golden["op/ld/R"]
    assert_close(golden["op/ld/diag"], O.logdensity_norm_diag(t(y), t(ym), t(R)).numpy(), 1e-14)
    assert_close(golden["op/ld/diag_nonvec"], O.logdensity_norm_diag_nonvec(t(y), t(ym), t(R)).numpy(), 1e-14)
    assert_close(golden["op/ld/full"], O.logdensity_norm(t(y), t(ym), t(golden["op/ld/Rfull"])).numpy(), 1e-14)


@pytest.mark.parametrize("collapsed", (False, True))
@pytest.mark.parametrize("kind", (0, 1))
def test_hand_derived_gradients_match_autograd(packed, collapsed, kind):
    """The algorithm the CUDA kernels implement (oracle/handgrad.py) == reverse-mode AD of the graph."""
    prob = copy.copy(packed[0]["gas_furnace/3"])
    prob.X = packed[1]["gas_furnace/0"][:2] if kind == 0 else prob.X
    if kind == 1:
        prob.kind, prob.logl, prob.logv = 1, None, np.zeros(4)
    a = O.nll_and_grads(prob, collapsed=collapsed)
    b = handgrad.nll_and_grads(prob, collapsed=collapsed)
    for k in a:
        assert_close(a[k], b[k], 2e-10, k)


def test_oracle_gradients_finite_differences():
    prob = fixtures.synthetic_problem(T=40, M=12, D=2, S=1)
    rng = np.random.default_rng(1)
    for collapsed in (False, True):
        base = O.nll_and_grads(prob, collapsed=collapsed)
        for name in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR"):
            if collapsed and name == "U":
                continue
            arr = getattr(prob, name)
            direction = rng.standard_normal(arr.shape)
            h = 1e-6
            vals = []
            for sgn in (+1, -1):
                q = copy.copy(prob)
                setattr(q, name, arr + sgn * h * direction)
                vals.append(O.nll_and_grads(q, collapsed=collapsed)["nll"])
            fd = (vals[0] - vals[1]) / (2 * h)
            an = float(np.sum(base["g_" + name] * direction))
            assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)), (collapsed, name, fd, an)


def test_collapsed_equals_uncollapsed_at_optimal_u(packed):
    """nll_collapsed == nll_uncollapsed(U*) + 1/2 sum_d logdet(H_d)/T  and  d nll_unc / dU = 0 at U*
     -- an identity between the two graphs that needs neither golden nor CUDA."""
    import torch
    prob = copy.copy(packed[0]["actuator/0"])
    T = prob.Y.shape[0]
    p = {k: torch.as_tensor(getattr(prob, k), dtype=torch.float64) for k in O.PARAM_NAMES}
    kerns = O._make_kernels(p["logv"], p["logl"], 0, 5)
    Q = torch.exp(p["logQ"])
    Xc = torch.cat((p["X"][:T], torch.as_tensor(prob.ctrl)), dim=1)
    Linv = O.kernel_pre_cal(p["Z"], kerns)
    Ustar, LHinv = O.collapse_u_mean_after_kernel_precalculation(Linv, Xc, p["X"], p["Z"], kerns, Q)
    logdet = sum(-2.0 * torch.sum(torch.log(torch.diagonal(LHinv[d]))) for d in range(4))
    q = copy.copy(prob)
    q.U = Ustar.numpy()
    unc = O.nll_and_grads(q, collapsed=False)
    col = O.nll_and_grads(prob, collapsed=True)
    # the uncollapsed objective includes prior_U = -1/2 |U|^2, which is exactly the N(0,I) factor
    # that the collapsed bound integrates against, so the identity holds with the prior in place
    lhs = col["nll"]
    rhs = unc["nll"] + 0.5 * float(logdet) / T
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
    assert np.max(np.abs(unc["g_U"])) <= 1e-9      # U* is the stationary point


def test_adam_matches_closed_form():
    rng = np.random.default_rng(0)
    th, gr = rng.standard_normal(7), rng.standard_normal(7)
    t1, m1, v1 = O.adam_update(th, gr, np.zeros(7), np.zeros(7), step=1, lr=0.003)
    # first TF1 Adam step moves every coordinate by lr * sign(grad) (up to epsilon)
    assert np.allclose(t1, th - 0.003 * np.sign(gr), rtol=0, atol=1e-6)
    assert O.adam_learning_rate(1000) == pytest.approx(0.003 * 0.95)
