"""GPU parity tests: the CUDA path, called through the C ABI (ctypes + DLPack), against the CPU
oracle and against the committed golden vectors.  Tolerance (BASELINE north star): <= 1e-9
relative in float64, measured relative to each tensor's max-norm."""
import copy

import numpy as np
import pytest

from util import assert_close, load_golden, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-9
PKEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
GKEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")


@pytest.fixture(scope="module")
def env():
    import torch
    import ffvd_b200
    from oracle import fixtures
    dev = torch.device("cuda:0")
    ctx = ffvd_b200.Context(0, torch.cuda.current_stream(0).cuda_stream)
    packed = fixtures.load_packed()
    return dict(torch=torch, ffvd=ffvd_b200, dev=dev, ctx=ctx, problems=packed["problems"],
                byname={p.name: p for p in packed["problems"]}, extra=packed["extra"])


def dev_problem(env, prob):
    torch = env["torch"]
    return {k: (None if getattr(prob, k) is None else torch.as_tensor(np.ascontiguousarray(getattr(prob, k)),
                                                                      dtype=torch.float64, device=env["dev"]))
            for k in PKEYS}


def alloc_out(env, p):
    torch = env["torch"]
    S = 1 if p["X"].dim() == 2 else p["X"].shape[0]
    o = {"nll": torch.full((S,), float("nan"), dtype=torch.float64, device=env["dev"]),
         "terms": torch.full((S, 6), float("nan"), dtype=torch.float64, device=env["dev"])}
    for k in GKEYS:
        if p[k] is not None:
            o["g_" + k] = torch.full_like(p[k], float("nan"))
    return o


def run_cuda(env, prob, collapsed, flags=None):
    p = dev_problem(env, prob)
    o = alloc_out(env, p)
    flags = env["ffvd"].FLAG_PRIOR_Z_NORMAL if flags is None else flags
    env["ctx"].nll_grads(prob.kind, collapsed, p, o, flags=flags)
    env["torch"].cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in o.items()}
    if p["X"].dim() == 2:
        res["nll"], res["terms"] = res["nll"][0], res["terms"][0]
    return res


def check(ref, got, tol=TOL, what=""):
    for k in ref:
        if k in got:
            assert_close(ref[k], got[k], tol, "%s %s" % (what, k))


# ------------------------------------------------------------------------------------------------
def test_all_95_fixtures_vs_oracle(env):
    """Every bundled warm start x {uncollapsed, collapsed}: nll, six terms and all gradients."""
    from oracle import ffvd_oracle as O
    worst = 0.0
    for prob in env["problems"]:
        for collapsed in (False, True):
            ref = O.nll_and_grads(prob, collapsed=collapsed)
            got = run_cuda(env, prob, collapsed)
            for k in ref:
                e = relerr(ref[k], got[k])
                worst = max(worst, e)
                assert e <= TOL, (prob.name, collapsed, k, e)
    print("worst max-norm relative error over 95 x 2 evaluations: %.3e" % worst)


@pytest.mark.parametrize("ds", ("dryer", "drive", "gas_furnace", "actuator", "flutter", "ballbeam"))
@pytest.mark.parametrize("mode", ("collapsed", "uncollapsed"))
def test_vs_reference_source_golden(env, ds, mode):
    """CUDA vs the vectors the reference's own source produced (tests/golden/reference_shim_golden.npz)."""
    golden = load_golden()
    for idx in (0, 1):
        got = run_cuda(env, env["byname"]["%s/%d" % (ds, idx)], mode == "collapsed")
        key = "%s/%d/%s/" % (ds, idx, mode)
        assert abs(got["nll"] - float(golden[key + "nll"])) <= TOL * abs(float(golden[key + "nll"]))
        assert_close(golden[key + "terms"], got["terms"], TOL, key + "terms")
        for g in GKEYS:
            assert_close(golden[key + "g_" + g].reshape(got["g_" + g].shape), got["g_" + g], TOL, key + g)


@pytest.mark.parametrize("collapsed", (False, True))
def test_sample_axis(env, collapsed):
    """S trajectories sharing Z,U,hypers: per-sample nll/terms/x-bar equal single-sample runs and
    shared gradients are the sum over samples (SURVEY 8 batch-axis contract)."""
    from oracle import ffvd_oracle as O
    prob = copy.copy(env["byname"]["flutter/0"])
    prob.X = env["extra"]["flutter/0"]
    ref = O.nll_and_grads(prob, collapsed=collapsed)
    got = run_cuda(env, prob, collapsed)
    check(ref, got, what="S=4")
    singles = []
    for s in range(prob.X.shape[0]):
        q = copy.copy(prob); q.X = prob.X[s]
        singles.append(run_cuda(env, q, collapsed))
    assert_close(np.array([r["nll"] for r in singles]), got["nll"], 1e-12)
    assert_close(np.stack([r["g_X"] for r in singles]), got["g_X"], 1e-11)
    assert_close(sum(r["g_Z"] for r in singles), got["g_Z"], 1e-10)
    # FFVD_FLAG_PRIOR_ONCE: shared priors counted once
    once = run_cuda(env, prob, collapsed, flags=env["ffvd"].FLAG_PRIOR_Z_NORMAL | env["ffvd"].FLAG_PRIOR_ONCE)
    T = prob.Y.shape[0]
    assert_close(got["g_Z"] - 3 * prob.Z / T, once["g_Z"], 1e-10)


@pytest.mark.parametrize("collapsed", (False, True))
def test_linear_kernel(env, collapsed):
    from oracle import ffvd_oracle as O
    prob = copy.copy(env["byname"]["gas_furnace/0"])
    prob.kind, prob.logl, prob.logv = 1, None, np.log(np.array([1.0, 0.7, 1.3, 0.9]))
    check(O.nll_and_grads(prob, collapsed=collapsed), run_cuda(env, prob, collapsed), what="LinearK")


@pytest.mark.parametrize("T,M,D,S", [(1, 5, 1, 1), (7, 3, 2, 2), (64, 128, 2, 1), (65, 129, 2, 1), (130, 200, 3, 2),
                                      (97, 300, 2, 1), (80, 512, 2, 1), (40, 700, 2, 1), (24, 1100, 1, 2), (12, 2048, 1, 1)])
@pytest.mark.parametrize("collapsed", (False, True))
def test_ragged_shapes(env, T, M, D, S, collapsed):
    """Tile-ragged T (1, prime, multiple and multiple+1 of the tile), every padded-M template."""
    from oracle import fixtures, ffvd_oracle as O
    prob = fixtures.synthetic_problem(T=T, M=M, D=D, S=S, seed=T * 1000 + M)
    check(O.nll_and_grads(prob, collapsed=collapsed), run_cuda(env, prob, collapsed), what="T%d M%d" % (T, M))


@pytest.mark.parametrize("M", (1, 2, 4, 7, 8, 9, 31, 32, 33, 64, 97, 116, 118, 119, 120, 128, 129, 140, 152, 153, 168))
@pytest.mark.parametrize("collapsed", (False, True))
def test_small_m_factorisation_paths(env, M, collapsed):
    """Every branch of the K(Z,Z) / H factorisation: the register-resident single-CTA path (M <= 119: column form below
    M = 8, panel form above, one tile, partial tiles, one warp) and the blocked multi-kernel path above it (one, two
    and three 64-blocks, ragged last block)."""
    from oracle import fixtures, ffvd_oracle as O
    prob = fixtures.synthetic_problem(T=37, M=M, D=2, S=2, seed=4242 + M)
    check(O.nll_and_grads(prob, collapsed=collapsed), run_cuda(env, prob, collapsed), what="M%d" % M)


@pytest.mark.parametrize("M", (6, 100, 140, 300))
def test_cholesky_failure_is_reported(env, M):
    """A negative "jitter" larger than the kernel variance makes the first pivot of chol(K(Z,Z)) negative (an exact
    zero pivot from duplicated inducing points is rounding dependent, here as in LAPACK).  The reference's
    tf.linalg.cholesky raises; every factorisation path here must return the not-positive-definite status, not NaNs."""
    from oracle import fixtures
    prob = fixtures.synthetic_problem(T=20, M=M, D=2, S=1, seed=11)
    p = dev_problem(env, prob)
    o = alloc_out(env, p)
    with pytest.raises(env["ffvd"].NotPositiveDefinite) as ei:
        env["ctx"].nll_grads(prob.kind, False, p, o, flags=env["ffvd"].FLAG_PRIOR_Z_NORMAL, jitter=-1.0e3)
    assert "positive definite" in str(ei.value) and ei.value.pivot == 1
    # the context stays usable
    good = fixtures.synthetic_problem(T=20, M=M, D=2, S=1, seed=12)
    from oracle import ffvd_oracle as O
    check(O.nll_and_grads(good, collapsed=False), run_cuda(env, good, False), what="after failure M%d" % M)


@pytest.mark.parametrize("T,M,D,n_ctrl", [(70, 40, 16, 1), (90, 200, 12, 3), (33, 130, 14, 1), (40, 60, 20, 11), (50, 500, 16, 1)])
@pytest.mark.parametrize("collapsed,kind", ((False, 0), (True, 0), (False, 1)))
def test_wide_inputs(env, T, M, D, n_ctrl, collapsed, kind):
    """Din = 15 / 16 (the boundary between two and four 8-column blocks of [X,1] / [Z,1] in the W back-propagation), 17
    (the C5 shape: D = 16 + one control) and 31 (the build's limit)."""
    from oracle import fixtures, ffvd_oracle as O
    prob = fixtures.synthetic_problem(T=T, M=M, D=D, S=1, n_ctrl=n_ctrl, seed=7 * T + M, kind=kind)
    check(O.nll_and_grads(prob, collapsed=collapsed), run_cuda(env, prob, collapsed), what="Din%d M%d" % (D + n_ctrl, M))


def test_no_control_inputs_and_multi_output_y(env):
    from oracle import fixtures, ffvd_oracle as O
    prob = fixtures.synthetic_problem(T=50, M=20, D=3, S=1, n_ctrl=0, Dy=2)
    prob.ctrl = np.zeros((50, 0))
    ref = O.nll_and_grads(prob, collapsed=False)
    q = copy.copy(prob); q.ctrl = None
    check(ref, run_cuda(env, q, False), what="nc=0 Dy=2")


def test_batched_chains(env):
    """BASELINE config 4: independent chains with ragged T in one call == one call per chain."""
    ctx, torch = env["ctx"], env["torch"]
    names = ["actuator/3", "gas_furnace/5", "drive/2", "dryer/7", "ballbeam/1", "flutter/9"]
    for collapsed in (False, True):
        probs = [dev_problem(env, env["byname"][n]) for n in names]
        outs = [alloc_out(env, p) for p in probs]
        ctx.nll_grads_batched(0, collapsed, probs, outs)
        torch.cuda.synchronize()
        for n, o in zip(names, outs):
            single = run_cuda(env, env["byname"][n], collapsed)
            for k, v in o.items():
                g = v.cpu().numpy()
                assert_close(single[k], g.reshape(np.shape(single[k])), 1e-10, "%s %s" % (n, k))   # FP64 atomics: summation order differs run to run


def test_forward_only_flag(env):
    prob = env["byname"]["actuator/0"]
    for collapsed in (False, True):
        full = run_cuda(env, prob, collapsed)
        p = dev_problem(env, prob)
        o = {"nll": env["torch"].empty(1, dtype=env["torch"].float64, device=env["dev"]),
             "terms": env["torch"].empty(1, 6, dtype=env["torch"].float64, device=env["dev"])}
        env["ctx"].nll_grads(0, collapsed, p, o, flags=env["ffvd"].FLAG_PRIOR_Z_NORMAL | env["ffvd"].FLAG_NO_GRADS)
        assert_close(full["terms"], o["terms"].cpu().numpy()[0], 1e-12)


def test_host_numpy_tensors_are_staged(env):
    """kDLCPU tensors in and out (explicit staging copies inside the library)."""
    from oracle import ffvd_oracle as O
    prob = env["byname"]["drive/0"]
    p = {k: (None if getattr(prob, k) is None else np.ascontiguousarray(getattr(prob, k), dtype=np.float64)) for k in PKEYS}
    o = {"nll": np.zeros(1), "terms": np.zeros((1, 6))}
    for k in GKEYS:
        o["g_" + k] = np.full_like(p[k], np.nan)
    env["ctx"].nll_grads(0, False, p, o)
    ref = O.nll_and_grads(prob, collapsed=False)
    assert abs(o["nll"][0] - ref["nll"]) <= TOL * abs(ref["nll"])
    for k in GKEYS:
        assert_close(ref["g_" + k], o["g_" + k], TOL, k)


def test_error_statuses(env):
    ffvd, torch, ctx = env["ffvd"], env["torch"], env["ctx"]
    prob = copy.copy(env["byname"]["drive/0"])
    # float16 / integer input -> ValueError (dtype); float32 is the widened float32 mode (test_float32_tensor_mode)
    p = dev_problem(env, prob); o = alloc_out(env, p)
    p["Z"] = p["Z"].to(torch.float16)
    with pytest.raises(ValueError):
        ctx.nll_grads(0, False, p, o)
    p["Z"] = p["Z"].to(torch.int64)
    with pytest.raises(ValueError):
        ctx.nll_grads(0, False, p, o)
    # shape mismatch
    p = dev_problem(env, prob); o = alloc_out(env, p)
    p["U"] = p["U"][:50].contiguous()
    with pytest.raises(ValueError):
        ctx.nll_grads(0, False, p, o)
    # non-contiguous
    p = dev_problem(env, prob); o = alloc_out(env, p)
    p["Z"] = p["Z"].t().contiguous().t()
    with pytest.raises(ValueError):
        ctx.nll_grads(0, False, p, o)
    # non-SPD Kzz (duplicate inducing points, zero jitter) -> NotPositiveDefinite, no abort
    q = copy.copy(prob); q.Z = prob.Z.copy(); q.Z[1] = q.Z[0]
    p = dev_problem(env, q); o = alloc_out(env, p)
    with pytest.raises(ffvd.NotPositiveDefinite):
        ctx.nll_grads(0, False, p, o, jitter=0.0)
    # and the context is still usable afterwards
    run_cuda(env, prob, False)


# ------------------------------------------------------------------------------------------------
def test_operator_level_vs_golden(env):
    """K / Kdiag / conditional / kernel_pre_cal / log-densities through the mirror API, on the golden inputs."""
    from ffvd_b200 import conditionals, conditionals_multi_output, likelihoods
    from ffvd_b200.kernels import LinearK
    from ffvd_b200.kernels_multi_output import SquaredExponential
    torch, dev = env["torch"], env["dev"]
    g = load_golden()
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64, device=dev)
    Z, Xn, f = (g["op/linear_single/" + k] for k in ("Z", "Xnew", "f"))
    lk = LinearK(3, variance=1.0)
    assert_close(g["op/linear/K"], lk.K(t(Xn), t(Z)).cpu().numpy(), 1e-13)
    assert_close(g["op/linear/Kdiag"], lk.Kdiag(t(Xn)).cpu().numpy(), 1e-13)
    kd = np.max(np.abs(g["op/linear/Kdiag"]))
    for white, sfx in ((True, ""), (False, "_nonwhite")):
        mu, var = conditionals.conditional(t(Xn), t(Z), lk, t(f), white=white)
        # rank-deficient Kzz + 1e-7 I: cond ~ 1e9, so the (non-)whitened mean carries ~cond*eps
        assert_close(g["op/linear_single/mean" + sfx], mu.cpu().numpy(), 1e-9 if white else 1e-6)
        # variance tolerance is relative to max|Kdiag| for LinearK
        assert np.max(np.abs(g["op/linear_single/var" + sfx] - var.cpu().numpy())) <= 1e-9 * kd
    mu, var = conditionals_multi_output.conditional(t(Xn), t(Z), [lk, lk], t(f), white=True)
    assert_close(g["op/linear_multi/mean"], mu.cpu().numpy(), 1e-9)
    assert np.max(np.abs(g["op/linear_multi/var"] - var.cpu().numpy())) <= 1e-9 * kd
    se = SquaredExponential(3, variance=0.37, lengthscales=np.array([0.9, 1.7, 2.6]), ARD=True)
    se2 = SquaredExponential(3, variance=0.11, lengthscales=np.array([1.3, 0.8, 3.1]), ARD=True)
    assert_close(g["op/se/K"], se.K(t(Xn), t(Z)).cpu().numpy(), 1e-13)
    assert_close(g["op/se/Kzz"], se.K(t(Z)).cpu().numpy(), 1e-13)
    assert_close(g["op/se/Kdiag"], se.Kdiag(t(Xn)).cpu().numpy(), 1e-15)
    mu, var = conditionals_multi_output.conditional(t(Xn), t(Z), [se, se2], t(f), white=True)
    assert_close(g["op/se_multi/mean"], mu.cpu().numpy(), TOL)
    assert_close(g["op/se_multi/var"], var.cpu().numpy(), TOL)
    Li = conditionals_multi_output.kernel_pre_cal(t(Z), [se, se2])
    assert_close(g["op/se_multi/LinvT"], torch.stack(Li).cpu().numpy(), TOL)
    y, ym, R = g["op/ld/y"], g["op/ld/ymean"], g["op/ld/R"]
    assert_close(g["op/ld/diag"], likelihoods.logdensity_norm_diag(t(y), t(ym), t(R)).cpu().numpy(), 1e-14)
    assert_close(g["op/ld/diag_nonvec"], likelihoods.logdensity_norm_diag_nonvec(t(y), t(ym), t(R)).cpu().numpy(), 1e-14)
    # numpy in -> numpy out
    K_np = se.K(np.asarray(Xn), np.asarray(Z))
    assert isinstance(K_np, np.ndarray)
    assert_close(g["op/se/K"], K_np, 1e-13)


def test_conditional_on_fixture_vs_oracle(env):
    """The regularizer's conditional (dgp_model.py:344) at fixture size, and collapse_after_kernel_precalculation."""
    import torch as th
    from oracle import ffvd_oracle as O
    from ffvd_b200 import conditionals_multi_output as cmo
    from ffvd_b200.kernels_multi_output import SquaredExponential
    prob = env["byname"]["ballbeam/4"]
    T = prob.Y.shape[0]
    kerns = [SquaredExponential(5, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(4)]
    Xc = np.concatenate([prob.X[:T], prob.ctrl], axis=1)
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    mu, var = cmo.conditional(t(Xc), t(prob.Z), kerns, t(prob.U), white=True)
    ok = O._make_kernels(th.as_tensor(prob.logv), th.as_tensor(prob.logl), 0, 5)
    rmu, rvar = O.conditional_multi_output(th.as_tensor(Xc), th.as_tensor(prob.Z), ok, th.as_tensor(prob.U), white=True)
    assert_close(rmu.numpy(), mu.cpu().numpy(), TOL)
    assert_close(rvar.numpy(), var.cpu().numpy(), TOL)
    Q = np.exp(prob.logQ)
    t1, t2, tr = cmo.collapse_after_kernel_precalculation(None, t(Xc), t(prob.X), t(prob.Z), kerns, t(Q), float(T), float(T))
    Linv = O.kernel_pre_cal(th.as_tensor(prob.Z), ok)
    r1, r2, rtr = O.collapse_after_kernel_precalculation(Linv, th.as_tensor(Xc), th.as_tensor(prob.X), th.as_tensor(prob.Z), ok,
                                                         th.as_tensor(Q), float(T), float(T))
    for a, b in ((r1, t1), (r2, t2), (rtr, tr)):
        assert abs(float(a) - float(b)) <= TOL * abs(float(a))


@pytest.mark.parametrize("case", (2, 7))
def test_sghmc_kernel_vs_reference_golden(env, case):
    torch, ctx, dev = env["torch"], env["ctx"], env["dev"]
    g = load_golden()
    key = "sghmc/case%d" % case
    X_N = float(g[key + "/X_N"])
    for i in range(int(g[key + "/nvars"])):
        pre = "%s/var%d/" % (key, i)
        for burn_in in (True, False):
            ten = {n: torch.as_tensor(np.array(g[pre + n], dtype=np.float64), device=dev).clone().reshape(-1)
                   for n in ("theta", "grad", "noise", "xi", "g", "g2", "p")}
            ctx.sghmc_update(ten["theta"], ten["grad"], ten["noise"], ten["xi"], ten["g"], ten["g2"], ten["p"], 0.01, 0.05, X_N, burn_in)
            torch.cuda.synchronize()
            assert_close(g[pre + "theta_t"].reshape(-1), ten["theta"].cpu().numpy(), 1e-14)
            assert_close(g[pre + "p_t"].reshape(-1), ten["p"].cpu().numpy(), 1e-13)
            for n in ("xi", "g", "g2"):
                expect = g[pre + n + "_t"] if burn_in else g[pre + n]
                assert_close(expect.reshape(-1), ten[n].cpu().numpy(), 1e-14)


def test_adam_kernel_vs_oracle(env):
    from oracle import ffvd_oracle as O
    torch, ctx, dev = env["torch"], env["ctx"], env["dev"]
    rng = np.random.default_rng(3)
    th, m, v = rng.standard_normal(1001), np.zeros(1001), np.zeros(1001)
    d_th, d_m, d_v = (torch.as_tensor(a.copy(), device=dev) for a in (th, m, v))
    for step in range(1, 5):
        gr = rng.standard_normal(1001)
        lr = O.adam_learning_rate(step)
        th, m, v = O.adam_update(th, gr, m, v, step=step, lr=lr)
        ctx.adam_update(d_th, torch.as_tensor(gr, device=dev), d_m, d_v, lr, step=step)
    torch.cuda.synchronize()
    assert_close(th, d_th.cpu().numpy(), 1e-14)
    assert_close(v, d_v.cpu().numpy(), 1e-14)


def test_dgpssm_mirror_and_sghmc_chain(env):
    """The `vfegpssm`-shaped API end to end: DGPSSM(...).nll, the SG-HMC variable set of case 2, and
    three scheduled updates with injected noise against the oracle driven by the same noise."""
    from oracle import ffvd_oracle as O
    from ffvd_b200.dgp_model import DGPSSM
    from ffvd_b200.kernels_multi_output import SquaredExponential
    from ffvd_b200.likelihoods import Gaussian
    torch = env["torch"]
    prob = copy.copy(env["byname"]["gas_furnace/2"])
    T = prob.Y.shape[0]
    kerns = [[SquaredExponential(5, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(4)]]
    lik = Gaussian(1, 4, CC=prob.C, DD=prob.d, RR_chol=np.exp(prob.logR))
    model = DGPSSM(prob.Y, [4], 100, kerns, lik, minibatch_size=T, window_size=64, prior_type="normal",
                   QQ_chol=np.exp(0.5 * prob.logQ), ZZ=prob.Z, control_inputs=prob.ctrl, U_ini=prob.U, X_0_ini=prob.X[0],
                   X_train_ini=prob.X[1:], kernel_optimization=False, U_optimization=False, U_collapse=False,
                   Z_optimization=True, case_val=2)
    assert model.vars == ["logv", "logl", "U"] and model.X_N == T + 1
    assert set(model.trainable) == {"X", "Z", "logQ", "C", "d", "logR"}
    ref = O.nll_and_grads(prob, collapsed=False)
    assert abs(float(model.nll) - ref["nll"]) <= TOL * abs(ref["nll"])
    # chain: burn-in, burn-in, sample with injected noise
    rng = np.random.default_rng(11)
    state = {n: dict(xi=np.ones_like(getattr(prob, n)), g=np.ones_like(getattr(prob, n)), g2=np.ones_like(getattr(prob, n)),
                     p=np.zeros_like(getattr(prob, n))) for n in model.vars}
    for burn_in in (True, True, False):
        noise = {n: rng.standard_normal(getattr(prob, n).shape) for n in model.vars}
        model._run_update(burn_in, noise)
        grads = O.nll_and_grads(prob, collapsed=False)
        for n in model.vars:
            s = state[n]
            th, s["xi"], s["g"], s["g2"], s["p"] = O.sghmc_update(getattr(prob, n), grads["g_" + n], noise[n], s["xi"], s["g"],
                                                                    s["g2"], s["p"], epsilon=0.01, mdecay=0.05, X_N=T + 1, burn_in=burn_in)
            setattr(prob, n, th)
    torch.cuda.synchronize()
    for n in model.vars:
        assert_close(getattr(prob, n), model.params[n].cpu().numpy(), 1e-9, n)
    out = model.train_hypers()
    assert torch.isfinite(out["nll"]).all()


def test_full_size_properties(env):
    """At a BASELINE-config-3-shaped size that the oracle cannot finish (M=256, D=8, T=4096, S=4):
    S-batched == per-sample, x-bar rows of a prefix are unaffected by dropping the tail's emission... 
    and linearity of the shared gradients over samples."""
    from oracle import fixtures
    prob = fixtures.synthetic_problem(T=4096, M=256, D=8, S=4, seed=5)
    got = run_cuda(env, prob, False)
    assert np.all(np.isfinite(got["nll"]))
    acc = None
    for s in range(4):
        q = copy.copy(prob); q.X = prob.X[s]
        r = run_cuda(env, q, False)
        assert abs(r["nll"] - got["nll"][s]) <= 1e-11 * abs(r["nll"])
        assert_close(r["g_X"], got["g_X"][s], 1e-10)
        acc = r if acc is None else {k: acc[k] + r[k] for k in r}
    for k in ("g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR"):
        assert_close(acc[k], got[k], 1e-9, k)
    # oracle spot check on a short prefix of sample 0 (same Z,U,hypers)
    from oracle import ffvd_oracle as O
    q = copy.copy(prob); q.X = prob.X[0, :129]; q.Y = prob.Y[:128]; q.ctrl = prob.ctrl[:128]
    check(O.nll_and_grads(q, collapsed=False), run_cuda(env, q, False), what="prefix")


# ---- "next" row, SURVEY 8(f) rank 2: collapsed q(u) and the prediction-time conditional ---------------------------
@pytest.mark.parametrize("name", ("ballbeam/4", "actuator/0", "gas_furnace/0"))
def test_collapse_u_mean_and_predictive_conditional(env, name):
    """collapse_u_mean_after_kernel_precalculation (cmo:206-227) and conditional_after_kernel_precalculation with its
    q_sqrt term (cmo:306-387) against the op-for-op oracle, including the reference's first-output broadcasting of
    q_sqrt."""
    import torch as th
    from oracle import ffvd_oracle as O
    from ffvd_b200 import conditionals_multi_output as cmo
    from ffvd_b200.kernels_multi_output import SquaredExponential
    prob = env["byname"][name]
    T, D = prob.Y.shape[0], prob.X.shape[1]
    Din = prob.Z.shape[1]
    kerns = [SquaredExponential(Din, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(D)]
    ok = O._make_kernels(th.as_tensor(prob.logv), th.as_tensor(prob.logl), 0, Din)
    Xc = np.concatenate([prob.X[:T], prob.ctrl], axis=1)
    Q = np.exp(prob.logQ)
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    Linv_o = O.kernel_pre_cal(th.as_tensor(prob.Z), ok)
    rU, rL = O.collapse_u_mean_after_kernel_precalculation(Linv_o, th.as_tensor(Xc), th.as_tensor(prob.X), th.as_tensor(prob.Z), ok,
                                                           th.as_tensor(Q))
    U_mean, Lseq = cmo.collapse_u_mean_after_kernel_precalculation(None, t(Xc), t(prob.X), t(prob.Z), kerns, t(Q))
    assert tuple(U_mean.shape) == (1, prob.Z.shape[0], D)             # tf.transpose of the (D,M,1) stack
    assert_close(rU.numpy(), U_mean[0].cpu().numpy(), TOL)
    assert_close(rL.numpy(), Lseq.cpu().numpy(), TOL)
    # prediction-time conditional at a few new inputs, q(u) = N(U_mean, H^{-1})
    rng = np.random.default_rng(3)
    Xn = Xc[rng.integers(0, T, 37)] + 0.05 * rng.standard_normal((37, Din))
    rmu, rvar = O.conditional_after_kernel_precalculation(Linv_o, th.as_tensor(Xn), th.as_tensor(prob.Z), ok, rU, white=True, q_sqrt=rL)
    mu, var = cmo.conditional_after_kernel_precalculation(None, t(Xn), t(prob.Z), kerns, U_mean[0], white=True, q_sqrt=Lseq)
    assert_close(rmu.numpy(), mu.cpu().numpy(), TOL)
    # the variance is a difference of O(v) terms: tolerance relative to max Kdiag (as for the LinearK variances)
    scale = float(np.max(np.exp(prob.logv)))
    assert np.max(np.abs(rvar.numpy() - var.cpu().numpy())) <= TOL * scale
    # NumPy (host) tensors go through the same C ABI with explicit staging copies
    mu_h, var_h = cmo.conditional_after_kernel_precalculation(None, Xn, prob.Z, kerns, U_mean[0].cpu().numpy(), white=True,
                                                              q_sqrt=Lseq.cpu().numpy())
    assert_close(mu.cpu().numpy(), mu_h, 1e-13)
    assert np.max(np.abs(var.cpu().numpy() - var_h)) <= 1e-13 * scale
    with pytest.raises(NotImplementedError):
        cmo.conditional_after_kernel_precalculation(None, t(Xn), t(prob.Z), kerns, U_mean[0], white=False)


def test_single_kernel_conditional_q_sqrt(env):
    """conditionals.conditional (one kernel, f (M,R)) with q_sqrt given as (R,M,M) factors and as (M,R) scales
    (conditionals.py:46-58): one factor per column of f."""
    import torch as th
    from oracle import ffvd_oracle as O
    from ffvd_b200 import conditionals
    from ffvd_b200.kernels_multi_output import SquaredExponential
    rng = np.random.default_rng(11)
    M, R, N, Din = 60, 3, 41, 4
    Z = rng.standard_normal((M, Din)) * 1.5
    Xn = rng.standard_normal((N, Din))
    f = rng.standard_normal((M, R))
    q3 = np.tril(rng.standard_normal((R, M, M))) * 0.2
    q2 = rng.uniform(0.1, 1.0, (M, R))
    ls = rng.uniform(1.0, 3.0, Din)
    kern = SquaredExponential(Din, variance=0.7, lengthscales=ls, ARD=True)
    ok = O.SquaredExponential(Din, variance=0.7, lengthscales=ls, ARD=True)
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    for q in (q3, q2):
        rmu, rvar = O.conditional(th.as_tensor(Xn), th.as_tensor(Z), ok, th.as_tensor(f), white=True, q_sqrt=th.as_tensor(q))
        mu, var = conditionals.conditional(t(Xn), t(Z), kern, t(f), white=True, q_sqrt=t(q))
        assert_close(rmu.numpy(), mu.cpu().numpy(), TOL)
        assert_close(rvar.numpy(), var.cpu().numpy(), TOL)
    with pytest.raises(ValueError):
        conditionals.conditional(t(Xn), t(Z), kern, t(f), white=True, q_sqrt=t(q2[:, 0]))


# ---- "next" rows, SURVEY 8(f) ranks 2-3: posterior roll-out, results file and the outer training loop ----------------
def _model_from_problem(env, prob, case_val, iterations=1, extra_ctrl=None):
    from ffvd_b200 import models
    ctrl = prob.ctrl if extra_ctrl is None else np.concatenate([prob.ctrl, extra_ctrl], axis=0)
    args = dict(CC=prob.C, DD=prob.d, QQ_chol=np.exp(0.5 * prob.logQ), RR_chol=np.exp(prob.logR), lengthscales=np.exp(prob.logl),
                variance=np.exp(prob.logv), UU_ini=prob.U, XX_0_ini=prob.X[0], x_initialization=prob.X[1:], ZZ=prob.Z)
    m = models.configure(models.RegressionModel("normal"), args, ctrl, case_val, iterations=iterations, window_size=64)
    return m, ctrl


@pytest.mark.parametrize("case_val", (4, 2))
def test_rollout_and_results_file(env, case_val, tmp_path):
    """collect_samples_formal (base_model.py:197-522) against the oracle's restatement of the roll-out with the same
    injected noise: case 4 (collapsed q(u), all samples share parameters -> batched roll-out) and case 2 (U, kernel
    hypers SG-HMC sampled -> per-sample roll-out after `spacing` sample updates)."""
    import torch as th
    from oracle import ffvd_oracle as O
    prob = env["byname"]["drive/0"]
    T, D = prob.Y.shape[0], prob.X.shape[1]
    rng = np.random.default_rng(5)
    L, num = 12, 3
    future = rng.standard_normal((L + 2, prob.ctrl.shape[1]))
    m, ctrl = _model_from_problem(env, prob, case_val, iterations=0, extra_ctrl=future)
    model = m.fit(prob.Y, kernel_type="SquaredExponential")
    noise = rng.standard_normal((num, L, D))
    spacing = 0 if case_val == 2 else 5          # spacing 0: the sampled variables stay at their initial values
    base = str(tmp_path / "run")
    py, pv = model.collect_samples_formal(num, spacing, ctrl, L, sghmc_var_len=len(model.vars), U_collapse=(case_val == 4),
                                          Y_test=rng.standard_normal((40, 1)), Y_train_std=2.0, save_path_file=base, Y_train=prob.Y,
                                          case="C%d" % case_val, noise=noise)
    # oracle
    tt = th.as_tensor
    ok = O._make_kernels(tt(prob.logv), tt(prob.logl), 0, prob.Z.shape[1])
    Linv = O.kernel_pre_cal(tt(prob.Z), ok)
    Q = tt(np.exp(prob.logQ))
    if case_val == 4:
        Xc = tt(np.concatenate([prob.X[:T], prob.ctrl], axis=1))
        U_val, q_sqrt = O.collapse_u_mean_after_kernel_precalculation(Linv, Xc, tt(prob.X), tt(prob.Z), ok, Q)
    else:
        U_val, q_sqrt = tt(prob.U), None
    xs, vs = zip(*[O.rollout(Linv, tt(prob.X[-1]), tt(ctrl[T:T + L]), tt(prob.Z), ok, U_val, q_sqrt, Q, tt(noise[i])) for i in range(num)])
    xs, vs = th.stack(xs).numpy(), th.stack(vs).numpy()
    # twelve sequential steps amplify rounding differences slightly: 1e-8 on the trajectories
    assert_close(xs, model.predict_x, 1e-8)
    assert_close(vs, model.predict_x_var, 1e-8)
    ref_y = (np.mean(np.einsum('ijk,kl->ijl', xs, prob.C), axis=0) + prob.d[None, :]).reshape(-1)
    assert_close(ref_y, py, 1e-8)
    ref_v = np.mean(np.einsum('ijk,kl->ijl', vs, prob.C ** 2), axis=0).reshape(-1) + np.exp(2 * prob.logR).reshape(-1)
    assert_close(ref_v, pv, 1e-8)
    res = np.load(base + "_results.npz", allow_pickle=True)
    for key in ("y_train_vfe", "y_test_vfe", "v_test_vfe_var", "Y_test_data", "Y_train_data", "Y_train_std", "CC_val", "DD_val",
                "log_R_cholesky", "log_QQ", "Z_val", "U_val", "X_val", "k_lengthscales", "k_log_variances", "case", "ll_seq",
                "running_time_seq", "PG_num", "mc_posterior_samples"):
        assert key in res.files, key                      # base_model.py:512-517
    assert_close((prob.X[1:] @ prob.C + prob.d).reshape(-1), res["y_train_vfe"], 1e-13)
    assert model.RMSE_val is not None and np.isfinite(model.RMSE_val)


def test_outer_training_loop(env):
    """Model._fit (models.py:142-168): each iteration = sghmc_step (21 evaluations; a no-op update set in case 4) +
    one Adam step on the trainable set.  The default case 4 must lower the collapsed nll."""
    prob = env["byname"]["actuator/0"]
    m, _ = _model_from_problem(env, prob, 4, iterations=5)
    model = m.fit(prob.Y, kernel_type="SquaredExponential")          # 2 * iterations = 10 outer iterations
    assert model.vars == [] and set(model.trainable) == {"X", "Z", "logv", "logl", "logQ", "C", "d", "logR"}
    assert model.adam_step == 10 and m.global_step == 10
    nll_end = float(model.nll)
    m0, _ = _model_from_problem(env, prob, 4, iterations=0)
    nll_start = float(m0.fit(prob.Y).nll)
    assert np.isfinite(nll_end) and nll_end < nll_start
    # case 2 samples {logv, logl, U} and keeps a window of snapshots
    m2, _ = _model_from_problem(env, prob, 2, iterations=1)
    model2 = m2.fit(prob.Y)
    assert model2.vars == ["logv", "logl", "U"] and len(model2.window) == 2
    assert np.isfinite(float(model2.nll))


def test_time_blocks_reassemble_full_trajectory(env):
    """Time sharding (SURVEY 8e, S < #GPUs): two blocks of one trajectory evaluated with FFVD_FLAG_NO_SHARED_PRIORS /
    FFVD_FLAG_NO_X0_PRIOR on the second, rescaled by T_b / T and stitched at the halo row, equal the full evaluation
    (the collective itself is covered by the gloo test; here the CUDA path produces the blocks)."""
    from oracle import fixtures
    from ffvd_b200 import distributed
    torch, ctx = env["torch"], env["ctx"]
    prob = fixtures.synthetic_problem(T=157, M=70, D=3, S=1)
    p = dev_problem(env, prob)
    full = alloc_out(env, p)
    ctx.nll_grads(0, False, p, full)
    T = prob.Y.shape[0]
    acc = {k: torch.zeros_like(v) for k, v in full.items() if k != "g_X"}
    gX = torch.zeros_like(full["g_X"])
    for rank in range(3):
        blk, a, b = distributed.time_block(p, rank, 3)
        blk = {k: (v.contiguous() if v is not None else None) for k, v in blk.items()}
        o = alloc_out(env, blk)
        ctx.nll_grads(0, False, blk, o, flags=env["ffvd"].FLAG_PRIOR_Z_NORMAL | distributed.time_block_flags(rank))
        sc = (b - a) / T
        for k in acc:
            acc[k] += sc * o[k]
        gX[a:b + 1] += sc * o["g_X"]
    torch.cuda.synchronize()
    for k in acc:
        assert_close(full[k].cpu().numpy(), acc[k].cpu().numpy(), 1e-11, k)
    assert_close(full["g_X"].cpu().numpy(), gX.cpu().numpy(), 1e-11, "g_X")


def test_prepared_call_sees_in_place_updates(env):
    """Context.prepare_nll_grads binds tensors once; re-running after an in-place parameter update must give the same
    result as a fresh (unbound) call -- single problem and batch."""
    from oracle import fixtures
    torch, ctx = env["torch"], env["ctx"]
    probs = [fixtures.synthetic_problem(T=90 + 10 * i, M=33, D=2, S=1, seed=i) for i in range(6)]
    Ps = [dev_problem(env, q) for q in probs]
    Os = [alloc_out(env, p) for p in Ps]
    call = ctx.prepare_nll_grads(0, True, Ps, Os)
    call.run()
    for p in Ps:
        p["X"].mul_(0.9); p["Z"].add_(0.01)
    call.run()
    torch.cuda.synchronize()
    fresh = [alloc_out(env, p) for p in Ps]
    ctx.nll_grads_batched(0, True, Ps, fresh)
    torch.cuda.synchronize()
    for a, b in zip(Os, fresh):
        for k in a:
            assert_close(b[k].cpu().numpy(), a[k].cpu().numpy(), 1e-10, k)      # FP64 RED order differs from run to run
    call.close()
    one = ctx.prepare_nll_grads(0, False, Ps[0], Os[0])
    one.run(); torch.cuda.synchronize()
    ref = alloc_out(env, Ps[0]); ctx.nll_grads(0, False, Ps[0], ref); torch.cuda.synchronize()
    for k in ref:
        assert_close(ref[k].cpu().numpy(), Os[0][k].cpu().numpy(), 1e-10, k)


def test_kzz_reuse_flag(env):
    """FFVD_FLAG_REUSE_KZZ: with Z and the kernel hyper-parameters unchanged, skipping the Cholesky preparation gives the
    same results; a changed Z pointer / shape falls back to a full preparation; a case-7 SG-HMC step (X, U sampled) with
    reuse equals the same step without it."""
    from oracle import fixtures
    torch, ctx, F = env["torch"], env["ctx"], env["ffvd"]
    prob = fixtures.synthetic_problem(T=120, M=90, D=3, S=2)
    p = dev_problem(env, prob)
    a, b = alloc_out(env, p), alloc_out(env, p)
    ctx.nll_grads(0, False, p, a)
    p["X"].mul_(0.97); p["U"].add_(0.05)                       # not Z / hypers
    ctx.nll_grads(0, False, p, a, flags=F.FLAG_PRIOR_Z_NORMAL | F.FLAG_REUSE_KZZ)
    ctx.nll_grads(0, False, p, b)
    torch.cuda.synchronize()
    for k in a:
        assert_close(b[k].cpu().numpy(), a[k].cpu().numpy(), 1e-10, k)
    # a different Z tensor (new address) must not reuse the old factors even if the flag is set
    p2 = dict(p); p2["Z"] = (p["Z"] * 1.1).contiguous()
    ctx.nll_grads(0, False, p2, a, flags=F.FLAG_PRIOR_Z_NORMAL | F.FLAG_REUSE_KZZ)
    ctx.nll_grads(0, False, p2, b)
    torch.cuda.synchronize()
    for k in a:
        assert_close(b[k].cpu().numpy(), a[k].cpu().numpy(), 1e-10, k)
    # end to end: one case-7 sghmc_step with and without the reuse
    fx = env["byname"]["drive/0"]
    res = []
    for reuse in (True, False):
        m, _ = _model_from_problem(env, fx, 7, iterations=0)
        model = m.fit(fx.Y)
        rng = np.random.default_rng(9)
        noises = [{n: rng.standard_normal(tuple(model.params[n].shape)) for n in model.vars} for _ in range(21)]
        if not reuse:
            model.invalidate_kzz()
            orig = model._run_update
            def no_reuse(burn_in, noise=None, _o=orig, _m=model):
                _m.invalidate_kzz()
                return _o(burn_in, noise)
            model._run_update = no_reuse
        model.sghmc_step(noise_fn=lambda k: noises[k])
        res.append({n: model.params[n].cpu().numpy().copy() for n in model.vars})
    assert sorted(res[0]) == ["U", "X"]
    for n in res[0]:
        assert_close(res[1][n], res[0][n], 1e-9, n)


def test_particle_gibbs_sweep(env):
    """PG_for_X (base_model.py:29-75): the conditional-SMC sweep with injected randomness against the
    oracle's literal restatement (whole trajectories gathered at every resampling)."""
    import torch as th
    from oracle import ffvd_oracle as O
    prob = env["byname"]["gas_furnace/0"]
    T, D = prob.Y.shape[0], prob.X.shape[1]
    m, ctrl = _model_from_problem(env, prob, 6, iterations=0)
    model = m.fit(prob.Y)
    P = 12
    rng = np.random.default_rng(21)
    normals, eps, uni = rng.standard_normal((P - 1, D)), rng.standard_normal((T, P - 1, D)), rng.random((T, P - 1))
    path = model.PG_for_X(prob.ctrl, P, normals=normals, eps=eps, uniforms=uni, assign=False).cpu().numpy()
    tt = th.as_tensor
    ok = O._make_kernels(tt(prob.logv), tt(prob.logl), 0, prob.Z.shape[1])
    ref = O.pg_for_x(tt(prob.X), tt(prob.Y), tt(prob.ctrl), tt(prob.Z), ok, tt(prob.U), tt(np.exp(prob.logQ)), tt(prob.C), tt(prob.d),
                     tt(np.exp(prob.logR)), P, tt(normals), tt(eps), tt(uni)).numpy()
    assert path.shape == ref.shape == (T + 1, D)
    assert_close(ref, path, 1e-8)
    # the sweep must be able to return the reference trajectory itself and free particles alike: with all the resampling
    # mass forced onto the last candidate (u -> 1) the reference particle survives everywhere
    same = model.PG_for_X(prob.ctrl, P, normals=normals, eps=eps, uniforms=np.full((T, P - 1), 1.0 - 1e-12), assign=False)
    assert_close(prob.X, same.cpu().numpy(), 1e-15)
    # case 6 outer loop runs: sghmc_step (nothing sampled) + PG sweep + Adam
    m6, _ = _model_from_problem(env, prob, 6, iterations=1)
    m6.ARGS.PG_particles = 8
    mod6 = m6.fit(prob.Y)
    assert np.isfinite(float(mod6.nll))


# ---- SURVEY 8(f) rows against golden vectors from the REFERENCE'S OWN SOURCE (reference_shim_golden_next.npz) -----------
@pytest.fixture(scope="module")
def gold_next():
    import os
    from util import GOLDEN
    return np.load(os.path.join(GOLDEN, "reference_shim_golden_next.npz"), allow_pickle=False)


def test_f2_collapsed_qu_and_conditional_vs_reference_source(env, gold_next):
    """cmo:206-227 and cmo:306-387 (q_sqrt 3-d with the reference's first-output broadcasting, and None) -- CUDA vs the
    reference source's outputs."""
    import torch as th
    from ffvd_b200 import conditionals_multi_output as cmo
    from ffvd_b200.kernels_multi_output import SquaredExponential
    g = gold_next
    prob = env["byname"]["actuator/0"]
    T, D, Din = prob.Y.shape[0], prob.X.shape[1], prob.Z.shape[1]
    kerns = [SquaredExponential(Din, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(D)]
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    Xc = np.concatenate([prob.X[:T], prob.ctrl], axis=1)
    U_mean, Lseq = cmo.collapse_u_mean_after_kernel_precalculation(None, t(Xc), t(prob.X), t(prob.Z), kerns, t(np.exp(prob.logQ)))
    assert_close(g["f2/U_mean"], U_mean.cpu().numpy(), TOL, "U_mean")
    assert_close(g["f2/LHinvT"], Lseq.cpu().numpy(), TOL, "LHinvT")
    scale = float(np.max(np.exp(prob.logv)))
    for tag, q in (("q3", t(g["f2/LHinvT"])), ("qnone", None)):
        mu, var = cmo.conditional_after_kernel_precalculation(None, t(g["f2/Xnew"]), t(prob.Z), kerns, t(g["f2/U_mean"][0]), white=True, q_sqrt=q)
        assert_close(g["f2/cond_%s/mean" % tag], mu.cpu().numpy(), TOL, tag)
        assert np.max(np.abs(g["f2/cond_%s/var" % tag] - var.cpu().numpy())) <= TOL * scale


@pytest.mark.parametrize("case_val,tag", ((4, "collapsed"), (2, "uncollapsed")))
def test_f2_rollout_vs_reference_source(env, gold_next, case_val, tag, tmp_path):
    """`collect_samples_formal` (base_model.py:197-522) on the CUDA path vs the same method executed from the reference
    source with its logged noise: predict_y, predict_y_var, fit_y, RMSE and the results-file keys."""
    g = gold_next
    prob = env["byname"]["actuator/0"]
    m, ctrl = _model_from_problem(env, prob, case_val, iterations=0, extra_ctrl=g["f2/ctrl_future"])
    model = m.fit(prob.Y, kernel_type="SquaredExponential")
    base = str(tmp_path / "run")
    py, pv = model.collect_samples_formal(3, 1, ctrl, 30, sghmc_var_len=0, U_collapse=(case_val == 4), Y_test=g["f2/Y_test"], Y_train_std=1.7,
                                          save_path_file=base, Y_train=prob.Y, case="C%d" % case_val, noise=g["f2/rollout_%s/noise" % tag])
    assert_close(g["f2/rollout_%s/predict_y" % tag], py, TOL, "predict_y")
    assert_close(g["f2/rollout_%s/predict_y_var" % tag].reshape(-1), np.asarray(pv).reshape(-1), TOL, "predict_y_var")
    assert_close(g["f2/rollout_%s/fit_y" % tag], model.fit_y, 1e-12, "fit_y")
    assert abs(model.RMSE_val - float(g["f2/rollout_%s/RMSE" % tag])) <= TOL * model.RMSE_val
    res = np.load(base + "_results.npz", allow_pickle=True)
    assert sorted(res.files) == sorted(str(k) for k in g["f2/rollout_%s/file_keys" % tag])
    for k in ("y_train_vfe", "y_test_vfe", "X_val", "Z_val", "log_QQ"):
        assert_close(g["f2/rollout_%s/file/%s" % (tag, k)], np.asarray(res[k], dtype=np.float64), TOL, k)


def test_f4_particle_gibbs_vs_reference_source(env, gold_next):
    """`BaseModel.PG_for_X` (base_model.py:29-75) on the CUDA path vs the reference source's sweep with its logged draws."""
    g = gold_next
    prob = env["byname"]["gas_furnace/0"]
    m, _ = _model_from_problem(env, prob, 6, iterations=0)
    model = m.fit(prob.Y)
    path = model.PG_for_X(prob.ctrl, int(g["f4/P"]), normals=g["f4/normals"], eps=g["f4/eps"], uniforms=g["f4/uniforms"], assign=False)
    assert_close(g["f4/X_after"], path.cpu().numpy(), 1e-8, "PG trajectory")


@pytest.mark.parametrize("case_val", (2, 4))
def test_f3_outer_loop_vs_reference_source(env, gold_next, case_val):
    """`Model._fit` (models.py:142-168: sghmc_step + train_hypers per outer iteration) on the CUDA path vs the reference's
    update expressions re-executed per session.run: final value of EVERY parameter after 2 (case 2: kernel hypers and U
    SG-HMC sampled, 42 SG-HMC evaluations) / 4 (case 4, the CLI default: Adam on everything, collapsed bound) iterations."""
    from oracle import loop
    g = gold_next
    key = "f3/case%d" % case_val
    prob = env["byname"]["actuator/0"]
    iters = int(g[key + "/iters"])
    D = prob.X.shape[1]
    nf = loop.reference_noise_fn(g, key, D)
    m, _ = _model_from_problem(env, prob, case_val, iterations=iters // 2)
    widx = [int(i) for i in g[key + "/window_index"]]
    noise = (lambda it, k: {n: nf(21 * it + k, n) for n in ("logv", "logl", "U")}) if case_val == 2 else None
    model = m.fit(prob.Y, kernel_type="SquaredExponential", sghmc_noise=noise, window_index=lambda it: widx[it])
    assert m.global_step == iters
    for k in GKEYS:
        assert_close(g["%s/%s" % (key, k)], model.params[k].cpu().numpy(), 5e-9, "%s %s" % (key, k))
    assert abs(float(model.nll) - float(g[key + "/nll_final"])) <= 1e-9 * abs(float(g[key + "/nll_final"]))


@pytest.mark.parametrize("M", (90, 200))
def test_reuse_kzz_content_guard(env, M):
    """FFVD_FLAG_REUSE_KZZ is keyed on tensor addresses; the device-side content hash of Z / logv / logl must catch an
    IN-PLACE change: the reusing call returns FFVD_E_STALE (nll = NaN under FFVD_FLAG_ASYNC) instead of silently using the
    factors of another Z, and the factors of a FAILED factorisation are never reused (single-CTA and blocked paths)."""
    from oracle import fixtures, ffvd_oracle as O
    torch, ctx, F = env["torch"], env["ctx"], env["ffvd"]
    prob = fixtures.synthetic_problem(T=100, M=M, D=3, S=1, seed=5)
    p = dev_problem(env, prob)
    o = alloc_out(env, p)
    base = F.FLAG_PRIOR_Z_NORMAL
    ctx.nll_grads(0, False, p, o, flags=base)
    ctx.nll_grads(0, False, p, o, flags=base | F.FLAG_REUSE_KZZ)            # legitimate reuse
    assert np.isfinite(o["nll"].cpu().numpy()).all()
    p["Z"].mul_(1.01)                                                        # same address, new contents
    with pytest.raises(F.StaleFactorsError):
        ctx.nll_grads(0, False, p, o, flags=base | F.FLAG_REUSE_KZZ)
    assert np.isnan(o["nll"].cpu().numpy()).all()
    ctx.nll_grads(0, False, p, o, flags=base | F.FLAG_REUSE_KZZ)            # the host dropped the record: a full preparation
    prob2 = copy.deepcopy(prob); prob2.Z = p["Z"].cpu().numpy()
    check(O.nll_and_grads(prob2, collapsed=False), {k: (v.cpu().numpy()[0] if k in ("nll", "terms") else v.cpu().numpy()) for k, v in o.items()},
          what="after stale")
    # asynchronous caller: no status is read, the NaN is the signal
    p["logl"].add_(0.01)
    ctx.nll_grads(0, False, p, o, flags=base | F.FLAG_REUSE_KZZ | F.FLAG_ASYNC)
    torch.cuda.synchronize()
    assert np.isnan(o["nll"].cpu().numpy()).all()
    ctx.nll_grads(0, False, p, o, flags=base)
    # a failed factorisation (asynchronous call, status never read) must not be reusable
    ctx.nll_grads(0, False, p, o, flags=base | F.FLAG_ASYNC, jitter=-1.0e3)
    ctx.nll_grads(0, False, p, o, flags=base | F.FLAG_REUSE_KZZ | F.FLAG_ASYNC, jitter=-1.0e3)
    torch.cuda.synchronize()
    assert np.isnan(o["nll"].cpu().numpy()).all()
    ctx.nll_grads(0, False, p, o, flags=base)
    assert np.isfinite(o["nll"].cpu().numpy()).all()


def test_device_data_generator(env):
    """bench.py's device-side AR(1) generator (shapes too large to ship from the host) against scipy's filter
    on the same innovations, across block boundaries."""
    import bench
    from scipy.signal import lfilter
    torch = env["torch"]
    eps = torch.randn((3, 1037, 4), dtype=torch.float64, device=env["dev"])
    x = bench.ar1_filter_device(eps.clone()).cpu().numpy()
    ref = lfilter([1.0], [1.0, -0.95], eps.cpu().numpy(), axis=1)
    assert np.max(np.abs(x - ref)) <= 1e-12
    d = bench.make_device_data(500, 64, 3, 2, seed=1, dev=env["dev"])
    assert tuple(d["X"].shape) == (2, 501, 3) and tuple(d["Y"].shape) == (500, 1) and tuple(d["ctrl"].shape) == (500, 1)
    assert abs(float(d["X"].std()) - 1.0) < 0.3


def test_logdensity_norm_full_factor_and_multi_output_gaussian(env):
    """likelihoods.py:114-127 with a full lower-triangular factor (device forward substitution) vs the reference-source
    golden; a single y row broadcast over ymean (base_model.py:62-66); Gaussian(Y_dim > 1) carries the reference's fixed
    masked factor (likelihoods.py:56-61)."""
    import torch as th
    from ffvd_b200 import likelihoods
    g = load_golden()
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    y, ym, Rc = g["op/ld/y"], g["op/ld/ymean"], g["op/ld/Rfull"]
    out = likelihoods.logdensity_norm(t(y), t(ym), t(Rc))
    assert_close(g["op/ld/full"], out.cpu().numpy(), 1e-13, "logdensity_norm full factor")
    assert_close(g["op/ld/full"], likelihoods.logdensity_norm(y, ym, Rc), 1e-13, "host tensors")       # NumPy in, NumPy out
    # garbage above the diagonal is ignored (triangular_solve(lower=True))
    Rg = Rc.copy(); Rg[0, 1] = 123.0
    assert_close(g["op/ld/full"], likelihoods.logdensity_norm(t(y), t(ym), t(Rg)).cpu().numpy(), 1e-13)
    # one row of y against many rows of ymean
    ref = np.array([-0.5 * np.sum(np.linalg.solve(np.tril(Rc), (y[3] - ym[n])) ** 2) - np.sum(np.log(np.diag(Rc))) for n in range(ym.shape[0])])
    assert_close(ref, likelihoods.logdensity_norm(t(y[3]), t(ym), t(Rc)).cpu().numpy(), 1e-12, "broadcast row")
    # a 5-output factor
    rng = np.random.default_rng(3)
    L5 = np.tril(rng.standard_normal((5, 5))) + 3.0 * np.eye(5)
    y5, m5 = rng.standard_normal((40, 5)), rng.standard_normal((40, 5))
    ref5 = -0.5 * np.sum(np.linalg.solve(L5, (y5 - m5).T) ** 2, axis=0) - np.sum(np.log(np.diag(L5)))
    assert_close(ref5, likelihoods.logdensity_norm(t(y5), t(m5), t(L5)).cpu().numpy(), 1e-12, "Dy=5")
    lik = likelihoods.Gaussian(3, 4)
    assert lik.Rchols.shape == (3, 3) and np.allclose(np.diag(lik.Rchols), 1.5) and lik.Rchols[1, 0] == 1.0 and lik.Rchols[0, 1] == 0.0
    assert not hasattr(lik, "log_Rchols")
    pm = lik.predict_mean(t(rng.standard_normal((7, 4))))
    assert tuple(pm.shape) == (7, 3)


def test_device_exp_routine_accuracy(env):
    """The branch-free table exp of the K tile (kernels_multi_output.py:246-247: K_r2 = exp(-r2/2)) against libm: <= 2 ulp
    over the whole range the SE argument can take, exact 1 at 0, saturation (not garbage) for hugely negative arguments."""
    torch, ctx = env["torch"], env["ctx"]
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.uniform(0, 50, 200000), -rng.uniform(0, 1e-3, 20000), -rng.uniform(50, 700, 50000), [0.0, -1e-300, -0.5, -708.0]])
    xd = torch.as_tensor(x, device=env["dev"])
    out = ctx.debug_exp(xd, torch.empty_like(xd)).cpu().numpy()
    ref = np.exp(x)
    ulp = np.abs(out - ref) / np.spacing(ref)
    assert ulp.max() <= 2.0, ulp.max()
    assert out[-4] == 1.0
    big = torch.as_tensor(np.array([-746.0, -1e4, -1e9, -1e15, -1e300, -np.inf]), device=env["dev"])
    sat = ctx.debug_exp(big, torch.empty_like(big)).cpu().numpy()
    assert np.all(np.isfinite(sat)) and np.all(sat >= 0) and np.all(sat < 1e-300)


@pytest.mark.parametrize("M", (70, 150))
def test_collapsed_time_blocks_reassemble_full_trajectory(env, M):
    """The collapsed bound under TIME sharding (cmo:246-254 sums F^T F over all transitions): three blocks, each on its
    own context (the ranks), pass 1 -> summed statistics (ffvd_collapsed_stats_get / _set) -> resume with
    FFVD_FLAG_NO_REPLICATED on all but the first -> the T_b/T-weighted sum equals the single evaluation of the whole
    trajectory (nll, all six terms, every gradient)."""
    from oracle import fixtures
    from ffvd_b200 import distributed
    torch, F = env["torch"], env["ffvd"]
    prob = fixtures.synthetic_problem(T=211, M=M, D=3, S=1)
    p = dev_problem(env, prob)
    full = alloc_out(env, p)
    env["ctx"].nll_grads(0, True, p, full)
    T = prob.Y.shape[0]
    world = 3
    ctxs = [F.Context(0, torch.cuda.current_stream(0).cuda_stream) for _ in range(world)]
    blks, outs = [], []
    for rank in range(world):
        blk, a, b = distributed.time_block(p, rank, world)
        blk = {k: (v.contiguous() if v is not None else None) for k, v in blk.items()}
        o = alloc_out(env, blk)
        fl = F.FLAG_PRIOR_Z_NORMAL | distributed.time_block_flags(rank)
        ctxs[rank].nll_grads(0, True, blk, o, flags=fl | F.FLAG_COLLAPSED_P1_ONLY)
        blks.append((blk, a, b, fl)); outs.append(o)
    nb, Mp = ctxs[0].collapsed_stats_shape()
    assert nb == 3
    Ssum = torch.zeros((nb, Mp, Mp), dtype=torch.float64, device=env["dev"]); bsum = torch.zeros((nb, Mp), dtype=torch.float64, device=env["dev"])
    for rank in range(world):
        S = torch.empty_like(Ssum); bv = torch.empty_like(bsum)
        ctxs[rank].collapsed_stats_get(S, bv)
        Ssum += S; bsum += bv
    acc = {k: torch.zeros_like(v) for k, v in full.items() if k != "g_X"}
    gX = torch.zeros_like(full["g_X"])
    for rank in range(world):
        blk, a, b, fl = blks[rank]
        ctxs[rank].collapsed_stats_set(Ssum, bsum)
        ctxs[rank].nll_grads(0, True, blk, outs[rank], flags=fl | F.FLAG_COLLAPSED_RESUME | (F.FLAG_NO_REPLICATED if rank else 0))
        sc = (b - a) / T
        for k in acc:
            acc[k] += sc * outs[rank][k]
        gX[a:b + 1] += sc * outs[rank]["g_X"]
    torch.cuda.synchronize()
    for k in acc:
        assert_close(full[k].cpu().numpy(), acc[k].cpu().numpy(), 1e-10, k)
    assert_close(full["g_X"].cpu().numpy(), gX.cpu().numpy(), 1e-10, "g_X")
    # protocol errors are reported, not silently computed
    with pytest.raises(ValueError):
        ctxs[0].nll_grads(0, True, blks[0][0], outs[0], flags=blks[0][3] | F.FLAG_COLLAPSED_RESUME)       # nothing pending any more
    with pytest.raises(ValueError):
        ctxs[0].nll_grads(0, False, blks[0][0], outs[0], flags=blks[0][3] | F.FLAG_COLLAPSED_P1_ONLY)     # uncollapsed


def test_cuda_graph_sghmc_step_and_capture_rules(env):
    """CUDA-graph form of `sghmc_step` (21 evaluations + updates as ONE launch, base_model.py:915-933): identical result to
    the launch-by-launch step with the same injected noise, for case 7 (X and U sampled, factors reused inside the graph)
    and case 2 (kernel hyper-parameters sampled: full preparation in every evaluation); replay after an in-place
    parameter change; the capture rules are enforced."""
    torch, F = env["torch"], env["ffvd"]
    fx = env["byname"]["drive/0"]
    for case_val in (7, 2):
        res = []
        for graph in (False, True):
            m, _ = _model_from_problem(env, fx, case_val, iterations=0)
            model = m.fit(fx.Y)
            rng = np.random.default_rng(9)
            noises = [[{n: rng.standard_normal(tuple(model.params[n].shape)) for n in model.vars} for _ in range(21)] for _ in range(2)]
            if graph:
                model.enable_graph(True)
            for it in range(2):                          # second step = a replay of the captured graph
                model.sghmc_step(noise_fn=lambda k, _it=it: noises[_it][k])
            if graph:
                assert model._graph.kernels >= 21 * 5
            res.append({n: model.params[n].cpu().numpy().copy() for n in model.vars})
            assert len(model.window) == 2
        for n in res[0]:
            assert_close(res[0][n], res[1][n], 1e-9, "case %d %s" % (case_val, n))
    # capture rules
    from oracle import fixtures
    ctx = F.Context(0, torch.cuda.current_stream(0).cuda_stream)
    p = dev_problem(env, fixtures.synthetic_problem(T=40, M=20, D=2, S=1))
    o = alloc_out(env, p)
    with pytest.raises(ValueError):                      # no warm-up call: the workspace does not exist yet
        ctx.capture(lambda: ctx.nll_grads(0, False, p, o, flags=F.FLAG_PRIOR_Z_NORMAL | F.FLAG_ASYNC))
    ctx.nll_grads(0, False, p, o)
    with pytest.raises(ValueError):                      # status read-back inside a graph
        ctx.capture(lambda: ctx.nll_grads(0, False, p, o, flags=F.FLAG_PRIOR_Z_NORMAL))
    g = ctx.capture(lambda: ctx.nll_grads(0, False, p, o, flags=F.FLAG_PRIOR_Z_NORMAL | F.FLAG_ASYNC))
    ref = {k: v.clone() for k, v in o.items()}
    for v in o.values():
        v.fill_(float("nan"))
    g.launch(); torch.cuda.synchronize()
    for k in ref:
        assert_close(ref[k].cpu().numpy(), o[k].cpu().numpy(), 1e-10, k)
    g.close()


def test_mirrors_reject_inconsistent_precalculated_inputs(env):
    """The fused path recomputes the K(Z,Z) factors and rebuilds [x_t, c_t] from X; `Lm_inverse_seq` / `X_combine` that do
    not match (Z, kern) / X must raise instead of being silently ignored -- and consistent ones (a `kernel_pre_cal`
    result, or plain matrices equal to it) are accepted."""
    import torch as th
    from ffvd_b200 import conditionals_multi_output as cmo
    from ffvd_b200.kernels_multi_output import SquaredExponential
    prob = env["byname"]["gas_furnace/0"]
    T, D, Din = prob.Y.shape[0], prob.X.shape[1], prob.Z.shape[1]
    kerns = [SquaredExponential(Din, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(D)]
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    X, Z, Q = t(prob.X), t(prob.Z), t(np.exp(prob.logQ))
    Xc = t(np.concatenate([prob.X[:T], prob.ctrl], axis=1))
    fac = cmo.kernel_pre_cal(Z, kerns)
    base = cmo.collapse_after_kernel_precalculation(None, Xc, X, Z, kerns, Q, T, T)
    for L in (fac, [a.clone() for a in fac], [a.cpu().numpy() for a in fac]):
        got = cmo.collapse_after_kernel_precalculation(L, Xc, X, Z, kerns, Q, T, T)
        for a, b in zip(base, got):
            assert abs(float(a) - float(b)) <= 1e-12 * abs(float(a))
    bad_fac = [a * 1.01 for a in fac]
    Xc_bad = Xc.clone(); Xc_bad[3, 0] += 0.5
    with pytest.raises(ValueError):
        cmo.collapse_after_kernel_precalculation(bad_fac, Xc, X, Z, kerns, Q, T, T)
    with pytest.raises(ValueError):
        cmo.collapse_after_kernel_precalculation(fac, Xc_bad, X, Z, kerns, Q, T, T)
    with pytest.raises(ValueError):
        cmo.collapse_u_mean_after_kernel_precalculation(cmo.kernel_pre_cal(Z * 1.1, kerns), Xc, X, Z, kerns, Q)
    with pytest.raises(ValueError):
        cmo.conditional_after_kernel_precalculation(bad_fac, Xc[:5], Z, kerns, t(prob.U), white=True)
    mu0, _ = cmo.conditional_after_kernel_precalculation(None, Xc[:5].contiguous(), Z, kerns, t(prob.U), white=True)
    mu1, _ = cmo.conditional_after_kernel_precalculation([a.clone() for a in fac], Xc[:5].contiguous(), Z, kerns, t(prob.U), white=True)
    assert_close(mu0.cpu().numpy(), mu1.cpu().numpy(), 1e-12)


def test_bounds_checked_build():
    """tools/bounds_check.py: the parity sweep on the -DFFVD_BOUNDS_CHECK library (device-side assertions on the fused
    kernels' index arithmetic).  Skipped if `make check` was not run."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "ffvd_b200", "lib", "libffvd_b200_check.so")):
        pytest.skip("make check not run")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bounds_check.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout


def test_conditional_option_branches_vs_reference_source(env, gold_next):
    """The dense conditional (ffvd_conditional_dense): full_cov, q_sqrt 2-d / 3-d whitened or not, return_Lm
    (conditionals.py:27-66) and the multi-output full_cov quirk shape (cmo:119-120) vs the reference source's outputs."""
    import torch as th
    from ffvd_b200 import conditionals, conditionals_multi_output as cmo
    from ffvd_b200.kernels_multi_output import SquaredExponential
    g = gold_next
    t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
    Z, Xn, f, q3, q2 = (t(g["cond/" + k]) for k in ("Z", "Xnew", "f", "q3", "q2"))
    se = SquaredExponential(Z.shape[1], variance=0.7, lengthscales=g["cond/ls"], ARD=True)
    for tag, kw in (("full", dict(full_cov=True, white=True)), ("full_q3", dict(full_cov=True, white=True, q_sqrt=q3)),
                    ("full_q2", dict(full_cov=True, white=True, q_sqrt=q2)), ("nonwhite_q3", dict(white=False, q_sqrt=q3)),
                    ("nonwhite_q2", dict(white=False, q_sqrt=q2)), ("full_nonwhite_q3", dict(full_cov=True, white=False, q_sqrt=q3))):
        mu, var = conditionals.conditional(Xn, Z, se, f, **kw)
        assert tuple(var.shape) == tuple(g["cond/%s/var" % tag].shape)
        assert_close(g["cond/%s/mean" % tag], mu.cpu().numpy(), TOL, tag)
        assert_close(g["cond/%s/var" % tag], var.cpu().numpy(), TOL, tag)
    mu, var, Lm = conditionals.conditional(Xn, Z, se, f, white=True, return_Lm=True)
    assert_close(g["cond/return_Lm/mean"], mu.cpu().numpy(), TOL)
    assert_close(g["cond/return_Lm/var"], var.cpu().numpy(), TOL)
    assert_close(g["cond/return_Lm/Lm"], Lm.cpu().numpy(), 1e-11, "Lm")
    se2 = SquaredExponential(Z.shape[1], variance=0.3, lengthscales=g["cond/ls"][::-1].copy(), ARD=True)
    mu, var = cmo.conditional(Xn, Z, [se, se2], f[:, :2].contiguous(), white=True, full_cov=True)
    assert tuple(var.shape) == (Xn.shape[0], 1, 2)
    assert_close(g["cond/multi_full/mean"], mu.cpu().numpy(), TOL)
    assert_close(g["cond/multi_full/var"], var.cpu().numpy(), TOL)
    # NumPy tensors through the same path
    mu_h, var_h = conditionals.conditional(g["cond/Xnew"], g["cond/Z"], se, g["cond/f"], full_cov=True, white=True)
    assert_close(g["cond/full/var"], var_h, TOL)


def test_float32_tensor_mode(env):
    """float32 mode of the north star (<= 1e-4): every tensor crosses the ABI as float32 (device or host), is widened on the
    device, computed in float64 and narrowed once.  Against the float64 oracle on the SAME (float32-representable) inputs:
    nll + all gradients on a bundled warm start (both forms), the SG-HMC update in place, and mixed float32 / float64 calls."""
    import torch as th
    from oracle import ffvd_oracle as O
    torch, ctx, F = env["torch"], env["ctx"], env["ffvd"]
    prob = copy.deepcopy(env["byname"]["dryer/0"])
    for k in PKEYS:
        setattr(prob, k, np.asarray(getattr(prob, k), dtype=np.float32).astype(np.float64))      # float32-representable inputs
    for collapsed in (False, True):
        ref = O.nll_and_grads(prob, collapsed=collapsed)
        p32 = {k: th.as_tensor(np.ascontiguousarray(getattr(prob, k)), dtype=th.float32, device=env["dev"]) for k in PKEYS}
        o32 = {"nll": th.empty(1, dtype=th.float32, device=env["dev"]), "terms": th.empty(1, 6, dtype=th.float32, device=env["dev"])}
        for k in GKEYS:
            o32["g_" + k] = th.full_like(p32[k], float("nan"))
        ctx.nll_grads(0, collapsed, p32, o32)
        torch.cuda.synchronize()
        for k, v in o32.items():
            assert v.dtype == th.float32
            r = ref[k] if k not in ("nll", "terms") else np.asarray(ref[k]).reshape(v.shape)
            assert relerr(r, v.cpu().numpy().astype(np.float64)) <= 1e-4, (collapsed, k)
            assert relerr(r, v.cpu().numpy().astype(np.float64)) <= 5e-7, (collapsed, k)      # in fact only the final rounding
        # host float32 (NumPy) in and out, float64 outputs mixed in
        ph = {k: np.ascontiguousarray(getattr(prob, k), dtype=np.float32) for k in PKEYS}
        oh = {"nll": np.empty(1, dtype=np.float64), "g_X": np.empty(prob.X.shape, dtype=np.float32), "g_Z": np.empty(prob.Z.shape, dtype=np.float64)}
        ctx.nll_grads(0, collapsed, ph, oh)
        assert relerr(ref["g_X"], oh["g_X"].astype(np.float64)) <= 5e-7 and relerr(ref["g_Z"], oh["g_Z"]) <= 1e-9
        assert abs(oh["nll"][0] - ref["nll"]) <= 1e-9 * abs(ref["nll"])
    # SG-HMC update in place on float32 state
    rng = np.random.default_rng(1)
    n = 1001
    arrs = [rng.standard_normal(n).astype(np.float32) for _ in range(3)] + [np.ones(n, np.float32), np.ones(n, np.float32),
                                                                                np.full(n, 1.5, np.float32), (0.1 * rng.standard_normal(n)).astype(np.float32)]
    ref = O.sghmc_update(*[a.astype(np.float64) for a in arrs], epsilon=0.01, mdecay=0.05, X_N=201, burn_in=True)
    dev = [th.as_tensor(a, device=env["dev"]) for a in arrs]
    ctx.sghmc_update(*dev, 0.01, 0.05, 201.0, True)
    torch.cuda.synchronize()
    th_, gr, nz, xi, g, g2, pm = dev
    for a, b in zip(ref, (th_, xi, g, g2, pm)):
        assert relerr(a, b.cpu().numpy().astype(np.float64)) <= 5e-7
    with pytest.raises(ValueError):
        ctx.sghmc_update(*[d.to(th.float16) for d in dev], 0.01, 0.05, 201.0, True)


@pytest.mark.parametrize("collapsed,kind", ((False, 0), (True, 0), (False, 1), (True, 1)))
def test_block_of_dims_work_items(env, collapsed, kind):
    """The work-item form the large problems use (one item per (sample, tile, block of dims): x tile staged once, CTA
    loops over the dims) is chosen by problem size; force it (FFVD_DL=1) on small problems, with blocks that do not divide D
    (FFVD_DBLK=3, D=7), against the oracle -- both bounds, both kernels, and the conditional."""
    import os
    from oracle import fixtures, ffvd_oracle as O
    old = {k: os.environ.get(k) for k in ("FFVD_DL", "FFVD_DBLK")}
    try:
        for dblk in ("3", "8"):
            os.environ["FFVD_DL"] = "1"; os.environ["FFVD_DBLK"] = dblk
            for (T, M, D, S) in ((150, 40, 7, 2), (70, 150, 5, 1)):
                prob = fixtures.synthetic_problem(T=T, M=M, D=D, S=S, kind=kind, seed=3)
                check(O.nll_and_grads(prob, collapsed=collapsed), run_cuda(env, prob, collapsed), what="dl dblk=%s T%d M%d D%d" % (dblk, T, M, D))
        if kind == 0 and not collapsed:
            import torch as th
            from ffvd_b200 import conditionals_multi_output as cmo
            from ffvd_b200.kernels_multi_output import SquaredExponential
            prob = fixtures.synthetic_problem(T=90, M=50, D=7, S=1, seed=4)
            Din = prob.Z.shape[1]
            kerns = [SquaredExponential(Din, variance=np.exp(prob.logv[k]), lengthscales=np.exp(prob.logl[k]), ARD=True) for k in range(7)]
            ok = O._make_kernels(th.as_tensor(prob.logv), th.as_tensor(prob.logl), 0, Din)
            Xc = np.concatenate([prob.X[:-1], prob.ctrl], axis=1)
            t = lambda a: th.as_tensor(np.ascontiguousarray(a), dtype=th.float64, device=env["dev"])
            mu, var = cmo.conditional(t(Xc), t(prob.Z), kerns, t(prob.U), white=True)
            rmu, rvar = O.conditional_multi_output(th.as_tensor(Xc), th.as_tensor(prob.Z), ok, th.as_tensor(prob.U), white=True)
            assert_close(rmu.numpy(), mu.cpu().numpy(), TOL, "cond mean (dl)")
            assert np.max(np.abs(rvar.numpy() - var.cpu().numpy())) <= TOL * float(np.max(np.exp(prob.logv)))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("collapsed,kind,M", ((False, 0, 100), (True, 0, 100), (False, 1, 60), (True, 0, 200), (False, 0, 300)))
def test_deterministic_flag_is_bitwise_repeatable(env, collapsed, kind, M):
    """FFVD_FLAG_DETERMINISTIC (two-level reduction option): private accumulator copies summed in index order.
    Two evaluations of the same inputs are bit-identical in every output, the result agrees with the default (RED-ordered)
    path to rounding and with the oracle to the parity tolerance.  Enough work items that every CTA of the grid contributes."""
    from oracle import fixtures, ffvd_oracle as O
    F = env["ffvd"]
    prob = fixtures.synthetic_problem(T=3000, M=M, D=4, S=3, kind=kind, seed=11 + M)
    base = F.FLAG_PRIOR_Z_NORMAL
    a = run_cuda(env, prob, collapsed, flags=base | F.FLAG_DETERMINISTIC)
    plain = run_cuda(env, prob, collapsed, flags=base)
    b = run_cuda(env, prob, collapsed, flags=base | F.FLAG_DETERMINISTIC)
    for k in a:
        assert np.array_equal(a[k], b[k]), (k, float(np.max(np.abs(a[k] - b[k]))))
        assert relerr(plain[k], a[k]) <= TOL, k        # the default path differs by its RED order (1e-10 where the bound cancels: LinearK, collapsed M=200)
    check(O.nll_and_grads(prob, collapsed=collapsed), a, what="deterministic M%d" % M)
    # forward-only evaluation under the flag
    f1 = run_cuda(env, prob, collapsed, flags=base | F.FLAG_DETERMINISTIC | F.FLAG_NO_GRADS)
    f2 = run_cuda(env, prob, collapsed, flags=base | F.FLAG_DETERMINISTIC | F.FLAG_NO_GRADS)
    assert np.array_equal(f1["nll"], f2["nll"]) and relerr(a["nll"], f1["nll"]) <= 1e-12


def test_deterministic_flag_batched_chains_and_refusals(env):
    """The flag on a batched call (ragged T), and what it refuses: graph capture and the split collapsed evaluation."""
    import torch
    from oracle import fixtures
    F = env["ffvd"]
    probs = [fixtures.synthetic_problem(T=T, M=48, D=3, S=1, seed=T) for T in (700, 1500, 333)]
    ps = [dev_problem(env, p) for p in probs]
    res = []
    for _ in range(2):
        outs = [alloc_out(env, p) for p in ps]
        env["ctx"].nll_grads_batched(0, False, ps, outs, flags=F.FLAG_PRIOR_Z_NORMAL | F.FLAG_DETERMINISTIC)
        torch.cuda.synchronize()
        res.append([{k: v.cpu().numpy() for k, v in o.items()} for o in outs])
    for r0, r1, p in zip(res[0], res[1], probs):
        single = run_cuda(env, p, False)
        for k in r0:
            assert np.array_equal(r0[k], r1[k]), k
            assert relerr(single[k].reshape(r0[k].shape), r0[k]) <= 1e-11, k
    p = ps[0]
    with pytest.raises(F.FFVDError):
        env["ctx"].nll_grads(0, True, p, alloc_out(env, p), flags=F.FLAG_DETERMINISTIC | F.FLAG_COLLAPSED_P1_ONLY)


@pytest.mark.parametrize("collapsed", (False, True))
def test_extreme_kernel_variance_and_distance(env, collapsed):
    """The K tile's exp routine works on a table pre-scaled by v_d with an exponent-shift bound (exp_nmin) and an integer
    clamp of the argument: kernel variances from 4e-18 to 20 and a trajectory row 1e5 length-scales away from every inducing
    point (argument ~ -1e9: exact 0 in the reference, v * exp(-745) here) must give the oracle's nll and gradients."""
    from oracle import fixtures, ffvd_oracle as O
    prob = fixtures.synthetic_problem(T=200, M=40, D=3, S=1, seed=5)
    prob.logv = np.array([-40.0, 3.0, np.log(0.3)])
    prob.X = prob.X.copy()
    # the far row only under the uncollapsed bound (same K-tile code in both): its 1e10-sized residuals make the collapsed
    # bound's sums cancel to a level where the RED order of two runs differs by more than the tolerance
    if not collapsed:
        prob.X[7] *= 1.0e5
    ref = O.nll_and_grads(prob, collapsed=collapsed)
    got = run_cuda(env, prob, collapsed)
    assert np.all(np.isfinite(got["nll"])) and all(np.all(np.isfinite(got[k])) for k in got)
    check(ref, got, what="extreme v / distance")
