"""Oracle-driven outer training loop.  TEST INFRASTRUCTURE ONLY (never imported by the product).

Restates the schedule of `Model._fit` (models.py:142-168): per outer iteration one `sghmc_step`
(base_model.py:915-933: 1 burn-in update, then 10 x (burn-in, sample), then a window snapshot) and one
`train_hypers` (base_model.py:944-950: TF1 Adam on the trainable set with the sampled variables fed from a window
entry, at the constant rate of `get_minibatch()`), on top of `ffvd_oracle.nll_and_grads / sghmc_update / adam_update`.
Pinned against the reference's own source re-executed per `session.run`
(tests/golden/make_reference_golden_next.py, keys f3/*) in tests/test_oracle_golden_next.py.
"""
from __future__ import annotations

import copy
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import ffvd_oracle as O

# FFVD_Main.py:273-324: case_val -> (kernel_optimization, U_optimization, Z_optimization, U_collapse)
CASES = {1: (True, True, True, False), 2: (False, False, True, False), 3: (False, False, False, False),
         4: (True, False, True, True), 5: (False, False, True, True), 6: (True, True, True, False), 7: (False, False, False, False)}


def sampled_names(case_val: int, kind: int = 0) -> List[str]:
    """`dgp_model.py:213-244` (kernel_train_flag=True, hyperparameter_sampling=False)."""
    ko, uo, zo, uc = CASES[case_val]
    if case_val == 7:
        return ["U", "X"]
    v: List[str] = []
    if not ko:
        v += ["logv", "logl"] if kind == 0 else ["logv"]
    if not uo and not uc:
        v += ["U"]
    if not zo:
        v += ["Z"]
    return v


def trainable_names(case_val: int, kind: int = 0) -> List[str]:
    """tf `trainable=` flags: dgp_model.py:62-69 (X, U, Z), kernels_multi_output.py:156-160, dgp_model.py:182-184 (log_Q),
    likelihoods.py:17-24,50-54 (C, d, log_Rchols)."""
    ko, uo, zo, uc = CASES[case_val]
    tr: List[str] = []
    if case_val not in (6, 7):
        tr.append("X")
    if uo:
        tr.append("U")
    if zo:
        tr.append("Z")
    if ko and kind == 0:
        tr += ["logv", "logl"]
    if case_val != 7:
        tr.append("logQ")
    tr += ["C", "d", "logR"]
    return tr


def outer_loop(prob: O.Problem, case_val: int, iters: int, noise_fn: Optional[Callable[[int, str], np.ndarray]] = None,
               window_index: Optional[Sequence[int]] = None, epsilon: float = 0.01, mdecay: float = 0.05) -> Dict[str, object]:
    """Run `iters` outer iterations; returns {"params": final parameters, "nll": final nll, "evals": SG-HMC evaluations}.
    `noise_fn(e, name)` is the N(0,1) draw of SG-HMC evaluation e (0-based over the whole run) for parameter `name`;
    `window_index[it]` the window entry `train_hypers` feeds in iteration it."""
    prob = copy.deepcopy(prob)
    collapsed = CASES[case_val][3]
    vars_ = sampled_names(case_val, prob.kind)
    train = trainable_names(case_val, prob.kind)
    X_N = prob.X.shape[0]
    st = {n: [np.ones_like(getattr(prob, n)), np.ones_like(getattr(prob, n)), np.ones_like(getattr(prob, n)), np.zeros_like(getattr(prob, n))]
          for n in vars_}
    adam = {n: (np.zeros_like(getattr(prob, n)), np.zeros_like(getattr(prob, n))) for n in train}
    window: List[Dict[str, np.ndarray]] = []
    evals = 0
    rng = np.random.default_rng(0)

    def update(burn_in: bool):
        nonlocal evals
        if not vars_:
            return
        res = O.nll_and_grads(prob, collapsed=collapsed)
        new = {}
        for n in vars_:                                   # Jacobi: all gradients come from the pre-step values
            nz = noise_fn(evals, n) if noise_fn is not None else rng.standard_normal(getattr(prob, n).shape)
            xi, g, g2, p = st[n]
            th, xi, g, g2, p = O.sghmc_update(getattr(prob, n), res["g_" + n], nz, xi, g, g2, p, epsilon=epsilon, mdecay=mdecay,
                                              X_N=X_N, burn_in=burn_in)
            st[n] = [xi, g, g2, p]
            new[n] = th
        for n, th in new.items():
            setattr(prob, n, th)
        evals += 1

    for it in range(iters):
        update(True)
        for _ in range(10):
            update(True)
            update(False)
        window.append({n: np.array(getattr(prob, n), copy=True) for n in vars_})
        # train_hypers: the window feed is temporary (feed_dict), the Adam update applies to the stored values
        i = window_index[it] if window_index is not None else int(rng.integers(len(window)))
        saved = {n: getattr(prob, n) for n in vars_}
        for n, v in window[i].items():
            setattr(prob, n, v)
        res = O.nll_and_grads(prob, collapsed=collapsed)
        for n, v in saved.items():
            setattr(prob, n, v)
        lr = O.adam_learning_rate()                       # get_minibatch() default: global_step = 1 (base_model.py:945)
        for n in train:
            m, v = adam[n]
            th, m, v = O.adam_update(getattr(prob, n), res["g_" + n], m, v, step=it + 1, lr=lr)
            adam[n] = (m, v)
            setattr(prob, n, th)
    final = O.nll_and_grads(prob, collapsed=collapsed)
    return {"params": {k: np.array(getattr(prob, k), copy=True) for k in O.PARAM_NAMES if getattr(prob, k) is not None},
            "nll": float(final["nll"]), "evals": evals}


def reference_noise_fn(gold, key: str, D: int):
    """The golden noise of a f3 run, re-keyed by parameter name: the reference's variable order is
    [logv_0, logl_0, logv_1, logl_1, ..., U] (dgp_model.py:225-233)."""
    def fn(e, name):
        if name == "logv":
            return np.array([float(gold["%s/noise_var%d" % (key, 2 * k)][e]) for k in range(D)])
        if name == "logl":
            return np.stack([gold["%s/noise_var%d" % (key, 2 * k + 1)][e] for k in range(D)])
        if name == "U":
            return gold["%s/noise_var%d" % (key, 2 * D)][e]
        raise KeyError(name)
    return fn
