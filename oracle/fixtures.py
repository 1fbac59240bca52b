"""Fixture loading for the oracle / parity tests.  TEST INFRASTRUCTURE ONLY.

Two sources:
  * ``load_reference_problem`` reads the reference tree directly (datasets under
    ``data/`` and warm starts under ``Factnonlin_ini/``) and restates
    `FFVD_Main.py:134-171` (create_dataset) and the npz -> parameter mapping of
    `FFVD_Main.py:212-254,340`, `dgp_model.py:56-69,182`, `likelihoods.py:20-24,54`.
    Only usable where ``/root/reference`` exists (the build container).
  * ``load_packed`` reads ``tests/golden/fixtures.npz`` -- the compact copy of the
    SAME inputs (only the fields the hot path consumes) that travels to the GPU box.
    ``tests/golden/make_fixtures.py`` writes it from the first source.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List, Tuple

import numpy as np

from .ffvd_oracle import Problem

DATASETS = ("dryer", "drive", "gas_furnace", "actuator", "flutter", "ballbeam")   # FFVD_Main.py:383
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def create_dataset(name: str, data_dir: str):
    """`FFVD_Main.py:134-171`: returns (Y_train (T,1), Y_test, control_inputs (2T,1))."""
    import pandas as pd
    import scipy.io
    if name in ("ballbeam", "dryer", "flutter"):
        data = pd.read_csv(os.path.join(data_dir, name + ".dat"), sep="\t", header=None)
        xx = data.values[:, 0][:, None]
        observations = data.values[:, 1][:, None]
    elif name == "actuator":
        mat = scipy.io.loadmat(os.path.join(data_dir, "actuator.mat"))
        xx, observations = mat["u"], mat["p"]
    elif name == "drive":
        mat = scipy.io.loadmat(os.path.join(data_dir, "drive.mat"))
        xx, observations = mat["u1"], mat["z1"]
    elif name == "gas_furnace":
        data = pd.read_csv(os.path.join(data_dir, "gas_furnace.csv"), sep=",", header=0)
        xx = data.values[:, 0][:, None]
        observations = data.values[:, 1][:, None]
    else:
        raise ValueError(name)
    xx = np.asarray(xx, dtype=np.float64)
    observations = np.asarray(observations, dtype=np.float64)
    control_inputs = (xx - np.mean(xx)) / np.std(xx)
    lens = observations.shape[0]
    Y_train_std = np.std(observations[: int(lens / 2)])
    Y_train_mean = np.mean(observations[: int(lens / 2)])
    observations = (observations - Y_train_mean) / Y_train_std
    return observations[: int(lens / 2)], observations[int(lens / 2):], control_inputs


def init_files(name: str, ref_root: str) -> List[str]:
    # sorted: the reference's glob order is unspecified
    return sorted(glob.glob(os.path.join(ref_root, "Factnonlin_ini", "factnonlin_initialized_10000_" + name + "*.npz")))


def problem_from_npz(fn, Y_train, control_inputs, name="", n_extra_samples: int = 0) -> Tuple[Problem, np.ndarray]:
    """npz -> Problem following `FFVD_Main.py:212-254,340`.  Also returns up to
    ``n_extra_samples`` un-averaged trajectories (S,T+1,D) for S>1 tests."""
    f = np.load(fn, allow_pickle=True)
    T = Y_train.shape[0]
    xs = f["x_samples_training"]                      # (T, 100, D)
    D = xs.shape[2]
    X = np.zeros((T + 1, D))
    X[0] = f["qx1_mu_ini"]                             # dgp_model.py:56-58
    X[1:] = np.mean(xs, axis=1)                       # FFVD_Main.py:226
    prob = Problem(
        X=X,
        Z=np.array(f["Z_val"], dtype=np.float64),                      # :224,340
        U=np.array(f["Umu_ini"].T, dtype=np.float64),                  # :215,253
        logv=np.log(f["kernel_variance"]).astype(np.float64),          # models.py:58
        logl=np.log(f["kernel_lengthscales"]).astype(np.float64),
        logQ=2.0 * np.log(f["Q_sqrt_ini"]).astype(np.float64),         # dgp_model.py:182
        C=np.array(f["C_val"].T, dtype=np.float64),                    # FFVD_Main.py:245
        d=np.array(f["d_val"], dtype=np.float64),
        logR=np.log(f["R_chol_val"]).astype(np.float64),               # likelihoods.py:54
        Y=np.array(Y_train, dtype=np.float64),
        ctrl=np.array(control_inputs[:T], dtype=np.float64),           # dgp_model.py:255
        name=name,
    )
    extra = np.zeros((n_extra_samples, T + 1, D))
    for s in range(n_extra_samples):
        extra[s, 0] = f["qx1_mu_ini"]
        extra[s, 1:] = xs[:, s, :]
    return prob, extra


def load_reference_problem(name: str, index: int = 0, ref_root: str = "/root/reference", n_extra_samples: int = 0):
    Y_train, _, ctrl = create_dataset(name, os.path.join(ref_root, "data"))
    fn = init_files(name, ref_root)[index]
    return problem_from_npz(fn, Y_train, ctrl, name="%s/%d" % (name, index), n_extra_samples=n_extra_samples)


# ---------------------------------------------------------------- packed copy
_PACK_KEYS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")


def load_packed(path: str = None) -> Dict[str, object]:
    """Returns {"problems": [Problem...], "extra": {name: (S,T+1,D)}} from fixtures.npz."""
    path = path or os.path.join(GOLDEN_DIR, "fixtures.npz")
    z = np.load(path, allow_pickle=False)
    names = [str(n) for n in z["names"]]
    problems = []
    for i, nm in enumerate(names):
        ds = nm.split("/")[0]
        kw = {k: np.array(z["%s__%s" % (nm, k)], dtype=np.float64) for k in _PACK_KEYS}
        problems.append(Problem(Y=np.array(z["data__%s__Y" % ds], dtype=np.float64),
                                ctrl=np.array(z["data__%s__ctrl" % ds], dtype=np.float64), name=nm, **kw))
    extra = {str(k)[len("extra__"):]: np.array(z[k], dtype=np.float64) for k in z.files if k.startswith("extra__")}
    return {"problems": problems, "extra": extra}


def synthetic_problem(T: int, M: int, D: int, S: int = 1, n_ctrl: int = 1, Dy: int = 1, seed: int = 20230209,
                      kind: int = 0) -> Problem:
    """Synthetic GPSSM inputs, SURVEY section 8(d) recipe (AR(1) trajectories etc.)."""
    rng = np.random.default_rng(seed)
    Din = D + n_ctrl
    ctrl = rng.standard_normal((T, n_ctrl))
    X = np.zeros((S, T + 1, D))
    X[:, 0] = rng.standard_normal((S, D))
    eps = rng.standard_normal((S, T, D)) * np.sqrt(1 - 0.95 ** 2)
    for t in range(T):
        X[:, t + 1] = 0.95 * X[:, t] + eps[:, t]
    Z = rng.standard_normal((M, Din)) * 1.5
    logl = np.log(rng.uniform(1.0, 4.0, (D, Din)))
    logv = np.log(rng.uniform(0.05, 0.8, D))
    logQ = 2.0 * np.log(rng.uniform(0.2, 0.8, D))
    logR = np.log(np.full((Dy, Dy), 0.4))
    C = rng.standard_normal((D, Dy)) * 0.3
    d = np.zeros(Dy)
    U = rng.standard_normal((M, D))
    Y = X[0, 1:] @ C + d + 0.4 * rng.standard_normal((T, Dy))
    if kind == 1:
        logv = np.zeros(D)
        logl = None
    return Problem(X=X if S > 1 else X[0], Z=Z, U=U, logv=logv, logl=logl, logQ=logQ, C=C, d=d, logR=logR,
                   Y=Y, ctrl=ctrl, kind=kind, name="synthetic_T%d_M%d_D%d_S%d" % (T, M, D, S))
