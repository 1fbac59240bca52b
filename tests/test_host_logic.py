"""CPU tests of the host-side mirror of the reference interface (no device work)."""
import numpy as np
import pytest

from ffvd_b200 import _capi, distributed
from ffvd_b200.dgp_model import sghmc_variable_names
from ffvd_b200.kernels import LinearK, SquaredExponential as SE_U
from ffvd_b200.kernels_multi_output import SquaredExponential


def test_kernel_constructor_matches_reference_attributes():
    k = SquaredExponential(5, variance=0.3, lengthscales=np.array([1., 2., 3., 4., 5.]), ARD=True, kernel_optimization=True)
    assert k.input_dim == 5 and k.ARD is True and k.trainable
    assert np.isclose(float(k.variance), 0.3) and np.allclose(k.lengthscales, [1, 2, 3, 4, 5])
    assert np.isclose(float(k.logvariance), np.log(0.3))
    k2 = SE_U(2, variance=0.1, lengthscales=1.0, U_kernel_optimization=False)    # kernels.py spelling of the kwarg
    assert k2.ARD is False and k2.loglengthscales.shape == ()
    with pytest.raises(TypeError):
        SquaredExponential(2, U_kernel_optimization=True)


def test_linear_kernel_reference_failure_modes():
    # ARD=False with a vector variance raises exactly like the reference
    with pytest.raises(ValueError, match="shape of variance does not match input_dim"):
        LinearK(5, ARD=False, variance=np.ones(4))
    assert float(LinearK(5).variance) == 1.0
    # Q1(b): a bare kernel object instead of a list is not subscriptable
    from ffvd_b200.dgp_model import Layer
    with pytest.raises(TypeError):
        Layer(None, None, None, None, LinearK(5), 4, 10, False, 4, 10, False)


def test_ard_shape_validation():
    with pytest.raises(ValueError, match="shape of lengthscales does not match input_dim"):
        SquaredExponential(1, lengthscales=np.ones(2))


@pytest.mark.parametrize("case_val,expected", [
    (1, []), (2, ["logv", "logl", "U"]), (3, ["logv", "logl", "U", "Z"]), (4, []), (5, ["logv", "logl"]),
    (6, []), (7, ["U", "X"])])
def test_sghmc_variable_selection(case_val, expected):
    # FFVD_Main.py:273-324 flag table -> dgp_model.py:213-244 selection
    table = {1: (True, True, True, False), 2: (False, False, True, False), 3: (False, False, False, False),
             4: (True, False, True, True), 5: (False, False, True, True), 6: (True, True, True, False),
             7: (False, False, False, False)}
    ko, uo, zo, uc = table[case_val]
    assert sghmc_variable_names(_capi.KERNEL_SE, case_val, ko, True, uo, uc, zo) == expected


def test_shard_helpers():
    for n in (1, 7, 64, 95, 256):
        for w in (1, 2, 4, 8):
            blocks = [distributed.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
            rr = sorted(sum((distributed.round_robin(n, r, w) for r in range(w)), []))
            assert rr == list(range(n))


def test_create_dataset_and_case_table(tmp_path):
    """datasets.create_dataset restates FFVD_Main.py:134-171 (checked against the oracle-side fixture loader on a
    synthetic file in the reference's gas_furnace.csv layout) and CASE_TABLE restates :273-324."""
    from ffvd_b200 import datasets
    from oracle import fixtures
    rng = np.random.default_rng(0)
    raw = np.column_stack([rng.standard_normal(60) * 3 + 1, rng.standard_normal(60) * 2 - 5])
    with open(tmp_path / "gas_furnace.csv", "w") as f:
        f.write("u,y\n" + "\n".join("%.17g,%.17g" % (a, b) for a, b in raw))
    Ytr, Yte, ctrl, ystd, ymean, cmean, cstd = datasets.create_dataset("gas_furnace/", str(tmp_path))
    rYtr, rYte, rctrl = fixtures.create_dataset("gas_furnace", str(tmp_path))
    assert np.array_equal(Ytr, rYtr) and np.array_equal(Yte, rYte) and np.array_equal(ctrl, rctrl)
    assert Ytr.shape == (30, 1) and Yte.shape == (30, 1) and ctrl.shape == (60, 1)
    assert abs(np.mean(Ytr)) < 1e-12 and abs(np.std(Ytr) - 1) < 1e-12 and abs(np.mean(ctrl)) < 1e-12
    assert np.isclose(ystd, np.std(raw[:30, 1])) and np.isclose(ymean, np.mean(raw[:30, 1]))
    with pytest.raises(ValueError):
        datasets.create_dataset("nope/", str(tmp_path))
    expect = {1: [], 2: ["logv", "logl", "U"], 3: ["logv", "logl", "U", "Z"], 4: [], 5: ["logv", "logl"], 6: [], 7: ["U", "X"]}
    for case_val, (ko, uo, zo, uc, pg) in datasets.CASE_TABLE.items():
        assert sghmc_variable_names(_capi.KERNEL_SE, case_val, ko, True, uo, uc, zo) == expect[case_val]
        assert pg == (case_val == 6)
    f = dict(C_val=np.ones((1, 4)), d_val=np.zeros(1), Q_sqrt_ini=np.ones(4), R_chol_val=np.ones((1, 1)), kernel_lengthscales=np.ones((4, 5)),
             kernel_variance=np.ones(4), Umu_ini=np.zeros((4, 100)), qx1_mu_ini=np.zeros(4), x_samples_training=np.zeros((30, 7, 4)),
             Z_val=np.zeros((100, 5)))
    a = datasets.arguments_from_factnonlin(f)
    assert a["CC"].shape == (4, 1) and a["UU_ini"].shape == (100, 4) and a["x_initialization"].shape == (30, 4)


def test_flag_and_status_constants_match_header():
    """The Python constants are the header's (a drifted flag value would silently select another code path)."""
    import os, re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ffvd_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(FFVD_[A-Z0-9_]+)\s+(-?\d+)\b", hdr)}
    pairs = {"FFVD_FLAG_PRIOR_Z_NORMAL": _capi.FLAG_PRIOR_Z_NORMAL, "FFVD_FLAG_PRIOR_ONCE": _capi.FLAG_PRIOR_ONCE,
             "FFVD_FLAG_NO_GRADS": _capi.FLAG_NO_GRADS, "FFVD_FLAG_ASYNC": _capi.FLAG_ASYNC,
             "FFVD_FLAG_NO_SHARED_PRIORS": _capi.FLAG_NO_SHARED_PRIORS, "FFVD_FLAG_NO_X0_PRIOR": _capi.FLAG_NO_X0_PRIOR,
             "FFVD_FLAG_REUSE_KZZ": _capi.FLAG_REUSE_KZZ, "FFVD_FLAG_COLLAPSED_P1_ONLY": _capi.FLAG_COLLAPSED_P1_ONLY,
             "FFVD_FLAG_COLLAPSED_RESUME": _capi.FLAG_COLLAPSED_RESUME, "FFVD_FLAG_NO_REPLICATED": _capi.FLAG_NO_REPLICATED,
             "FFVD_KERNEL_SE": _capi.KERNEL_SE, "FFVD_KERNEL_LINEAR": _capi.KERNEL_LINEAR}
    for name, val in pairs.items():
        assert defs[name] == val, name
    flags = [v for k, v in defs.items() if k.startswith("FFVD_FLAG_")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags)      # distinct single bits
    assert defs["FFVD_E_STALE"] == -8


def test_shared_noise_is_rank_independent_and_step_dependent():
    a = distributed.shared_noise((3, 4), step=5, seed=2)
    b = distributed.shared_noise((3, 4), step=5, seed=2)
    c = distributed.shared_noise((3, 4), step=6, seed=2)
    assert a.dtype.is_floating_point and a.shape == (3, 4)
    assert bool((a == b).all()) and not bool((a == c).all())


def test_oracle_loop_variable_sets_match_the_mirror():
    """oracle/loop.py restates dgp_model.py:213-244 and the trainable flags independently of the product mirror."""
    from oracle import loop
    from ffvd_b200.datasets import CASE_TABLE
    for case_val, (ko, uo, zo, uc, xpg) in CASE_TABLE.items():
        assert loop.CASES[case_val] == (ko, uo, zo, uc)
        assert loop.sampled_names(case_val) == sghmc_variable_names(_capi.KERNEL_SE, case_val, ko, True, uo, uc, zo)
