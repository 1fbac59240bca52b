"""Mirror of the data handling in `FFVD_Main.py`: `create_dataset` (:134-171) and the mapping from a
`Factnonlin_ini/*.npz` warm start to the model arguments (:212-254, 340).  Host-side NumPy only."""
from __future__ import annotations

import glob
import os

import numpy as np

DATASETS = ('dryer', 'drive', 'gas_furnace', 'actuator', 'flutter', 'ballbeam')      # FFVD_Main.py:383


def create_dataset(file_path, data_dir='data/'):
    """`FFVD_Main.py:134-171`.  `file_path` is the reference's 'name/' string.  Returns (Y_train, Y_test,
    control_inputs, Y_train_std, Y_train_mean, control_inputs_mean, control_inputs_std): inputs standardised over
    the whole series, observations by the first half, train = first half."""
    import pandas as pd
    import scipy.io
    name = file_path[:-1] if file_path.endswith('/') else file_path
    if name in ('ballbeam', 'dryer', 'flutter'):
        data = pd.read_csv(os.path.join(data_dir, name + '.dat'), sep='\t', header=None)
        xx = data.values[:, 0][:, None]
        observations = data.values[:, 1][:, None]
    elif name == 'actuator':
        mat = scipy.io.loadmat(os.path.join(data_dir, name + '.mat'))
        xx, observations = mat['u'], mat['p']
    elif name == 'drive':
        mat = scipy.io.loadmat(os.path.join(data_dir, name + '.mat'))
        xx, observations = mat['u1'], mat['z1']
    elif name == 'gas_furnace':
        data = pd.read_csv(os.path.join(data_dir, name + '.csv'), sep=',', header=0)
        xx = data.values[:, 0][:, None]
        observations = data.values[:, 1][:, None]
    else:
        raise ValueError("unknown dataset %r" % file_path)
    xx = np.asarray(xx, dtype=np.float64)
    observations = np.asarray(observations, dtype=np.float64)
    control_inputs_mean, control_inputs_std = np.mean(xx), np.std(xx)
    control_inputs = (xx - control_inputs_mean) / control_inputs_std
    lens = observations.shape[0]
    Y_train_std = np.std(observations[:int(lens / 2)])
    Y_train_mean = np.mean(observations[:int(lens / 2)])
    observations = (observations - Y_train_mean) / Y_train_std
    return (observations[:int(lens / 2)], observations[int(lens / 2):], control_inputs, Y_train_std, Y_train_mean,
            control_inputs_mean, control_inputs_std)


def init_files(file_path, root='.'):
    """The warm starts of one dataset, sorted (the reference's glob order is unspecified)."""
    name = file_path[:-1] if file_path.endswith('/') else file_path
    return sorted(glob.glob(os.path.join(root, 'Factnonlin_ini', 'factnonlin_initialized_10000_' + name + '*.npz')))


def arguments_from_factnonlin(factnonlin):
    """`FFVD_Main.py:212-254`: the fields of a `Factnonlin_ini` npz (or a dict with the same keys) the driver feeds to
    `model.ARGS` -- CC (D,Dy), DD, QQ_chol, RR_chol, lengthscales, variance, UU_ini (M,D), XX_0_ini,
    x_initialization (T,D), ZZ (M,Din)."""
    f = factnonlin
    return dict(CC=np.asarray(f['C_val'], dtype=np.float64).T, DD=np.asarray(f['d_val'], dtype=np.float64),
                QQ_chol=np.asarray(f['Q_sqrt_ini'], dtype=np.float64), RR_chol=np.asarray(f['R_chol_val'], dtype=np.float64),
                lengthscales=np.asarray(f['kernel_lengthscales'], dtype=np.float64),
                variance=np.asarray(f['kernel_variance'], dtype=np.float64), UU_ini=np.asarray(f['Umu_ini'], dtype=np.float64).T,
                XX_0_ini=np.asarray(f['qx1_mu_ini'], dtype=np.float64),
                x_initialization=np.mean(np.asarray(f['x_samples_training'], dtype=np.float64), axis=1),
                ZZ=np.asarray(f['Z_val'], dtype=np.float64))


# FFVD_Main.py:273-324: case_val -> (kernel_optimization, U_optimization, Z_optimization, U_collapse, X_PG)
CASE_TABLE = {1: (True, True, True, False, False), 2: (False, False, True, False, False), 3: (False, False, False, False, False),
              4: (True, False, True, True, False), 5: (False, False, True, True, False), 6: (True, True, True, False, True),
              7: (False, False, False, False, False)}


def load_packed_problems(path):
    """The packed copy of the 95 bundled warm starts x 6 datasets (`tests/golden/fixtures.npz`, written by
    `tests/golden/make_fixtures.py` with the `FFVD_Main.py:212-254` mapping) as a list of problem dicts
    {X, Z, U, logv, logl, logQ, C, d, logR, Y, ctrl} of NumPy arrays -- BASELINE configs[3]'s many-small-chains batch."""
    z = np.load(path, allow_pickle=False)
    out = []
    for nm in [str(n) for n in z["names"]]:
        ds = nm.split("/")[0]
        p = {k: np.array(z["%s__%s" % (nm, k)], dtype=np.float64) for k in ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR")}
        p["Y"] = np.array(z["data__%s__Y" % ds], dtype=np.float64)
        p["ctrl"] = np.array(z["data__%s__ctrl" % ds], dtype=np.float64)
        p["name"] = nm
        out.append(p)
    return out
