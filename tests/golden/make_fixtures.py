"""Pack the hot-path inputs of the reference's 95 warm starts + 6 datasets into fixtures.npz.

Run in the build container (needs /root/reference):  python tests/golden/make_fixtures.py
Only the fields the hot path consumes are kept (FFVD_Main.py:212-254 mapping, see
oracle/fixtures.py); x_samples_training is reduced to its mean trajectory plus, for the first
init of each dataset, 4 un-averaged sample trajectories for S>1 tests.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fixtures  # noqa: E402

REF = "/root/reference"


def main():
    out = {}
    names = []
    for ds in fixtures.DATASETS:
        Y_train, _, ctrl = fixtures.create_dataset(ds, os.path.join(REF, "data"))
        T = Y_train.shape[0]
        out["data__%s__Y" % ds] = Y_train
        out["data__%s__ctrl" % ds] = ctrl[:T]
        for i, fn in enumerate(fixtures.init_files(ds, REF)):
            prob, extra = fixtures.problem_from_npz(fn, Y_train, ctrl, name="%s/%d" % (ds, i), n_extra_samples=4 if i == 0 else 0)
            names.append(prob.name)
            for k in fixtures._PACK_KEYS:
                out["%s__%s" % (prob.name, k)] = getattr(prob, k)
            if i == 0:
                out["extra__%s" % prob.name] = extra
    out["names"] = np.array(names)
    path = os.path.join(fixtures.GOLDEN_DIR, "fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB,", len(names), "problems")


if __name__ == "__main__":
    main()
