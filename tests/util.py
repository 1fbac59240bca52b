import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(ref, got):
    """max-norm relative error (tolerance is relative to each tensor's max-norm)."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    assert ref.shape == got.shape, (ref.shape, got.shape)
    if not np.all(np.isfinite(got)):
        return float("inf")
    return float(np.max(np.abs(ref - got)) / max(np.max(np.abs(ref)), 1e-300)) if ref.size else 0.0


def assert_close(ref, got, tol, what=""):
    e = relerr(ref, got)
    assert e <= tol, "%s: max-norm relative error %.3e > %.1e" % (what, e, tol)


def load_golden():
    return np.load(os.path.join(GOLDEN, "reference_shim_golden.npz"), allow_pickle=False)
