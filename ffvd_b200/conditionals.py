"""Mirror of `vfegpssm/conditionals.py`: the single-kernel GPflow conditional (jitter 1e-7).
It is the only conditional a bare `LinearK` object can go through."""
from __future__ import annotations

from ._tensor import as_f64, context_for, empty_like_lib, to_lib

JITTER = 1e-7          # conditionals.py:101


def conditional(Xnew, X, kern, f, *, full_cov=False, q_sqrt=None, white=False, return_Lm=False):
    """`conditionals.py:69-107`: one kernel shared by the R columns of f (M,R) -> mean, var (N,R).
    `q_sqrt`: (M,R) scales or (R,M,M) factors, one per column of f (conditionals.py:46-58); needs white=True."""
    if full_cov or return_Lm:
        raise NotImplementedError("full_cov / return_Lm are not built (the driver runs with full_cov=False, FFVD_Main.py:267)")
    if q_sqrt is not None and not white:
        raise NotImplementedError("q_sqrt with white=False is not built")
    Xnew, f = as_f64(Xnew), as_f64(f)
    Xs, Zs = kern._slice(Xnew, to_lib(Xnew, as_f64(X)))
    f = to_lib(Xnew, f)
    logv, logl = kern._hyper(Xnew)
    q = None if q_sqrt is None else to_lib(Xnew, as_f64(q_sqrt))
    if q is not None and q.ndim not in (2, 3):
        raise ValueError("Bad dimension for q_sqrt: %s" % str(q.ndim))       # conditionals.py:55-57
    mean = empty_like_lib(Xnew, (Xnew.shape[0], f.shape[1]))
    var = empty_like_lib(Xnew, (Xnew.shape[0], f.shape[1]))
    context_for(Xnew).conditional(kern.kind, True, Xs, Zs, logv, logl, f, q, white, False, JITTER, mean, var)
    return mean, var
