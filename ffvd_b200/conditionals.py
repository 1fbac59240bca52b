"""Mirror of `vfegpssm/conditionals.py`: the single-kernel GPflow conditional (jitter 1e-7).
It is the only conditional a bare `LinearK` object can go through."""
from __future__ import annotations

from ._tensor import as_f64, context_for, empty_like_lib, to_lib

JITTER = 1e-7          # conditionals.py:101


def conditional(Xnew, X, kern, f, *, full_cov=False, q_sqrt=None, white=False, return_Lm=False):
    """`conditionals.py:69-107`: one kernel shared by the R columns of f (M,R) -> mean, var (N,R)."""
    if full_cov or q_sqrt is not None or return_Lm:
        raise NotImplementedError("full_cov / q_sqrt / return_Lm belong to the prediction path (SURVEY 8f)")
    Xnew, f = as_f64(Xnew), as_f64(f)
    Xs, Zs = kern._slice(Xnew, to_lib(Xnew, as_f64(X)))
    f = to_lib(Xnew, f)
    logv, logl = kern._hyper(Xnew)
    mean = empty_like_lib(Xnew, (Xnew.shape[0], f.shape[1]))
    var = empty_like_lib(Xnew, (Xnew.shape[0], f.shape[1]))
    context_for(Xnew).conditional(kern.kind, True, Xs, Zs, logv, logl, f, None, white, False, JITTER, mean, var)
    return mean, var
