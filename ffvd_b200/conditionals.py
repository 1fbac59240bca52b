"""Mirror of `vfegpssm/conditionals.py`: the single-kernel GPflow conditional (jitter 1e-7).
It is the only conditional a bare `LinearK` object can go through."""
from __future__ import annotations

from ._tensor import as_f64, context_for, empty_like_lib, to_lib

JITTER = 1e-7          # conditionals.py:101


def conditional(Xnew, X, kern, f, *, full_cov=False, q_sqrt=None, white=False, return_Lm=False):
    """`conditionals.py:69-107`: one kernel shared by the R columns of f (M,R) -> mean (N,R), var (N,R) or, with
    `full_cov`, (R,N,N).  `q_sqrt`: (M,R) scales or (R,M,M) factors, one per column of f (conditionals.py:46-58), whitened
    or not; `return_Lm=True` also returns chol(K(X,X) + 1e-7 I) (:60-66).
    The hot-path branch (diagonal variances, no q_sqrt or a whitened one) runs in the fused tile kernel; `full_cov`,
    `return_Lm` and q_sqrt with white=False go through the op-for-op dense path (ffvd_conditional_dense), meant for a
    handful of prediction points."""
    Xnew, f = as_f64(Xnew), as_f64(f)
    Xs, Zs = kern._slice(Xnew, to_lib(Xnew, as_f64(X)))
    f = to_lib(Xnew, f)
    logv, logl = kern._hyper(Xnew)
    q = None if q_sqrt is None else to_lib(Xnew, as_f64(q_sqrt))
    if q is not None and q.ndim not in (2, 3):
        raise ValueError("Bad dimension for q_sqrt: %s" % str(q.ndim))       # conditionals.py:55-57
    N, R, M = Xnew.shape[0], f.shape[1], Zs.shape[0]
    mean = empty_like_lib(Xnew, (N, R))
    if full_cov or return_Lm or (q is not None and not white):
        var = empty_like_lib(Xnew, (R, N, N) if full_cov else (N, R))
        Lm = empty_like_lib(Xnew, (1, M, M)) if return_Lm else None
        context_for(Xnew).conditional_dense(kern.kind, True, Xs, Zs, logv, logl, f, q, white, full_cov, JITTER, mean, var, Lm)
        return (mean, var, Lm[0]) if return_Lm else (mean, var)
    var = empty_like_lib(Xnew, (N, R))
    context_for(Xnew).conditional(kern.kind, True, Xs, Zs, logv, logl, f, q, white, False, JITTER, mean, var)
    return mean, var
