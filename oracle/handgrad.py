"""Hand-derived gradients of the FFVD nll, in numpy.  TEST INFRASTRUCTURE ONLY.

This is the *algorithm the CUDA path implements* (DESIGN.md section "Math"), written
densely in numpy so that the derivation can be checked on the CPU against the autograd
oracle (`oracle/ffvd_oracle.py`, which follows the reference graph op for op) and
against finite differences -- an independent cross-check of both.

Notation (per output dim d): L = chol(Kzz + jI), a_t = L^{-1} k_t, J = objective to
maximise, nll = -(J + priors)/T.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from .ffvd_oracle import Problem, TERM_NAMES


def _kern(kind, X, Z, logv, logl):
    v = math.exp(logv)
    if kind == 0:
        il = np.exp(-logl)
        diff = X[:, None, :] * il - Z[None, :, :] * il
        return v * np.exp(-0.5 * np.sum(diff * diff, axis=2))
    return v * (X @ Z.T)


def _kzz_backward(kind, Kbar, Z, logv, logl):
    """dJ/dZ, dJ/dlogl, dJ/dlogv from a symmetric adjoint Kbar of Kzz (no jitter)."""
    M, Din = Z.shape
    Kzz = _kern(kind, Z, Z, logv, logl)
    Wz = Kbar * Kzz
    if kind == 0:
        il2 = np.exp(-2.0 * logl)
        diff = Z[:, None, :] - Z[None, :, :]                     # (m,n,j)
        gZ = -2.0 * np.einsum("mn,mnj->mj", Wz, diff) * il2
        gl = np.einsum("mn,mnj->j", Wz, diff * diff) * il2
        gv = np.sum(Wz)
        return gZ, gl, gv
    v = math.exp(logv)
    gZ = 2.0 * v * (Kbar @ Z)
    return gZ, None, np.sum(Wz)


def nll_and_grads(prob: Problem, *, collapsed: bool, jitter: float = 1e-5, prior_once: bool = False) -> Dict[str, np.ndarray]:
    X = np.asarray(prob.X, dtype=np.float64)
    single = X.ndim == 2
    Xs = X[None] if single else X
    S, T1, D = Xs.shape
    T = T1 - 1
    Z, U = prob.Z, prob.U
    M, Din = Z.shape
    nc = Din - D
    kind = prob.kind
    Q = np.exp(prob.logQ)
    R = np.exp(prob.logR[0])
    Dy = prob.Y.shape[1]
    ctrl = prob.ctrl[:T]
    log005 = math.log(0.05)

    # raw J-gradients (sum over samples)
    gX = np.zeros_like(Xs)
    gZ = np.zeros((M, Din)); gU = np.zeros((M, D)); gv = np.zeros(D)
    gl = np.zeros((D, Din)) if kind == 0 else None
    gQ = np.zeros(D); gC = np.zeros((D, Dy)); gd = np.zeros(Dy); gR = np.zeros_like(prob.logR)
    terms = np.zeros((S, len(TERM_NAMES)))

    # ---- per output dim prep: L, L^{-1}
    Linv = []
    for d in range(D):
        Kzz = _kern(kind, Z, Z, prob.logv[d], None if kind else prob.logl[d]) + jitter * np.eye(M)
        L = np.linalg.cholesky(Kzz)
        Linv.append(np.linalg.solve(L, np.eye(M)))

    # ---- emission (likelihoods.py:76-79,96-111) per sample
    for s in range(S):
        Xn = Xs[s, 1:]
        res = (prob.Y[:T] - (Xn @ prob.C + prob.d)) / R[None, :]
        terms[s, 1] = -(np.sum(-0.5 * np.sum(res * res, axis=1)) - T * np.sum(np.log(R))) / T
        dyhat = res / R[None, :]
        gX[s, 1:] += dyhat @ prob.C.T
        gC += Xn.T @ dyhat
        gd += dyhat.sum(axis=0)
        gR[0] += np.sum(res * res, axis=0) - T

    for d in range(D):
        logl_d = None if kind else prob.logl[d]
        v = math.exp(prob.logv[d])
        Li = Linv[d]
        ubar = np.zeros(M)
        Ssum = np.zeros((M, M))
        bsum = np.zeros((S, M))
        kd_sum = 0.0
        Ks, As = [], []
        for s in range(S):
            Xc = np.concatenate([Xs[s, :T], ctrl], axis=1) if nc else Xs[s, :T]
            K = _kern(kind, Xc, Z, prob.logv[d], logl_d)
            A = K @ Li.T                                         # rows a_t
            Ks.append((Xc, K)); As.append(A)
        if not collapsed:
            u = U[:, d]
            G = np.zeros((M, M))
            for s in range(S):
                Xc, K = Ks[s]; A = As[s]
                kdiag = np.full(T, v) if kind == 0 else v * np.sum(Xc * Xc, axis=1)
                mu = Xs[s, :T, d] + A @ u
                sig2 = kdiag - np.sum(A * A, axis=1)
                r = Xs[s, 1:, d] - mu
                e = r / Q[d]
                terms[s, 2] += -np.sum(-0.5 * r * r / Q[d] - 0.5 * prob.logQ[d]) / T
                terms[s, 3] += -np.sum(-0.5 * sig2 / Q[d]) / T
                gX[s, :T, d] += e
                gX[s, 1:, d] -= e
                gQ[d] += np.sum(0.5 * r * r / Q[d] - 0.5 + 0.5 * sig2 / Q[d])
                ubar += A.T @ e
                Ssum += A.T @ A
                Kbar_t = (np.outer(e, u) + A / Q[d]) @ Li        # k-bar rows = L^{-T} a-bar
                _accum_kxz(kind, Kbar_t, K, Xc, Z, prob.logv[d], logl_d, gX[s], gZ, gl, gv, d, D)
                # Kdiag contributions
                if kind == 0:
                    gv[d] += -0.5 * T * v / Q[d]
                else:
                    gv[d] += -0.5 * np.sum(kdiag) / Q[d]
                    gX[s, :T, :D] += (-v / Q[d]) * Xc[:, :D]
            gU[:, d] += ubar
            G = np.outer(u, ubar) + Ssum / Q[d]
        else:
            # collapsed bound (cmo:230-257): pass 1 sums, small solve, pass 2 with N
            H_list = []
            for s in range(S):
                A = As[s]
                delta = Xs[s, 1:, d] - Xs[s, :T, d]
                Ss = A.T @ A
                H = Ss / Q[d] + np.eye(M)
                b = (A.T @ delta) / Q[d]
                Hinv = np.linalg.inv(H)
                c = Hinv @ b
                sign, logdet = np.linalg.slogdet(H)
                Xc, K = Ks[s]
                kdiag = np.full(T, v) if kind == 0 else v * np.sum(Xc * Xc, axis=1)
                sig2 = kdiag - np.sum(A * A, axis=1)
                terms[s, 4] += 0.5 * logdet / T
                terms[s, 5] += -0.5 * (b @ c) / T
                terms[s, 3] += 0.5 * np.sum(sig2) / Q[d] / T
                terms[s, 2] += -np.sum(-0.5 * delta * delta / Q[d] - 0.5 * prob.logQ[d]) / T
                Mat = (np.eye(M) - Hinv - np.outer(c, c)) / Q[d]          # F-bar = F Mat + delta c^T/Q
                N = Li.T @ Mat @ Li
                w = (Li.T @ c) / Q[d]
                Kbar_t = K @ N + np.outer(delta, w)
                _accum_kxz(kind, Kbar_t, K, Xc, Z, prob.logv[d], logl_d, gX[s], gZ, gl, gv, d, D)
                dbar = K @ w                                              # = F c / Q
                gX[s, 1:, d] += dbar - delta / Q[d]
                gX[s, :T, d] -= dbar - delta / Q[d]
                if kind == 0:
                    gv[d] += -0.5 * T * v / Q[d]
                else:
                    gv[d] += -0.5 * np.sum(kdiag) / Q[d]
                    gX[s, :T, :D] += (-v / Q[d]) * Xc[:, :D]
                # dJ/dlogQ = Q dJ/dQ
                dJdQ = (0.5 * (M - np.trace(Hinv)) / Q[d] - (c @ b) / Q[d] + 0.5 * (c @ ((H - np.eye(M)) @ c)) / Q[d]
                        + 0.5 * np.sum(sig2) / Q[d] ** 2 + 0.5 * np.sum(delta * delta) / Q[d] ** 2 - 0.5 * T / Q[d])
                gQ[d] += Q[d] * dJdQ
                # G = sum_t fbar_t f_t^T = Mat S + c b^T  (b already has the 1/Q)
                H_list.append(Mat @ Ss + np.outer(c, b))
            G = sum(H_list)
        # ---- Cholesky backward: Kbar_zz = -1/2 L^{-T} (tril(G) + tril(G)^T - diag(G)) L^{-1}
        Gs = np.tril(G) + np.tril(G, -1).T
        Kbar_zz = -0.5 * Li.T @ Gs @ Li
        dZ, dl, dv = _kzz_backward(kind, Kbar_zz, Z, prob.logv[d], logl_d)
        gZ += dZ
        gv[d] += dv
        if kind == 0:
            gl[d] += dl

    # ---- priors and final scaling: nll = -(J + priors)/T
    npri = 1.0 if prior_once else float(S)
    res = {}
    pz = -0.5 * np.sum(Z * Z) if prob.prior_type == "normal" else 0.0
    ph = (-0.5 * np.sum(prob.logl ** 2) if kind == 0 else 0.0) - 0.5 * np.sum((prob.logv - log005) ** 2)
    pu = 0.0 if collapsed else -0.5 * np.sum(U * U)
    hyp = -0.5 * (np.sum(prob.logQ ** 2) + np.sum(prob.C ** 2) + np.sum(prob.d ** 2) + np.sum(prob.logR ** 2))
    for s in range(S):
        px0 = -0.5 * np.sum(Xs[s, 0] ** 2)
        terms[s, 0] = -(pu + ph + pz + px0 + hyp) / T
        gX[s, 0] += -Xs[s, 0]
    nll = terms.sum(axis=1)
    sc = -1.0 / T
    res["nll"] = nll
    res["terms"] = terms
    res["g_X"] = sc * gX
    res["g_Z"] = sc * (gZ + npri * (-Z if prob.prior_type == "normal" else 0.0))
    res["g_U"] = sc * (gU + (0.0 if collapsed else npri * (-U)))
    res["g_logv"] = sc * (gv + npri * (-(prob.logv - log005)))
    if kind == 0:
        res["g_logl"] = sc * (gl + npri * (-prob.logl))
    res["g_logQ"] = sc * (gQ + npri * (-prob.logQ))
    res["g_C"] = sc * (gC + npri * (-prob.C))
    res["g_d"] = sc * (gd + npri * (-prob.d))
    res["g_logR"] = sc * (gR + npri * (-prob.logR))
    if single:
        res["nll"] = res["nll"][0]; res["terms"] = res["terms"][0]; res["g_X"] = res["g_X"][0]
    return res


def _accum_kxz(kind, Kbar_t, K, Xc, Z, logv, logl, gX_s, gZ, gl, gv, d, D):
    """Back-propagate k-bar (T x M) through K(Xc, Z)."""
    T = Xc.shape[0]
    if kind == 0:
        il2 = np.exp(-2.0 * logl)
        W = Kbar_t * K
        rs = W.sum(axis=1); cs = W.sum(axis=0)
        WZ = W @ Z; WtX = W.T @ Xc
        xbar = -(Xc * rs[:, None] - WZ) * il2
        zbar = (WtX - cs[:, None] * Z) * il2
        gX_s[:T, :D] += xbar[:, :D]
        gZ += zbar
        gl[d] += -np.sum(Xc * xbar, axis=0) - np.sum(Z * zbar, axis=0)
        gv[d] += W.sum()
    else:
        v = math.exp(logv)
        xbar = v * (Kbar_t @ Z)
        gX_s[:T, :D] += xbar[:, :D]
        gZ += v * (Kbar_t.T @ Xc)
        gv[d] += np.sum(Kbar_t * K)
