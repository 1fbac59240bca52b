"""ctypes binding of libffvd_b200.so (C ABI in include/ffvd_b200.h).

Tensors cross the boundary as DLPack capsules: anything with ``__dlpack__`` (torch CPU/CUDA
tensors, NumPy arrays, TF tensors via ``tf.experimental.dlpack``) is accepted.  The capsule is
kept alive for the duration of the call and released afterwards (the library only borrows).
There is no CPU fallback: if the shared library is missing or no B200 is visible, calls fail.
"""
from __future__ import annotations

import ctypes
import os
from typing import Any, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFVD_B200_LIB selects another build of the same library (e.g. the -DFFVD_PHASE_TIMING diagnostic build)
LIB_PATH = os.environ.get("FFVD_B200_LIB") or os.path.join(_HERE, "lib", "libffvd_b200.so")

KERNEL_SE = 0
KERNEL_LINEAR = 1
FLAG_PRIOR_Z_NORMAL = 1
FLAG_PRIOR_ONCE = 2
FLAG_NO_SHARED_PRIORS = 16
FLAG_NO_X0_PRIOR = 32
FLAG_REUSE_KZZ = 64
FLAG_NO_GRADS = 4
FLAG_ASYNC = 8
FLAG_COLLAPSED_P1_ONLY = 128
FLAG_COLLAPSED_RESUME = 256
FLAG_NO_REPLICATED = 512
FLAG_DETERMINISTIC = 1024

_PROBLEM_FIELDS = ("X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl")
_OUTPUT_FIELDS = ("nll", "terms", "g_X", "g_Z", "g_U", "g_logv", "g_logl", "g_logQ", "g_C", "g_d", "g_logR")


class FFVDError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("libffvd_b200 status %d: %s" % (status, message))
        self.status = status


class StaleFactorsError(FFVDError):
    """FLAG_REUSE_KZZ was set although Z / logv / logl changed since the factors were built (status -8)."""


class NotPositiveDefinite(FloatingPointError):
    """Raised when a Cholesky factorisation fails (the reference raises InvalidArgumentError)."""

    def __init__(self, pivot: int, message: str):
        super().__init__(message)
        self.pivot = pivot


class _Problem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _PROBLEM_FIELDS]


class _Outputs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _OUTPUT_FIELDS]


_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libffvd_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64
    lib.ffvd_version.restype = ci
    lib.ffvd_status_string.restype = ctypes.c_char_p
    lib.ffvd_status_string.argtypes = [ci]
    lib.ffvd_last_error.restype = ctypes.c_char_p
    lib.ffvd_ctx_create.argtypes = [ci, vp, ctypes.POINTER(vp)]
    lib.ffvd_ctx_destroy.argtypes = [vp]
    lib.ffvd_ctx_synchronize.argtypes = [vp]
    lib.ffvd_ctx_launch_count.argtypes = [vp]
    lib.ffvd_ctx_launch_count.restype = cll
    lib.ffvd_ctx_fused_time.argtypes = [vp, ci, ctypes.POINTER(cd), ctypes.POINTER(cll)]
    lib.ffvd_ctx_fused_time.restype = ci
    lib.ffvd_debug_phase_clocks.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_uint64)]
    lib.ffvd_kernel_K.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    lib.ffvd_kernel_Kdiag.argtypes = [vp, ci, vp, vp, vp, vp]
    lib.ffvd_kernel_pre_cal.argtypes = [vp, ci, vp, vp, vp, cd, vp]
    lib.ffvd_conditional.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, ci, ci, cd, vp, vp]
    lib.ffvd_conditional_ex.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, ci, ci, cd, ci, vp, vp]
    lib.ffvd_conditional_dense.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, ci, ci, cd, vp, vp, vp]
    lib.ffvd_collapse_u_mean.argtypes = [vp, ci, ctypes.POINTER(_Problem), cd, vp, vp]
    lib.ffvd_logdensity_norm_diag.argtypes = [vp, vp, vp, vp, ci, vp]
    lib.ffvd_logdensity_norm.argtypes = [vp, vp, vp, vp, vp]
    lib.ffvd_debug_exp.argtypes = [vp, vp, vp]
    lib.ffvd_comm_unique_id.argtypes = [vp]
    lib.ffvd_comm_init.argtypes = [vp, vp, ci, ci]
    lib.ffvd_comm_destroy.argtypes = [vp]
    lib.ffvd_comm_info.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.ffvd_allreduce_shared.argtypes = [vp, ctypes.POINTER(_Outputs), ci]
    lib.ffvd_allreduce.argtypes = [vp, vp]
    lib.ffvd_graph_capture_begin.argtypes = [vp]
    lib.ffvd_graph_capture_end.argtypes = [vp, ctypes.POINTER(vp)]
    lib.ffvd_graph_launch.argtypes = [vp, vp]
    lib.ffvd_graph_kernel_count.argtypes = [vp]
    lib.ffvd_graph_kernel_count.restype = cll
    lib.ffvd_graph_destroy.argtypes = [vp, vp]
    lib.ffvd_collapsed_stats_shape.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.ffvd_collapsed_stats_allreduce.argtypes = [vp]
    lib.ffvd_collapsed_stats_get.argtypes = [vp, vp, vp]
    lib.ffvd_collapsed_stats_set.argtypes = [vp, vp, vp]
    lib.ffvd_nll_grads_uncollapsed.argtypes = [vp, ci, ctypes.POINTER(_Problem), ci, cd, ctypes.POINTER(_Outputs)]
    lib.ffvd_nll_grads_collapsed.argtypes = [vp, ci, ctypes.POINTER(_Problem), ci, cd, ctypes.POINTER(_Outputs)]
    lib.ffvd_nll_grads_batched.argtypes = [vp, ci, ci, ci, ctypes.POINTER(_Problem), ci, cd, ctypes.POINTER(_Outputs)]
    lib.ffvd_sghmc_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, cd, cd, cd, ci]
    lib.ffvd_adam_update.argtypes = [vp, vp, vp, vp, vp, cd, cd, cd, cd, cll]
    for name in ("ffvd_ctx_create", "ffvd_ctx_destroy", "ffvd_ctx_synchronize", "ffvd_kernel_K", "ffvd_kernel_Kdiag",
                 "ffvd_kernel_pre_cal", "ffvd_conditional", "ffvd_logdensity_norm_diag", "ffvd_nll_grads_uncollapsed",
                 "ffvd_nll_grads_collapsed", "ffvd_nll_grads_batched", "ffvd_sghmc_update", "ffvd_adam_update",
                 "ffvd_collapse_u_mean", "ffvd_debug_phase_clocks", "ffvd_conditional_ex", "ffvd_logdensity_norm", "ffvd_debug_exp", "ffvd_comm_unique_id", "ffvd_comm_init",
                 "ffvd_comm_destroy", "ffvd_comm_info", "ffvd_allreduce_shared", "ffvd_allreduce", "ffvd_collapsed_stats_shape",
                 "ffvd_collapsed_stats_allreduce", "ffvd_collapsed_stats_get", "ffvd_collapsed_stats_set", "ffvd_graph_capture_begin",
                 "ffvd_graph_capture_end", "ffvd_graph_launch", "ffvd_graph_destroy", "ffvd_conditional_dense"):
        getattr(lib, name).restype = ci
    _lib = lib
    return lib


# ---- DLPack capsule plumbing ---------------------------------------------------------------
_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]


class _Borrow:
    """Holds DLPack capsules alive while the C call runs."""

    def __init__(self):
        self.capsules = []

    def ptr(self, obj: Any) -> Optional[int]:
        if obj is None:
            return None
        if not hasattr(obj, "__dlpack__"):
            raise TypeError("expected an object with __dlpack__ (torch tensor, numpy array, ...), got %r" % type(obj))
        cap = obj.__dlpack__()
        self.capsules.append((cap, obj))
        return _PyCapsule_GetPointer(cap, b"dltensor")

    def release(self):
        # un-consumed capsules run the producer's deleter when they are garbage collected
        self.capsules.clear()


def _check(status: int):
    if status == 0:
        return
    lib = load_library()
    msg = lib.ffvd_last_error().decode() or lib.ffvd_status_string(status).decode()
    if status > 0:
        raise NotPositiveDefinite(status, msg)
    if status in (-1, -2, -3, -4):
        raise ValueError("libffvd_b200 status %d: %s" % (status, msg))
    if status == -8:
        raise StaleFactorsError(status, msg)
    raise FFVDError(status, msg)


class PreparedCall:
    """A bound `ffvd_nll_grads_*` call (see `Context.prepare_nll_grads`).  Keeps the DLPack capsules -- and through them
    the tensors -- alive until `close()`."""

    def __init__(self, ctx, kind, collapsed, problems, outputs, flags, jitter):
        self._ctx, self._kind, self._collapsed, self._flags, self._jitter = ctx, int(kind), bool(collapsed), int(flags), float(jitter)
        self._b = _Borrow()
        self._single = isinstance(problems, dict)
        plist = [problems] if self._single else list(problems)
        olist = [outputs] if self._single else list(outputs)
        self.n = len(plist)
        self._PA, self._OA = (_Problem * self.n)(), (_Outputs * self.n)()
        for i in range(self.n):
            ctx._fill(self._b, self._PA[i], _PROBLEM_FIELDS, plist[i])
            ctx._fill(self._b, self._OA[i], _OUTPUT_FIELDS, olist[i])
        self.outputs = outputs

    def run(self):
        lib = self._ctx._lib
        _check(lib.ffvd_nll_grads_batched(self._ctx._h, self._kind, int(self._collapsed), self.n, self._PA, self._flags,
                                          self._jitter, self._OA))
        return self.outputs

    def close(self):
        self._b.release()


class Graph:
    """A captured sequence of evaluations / updates (`Context.capture`); `launch()` replays it as one CUDA-graph launch.
    Keeps the tensors of the captured calls alive."""

    def __init__(self, ctx, handle, keep):
        self._ctx, self._h, self._keep = ctx, handle, keep

    @property
    def kernels(self) -> int:
        return int(self._ctx._lib.ffvd_graph_kernel_count(self._h))

    def launch(self):
        _check(self._ctx._lib.ffvd_graph_launch(self._ctx._h, self._h))

    def close(self):
        if self._h:
            self._ctx._lib.ffvd_graph_destroy(self._ctx._h, self._h)
            self._h = None
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One context per device; work is enqueued on ``stream`` (a raw cudaStream_t handle,
    e.g. ``torch.cuda.current_stream().cuda_stream``) or on a context-owned stream."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = load_library()
        h = ctypes.c_void_p()
        # a raw handle of 0 is CUDA's legacy default stream (torch's default): pass cudaStreamLegacy (0x1),
        # because NULL means "create a stream owned by the context" in the C ABI
        sarg = None if stream is None else ctypes.c_void_p(int(stream) if int(stream) != 0 else 1)
        _check(self._lib.ffvd_ctx_create(int(device), sarg, ctypes.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ffvd_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(self._lib.ffvd_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.ffvd_ctx_launch_count(self._h))

    def fused_time(self, reset: bool = True):
        """(total_ms, launches) of the fused tile kernel since the last reset; synchronises."""
        tot, cnt = ctypes.c_double(0.0), ctypes.c_int64(0)
        _check(self._lib.ffvd_ctx_fused_time(self._h, int(reset), ctypes.byref(tot), ctypes.byref(cnt)))
        return tot.value, cnt.value

    def phase_clocks(self, reset: bool = True):
        """Diagnostic builds only (-DFFVD_PHASE_TIMING): per-phase SM clock totals of the fused kernel."""
        buf = (ctypes.c_uint64 * 16)()
        _check(self._lib.ffvd_debug_phase_clocks(self._h, int(reset), buf))
        return list(buf)

    # ---- operators
    def kernel_K(self, kind, X, X2, logv, logl, out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_kernel_K(self._h, kind, b.ptr(X), b.ptr(X2), b.ptr(logv), b.ptr(logl), b.ptr(out)))
        finally:
            b.release()
        return out

    def kernel_Kdiag(self, kind, X, logv, logl, out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_kernel_Kdiag(self._h, kind, b.ptr(X), b.ptr(logv), b.ptr(logl), b.ptr(out)))
        finally:
            b.release()
        return out

    def kernel_pre_cal(self, kind, Z, logv, logl, jitter, out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_kernel_pre_cal(self._h, kind, b.ptr(Z), b.ptr(logv), b.ptr(logl), float(jitter), b.ptr(out)))
        finally:
            b.release()
        return out

    def conditional(self, kind, shared_kernel, Xnew, Z, logv, logl, f, q_sqrt, white, full_cov, jitter, mean_out, var_out, flags=0):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_conditional_ex(self._h, kind, int(bool(shared_kernel)), b.ptr(Xnew), b.ptr(Z), b.ptr(logv),
                                                 b.ptr(logl), b.ptr(f), b.ptr(q_sqrt), int(bool(white)), int(bool(full_cov)),
                                                 float(jitter), int(flags), b.ptr(mean_out), b.ptr(var_out)))
        finally:
            b.release()
        return mean_out, var_out

    def conditional_dense(self, kind, shared_kernel, Xnew, Z, logv, logl, f, q_sqrt, white, full_cov, jitter, mean_out, var_out, Lm_out=None):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_conditional_dense(self._h, kind, int(bool(shared_kernel)), b.ptr(Xnew), b.ptr(Z), b.ptr(logv), b.ptr(logl),
                                                    b.ptr(f), b.ptr(q_sqrt), int(bool(white)), int(bool(full_cov)), float(jitter),
                                                    b.ptr(mean_out), b.ptr(var_out), b.ptr(Lm_out)))
        finally:
            b.release()
        return mean_out, var_out, Lm_out

    def collapse_u_mean(self, kind: int, problem: dict, U_mean_out, LHinvT_out=None, jitter: float = 1e-5):
        b = _Borrow()
        try:
            P = _Problem()
            self._fill(b, P, _PROBLEM_FIELDS, problem)
            _check(self._lib.ffvd_collapse_u_mean(self._h, kind, ctypes.byref(P), float(jitter), b.ptr(U_mean_out), b.ptr(LHinvT_out)))
        finally:
            b.release()
        return U_mean_out, LHinvT_out

    def logdensity_norm_diag(self, y, ymean, Rchols, vec, out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_logdensity_norm_diag(self._h, b.ptr(y), b.ptr(ymean), b.ptr(Rchols), int(bool(vec)), b.ptr(out)))
        finally:
            b.release()
        return out

    # ---- CUDA graphs
    def capture(self, fn, keep=None) -> "Graph":
        """Run `fn()` (calls of nll_grads / sghmc_update / adam_update on this context, FLAG_ASYNC, device tensors) under
        stream capture and return the instantiated graph.  Make the same calls once BEFORE capturing (workspace layout).
        `keep`: objects (tensors) that must stay alive as long as the graph."""
        _check(self._lib.ffvd_graph_capture_begin(self._h))
        h = ctypes.c_void_p()
        try:
            fn()
        except BaseException:
            self._lib.ffvd_graph_capture_end(self._h, ctypes.byref(h))
            if h:
                self._lib.ffvd_graph_destroy(self._h, h)
            raise
        _check(self._lib.ffvd_graph_capture_end(self._h, ctypes.byref(h)))
        return Graph(self, h, keep)

    # ---- multi-GPU: NCCL communicator behind the C ABI
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL unique id (create on rank 0, ship to the other ranks, pass to `comm_init` everywhere)."""
        buf = ctypes.create_string_buffer(128)
        _check(load_library().ffvd_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        if len(unique_id) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        _check(self._lib.ffvd_comm_init(self._h, ctypes.c_char_p(unique_id), int(rank), int(nranks)))

    def comm_destroy(self):
        _check(self._lib.ffvd_comm_destroy(self._h))

    def comm_info(self):
        r, n, v = ctypes.c_int(0), ctypes.c_int(1), ctypes.c_int(0)
        _check(self._lib.ffvd_comm_info(self._h, ctypes.byref(r), ctypes.byref(n), ctypes.byref(v)))
        return r.value, n.value, v.value

    def allreduce_shared(self, outputs: dict, with_scalars: bool = False):
        """ONE NCCL all-reduce of the packed shared-parameter gradients (in place); g_X stays on its owner."""
        b = _Borrow()
        try:
            O = _Outputs()
            self._fill(b, O, _OUTPUT_FIELDS, {k: v for k, v in outputs.items() if k != "g_X"})
            _check(self._lib.ffvd_allreduce_shared(self._h, ctypes.byref(O), int(bool(with_scalars))))
        finally:
            b.release()
        return outputs

    def allreduce(self, tensor):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_allreduce(self._h, b.ptr(tensor)))
        finally:
            b.release()
        return tensor

    # ---- collapsed bound under time sharding: statistics of a pending FLAG_COLLAPSED_P1_ONLY evaluation
    def collapsed_stats_shape(self):
        nb, Mp = ctypes.c_int(0), ctypes.c_int(0)
        _check(self._lib.ffvd_collapsed_stats_shape(self._h, ctypes.byref(nb), ctypes.byref(Mp)))
        return nb.value, Mp.value

    def collapsed_stats_allreduce(self):
        _check(self._lib.ffvd_collapsed_stats_allreduce(self._h))

    def collapsed_stats_get(self, S_out, b_out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_collapsed_stats_get(self._h, b.ptr(S_out), b.ptr(b_out)))
        finally:
            b.release()
        return S_out, b_out

    def collapsed_stats_set(self, S_in, b_in):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_collapsed_stats_set(self._h, b.ptr(S_in), b.ptr(b_in)))
        finally:
            b.release()

    def debug_exp(self, x, out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_debug_exp(self._h, b.ptr(x), b.ptr(out)))
        finally:
            b.release()
        return out

    def logdensity_norm(self, y, ymean, Rchols, out):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_logdensity_norm(self._h, b.ptr(y), b.ptr(ymean), b.ptr(Rchols), b.ptr(out)))
        finally:
            b.release()
        return out

    def _fill(self, b: _Borrow, struct, fields, d: dict):
        for n in fields:
            p = b.ptr(d.get(n))
            setattr(struct, n, p)

    def nll_grads(self, kind: int, collapsed: bool, problem: dict, outputs: dict, flags: int = FLAG_PRIOR_Z_NORMAL,
                  jitter: float = 1e-5):
        b = _Borrow()
        try:
            P, O = _Problem(), _Outputs()
            self._fill(b, P, _PROBLEM_FIELDS, problem)
            self._fill(b, O, _OUTPUT_FIELDS, outputs)
            fn = self._lib.ffvd_nll_grads_collapsed if collapsed else self._lib.ffvd_nll_grads_uncollapsed
            _check(fn(self._h, kind, ctypes.byref(P), int(flags), float(jitter), ctypes.byref(O)))
        finally:
            b.release()
        return outputs

    def prepare_nll_grads(self, kind: int, collapsed: bool, problems, outputs, flags: int = FLAG_PRIOR_Z_NORMAL,
                          jitter: float = 1e-5) -> "PreparedCall":
        """Bind the tensors of one (dict) or many (sequence of dicts) problems ONCE; `PreparedCall.run()` then re-issues
        the evaluation without re-exporting ~22 DLPack capsules per problem (the capsules describe memory, not values:
        in-place updates of the tensors are seen).  For the 95-chain batch of BASELINE config 4 the per-call export cost
        (~4 ms of Python) exceeded the 1.9 ms of device work."""
        return PreparedCall(self, kind, collapsed, problems, outputs, flags, jitter)

    def nll_grads_batched(self, kind: int, collapsed: bool, problems: Sequence[dict], outputs: Sequence[dict],
                          flags: int = FLAG_PRIOR_Z_NORMAL, jitter: float = 1e-5):
        n = len(problems)
        b = _Borrow()
        try:
            PA, OA = (_Problem * n)(), (_Outputs * n)()
            for i in range(n):
                self._fill(b, PA[i], _PROBLEM_FIELDS, problems[i])
                self._fill(b, OA[i], _OUTPUT_FIELDS, outputs[i])
            _check(self._lib.ffvd_nll_grads_batched(self._h, kind, int(bool(collapsed)), n, PA, int(flags), float(jitter), OA))
        finally:
            b.release()
        return outputs

    def sghmc_update(self, theta, grad, noise, xi, g, g2, p, epsilon, mdecay, X_N, burn_in):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_sghmc_update(self._h, b.ptr(theta), b.ptr(grad), b.ptr(noise), b.ptr(xi), b.ptr(g), b.ptr(g2),
                                               b.ptr(p), float(epsilon), float(mdecay), float(X_N), int(bool(burn_in))))
        finally:
            b.release()

    def adam_update(self, theta, grad, m, v, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=1):
        b = _Borrow()
        try:
            _check(self._lib.ffvd_adam_update(self._h, b.ptr(theta), b.ptr(grad), b.ptr(m), b.ptr(v), float(lr), float(beta1),
                                              float(beta2), float(eps), int(step)))
        finally:
            b.release()
