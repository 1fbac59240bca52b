"""Array plumbing shared by the operator mirrors: contexts per device and output allocation.
PyTorch is used only for device memory and streams; NumPy arrays are accepted and returned
as NumPy (the C ABI stages them with explicit copies)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import _capi

_CONTEXTS: Dict[int, _capi.Context] = {}


def is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def context_for(x=None, device: Optional[int] = None) -> _capi.Context:
    """Context bound to the device of ``x`` (torch CUDA tensor) or to ``device`` (default 0).
    The context enqueues on torch's current stream of that device so that torch ops and
    ffvd kernels are ordered without extra synchronisation."""
    import torch
    if device is None:
        device = x.device.index if (x is not None and is_torch(x) and x.is_cuda) else torch.cuda.current_device() if torch.cuda.is_available() else 0
    stream = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else None
    key = (device, stream)
    ctx = _CONTEXTS.get(key)
    if ctx is None:
        ctx = _capi.Context(device, stream)
        _CONTEXTS[key] = ctx
    return ctx


def as_f64(x):
    """float64, C-contiguous view/copy of x in its own array library (None passes through)."""
    if x is None:
        return None
    if is_torch(x):
        import torch
        return x.detach().to(torch.float64).contiguous()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def empty_like_lib(ref, shape):
    if is_torch(ref):
        import torch
        return torch.empty(tuple(shape), dtype=torch.float64, device=ref.device)
    return np.empty(tuple(shape), dtype=np.float64)


def to_lib(ref, value):
    """value (python scalar / numpy) as a float64 array in ref's library and device."""
    if is_torch(ref):
        import torch
        if is_torch(value):
            return value.detach().to(dtype=torch.float64, device=ref.device).contiguous()
        return torch.as_tensor(np.asarray(value, dtype=np.float64), device=ref.device)
    if is_torch(value):
        return value.detach().cpu().numpy().astype(np.float64)
    return np.ascontiguousarray(np.asarray(value, dtype=np.float64))
