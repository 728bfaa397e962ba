"""Device-side plumbing: torch tensors as HBM buffers, streams, and the gsb_ctx cache.

torch is used for memory and streams only; every computation is a libgsb200 kernel.
"""
from __future__ import annotations

import ctypes
import threading
from ctypes import POINTER, c_double, c_void_p

import numpy as np

from . import _lib

_dp = POINTER(c_double)
_PINNED_STAGE_MAX = 8 << 20


def torch_mod():
    import torch

    return torch


def current_device() -> int:
    _lib.require_device()
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise _lib.GsbError("torch sees no CUDA device: scpn_fusion_core_b200 has no CPU fallback")
    return torch.cuda.current_device()


def stream_ptr() -> c_void_p:
    return c_void_p(torch_mod().cuda.current_stream().cuda_stream)


def to_device(a, device: int):
    """float64 C-contiguous device tensor from numpy / torch input (copy)."""
    torch = torch_mod()
    if isinstance(a, torch.Tensor):
        return a.to(device=f"cuda:{device}", dtype=torch.float64).contiguous().clone()
    h = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    t = torch.from_numpy(h)
    if 0 < h.nbytes <= _PINNED_STAGE_MAX:
        # per-sample parameter arrays (KBs): stage through pinned host memory so the copy is one asynchronous DMA;
        # large fields go straight from the caller's pageable array (pinning them would cost more than it saves)
        return t.pin_memory().to(f"cuda:{device}", non_blocking=True)
    return t.to(f"cuda:{device}")


def empty(shape, device: int, dtype=None):
    torch = torch_mod()
    return torch.empty(shape, dtype=dtype or torch.float64, device=f"cuda:{device}")


def zeros(shape, device: int, dtype=None):
    torch = torch_mod()
    return torch.zeros(shape, dtype=dtype or torch.float64, device=f"cuda:{device}")


def ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr())


def np_ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class Context:
    """Owns one gsb_ctx (grid geometry + workspace for `batch_cap` equilibria on one device)."""

    def __init__(self, nz: int, nr: int, r_row: np.ndarray, z_axis, dr: float, dz: float,
                 batch_cap: int, device: int):
        self.lib = _lib.load()
        self.nz, self.nr, self.dr, self.dz = int(nz), int(nr), float(dr), float(dz)
        self.batch_cap, self.device = int(batch_cap), int(device)
        self.r_row = np.ascontiguousarray(r_row, dtype=np.float64)
        self.z_axis = None if z_axis is None else np.ascontiguousarray(z_axis, dtype=np.float64)
        h = c_void_p()
        rc = self.lib.gsb_create(ctypes.byref(h), self.nz, self.nr, np_ptr(self.r_row),
                                 None if self.z_axis is None else np_ptr(self.z_axis),
                                 self.dr, self.dz, self.batch_cap, self.device)
        _lib.check(rc, "gsb_create")
        self.handle = h

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.gsb_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


_cache: dict = {}
_cache_lock = threading.Lock()


def get_context(nz, nr, r_row, z_axis, dr, dz, batch: int, device: int) -> Context:
    """Cached context keyed by geometry; re-created with a larger workspace when needed."""
    r_row = np.ascontiguousarray(r_row, dtype=np.float64)
    zkey = b"" if z_axis is None else np.ascontiguousarray(z_axis, dtype=np.float64).tobytes()
    key = (int(nz), int(nr), r_row.tobytes(), zkey, float(dr), float(dz), int(device))
    with _cache_lock:
        ctx = _cache.get(key)
        if ctx is None or ctx.batch_cap < batch:
            if ctx is not None:
                ctx.close()
            ctx = Context(nz, nr, r_row, z_axis, dr, dz, max(int(batch), 1), device)
            _cache[key] = ctx
        return ctx


def clear_cache() -> None:
    with _cache_lock:
        for c in _cache.values():
            c.close()
        _cache.clear()
