"""GPU mirror of the reference's ``scpn_fusion.core.multigrid_solve`` free functions.

Same names, argument meaning and error behaviour as
``src/scpn_fusion/core/multigrid_solve.py`` (reference), executed by the sm_100a
kernels of libgsb200 (``csrc/gsb_mg.cu``).  Inputs may be NumPy arrays (results
come back as NumPy, like the reference) or CUDA ``torch`` tensors, optionally
with a leading batch axis ``(B, nz, nr)`` (results stay on the device).

The reference passes a full ``r_grid`` meshgrid; its rows are identical, so the
kernels take one interior row and derive the per-column stencil coefficients
(``multigrid_solve.py:184-189``) on the fly.
"""
from __future__ import annotations

from ctypes import c_void_p

import numpy as np

from . import _device as D
from . import _lib

__all__ = [
    "validate_sor_omega", "restrict_full_weight", "prolongate_bilinear", "mg_smooth", "mg_residual",
    "multigrid_vcycle", "residual_linf", "multigrid_solve",
]


def validate_sor_omega(omega: float) -> float:
    """multigrid_solve.py:33-54."""
    w = float(omega)
    if not np.isfinite(w) or w < 1.0 or w >= 2.0:
        raise ValueError("omega must be finite and satisfy 1.0 <= omega < 2.0")
    return w


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _r_row(r_grid, nz: int, nr: int) -> np.ndarray:
    """One interior row of the reference's r_grid (all interior rows must be identical)."""
    rg = r_grid.detach().cpu().numpy() if _is_torch(r_grid) else np.asarray(r_grid, dtype=np.float64)
    if rg.ndim == 1:
        if rg.shape[0] != nr:
            raise ValueError("r_grid row must have nr entries")
        return np.ascontiguousarray(rg)
    if rg.shape != (nz, nr):
        raise ValueError(f"r_grid must have shape ({nz}, {nr}); got {rg.shape}")
    row = rg[1] if nz > 2 else rg[0]
    if nz > 3 and not np.all(rg[1:-1] == row):
        raise NotImplementedError("r_grid with non-identical rows is not a meshgrid of an R axis")
    return np.ascontiguousarray(row)


class _Fields:
    """Normalises (psi, ...) inputs to device tensors [B, nz, nr] and back."""

    def __init__(self, first, device=None):
        self.torch_in = _is_torch(first)
        if self.torch_in:
            if not first.is_cuda:
                raise ValueError("torch inputs must be CUDA tensors")
            self.device = first.device.index
        else:
            self.device = D.current_device() if device is None else device
        shape = tuple(first.shape)
        if len(shape) == 2:
            self.batched, self.B = False, 1
            self.nz, self.nr = shape
        elif len(shape) == 3:
            self.batched, self.B = True, shape[0]
            self.nz, self.nr = shape[1:]
        else:
            raise ValueError("fields must have shape (nz, nr) or (B, nz, nr)")

    def dev(self, a, copy=True):
        torch = D.torch_mod()
        if _is_torch(a):
            t = a.to(dtype=torch.float64).contiguous()
            if copy and t.data_ptr() == a.data_ptr():
                t = t.clone()
        else:
            t = D.to_device(a, self.device)
        return t.reshape(self.B, self.nz, self.nr) if t.dim() != 3 else t

    def out(self, t):
        if not self.batched:
            t = t.reshape(t.shape[-2], t.shape[-1])
        return t if self.torch_in else t.cpu().numpy()


def restrict_full_weight(fine):
    """multigrid_solve.py:57-99 (9-point full weighting, walls injected)."""
    f = _Fields(fine)
    lib = _lib.load()
    x = f.dev(fine, copy=False)
    nzc, nrc = (f.nz + 1) // 2, (f.nr + 1) // 2
    out = D.empty((f.B, nzc, nrc), f.device)
    _lib.check(lib.gsb_restrict_full_weight(D.ptr(x), D.ptr(out), f.nz, f.nr, f.B, D.stream_ptr()),
               "gsb_restrict_full_weight")
    return f.out(out)


def prolongate_bilinear(coarse, nz_f: int, nr_f: int):
    """multigrid_solve.py:102-145."""
    f = _Fields(coarse)
    lib = _lib.load()
    x = f.dev(coarse, copy=False)
    out = D.empty((f.B, int(nz_f), int(nr_f)), f.device)
    _lib.check(lib.gsb_prolong_bilinear(D.ptr(x), D.ptr(out), f.nz, f.nr, int(nz_f), int(nr_f), f.B,
                                        D.stream_ptr()), "gsb_prolong_bilinear")
    return f.out(out)


def _ctx(f: _Fields, r_grid, dr, dz, z_axis=None):
    return D.get_context(f.nz, f.nr, _r_row(r_grid, f.nz, f.nr), z_axis, float(dr), float(dz), f.B, f.device)


def mg_smooth(psi, source, r_grid, dr: float, dz: float, omega: float, n_sweeps: int, *, fuse: int = 3):
    """multigrid_solve.py:148-208.  NumPy input is updated in place like the reference.

    ``fuse`` (not in the reference) selects the kernel: 0 = one launch per colour pass, 1..3 = that
    many sweeps per pass over HBM (temporally blocked, identical results; default 3)."""
    omega = validate_sor_omega(omega)
    f = _Fields(psi)
    ctx = _ctx(f, r_grid, dr, dz)
    p = f.dev(psi, copy=False)
    s = f.dev(source, copy=False)
    _lib.check(ctx.lib.gsb_smooth_ex(ctx.handle, D.ptr(p), D.ptr(s), f.B, omega, int(n_sweeps), 0, int(fuse),
                                     D.stream_ptr()), "gsb_smooth")
    if f.torch_in:
        if p.data_ptr() != psi.data_ptr():
            psi.copy_(p.reshape(psi.shape))
        return psi
    res = p.cpu().numpy().reshape(np.shape(psi))
    if isinstance(psi, np.ndarray) and psi.dtype == np.float64:
        psi[...] = res
        return psi
    return res


def mg_residual(psi, source, r_grid, dr: float, dz: float):
    """multigrid_solve.py:211-249."""
    f = _Fields(psi)
    ctx = _ctx(f, r_grid, dr, dz)
    p, s = f.dev(psi, copy=False), f.dev(source, copy=False)
    out = D.empty((f.B, f.nz, f.nr), f.device)
    _lib.check(ctx.lib.gsb_residual(ctx.handle, D.ptr(p), D.ptr(s), D.ptr(out), f.B, D.stream_ptr()), "gsb_residual")
    return f.out(out)


def residual_linf(psi, source, r_grid, dr: float, dz: float):
    """multigrid_solve.py:338-349; float for one field, array for a batch."""
    f = _Fields(psi)
    ctx = _ctx(f, r_grid, dr, dz)
    p, s = f.dev(psi, copy=False), f.dev(source, copy=False)
    out = D.empty((f.B,), f.device)
    _lib.check(ctx.lib.gsb_residual_norms(ctx.handle, D.ptr(p), D.ptr(s), D.ptr(out), c_void_p(), f.B,
                                          D.stream_ptr()), "gsb_residual_norms")
    if f.batched:
        return out if f.torch_in else out.cpu().numpy()
    return float(out.cpu().numpy()[0])


def multigrid_vcycle(psi, source, r_grid, dr: float, dz: float, *, omega: float = 1.0, pre_smooth: int = 3,
                     post_smooth: int = 3, min_grid: int = 5):
    """multigrid_solve.py:252-335: returns the improved estimate, input untouched."""
    omega = validate_sor_omega(omega)
    f = _Fields(psi)
    ctx = _ctx(f, r_grid, dr, dz)
    p = f.dev(psi, copy=True)
    s = f.dev(source, copy=False)
    _lib.check(ctx.lib.gsb_vcycle(ctx.handle, D.ptr(p), D.ptr(s), f.B, omega, int(pre_smooth), int(post_smooth),
                                  int(min_grid), D.stream_ptr()), "gsb_vcycle")
    return f.out(p)


def multigrid_solve(source, psi_bc, r_min: float, r_max: float, z_min: float, z_max: float, nr: int, nz: int, *,
                    tol: float = 1e-6, max_cycles: int = 500, omega: float = 1.0, pre_smooth: int = 3,
                    post_smooth: int = 3, min_grid: int = 5):
    """multigrid_solve.py:352-463: ``(psi, residual, n_cycles, converged)``.

    With a leading batch axis the last three are arrays.
    """
    f = _Fields(psi_bc)
    src_shape = tuple(source.shape)
    if (f.nz, f.nr) != (nz, nr) or src_shape[-2:] != (nz, nr) or len(src_shape) != len(tuple(psi_bc.shape)):
        raise ValueError(
            f"source and psi_bc must have shape (nz, nr) = ({nz}, {nr}); "
            f"got source={src_shape}, psi_bc={tuple(psi_bc.shape)}."
        )
    if not (np.isfinite(tol) and tol > 0.0):
        raise ValueError("tol must be finite and > 0.")
    if max_cycles < 1:
        raise ValueError("max_cycles must be >= 1.")
    omega = validate_sor_omega(omega)
    r_axis = np.linspace(r_min, r_max, nr)
    z_axis = np.linspace(z_min, z_max, nz)
    dr = float(r_axis[1] - r_axis[0]) if nr > 1 else 1.0
    dz = float(z_axis[1] - z_axis[0]) if nz > 1 else 1.0
    torch = D.torch_mod()
    ctx = D.get_context(nz, nr, r_axis, None, dr, dz, f.B, f.device)
    psi = f.dev(psi_bc, copy=True)
    src = f.dev(source, copy=False)
    res = D.empty((f.B,), f.device)
    cyc = D.empty((f.B,), f.device, torch.int32)
    conv = D.empty((f.B,), f.device, torch.int32)
    _lib.check(ctx.lib.gsb_mg_solve(ctx.handle, D.ptr(src), D.ptr(psi), f.B, float(tol), int(max_cycles), omega,
                                    int(pre_smooth), int(post_smooth), int(min_grid), D.ptr(res), D.ptr(cyc),
                                    D.ptr(conv), D.stream_ptr()), "gsb_mg_solve")
    if f.batched:
        if f.torch_in:
            return psi, res, cyc, conv.bool()
        return psi.cpu().numpy(), res.cpu().numpy(), cyc.cpu().numpy(), conv.cpu().numpy().astype(bool)
    r, c, k = res.cpu().numpy(), cyc.cpu().numpy(), conv.cpu().numpy()
    return f.out(psi), float(r[0]), int(c[0]), bool(k[0])
