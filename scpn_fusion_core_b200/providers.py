"""GPU-tier providers for the reference's kernel registry.

The reference dispatches ``gs_rb_sor_smooth`` and ``multigrid_solve`` through
``scpn_fusion.core._multi_compat`` (``register_kernel(name, tier, fn)``, ``_multi_compat.py:240``)
and ``benchmarks/bench_gpu_gs_solver.py`` calls ``providers._gpu_gs_rb_sor_smooth`` by name.
These callables have the reference providers' signatures
(``_multi_compat_providers.py:322-344,723-792``) but run in FP64 on the B200, so the
cross-tier agreement is no longer f32-bounded.

``PyGpuSolver`` / ``py_gpu_available`` / ``py_gpu_info`` reproduce the surface of the
reference's PyO3 module ``scpn_fusion_rs`` (``fusion-python/src/bindings/gpu.rs:19-90``) so an
UNMODIFIED ``bench_gpu_gs_solver.py`` can be pointed at this package.
``PyFusionKernel`` / ``PyEquilibriumResult`` / ``multigrid_vcycle`` are the equilibrium names of the same
module (``fusion-python/src/bindings/equilibrium.rs:14-116,368-401``), i.e. what the reference's Rust tier
(``_rust_compat.RustAcceleratedKernel``, ``_multi_compat_providers._rust_multigrid_solve``) binds.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from . import _device as D
from . import _lib
import importlib

_mg = importlib.import_module(__package__ + ".multigrid_solve")  # the package re-exports a same-named function


def _gpu_gs_rb_sor_smooth(psi: Any, source: Any, r_left: float, r_right: float, z_bottom: float, z_top: float, *,
                          omega: float = 1.3, n_sweeps: int = 50) -> Any:
    """``gs_rb_sor_smooth`` GPU tier: n_sweeps RB-SOR sweeps on a copy of *psi*, float64 out.

    Geometry exactly as the NumPy tier builds it (``_multi_compat_providers.py:744-751``):
    r_grid from ``linspace`` but ``dr = (r_right - r_left)/(nr - 1)``.
    """
    psi_arr = np.array(psi, dtype=np.float64, copy=True)
    source_arr = np.asarray(source, dtype=np.float64)
    nz, nr = psi_arr.shape
    r_axis = np.linspace(r_left, r_right, nr)
    dr = (r_right - r_left) / (nr - 1)
    dz = (z_top - z_bottom) / (nz - 1)
    return _mg.mg_smooth(psi_arr, source_arr, r_axis, dr, dz, omega, n_sweeps)


def _gpu_multigrid_solve(source: Any, psi_bc: Any, r_min: float, r_max: float, z_min: float, z_max: float,
                         nr: int, nz: int, *, tol: float = 1e-6, max_cycles: int = 500) -> Any:
    """``multigrid_solve`` GPU tier (``_multi_compat_providers.py:322-344``)."""
    return _mg.multigrid_solve(source, psi_bc, r_min, r_max, z_min, z_max, nr, nz, tol=tol, max_cycles=max_cycles)


def _gpu_equilibrium_kernel_loader():
    """Class loader for ``register_kernel_class("equilibrium_kernel", GPU, loader)``."""
    from .fusion_kernel import FusionKernel

    return FusionKernel


def py_gpu_available() -> bool:
    """``scpn_fusion_rs.py_gpu_available`` analogue: True when libgsb200 sees a CUDA device."""
    try:
        return _lib.load().gsb_device_count() > 0
    except Exception:
        return False


def py_gpu_info() -> str:
    torch = D.torch_mod()
    if not py_gpu_available() or not torch.cuda.is_available():
        return "none"
    p = torch.cuda.get_device_properties(torch.cuda.current_device())
    return f"{p.name} (sm_{p.major}{p.minor}, {p.total_memory // 2**20} MiB, libgsb200 FP64)"


class PyGpuSolver:
    """Surface of the reference's PyO3 ``PyGpuSolver`` (``bindings/gpu.rs:19-81``).

    ``solve(psi, source, iterations, omega)`` takes flat sequences and returns a flat float32
    array like the wgpu tier, but the sweeps themselves run in FP64.
    """

    def __init__(self, nr: int, nz: int, r_left: float, r_right: float, z_bottom: float, z_top: float):
        if nr < 3 or nz < 3:
            raise ValueError("grid must be at least 3x3")
        self.nr, self.nz = int(nr), int(nz)
        self.box = (float(r_left), float(r_right), float(z_bottom), float(z_top))

    def solve(self, psi, source, iterations: int, omega: float):
        p = np.asarray(psi, dtype=np.float64).reshape(self.nz, self.nr)
        s = np.asarray(source, dtype=np.float64).reshape(self.nz, self.nr)
        out = _gpu_gs_rb_sor_smooth(p, s, *self.box, omega=omega, n_sweeps=int(iterations))
        return out.astype(np.float32).ravel()


class PyEquilibriumResult:
    """Read-only result record of ``PyFusionKernel.solve_equilibrium`` (``bindings/equilibrium.rs:119-152``)."""

    __slots__ = ("converged", "iterations", "residual", "axis_r", "axis_z", "x_point_r", "x_point_z", "psi_axis",
                 "psi_boundary", "solve_time_ms")

    def __init__(self, **kw: Any) -> None:
        for k in self.__slots__:
            object.__setattr__(self, k, kw[k])

    def __setattr__(self, name: str, value: Any) -> None:
        raise AttributeError("PyEquilibriumResult is read-only")

    def __repr__(self) -> str:
        return f"EquilibriumResult(converged={self.converged}, iters={self.iterations}, residual={self.residual:.2e})"


class PyFusionKernel:
    """Surface of the reference's PyO3 ``PyFusionKernel`` (``bindings/equilibrium.rs:14-116``) on the B200.

    The arithmetic is this package's ``FusionKernel`` (parity with the reference's NumPy lane; the Rust lane
    itself differs from NumPy in documented details).  ``calculate_thermodynamics`` is not part
    of the equilibrium path and raises ``NotImplementedError``.
    """

    _METHODS = {"sor": "sor", "picard_sor": "sor", "multigrid": "multigrid", "picard_multigrid": "multigrid", "mg": "multigrid"}

    def __init__(self, config_path: str) -> None:
        from .fusion_kernel import FusionKernel
        try:
            self._k = FusionKernel(str(config_path))
        except (FileNotFoundError, ValueError, KeyError) as exc:  # PyO3 maps every load failure to PyIOError
            raise OSError(str(exc)) from exc

    def solve_equilibrium(self) -> PyEquilibriumResult:
        import time
        t0 = time.perf_counter()
        try:
            r = self._k.solve_equilibrium()
        except Exception as exc:  # PyO3: PyRuntimeError(e.to_string())
            raise RuntimeError(str(exc)) from exc
        iz, ir, psi_ax = self._k._find_magnetic_axis()
        (rx, zx), psi_b = self._k.find_x_point(self._k.Psi)
        return PyEquilibriumResult(converged=bool(r["converged"]), iterations=int(r["iterations"]), residual=float(r["residual"]),
                                   axis_r=float(self._k.R[ir]), axis_z=float(self._k.Z[iz]), x_point_r=float(rx),
                                   x_point_z=float(zx), psi_axis=float(psi_ax), psi_boundary=float(psi_b),
                                   solve_time_ms=(time.perf_counter() - t0) * 1e3)

    def calculate_thermodynamics(self, p_aux_mw: float):
        raise NotImplementedError("calculate_thermodynamics is outside the B200 equilibrium path (SURVEY.md 8)")

    def get_psi(self) -> np.ndarray:
        return np.array(self._k.Psi, dtype=np.float64, copy=True)

    def get_j_phi(self) -> np.ndarray:
        return np.array(self._k.J_phi, dtype=np.float64, copy=True)

    def get_r(self) -> np.ndarray:
        return np.array(self._k.R, dtype=np.float64, copy=True)

    def get_z(self) -> np.ndarray:
        return np.array(self._k.Z, dtype=np.float64, copy=True)

    def grid_shape(self) -> tuple[int, int]:
        return int(self._k.NR), int(self._k.NZ)

    def set_solver_method(self, method: str) -> None:
        m = self._METHODS.get(str(method).lower())
        if m is None:
            raise ValueError(f"Unknown solver method '{method}'. Use 'sor' or 'multigrid'.")
        self._k.cfg["solver"]["solver_method"] = m

    def solver_method(self) -> str:
        m = self._k.cfg["solver"].get("solver_method", "multigrid")
        return "sor" if m == "sor" else "multigrid"


def multigrid_vcycle(source: Any, psi_bc: Any, r_min: float, r_max: float, z_min: float, z_max: float, nr: int, nz: int,
                     tol: float = 1e-6, max_cycles: int = 500):
    """``scpn_fusion_rs.multigrid_vcycle`` (``bindings/equilibrium.rs:368-401``): despite the name a full
    V-cycle solve, positional ``tol`` / ``max_cycles``; returns ``(psi, residual, cycles, converged)``."""
    return _gpu_multigrid_solve(source, psi_bc, r_min, r_max, z_min, z_max, int(nr), int(nz), tol=tol, max_cycles=int(max_cycles))


def register(multi: Any, replace: bool = True) -> None:
    """Register the GPU tier into the reference's registry module (``_multi_compat``).

    The reference already registers its own wgpu provider for ``gs_rb_sor_smooth`` at ``BackendTier.GPU``
    (``_multi_compat_providers.py:828-847``); entries of one tier keep their registration order, so with
    ``replace`` (default) that entry is dropped first - both cannot serve the same tier, and the wgpu one needs the
    ``scpn_fusion_rs`` wheel this package stands in for.
    """
    tier = multi.BackendTier.GPU
    if replace:
        with multi._registry_lock:
            for name in ("gs_rb_sor_smooth", "multigrid_solve"):
                if name in multi._registry:
                    multi._registry[name] = [(t, f) for t, f in multi._registry[name] if t != tier]
                multi._dispatch_cache.pop(name, None)
        reg = getattr(multi, "_class_registry", None)
        if reg is not None and "equilibrium_kernel" in reg:
            with multi._class_registry_lock:
                reg["equilibrium_kernel"] = [(t, f) for t, f in reg["equilibrium_kernel"] if t != tier]
                multi._class_dispatch_cache.pop("equilibrium_kernel", None)
    multi.register_kernel("gs_rb_sor_smooth", tier, _gpu_gs_rb_sor_smooth)
    multi.register_kernel("multigrid_solve", tier, _gpu_multigrid_solve)
    multi.register_kernel_class("equilibrium_kernel", tier, _gpu_equilibrium_kernel_loader)
