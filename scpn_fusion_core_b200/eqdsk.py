"""G-EQDSK (EFIT) equilibrium files: the data format on either side of the equilibrium path.

Drop-in for the reference's ``src/scpn_fusion/core/eqdsk.py`` (``GEqdsk``, ``read_geqdsk``,
``write_geqdsk``): same container fields and derived grids, same accepted inputs (free-format or
fixed-width ``5e16.9`` numbers, including run-together values such as ``2.385E+00-1.216E+01`` and Fortran
``D`` exponents), same safety limits and ``ValueError`` conditions, byte-identical output files, and the
same ``to_config()`` dictionary (``eqdsk.py:128-186``) that seeds a ``FusionKernel`` with the EFIT
boundary as an isoflux shape target.  ``from_kernel`` goes the other way: a solved ``FusionKernel`` ->
``GEqdsk`` ready to write.  Pure host-side IO; nothing here touches the GPU.

Format limit shared with the reference: a value with a three-digit decimal exponent fills all 24 columns of its
cell, so neighbouring cells run together and cannot be split again; keep |x| within [1e-99, 1e99] (or 0).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np

MAX_GEQDSK_BYTES = 10 * 1024 * 1024
MAX_GEQDSK_GRID_POINTS = 1_000_000
MAX_GEQDSK_CONTOUR_POINTS = 100_000
MAX_GEQDSK_NUMERIC_TOKENS = 4_500_000

GEQDSK_SOURCE_CONVENTION_MODES = {"raw_canonical": "raw_canonical",
                                  "public_sparc_named_adapter": "public_sparc_named_adapter"}
# public SPARC files whose p'/FF' are stored per (psi/2pi) (eqdsk.py:44-53)
GEQDSK_PUBLIC_SPARC_SOURCE_ADAPTERS = {f"sparc_{n}.eqdsk": "scaled_by_2pi" for n in (1305, 1310, 1315, 1349)}
GEQDSK_SOURCE_CONVENTION_ADAPTERS = {"scaled_by_2pi": 2.0 * np.pi}

_SCALARS = ("rdim", "zdim", "rcentr", "rleft", "zmid", "rmaxis", "zmaxis", "simag", "sibry", "bcentr", "current")
_PROFILES = ("fpol", "pres", "ffprime", "pprime", "qpsi")
_NUMBER = re.compile(r"[+-]?\d*\.?\d+(?:[eEdD][+-]?\d+)?")


def _empty() -> np.ndarray:
    return np.array([], dtype=np.float64)


@dataclass
class GEqdsk:
    """Everything a G-EQDSK file holds (Lao et al., Nucl. Fusion 25 (1985) 1611)."""

    description: str = ""
    nw: int = 0   # R points
    nh: int = 0   # Z points
    rdim: float = 0.0
    zdim: float = 0.0
    rcentr: float = 0.0
    rleft: float = 0.0
    zmid: float = 0.0
    rmaxis: float = 0.0
    zmaxis: float = 0.0
    simag: float = 0.0    # psi on axis (Wb/rad)
    sibry: float = 0.0    # psi on the boundary
    bcentr: float = 0.0
    current: float = 0.0  # A
    fpol: np.ndarray = field(default_factory=_empty)
    pres: np.ndarray = field(default_factory=_empty)
    ffprime: np.ndarray = field(default_factory=_empty)
    pprime: np.ndarray = field(default_factory=_empty)
    qpsi: np.ndarray = field(default_factory=_empty)
    psirz: np.ndarray = field(default_factory=_empty)   # (nh, nw)
    rbdry: np.ndarray = field(default_factory=_empty)
    zbdry: np.ndarray = field(default_factory=_empty)
    rlim: np.ndarray = field(default_factory=_empty)
    zlim: np.ndarray = field(default_factory=_empty)
    source_convention: str = "raw_canonical"
    source_convention_adapter: str = "not_applied"
    source_convention_adapter_pass: bool = False
    source_convention_metadata: dict = field(default_factory=dict)

    @property
    def r(self) -> np.ndarray:
        return np.linspace(self.rleft, self.rleft + self.rdim, self.nw)

    @property
    def z(self) -> np.ndarray:
        return np.linspace(self.zmid - self.zdim / 2, self.zmid + self.zdim / 2, self.nh)

    @property
    def psi_norm(self) -> np.ndarray:
        return np.linspace(0.0, 1.0, self.nw)

    def psi_to_norm(self, psi) -> np.ndarray:
        return (psi - self.simag) / (self.sibry - self.simag)

    def to_config(self, name: str = "eqdsk") -> dict[str, Any]:
        """FusionKernel config of the same box and grid; the EFIT boundary becomes the isoflux target
        (``free_boundary.target_flux_points/values``), the limiter is carried along; no coils in a GEQDSK."""
        fb: dict[str, Any] = {"magnetic_axis": [float(self.rmaxis), float(self.zmaxis)], "psi_axis": float(self.simag),
                              "psi_boundary": float(self.sibry), "boundary_source": "geqdsk_rbdry_zbdry"}
        if self.rbdry.size and self.zbdry.size:
            if self.rbdry.shape != self.zbdry.shape:
                raise ValueError("GEQDSK boundary R/Z arrays must have matching lengths.")
            if not (np.all(np.isfinite(self.rbdry)) and np.all(np.isfinite(self.zbdry))):
                raise ValueError("GEQDSK boundary contour must contain finite values only.")
            pts = np.column_stack([self.rbdry, self.zbdry]).astype(np.float64)
            fb["target_flux_points"] = pts.tolist()
            fb["target_flux_values"] = np.full(pts.shape[0], float(self.sibry), dtype=np.float64).tolist()
        if self.rlim.size or self.zlim.size:
            if self.rlim.shape != self.zlim.shape:
                raise ValueError("GEQDSK limiter R/Z arrays must have matching lengths.")
            if not (np.all(np.isfinite(self.rlim)) and np.all(np.isfinite(self.zlim))):
                raise ValueError("GEQDSK limiter contour must contain finite values only.")
            fb["limiter_points"] = np.column_stack([self.rlim, self.zlim]).astype(np.float64).tolist()
        r, z = self.r, self.z
        return {"reactor_name": name, "grid_resolution": [self.nw, self.nh],
                "dimensions": {"R_min": float(r[0]), "R_max": float(r[-1]), "Z_min": float(z[0]), "Z_max": float(z[-1])},
                "physics": {"plasma_current_target": float(self.current / 1e6), "vacuum_permeability": 1.0},
                "coils": [], "free_boundary": fb,
                "solver": {"max_iterations": 1000, "convergence_threshold": 1e-4, "relaxation_factor": 0.1}}


# -- validation (eqdsk.py:213-287) -----------------------------------------------------------------

def _check_grid(nw: int, nh: int) -> None:
    if nw < 2 or nh < 2:
        raise ValueError(f"GEQDSK grid dimensions must be >= 2x2, got {(nw, nh)}")
    if nw * nh > MAX_GEQDSK_GRID_POINTS:
        raise ValueError(f"GEQDSK grid dimensions exceed safety limit {MAX_GEQDSK_GRID_POINTS}: got {nw}x{nh}")


def _check_count(name: str, n: int) -> None:
    if n < 0:
        raise ValueError(f"GEQDSK {name} count must be non-negative")
    if n > MAX_GEQDSK_CONTOUR_POINTS:
        raise ValueError(f"GEQDSK {name} count exceeds safety limit {MAX_GEQDSK_CONTOUR_POINTS}")


def _finite(name: str, a: np.ndarray) -> None:
    if not np.all(np.isfinite(a)):
        raise ValueError(f"GEQDSK {name} must contain finite values only.")


def validate_geqdsk(eq: GEqdsk) -> None:
    _check_grid(eq.nw, eq.nh)
    if eq.rdim <= 0.0 or eq.zdim <= 0.0:
        raise ValueError("GEQDSK rdim and zdim must be positive.")
    if eq.rcentr <= 0.0:
        raise ValueError("GEQDSK rcentr must be positive.")
    if eq.sibry == eq.simag:
        raise ValueError("GEQDSK psi boundary must differ from psi axis.")
    for name in _SCALARS:
        if not math.isfinite(float(getattr(eq, name))):
            raise ValueError(f"GEQDSK scalar {name} must be finite.")
    for name in _PROFILES:
        a = getattr(eq, name)
        if a.shape != (eq.nw,):
            raise ValueError(f"GEQDSK {name} shape must be {(eq.nw,)}, got {a.shape}.")
        _finite(name, a)
    if eq.psirz.shape != (eq.nh, eq.nw):
        raise ValueError(f"GEQDSK psirz shape must be {(eq.nh, eq.nw)}, got {eq.psirz.shape}.")
    _finite("psirz", eq.psirz)
    for rn, zn in (("rbdry", "zbdry"), ("rlim", "zlim")):
        ra, za = getattr(eq, rn), getattr(eq, zn)
        if ra.shape != za.shape:
            raise ValueError(f"GEQDSK {rn}/{zn} contours must have matching lengths.")
        _finite(rn, ra)
        _finite(zn, za)


# -- reader (eqdsk.py:349-536) ------------------------------------------------------------------------

class _Tokens:
    """Cursor over the numeric tokens of the file body."""

    def __init__(self, tokens: list[str]):
        self.t, self.i = tokens, 0

    def _value(self, k: int) -> float:
        try:
            v = float(self.t[k].replace("D", "E").replace("d", "e"))
        except ValueError as exc:
            raise ValueError(f"GEQDSK token[{k}] is not a valid finite float.") from exc
        if not math.isfinite(v):
            raise ValueError(f"GEQDSK token[{k}] must be finite.")
        return v

    def one(self) -> float:
        if self.i >= len(self.t):
            raise ValueError("GEQDSK file ended before all required values were present")
        self.i += 1
        return self._value(self.i - 1)

    def many(self, n: int) -> np.ndarray:
        if n < 0:
            raise ValueError("GEQDSK array length must be non-negative")
        if self.i + n > len(self.t):
            raise ValueError("GEQDSK file ended before all required array values were present")
        out = np.array([self._value(self.i + k) for k in range(n)], dtype=np.float64)
        self.i += n
        return out

    def left(self) -> int:
        return len(self.t) - self.i


def read_geqdsk(path, *, source_convention_mode: str = "raw_canonical") -> GEqdsk:
    """Parse a G-EQDSK file.  ``source_convention_mode="public_sparc_named_adapter"`` rescales p'/FF' of
    the four recognised public SPARC files by 2 pi and records the provenance; everything else is raw."""
    path = Path(path)
    if source_convention_mode not in GEQDSK_SOURCE_CONVENTION_MODES:
        raise ValueError(f"Unsupported source_convention_mode: {source_convention_mode}. Expected one of: "
                         + ", ".join(sorted(GEQDSK_SOURCE_CONVENTION_MODES)))
    size = path.stat().st_size
    if size > MAX_GEQDSK_BYTES:
        raise ValueError(f"GEQDSK file too large: {size} bytes exceeds {MAX_GEQDSK_BYTES}")
    with open(path, "r", encoding="utf-8") as fh:
        lines = fh.readlines()
    if not lines:
        raise ValueError("GEQDSK file is empty")
    head = lines[0].split()
    if len(head) < 3:
        raise ValueError("GEQDSK header must contain idum, nw, and nh")
    nw, nh = int(head[-2]), int(head[-1])
    desc = " ".join(head[:-3]) if len(head) > 3 else ""
    _check_grid(nw, nh)
    tokens: list[str] = []
    for line in lines[1:]:
        tokens.extend(_NUMBER.findall(line))
        if len(tokens) > MAX_GEQDSK_NUMERIC_TOKENS:
            raise ValueError(f"GEQDSK numeric token count exceeds safety limit {MAX_GEQDSK_NUMERIC_TOKENS}")
    tk = _Tokens(tokens)
    block = [tk.one() for _ in range(20)]  # 11 scalars, then duplicates and padding
    eq = GEqdsk(description=desc, nw=nw, nh=nh, **dict(zip(_SCALARS, block[:11])))
    eq.fpol, eq.pres, eq.ffprime, eq.pprime = (tk.many(nw) for _ in range(4))
    eq.psirz = tk.many(nh * nw).reshape(nh, nw)
    eq.qpsi = tk.many(nw)
    nbdry, nlim = int(tk.one()), int(tk.one())
    _check_count("boundary", nbdry)
    _check_count("limiter", nlim)
    if 2 * (nbdry + nlim) > tk.left():
        raise ValueError("GEQDSK file ended before all required contour values were present")
    bd = tk.many(2 * nbdry).reshape(nbdry, 2)
    lm = tk.many(2 * nlim).reshape(nlim, 2)
    eq.rbdry, eq.zbdry = bd[:, 0].copy(), bd[:, 1].copy()
    eq.rlim, eq.zlim = lm[:, 0].copy(), lm[:, 1].copy()
    validate_geqdsk(eq)
    if source_convention_mode == "public_sparc_named_adapter":
        _adapt_source(eq, path, source_convention_mode)
    return eq


def _adapt_source(eq: GEqdsk, path: Path, mode: str) -> None:
    adapter = GEQDSK_PUBLIC_SPARC_SOURCE_ADAPTERS.get(path.name.lower())
    if adapter is None:
        eq.source_convention, eq.source_convention_adapter, eq.source_convention_adapter_pass = \
            "raw_canonical", "no_named_adapter", False
        eq.source_convention_metadata = {"requested_mode": mode, "source_file": path.name, "adapter": "no_named_adapter",
                                         "applied_scale": 1.0, "provenance": "public_sparc_named_adapter_no_match",
                                         "error": "no recognized public SPARC convention mapping"}
        return
    scale = GEQDSK_SOURCE_CONVENTION_ADAPTERS[adapter]
    eq.ffprime = np.round(scale * eq.ffprime, decimals=11)
    eq.pprime = np.round(scale * eq.pprime, decimals=11)
    eq.source_convention, eq.source_convention_adapter, eq.source_convention_adapter_pass = "canonical", adapter, True
    eq.source_convention_metadata = {"requested_mode": mode, "source_file": path.name, "adapter": adapter,
                                     "applied_scale": float(scale), "provenance": "public_sparc_named_adapter",
                                     "public_case": path.name}


# -- writer (eqdsk.py:542-632): 24.17e fields, five per line, byte-identical to the reference's output ----

def _rows(values) -> str:
    cells = [f"{float(v):24.17e}" for v in values]
    return "".join("".join(cells[i:i + 5]) + "\n" for i in range(0, len(cells), 5))


def _pairs(rs, zs) -> str:
    """(R, Z) pairs: a line break after every pair count that fills a multiple of five cells, after the
    last pair, and once more when the total cell count is not a multiple of five (the reference's layout)."""
    n = len(rs)
    out = []
    for i in range(n):
        out.append(f"{float(rs[i]):24.17e}{float(zs[i]):24.17e}")
        if ((i + 1) * 2) % 5 == 0 or i == n - 1:
            out.append("\n")
    if n > 0 and (n * 2) % 5 != 0:
        out.append("\n")
    return "".join(out)


def format_geqdsk(eq: GEqdsk) -> str:
    head = f"{eq.description[:48].ljust(48)}   0 {eq.nw:4d} {eq.nh:4d}\n"
    block = [getattr(eq, n) for n in _SCALARS] + [eq.simag, 0.0, eq.rmaxis, 0.0, eq.zmaxis, 0.0, eq.sibry, 0.0, 0.0]
    body = [_rows(block)] + [_rows(np.ravel(getattr(eq, n))) for n in ("fpol", "pres", "ffprime", "pprime", "psirz", "qpsi")]
    counts = f"{len(eq.rbdry):5d}{len(eq.rlim):5d}\n"
    return head + "".join(body) + counts + _pairs(eq.rbdry, eq.zbdry) + _pairs(eq.rlim, eq.zlim)


def write_geqdsk(eq: GEqdsk, path) -> None:
    with open(Path(path), "w") as fh:
        fh.write(format_geqdsk(eq))


# -- solved kernel -> GEQDSK ----------------------------------------------------------------------------------

def from_kernel(kernel: Any, *, description: str = "scpn_fusion_core_b200", bcentr: float = 0.0, rcentr: float | None = None,
                boundary=None, limiter=None) -> GEqdsk:
    """Package a solved ``FusionKernel`` (``Psi`` on its (Z, R) grid, axis and X-point from the device
    topology search) as a ``GEqdsk``.  1-D profiles the kernel does not carry (F, p, q) are zero-filled,
    q is one; ``current`` is the configured plasma current in amperes."""
    nw, nh = int(kernel.NR), int(kernel.NZ)
    iz, ir, psi_ax = kernel._find_magnetic_axis()
    _, psi_b = kernel.find_x_point(kernel.Psi)
    if psi_b == psi_ax:
        psi_b = psi_ax * 0.1 if psi_ax != 0.0 else 1.0
    bd = np.zeros((0, 2)) if boundary is None else np.asarray(boundary, dtype=np.float64).reshape(-1, 2)
    lm = np.zeros((0, 2)) if limiter is None else np.asarray(limiter, dtype=np.float64).reshape(-1, 2)
    zero = np.zeros(nw)
    eq = GEqdsk(description=description, nw=nw, nh=nh, rdim=float(kernel.R[-1] - kernel.R[0]),
                zdim=float(kernel.Z[-1] - kernel.Z[0]),
                rcentr=float(0.5 * (kernel.R[0] + kernel.R[-1])) if rcentr is None else float(rcentr),
                rleft=float(kernel.R[0]), zmid=float(0.5 * (kernel.Z[0] + kernel.Z[-1])), rmaxis=float(kernel.R[ir]),
                zmaxis=float(kernel.Z[iz]), simag=float(psi_ax), sibry=float(psi_b), bcentr=float(bcentr),
                current=float(kernel.cfg["physics"]["plasma_current_target"]) * 1e6,
                fpol=zero.copy(), pres=zero.copy(), ffprime=zero.copy(), pprime=zero.copy(), qpsi=np.ones(nw),
                psirz=np.asarray(kernel.Psi, dtype=np.float64).copy(), rbdry=bd[:, 0].copy(), zbdry=bd[:, 1].copy(),
                rlim=lm[:, 0].copy(), zlim=lm[:, 1].copy())
    validate_geqdsk(eq)
    return eq
