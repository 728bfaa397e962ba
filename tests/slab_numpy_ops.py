"""Oracle-backed compute backend for scpn_fusion_core_b200.slab (TEST INFRASTRUCTURE).

Implements the ``ops`` interface of ``SlabMultigrid`` with the NumPy oracle so that the slab
orchestration (row partition, halo exchange, row offsets, gather at the coarse levels) can be checked
on CPU ranks over gloo against a single-process oracle solve."""
from __future__ import annotations

import numpy as np
import torch

import gs_oracle as G


def rb_sor_smooth_offset(psi, source, r_grid, dr, dz, omega, n_sweeps, par_off):
    """gs_oracle.rb_sor_smooth with the colour order given by the GLOBAL row parity: local colour
    (iz+ir)%2 is global colour ((iz+par_off+ir)%2), and global colour 0 is relaxed first."""
    nz, nr = psi.shape
    a_e, a_w, a_ns, a_c = G._stencil_coeffs(r_grid, dr, dz)
    order = (0, 1) if par_off % 2 == 0 else (1, 0)
    for _ in range(int(n_sweeps)):
        for parity in order:
            for zs, rs in G._colour_slices(nz, nr, parity):
                zi = slice(zs.start - 1, nz - 2, 2)
                ri = slice(rs.start - 1, nr - 2, 2)
                east = psi[zs, rs.start + 1: nr: 2]
                west = psi[zs, rs.start - 1: nr - 2: 2]
                south = psi[zs.start - 1: nz - 2: 2, rs]
                north = psi[zs.start + 1: nz: 2, rs]
                gs = (a_e[zi, ri] * east + a_w[zi, ri] * west + a_ns * south + a_ns * north - source[zs, rs]) / a_c
                psi[zs, rs] = (1.0 - omega) * psi[zs, rs] + omega * gs
    return psi


class NumpySlabOps:
    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float64)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).clone()

    def to_numpy(self, t):
        return t.numpy()

    @staticmethod
    def _rg(L, rows):
        return np.tile(L.r_row, (rows, 1))

    def smooth(self, L, x, f, omega, sweeps, out=None):
        a = x.numpy()
        rb_sor_smooth_offset(a, f.numpy(), self._rg(L, a.shape[0]), L.dr, L.dz, omega, sweeps, L.row0)
        return x

    def residual_restrict(self, L, C, x, f, d, roff, ci0, ci1):
        if ci0 >= ci1:
            return
        xa, fa = x.numpy(), f.numpy()
        defect = -G.gs_residual(xa, fa, self._rg(L, xa.shape[0]), L.dr, L.dz)
        lo = 2 * ci0 + roff - 2
        hi = 2 * (ci1 - 1) + roff + 2
        sub = defect[lo:hi + 1]
        c = G.restrict_full_weight(sub)
        d.numpy()[ci0:ci1] = c[1:-1]
        d.numpy()[ci0:ci1, 0] = 0.0
        d.numpy()[ci0:ci1, -1] = 0.0

    def prolong_add(self, L, x, e, roff, fi0, fi1):
        if fi0 >= fi1:
            return
        ea = e.numpy()
        i0 = (fi0 - roff) >> 1
        i1 = ((fi1 - 1 - roff) + 1) >> 1
        p = G.prolong_bilinear(ea[i0:i1 + 1], 2 * (i1 - i0) + 1, L.nr)
        xa = x.numpy()
        for lf in range(fi0, fi1):
            xa[lf, 1:-1] = xa[lf, 1:-1] + p[(lf - roff) - 2 * i0, 1:-1]

    def residual_linf(self, L, x, f, row0, row1):
        if row0 >= row1:
            return 0.0
        xa, fa = x.numpy(), f.numpy()
        r = G.gs_residual(xa, fa, self._rg(L, xa.shape[0]), L.dr, L.dz)[row0:row1, 1:-1]
        return float(np.max(np.abs(r)))

    def coarse_vcycle(self, Gd, d_full, omega, pre, post, min_grid):
        rg = np.tile(Gd["r_row"], (Gd["nz"], 1))
        d = d_full.numpy()
        out = G.vcycle(np.zeros_like(d), d, rg, Gd["dr"], Gd["dz"], omega=omega, pre=pre, post=post, min_grid=min_grid)
        return torch.from_numpy(out)
