"""GPU: the temporally blocked Jacobi kernel (gsb_jacobi_steps: five _jacobi_step's per pass over HBM,
fusion_kernel_iterative_solver.py:54-95) must be bit-identical to the same number of single steps (gsb_jacobi, which is
itself bit-identical to the reference step, tests/test_gpu_picard.py) - including non-finite input, walls, even sizes,
multi-strip / multi-band tilings and step counts that are not a multiple of five - and to the oracle's jacobi_step."""
from __future__ import annotations

import numpy as np
import pytest

import gs_oracle as G

pytestmark = pytest.mark.gpu


def _run(nz, nr, B, steps, seed, poison):
    import torch
    from scpn_fusion_core_b200 import _device as D, _lib
    rng = np.random.default_rng(seed)
    R = np.linspace(1.5, 4.5, nr)
    Z = np.linspace(-2.0, 2.0, nz)
    psi = rng.normal(size=(B, nz, nr))
    src = rng.normal(size=(B, nz, nr)) * 3.0
    if poison:
        for arr in (psi, src):
            flat = arr.reshape(-1)
            idx = rng.choice(flat.size, size=max(6, flat.size // 400), replace=False)
            flat[idx[0::3]] = np.nan
            flat[idx[1::3]] = np.inf
            flat[idx[2::3]] = -1e300
        psi[0, 0, 1] = np.nan       # wall points too
        psi[0, nz - 1, nr - 2] = -np.inf
        psi[0, nz // 2, 0] = np.inf
    ctx = D.get_context(nz, nr, R, Z, float(R[1] - R[0]), float(Z[1] - Z[0]), B, 0)
    st = D.stream_ptr()
    d_src = D.to_device(src, 0)
    a = D.to_device(psi, 0)
    tmp = torch.empty_like(a)
    _lib.check(ctx.lib.gsb_jacobi_steps(ctx.handle, D.ptr(a), D.ptr(d_src), D.ptr(tmp), steps, B, st), "gsb_jacobi_steps")
    x = D.to_device(psi, 0)
    y = torch.empty_like(x)
    for _ in range(steps):
        _lib.check(ctx.lib.gsb_jacobi(ctx.handle, D.ptr(x), D.ptr(d_src), D.ptr(y), B, st), "gsb_jacobi")
        x, y = y, x
    torch.cuda.synchronize()
    return a.cpu().numpy(), x.cpu().numpy(), psi, src, R, Z


@pytest.mark.parametrize("nz,nr,B,steps,poison", [
    (33, 33, 2, 5, False), (33, 33, 2, 50, True), (65, 129, 3, 10, True), (129, 129, 5, 50, False),
    (130, 70, 2, 12, True), (257, 257, 2, 50, True), (16, 300, 1, 7, True), (300, 16, 2, 15, False), (3, 3, 1, 5, True),
    (5, 64, 1, 5, True), (513, 513, 1, 20, False),
])
def test_blocked_jacobi_equals_single_steps(nz, nr, B, steps, poison):
    got, want, *_ = _run(nz, nr, B, steps, 7 + nz + nr, poison)
    np.testing.assert_array_equal(got, want)  # NaN positions must agree as well (assert_array_equal treats NaN == NaN)


def test_blocked_jacobi_equals_the_oracle():
    got, _, psi, src, R, Z = _run(65, 49, 2, 10, 3, True)
    rr = np.meshgrid(R, Z)[0]
    for b in range(2):
        x = psi[b]
        for _ in range(10):
            x = G.jacobi_step(x, src[b], rr, float(R[1] - R[0]), float(Z[1] - Z[0]))
        np.testing.assert_array_equal(got[b], x)
