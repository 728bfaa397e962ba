"""GPU parity of the reference's native solver ABI (libscpn_solver.so symbols) and lane-C wall pieces."""
from __future__ import annotations

import ctypes

import numpy as np
import pytest

import gs_oracle as G
from conftest import golden, rel_l2
from scpn_fusion_core_b200 import _lib

pytestmark = pytest.mark.gpu
dp = ctypes.POINTER(ctypes.c_double)


def test_run_step_matches_compiled_reference_and_oracle():
    """Golden: the UNMODIFIED solver.cpp compiled with its own build line (FMA contraction allowed
    there, hence 1e-12); oracle: same operand order without contraction, bit-exact."""
    lib = _lib.load()
    z = golden("hpc_solver")
    j = np.ascontiguousarray(z["j"])
    nz, nr = j.shape
    h = lib.create_solver(nr, nz, 2.0, 10.0, -4.0, 4.0)
    assert h
    psi = np.zeros((nz, nr))
    lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 7)
    assert rel_l2(psi, z["psi_7"]) < 1e-12
    ref, _ = G.hpc_run_step(np.zeros((nz, nr)), j, 2.0, 10.0, -4.0, 4.0, 7)
    np.testing.assert_array_equal(psi, ref)
    lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 5)  # warm start persists
    assert rel_l2(psi, z["psi_12"]) < 1e-12
    lib.set_boundary_dirichlet(h, 0.25)
    delta = ctypes.c_double(-1.0)
    n = lib.run_step_converged(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 400, 1.5, 1e-9,
                               ctypes.byref(delta))
    assert n == int(z["conv_meta"][0])  # the reference also exhausts its 400 sweeps here
    assert rel_l2(psi, z["psi_conv"]) < 1e-9
    assert abs(delta.value - z["conv_meta"][1]) <= 1e-6 * z["conv_meta"][1]
    assert np.all(psi[0, :] == 0.25) and np.all(psi[:, -1] == 0.25)
    # early stop: a loose tolerance must end before max_iterations with delta <= tol
    n2 = lib.run_step_converged(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 400, 1.5, 1e-3,
                                ctypes.byref(delta))
    assert 1 <= n2 < 400 and delta.value <= 1e-3
    # wrong size: silent no-op / 0
    before = psi.copy()
    lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr - 1, 3)
    np.testing.assert_array_equal(psi, before)
    assert lib.run_step_converged(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), 5, 3, 1.5, 0.0, None) == 0
    lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 0)  # iterations clamped to 1
    lib.delete_solver(h)


def test_wall_response_matrix_and_flux(monkeypatch):
    """Lane C (parity unpinned by a reference run): device matrix vs the NumPy restatement, and the
    FP64 tensor-core contraction vs M @ (J*dA)."""
    import torch
    from scpn_fusion_core_b200 import _device as D
    nz, nr = 33, 29
    R = np.linspace(1.0, 3.0, nr)
    Z = np.linspace(-1.5, 1.5, nz)
    dr, dz = float(R[1] - R[0]), float(Z[1] - Z[0])
    ctx = D.get_context(nz, nr, R, Z, dr, dz, 300, 0)
    M_ref, b_idx, s_idx = G.wall_response_matrix(R, Z)
    m = D.empty(M_ref.shape, 0)
    _lib.check(ctx.lib.gsb_wall_matrix(ctx.handle, G.MU0_SI, D.ptr(m), D.stream_ptr()))
    np.testing.assert_allclose(m.cpu().numpy(), M_ref, rtol=2e-14, atol=0)
    rng = np.random.default_rng(1)
    for B in (1, 7, 80, 128, 300):  # >= 128 takes the large-batch tile kernel
        J = rng.normal(size=(B, nz, nr))
        dA = dr * dz
        Jd = D.to_device(J, 0)
        wall = D.empty((B, M_ref.shape[0]), 0)
        _lib.check(ctx.lib.gsb_wall_flux(ctx.handle, D.ptr(m), D.ptr(Jd), dA, D.ptr(wall), B, D.stream_ptr()))
        ref = np.stack([G.plasma_wall_flux(M_ref, s_idx, J[b], dA) for b in range(B)])
        scale = np.abs(M_ref).sum(axis=1).max() * np.abs(J).max() * dA
        assert np.max(np.abs(wall.cpu().numpy() - ref)) < 1e-13 * scale
        bc = D.zeros((B, nz, nr), 0)
        _lib.check(ctx.lib.gsb_wall_scatter(ctx.handle, D.ptr(wall), D.ptr(bc), 0, B, D.stream_ptr()))
        got = bc.cpu().numpy().reshape(B, -1)
        np.testing.assert_array_equal(got[:, b_idx], wall.cpu().numpy())
        assert np.all(got[:, s_idx] == 0.0)
