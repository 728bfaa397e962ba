"""Solov'ev analytic equilibrium on the B200 path (reference validation/validate_grad_shafranov_solovev.py):
discrete operator truncation error, SOR reconstruction through FusionKernel._sor_step and multigrid
reconstruction through multigrid_solve, against the numbers the reference's own script produces
(tests/golden/solovev.json) and bit-for-bit against the oracle."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import gs_oracle as G
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _fixture():
    return json.load(open(os.path.join(GOLDEN, "solovev.json")))


def _mesh(geo, n):
    R, Z = np.linspace(geo["r_min"], geo["r_max"], n), np.linspace(geo["z_min"], geo["z_max"], n)
    rr, zz = np.meshgrid(R, Z)
    return R, Z, rr, geo["c1"] * rr ** 4 / 8.0 + geo["c2"] * zz ** 2, geo["c1"] * rr ** 2 + 2.0 * geo["c2"]


def _nrmse(num, exact):
    return float(np.sqrt(np.mean((num[1:-1, 1:-1] - exact[1:-1, 1:-1]) ** 2))) / max(float(exact.max() - exact.min()), 1e-15)


def test_operator_truncation_error():
    import scpn_fusion_core_b200 as pkg
    fx = _fixture()
    for n in (33, 49, 65, 97):
        R, Z, rr, exact, src = _mesh(fx["geometry"], n)
        dr, dz = float(R[1] - R[0]), float(Z[1] - Z[0])
        r = np.asarray(pkg.mg_residual(exact, src, rr, dr, dz))
        np.testing.assert_array_equal(r, G.gs_residual(exact, src, rr, dr, dz))
        assert float(np.max(np.abs(r[1:-1, 1:-1]))) == pytest.approx(fx["operator_errors"][str(n)], rel=1e-6)


def test_sor_and_multigrid_reconstruction():
    import scpn_fusion_core_b200 as pkg
    fx = _fixture()
    geo = fx["geometry"]
    # SOR through the kernel method the reference's script drives (33^2: 651 sweeps, residual checked every 50)
    n = 33
    R, Z, rr, exact, src = _mesh(geo, n)
    cfg = {"reactor_name": "solovev", "grid_resolution": [n, n],
           "dimensions": {"R_min": geo["r_min"], "R_max": geo["r_max"], "Z_min": geo["z_min"], "Z_max": geo["z_max"]},
           "physics": {"plasma_current_target": 1.0, "vacuum_permeability": 1.0},
           "coils": [{"name": "PF1", "r": geo["r_min"], "z": geo["z_max"], "current": 1.0}],
           "solver": {"max_iterations": 1, "convergence_threshold": 1e-10, "relaxation_factor": 0.1, "solver_method": "sor",
                      "sor_omega": 1.6}}
    k = pkg.FusionKernel(cfg)
    np.testing.assert_array_equal(k.RR, rr)
    psi = np.zeros_like(exact)
    G.copy_wall(psi, exact)
    its, res, conv = 0, np.inf, False
    for sweep in range(2000):
        psi = np.asarray(k._sor_step(psi, src, omega=1.6), dtype=np.float64)
        G.copy_wall(psi, exact)
        its = sweep + 1
        if sweep % 50 == 0:
            res = float(np.max(np.abs(np.asarray(pkg.mg_residual(psi, src, rr, k.dR, k.dZ))[1:-1, 1:-1])))
            if res < 1e-9:
                conv = True
                break
    want = fx["sor"]["33"]
    assert (its, conv) == (want["iterations"], want["converged"])
    assert _nrmse(psi, exact) == pytest.approx(want["nrmse"], rel=1e-9)
    # multigrid at 97^2 (levels 97 -> 49 -> 25 -> 13 -> 7 -> 4): the reference's NumPy-tier record
    R, Z, rr, exact, src = _mesh(geo, 97)
    bc = np.zeros_like(exact)
    G.copy_wall(bc, exact)
    box = (geo["r_min"], geo["r_max"], geo["z_min"], geo["z_max"])
    psi, res, cycles, conv = pkg.multigrid_solve(src, bc, *box, 97, 97, tol=1e-9, max_cycles=200)
    o_psi, o_res, o_cycles, o_conv = G.mg_solve(src, bc, *box, 97, 97, tol=1e-9, max_cycles=200)
    np.testing.assert_array_equal(np.asarray(psi), o_psi)
    assert (res, cycles, conv) == (o_res, o_cycles, o_conv)
    want = fx["multigrid_numpy_97"]
    assert (cycles, conv) == (want["cycles"], want["converged"])
    assert _nrmse(np.asarray(psi), exact) == pytest.approx(want["nrmse"], rel=1e-9)
