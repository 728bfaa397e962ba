"""Solov'ev analytic equilibrium psi = c1 R^4/8 + c2 Z^2 (SURVEY.md 8c known answers; reference
validation/validate_grad_shafranov_solovev.py and its sealed report): the oracle's operator, SOR sweep and
multigrid solve reproduce the reference's truncation errors, sweep counts and reconstruction errors."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import gs_oracle as G
from conftest import GOLDEN


def _fixture():
    return json.load(open(os.path.join(GOLDEN, "solovev.json")))


def _mesh(geo, n):
    R, Z = np.linspace(geo["r_min"], geo["r_max"], n), np.linspace(geo["z_min"], geo["z_max"], n)
    rr, zz = np.meshgrid(R, Z)
    exact = geo["c1"] * rr ** 4 / 8.0 + geo["c2"] * zz ** 2
    src = geo["c1"] * rr ** 2 + 2.0 * geo["c2"]
    return R, Z, rr, exact, src


def _nrmse(num, exact):
    return float(np.sqrt(np.mean((num[1:-1, 1:-1] - exact[1:-1, 1:-1]) ** 2))) / max(float(exact.max() - exact.min()), 1e-15)


def test_operator_truncation_error_is_second_order():
    fx = _fixture()
    errs, hs = [], []
    for n in (33, 49, 65, 97):
        R, Z, rr, exact, src = _mesh(fx["geometry"], n)
        dr, dz = float(R[1] - R[0]), float(Z[1] - Z[0])
        e = float(np.max(np.abs(G.gs_operator(exact, rr, dr, dz)[1:-1, 1:-1] - src[1:-1, 1:-1])))
        assert e == pytest.approx(fx["operator_errors"][str(n)], rel=1e-9)
        sealed = [r["error"] for r in fx["sealed"]["operator_records"] if r["resolution"] == n][0]
        assert e == pytest.approx(sealed, rel=1e-8)   # sealed report: other numpy build, last digits differ
        errs.append(e)
        hs.append(dr)
    order = float(np.polyfit(np.log(hs), np.log(errs), 1)[0])
    assert order == pytest.approx(fx["sealed"]["operator_order"], abs=1e-6) and errs[-1] < fx["sealed"]["operator_error_gate"]


@pytest.mark.parametrize("n", [33, 49])
def test_sor_reconstruction_sweep_count_and_error(n):
    fx = _fixture()
    R, Z, rr, exact, src = _mesh(fx["geometry"], n)
    dr, dz = float(R[1] - R[0]), float(Z[1] - Z[0])
    psi = np.zeros_like(exact)
    G.copy_wall(psi, exact)
    its, res, conv = 0, np.inf, False
    for sweep in range(40000):
        psi = G.sor_step(psi, src, rr, dr, dz, omega=1.6)
        G.copy_wall(psi, exact)
        its = sweep + 1
        if sweep % 50 == 0:
            res = float(np.max(np.abs((G.gs_operator(psi, rr, dr, dz) - src)[1:-1, 1:-1])))
            if res < 1e-9:
                conv = True
                break
    want = fx["sor"][str(n)]
    assert (its, conv) == (want["iterations"], want["converged"])
    assert _nrmse(psi, exact) == pytest.approx(want["nrmse"], rel=1e-9)
    assert res == pytest.approx(want["residual_inf"], rel=1e-3)


def test_multigrid_reconstruction_97():
    fx = _fixture()
    geo = fx["geometry"]
    R, Z, rr, exact, src = _mesh(geo, 97)
    bc = np.zeros_like(exact)
    G.copy_wall(bc, exact)
    psi, res, cycles, conv = G.mg_solve(src, bc, geo["r_min"], geo["r_max"], geo["z_min"], geo["z_max"], 97, 97, tol=1e-9,
                                        max_cycles=200)
    want, sealed = fx["multigrid_numpy_97"], fx["sealed"]["multigrid_numpy_record"]
    assert (cycles, conv) == (want["cycles"], want["converged"]) == (sealed["cycles"], sealed["converged"])
    assert res == pytest.approx(want["residual"], rel=1e-6)
    assert _nrmse(psi, exact) == pytest.approx(want["nrmse"], rel=1e-9)
    assert _nrmse(psi, exact) == pytest.approx(sealed["nrmse"], rel=1e-8) and _nrmse(psi, exact) < fx["sealed"]["reconstruction_nrmse_gate"]
