"""Host-side logic of scpn_fusion_core_b200.free_boundary on CPU: the device Green's-function launch
(`_mutual_device`) is replaced by the oracle's `mutual_matrix`, everything else - argument checks, bounded
least squares, probe differencing, contour diagnostics, flux sampling - is the product code, compared with
values produced by the UNMODIFIED reference (tests/golden/free_boundary_shape.npz)."""
from __future__ import annotations

import json
from types import SimpleNamespace

import numpy as np
import pytest

import gs_oracle as G
from conftest import golden


@pytest.fixture()
def fb(monkeypatch):
    from scpn_fusion_core_b200 import free_boundary as mod
    monkeypatch.setattr(mod, "_mutual_device", lambda device, positions, turns, obs: G.mutual_matrix(
        [tuple(p) for p in positions], [int(t) for t in turns], obs))
    return mod


def _kernel_and_coils(z):
    from scpn_fusion_core_b200 import CoilSet
    cfg = json.loads(str(z["cfg"]))
    d, (nr, nz) = cfg["dimensions"], cfg["grid_resolution"]
    R, Z = np.linspace(d["R_min"], d["R_max"], nr), np.linspace(d["Z_min"], d["Z_max"], nz)
    k = SimpleNamespace(R=R, Z=Z, NR=nr, NZ=nz, dR=float(R[1] - R[0]), dZ=float(Z[1] - Z[0]), Psi=z["psi"].copy(), device=0)
    coils = CoilSet(positions=[(c["r"], c["z"]) for c in cfg["coils"]], currents=z["currents0"].copy(),
                    turns=[1] * len(cfg["coils"]))
    return k, coils


def test_flux_sampling_and_shape_targets(fb):
    z = golden("free_boundary_shape")
    k, coils = _kernel_and_coils(z)
    np.testing.assert_allclose(fb.sample_flux_at_points(k, z["sample_pts"]), z["sample_psi"], rtol=1e-13, atol=1e-15)
    coils.target_flux_points = z["pts"]
    # the fixture's target was sampled from the flux map of the iteration before the last current update, so only
    # the isoflux property and the explicit-value path are checked against it
    t = fb.resolve_shape_target_flux(k, coils)
    assert t.shape == (z["pts"].shape[0],) and float(np.ptp(t)) == 0.0
    assert t[0] == np.mean([fb.interp_psi(k, r, zz) for r, zz in z["pts"]])
    coils.target_flux_values = z["explicit_targets"]
    np.testing.assert_array_equal(fb.resolve_shape_target_flux(k, coils), z["explicit_targets"])
    coils.target_flux_values = z["explicit_targets"][:-1]
    with pytest.raises(ValueError):
        fb.resolve_shape_target_flux(k, coils)
    coils.target_flux_points = None
    with pytest.raises(ValueError):
        fb.resolve_shape_target_flux(k, coils)


def test_bounded_fit_matches_reference_last_update(fb):
    """The last outer iteration of the fixture solved exactly this problem: target so_target, alpha, limits."""
    z = golden("free_boundary_shape")
    k, coils = _kernel_and_coils(z)
    coils.target_flux_points, coils.current_limits = z["pts"], z["limits"]
    got = fb.optimize_coil_currents(k, coils, z["so_target"], tikhonov_alpha=float(z["alpha"]))
    np.testing.assert_allclose(got, z["currents"], rtol=1e-9)
    M = fb.build_mutual_inductance_matrix(k, coils, z["pts"])
    np.testing.assert_allclose(M.T @ got, z["so_achieved"], rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        fb.optimize_coil_currents(k, coils, z["so_target"][:-1])
    with pytest.raises(ValueError):
        fb.optimize_coil_currents(k, coils, z["so_target"], tikhonov_alpha=float("nan"))
    coils.current_limits = z["limits"][:-1]
    with pytest.raises(ValueError):
        fb.optimize_coil_currents(k, coils, z["so_target"])


def test_probe_response_and_reconstruction(fb):
    z = golden("free_boundary_shape")
    k, coils = _kernel_and_coils(z)
    coils.current_limits = z["limits"]
    fl, bp, dirs = z["probe_flux_pts"], z["probe_b_pts"], [str(d) for d in z["probe_dirs"]]
    resp = fb.build_magnetic_probe_response_matrix(k, coils, flux_points=fl, b_probe_points=bp, b_probe_directions=dirs)
    np.testing.assert_array_equal(resp, z["probe_response"])   # same G values, same differencing arithmetic
    meas, nf = z["probe_meas"], fl.shape[0]
    rec = fb.reconstruct_coil_currents_from_magnetic_probes(
        k, coils, flux_points=fl, flux_measurements=meas[:nf], b_probe_points=bp, b_probe_directions=dirs,
        b_probe_measurements=meas[nf:], measurement_sigma=z["probe_sigma"], tikhonov_alpha=1e-6)
    np.testing.assert_allclose(rec["coil_currents"], z["probe_currents"], rtol=1e-10)
    np.testing.assert_allclose(rec["residual"], z["probe_residual"], rtol=1e-6, atol=1e-12)
    ps = z["probe_scalars"]
    assert (rec["response_rank"], rec["active_bounds"]) == (int(ps[2]), int(ps[4]))
    np.testing.assert_allclose([rec["residual_rms"], rec["weighted_residual_rms"], rec["response_condition"]],
                               [ps[0], ps[1], ps[3]], rtol=1e-6)
    for kw in ({"flux_points": fl}, {"flux_measurements": meas[:nf]},
               {"b_probe_points": bp, "b_probe_directions": dirs}, {"b_probe_measurements": meas[nf:]},
               {"flux_points": fl, "flux_measurements": meas[:nf], "measurement_sigma": -z["probe_sigma"][:nf]},
               {"flux_points": fl, "flux_measurements": meas[:nf], "tikhonov_alpha": -1.0}):
        with pytest.raises(ValueError):
            fb.reconstruct_coil_currents_from_magnetic_probes(k, coils, **kw)
    with pytest.raises(ValueError):
        fb.build_magnetic_probe_response_matrix(k, coils, b_probe_points=bp, b_probe_directions=["R", "Q"] + dirs[2:])
    with pytest.raises(ValueError):  # a Z-direction probe on the axis would difference across R <= 0
        fb.build_magnetic_probe_response_matrix(k, coils, b_probe_points=np.array([[0.0, 0.1]]), b_probe_directions=["Z"])


def test_contour_reconstruction_diagnostics(fb):
    z = golden("free_boundary_shape")
    k, coils = _kernel_and_coils(z)
    coils.currents = z["currents"].copy()   # the fixture's final currents
    pts = fb._kernel_boundary_points(k)
    np.testing.assert_array_equal(pts, z["br_points"])
    lim = np.array([[3.9, -4.6], [8.6, -4.6], [8.6, 4.6], [3.9, 4.6]])
    d = fb.reconstruct_boundary_flux_from_coils(k, coils, boundary_points=pts, limiter_points=lim, axis_point=np.array([6.2, 0.0]),
                                                x_points=np.array([[5.0, -3.4], [5.0, 3.4]]), target_flux=z["br_flux"])
    bs = z["br_scalars"]
    np.testing.assert_allclose(d["reconstructed_flux"], z["br_flux"], rtol=1e-12, atol=1e-15)
    assert (d["response_rank"], d["point_count"], d["coil_count"], d["limiter_point_count"], d["x_point_count"]) == \
        (int(bs[0]), int(bs[1]), int(bs[2]), int(bs[3]), int(bs[8]))
    assert d["min_limiter_distance_m"] == bs[4] and d["boundary_containment_fraction"] == bs[5]
    assert float(d["boundary_containment_pass"]) == bs[6]
    np.testing.assert_allclose([d["axis_flux"], d["x_point_flux_span"], d["x_point_pair_symmetry_abs_error"]],
                               [bs[7], bs[9], bs[10]], rtol=1e-10)
    np.testing.assert_allclose(d["limiter_flux"], z["br_limiter_flux"], rtol=1e-12)
    np.testing.assert_allclose(d["x_point_flux"], z["br_x_flux"], rtol=1e-12)
    assert d["max_abs_error"] <= 1e-12 and d["rmse"] <= 1e-12
    inside = fb._points_inside_polygon(np.array([[5.0, 0.0], [9.0, 0.0], [3.9, 10.0]]), lim)
    assert inside.tolist() == [True, False, False]
    with pytest.raises(ValueError):
        fb._points_inside_polygon(np.array([[5.0, 0.0]]), lim[:2])
    with pytest.raises(ValueError):
        fb.reconstruct_boundary_flux_from_coils(k, coils, boundary_points=pts, target_flux=z["br_flux"][:-1])


def test_green_function_argument_checks(fb):
    assert fb.green_function(6.2, 0.5, 4.0, -1.0, device=0) == golden("free_boundary_shape")["green_scalar"][0]
    for bad in ((0.0, 0.0, 1.0, 1.0), (1.0, 0.0, -1.0, 1.0), (1.0, float("nan"), 1.0, 1.0)):
        with pytest.raises(ValueError):
            fb.green_function(*bad, device=0)
