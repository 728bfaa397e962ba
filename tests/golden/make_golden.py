"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array written here is the output of a reference function called through
the reference's own API (``oracle/ref_shim.py`` only patches the numpy>=2 /
matplotlib import problems).  The tests never read
/root/reference; they read these .npz/.json files.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_shim  # noqa: E402

ref_shim.install()

from scpn_fusion.core import multigrid_solve as ref_mg  # noqa: E402
from scpn_fusion.core import _multi_compat_providers as ref_prov  # noqa: E402
from scpn_fusion.core.fusion_kernel import FusionKernel  # noqa: E402
from scpn_fusion.core import fusion_kernel_free_boundary as ref_fb  # noqa: E402

REF = ref_shim.REFERENCE_ROOT


def _save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


def _kernel(cfg):
    fd, p = tempfile.mkstemp(suffix=".json")
    with os.fdopen(fd, "w") as f:
        json.dump(cfg, f)
    k = FusionKernel(p)
    os.unlink(p)
    return k


def _cfg(name, n=None, **solver):
    cfg = json.load(open(os.path.join(REF, "validation", name)))
    cfg.pop("_reference", None)
    if n is not None:
        cfg["grid_resolution"] = [n[0], n[1]] if isinstance(n, (tuple, list)) else [n, n]
    cfg["solver"].update(solver)
    return cfg


# -- 1. element-wise multigrid operators on random fields, several shapes ------

def gen_ops():
    out = {}
    shapes = [(33, 33), (21, 37), (32, 32), (20, 36), (9, 9), (5, 5), (6, 7), (17, 9)]
    out["shapes"] = np.array(shapes)
    for i, (nz, nr) in enumerate(shapes):
        rng = np.random.default_rng(1000 + i)
        psi = rng.normal(size=(nz, nr))
        src = rng.normal(size=(nz, nr))
        r_axis = np.linspace(1.5, 4.25, nr)
        z_axis = np.linspace(-1.0, 2.0, nz)
        r_grid, _ = np.meshgrid(r_axis, z_axis)
        dr = float(r_axis[1] - r_axis[0])
        dz = float(z_axis[1] - z_axis[0])
        t = f"s{i}_"
        out[t + "psi"], out[t + "src"], out[t + "r_grid"] = psi, src, r_grid
        out[t + "drdz"] = np.array([dr, dz])
        out[t + "smooth_w13_2"] = ref_mg.mg_smooth(psi.copy(), src, r_grid, dr, dz, 1.3, 2)
        out[t + "smooth_w10_1"] = ref_mg.mg_smooth(psi.copy(), src, r_grid, dr, dz, 1.0, 1)
        out[t + "residual"] = ref_mg.mg_residual(psi, src, r_grid, dr, dz)
        out[t + "restrict"] = ref_mg.restrict_full_weight(psi)
        out[t + "restrict_r"] = ref_mg.restrict_full_weight(r_grid)
        coarse = rng.normal(size=((nz + 1) // 2, (nr + 1) // 2))
        out[t + "coarse"] = coarse
        out[t + "prolong"] = ref_mg.prolongate_bilinear(coarse, nz, nr)
        out[t + "vcycle_w10"] = ref_mg.multigrid_vcycle(psi.copy(), src, r_grid, dr, dz, omega=1.0)
        out[t + "vcycle_w16"] = ref_mg.multigrid_vcycle(psi.copy(), src, r_grid, dr, dz, omega=1.6)
        out[t + "linf"] = np.array(ref_mg.residual_linf(psi, src, r_grid, dr, dz))
    _save("mg_ops", **out)


# -- 2. multigrid_solve known answers ------------------------------------------

def gen_mg_solve():
    out = {}
    # the pinned 33^2 checksum case, benchmarks/bench_dispatcher_kernel_tiers.py:69-76
    nr = nz = 33
    r_min, r_max, z_min, z_max = 1.2, 2.2, -0.5, 0.5
    rr, zz = np.meshgrid(np.linspace(r_min, r_max, nr), np.linspace(z_min, z_max, nz))
    source = np.asarray(-rr * np.exp(-((rr - 1.7) ** 2 + zz ** 2) / 0.05), dtype=np.float64)
    psi, res, n, conv = ref_mg.multigrid_solve(source, np.zeros((nz, nr)), r_min, r_max, z_min,
                                               z_max, nr, nz, tol=1e-6, max_cycles=120)
    out["c33_source"], out["c33_psi"] = source, psi
    out["c33_meta"] = np.array([res, n, float(conv), float(np.sum(psi)) + res + n + float(conv)])
    # Gaussian source with a non-zero Dirichlet ring, rectangular and even grids
    for tag, (nz, nr) in {"g65": (65, 65), "g40x72": (40, 72), "g49x97": (49, 97)}.items():
        rng = np.random.default_rng(7)
        r_min, r_max, z_min, z_max = 4.0, 8.0, -4.0, 4.0
        rr, zz = np.meshgrid(np.linspace(r_min, r_max, nr), np.linspace(z_min, z_max, nz))
        source = -np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5)
        bc = 0.01 * rng.normal(size=(nz, nr))
        psi, res, n, conv = ref_mg.multigrid_solve(source, bc, r_min, r_max, z_min, z_max, nr, nz,
                                                   tol=1e-8, max_cycles=60)
        out[tag + "_source"], out[tag + "_bc"], out[tag + "_psi"] = source, bc, psi
        out[tag + "_meta"] = np.array([res, n, float(conv)])
    _save("mg_solve", **out)


# -- 3. the bench_gpu_gs_solver smoother problem --------------------------------

def gen_bench_smooth():
    out = {}
    for n, sweeps in ((65, 50), (129, 200)):
        rng = np.random.default_rng(2026)
        r_axis = np.linspace(4.0, 8.0, n)
        z_axis = np.linspace(-4.0, 4.0, n)
        r_grid, z_grid = np.meshgrid(r_axis, z_axis)
        source = -np.exp(-((r_grid - 6.0) ** 2 + z_grid ** 2) / 0.5)
        psi0 = rng.normal(0.0, 1e-3, size=(n, n))
        psi0[0, :] = psi0[-1, :] = psi0[:, 0] = psi0[:, -1] = 0.0
        sol = ref_prov._numpy_gs_rb_sor_smooth(psi0, source, 4.0, 8.0, -4.0, 4.0, omega=1.3,
                                               n_sweeps=sweeps)
        out[f"n{n}_out"] = sol
        out[f"n{n}_sweeps"] = np.array(sweeps)
    _save("bench_smooth", **out)


# -- 4. Picard pieces on real iterates ------------------------------------------

def gen_picard_pieces():
    out = {}
    cfg = _cfg("iter_validated_config.json", 65, max_iterations=40)
    k = _kernel(cfg)
    k.solve_equilibrium()
    psi = k.Psi.copy()
    out["psi"] = psi
    out["RZ"] = np.stack([k.R, k.Z])
    iz, ir, pax = k._find_magnetic_axis()
    out["axis"] = np.array([iz, ir, pax])
    (rx, zx), px = k.find_x_point(psi)
    out["xpoint"] = np.array([rx, zx, px])
    k.cfg["solver"]["xpoint_use_saddle_detection"] = True
    (rx, zx), px = k.find_x_point(psi)
    out["xpoint_saddle"] = np.array([rx, zx, px])
    k.cfg["solver"]["xpoint_use_saddle_detection"] = False
    pb = px if abs(pax - px) >= 0.1 else pax * 0.1
    out["axis_bnd"] = np.array([pax, pb])
    out["j_lmode"] = k.update_plasma_source_nonlinear(pax, pb).copy()
    k.profile_mode = "h-mode"
    k.ped_params_p.update({"ped_top": 0.9, "ped_width": 0.06, "ped_height": 1.1, "core_alpha": 0.25})
    out["j_hmode"] = k.update_plasma_source_nonlinear(pax, pb).copy()
    out["ped_p"] = np.array([0.9, 0.06, 1.1, 0.25])
    out["ped_ff"] = np.array([0.92, 0.05, 1.0, 0.3])
    k.profile_mode = "l-mode"
    src = -1.0 * k.RR * out["j_lmode"]
    out["jacobi"] = k._jacobi_step(psi, src)
    out["sor16"] = k._sor_step(psi, src, omega=1.6)
    out["gs_rms"] = np.array(k._compute_gs_residual_rms(src))
    k.compute_b_field()
    out["b_r"], out["b_z"] = k.B_R, k.B_Z
    out["vacuum"] = k.calculate_vacuum_field()
    coils = k.build_coilset_from_config()
    coils.turns = [1, 2, 1, 3, 1, 1, 4]
    out["ext_flux"] = ref_fb.compute_external_flux(k, coils)
    pts = np.array([[5.0, 0.5], [6.5, -2.0], [3.9, 7.6], [8.0, 0.0]])
    out["mutual_pts"] = pts
    out["mutual"] = ref_fb.build_mutual_inductance_matrix(k, coils, pts)
    out["turns"] = np.array(coils.turns)
    _save("picard_pieces", **out)


# -- 5. full solve_equilibrium runs ---------------------------------------------

def _solve_case(out, tag, cfg, keep_psi=True):
    k = _kernel(cfg)
    r = k.solve_equilibrium()
    (rx, zx), px = k.find_x_point(k.Psi)
    iz, ir, pax = k._find_magnetic_axis()
    if keep_psi:
        out[tag + "_psi"] = r["psi"]
        out[tag + "_jphi"] = k.J_phi
    out[tag + "_hist"] = np.array(r["residual_history"])
    out[tag + "_gshist"] = np.array(r["gs_residual_history"])
    out[tag + "_meta"] = np.array([r["iterations"], float(r["converged"]), r["residual"],
                                   r["gs_residual"], r["gs_residual_best"]])
    out[tag + "_topo"] = np.array([k.R[ir], k.Z[iz], pax, rx, zx, px])
    out[tag + "_cfg"] = np.array(json.dumps(cfg))
    print(f"  {tag}: iters={r['iterations']} conv={r['converged']} t={r['wall_time_s']:.2f}s "
          f"axis=({k.R[ir]:.3f},{k.Z[iz]:.3f}) x=({rx:.3f},{zx:.3f})")
    return k


def gen_solves():
    out = {}
    _solve_case(out, "iter65", _cfg("iter_config.json", 65))
    _solve_case(out, "iter129", _cfg("iter_config.json", 129))
    _solve_case(out, "iter64", _cfg("iter_config.json", 64))
    _solve_case(out, "iter48x80", _cfg("iter_config.json", (48, 80)))
    _solve_case(out, "iterval65", _cfg("iter_validated_config.json", 65))
    _solve_case(out, "diiid65", _cfg("diiid_config.json", 65))
    _solve_case(out, "diiid65s", _cfg("diiid_config.json", 65, xpoint_use_saddle_detection=True))
    _solve_case(out, "iter65sor", _cfg("iter_config.json", 65, solver_method="sor", max_iterations=60))
    _solve_case(out, "iter65jac", _cfg("iter_config.json", 65, solver_method="jacobi", max_iterations=60))
    _solve_case(out, "iter65gs", _cfg("iter_config.json", 65, require_gs_residual=True,
                                       gs_residual_threshold=5e-3, max_iterations=200))
    # H-mode UQ samples, recipe of SURVEY.md 8d config 3 / tools/parallel_gen_iter.py:96-101
    for ksample in range(3):
        cfg = _cfg("iter_config.json", 65)
        rng = np.random.default_rng(2026 + ksample)
        for c in cfg["coils"]:
            c["current"] = c["current"] * float(rng.uniform(0.85, 1.15))
        cfg["physics"]["plasma_current_target"] *= float(rng.uniform(0.8, 1.2))
        ped = {"ped_top": 0.92 * float(rng.uniform(0.97, 1.03)),
               "ped_width": 0.05 * float(rng.uniform(0.9, 1.1)),
               "ped_height": 1.0 * float(rng.uniform(0.9, 1.1)),
               "core_alpha": 0.3 * float(rng.uniform(0.9, 1.1))}
        cfg["physics"]["profiles"] = {"mode": "h-mode", "p_prime": ped, "ff_prime": dict(ped)}
        _solve_case(out, f"uq65_{ksample}", cfg)
    _save("solves", **out)


def gen_solve_129_validated():
    out = {}
    _solve_case(out, "iterval129", _cfg("iter_validated_config.json", 129))
    _save("solve_iterval129", **out)


# -- 6. free boundary -------------------------------------------------------------

def gen_free_boundary():
    out = {}
    cfg = _cfg("iter_validated_config.json", 33)
    k = _kernel(cfg)
    coils = k.build_coilset_from_config()
    coils.currents = coils.currents * 1.0e6  # SI amperes so the SI-mu0 wall flux is O(1)
    res = k.solve_free_boundary(coils, max_outer_iter=4, tol=1e-4)
    out["psi"] = k.Psi
    out["meta"] = np.array([res["outer_iterations"], res["final_diff"]])
    out["currents"] = coils.currents
    out["cfg"] = np.array(json.dumps(cfg))
    print("  free boundary:", res["outer_iterations"], res["final_diff"])
    _save("free_boundary", **out)


SHAPE_ALPHA = 1e-13  # currents are SI amperes (1e6-1e7): the default 1e-4 would regularise them to ~0


def _shape_setup():
    """Shared by the generator and by the tests (tests read the same arrays from the fixture)."""
    th = np.linspace(0.0, 2.0 * np.pi, 9)[:-1] + 0.2
    pts = np.column_stack([6.2 + 1.9 * np.cos(th), 3.1 * np.sin(th)])
    limits = np.array([2.0e7, 6.0e6, 6.0e6, 6.0e6, 6.0e6, 2.0e7, 2.0e7])
    return pts, limits


def gen_free_boundary_shape():
    """solve_free_boundary(optimize_shape=True) plus the coil-diagnostic functions built on the same
    Green's function (fusion_kernel_free_boundary.py:156-739)."""
    out = {}
    cfg = _cfg("iter_validated_config.json", 33)
    k = _kernel(cfg)
    coils = k.build_coilset_from_config()
    coils.currents = coils.currents * 1.0e6
    pts, limits = _shape_setup()
    coils.target_flux_points = pts
    coils.current_limits = limits
    out["currents0"], out["pts"], out["limits"] = coils.currents.copy(), pts, limits
    # (i) isoflux target inferred from the solution, bounded currents
    res = k.solve_free_boundary(coils, max_outer_iter=3, tol=1e-4, optimize_shape=True, tikhonov_alpha=SHAPE_ALPHA,
                                limiter_points=np.array([[3.9, -4.6], [8.6, -4.6], [8.6, 4.6], [3.9, 4.6]]),
                                axis_point=np.array([6.2, 0.0]), x_points=np.array([[5.0, -3.4], [5.0, 3.4]]))
    so, br = res["shape_optimization"], res["boundary_reconstruction"]
    out["psi"], out["currents"] = k.Psi.copy(), res["coil_currents"]
    out["meta"] = np.array([res["outer_iterations"], res["final_diff"], res["vacuum_boundary_abs_error"]])
    out["so_scalars"] = np.array([so["target_point_count"], so["coil_count"], so["response_rank"], so["response_condition"],
                                  so["flux_rmse"], so["flux_relative_rmse"], so["max_abs_flux_residual"],
                                  so["active_current_bounds"]], dtype=np.float64)
    out["so_target"], out["so_achieved"] = so["target_flux"], so["achieved_flux"]
    out["br_points"], out["br_flux"] = br["boundary_points"], br["reconstructed_flux"]
    out["br_scalars"] = np.array([br["response_rank"], br["point_count"], br["coil_count"], br["limiter_point_count"],
                                  br["min_limiter_distance_m"], br["boundary_containment_fraction"],
                                  float(br["boundary_containment_pass"]), br["axis_flux"], br["x_point_count"],
                                  br["x_point_flux_span"], br["x_point_pair_symmetry_abs_error"], br["rmse"],
                                  br["max_abs_error"]], dtype=np.float64)
    out["br_limiter_flux"], out["br_x_flux"] = br["limiter_flux"], br["x_point_flux"]
    out["alpha"] = np.array(SHAPE_ALPHA)
    print("  shape:", res["outer_iterations"], res["final_diff"], so["flux_rmse"], so["active_current_bounds"], res["coil_currents"])
    # (ii) explicit target values, one outer iteration, unbounded
    k2 = _kernel(cfg)
    c2 = k2.build_coilset_from_config()
    c2.currents = out["currents0"].copy()
    c2.target_flux_points = pts
    c2.target_flux_values = np.linspace(-0.4, 0.3, pts.shape[0])
    r2 = k2.solve_free_boundary(c2, max_outer_iter=1, tol=0.0, optimize_shape=True, tikhonov_alpha=1e-6)
    out["explicit_targets"], out["explicit_currents"] = c2.target_flux_values, r2["coil_currents"]
    out["explicit_psi"] = k2.Psi.copy()
    # (iii) interpolation and point sampling of the final flux map
    sample = np.array([[2.0, -6.0], [10.0, 6.0], [6.123, 0.456], [9.99, -5.99], [1.0, 0.0], [12.0, 7.0], [4.25, 3.0]])
    out["sample_pts"] = sample
    out["sample_psi"] = ref_fb.sample_flux_at_points(k, sample)
    # (iv) magnetic-probe response and bounded inverse reconstruction
    c3 = k.build_coilset_from_config()
    c3.currents = out["currents0"].copy()
    c3.current_limits = limits
    fl = np.array([[3.0, 0.0], [4.0, 3.5], [8.5, 2.0], [8.8, -1.5], [4.2, -3.6], [6.0, 4.4]])
    bp = np.array([[3.2, 1.0], [3.2, -1.0], [8.9, 0.5], [6.5, 4.2], [6.5, -4.2], [0.0, 0.3]])
    bd = ["R", "Z", "z", "r", "Z", "R"]
    resp = ref_fb.build_magnetic_probe_response_matrix(k, c3, flux_points=fl, b_probe_points=bp, b_probe_directions=bd)
    truth = out["currents0"] * np.array([1.1, 0.9, 1.0, 1.0, 1.2, 0.8, 1.05]) + np.array([0, 0, 2e5, -3e5, 0, 0, 0])
    meas = resp @ truth
    sigma = np.concatenate([np.full(len(fl), 1e-3), np.full(len(bp), 2e-3)])
    rec = ref_fb.reconstruct_coil_currents_from_magnetic_probes(
        k, c3, flux_points=fl, flux_measurements=meas[:len(fl)], b_probe_points=bp, b_probe_directions=bd,
        b_probe_measurements=meas[len(fl):], measurement_sigma=sigma, tikhonov_alpha=1e-6)
    out["probe_flux_pts"], out["probe_b_pts"], out["probe_dirs"] = fl, bp, np.array(bd)
    out["probe_response"], out["probe_meas"], out["probe_sigma"] = resp, meas, sigma
    out["probe_currents"], out["probe_residual"] = rec["coil_currents"], rec["residual"]
    out["probe_scalars"] = np.array([rec["residual_rms"], rec["weighted_residual_rms"], rec["response_rank"],
                                     rec["response_condition"], rec["active_bounds"]], dtype=np.float64)
    out["green_scalar"] = np.array([ref_fb.green_function(6.2, 0.5, 4.0, -1.0), ref_fb.green_function(6.2, 0.5, 6.2, 0.5),
                                    ref_fb.green_function(1.7, 0.0, 9.0, 5.0)])
    out["cfg"] = np.array(json.dumps(cfg))
    _save("free_boundary_shape", **out)


def gen_dataset():
    """tools/parallel_gen_iter.py::generate_chunk - the reference's UQ/dataset producer."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_parallel_gen_iter", os.path.join(REF, "tools", "parallel_gen_iter.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    cases = [("val65", "iter_validated_config.json", 65, 2, 42, False), ("val", "iter_validated_config.json", 33, 3, 42, True),
             ("iter", "iter_config.json", 33, 2, 43, False), ("iter_allow", "iter_config.json", 33, 2, 43, True)]
    for tag, name, grid, n, seed, allow in cases:
        cfg = _cfg(name, grid)
        fd, path = tempfile.mkstemp(suffix=".json")
        with os.fdopen(fd, "w") as f:
            json.dump(cfg, f)
        X, Y, rej, failed = mod.generate_chunk(n, path, seed, allow)
        os.unlink(path)
        out[tag + "_cfg"] = np.array(json.dumps(cfg))
        out[tag + "_X"], out[tag + "_Y"] = X, Y
        out[tag + "_meta"] = np.array([n, seed, int(allow), rej, failed])
        print(f"  {tag}: X{X.shape} Y{Y.shape} rejected={rej} failed={failed}")
    out["boundary_checks"] = np.array([mod.is_boundary_xpoint(2.05, 0.0, 2.0, 10.0, -6.0, 6.0),
                                       mod.is_boundary_xpoint(2.09, 0.0, 2.0, 10.0, -6.0, 6.0),
                                       mod.is_boundary_xpoint(5.0, -5.87, 2.0, 10.0, -6.0, 6.0),
                                       mod.is_boundary_xpoint(5.0, -5.89, 2.0, 10.0, -6.0, 6.0)])
    _save("dataset", **out)


def _fortran_body(values, width=16, prec=9, dexp=False):
    """Fixed-width Fortran (5e16.9) text with NO separating blanks: negative values run into their neighbours."""
    cells = []
    for v in values:
        c = f"{v:{width}.{prec}E}"
        cells.append(c.replace("E", "D") if dexp else c)
    return "".join("".join(cells[i:i + 5]).lstrip() + "\n" for i in range(0, len(cells), 5))


def gen_eqdsk():
    """core/eqdsk.py: writer bytes, reader values (free-format, run-together fixed-width, D exponents),
    to_config(), the public-SPARC adapter and the rejected inputs."""
    from scpn_fusion.core import eqdsk as ref_eq
    out = {}
    rng = np.random.default_rng(77)
    nw, nh, nb, nl = 9, 7, 7, 5
    eq = ref_eq.GEqdsk(description="synthetic case for the B200 reader/writer parity test, long description", nw=nw, nh=nh,
                       rdim=2.5, zdim=4.0, rcentr=1.85, rleft=0.6, zmid=0.1, rmaxis=1.9, zmaxis=0.05, simag=-0.42,
                       sibry=0.13, bcentr=-12.2, current=8.7e6, fpol=rng.normal(size=nw), pres=rng.uniform(0, 1e5, nw),
                       ffprime=rng.normal(size=nw), pprime=rng.normal(size=nw) * 1e3, qpsi=rng.uniform(1, 5, nw),
                       psirz=rng.normal(size=(nh, nw)), rbdry=rng.uniform(1, 2.5, nb), zbdry=rng.uniform(-1, 1, nb),
                       rlim=rng.uniform(0.7, 3.0, nl), zlim=rng.uniform(-1.9, 2.0, nl))
    tmp = tempfile.mkdtemp()
    pa = os.path.join(tmp, "a.geqdsk")
    ref_eq.write_geqdsk(eq, pa)
    out["a_bytes"] = np.frombuffer(open(pa, "rb").read(), dtype=np.uint8)
    back = ref_eq.read_geqdsk(pa)
    for n in ("fpol", "pres", "ffprime", "pprime", "qpsi", "psirz", "rbdry", "zbdry", "rlim", "zlim"):
        out["a_" + n] = getattr(back, n)
        np.testing.assert_array_equal(getattr(back, n), getattr(eq, n))  # 17 significant digits round-trip
    out["a_scalars"] = np.array([getattr(back, n) for n in ("rdim", "zdim", "rcentr", "rleft", "zmid", "rmaxis", "zmaxis",
                                                             "simag", "sibry", "bcentr", "current")])
    out["a_desc"] = np.array(back.description)
    out["a_config"] = np.array(json.dumps(back.to_config("case_a")))
    out["a_r"], out["a_z"] = back.r, back.z
    out["a_psin"] = back.psi_to_norm(back.psirz)
    # no contours at all, and an odd pair count (line-break rule of the writer)
    for tag, nb2, nl2 in (("b", 0, 0), ("c", 3, 1)):
        e2 = ref_eq.GEqdsk(**{**eq.__dict__, "rbdry": eq.rbdry[:nb2], "zbdry": eq.zbdry[:nb2], "rlim": eq.rlim[:nl2],
                              "zlim": eq.zlim[:nl2], "description": "short"})
        p2 = os.path.join(tmp, tag + ".geqdsk")
        ref_eq.write_geqdsk(e2, p2)
        out[tag + "_bytes"] = np.frombuffer(open(p2, "rb").read(), dtype=np.uint8)
        out[tag + "_config"] = np.array(json.dumps(ref_eq.read_geqdsk(p2).to_config()))
    # fixed-width body without blanks (+ D exponents), as legacy EFIT writers produce
    scal = [2.5, 4.0, 1.85, 0.6, -0.1, 1.9, -0.05, -0.42, 0.13, -12.2, -8.7e6, -0.42, 0.0, 1.9, 0.0, -0.05, 0.0, 0.13, 0.0, 0.0]
    vals = np.concatenate([eq.fpol, eq.pres, -np.abs(eq.ffprime), eq.pprime, -np.abs(eq.psirz).ravel(), eq.qpsi])
    for tag, dexp in (("f", False), ("d", True)):
        text = f"  EFITD    01/01/2026    #123456  1000ms           3 {nw} {nh}\n" + _fortran_body(scal, dexp=dexp) \
            + _fortran_body(vals, dexp=dexp) + f"{nb:5d}{nl:5d}\n" \
            + _fortran_body(np.column_stack([eq.rbdry, -np.abs(eq.zbdry)]).ravel(), dexp=dexp) \
            + _fortran_body(np.column_stack([eq.rlim, eq.zlim]).ravel(), dexp=dexp)
        pf = os.path.join(tmp, tag + ".geqdsk")
        open(pf, "w").write(text)
        e = ref_eq.read_geqdsk(pf)
        out[tag + "_bytes"] = np.frombuffer(text.encode(), dtype=np.uint8)
        out[tag + "_desc"] = np.array(e.description)
        out[tag + "_scalars"] = np.array([getattr(e, n) for n in ("rdim", "zdim", "rcentr", "rleft", "zmid", "rmaxis", "zmaxis",
                                                                   "simag", "sibry", "bcentr", "current")])
        for n in ("fpol", "pres", "ffprime", "pprime", "qpsi", "psirz", "rbdry", "zbdry", "rlim", "zlim"):
            out[tag + "_" + n] = getattr(e, n)
    # public SPARC named adapter
    ps = os.path.join(tmp, "sparc_1305.eqdsk")
    ref_eq.write_geqdsk(eq, ps)
    e = ref_eq.read_geqdsk(ps, source_convention_mode="public_sparc_named_adapter")
    out["sparc_ffprime"], out["sparc_pprime"] = e.ffprime, e.pprime
    out["sparc_meta"] = np.array(json.dumps({"convention": e.source_convention, "adapter": e.source_convention_adapter,
                                             "ok": e.source_convention_adapter_pass, "meta": e.source_convention_metadata}))
    e = ref_eq.read_geqdsk(pa, source_convention_mode="public_sparc_named_adapter")
    out["nomatch_meta"] = np.array(json.dumps({"convention": e.source_convention, "adapter": e.source_convention_adapter,
                                               "ok": e.source_convention_adapter_pass, "meta": e.source_convention_metadata}))
    # inputs the reference rejects with ValueError
    good = open(pa).read()
    lines = good.splitlines(keepends=True)
    bad = {"empty": "", "short_header": "x 9\n", "tiny_grid": lines[0][:-10] + "    1    7\n" + "".join(lines[1:]),
           "truncated": "".join(lines[:12]), "nan_token": good.replace(lines[3][:24], "                     nan", 1),
           "huge_grid": lines[0][:-10] + " 2000 2000\n" + "".join(lines[1:]),
           "neg_count": good.replace(f"{nb:5d}{nl:5d}\n", f"{-1:5d}{nl:5d}\n"),
           "equal_psi": None}
    e3 = ref_eq.GEqdsk(**{**eq.__dict__, "sibry": eq.simag})
    p3 = os.path.join(tmp, "eq.geqdsk")
    ref_eq.write_geqdsk(e3, p3)
    bad["equal_psi"] = open(p3).read()
    names = []
    for k, text in bad.items():
        pb = os.path.join(tmp, "bad_" + k)
        open(pb, "w").write(text)
        try:
            ref_eq.read_geqdsk(pb)
        except ValueError as exc:
            names.append(k)
            out["bad_" + k] = np.frombuffer(text.encode(), dtype=np.uint8)
            print("   rejects", k, "->", str(exc)[:60])
        else:
            print("   ACCEPTS", k)
    out["bad_names"] = np.array(names)
    _save("eqdsk", **out)


def gen_anderson():
    """solver_method="anderson" (SOR inner sweep + Anderson mixing every third iterate)."""
    out = {}
    for tag, name, n, kw in (("val33", "iter_validated_config.json", 33, {}),
                             ("iter49", "iter_config.json", 49, {"anderson_depth": 3, "max_iterations": 200})):
        cfg = _cfg(name, n, solver_method="anderson", **kw)
        k = _kernel(cfg)
        r = k.solve_equilibrium()
        out[tag + "_cfg"] = np.array(json.dumps(cfg))
        out[tag + "_psi"] = k.Psi.copy()
        out[tag + "_meta"] = np.array([r["iterations"], float(r["converged"]), r["residual"], r["gs_residual"]])
        out[tag + "_hist"] = np.array(r["residual_history"])
        out[tag + "_gshist"] = np.array(r["gs_residual_history"])
        print(f"  {tag}: its={r['iterations']} conv={r['converged']} res={r['residual']:.3e}")
    _save("anderson", **out)


def gen_solovev():
    """validation/validate_grad_shafranov_solovev.py re-run here (operator truncation error, SOR reconstruction,
    NumPy-tier multigrid reconstruction of psi = c1 R^4/8 + c2 Z^2) next to the numbers of the reference's sealed
    report validation/reports/grad_shafranov_solovev.json."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_solovev", os.path.join(REF, "validation", "validate_grad_shafranov_solovev.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_solovev"] = mod
    spec.loader.exec_module(mod)
    geo = mod.SolovevGeometry.from_aspect()
    sealed = json.load(open(os.path.join(REF, "validation", "reports", "grad_shafranov_solovev.json")))
    out = {"geometry": {k: getattr(geo, k) for k in ("r0", "a", "r_min", "r_max", "z_min", "z_max", "c1", "c2")},
           "sealed": {k: sealed[k] for k in ("operator_records", "operator_order", "reconstruction_records",
                                             "multigrid_numpy_record", "operator_error_gate", "reconstruction_nrmse_gate")}}
    out["operator_errors"] = {str(n): mod.operator_truncation_error(geo, n) for n in (33, 49, 65, 97)}
    out["sor"] = {}
    for n in (33, 49):
        r = mod.sor_reconstruction(geo, n)
        out["sor"][str(n)] = {"nrmse": r.nrmse, "iterations": r.iterations, "converged": r.converged, "residual_inf": r.residual_inf}
    m = mod.dispatched_multigrid_reconstruction(geo, 97, tier="numpy", analytic_tolerance=1e-4)
    out["multigrid_numpy_97"] = {"nrmse": m.nrmse, "residual": m.residual, "cycles": m.cycles, "converged": m.converged}
    for n in (33, 49, 65, 97):
        sealed_err = [r["error"] for r in sealed["operator_records"] if r["resolution"] == n][0]
        print(f"  operator {n}: here {out['operator_errors'][str(n)]:.16e} sealed {sealed_err:.16e}")
    print("  sor:", out["sor"], "\n  mg:", out["multigrid_numpy_97"], "sealed", sealed["multigrid_numpy_record"])
    path = os.path.join(HERE, "solovev.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote solovev.json")


# -- 7. the reference's compiled C++ solver (hpc/solver.cpp) ----------------------

def gen_hpc():
    so = os.path.join(ROOT, "oracle", "_ref", "libscpn_solver.so")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    lib = ctypes.CDLL(so)
    lib.create_solver.restype = ctypes.c_void_p
    lib.create_solver.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4
    dp = ctypes.POINTER(ctypes.c_double)
    lib.run_step.argtypes = [ctypes.c_void_p, dp, dp, ctypes.c_int, ctypes.c_int]
    lib.run_step_converged.argtypes = [ctypes.c_void_p, dp, dp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_double, dp]
    lib.run_step_converged.restype = ctypes.c_int
    lib.set_boundary_dirichlet.argtypes = [ctypes.c_void_p, ctypes.c_double]
    lib.destroy_solver.argtypes = [ctypes.c_void_p]
    out = {}
    nz, nr = 33, 41
    rng = np.random.default_rng(5)
    j = rng.normal(size=(nz, nr))
    h = lib.create_solver(nr, nz, 2.0, 10.0, -4.0, 4.0)
    psi = np.zeros((nz, nr))
    lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 7)
    out["j"], out["psi_7"] = j, psi.copy()
    lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 5)  # warm: 12 total
    out["psi_12"] = psi.copy()
    lib.set_boundary_dirichlet(h, 0.25)
    delta = ctypes.c_double(0.0)
    n = lib.run_step_converged(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), nz * nr, 400, 1.5,
                               1e-9, ctypes.byref(delta))
    out["psi_conv"] = psi.copy()
    out["conv_meta"] = np.array([n, delta.value])
    lib.destroy_solver(h)
    _save("hpc_solver", **out)


def gen_external():
    """external_profile_mode (the transport-coupling caller, _integrated_transport_solver_init.py:28): the Picard
    loop does not update J_phi from psi (fusion_kernel_newton_solver.py:509); _seed_plasma still overwrites it."""
    out = {}
    for tag, name, n in (("iter65x", "iter_config.json", 65), ("iterval33x", "iter_validated_config.json", 33)):
        cfg = _cfg(name, n)
        k = _kernel(cfg)
        k.external_profile_mode = True
        k.J_phi = np.ones_like(k.Psi)  # what the external caller put there; the seed replaces it
        r = k.solve_equilibrium()
        out[tag + "_cfg"] = np.array(json.dumps(cfg))
        out[tag + "_psi"], out[tag + "_jphi"] = k.Psi, k.J_phi
        out[tag + "_meta"] = np.array([r["iterations"], float(r["converged"]), r["residual"], r["gs_residual"]])
        out[tag + "_hist"] = np.array(r["residual_history"])
        print(" ", tag, r["iterations"], r["converged"], r["residual"])
    _save("solve_external", **out)


def gen_mg_4097():
    """BASELINE configs[4] / SURVEY.md 8d config 5: the reference's multigrid_solve on the 4097^2 bench problem
    (bench_gpu_gs_solver._problem source, psi_bc = 0, tol 1e-8, omega 1.0, 3/3, min_grid 5; 13 levels).  Run once
    (minutes of NumPy time); the fixture keeps the scalars, sum(psi), a 16-row stripe through the source peak and
    a 32-strided sample of the whole field."""
    n = 4097
    r_axis = np.linspace(4.0, 8.0, n)
    z_axis = np.linspace(-4.0, 4.0, n)
    rr, zz = np.meshgrid(r_axis, z_axis)
    src = -np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5)
    del rr, zz
    psi, res, cycles, conv = ref_mg.multigrid_solve(src, np.zeros((n, n)), 4.0, 8.0, -4.0, 4.0, n, n, tol=1e-8,
                                                   max_cycles=50)
    _save("mg_solve_4097", meta=np.array([res, cycles, float(conv), float(np.sum(psi))]),
          stripe_rows=np.array([2040, 2056]), stripe=psi[2040:2056].copy(), strided=psi[::32, ::32].copy())


def gen_elliptic():
    tab = json.load(open(os.path.join(REF, "scpn-fusion-rs", "tests", "reference", "reference_elliptic.json")))
    m = np.array([float(x) for x in tab["K"].keys()])
    _save("elliptic", m=m, K=np.array(list(tab["K"].values())), E=np.array(list(tab["E"].values())))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "mg_solve", "bench_smooth", "picard_pieces", "solves",
                             "solve_129_validated", "free_boundary", "free_boundary_shape", "dataset", "eqdsk", "anderson", "solovev", "hpc", "elliptic"]
    # "mg_4097" is generated on request only (several minutes of NumPy time)
    for w in which:
        print("==", w)
        globals()["gen_" + w]()
