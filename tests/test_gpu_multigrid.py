"""GPU parity: multigrid operators and multigrid_solve vs the oracle and the golden fixtures.

Bit-exact: the kernels keep NumPy's operand order without FMA contraction, and the only
divisions are by per-level constants evaluated as correctly rounded quotients.
"""
from __future__ import annotations

import numpy as np
import pytest

import gs_oracle as G
from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import scpn_fusion_core_b200 as pkg
    return pkg


def _problem(nz, nr, seed, rmin=1.5, rmax=4.25):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=(nz, nr))
    src = rng.normal(size=(nz, nr))
    r_axis = np.linspace(rmin, rmax, nr)
    z_axis = np.linspace(-1.0, 2.0, nz)
    rg, _ = np.meshgrid(r_axis, z_axis)
    return psi, src, rg, float(r_axis[1] - r_axis[0]), float(z_axis[1] - z_axis[0])


def test_operators_match_golden_bit_exact(mg):
    z = golden("mg_ops")
    for i, (nz, nr) in enumerate(z["shapes"]):
        t = f"s{i}_"
        psi, src, rg = z[t + "psi"], z[t + "src"], z[t + "r_grid"]
        dr, dz = z[t + "drdz"]
        np.testing.assert_array_equal(mg.mg_smooth(psi.copy(), src, rg, dr, dz, 1.3, 2), z[t + "smooth_w13_2"])
        np.testing.assert_array_equal(mg.mg_smooth(psi.copy(), src, rg, dr, dz, 1.0, 1), z[t + "smooth_w10_1"])
        np.testing.assert_array_equal(mg.mg_residual(psi, src, rg, dr, dz), z[t + "residual"])
        np.testing.assert_array_equal(mg.restrict_full_weight(psi), z[t + "restrict"])
        np.testing.assert_array_equal(mg.prolongate_bilinear(z[t + "coarse"], int(nz), int(nr)), z[t + "prolong"])
        np.testing.assert_array_equal(mg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.0), z[t + "vcycle_w10"])
        np.testing.assert_array_equal(mg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.6), z[t + "vcycle_w16"])
        assert mg.residual_linf(psi, src, rg, dr, dz) == float(z[t + "linf"])


@pytest.mark.parametrize("shape", [(129, 129), (128, 128), (65, 97), (7, 200), (200, 7), (3, 3), (4, 4), (257, 257)])
def test_operators_match_oracle_on_seeded_inputs(mg, shape):
    nz, nr = shape
    psi, src, rg, dr, dz = _problem(nz, nr, 11)
    np.testing.assert_array_equal(mg.mg_smooth(psi.copy(), src, rg, dr, dz, 1.6, 3),
                                  G.rb_sor_smooth(psi.copy(), src, rg, dr, dz, 1.6, 3))
    np.testing.assert_array_equal(mg.mg_residual(psi, src, rg, dr, dz), G.gs_residual(psi, src, rg, dr, dz))
    np.testing.assert_array_equal(mg.restrict_full_weight(psi), G.restrict_full_weight(psi))
    coarse = np.random.default_rng(3).normal(size=((nz + 1) // 2, (nr + 1) // 2))
    np.testing.assert_array_equal(mg.prolongate_bilinear(coarse, nz, nr), G.prolong_bilinear(coarse, nz, nr))
    np.testing.assert_array_equal(mg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.6),
                                  G.vcycle(psi.copy(), src, rg, dr, dz, omega=1.6))
    assert mg.residual_linf(psi, src, rg, dr, dz) == G.residual_linf(psi, src, rg, dr, dz)


def test_vcycle_options_and_batch(mg):
    import torch
    nz, nr = 65, 49
    fields = [_problem(nz, nr, s) for s in range(5)]
    rg, dr, dz = fields[0][2], fields[0][3], fields[0][4]
    psi = torch.tensor(np.stack([f[0] for f in fields]), device="cuda")
    src = torch.tensor(np.stack([f[1] for f in fields]), device="cuda")
    out = mg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.2, pre_smooth=2, post_smooth=1, min_grid=9)
    assert out.shape == (5, nz, nr) and out.is_cuda
    for b, f in enumerate(fields):
        ref = G.vcycle(f[0].copy(), f[1], rg, dr, dz, omega=1.2, pre=2, post=1, min_grid=9)
        np.testing.assert_array_equal(out[b].cpu().numpy(), ref)
    np.testing.assert_array_equal(psi[0].cpu().numpy(), fields[0][0])  # input untouched


def test_smoother_fixed_point_and_wall(mg):
    """tests/test_fusion_kernel_solver_mixins.py (reference): a sweep leaves the discrete solution fixed."""
    nz = nr = 33
    r = np.linspace(1.0, 2.0, nr)
    zc = np.linspace(-0.5, 0.5, nz)
    rr, zz = np.meshgrid(r, zc)
    psi = 0.03125 * rr ** 4 - 0.125 * zz ** 2 + 0.05 * rr ** 2 * zz ** 2
    dr, dz = float(r[1] - r[0]), float(zc[1] - zc[0])
    src = G.gs_operator(psi, rr, dr, dz)
    out = mg.mg_smooth(psi.copy(), src, rr, dr, dz, 1.0, 5)
    assert np.max(np.abs(out - psi)) < 1e-13
    for sl in (np.s_[0, :], np.s_[-1, :], np.s_[:, 0], np.s_[:, -1]):
        np.testing.assert_array_equal(out[sl], psi[sl])


def test_omega_rejection(mg):
    s = np.zeros((9, 9))
    for bad in (0.99, 2.0, float("nan"), float("inf")):
        with pytest.raises(ValueError):
            mg.mg_smooth(s.copy(), s, np.ones((9, 9)), 0.1, 0.1, bad, 1)
        with pytest.raises(ValueError):
            mg.multigrid_vcycle(s, s, np.ones((9, 9)), 0.1, 0.1, omega=bad)


def test_mg_solve_pinned_checksum(mg):
    """validation/reports/dispatcher_kernel_tiers_benchmark.json: output_checksum 22.46587229431361."""
    z = golden("mg_solve")
    psi, res, n, conv = mg.multigrid_solve(z["c33_source"], np.zeros((33, 33)), 1.2, 2.2, -0.5, 0.5, 33, 33,
                                           tol=1e-6, max_cycles=120)
    np.testing.assert_array_equal(psi, z["c33_psi"])
    assert (res, n, conv) == (z["c33_meta"][0], 5, True)
    assert abs(float(np.sum(psi)) + res + n + 1.0 - 22.46587229431361) < 1e-12


@pytest.mark.parametrize("tag,shape", [("g65", (65, 65)), ("g40x72", (40, 72)), ("g49x97", (49, 97))])
def test_mg_solve_golden_cases(mg, tag, shape):
    z = golden("mg_solve")
    nz, nr = shape
    bc = z[tag + "_bc"]
    psi, res, n, conv = mg.multigrid_solve(z[tag + "_source"], bc, 4.0, 8.0, -4.0, 4.0, nr, nz, tol=1e-8, max_cycles=60)
    np.testing.assert_array_equal(psi, z[tag + "_psi"])
    assert (res, n, float(conv)) == tuple(z[tag + "_meta"])
    for sl in (np.s_[0, :], np.s_[-1, :], np.s_[:, 0], np.s_[:, -1]):  # Dirichlet ring preserved exactly
        np.testing.assert_array_equal(psi[sl], bc[sl])


def test_mg_solve_batch_with_different_cycle_counts(mg):
    import torch
    nz = nr = 33
    rr, zz = np.meshgrid(np.linspace(1.2, 2.2, nr), np.linspace(-0.5, 0.5, nz))
    base = -rr * np.exp(-((rr - 1.7) ** 2 + zz ** 2) / 0.05)
    scales = [1.0, 1e-3, 0.0, 50.0]
    src = np.stack([s * base for s in scales])
    psi, res, cyc, conv = mg.multigrid_solve(torch.tensor(src, device="cuda"), torch.zeros((4, nz, nr), dtype=torch.float64,
                                             device="cuda"), 1.2, 2.2, -0.5, 0.5, nr, nz, tol=1e-6, max_cycles=120)
    for b, s in enumerate(scales):
        p, r, n, c = G.mg_solve(s * base, np.zeros((nz, nr)), 1.2, 2.2, -0.5, 0.5, nr, nz, tol=1e-6, max_cycles=120)
        np.testing.assert_array_equal(psi[b].cpu().numpy(), p)
        assert (float(res[b]), int(cyc[b]), bool(conv[b])) == (r, n, c)
    assert len({int(c) for c in cyc}) > 1


def test_mg_solve_manufactured_fixed_point_and_validation(mg):
    nz = nr = 33
    r = np.linspace(1.0, 2.0, nr)
    zc = np.linspace(-0.5, 0.5, nz)
    rr, zz = np.meshgrid(r, zc)
    psi = 0.03125 * rr ** 4 - 0.125 * zz ** 2 + 0.05 * rr ** 2 * zz ** 2
    src = G.gs_operator(psi, rr, float(r[1] - r[0]), float(zc[1] - zc[0]))
    out, res, n, conv = mg.multigrid_solve(src, psi, 1.0, 2.0, -0.5, 0.5, nr, nz, tol=1e-12)
    assert conv and n == 0 and res < 1e-12
    np.testing.assert_array_equal(out, psi)
    s = np.zeros((9, 9))
    with pytest.raises(ValueError):
        mg.multigrid_solve(s, np.zeros((9, 8)), 1, 2, 0, 1, 9, 9)
    with pytest.raises(ValueError):
        mg.multigrid_solve(s, s, 1, 2, 0, 1, 9, 9, tol=0.0)
    with pytest.raises(ValueError):
        mg.multigrid_solve(s, s, 1, 2, 0, 1, 9, 9, max_cycles=0)


def test_mg_solve_max_cycles_cap(mg):
    z = golden("mg_solve")
    psi, res, n, conv = mg.multigrid_solve(z["c33_source"], np.zeros((33, 33)), 1.2, 2.2, -0.5, 0.5, 33, 33,
                                           tol=1e-14, max_cycles=2)
    p, r, k, c = G.mg_solve(z["c33_source"], np.zeros((33, 33)), 1.2, 2.2, -0.5, 0.5, 33, 33, tol=1e-14, max_cycles=2)
    np.testing.assert_array_equal(psi, p)
    assert (res, n, conv) == (r, 2, False)


def test_bench_smoother_provider_matches_golden(mg):
    """bench_gpu_gs_solver._problem (seed 2026), the `gs_rb_sor_smooth` tier contract."""
    from scpn_fusion_core_b200 import providers
    z = golden("bench_smooth")
    for n in (65, 129):
        rng = np.random.default_rng(2026)
        rg, zg = np.meshgrid(np.linspace(4.0, 8.0, n), np.linspace(-4.0, 4.0, n))
        source = -np.exp(-((rg - 6.0) ** 2 + zg ** 2) / 0.5)
        psi0 = rng.normal(0.0, 1e-3, size=(n, n))
        psi0[0, :] = psi0[-1, :] = psi0[:, 0] = psi0[:, -1] = 0.0
        keep = psi0.copy()
        out = providers._gpu_gs_rb_sor_smooth(psi0, source, 4.0, 8.0, -4.0, 4.0, omega=1.3, n_sweeps=int(z[f"n{n}_sweeps"]))
        np.testing.assert_array_equal(out, z[f"n{n}_out"])
        np.testing.assert_array_equal(psi0, keep)
        again = providers._gpu_gs_rb_sor_smooth(psi0, source, 4.0, 8.0, -4.0, 4.0, omega=1.3, n_sweeps=int(z[f"n{n}_sweeps"]))
        np.testing.assert_array_equal(out, again)  # bit-identical repeat (reference test :414-439)
    solver = providers.PyGpuSolver(65, 65, 4.0, 8.0, -4.0, 4.0)
    rg, zg = np.meshgrid(np.linspace(4.0, 8.0, 65), np.linspace(-4.0, 4.0, 65))
    src = -np.exp(-((rg - 6.0) ** 2 + zg ** 2) / 0.5)
    flat = solver.solve(np.zeros(65 * 65, np.float32).tolist(), src.astype(np.float32).ravel().tolist(), 50, 1.3)
    ref = G.provider_rb_sor_smooth(np.zeros((65, 65)), src, 4.0, 8.0, -4.0, 4.0, omega=1.3, n_sweeps=50)
    assert flat.dtype == np.float32
    assert np.linalg.norm(flat.reshape(65, 65) - ref) / np.linalg.norm(ref) < 1e-4  # reference GPU-tier bar


def test_large_grid_properties(mg):
    """513^2: linearity of the smoother in (psi, source) and idempotent Dirichlet ring."""
    import torch
    n = 513
    psi, src, rg, dr, dz = _problem(n, n, 5, 4.0, 8.0)
    a = mg.mg_smooth(psi.copy(), src, rg, dr, dz, 1.0, 2)
    b = mg.mg_smooth(2.0 * psi, 2.0 * src, rg, dr, dz, 1.0, 2)
    np.testing.assert_array_equal(b, 2.0 * a)  # scaling by 2 is exact in binary FP
    np.testing.assert_array_equal(a, G.rb_sor_smooth(psi.copy(), src, rg, dr, dz, 1.0, 2))
    for sl in (np.s_[0, :], np.s_[-1, :], np.s_[:, 0], np.s_[:, -1]):
        np.testing.assert_array_equal(a[sl], psi[sl])


@pytest.mark.parametrize("shape,batch", [((129, 129), 3), ((64, 200), 2), ((300, 700), 2), ((1025, 1025), 1),
                                         ((37, 1000), 1), ((2049, 65), 1)])
def test_fused_sweeps_equal_per_colour_passes(mg, shape, batch):
    """The temporally blocked kernel (1..3 sweeps per pass over HBM, single- and multi-tile grids,
    row bands and column strips, in place and ping-pong) is bit-identical to per-colour launches and
    to the oracle."""
    import torch
    nz, nr = shape
    rng = np.random.default_rng(nz * 1000 + nr)
    R = np.linspace(1.0, 3.0, nr)
    rg = np.tile(R, (nz, 1))
    dr, dz = float(R[1] - R[0]), 2.0 / (nz - 1)
    psi = rng.normal(size=(batch, nz, nr))
    src = rng.normal(size=(batch, nz, nr))
    for sweeps in (1, 2, 3, 5):
        ref = mg.mg_smooth(torch.tensor(psi, device="cuda"), torch.tensor(src, device="cuda"), rg, dr, dz, 1.3, sweeps,
                           fuse=0).cpu().numpy()
        for fuse in (1, 2, 3):
            out = mg.mg_smooth(torch.tensor(psi, device="cuda"), torch.tensor(src, device="cuda"), rg, dr, dz, 1.3,
                               sweeps, fuse=fuse).cpu().numpy()
            np.testing.assert_array_equal(out, ref, err_msg=f"sweeps={sweeps} fuse={fuse}")
    if nz * nr <= 300 * 700:
        o = G.rb_sor_smooth(psi[0].copy(), src[0], rg, dr, dz, 1.3, 3)
        np.testing.assert_array_equal(mg.mg_smooth(psi[0].copy(), src[0], rg, dr, dz, 1.3, 3), o)


@pytest.mark.parametrize("shape", [(3, 3), (4, 7), (5, 64), (6, 65), (33, 33), (64, 64), (128, 128), (7, 130)])
def test_fused_sweeps_small_and_even_grids_vs_oracle(mg, shape):
    """Edge shapes of the temporally blocked kernel (fewer rows than the pipeline depth, even sizes, exactly
    one 64-column strip, one column more than a strip) against the oracle, bit-exact."""
    nz, nr = shape
    rng = np.random.default_rng(nz * 131 + nr)
    R = np.linspace(0.5, 2.5, nr)
    rg = np.tile(R, (nz, 1))
    dr, dz = float(R[1] - R[0]), 1.0 / max(nz - 1, 1)
    psi = rng.normal(size=(nz, nr))
    src = rng.normal(size=(nz, nr))
    for sweeps in (1, 2, 3, 4, 7):
        ref = G.rb_sor_smooth(psi.copy(), src, rg, dr, dz, 1.7, sweeps)
        for fuse in (0, 1, 3):
            out = mg.mg_smooth(psi.copy(), src, rg, dr, dz, 1.7, sweeps, fuse=fuse)
            np.testing.assert_array_equal(out, ref, err_msg=f"shape={shape} sweeps={sweeps} fuse={fuse}")


@pytest.mark.parametrize("shape", [(200, 300), (257, 129), (130, 515)])
def test_vcycle_multi_tile_levels_vs_oracle(mg, shape):
    """V-cycles whose finest levels span several 64-column strips (out-of-place fused sweeps with the
    ping-pong partner buffers), even and odd level sizes, batch of 2: bit-identical to the oracle."""
    import torch
    nz, nr = shape
    rng = np.random.default_rng(nz + 7 * nr)
    R = np.linspace(1.5, 4.5, nr)
    rg = np.tile(R, (nz, 1))
    dr, dz = float(R[1] - R[0]), 3.0 / (nz - 1)
    psi = rng.normal(size=(2, nz, nr))
    src = rng.normal(size=(2, nz, nr))
    out = mg.multigrid_vcycle(torch.tensor(psi, device="cuda"), torch.tensor(src, device="cuda"), rg, dr, dz,
                              omega=1.4, pre_smooth=3, post_smooth=3).cpu().numpy()
    for b in range(2):
        ref = G.vcycle(psi[b], src[b], rg, dr, dz, omega=1.4, pre=3, post=3, min_grid=5)
        np.testing.assert_array_equal(out[b], ref, err_msg=f"shape={shape} b={b}")
    # a second cycle from the first one's output, odd smoothing counts (multi-launch ping-pong parity)
    out2 = mg.multigrid_vcycle(torch.tensor(out[0], device="cuda"), torch.tensor(src[0], device="cuda"), rg, dr, dz,
                               omega=1.4, pre_smooth=4, post_smooth=2).cpu().numpy()
    ref2 = G.vcycle(out[0], src[0], rg, dr, dz, omega=1.4, pre=4, post=2, min_grid=5)
    np.testing.assert_array_equal(out2, ref2)
