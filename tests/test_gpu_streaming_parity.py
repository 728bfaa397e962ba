"""GPU parity of the code paths BASELINE configs[1]-[3] actually take, end to end against the oracle.

Grids above 129^2 do not fit one SM, so `gsb_picard_solve` runs the STREAMING Picard loop (one launch
sequence per iteration, `gsb_picard.cu`) with the streaming multigrid V-cycle; 129^2 and below run the
persistent shared-memory-resident kernel.  Both are compared here with `gs_oracle.picard_solve` (the
NumPy restatement pinned to the unmodified reference, tests/test_oracle_vs_golden.py) on identical
inputs: relative L2 of psi <= 1e-9, axis / X-point within 1e-6 m, Picard iterations within +-1
(BASELINE.json north_star), histories to 1e-7.

Reference: fusion_kernel_newton_solver.py:499-571 (loop), fusion_kernel.py:295-337 (saddle filter).
"""
from __future__ import annotations

import json
import os

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


class _streaming:
    """Force the streaming Picard loop for grids that would otherwise take the resident kernel."""

    def __enter__(self):
        os.environ["GSB_PICARD_STREAMING"] = "1"

    def __exit__(self, *a):
        os.environ.pop("GSB_PICARD_STREAMING", None)


def _check_against_oracle(pkg, cfg, *, psi_tol=PSI_TOL, expect_streaming=None):
    k = pkg.FusionKernel(cfg)
    l0 = pkg._lib.launch_count()
    r = k.solve_equilibrium()
    launches = pkg._lib.launch_count() - l0
    prob = G.PicardProblem(pkg.validate_config(cfg))
    ro = G.picard_solve(prob)
    assert abs(r["iterations"] - ro["iterations"]) <= 1, (r["iterations"], ro["iterations"])
    assert r["converged"] == ro["converged"]
    assert rel_l2(r["psi"], ro["psi"]) <= psi_tol
    assert rel_l2(k.J_phi, prob.J_phi) <= 1e-8
    n = min(len(r["residual_history"]), len(ro["residual_history"]))
    np.testing.assert_allclose(r["residual_history"][:n], ro["residual_history"][:n], rtol=1e-7)
    np.testing.assert_allclose(r["gs_residual_history"][:n], ro["gs_residual_history"][:n], rtol=1e-7)
    # topology of the final state: same grid indices => positions equal to the last bit
    iz, ir, pax = k._find_magnetic_axis()
    oiz, oir, opax = G.find_axis(ro["psi"])
    assert abs(k.R[ir] - prob.R[oir]) <= POS_TOL and abs(k.Z[iz] - prob.Z[oiz]) <= POS_TOL
    saddle = bool(cfg["solver"].get("xpoint_use_saddle_detection", False))
    (rx, zx), px = k.find_x_point(k.Psi)
    (orx, ozx), opx = G.find_x_point(ro["psi"], prob.R, prob.Z, prob.dR, prob.dZ, prob.cfg["dimensions"]["Z_min"],
                                     saddle=saddle)
    assert abs(rx - orx) <= POS_TOL and abs(zx - ozx) <= POS_TOL
    if expect_streaming is not None:
        # the resident kernel is ONE launch for the whole solve; the streaming loop issues tens per iteration
        assert (launches > 10 * r["iterations"]) == expect_streaming, (launches, r["iterations"])
    return r, ro


def test_streaming_multigrid_picard_257_iter_lmode(pkg):
    """BASELINE configs[1] shape: ITER-like 257^2, Picard + streaming multigrid V-cycle (oracle: 101 iterations)."""
    cfg = golden_cfg(golden("solves"), "iter129")
    cfg["grid_resolution"] = [257, 257]
    r, ro = _check_against_oracle(pkg, cfg, expect_streaming=True)
    assert ro["iterations"] == 101 and ro["converged"]


def test_streaming_multigrid_picard_129_diiid_saddle_forced(pkg):
    """DIII-D 129^2 with the Hessian saddle filter on, forced onto the streaming loop: exercises
    k_topo + k_saddle_cand/k_saddle_pick (or k_xpoint_saddle) + streaming V-cycle."""
    cfg = golden_cfg(golden("solves"), "diiid65s")
    cfg["grid_resolution"] = [129, 129]
    assert cfg["solver"]["xpoint_use_saddle_detection"]
    with _streaming():
        _check_against_oracle(pkg, cfg, expect_streaming=True)
    # and the resident kernel on the same problem gives the same answer
    _check_against_oracle(pkg, cfg, expect_streaming=False)


def test_streaming_multigrid_picard_257_diiid_saddle(pkg):
    cfg = golden_cfg(golden("solves"), "diiid65s")
    cfg["grid_resolution"] = [257, 257]
    _check_against_oracle(pkg, cfg, expect_streaming=True)


def test_streaming_multigrid_picard_513_diiid_saddle_fixed_iterations(pkg):
    """BASELINE configs[3] shape (513^2, DIII-D-like, saddle search) at a fixed iteration count."""
    cfg = golden_cfg(golden("solves"), "diiid65s")
    cfg["grid_resolution"] = [513, 513]
    cfg["solver"]["max_iterations"] = 12
    r, ro = _check_against_oracle(pkg, cfg, expect_streaming=True)
    assert r["iterations"] == 12 and not r["converged"]


def _uq_inputs(cfg, B, seed0=2026):
    """SURVEY.md 8d config 3 recipe (tools/parallel_gen_iter.py:96-101 + pedestal jitter) - bench.py's inputs."""
    base = np.array([c["current"] for c in cfg["coils"]])
    cc, ip, ped = [], [], []
    for ks in range(B):
        rng = np.random.default_rng(seed0 + ks)
        cc.append(base * rng.uniform(0.85, 1.15, size=len(base)))
        ip.append(cfg["physics"]["plasma_current_target"] * rng.uniform(0.8, 1.2))
        ped.append([0.92 * rng.uniform(0.97, 1.03), 0.05 * rng.uniform(0.9, 1.1), 1.0 * rng.uniform(0.9, 1.1),
                    0.3 * rng.uniform(0.9, 1.1)])
    return np.array(cc), np.array(ip), np.array(ped)


def _oracle_sample(cfg, cc, ip, ped):
    c = json.loads(json.dumps(cfg))
    for coil, cur in zip(c["coils"], cc):
        coil["current"] = float(cur)
    c["physics"]["plasma_current_target"] = float(ip)
    pd = dict(zip(("ped_top", "ped_width", "ped_height", "core_alpha"), (float(v) for v in ped)))
    c["physics"]["profiles"] = {"mode": "h-mode", "p_prime": pd, "ff_prime": dict(pd)}
    prob = G.PicardProblem(c)
    return prob, G.picard_solve(prob)


@pytest.mark.parametrize("n,B,streaming", [(129, 4, False), (129, 2, True), (257, 4, True)])
def test_hmode_uq_samples_vs_oracle(pkg, n, B, streaming):
    """The headline workload itself (BASELINE configs[2]: 129^2 H-mode, perturbed coils / Ip / pedestal),
    sample by sample against the oracle - on the resident kernel, forced onto the streaming loop, and at 257^2."""
    cfg = golden_cfg(golden("solves"), "iter129")
    cfg["grid_resolution"] = [n, n]
    cfg["physics"]["profiles"] = {"mode": "h-mode"}
    cc, ip, ped = _uq_inputs(cfg, B)
    bk = pkg.BatchedFusionKernel(cfg)
    if streaming and n <= 129:
        with _streaming():
            res = bk.solve(cc, ip, ped, ped, want_history=True)
    else:
        res = bk.solve(cc, ip, ped, ped, want_history=True)
    for ks in range(B):
        prob, ro = _oracle_sample(cfg, cc[ks], ip[ks], ped[ks])
        assert abs(int(res["iterations"][ks]) - ro["iterations"]) <= 1
        assert bool(res["converged"][ks]) == ro["converged"]
        assert rel_l2(res["psi"][ks], ro["psi"]) <= PSI_TOL
        assert rel_l2(res["j_phi"][ks], prob.J_phi) <= 1e-8
        m = min(int(res["iterations"][ks]), ro["iterations"])
        np.testing.assert_allclose(res["residual_history"][ks][:m], ro["residual_history"][:m], rtol=1e-7)
        oiz, oir, _ = G.find_axis(ro["psi"])
        assert abs(res["axis_R"][ks] - prob.R[oir]) <= POS_TOL and abs(res["axis_Z"][ks] - prob.Z[oiz]) <= POS_TOL


def test_second_device_in_one_process(pkg):
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: a solve on cuda:1 after cuda:0 in the
    same process must launch the >48 KB shared-memory kernels there too."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    cfg = golden_cfg(golden("solves"), "iter65")
    a = pkg.FusionKernel(cfg, device=0).solve_equilibrium()
    with torch.cuda.device(1):
        b = pkg.FusionKernel(cfg, device=1).solve_equilibrium()
    np.testing.assert_array_equal(a["psi"], b["psi"])
    assert a["iterations"] == b["iterations"]
