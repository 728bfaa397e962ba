"""Deviation of the device solver_method="anderson" run from the oracle, per iteration (calibration of the
tolerances in tests/test_gpu_anderson.py; run on a GPU box)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import gs_oracle as G  # noqa: E402
import scpn_fusion_core_b200 as pkg  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "anderson.npz"), allow_pickle=True)
for tag in ("val33", "iter49"):
    cfg = json.loads(str(z[tag + "_cfg"]))
    k = pkg.FusionKernel(cfg)
    t = time.time()
    r = k.solve_equilibrium()
    dt = time.time() - t
    ro = G.picard_solve(G.PicardProblem(pkg.validate_config(cfg)))
    h, ho, hf = np.array(r["residual_history"]), np.array(ro["residual_history"]), z[tag + "_hist"]
    n = min(len(h), len(ho))
    dev = np.abs(h[:n] - ho[:n]) / np.abs(ho[:n])
    devf = np.abs(h[:n] - hf[:n]) / np.abs(hf[:n])
    print(tag, "iterations", r["iterations"], ro["iterations"], "wall %.3f s" % dt)
    for a in (3, 12, 30, 60, 100, 200, 500, 1000):
        if a <= n:
            print("  hist rel dev vs oracle up to %4d: %.2e   vs reference fixture: %.2e" % (a, dev[:a].max(), devf[:a].max()))
    num = np.linalg.norm(r["psi"] - ro["psi"]) / np.linalg.norm(ro["psi"])
    numf = np.linalg.norm(r["psi"] - z[tag + "_psi"]) / np.linalg.norm(z[tag + "_psi"])
    print("  psi rel L2 vs oracle %.2e   vs reference fixture %.2e   oracle vs fixture %.2e" % (
        num, numf, np.linalg.norm(ro["psi"] - z[tag + "_psi"]) / np.linalg.norm(z[tag + "_psi"])))
