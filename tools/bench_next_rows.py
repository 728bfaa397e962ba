"""Timing of the SURVEY.md 8(f) rows built on top of the hot path (one B200):
  * dataset producer (scpn_fusion_core_b200.dataset = reference tools/parallel_gen_iter.py): samples/s for a
    4096-sample 129^2 request incl. the D2H of every flux map and the host-side feature assembly, beside the
    NumPy oracle's chunk generator on one host core (bounded sample);
  * solve_free_boundary(optimize_shape=True): seconds per call at 33^2 / 65^2 / 129^2, beside the oracle at 33^2.
Prints one JSON object.  Usage: python tools/bench_next_rows.py [samples]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import bench
import gs_oracle as G
import scpn_fusion_core_b200 as pkg
from scpn_fusion_core_b200 import dataset as ds

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out = {}

# ---- dataset producer ----------------------------------------------------------------------------
cfg = bench.base_config(129)
cfg["physics"].pop("profiles")          # the reference tool runs the L-mode demo config
ds.generate_dataset(cfg, 256, 2, True)  # warm-up: context, Green's tables, kernels
torch.cuda.synchronize()
t0 = time.perf_counter()
X, Y, rej, failed = ds.generate_dataset(cfg, samples, 12, True)
dt = time.perf_counter() - t0
t0 = time.perf_counter()
Xc, Yc, _, _ = G.dataset_chunk(cfg, 2, 42, True)
dtc = time.perf_counter() - t0
out["dataset"] = {"config": "iter_config 129x129 L-mode, 12 worker seeds, boundary X-points kept", "samples": samples,
                  "valid": int(len(X)), "rejected": rej, "failed": failed, "seconds": dt, "samples_per_s": samples / dt,
                  "y_mbytes": Y.nbytes / 1e6,
                  "cpu_oracle_1core": {"samples": 2, "seconds": dtc, "samples_per_s": 2 / dtc},
                  "first_chunk_matches_oracle_rel_l2": float(np.linalg.norm(Y[:2] - Yc) / np.linalg.norm(Yc)),
                  "features_max_abs_diff": float(np.max(np.abs(X[:2] - Xc)))}

# ---- free boundary with shape optimisation ---------------------------------------------------------
base = {"reactor_name": "ITER-Validated", "grid_resolution": [65, 65],
        "dimensions": {"R_min": 2.0, "R_max": 10.0, "Z_min": -6.0, "Z_max": 6.0},
        "physics": {"plasma_current_target": 15.0, "vacuum_permeability": 1.0},
        "coils": [{"r": 3.9, "z": 7.6, "current": 5.0}, {"r": 8.2, "z": 6.7, "current": -1.0}, {"r": 12.0, "z": 2.7, "current": 0.0},
                  {"r": 12.6, "z": -2.3, "current": 0.0}, {"r": 8.4, "z": -6.7, "current": -1.0},
                  {"r": 4.3, "z": -7.6, "current": 8.0}, {"r": 1.7, "z": 0.0, "current": -5.0}],
        "solver": {"max_iterations": 1000, "convergence_threshold": 1e-4, "relaxation_factor": 0.1, "fail_on_diverge": True}}
th = np.linspace(0.0, 2.0 * np.pi, 9)[:-1] + 0.2
pts = np.column_stack([6.2 + 1.9 * np.cos(th), 3.1 * np.sin(th)])
limits = np.array([2.0e7, 6.0e6, 6.0e6, 6.0e6, 6.0e6, 2.0e7, 2.0e7])
fb = {}
for n in (33, 65, 129):
    cfg = json.loads(json.dumps(base))
    cfg["grid_resolution"] = [n, n]
    times = []
    for rep in range(2):
        k = pkg.FusionKernel(cfg)
        coils = k.build_coilset_from_config()
        coils.currents = coils.currents * 1e6
        coils.target_flux_points, coils.current_limits = pts, limits
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = k.solve_free_boundary(coils, max_outer_iter=3, tol=1e-4, optimize_shape=True, tikhonov_alpha=1e-13)
        times.append(time.perf_counter() - t0)
    fb[f"{n}x{n}"] = {"seconds": min(times), "outer_iterations": r["outer_iterations"], "final_diff": r["final_diff"],
                      "flux_rmse": r["shape_optimization"]["flux_rmse"]}
    if n == 33:
        prob = G.PicardProblem(cfg)
        pos = [(c["r"], c["z"]) for c in cfg["coils"]]
        cur = np.array([c["current"] for c in cfg["coils"]]) * 1e6
        t0 = time.perf_counter()
        ro = G.free_boundary_solve(prob, pos, cur, [1] * 7, max_outer_iter=3, tol=1e-4, optimize_shape=True,
                                   tikhonov_alpha=1e-13, target_points=pts, current_limits=limits)
        fb["33x33"]["cpu_oracle_seconds"] = time.perf_counter() - t0
        fb["33x33"]["psi_rel_l2_vs_oracle"] = float(np.linalg.norm(k.Psi - ro["psi"]) / np.linalg.norm(ro["psi"]))
        fb["33x33"]["currents_max_rel_diff"] = float(np.max(np.abs(r["coil_currents"] - ro["coil_currents"]) / np.abs(ro["coil_currents"])))
out["free_boundary_shape"] = fb
print(json.dumps(out), flush=True)
