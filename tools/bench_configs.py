"""BASELINE.json configs[0], [1], [3] (parity-test cases, measured here for the record).
  c1: one 129^2 ITER-like fixed-boundary solve (latency) - FusionKernel.solve_equilibrium
  c2: 257^2 ITER-like: one solve (latency), solve_free_boundary, and a batch of 256 (streaming Picard path)
  c4: 256 x 513^2 DIII-D-shaped with X-point saddle detection, coil currents x U(0.95,1.05), seeds 145419+k"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
import scpn_fusion_core_b200 as pkg

def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r

cfg = bench.base_config(129); cfg["physics"].pop("profiles")
k = pkg.FusionKernel(cfg)
dt, r = t(lambda: k.solve_equilibrium())
print(f"c1 129^2 L-mode single solve: {dt*1e3:.1f} ms, iterations {r['iterations']}, converged {r['converged']}")
cfg = bench.base_config(257); cfg["physics"].pop("profiles")
k = pkg.FusionKernel(cfg)
dt, r = t(lambda: k.solve_equilibrium(), 1)
print(f"c2 257^2 single solve: {dt*1e3:.1f} ms, iterations {r['iterations']}, converged {r['converged']}")
coils = k.build_coilset_from_config()
dt, r = t(lambda: k.solve_free_boundary(coils, max_outer_iter=20, tol=1e-4), 1)
print(f"c2 257^2 solve_free_boundary: {dt*1e3:.1f} ms, outer iterations {r.get('outer_iterations')}, final_diff {r.get('final_diff')}")
B = 256
bk = pkg.BatchedFusionKernel(bench.base_config(257))
cc, ip, ped = bench.uq_inputs(B)
dt, r = t(lambda: bk.solve(cc, ip, ped, ped, to_host=False), 1)
print(f"c2 257^2 x {B} H-mode batch (streaming Picard): {dt*1e3:.1f} ms = {B/dt:.0f} eq/s, iterations {r['iterations'].mean():.1f}, converged {int(r['converged'].sum())}")
z = np.load(os.path.join(ROOT, "tests", "golden", "solves.npz"))
dcfg = json.loads(str(z["diiid65_cfg"]))
dcfg["grid_resolution"] = [513, 513]
dcfg.setdefault("solver", {})["xpoint_use_saddle_detection"] = True
bk = pkg.BatchedFusionKernel(dcfg)
base = np.array([c["current"] for c in dcfg["coils"]])
cc = np.stack([base * np.random.default_rng(145419 + i).uniform(0.95, 1.05, size=base.size) for i in range(B)])
dt, r = t(lambda: bk.solve(cc, to_host=False), 1)
print(f"c4 513^2 x {B} DIII-D-shaped, saddle detection: {dt*1e3:.1f} ms = {B/dt:.1f} eq/s, iterations {r['iterations'].mean():.1f} (max {r['iterations'].max()}), converged {int(r['converged'].sum())}")
