"""One batched streaming-Picard solve for the ncu launch list: python tools/prof_streaming.py <n> <B> [saddle]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
import scpn_fusion_core_b200 as pkg
n, B = int(sys.argv[1]), int(sys.argv[2])
saddle = len(sys.argv) > 3
if saddle:
    z = np.load(os.path.join(ROOT, "tests", "golden", "solves.npz"))
    cfg = json.loads(str(z["diiid65_cfg"])); cfg["grid_resolution"] = [n, n]
    cfg.setdefault("solver", {})["xpoint_use_saddle_detection"] = True
    bk = pkg.BatchedFusionKernel(cfg)
    base = np.array([c["current"] for c in cfg["coils"]])
    cc = np.stack([base * np.random.default_rng(145419 + i).uniform(0.95, 1.05, size=base.size) for i in range(B)])
    r = bk.solve(cc, to_host=False)
else:
    bk = pkg.BatchedFusionKernel(bench.base_config(n))
    cc, ip, ped = bench.uq_inputs(B)
    r = bk.solve(cc, ip, ped, ped, to_host=False)
torch.cuda.synchronize()
print("iterations", r["iterations"].mean())
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
if saddle:
    r = bk.solve(cc, to_host=False)
else:
    r = bk.solve(cc, ip, ped, ped, to_host=False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"n={n} B={B} saddle={saddle}: {dt*1e3:.1f} ms = {B/dt:.0f} eq/s, iterations {r['iterations'].mean():.1f}, converged {int(r['converged'].sum())}")
