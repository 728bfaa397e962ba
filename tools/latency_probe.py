"""Single-equilibrium latency: resident kernel (one SM runs the whole solve) vs the streaming loop (all SMs, one
launch sequence per iteration replayed from a CUDA graph), for small batches."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import scpn_fusion_core_b200 as pkg
for n in (65, 129):
    cfg = bench.base_config(n); cfg["physics"].pop("profiles")
    for B in (1, 2, 4, 8, 16, 32):
        bk = pkg.BatchedFusionKernel(cfg)
        cc, ip, ped = bench.uq_inputs(B)
        out = []
        for mode in ("resident", "streaming"):
            if mode == "streaming":
                os.environ["GSB_PICARD_STREAMING"] = "1"
            else:
                os.environ.pop("GSB_PICARD_STREAMING", None)
            run = lambda: bk.solve(cc, ip, to_host=False)
            run(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3): r = run()
            torch.cuda.synchronize()
            out.append((time.perf_counter() - t0) / 3 * 1e3)
        os.environ.pop("GSB_PICARD_STREAMING", None)
        print(f"{n}^2 B={B:3d}: resident {out[0]:7.2f} ms   streaming+graph {out[1]:7.2f} ms   iterations {r['iterations'].mean():.0f}", flush=True)
