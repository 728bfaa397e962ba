"""BASELINE configs[1] shape as a batch: 256 x 257^2 H-mode equilibria on the streaming Picard path (one launch
sequence per iteration, replayed from a CUDA graph).  Prints eq/s and the 350 B/point/iteration roofline fraction.
usage: python tools/bench_batch257.py [n=257] [B=256] [max_iterations]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import scpn_fusion_core_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 257
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = bench.base_config(n)
if len(sys.argv) > 3:
    cfg["solver"]["max_iterations"] = int(sys.argv[3])
bk = pkg.BatchedFusionKernel(cfg)
cc, ip, ped = bench.uq_inputs(B)
run = lambda: bk.solve(cc, ip, ped, ped, to_host=False)
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); r = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
its = float(r["iterations"].sum())
gbs = 350.0 * n * n * its / dt / 1e9
print(f"{n}^2 x {B}: {dt*1e3:.1f} ms = {B/dt:.0f} eq/s, iterations mean {its/B:.1f}, converged {int(r['converged'].sum())}, "
      f"{gbs:.0f} GB/s algorithmic = {gbs/6552:.2f} of the HBM roofline")
