"""Where does the e2e leg spend its time? (D2H bandwidth, host prep, solve)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import scpn_fusion_core_b200 as pkg
B = 4096
bk = pkg.BatchedFusionKernel(bench.base_config(129), device=0)
cc, ip, ped = bench.uq_inputs(B)
host = torch.empty((B, 129, 129), dtype=torch.float64).pin_memory()
dev = torch.empty((B, 129, 129), dtype=torch.float64, device="cuda")
for _ in range(2):
    host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"D2H pinned 545 MB: {(t1-t0)*1e3:.1f} ms = {dev.numel()*8/(t1-t0)/1e9:.1f} GB/s")
for _ in range(2):
    r = bk.solve(cc, ip, ped, ped, to_host=False); torch.cuda.synchronize()
t0 = time.perf_counter(); r = bk.solve(cc, ip, ped, ped, to_host=False); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"solve(to_host=False): {(t1-t0)*1e3:.1f} ms")
t0 = time.perf_counter(); host.copy_(r["psi"], non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"copy result: {(t1-t0)*1e3:.1f} ms")
