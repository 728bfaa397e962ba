import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import scpn_fusion_core_b200 as pkg
from scpn_fusion_core_b200 import _device as D, _lib
B = 4096
bk = pkg.BatchedFusionKernel(bench.base_config(129), device=0)
ctx = bk._context(B)
st = D.stream_ptr()
psi = torch.randn((B,129,129), dtype=torch.float64, device="cuda")*1e-3
src = torch.randn((B,129,129), dtype=torch.float64, device="cuda")
pp, ps = D.ptr(psi), D.ptr(src)
def tsm(fuse, sweeps, reps):
    ctx.lib.gsb_smooth_ex(ctx.handle, pp, ps, B, 1.6, sweeps, 0, fuse, st); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps): ctx.lib.gsb_smooth_ex(ctx.handle, pp, ps, B, 1.6, sweeps, 0, fuse, st)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), (t1 - t0) * 1e3
for sweeps, reps in ((6, 10), (3, 20), (6, 10), (3, 20)):
    g, h = tsm(3, sweeps, reps)
    nl = reps * (sweeps // 3)
    print(f"{reps} calls x {sweeps} sweeps = {nl} launches: GPU {g:.3f} ms ({g/nl:.3f} ms/launch), host issue time {h:.3f} ms")
# isolated launches: one call, sync, repeat
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.lib.gsb_smooth_ex(ctx.handle, pp, ps, B, 1.6, 3, 0, 3, st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("isolated single launches (ms):", [round(t, 3) for t in ts])
# two single-launch calls issued back to back, then sync
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.lib.gsb_smooth_ex(ctx.handle, pp, ps, B, 1.6, 3, 0, 3, st); ctx.lib.gsb_smooth_ex(ctx.handle, pp, ps, B, 1.6, 3, 0, 3, st)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("pairs of single-launch calls (ms):", [round(t, 3) for t in ts])
