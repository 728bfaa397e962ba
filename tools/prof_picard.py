"""Phase breakdown of the persistent resident Picard kernel (needs libgsb200_prof.so:
`make -C scpn_fusion_core_b200/csrc prof`, then GSB200_LIB=.../libgsb200_prof.so python tools/prof_picard.py [B])."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import scpn_fusion_core_b200 as pkg
from scpn_fusion_core_b200 import _device as D, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = bench.base_config(129)
bk = pkg.BatchedFusionKernel(cfg, device=0)
cc, ip, ped = bench.uq_inputs(B)
ped8 = np.concatenate([ped, ped], axis=1)
w = (1.0 * cc) / (2.0 * np.pi)
w_dev, ip_dev, ped_dev = (D.to_device(a, 0) for a in (w, ip, ped8))
lib = _lib.load()
buf = (ctypes.c_longlong * 64)()
r = bk.solve_device(w_dev, ip_dev, ped_dev)
torch.cuda.synchronize()
lib.gsb_debug_phase_cycles(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    r = bk.solve_device(w_dev, ip_dev, ped_dev)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
s = r["summary"].cpu().numpy()
print(f"B={B}: {ms:.2f} ms per solve, {B/ms*1e3:.0f} eq/s, iters mean {s[:,0].mean():.1f} max {s[:,0].max():.0f}")
lib.gsb_debug_phase_cycles(buf, 1)
v = list(buf)
if v[53]:
    it = v[53]
    names = {48: "topology", 49: "source", 50: "vcycle", 51: "relax+diff", 52: "gsres+decide"}
    tot = sum(v[k] for k in names)
    print(f"CTA 0: {it} Picard iterations, {tot/it:.0f} cycles/iteration")
    for k, nm in names.items():
        print(f"  {nm:14s} {v[k]/it:9.0f} cycles  {100*v[k]/tot:5.1f} %")
    vn = ["pre", "res+restrict", "prolong", "post"]
    for l in range(8):
        row = v[4*l:4*l+4]
        if any(row):
            print(f"  level {l}: " + "  ".join(f"{n}={c/it:.0f}" for n, c in zip(vn, row)))
