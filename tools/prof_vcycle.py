"""Small driver for ncu: a few gsb_vcycle calls on a batch of 129^2 fields (one CTA per SM)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scpn_fusion_core_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
n = int(sys.argv[2]) if len(sys.argv) > 2 else 129
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(0)
R = np.linspace(2.0, 10.0, n); Z = np.linspace(-4.0, 4.0, n)
rg, _ = np.meshgrid(R, Z)
psi = torch.tensor(rng.normal(size=(B, n, n)), device="cuda")
src = torch.tensor(rng.normal(size=(B, n, n)), device="cuda")
dr, dz = float(R[1] - R[0]), float(Z[1] - Z[0])
out = pkg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.6)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = pkg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.6)
e1.record(); torch.cuda.synchronize()
print(f"B={B} n={n}: {e0.elapsed_time(e1)/reps*1e3:.1f} us per batched V-cycle call (incl. input clone)")

import ctypes
from scpn_fusion_core_b200 import _lib
lib = _lib.load()
buf = (ctypes.c_longlong * 64)()
lib.gsb_debug_phase_cycles(buf, 1)
out = pkg.multigrid_vcycle(psi, src, rg, dr, dz, omega=1.6); torch.cuda.synchronize()
lib.gsb_debug_phase_cycles(buf, 1)
v = list(buf)
if any(v):
    import math
    nb = max(1, math.ceil(B / 148))
    names = ["pre", "res+restrict", "prolong", "post"]
    tot = sum(v)
    print(f"phase cycles per V-cycle (CTA 0, {nb} equilibria), total {tot/nb:.0f} cycles")
    for l in range(8):
        row = v[4*l:4*l+4]
        if any(row):
            print(f"  level {l}: " + "  ".join(f"{n}={c/nb:.0f}" for n, c in zip(names, row)))

if any(v[40:45]) and v[44]:
    print("level nz==9 pass breakdown (cycles/pass): setup=%.0f coefs=%.0f rows=%.0f barrier=%.0f passes=%d  sor_point+store=%.0f" % (v[40]/v[44], v[41]/v[44], v[42]/v[44], v[43]/v[44], v[44], v[45]/v[44]))
