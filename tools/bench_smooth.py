"""Streaming RB-SOR smoother: per-colour launches vs the temporally blocked kernel (CUDA events).
Algorithmic bytes = 24 B per interior point per sweep."""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scpn_fusion_core_b200 import _device as D, _lib

PEAK = 6552.0
cases = [(129, 129, 4096), (257, 257, 1024), (513, 513, 256), (1025, 1025, 64), (4097, 4097, 1), (4097, 4097, 4)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for nz, nr, B in cases:
    R = np.linspace(4.0, 8.0, nr); Z = np.linspace(-4.0, 4.0, nz)
    ctx = D.get_context(nz, nr, R, Z, float(R[1] - R[0]), float(Z[1] - Z[0]), B, 0)
    psi = torch.randn((B, nz, nr), dtype=torch.float64, device="cuda") * 1e-3
    src = torch.randn((B, nz, nr), dtype=torch.float64, device="cuda")
    st = D.stream_ptr()
    n_int = (nz - 2) * (nr - 2)
    only = os.environ.get("FUSE_ONLY")
    for fuse, sweeps in ((0, 6), (1, 6), (2, 6), (3, 6)):
        if only is not None and int(only) != fuse:
            continue
        _lib.check(ctx.lib.gsb_smooth_ex(ctx.handle, D.ptr(psi), D.ptr(src), B, 1.3, sweeps, 0, fuse, st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            _lib.check(ctx.lib.gsb_smooth_ex(ctx.handle, D.ptr(psi), D.ptr(src), B, 1.3, sweeps, 0, fuse, st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 24.0 * n_int * B * sweeps / (ms * 1e-3) / 1e9
        glups = n_int * B * sweeps / (ms * 1e-3) / 1e9
        print(f"{nz}x{nr} B={B} fuse={fuse}: {ms:8.3f} ms for {sweeps} sweeps  {glups:7.1f} GLUPS  "
              f"{gbs:7.0f} GB/s algorithmic = {gbs/PEAK:5.2f} of HBM peak", flush=True)
    del psi, src
    D.clear_cache()
    torch.cuda.empty_cache()
