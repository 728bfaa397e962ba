"""a18: batched plasma->wall Green's-function contraction wall[b][w] = sum_s M[w][s] * (J[b][s]*dA)
(jax_free_boundary_predictive.py:498) as an FP64 tensor-core GEMM; compared with cuBLAS DGEMM on the
same shape (the FP64 tensor peak reference measured on this box)."""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scpn_fusion_core_b200 import _device as D, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
R = np.linspace(2.0, 10.0, n); Z = np.linspace(-4.0, 4.0, n)
ctx = D.get_context(n, n, R, Z, float(R[1] - R[0]), float(Z[1] - Z[0]), B, 0)
nw, ni = 2 * n + 2 * (n - 2), (n - 2) ** 2
M = D.empty((nw, ni), 0)
st = D.stream_ptr()
_lib.check(ctx.lib.gsb_wall_matrix(ctx.handle, 4e-7 * np.pi, D.ptr(M), st))
J = torch.randn((B, n, n), dtype=torch.float64, device="cuda")
wall = D.empty((B, nw), 0)
dA = float((R[1] - R[0]) * (Z[1] - Z[0]))
def run():
    _lib.check(ctx.lib.gsb_wall_flux(ctx.handle, D.ptr(M), D.ptr(J), dA, D.ptr(wall), B, st))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(run)
flops = 2.0 * B * nw * ni
X = (J[:, 1:-1, 1:-1].reshape(B, ni) * dA).contiguous()
ref = X @ M.T
err = float((wall - ref).abs().max() / ref.abs().max())
ms_cublas = timeit(lambda: torch.matmul(X, M.T))
a = torch.randn((8192, 8192), dtype=torch.float64, device="cuda"); b = torch.randn((8192, 8192), dtype=torch.float64, device="cuda")
ms_big = timeit(lambda: torch.matmul(a, b), 3)
print(f"wall GEMM {B} x {nw} x {ni}: k_wall_gemm {ms:.3f} ms = {flops/ms/1e9:.2f} TFLOP/s (gather fused); "
      f"cuBLAS DGEMM same shape {ms_cublas:.3f} ms = {flops/ms_cublas/1e9:.2f} TFLOP/s; rel err vs cuBLAS {err:.1e}")
print(f"cuBLAS DGEMM 8192^3: {ms_big:.2f} ms = {2*8192**3/ms_big/1e9:.2f} TFLOP/s (FP64 tensor peak reference)")
