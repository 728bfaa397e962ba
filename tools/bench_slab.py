"""Config 5 (BASELINE.json): one 4097x4097 multigrid solve, slab-decomposed over N GPUs.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/bench_slab.py [n] [tol]
N=1 runs the single-GPU multigrid_solve for comparison as well."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import scpn_fusion_core_b200 as pkg
from scpn_fusion_core_b200.slab import CudaSlabOps, SlabComm, SlabMultigrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
R = np.linspace(4.0, 8.0, n); Z = np.linspace(-4.0, 4.0, n)
comm = SlabComm(rank, world)
peer = False
if world > 1 and os.environ.get("SLAB_PEER", "1") != "0":
    peer = comm.enable_peer_halo(local, 12 * n)
mgs = SlabMultigrid(n, n, 4.0, 8.0, -4.0, 4.0, comm, CudaSlabOps(local), min_rows=int(os.environ.get("SLAB_MIN_ROWS", "128")),
                    use_graph=os.environ.get("SLAB_GRAPH", "1") != "0", strict_graph=os.environ.get("SLAB_GRAPH_STRICT", "0") == "1")
g0, g1 = mgs.owned_rows()
rr, zz = np.meshgrid(R, Z[g0:g1])
src = -np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5)   # bench_gpu_gs_solver._problem source, psi_bc = 0
bc = np.zeros_like(src)
srcd, bcd = mgs.ops.from_numpy(src), mgs.ops.from_numpy(bc)
for it in range(2):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    psi, res, cyc, conv = mgs.solve(srcd, bcd, tol=tol, max_cycles=30)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
chk = float(psi.sum().item())
if world > 1:
    t = torch.tensor([chk], dtype=torch.float64, device="cuda"); dist.all_reduce(t); chk = float(t.item())
if rank == 0:
    n_int = (n - 2) ** 2
    line = {"config": f"{n}x{n} multigrid_solve, slab decomposition", "n_gpus": world, "cycles": cyc, "residual": res,
            "converged": conv, "solve_ms": dt * 1e3, "ms_per_vcycle": dt * 1e3 / max(cyc, 1),
            "glups_per_vcycle": 8.0 * n_int * cyc / dt / 1e9, "levels_distributed": len(mgs.levels),
            "cuda_graph": mgs.used_graph, "native_driver": bool(mgs._state.get("native")), "halo_transport": "nvlink-peer-kernels" if peer else "nccl-p2p", "halo_messages": comm.messages, "halo_mbytes": comm.bytes_sent / 1e6, "psi_sum": chk}
    if world == 1:
        rr, zz = np.meshgrid(R, Z)
        s = -np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5)
        sd = torch.tensor(s, device="cuda"); b0 = torch.zeros_like(sd)
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            p, r, c, cv = pkg.multigrid_solve(sd, b0, 4.0, 8.0, -4.0, 4.0, n, n, tol=tol, max_cycles=30)
            torch.cuda.synchronize(); d1 = time.perf_counter() - t0
        line["single_gpu_multigrid_solve"] = {"cycles": c, "residual": r, "solve_ms": d1 * 1e3, "psi_sum": float(p.sum().item()),
                                              "bit_identical": bool(torch.equal(p, psi))}
    print(json.dumps(line), flush=True)
if world > 1:
    comm.disable_peer_halo()
    dist.destroy_process_group()
