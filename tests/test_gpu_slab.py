"""GPU: slab-decomposed multigrid through libgsb200's slab entry points.
A slab solve must be bit-identical to the single-GPU multigrid_solve - with one rank (row offsets
zero) and with two ranks sharing cuda:0 (halo rows staged through gloo)."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _problem(nz, nr):
    rr, zz = np.meshgrid(np.linspace(4.0, 8.0, nr), np.linspace(-4.0, 4.0, nz))
    src = -np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5)
    bc = np.random.default_rng(11).normal(scale=1e-3, size=(nz, nr))
    return src, bc


def _worker(rank, world, port, nz, nr, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    from scpn_fusion_core_b200.slab import CudaSlabOps, SlabComm, SlabMultigrid
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        src, bc = _problem(nz, nr)
        comm = SlabComm(rank, world)
        mgs = SlabMultigrid(nz, nr, 4.0, 8.0, -4.0, 4.0, comm, CudaSlabOps(0), min_rows=16)
        g0, g1 = mgs.owned_rows()
        psi, res, n, conv = mgs.solve(src[g0:g1], bc[g0:g1], tol=1e-9, max_cycles=30)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), psi=psi.cpu().numpy(), res=res, n=n, conv=conv, g0=g0, g1=g1,
                 nlev=len(mgs.levels))
    finally:
        dist.destroy_process_group()


def test_slab_one_rank_equals_multigrid_solve():
    import scpn_fusion_core_b200 as pkg
    from scpn_fusion_core_b200.slab import CudaSlabOps, SlabComm, SlabMultigrid
    for nz, nr in ((257, 129), (129, 513)):
        src, bc = _problem(nz, nr)
        mgs = SlabMultigrid(nz, nr, 4.0, 8.0, -4.0, 4.0, SlabComm(0, 1), CudaSlabOps(0), min_rows=16)
        psi, res, n, conv = mgs.solve(src, bc, tol=1e-9, max_cycles=30)
        p0, r0, n0, c0 = pkg.multigrid_solve(src, bc, 4.0, 8.0, -4.0, 4.0, nr, nz, tol=1e-9, max_cycles=30)
        np.testing.assert_array_equal(psi.cpu().numpy(), p0)
        assert (res, n, conv) == (r0, n0, c0)
        assert len(mgs.levels) >= 2


@pytest.mark.parametrize("world", [2, 4])
def test_slab_ranks_equal_single_gpu(tmp_path, world):
    import torch.multiprocessing as mp
    import scpn_fusion_core_b200 as pkg
    nz, nr = 513, 257
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, nz, nr, str(tmp_path)), nprocs=world, join=True)
    src, bc = _problem(nz, nr)
    p0, r0, n0, c0 = pkg.multigrid_solve(src, bc, 4.0, 8.0, -4.0, 4.0, nr, nz, tol=1e-9, max_cycles=30)
    full = np.empty((nz, nr))
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        full[int(z["g0"]):int(z["g1"])] = z["psi"]
        assert (float(z["res"]), int(z["n"]), bool(z["conv"])) == (r0, n0, c0)
        assert int(z["nlev"]) >= 3
    np.testing.assert_array_equal(full, p0)
