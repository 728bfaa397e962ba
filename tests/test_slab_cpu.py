"""Slab-decomposed multigrid on CPU ranks: the orchestration of
scpn_fusion_core_b200.slab (partition, halo exchange over gloo, row offsets, coarse gather) with an
oracle-backed compute backend must reproduce the single-process oracle solve bit for bit."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gs_oracle as G

HERE = os.path.dirname(os.path.abspath(__file__))


def _problem(nz, nr):
    rr, zz = np.meshgrid(np.linspace(4.0, 8.0, nr), np.linspace(-4.0, 4.0, nz))
    src = -np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5)
    rng = np.random.default_rng(7)
    bc = rng.normal(scale=1e-3, size=(nz, nr))
    return src, bc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nz, nr, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    from slab_numpy_ops import NumpySlabOps
    from scpn_fusion_core_b200.slab import SlabComm, SlabMultigrid
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        src, bc = _problem(nz, nr)
        comm = SlabComm(rank, world)
        mgs = SlabMultigrid(nz, nr, 4.0, 8.0, -4.0, 4.0, comm, NumpySlabOps(), min_rows=16, gather_nz=0)
        g0, g1 = mgs.owned_rows()
        psi, res, n, conv = mgs.solve(src[g0:g1], bc[g0:g1], tol=1e-9, max_cycles=30)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), psi=psi.numpy(), res=res, n=n, conv=conv, g0=g0, g1=g1,
                 nlev=len(mgs.levels), msgs=comm.messages)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_solve_equals_single_process_oracle(tmp_path, world):
    nz, nr = (129, 65) if world == 2 else (257, 33)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nz, nr, str(tmp_path)), nprocs=world, join=True)
    src, bc = _problem(nz, nr)
    p0, r0, n0, c0 = G.mg_solve(src, bc, 4.0, 8.0, -4.0, 4.0, nr, nz, tol=1e-9, max_cycles=30)
    full = np.empty((nz, nr))
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        full[int(z["g0"]):int(z["g1"])] = z["psi"]
        assert (float(z["res"]), int(z["n"]), bool(z["conv"])) == (r0, n0, c0)
        assert int(z["nlev"]) >= 2 and int(z["msgs"]) > 0
    np.testing.assert_array_equal(full, p0)


def test_slab_world_one_and_plan():
    sys.path.insert(0, HERE)
    from slab_numpy_ops import NumpySlabOps, rb_sor_smooth_offset
    from scpn_fusion_core_b200.slab import SlabComm, SlabMultigrid, plan_slab_levels
    nz, nr = 65, 65
    src, bc = _problem(nz, nr)
    # the offset smoother is the oracle smoother when the offset is even
    rg = np.tile(np.linspace(4.0, 8.0, nr), (nz, 1))
    a = rb_sor_smooth_offset(bc.copy(), src, rg, 0.0625, 0.125, 1.3, 2, 0)
    np.testing.assert_array_equal(a, G.rb_sor_smooth(bc.copy(), src, rg, 0.0625, 0.125, 1.3, 2))
    mgs = SlabMultigrid(nz, nr, 4.0, 8.0, -4.0, 4.0, SlabComm(0, 1), NumpySlabOps(), min_rows=16, gather_nz=0)
    psi, res, n, conv = mgs.solve(src, bc, tol=1e-9, max_cycles=30)
    p0, r0, n0, c0 = G.mg_solve(src, bc, 4.0, 8.0, -4.0, 4.0, nr, nz, tol=1e-9, max_cycles=30)
    np.testing.assert_array_equal(psi.numpy(), p0)
    assert (res, n, conv) == (r0, n0, c0)
    # partition contract: contiguous, aligned, the last rank owns the wall row
    for world in (2, 4, 8):
        lv = [plan_slab_levels(4097, 4097, 4.0, 8.0, -4.0, 4.0, world, r)[0] for r in range(world)]
        for l in range(len(lv[0])):
            assert lv[0][l].g0 == 0 and lv[-1][l].g1 == lv[0][l].nz
            for r in range(1, world):
                assert lv[r][l].g0 == lv[r - 1][l].g1 and lv[r][l].g0 % 2 == 0
    with pytest.raises(ValueError):
        plan_slab_levels(130, 129, 4.0, 8.0, -4.0, 4.0, 2, 0)
