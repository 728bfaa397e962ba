"""Batch sharding over ranks (mode 1) on CPU: two gloo ranks, a stub kernel in place of
the CUDA solve.  Checks the slice arithmetic, that no rank touches another rank's samples, and the
all-gather of the per-equilibrium scalars."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _StubKernel:
    """Deterministic stand-in for BatchedFusionKernel.solve (no GPU): results are functions of the inputs."""

    def __init__(self):
        self.seen = None

    def solve(self, cc, ip, ped_p, ped_ff):
        self.seen = cc.copy()
        return {"iterations": (cc.sum(axis=1) * 10).astype(np.int64), "converged": cc[:, 0] > 0,
                "residual": ip * 1e-5, "psi": np.repeat(cc[:, :1, None], 3, axis=2)}


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from scpn_fusion_core_b200 import shard_range, solve_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        total = 11
        cc = np.arange(total * 3, dtype=np.float64).reshape(total, 3)
        ip = np.linspace(10.0, 20.0, total)
        k = _StubKernel()
        res = solve_sharded(k, cc, ip)
        lo, hi = shard_range(total, world, rank)
        assert res["shard"] == (lo, hi)
        np.testing.assert_array_equal(k.seen, cc[lo:hi])           # only its own samples
        assert res["psi"].shape[0] == hi - lo                       # flux maps stay local
        g = res["global"]
        np.testing.assert_array_equal(g["iterations"], (cc.sum(axis=1) * 10).astype(np.int64))
        np.testing.assert_array_equal(g["residual"], ip * 1e-5)
        np.testing.assert_array_equal(g["converged"], cc[:, 0] > 0)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    from scpn_fusion_core_b200 import shard_range
    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_solve_sharded_two_ranks(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
