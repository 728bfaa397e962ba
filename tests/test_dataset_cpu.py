"""Dataset producer (reference tools/parallel_gen_iter.py) on CPU: the oracle restatement
against the reference-generated fixture, the host-side recipe of the product (draw order, chunk plan,
boundary rejection) and the two-rank chunk distribution over gloo with a stub in place of the CUDA solve."""
from __future__ import annotations

import json
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import gs_oracle as G
from conftest import golden, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_chunk_matches_reference():
    z = golden("dataset")
    for tag in ("val", "iter", "iter_allow"):
        n, seed, allow, rej, failed = (int(v) for v in z[tag + "_meta"])
        X, Y, r, f = G.dataset_chunk(json.loads(str(z[tag + "_cfg"])), n, seed, bool(allow))
        assert (r, f) == (rej, failed), tag
        assert X.shape == z[tag + "_X"].shape and Y.shape == z[tag + "_Y"].shape
        if X.size:
            np.testing.assert_allclose(X, z[tag + "_X"], rtol=1e-12, atol=0)
            assert rel_l2(Y, z[tag + "_Y"]) <= 1e-13
    for args, want in zip([(2.05, 0.0), (2.09, 0.0), (5.0, -5.87), (5.0, -5.89)], z["boundary_checks"]):
        assert G.is_boundary_xpoint(*args, 2.0, 10.0, -6.0, 6.0) == bool(want)


def test_host_recipe_matches_reference_draws():
    from scpn_fusion_core_b200 import dataset as ds
    z = golden("dataset")
    cfg = json.loads(str(z["val_cfg"]))
    cc, ip = ds.draw_perturbations(cfg, 3, 42)
    np.testing.assert_array_equal(ip / 1e6, z["val_X"][:, 0])       # feature 0 is Ip/1e6 of the same draws
    base = np.array([c["current"] for c in cfg["coils"]])
    assert cc.shape == (3, base.size)
    nz = base != 0
    assert np.all(cc[:, nz] / base[nz] >= 0.85) and np.all(cc[:, nz] / base[nz] <= 1.15) and np.all(cc[:, ~nz] == 0.0)
    assert ds.chunk_plan(10, 3) == [(4, 42), (3, 43), (3, 44)]
    assert ds.chunk_plan(2, 4) == [(1, 42), (1, 43), (0, 44), (0, 45)]
    with pytest.raises(ValueError):
        ds.chunk_plan(4, 0)
    for args, want in zip([(2.05, 0.0), (2.09, 0.0), (5.0, -5.87), (5.0, -5.89)], z["boundary_checks"]):
        assert ds.is_boundary_xpoint(*args, 2.0, 10.0, -6.0, 6.0) == bool(want)


def _fake_chunk(n, config, seed, allow, *, device=None):
    """Stub for the CUDA chunk: content is a function of (n, seed) only; odd seeds reject one sample."""
    keep = n - (1 if (seed % 2 and n) else 0)
    X = np.full((keep, 12), float(seed)) if keep else np.asarray([])
    Y = np.full((keep, 5), float(seed) + 0.5) if keep else np.asarray([])
    return X, Y, n - keep, 0


def _fake_chunks(specs, config, allow, *, device=None):
    return [_fake_chunk(n, config, seed, allow) for n, seed in specs]


def _expected(samples, workers):
    from scpn_fusion_core_b200 import dataset as ds
    parts = [_fake_chunk(n, None, s, False) for n, s in ds.chunk_plan(samples, workers)]
    X = np.concatenate([p[0] for p in parts if len(p[0])])
    Y = np.concatenate([p[1] for p in parts if len(p[0])])
    return X, Y, sum(p[2] for p in parts)


def test_generate_dataset_merges_chunks_in_worker_order(monkeypatch):
    from scpn_fusion_core_b200 import dataset as ds
    monkeypatch.setattr(ds, "generate_chunks", _fake_chunks)
    X, Y, rej, failed = ds.generate_dataset("unused", 11, 4)
    eX, eY, erej = _expected(11, 4)
    np.testing.assert_array_equal(X, eX)
    np.testing.assert_array_equal(Y, eY)
    assert (rej, failed) == (erej, 0)
    X, Y, rej, failed = ds.generate_dataset("unused", 0, 3)
    assert X.shape == (0, 12) and Y.shape == (0, 0) and (rej, failed) == (0, 0)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scpn_fusion_core_b200 import dataset as ds
    import test_dataset_cpu as me
    ds.generate_chunks = me._fake_chunks
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        out = ds.generate_dataset("unused", 13, 5, rank=rank, world=world)
        if rank == 0:
            eX, eY, erej = me._expected(13, 5)
            np.testing.assert_array_equal(out[0], eX)
            np.testing.assert_array_equal(out[1], eY)
            assert out[2] == erej
        else:
            assert out is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_generate_dataset_two_ranks(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


class _OracleBatchedKernel:
    """CPU stand-in for BatchedFusionKernel (the device solve and topology launch replaced by the oracle): lets the
    product's chunk assembly - draw order, boundary rejection, feature rows, chunk splitting - run on CPU."""

    def __init__(self, cfg, device=None):
        from scpn_fusion_core_b200.fusion_kernel import validate_config
        self.cfg = validate_config(cfg)
        p = G.PicardProblem(self.cfg)
        self.R, self.Z = p.R, p.Z

    def solve(self, cc, ip, to_host=False):
        import copy
        psi, status = [], []
        for s in range(cc.shape[0]):
            cfg = copy.deepcopy(self.cfg)
            for c, cur in zip(cfg["coils"], cc[s]):
                c["current"] = float(cur)
            cfg["physics"]["plasma_current_target"] = float(ip[s])
            prob = G.PicardProblem(cfg)
            r = G.picard_solve(prob)
            psi.append(prob.Psi.copy())
            status.append(1 if r["converged"] else 2)
        return {"psi": _FakeDeviceArray(np.stack(psi)), "status": np.array(status)}

    def topology(self, psi_dev):
        out = np.zeros((psi_dev.a.shape[0], 8))
        p = G.PicardProblem(self.cfg)
        for s, psi in enumerate(psi_dev.a):
            iz, ir, pax = G.find_axis(psi)
            (rx, zx), px = G.find_x_point(psi, p.R, p.Z, p.dR, p.dZ, self.cfg["dimensions"]["Z_min"])
            found = not (rx == 0.0 and zx == 0.0)
            izx = int(np.argmin(np.abs(p.Z - zx))) if found else 0
            irx = int(np.argmin(np.abs(p.R - rx))) if found else 0
            out[s] = [iz, ir, pax, izx, irx, px, 1.0 if found else 0.0, float(np.min(psi))]
        return out


class _FakeDeviceArray:
    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


def test_chunk_assembly_matches_reference_with_oracle_backed_kernel(monkeypatch):
    from scpn_fusion_core_b200 import dataset as ds
    monkeypatch.setattr(ds, "BatchedFusionKernel", _OracleBatchedKernel)
    z = golden("dataset")
    for tag in ("iter", "iter_allow"):
        n, seed, allow, rej, failed = (int(v) for v in z[tag + "_meta"])
        X, Y, r, f = ds.generate_chunk(n, json.loads(str(z[tag + "_cfg"])), seed, bool(allow))
        assert (r, f) == (rej, failed) and X.shape == z[tag + "_X"].shape
        if X.size:
            np.testing.assert_allclose(X, z[tag + "_X"], rtol=1e-12, atol=0)
            assert rel_l2(Y, z[tag + "_Y"]) <= 1e-13
    # two specs in one call = the two chunks computed separately
    cfg = json.loads(str(z["iter_allow_cfg"]))
    both = ds.generate_chunks([(1, 43), (1, 44)], cfg, True)
    one = ds.generate_chunk(1, cfg, 44, True)
    np.testing.assert_array_equal(both[1][0], one[0])
    np.testing.assert_array_equal(both[1][1], one[1])
    np.testing.assert_array_equal(both[0][0], z["iter_allow_X"][:1])
    with pytest.raises(ValueError):
        ds.generate_chunks([(-1, 42)], cfg, True)
    empty = ds.generate_chunks([(0, 42)], cfg, True)[0]
    assert empty[0].size == 0 and empty[2:] == (0, 0)
