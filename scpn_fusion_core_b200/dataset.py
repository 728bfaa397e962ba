"""ITER surrogate dataset producer on the GPU: the reference's ``tools/parallel_gen_iter.py``.

The reference generates training sets with a ``multiprocessing.Pool``: worker ``i`` seeds
``default_rng(42 + i)``, perturbs every coil current by U(0.85, 1.15) and Ip by U(0.8, 1.2) per sample,
runs ``FusionKernel.solve_equilibrium()`` and keeps 12 features + the flattened flux map unless the
X-point sits on the box edge (``parallel_gen_iter.py:72-141``).  Here one chunk is ONE batched device
solve (``BatchedFusionKernel``) followed by one batched topology launch on the final flux maps; the
random draws are made in the reference's order, so chunk ``i`` reproduces worker ``i``'s samples and
the ``--workers`` value only decides how the seeds partition the request (all chunks of a rank share one
batched solve).  Under torchrun the chunks
are dealt round-robin to the ranks (independent units, no data-path collective) and rank 0 writes the
same ``.npz`` (X, Y) and ``.report.json`` the reference writes.

    python -m scpn_fusion_core_b200.dataset --config validation/iter_config.json --samples 4096 --workers 12 \\
        --out data/iter_2d_high_fidelity.npz
"""
from __future__ import annotations

import argparse
import copy
import json
import logging
import os
import time
from pathlib import Path

import numpy as np

from . import _device as D
from . import _lib
from .fusion_kernel import BatchedFusionKernel, _load_config

logger = logging.getLogger(__name__)

N_FEATURES = 12
FIRST_SEED = 42  # parallel_gen_iter.py:173


def is_boundary_xpoint(r_x: float, z_x: float, r_min: float, r_max: float, z_min: float, z_max: float, *,
                       margin_fraction: float = 0.01) -> bool:
    """parallel_gen_iter.py:50-69: True when the X-point lies within 1 % of the box edge."""
    r_margin = max((r_max - r_min) * margin_fraction, 1.0e-12)
    z_margin = max((z_max - z_min) * margin_fraction, 1.0e-12)
    return bool(r_x <= r_min + r_margin or r_x >= r_max - r_margin or z_x <= z_min + z_margin or z_x >= z_max - z_margin)


def draw_perturbations(cfg: dict, n_samples: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """The reference's draw order (:96-101): per sample, one factor per coil, then the Ip factor."""
    base_i = np.array([float(c["current"]) for c in cfg["coils"]], dtype=np.float64)
    base_ip = float(cfg["physics"]["plasma_current_target"])
    rng = np.random.default_rng(seed)
    cc = np.empty((n_samples, base_i.size))
    ip = np.empty(n_samples)
    for s in range(n_samples):
        for c in range(base_i.size):
            cc[s, c] = base_i[c] * rng.uniform(0.85, 1.15)
        ip[s] = base_ip * rng.uniform(0.8, 1.2)
    return cc, ip


MAX_BATCH = 8192  # equilibria per device solve (129^2: ~1.1 GB per field)


def generate_chunks(specs, config_path, allow_boundary_xpoints: bool, *, device: int | None = None,
                    max_batch: int = MAX_BATCH):
    """Several workers' chunks ``[(n_samples, seed), ...]`` in as few batched device solves as possible.

    Returns one ``(X (n_valid, 12), Y (n_valid, nz*nr), rejected_boundary_xpoints, failed_solves)`` per spec.
    A sample whose solve diverges counts as failed when the config sets ``solver.fail_on_diverge`` (the
    reference's solve raises there, :103,138-140); otherwise its best state is kept, as the reference does.
    """
    specs = [(int(n), int(seed)) for n, seed in specs]
    if any(n < 0 for n, _ in specs):
        raise ValueError("n_samples must be >= 0")
    cfg = copy.deepcopy(_load_config(config_path))
    empty = (np.asarray([], dtype=np.float64), np.asarray([], dtype=np.float64), 0, 0)
    total = sum(n for n, _ in specs)
    if total == 0:
        return [empty for _ in specs]
    bk = BatchedFusionKernel(cfg, device=device)
    draws = [draw_perturbations(bk.cfg, n, seed) for n, seed in specs]
    cc = np.concatenate([d[0] for d in draws])
    ip = np.concatenate([d[1] for d in draws])
    fail_on_diverge = bool(bk.cfg["solver"].get("fail_on_diverge", False))
    r_min, r_max = float(np.min(bk.R)), float(np.max(bk.R))
    z_min, z_max = float(np.min(bk.Z)), float(np.max(bk.Z))
    rows: list = []  # per sample: None (failed), False (rejected) or (features, psi)
    for lo in range(0, total, max_batch):
        hi = min(total, lo + max_batch)
        res = bk.solve(cc[lo:hi], ip[lo:hi], to_host=False)
        topo = bk.topology(res["psi"])  # of the FINAL flux maps, like fk._find_magnetic_axis()/find_x_point(fk.Psi)
        psi = res["psi"].cpu().numpy()
        for s in range(hi - lo):
            if fail_on_diverge and int(res["status"][s]) == 3:
                rows.append(None)
                continue
            t = topo[s]
            if t[6] == 0.0:  # no row below 0.5*Z_min: the reference's ((0, 0), min psi) fallback
                rx, zx, psi_x = 0.0, 0.0, float(t[7])
            else:
                rx, zx, psi_x = float(bk.R[int(t[4])]), float(bk.Z[int(t[3])]), float(t[5])
            if not allow_boundary_xpoints and is_boundary_xpoint(rx, zx, r_min, r_max, z_min, z_max):
                rows.append(False)
                continue
            rows.append(([float(ip[lo + s] / 1e6), 5.3, float(bk.R[int(t[1])]), float(bk.Z[int(t[0])]), 1.0, 1.0, float(t[2]),
                          psi_x, 1.7, 0.33, 0.33, 3.0], psi[s].ravel()))
    out, at = [], 0
    for n, _ in specs:
        part = rows[at:at + n]
        at += n
        good = [r for r in part if r]
        out.append((np.asarray([g[0] for g in good], dtype=np.float64), np.asarray([g[1] for g in good], dtype=np.float64),
                    sum(1 for r in part if r is False), sum(1 for r in part if r is None)))
    return out


def generate_chunk(n_samples: int, config_path, seed: int, allow_boundary_xpoints: bool, *, device: int | None = None):
    """One worker's chunk (the reference's ``generate_chunk`` signature; ``config_path``: path or dict)."""
    return generate_chunks([(n_samples, seed)], config_path, allow_boundary_xpoints, device=device)[0]


def chunk_plan(samples: int, workers: int) -> list[tuple[int, int]]:
    """[(n_samples, seed)] per worker (:167-173)."""
    if workers < 1:
        raise ValueError("workers must be positive")
    if samples < 0:
        raise ValueError("samples must be >= 0")
    per, rem = divmod(samples, workers)
    return [(per + (1 if i < rem else 0), FIRST_SEED + i) for i in range(workers)]


def generate_dataset(config_path, samples: int, workers: int, allow_boundary_xpoints: bool = False, *, rank: int = 0,
                     world: int = 1, device: int | None = None, group=None):
    """All chunks; with world > 1 rank r runs chunks r, r+world, ... and rank 0 receives everything
    (``torch.distributed.gather_object`` of the host arrays; the solves themselves need no collective).
    Returns ``(X, Y, rejected, failed)`` on rank 0 and ``None`` elsewhere."""
    plan = chunk_plan(samples, workers)
    idx = [i for i in range(len(plan)) if i % world == rank]
    mine = dict(zip(idx, generate_chunks([plan[i] for i in idx], config_path, allow_boundary_xpoints, device=device)))
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0, group=group)
        if rank != 0:
            return None
        mine = {k: v for part in gathered for k, v in part.items()}
    results = [mine[i] for i in range(len(plan))]
    valid = [r for r in results if len(r[0]) > 0]
    if valid:
        X = np.concatenate([r[0] for r in valid])
        Y = np.concatenate([r[1] for r in valid])
    else:
        X = np.empty((0, N_FEATURES), dtype=np.float64)
        Y = np.empty((0, 0), dtype=np.float64)
    return X, Y, sum(int(r[2]) for r in results), sum(int(r[3]) for r in results)


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--config", required=True)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--workers", type=int, default=12, help="seed partition of the request (reference: pool size)")
    ap.add_argument("--out", default="data/iter_2d_high_fidelity.npz")
    ap.add_argument("--report", help="JSON generation report path; defaults beside --out")
    ap.add_argument("--allow-boundary-xpoints", action="store_true")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    if args.workers < 1:
        ap.error("workers must be positive")
    _lib.require_device()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch = D.torch_mod()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    t0 = time.perf_counter()
    out = generate_dataset(args.config, args.samples, args.workers, args.allow_boundary_xpoints, rank=rank, world=world,
                           device=local)
    elapsed = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    X, Y, rejected, failed = out
    report_path = Path(args.report) if args.report else Path(args.out).with_suffix(".report.json")
    report_path.parent.mkdir(parents=True, exist_ok=True)
    report = {"requested_samples": args.samples, "workers": args.workers, "n_gpus": world,
              "allow_boundary_xpoints": args.allow_boundary_xpoints, "valid_samples": int(len(X)),
              "rejected_boundary_xpoints": rejected, "failed_solves": failed, "elapsed_s": elapsed,
              "status": "passed" if len(X) > 0 else "failed_no_valid_samples"}
    report_path.write_text(json.dumps(report, indent=2, sort_keys=True) + "\n", encoding="utf-8")
    if len(X) == 0:
        raise RuntimeError(f"no valid samples generated; report written to {report_path}")
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    np.savez(args.out, X=X, Y=Y)
    logger.info("Saved %d samples to %s in %.2fs", len(X), args.out, elapsed)


if __name__ == "__main__":
    main()
