#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 Grad-Shafranov hot path.

Workload (BASELINE.json configs[2]): a UQ sweep of independent ITER-like 129x129 H-mode
equilibria with perturbed coil currents, plasma current and pedestal parameters
(SURVEY.md 8d config 3), `--batch` equilibria PER GPU (weak scaling; 4096 at N=1 is the
named configuration).  One "step" = one complete batched solve: vacuum field from the coil
currents, seed, and Picard iterations until every equilibrium has converged.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = converged equilibria/s with the per-sample inputs
already resident in HBM; `e2e` = same metric through the public host API
(BatchedFusionKernel.solve: pinned host inputs -> device, flux maps + summaries -> host).
`roofline` is for the dominant kernel (k_picard_resident, >95 % of the step), timed live with CUDA
events; `roofline_streaming_smoother` is the HBM-streaming level-0 RB-SOR colour pass; `cpu_baseline` / `--impl reference` time the NumPy port of the reference's CPU path
(oracle/gs_oracle.py - the reference itself is Python and cannot travel to the GPU box) on
all host cores the way the reference runs sweeps (tools/parallel_gen_iter.py: a process pool).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 129
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
# (profiles/), per launch at the headline configuration; None until a capture of the current kernel exists.
TRAFFIC_NCU = 147.845e9  # profiles/r1_v5_summary.md: 24.4 GB read + 123.4 GB written per launch at B = 4096 (ncu --set full)
ITER_COILS = [(3.5, 3.0, -1.0), (8.0, 3.0, 4.0), (9.5, 0.0, 6.0), (8.0, -3.0, 4.0), (3.5, -3.0, -1.0),
              (9.5, 3.0, 3.0), (2.1, 0.0, 0.0)]


def base_config(n: int = GRID) -> dict:
    """validation/iter_config.json of the reference, H-mode profiles, n x n grid."""
    return {
        "reactor_name": "ITER-Like-Demo", "grid_resolution": [n, n],
        "dimensions": {"R_min": 2.0, "R_max": 10.0, "Z_min": -4.0, "Z_max": 4.0},
        "physics": {"plasma_current_target": 15.0, "vacuum_permeability": 1.0, "profiles": {"mode": "h-mode"}},
        "coils": [{"name": f"C{i}", "r": r, "z": z, "current": c} for i, (r, z, c) in enumerate(ITER_COILS)],
        "solver": {"max_iterations": 500, "convergence_threshold": 1e-4, "relaxation_factor": 0.1},
    }


def uq_inputs(batch: int, seed0: int = 2026):
    """Per-sample perturbations (tools/parallel_gen_iter.py:96-101)."""
    base = np.array([c[2] for c in ITER_COILS])
    cc = np.empty((batch, len(base)))
    ip = np.empty(batch)
    ped = np.empty((batch, 4))
    for k in range(batch):
        rng = np.random.default_rng(seed0 + k)
        cc[k] = base * rng.uniform(0.85, 1.15, size=len(base))
        ip[k] = 15.0 * rng.uniform(0.8, 1.2)
        ped[k] = [0.92 * rng.uniform(0.97, 1.03), 0.05 * rng.uniform(0.9, 1.1), 1.0 * rng.uniform(0.9, 1.1),
                  0.3 * rng.uniform(0.9, 1.1)]
    return cc, ip, ped


# ------------------------------------------------------------------------------------ CPU arm
def _cpu_one(args):
    cfg, cc, ip, ped = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gs_oracle as G
    for coil, cur in zip(cfg["coils"], cc):
        coil["current"] = float(cur)
    cfg["physics"]["plasma_current_target"] = float(ip)
    pd = dict(zip(("ped_top", "ped_width", "ped_height", "core_alpha"), (float(v) for v in ped)))
    cfg["physics"]["profiles"] = {"mode": "h-mode", "p_prime": pd, "ff_prime": dict(pd)}
    r = G.picard_solve(G.PicardProblem(cfg))
    return r["iterations"], bool(r["converged"])


def cpu_sweep(n_samples: int, cores: int, seed0: int):
    """Time `n_samples` oracle solves on a `cores`-process pool; returns (eq/s, seconds, iters)."""
    import multiprocessing as mp
    cc, ip, ped = uq_inputs(n_samples, seed0)
    jobs = [(base_config(), cc[k], ip[k], ped[k]) for k in range(n_samples)]
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_one, jobs[:cores])  # warm the workers (imports, page-in)
        t0 = time.perf_counter()
        out = pool.map(_cpu_one, jobs)
        dt = time.perf_counter() - t0
    return n_samples / dt, dt, [o[0] for o in out]


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 4 * max(cores, 2)
    for _ in range(args.warmup):
        pass  # pool warm-up happens inside cpu_sweep; python has no further JIT state to warm
    times, rates = [], []
    for s in range(args.steps):
        rate, dt, _ = cpu_sweep(n, cores, 2026 + 1000 * s)
        times.append(dt)
        rates.append(rate)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "converged_equilibria_per_s", "value": value, "unit": "equilibria/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"UQ sweep of {args.batch} independent ITER-like {GRID}x{GRID} H-mode "
                               "equilibria per GPU (BASELINE configs[2]); coil currents x U(0.85,1.15), "
                               "Ip x U(0.8,1.2), pedestal params +-3-10 %",
                   "grid": [GRID, GRID], "batch_per_gpu": args.batch, "method": "picard+multigrid(3,3,omega=1.6)",
                   "tol": 1e-4, "sample": f"each step solves a bounded sample of {n} of the equilibria on the host cores"},
        "cpu_baseline": {"value": value, "unit": "equilibria/s", "cores": cores, "kind": "port",
                         "sample": f"{n} equilibria per step on a {cores}-process pool (NumPy port of the "
                                   "reference's FusionKernel.solve_equilibrium; the reference is Python and "
                                   "cannot travel to the GPU box)"},
        "e2e": {"value": value, "unit": "equilibria/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region.  Uses NVML in-process (a query costs
    microseconds); spawning nvidia-smi every 200 ms stalls the driver long enough to distort the
    host-driven e2e leg.  Falls back to nvidia-smi only if NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        return [str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) >= 6:
                        self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.05 if self.nvml is not None else 0.5)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy kernel)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_gpu_arm(args) -> None:
    import torch
    import scpn_fusion_core_b200 as pkg
    from scpn_fusion_core_b200 import _device as D
    from scpn_fusion_core_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    B = args.batch
    cfg = base_config(args.grid)
    bk = pkg.BatchedFusionKernel(cfg, device=local)
    cc, ip, ped = uq_inputs(B, 2026 + rank * B)
    ped8 = np.concatenate([ped, ped], axis=1)
    mu0 = cfg["physics"]["vacuum_permeability"]
    w = (mu0 * cc) / (2.0 * np.pi)
    # resident inputs
    w_dev, ip_dev, ped_dev = (D.to_device(a, local) for a in (w, ip, ped8))
    psi_buf = D.empty((B, args.grid, args.grid), local)
    j_buf = D.empty((B, args.grid, args.grid), local)
    # e2e leg: per-sample inputs start in PINNED host memory every step, results end in pinned host memory
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    w_h, ip_h, ped_h = pin(w), pin(ip), pin(ped8)
    w_e2e, ip_e2e, ped_e2e = (torch.empty_like(t, device=f"cuda:{local}") for t in (w_h, ip_h, ped_h))
    psi_host = torch.empty((B, args.grid, args.grid), dtype=torch.float64).pin_memory()
    summ_host = torch.empty((B, 16), dtype=torch.float64).pin_memory()

    def step_resident():
        return bk.solve_device(w_dev, ip_dev, ped_dev, psi_out=psi_buf, jphi_out=j_buf)

    def step_e2e():
        # H2D of this step's inputs (coil-current weights, Ip targets, pedestal parameters) ...
        w_e2e.copy_(w_h, non_blocking=True)
        ip_e2e.copy_(ip_h, non_blocking=True)
        ped_e2e.copy_(ped_h, non_blocking=True)
        r = bk.solve_device(w_e2e, ip_e2e, ped_e2e, psi_out=psi_buf, jphi_out=j_buf)  # the public batched entry point
        # ... and D2H of the step's results: every flux map and the per-equilibrium summary rows
        psi_host.copy_(r["psi"], non_blocking=True)
        summ_host.copy_(r["summary"], non_blocking=True)
        torch.cuda.synchronize()
        return r

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launch_count() - l0, last

    for _ in range(max(args.warmup, 3)):
        last = step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, last = timed(step_resident, args.steps)
    summ = last["summary"].cpu().numpy()
    n_conv = int((summ[:, 1] > 0.5).sum())
    iters = summ[:, 0]
    for _ in range(2):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- dominant kernel: k_picard_resident (one launch = the whole batch), CUDA events on its stream ----
    n_pts = args.grid ** 2
    n_int = (args.grid - 2) ** 2
    peak, peak_src = measured_peak_hbm()
    k_times = []
    for _ in range(max(1, min(args.steps, 3))):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        l0 = _lib.launch_count()
        rr = bk.solve_device(w_dev, ip_dev, ped_dev, psi_out=psi_buf, jphi_out=j_buf, events=ev)
        torch.cuda.synchronize()
        picard_launches = _lib.launch_count() - l0 - 1  # minus gsb_coil_flux
        k_times.append(ev[0].elapsed_time(ev[1]))
    k_ms = float(np.mean(k_times))
    it_sum = float(rr["summary"].cpu().numpy()[:, 0].sum())
    resident = picard_launches == 1
    BYTES_PER_POINT_ITER = 350.0  # 261 (V-cycle, all levels) + 89 (topology, source, relax, residual)
    alg_bytes = BYTES_PER_POINT_ITER * n_pts * it_sum
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # ---- secondary: the HBM-streaming smoother over the whole batch: one launch of the temporally
    # blocked kernel = 3 full RB-SOR sweeps in one pass over HBM (what a V-cycle's pre/post-smoothing
    # uses on grids that do not fit an SM), and the per-colour-pass kernel it replaces ----
    ctx = bk._context(B)
    src = j_buf  # any resident field of the right shape serves as the right-hand side
    st = D.stream_ptr()

    def time_smooth(fuse, sweeps, reps):
        _lib.check(ctx.lib.gsb_smooth_ex(ctx.handle, D.ptr(psi_buf), D.ptr(src), B, 1.6, sweeps, 0, fuse, st))
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(reps):
            _lib.check(ctx.lib.gsb_smooth_ex(ctx.handle, D.ptr(psi_buf), D.ptr(src), B, 1.6, sweeps, 0, fuse, st))
        f1.record()
        torch.cuda.synchronize()
        return f0.elapsed_time(f1) / reps

    s_ms = time_smooth(3, 6, 5) / 2.0        # per launch (3 sweeps each); two launches per call, back to back
    c_ms = time_smooth(0, 3, 10) / 6.0       # per colour-pass launch
    s_bytes = 24.0 * n_int * B * 3           # 24 B/LUP per sweep x 3 sweeps per launch
    s_achieved = s_bytes / (s_ms * 1e-3) / 1e9
    c_bytes = 12.0 * n_int * B
    c_achieved = c_bytes / (c_ms * 1e-3) / 1e9

    if rank == 0:
        total = world * B
        value = total * args.steps / (ms * 1e-3)
        e2e = total * args.steps / (ms_e2e * 1e-3)
        picard_iters = float(iters.mean())
        line = {
            "metric": "converged_equilibria_per_s", "value": value, "unit": "equilibria/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"UQ sweep of {B} independent ITER-like {args.grid}x{args.grid} H-mode "
                                   "equilibria per GPU (BASELINE configs[2]); coil currents x U(0.85,1.15), "
                                   "Ip x U(0.8,1.2), pedestal params +-3-10 %",
                       "grid": [args.grid, args.grid], "batch_per_gpu": B, "global_batch": total,
                       "method": "picard+multigrid(3,3,omega=1.6)", "tol": 1e-4,
                       "l2_policy": f"working set {5 * B * args.grid ** 2 * 8 / 2 ** 20:.0f} MiB per pass exceeds "
                                    "the 126 MB L2" if B * args.grid ** 2 * 8 > 126e6 else
                                    "working set fits L2 (small --batch run; not the headline configuration)",
                       "converged": n_conv, "picard_iterations_mean": picard_iters,
                       "picard_iterations_max": int(iters.max())},
            "e2e": {"value": e2e, "unit": "equilibria/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(w_h.numel() * 8 + ip_h.numel() * 8 + ped_h.numel() * 8),
                    "d2h_bytes_per_step": int(psi_host.numel() * 8 + summ_host.numel() * 8),
                    "api": "BatchedFusionKernel.solve_device with pinned-host -> device input copies and "
                           "device -> pinned-host copies of all flux maps and summaries inside the timed region"},
            "gpu_launches": int(launches),
            "glups_per_vcycle": None,
            "roofline": {"bound": "hbm",
                         "kernel": "k_picard_resident (persistent: one CTA per equilibrium, psi resident in "
                                   "shared memory for the whole solve)" if resident else
                                   "streaming Picard launch sequence (grid does not fit one SM)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": TRAFFIC_NCU, "launch_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_note": f"{BYTES_PER_POINT_ITER:.0f} B per grid point per Picard iteration "
                                                   f"x {n_pts} points x {it_sum:.0f} iterations (sum over the batch); "
                                                   "psi never leaves shared memory, so frac may exceed 1 - the "
                                                   "kernel is FP64-issue bound, not HBM bound",
                         "share_of_step": k_ms / (ms / args.steps)},
            "roofline_streaming_smoother": {"bound": "hbm", "kernel": "k_sweep_warp<6> (3 RB-SOR sweeps = 6 colour passes "
                                            "per pass over HBM, whole batch, warp-autonomous temporal blocking)",
                                            "achieved": s_achieved, "peak": peak, "unit": "GB/s",
                                            "frac": s_achieved / peak, "launch_ms": s_ms,
                                            "algorithmic_bytes_per_launch": s_bytes,
                                            "glups": (n_int * B * 3) / (s_ms * 1e-3) / 1e9,
                                            "per_colour_pass_kernel": {"kernel": "k_smooth_colour", "achieved": c_achieved,
                                                                       "frac": c_achieved / peak, "launch_ms": c_ms,
                                                                       "algorithmic_bytes_per_launch": c_bytes}},
            "clocks": sampler.summary(),
        }
        # GLUPS per V-cycle: 8 LUP per fine point per cycle over the Picard iterations
        lups = 8.0 * n_int * float(iters.sum()) * world
        line["glups_per_vcycle"] = lups * args.steps / (ms * 1e-3) / 1e9
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = 4 * max(cores, 2)
            rate, dt, its = cpu_sweep(n, cores, 2026)
            line["cpu_baseline"] = {"value": rate, "unit": "equilibria/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} of the {B} equilibria, {cores}-process pool, {dt:.1f} s "
                                              f"(NumPy port of the reference CPU path; iterations {min(its)}-{max(its)})"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="equilibria per GPU")
    ap.add_argument("--grid", type=int, default=GRID)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
