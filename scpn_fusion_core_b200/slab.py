"""Slab-decomposed multigrid_solve: ONE large (R,Z) grid split into Z-row slabs, one process per GPU.

The reference solves a single grid on one device (``multigrid_solve``, ``multigrid_solve.py:352-463``);
its only decomposition code for this path is the additive-Schwarz scaffold of
``scpn-fusion-rs/crates/fusion-core/src/mpi_domain.rs:48-965`` (threads + serial copies, not the same
iteration).  This module keeps the reference's iteration exactly - a slab solve is bit-identical to
the single-GPU solve - and exchanges halo rows between neighbouring ranks:

* rank r owns global rows [r*R, (r+1)*R) of every distributed level (the last rank also owns the
  wall row); local arrays carry ``halo`` rows on each inner side;
* smoothing = temporally blocked RB-SOR sweeps on the local array (``gsb_slab_smooth``): one halo
  exchange of 2*sweeps rows per smoothing phase instead of one row per colour pass;
* residual + full weighting and prolongation + add act on owned rows with one-/two-row halos;
* once a level has fewer than ``min_rows`` rows per rank its right-hand side is all-gathered and
  the rest of the V-cycle runs replicated on every rank (no scatter needed afterwards);
* the convergence norm is an all-reduce(MAX) of one double per cycle.

Communication goes through ``torch.distributed`` point-to-point ops (NCCL over NVLink on the GPU
box; with the gloo backend halo rows are staged through host memory, which is how the two-rank
tests run).  The compute backend is injected (``ops``): ``CudaSlabOps`` drives libgsb200; the CPU
tests inject an oracle-backed implementation to check the orchestration without a GPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _lib


class _LevelDesc(ctypes.Structure):  # gsb_slab_level_desc (include/gsb200.h)
    _fields_ = [("ctx", ctypes.c_void_p), ("x", ctypes.c_void_p), ("f", ctypes.c_void_p), ("alt", ctypes.c_void_p),
                ("cur", ctypes.c_void_p)] + [(n, ctypes.c_int) for n in (
                    "rows_loc", "nr", "own0", "own1", "has_up", "has_dn", "row0", "roff", "ci0", "ci1", "nzc_loc", "nrc",
                    "fi0", "fi1")]


class _HaloDesc(ctypes.Structure):  # gsb_slab_halo_desc
    _fields_ = [(n, ctypes.c_void_p) for n in ("inbox_up", "inbox_dn", "up_inbox_dn", "dn_inbox_up", "flags_local",
                                                "flags_up", "flags_dn", "counters", "epochs")] + [("cap", ctypes.c_longlong)]


@dataclass
class SlabLevel:
    """One distributed level as seen by one rank."""
    nz: int        # global rows
    nr: int        # columns
    g0: int        # first owned global row
    g1: int        # one past the last owned global row
    h_top: int     # halo rows above / below the owned rows
    h_bot: int
    dr: float
    dz: float
    r_row: np.ndarray

    @property
    def rows_loc(self) -> int:
        return self.h_top + (self.g1 - self.g0) + self.h_bot

    @property
    def row0(self) -> int:
        """global row index of local row 0"""
        return self.g0 - self.h_top

    def loc(self, g: int) -> int:
        return g - self.row0


def plan_slab_levels(nz: int, nr: int, r_min: float, r_max: float, z_min: float, z_max: float, world: int,
                     rank: int, *, halo: int = 12, min_rows: int = 32, min_grid: int = 5,
                     gather_nz: int = 129) -> tuple[list[SlabLevel], dict]:
    """Row partition of every distributed level plus the geometry of the first gathered level.

    Needs nz = 2^p + 1 rows with (nz-1) divisible by world; a level stays distributed while it has more than
    ``gather_nz`` rows in total, at least ``min_rows`` (and an even number of) rows per rank, and the
    reference's recursion continues.  ``gather_nz`` = 129: a 129-row level and everything below it fits the
    shared memory of one SM, so the replicated tail of the V-cycle is ONE resident-kernel launch
    (k_vcycle_resident) instead of several latency-bound streaming levels.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if (nz - 1) % world or (nz - 1) & (nz - 2) or (nr - 1) & (nr - 2):
        raise ValueError("slab mode needs a (2^p+1) x (2^q+1) grid and (nz-1) divisible by the world size")
    lib = _lib.load()
    dp = ctypes.POINTER(ctypes.c_double)
    r_axis = np.linspace(r_min, r_max, nr)
    z_axis = np.linspace(z_min, z_max, nz)
    dr, dz = float(r_axis[1] - r_axis[0]), float(z_axis[1] - z_axis[0])  # multigrid_solve.py:431-432

    def tables(level: int):
        n_r = (nr - 1) // (1 << level) + 1
        r = np.zeros(n_r)
        sc = np.zeros(4)
        n = lib.gsb_plan_level_tables(nz, nr, r_axis.ctypes.data_as(dp), dr, dz, min_grid, level,
                                      r.ctypes.data_as(dp), None, None, sc.ctypes.data_as(dp))
        if n != n_r:
            raise _lib.GsbError("gsb_plan_level_tables: unexpected level width")
        r[0], r[-1] = r_axis[0], r_axis[-1]  # wall columns are injected (multigrid_solve.py:93-98)
        return r, float(sc[0]), float(sc[1])

    levels: list[SlabLevel] = []
    lev = 0
    nz_l, nr_l = nz, nr
    while True:
        rows = (nz_l - 1) // world
        stop = min_grid >= nz_l or min_grid >= nr_l  # the reference's base case (multigrid_solve.py:292)
        if stop or rows < max(min_rows, 2 * halo) or rows % 2 or (levels and nz_l <= gather_nz):
            break
        r_row, dr_l, dz_l = tables(lev)
        g0 = rank * rows
        g1 = (rank + 1) * rows + (1 if rank == world - 1 else 0)
        levels.append(SlabLevel(nz_l, nr_l, g0, g1, 0 if rank == 0 else halo, 0 if rank == world - 1 else halo,
                                dr_l, dz_l, r_row))
        nz_l, nr_l = (nz_l + 1) // 2, (nr_l + 1) // 2
        lev += 1
    r_row, dr_l, dz_l = tables(lev)
    gathered = {"level": lev, "nz": nz_l, "nr": nr_l, "r_row": r_row, "dr": dr_l, "dz": dz_l,
                "rows_per_rank": (nz_l - 1) // world}
    return levels, gathered


class CudaSlabOps:
    """Compute backend on one GPU: libgsb200 slab entry points + torch CUDA tensors as buffers."""

    def __init__(self, device: int):
        from . import _device as D
        self.D = D
        self.device = device
        self.torch = D.torch_mod()
        self._ctx: dict = {}
        self._alt: dict = {}
        self._linf_out = None

    def zeros(self, shape):
        return self.D.zeros(shape, self.device)

    def from_numpy(self, a):
        return self.D.to_device(a, self.device)

    def to_numpy(self, t):
        return t.cpu().numpy()

    def _context(self, L: SlabLevel):
        key = (L.nz, L.nr, L.rows_loc, L.dr, L.dz, np.ascontiguousarray(L.r_row, dtype=np.float64).tobytes())
        if key not in self._ctx:
            self._ctx[key] = self.D.Context(L.rows_loc, L.nr, L.r_row, None, L.dr, L.dz, 1, self.device)
        return self._ctx[key]

    def smooth(self, L: SlabLevel, x, f, omega: float, sweeps: int, out=None):
        """``sweeps`` RB-SOR sweeps of the local array; returns the tensor holding the result
        (``out`` if given and usable for the last pass, so a V-cycle can end in the buffer it started in)."""
        D = self.D
        ctx = self._context(L)
        st = D.stream_ptr()
        cur = x
        left = sweeps
        while left > 0:
            s = min(left, 3)
            if ctx.lib.gsb_slab_single_tile(ctx.handle, s):
                dst = cur
            else:  # out of place: write into a cached buffer of this shape that is not the input
                pool = self._alt.setdefault((L.rows_loc, L.nr), [])
                last = left <= 3
                if last and out is not None and out.data_ptr() != cur.data_ptr():
                    dst = out
                else:
                    dst = next((b for b in pool if b.data_ptr() != cur.data_ptr()
                                and (out is None or b.data_ptr() != out.data_ptr())), None)
                    if dst is None:
                        dst = D.empty(tuple(x.shape), self.device)
                        pool.append(dst)
            _lib.check(ctx.lib.gsb_slab_smooth(ctx.handle, D.ptr(cur), D.ptr(dst), D.ptr(f), omega, s, L.row0, st),
                       "gsb_slab_smooth")
            cur = dst
            left -= s
        return cur  # may be a pooled buffer: callers rebind (the input tensor is then scratch)

    def residual_restrict(self, L: SlabLevel, C: SlabLevel | None, x, f, d, roff: int, ci0: int, ci1: int):
        D = self.D
        ctx = self._context(L)
        _lib.check(ctx.lib.gsb_slab_residual_restrict(ctx.handle, D.ptr(x), D.ptr(f), D.ptr(d), int(d.shape[0]),
                                                      int(d.shape[1]), roff, ci0, ci1, D.stream_ptr()),
                   "gsb_slab_residual_restrict")

    def prolong_add(self, L: SlabLevel, x, e, roff: int, fi0: int, fi1: int):
        D = self.D
        ctx = self._context(L)
        _lib.check(ctx.lib.gsb_slab_prolong_add(ctx.handle, D.ptr(e), int(e.shape[0]), int(e.shape[1]), D.ptr(x), roff,
                                                fi0, fi1, D.stream_ptr()), "gsb_slab_prolong_add")

    def residual_linf(self, L: SlabLevel, x, f, row0: int, row1: int) -> float:
        D = self.D
        ctx = self._context(L)
        out = self._linf_out
        if out is None:
            out = self._linf_out = D.zeros((1,), self.device)
        out.zero_()
        _lib.check(ctx.lib.gsb_slab_residual_linf(ctx.handle, D.ptr(x), D.ptr(f), row0, row1, D.ptr(out),
                                                  D.stream_ptr()), "gsb_slab_residual_linf")
        return float(out.item())

    def residual_linf_async(self, L: SlabLevel, x, f, row0: int, row1: int, out) -> None:
        """Stream-ordered form: max |L x - f| of the local rows into the device scalar `out` (no host read)."""
        D = self.D
        ctx = self._context(L)
        out.zero_()
        _lib.check(ctx.lib.gsb_slab_residual_linf(ctx.handle, D.ptr(x), D.ptr(f), row0, row1, D.ptr(out),
                                                  D.stream_ptr()), "gsb_slab_residual_linf")

    def coarse_vcycle(self, G: dict, d_full, omega: float, pre: int, post: int, min_grid: int):
        """The replicated tail of the V-cycle on the gathered level (zero initial guess): one gsb_vcycle
        call on a cached context of that level's geometry (multigrid_solve.py:252-335 from level `G`)."""
        D = self.D
        key = ("coarse", G["nz"], G["nr"], G["dr"], G["dz"], np.ascontiguousarray(G["r_row"], dtype=np.float64).tobytes())
        ctx = self._ctx.get(key)
        if ctx is None:
            ctx = self._ctx[key] = D.Context(G["nz"], G["nr"], G["r_row"], None, G["dr"], G["dz"], 1, self.device)
        x0 = self.torch.zeros_like(d_full)
        _lib.check(ctx.lib.gsb_vcycle(ctx.handle, D.ptr(x0), D.ptr(d_full), 1, omega, int(pre), int(post), int(min_grid),
                                      D.stream_ptr()), "gsb_vcycle")
        return x0


class SlabComm:
    """Halo exchange / gather / max-reduce over a torch.distributed process group (NCCL or gloo)."""

    def __init__(self, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.group = rank, world, group
        self.backend = dist.get_backend(group) if world > 1 else "none"
        self.bytes_sent = 0
        self.messages = 0
        self.peer = None
        self.gather = self.norm = None
        # bumped whenever the peer-memory transport is (re)created or torn down: a SlabMultigrid that cached raw
        # inbox / flag pointers (native descriptors, a captured graph) for an older generation must rebuild them
        self.peer_generation = 0

    def enable_peer_halo(self, device: int, max_doubles: int) -> bool:
        """Exchange halos over NVLink peer memory with libgsb200's own push/recv kernels instead of NCCL
        point-to-point launches (one process per GPU, buffers shared through CUDA IPC).  Collective:
        every rank of the group must call it.  Returns False (and keeps NCCL) if IPC is unavailable."""
        if getattr(self, "peer", None) is not None:
            self.disable_peer_halo()
        self.peer = None
        if self.world == 1:
            return False
        import ctypes as C
        torch, dist = self.torch, self.dist
        lib = _lib.load()
        dev = torch.device(f"cuda:{device}")
        # one IPC block per rank: [4 int64 flags | pad to 256 B | inbox_up | inbox_dn]
        inbox_bytes = ((max_doubles * 8 + 255) // 256) * 256
        total = 256 + 2 * inbox_bytes
        base = C.c_void_p()
        handle = C.create_string_buffer(64)
        ok = 1 if lib.gsb_ipc_alloc(device, total, C.byref(base), handle) == 0 else 0
        gathered: list = [None] * self.world
        dist.all_gather_object(gathered, (ok, bytes(handle.raw), device), group=self.group)
        if not all(g[0] for g in gathered):
            if ok:
                lib.gsb_ipc_free(base)
            return False
        peers = {}
        for name, r in (("up", self.rank - 1), ("dn", self.rank + 1)):
            if 0 <= r < self.world:
                p = C.c_void_p()
                if lib.gsb_ipc_open(device, gathered[r][1], C.byref(p)) != 0:
                    ok = 0
                    break
                peers[name] = {"flags": p.value, "inbox_up": p.value + 256, "inbox_dn": p.value + 256 + inbox_bytes}
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            # some rank could not map a neighbour: unmap what this rank opened and release its own block
            for nb in peers.values():
                lib.gsb_ipc_close(C.c_void_p(nb["flags"]))
            dist.barrier(group=self.group)  # nobody still maps this rank's block when it is freed
            lib.gsb_ipc_free(base)
            return False
        own = {"flags": base.value, "inbox_up": base.value + 256, "inbox_dn": base.value + 256 + inbox_bytes}
        self.peer = {"own": own, "peers": peers, "counters": torch.zeros(4, dtype=torch.int32, device=dev),
                     "epochs": torch.zeros(4, dtype=torch.int64, device=dev), "cap": max_doubles, "lib": lib}
        self.peer_generation += 1
        return True

    def _make_allgather(self, device: int, doubles: int):
        """One all-to-all block {world int64 flags | two halves of `doubles`} per rank, mapped by every rank through
        CUDA IPC (collective).  Returns None if IPC is unavailable on some rank."""
        import ctypes as C
        torch, dist = self.torch, self.dist
        lib = _lib.load()
        dev = torch.device(f"cuda:{device}")
        half = ((doubles * 8 + 255) // 256) * 256
        total = 256 + 2 * half
        base = C.c_void_p()
        handle = C.create_string_buffer(64)
        ok = 1 if lib.gsb_ipc_alloc(device, total, C.byref(base), handle) == 0 else 0
        gathered: list = [None] * self.world
        dist.all_gather_object(gathered, (ok, bytes(handle.raw)), group=self.group)
        if not all(g[0] for g in gathered):
            if ok:
                lib.gsb_ipc_free(base)
            return None
        ptrs: list = [None] * self.world
        opened = []
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = base.value
                continue
            p = C.c_void_p()
            if lib.gsb_ipc_open(device, gathered[r][1], C.byref(p)) != 0:
                ok = 0
                break
            ptrs[r] = p.value
            opened.append(p.value)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            for q in opened:
                lib.gsb_ipc_close(C.c_void_p(q))
            dist.barrier(group=self.group)
            lib.gsb_ipc_free(base)
            return None
        bufs = (C.c_void_p * self.world)(*[C.c_void_p(q + 256) for q in ptrs])
        flags = (C.c_void_p * self.world)(*[C.c_void_p(q) for q in ptrs])
        return {"base": base.value, "opened": opened, "bufs": bufs, "flags": flags, "half_doubles": half // 8,
                "cap": doubles, "counters": torch.zeros(self.world, dtype=torch.int32, device=dev),
                "epoch": torch.zeros(1, dtype=torch.int64, device=dev), "lib": lib}

    def _free_allgather(self, G) -> None:
        import ctypes as C
        for q in G["opened"]:
            G["lib"].gsb_ipc_close(C.c_void_p(q))
        G["lib"].gsb_ipc_free(C.c_void_p(G["base"]))

    def enable_peer_gather(self, device: int, level_doubles: int) -> bool:
        """Gather the coarsest distributed right-hand side AND the per-cycle convergence norm over NVLink peer memory
        (gsb_gather_push / gsb_gather_wait) instead of NCCL collectives: every rank maps every rank's blocks through
        CUDA IPC.  Collective.  With it (and peer halos) a V-cycle consists of libgsb200 kernels only and is replayed
        from a CUDA graph on every rank; the norm lands in pinned host memory and the host polls it (no stream
        synchronisation, no NCCL launch in the solve loop).  Returns False (NCCL keeps doing both) if IPC is
        unavailable or the group has more than 16 ranks."""
        if getattr(self, "gather", None) is not None:
            self.disable_peer_gather()
        self.gather = self.norm = None
        if self.world == 1 or self.world > 16:
            return False
        g = self._make_allgather(device, level_doubles)
        if g is None:
            return False
        m = self._make_allgather(device, self.world)
        if m is None:
            self._free_allgather(g)
            return False
        torch = self.torch
        m["host"] = torch.zeros(self.world, dtype=torch.float64).pin_memory()      # one norm per rank
        m["host_flag"] = torch.zeros(1, dtype=torch.int64).pin_memory()            # exchange number of `host`
        m["local"] = torch.zeros(1, dtype=torch.float64, device=f"cuda:{device}")  # this rank's max |r| (bit pattern)
        m["issued"] = 0
        self.gather, self.norm = g, m
        self.peer_generation += 1
        return True

    def disable_peer_gather(self) -> None:
        if getattr(self, "gather", None) is None:
            return
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        self._free_allgather(self.gather)
        self._free_allgather(self.norm)
        self.gather = self.norm = None
        self.peer_generation += 1

    def _allgather_rows(self, G, owned, n: int, off: int, n_total: int, out, host_flag=None) -> None:
        import ctypes as C
        st = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        vp = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(G["lib"].gsb_gather_push(vp(owned), n, off, G["half_doubles"], G["bufs"], G["flags"], self.world, self.rank,
                                            vp(G["counters"]), vp(G["epoch"]), st), "gsb_gather_push")
        _lib.check(G["lib"].gsb_gather_wait(G["bufs"][self.rank], G["half_doubles"], G["flags"][self.rank], self.world, n_total,
                                            vp(out), vp(G["epoch"]), None if host_flag is None else vp(host_flag), st),
                   "gsb_gather_wait")

    def gather_rows_peer(self, owned, rows_per_rank: int, nz: int, out) -> None:
        """Peer-memory form of gather_rows: this rank's owned rows go straight into every rank's copy of the level;
        `out` (nz, nr) receives the assembled level.  Stream-ordered, no host synchronisation."""
        G = self.gather
        nr = int(owned.shape[1])
        n_rows = rows_per_rank + (1 if self.rank == self.world - 1 else 0)
        if nz * nr > G["cap"]:
            raise _lib.GsbError("peer gather buffer too small for this level")
        self._allgather_rows(G, owned, n_rows * nr, self.rank * rows_per_rank * nr, nz * nr, out)
        self.bytes_sent += (self.world - 1) * n_rows * nr * 8
        self.messages += self.world - 1

    def norm_exchange_issue(self) -> None:
        """Stream-ordered: publish this rank's max |r| (``self.norm["local"]``, filled by the residual kernel) to every
        rank; the assembled per-rank values land in pinned host memory together with their exchange number."""
        M = self.norm
        self._allgather_rows(M, M["local"], 1, self.rank, self.world, M["host"], host_flag=M["host_flag"])

    def norm_exchange_wait(self, expected: int) -> float:
        """Host side: poll the pinned exchange number (no stream synchronisation), then max over the ranks (NaN wins)."""
        M = self.norm
        flag, host = M["host_flag"], M["host"]
        import time as _t
        t0 = _t.perf_counter()
        while int(flag[0]) < expected:
            if _t.perf_counter() - t0 > 120.0:
                raise _lib.GsbError("slab norm exchange timed out (a peer rank stopped issuing exchanges)")
        vals = host.numpy().copy()
        return float("nan") if np.isnan(vals).any() else float(vals.max())

    def disable_peer_halo(self) -> None:
        """Unmap the neighbours' IPC blocks and free this rank's (collective in spirit: call it on every rank
        once no exchange is in flight; NCCL point-to-point exchanges take over again).  Solvers that cached
        pointers into these blocks (native level descriptors, captured graphs) see the new
        ``peer_generation`` at their next ``solve()`` and rebuild their state."""
        P = getattr(self, "peer", None)
        if P is None:
            return
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)  # nobody may still be pushing into a block that is unmapped below
        import ctypes as C
        for nb in P["peers"].values():
            P["lib"].gsb_ipc_close(C.c_void_p(nb["flags"]))
        P["lib"].gsb_ipc_free(C.c_void_p(P["own"]["flags"]))
        self.peer = None
        self.peer_generation += 1

    def _exchange_peer(self, x, L: SlabLevel, k: int) -> None:
        import ctypes as C
        P = self.peer
        n = k * L.nr
        if n > P["cap"]:
            raise _lib.GsbError("peer halo inbox too small for this exchange")
        own0, own1 = L.h_top, L.h_top + (L.g1 - L.g0)
        vp = lambda t: C.c_void_p(t if isinstance(t, int) else t.data_ptr())
        null = C.c_void_p()
        up, dn = P["peers"].get("up"), P["peers"].get("dn")
        has_up, has_dn = bool(L.h_top) and up is not None, bool(L.h_bot) and dn is not None
        st = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        lib, own = P["lib"], P["own"]
        _lib.check(lib.gsb_halo_push(vp(x[own0:own0 + k]) if has_up else null, vp(x[own1 - k:own1]) if has_dn else null, n,
                                     vp(up["inbox_dn"]) if has_up else null, vp(dn["inbox_up"]) if has_dn else null,
                                     vp(own["flags"]), vp(up["flags"]) if has_up else null,
                                     vp(dn["flags"]) if has_dn else null, vp(P["counters"]), vp(P["epochs"]), st), "gsb_halo_push")
        _lib.check(lib.gsb_halo_recv(vp(x[own0 - k:own0]) if has_up else null, vp(x[own1:own1 + k]) if has_dn else null, n,
                                     vp(own["inbox_up"]) if has_up else null, vp(own["inbox_dn"]) if has_dn else null,
                                     vp(own["flags"]), vp(up["flags"]) if has_up else null,
                                     vp(dn["flags"]) if has_dn else null, vp(P["counters"]), vp(P["epochs"]), st), "gsb_halo_recv")
        self.bytes_sent += (int(has_up) + int(has_dn)) * n * 8
        self.messages += int(has_up) + int(has_dn)

    def _stage(self, t):
        # gloo cannot move CUDA tensors point-to-point: stage through host memory
        return t.cpu() if (self.backend == "gloo" and t.is_cuda) else t

    def exchange(self, x, L: SlabLevel, k: int) -> None:
        """Fill the innermost k halo rows on both sides of x from the neighbours' owned rows."""
        if self.world == 1 or k == 0:
            return
        if getattr(self, "peer", None) is not None and x.is_cuda:
            return self._exchange_peer(x, L, k)
        dist = self.dist
        ops, recvs = [], []
        own0, own1 = L.h_top, L.h_top + (L.g1 - L.g0)
        direct = not (self.backend == "gloo" and x.is_cuda)  # row blocks are contiguous views: no staging copies

        def post(send_rows, recv_rows, peer):
            send = x[send_rows] if direct else x[send_rows].cpu()
            recv = x[recv_rows] if direct else self.torch.empty_like(send)
            ops.append(dist.P2POp(dist.isend, send, peer, self.group))
            ops.append(dist.P2POp(dist.irecv, recv, peer, self.group))
            recvs.append((recv, recv_rows))

        if L.h_top:  # upper neighbour: send my first k owned rows, receive its last k owned rows
            post(slice(own0, own0 + k), slice(own0 - k, own0), self.rank - 1)
        if L.h_bot:
            post(slice(own1 - k, own1), slice(own1, own1 + k), self.rank + 1)
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        for recv, sl in recvs:
            if not direct:
                x[sl].copy_(recv)
            self.bytes_sent += recv.numel() * 8
            self.messages += 1

    def gather_rows(self, owned, rows_per_rank: int, nz: int):
        """All-gather the owned rows of a level into the full (nz, nr) array on every rank."""
        if self.world == 1:
            return owned
        t = self._stage(owned[:rows_per_rank].contiguous())
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t, group=self.group)
        last = self._stage(owned[rows_per_rank:rows_per_rank + 1].contiguous()) if self.rank == self.world - 1 \
            else self.torch.empty((1, owned.shape[1]), dtype=owned.dtype, device=t.device)
        self.dist.broadcast(last, src=self._global_rank(self.world - 1), group=self.group)
        full = self.torch.cat(parts + [last], dim=0)
        assert full.shape[0] == nz
        return full.to(owned.device)

    def _global_rank(self, group_rank: int) -> int:
        return self.dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    def max(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64)
        if self.backend == "nccl":
            t = t.cuda()
        # NaN must win like np.max: reduce a NaN flag alongside
        flag = self.torch.tensor([1.0 if v != v else 0.0], dtype=self.torch.float64, device=t.device)
        t = self.torch.nan_to_num(t, nan=0.0)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX, group=self.group)
        return float("nan") if flag.item() > 0 else float(t.item())


class SlabMultigrid:
    """``multigrid_solve`` for one rank's slab.  API mirrors ``multigrid_solve.py:352`` except that
    ``source`` / ``psi_bc`` are this rank's OWNED rows (``owned_rows()``) of the global arrays."""

    def __init__(self, nz: int, nr: int, r_min: float, r_max: float, z_min: float, z_max: float, comm: SlabComm,
                 ops: Any, *, halo: int | None = None, min_rows: int = 32, omega: float = 1.0, pre_smooth: int = 3,
                 post_smooth: int = 3, min_grid: int = 5, use_graph: bool = True, strict_graph: bool = False,
                 gather_nz: int = 129):
        self.use_graph, self.strict_graph, self.used_graph = use_graph, strict_graph, False
        import os as _os
        self.use_native = _os.environ.get("GSB_SLAB_NATIVE", "1") != "0"  # C driver of the distributed levels
        self._native = None
        self._state: dict = {}
        # one halo exchange per level and V-cycle: the halo must survive the pre-smoothing (2 rows per
        # sweep become invalid), still hold the 2 rows the residual needs and the 2*post rows the
        # post-smoothing consumes (those rows receive the prolonged correction redundantly)
        need = 2 * pre_smooth + max(2, 2 * post_smooth)
        halo = need if halo is None else halo
        if halo < need:
            raise ValueError(f"halo must be >= 2*pre_smooth + max(2, 2*post_smooth) = {need}")
        self.comm, self.ops = comm, ops
        self.omega, self.pre, self.post, self.min_grid, self.halo = omega, pre_smooth, post_smooth, min_grid, halo
        self.levels, self.gathered = plan_slab_levels(nz, nr, r_min, r_max, z_min, z_max, comm.world, comm.rank,
                                                      halo=halo, min_rows=min_rows, min_grid=min_grid,
                                                      gather_nz=gather_nz)
        if not self.levels:
            raise ValueError("grid too small for a slab decomposition at this world size (use multigrid_solve)")
        self.nz, self.nr = nz, nr

    def owned_rows(self) -> tuple[int, int]:
        L = self.levels[0]
        return L.g0, L.g1

    # ---- one V-cycle on distributed level l; x and f are local arrays (halo rows included) ----
    def _vcycle(self, l: int, x, f):
        L = self.levels[l]
        ops, comm = self.ops, self.comm
        # precondition: x and f hold `halo` valid rows beyond the owned rows (level 0: exchanged by solve();
        # coarse levels: x starts as zeros everywhere, f = d was exchanged by the caller)
        x_home = x
        x = ops.smooth(L, x, f, self.omega, self.pre)  # 2*pre halo rows per side are now invalid
        nzc, nrc = (L.nz + 1) // 2, (L.nr + 1) // 2
        dist_next = l + 1 < len(self.levels)
        if dist_next:
            C = self.levels[l + 1]
            c_rows, c_row0, c_htop = C.rows_loc, C.row0, C.h_top
            cg0, cg1 = C.g0, C.g1
        else:
            G = self.gathered
            rpr = G["rows_per_rank"]
            cg0 = comm.rank * rpr
            cg1 = (comm.rank + 1) * rpr + (1 if comm.rank == comm.world - 1 else 0)
            c_htop = 0
            c_rows, c_row0 = cg1 - cg0, cg0
        roff = 2 * c_row0 - L.row0  # fine local row of coarse local row I is 2*I + roff
        d = ops.zeros((c_rows, nrc))
        ci0, ci1 = max(1, cg0) - c_row0, min(nzc - 1, cg1) - c_row0  # owned coarse rows that are not global walls
        ops.residual_restrict(L, None, x, f, d, roff, ci0, ci1)
        if dist_next:
            comm.exchange(d, C, self.halo)  # the halo rows are re-smoothed redundantly and need their rhs
            e = ops.zeros((c_rows, nrc))
            e = self._vcycle(l + 1, e, d)
            comm.exchange(e, C, self.post + 1)  # enough coarse rows to prolong into 2*post fine halo rows
            e_loc, e_row0 = e, c_row0
        else:
            d_full = comm.gather_rows(d, rpr, nzc)
            e_full = ops.coarse_vcycle(self.gathered, d_full, self.omega, self.pre, self.post, self.min_grid)
            lo, hi = max(0, cg0 - self.post - 1), min(nzc, cg1 + self.post + 1)
            e_loc, e_row0 = e_full[lo:hi].contiguous(), lo
            roff = 2 * e_row0 - L.row0
        # owned rows plus the 2*post halo rows the post-smoothing will consume (the neighbour adds the same
        # correction to the same rows: redundant, bit-identical), never global walls
        ext_top = 2 * self.post if L.h_top else 0
        ext_bot = 2 * self.post if L.h_bot else 0
        fi0, fi1 = max(1, L.g0 - ext_top) - L.row0, min(L.nz - 1, L.g1 + ext_bot) - L.row0
        ops.prolong_add(L, x, e_loc, roff, fi0, fi1)
        x = ops.smooth(L, x, f, self.omega, self.post, out=x_home)
        return x

    # ---- native driver: all launches of the distributed levels from two C calls per V-cycle ----
    def _coarse_geometry(self, l: int):
        """(c_rows, c_row0, cg0, cg1, nzc, nrc) of the array that level l restricts into."""
        L, comm = self.levels[l], self.comm
        nzc, nrc = (L.nz + 1) // 2, (L.nr + 1) // 2
        if l + 1 < len(self.levels):
            C = self.levels[l + 1]
            return C.rows_loc, C.row0, C.g0, C.g1, nzc, nrc
        rpr = self.gathered["rows_per_rank"]
        cg0 = comm.rank * rpr
        cg1 = (comm.rank + 1) * rpr + (1 if comm.rank == comm.world - 1 else 0)
        return cg1 - cg0, cg0, cg0, cg1, nzc, nrc

    def _native_setup(self, x0, f0) -> bool:
        ops, comm = self.ops, self.comm
        if not isinstance(ops, CudaSlabOps) or not self.use_native:
            return False
        if comm.world > 1 and getattr(comm, "peer", None) is None:
            return False  # NCCL point-to-point exchanges are issued from Python
        D, lib = ops.D, _lib.load()
        n = len(self.levels)
        descs = (_LevelDesc * n)()
        keep = []
        for l, L in enumerate(self.levels):
            ctx = ops._context(L)
            x = x0 if l == 0 else ops.zeros((L.rows_loc, L.nr))
            f = f0 if l == 0 else ops.zeros((L.rows_loc, L.nr))
            single = all(lib.gsb_slab_single_tile(ctx.handle, s) for s in (1, 2, 3))
            alt = None if single else D.empty((L.rows_loc, L.nr), ops.device)
            c_rows, c_row0, cg0, cg1, nzc, nrc = self._coarse_geometry(l)
            ext_top = 2 * self.post if L.h_top else 0
            ext_bot = 2 * self.post if L.h_bot else 0
            d = descs[l]
            d.ctx, d.x, d.f = ctx.handle.value if hasattr(ctx.handle, "value") else ctx.handle, x.data_ptr(), f.data_ptr()
            d.alt, d.cur = (alt.data_ptr() if alt is not None else None), x.data_ptr()
            d.rows_loc, d.nr, d.own0, d.own1 = L.rows_loc, L.nr, L.h_top, L.h_top + (L.g1 - L.g0)
            d.has_up, d.has_dn, d.row0 = int(bool(L.h_top)), int(bool(L.h_bot)), L.row0
            d.roff, d.ci0, d.ci1 = 2 * c_row0 - L.row0, max(1, cg0) - c_row0, min(nzc - 1, cg1) - c_row0
            d.nzc_loc, d.nrc = c_rows, nrc
            d.fi0, d.fi1 = max(1, L.g0 - ext_top) - L.row0, min(L.nz - 1, L.g1 + ext_bot) - L.row0
            keep += [ctx, x, f, alt]
        c_rows, c_row0, cg0, cg1, nzc, nrc = self._coarse_geometry(n - 1)
        d_last = ops.zeros((c_rows, nrc))
        lo, hi = max(0, cg0 - self.post - 1), min(nzc, cg1 + self.post + 1)
        halo = None
        if comm.world > 1:
            P = comm.peer
            up, dn = P["peers"].get("up"), P["peers"].get("dn")
            halo = _HaloDesc()
            halo.inbox_up, halo.inbox_dn = P["own"]["inbox_up"], P["own"]["inbox_dn"]
            halo.up_inbox_dn = up["inbox_dn"] if up else None
            halo.dn_inbox_up = dn["inbox_up"] if dn else None
            halo.flags_local = P["own"]["flags"]
            halo.flags_up = up["flags"] if up else None
            halo.flags_dn = dn["flags"] if dn else None
            halo.counters, halo.epochs, halo.cap = P["counters"].data_ptr(), P["epochs"].data_ptr(), P["cap"]
        self._native = {"descs": descs, "keep": keep, "d_last": d_last, "lo": lo, "hi": hi, "nzc": nzc,
                        "roff_last": 2 * lo - self.levels[-1].row0, "halo": halo, "lib": lib,
                        "d_full": ops.zeros((nzc, nrc)) if getattr(comm, "gather", None) is not None else None}
        return True

    def _vcycle_native(self) -> None:
        N, ops, comm = self._native, self.ops, self.comm
        D, lib = ops.D, N["lib"]
        n = len(self.levels)
        hp = ctypes.byref(N["halo"]) if N["halo"] is not None else None
        st = D.stream_ptr()
        _lib.check(lib.gsb_slab_down(N["descs"], n, D.ptr(N["d_last"]), hp, self.halo, self.omega, self.pre, st),
                   "gsb_slab_down")
        if N["d_full"] is not None:  # NVLink peer-memory gather: own kernels only (graph replayable on every rank)
            d_full = N["d_full"]
            comm.gather_rows_peer(N["d_last"], self.gathered["rows_per_rank"], N["nzc"], d_full)
        else:
            d_full = comm.gather_rows(N["d_last"], self.gathered["rows_per_rank"], N["nzc"])
        e_full = ops.coarse_vcycle(self.gathered, d_full, self.omega, self.pre, self.post, self.min_grid)
        e_loc = e_full[N["lo"]:N["hi"]]  # whole rows of a contiguous array: no copy
        N["e_keep"] = e_loc
        _lib.check(lib.gsb_slab_up(N["descs"], n, D.ptr(e_loc), N["hi"] - N["lo"], N["roff_last"], hp, self.post + 1,
                                   self.omega, self.post, st), "gsb_slab_up")

    def solve(self, source_owned, psi_bc_owned, *, tol: float = 1e-6, max_cycles: int = 500):
        """Returns (psi_owned_rows, residual_linf, n_cycles, converged) - residual/cycles/converged are
        global (identical on every rank), like multigrid_solve.py:352-463."""
        if not (np.isfinite(tol) and tol > 0):
            raise ValueError("tol must be finite and > 0")
        if max_cycles < 1:
            raise ValueError("max_cycles must be >= 1")
        L = self.levels[0]
        ops, comm = self.ops, self.comm
        n_own = L.g1 - L.g0
        own = slice(L.h_top, L.h_top + n_own)
        # persistent level-0 buffers: a captured graph (below) refers to their addresses, so repeated
        # solves on the same SlabMultigrid reuse one graph
        st = self._state
        gen = getattr(comm, "peer_generation", 0)
        if st and st.get("peer_generation") != gen:
            # the halo transport changed under us (enable/disable_peer_halo): descriptors and graphs built for
            # the old inboxes / flags would touch freed or unmapped memory - drop them
            st.clear()
            self._native = None
        if not st:
            st["f"], st["bc"], st["x"] = (ops.zeros((L.rows_loc, L.nr)) for _ in range(3))
            st["graph"] = None
            st["peer_generation"] = gen
            st["native"] = self._native_setup(st["x"], st["f"])
        f, bc = st["f"], st["bc"]
        f.zero_()
        bc.zero_()
        f[own] = ops.from_numpy(np.asarray(source_owned, dtype=np.float64)) if isinstance(source_owned, np.ndarray) else source_owned
        bc[own] = ops.from_numpy(np.asarray(psi_bc_owned, dtype=np.float64)) if isinstance(psi_bc_owned, np.ndarray) else psi_bc_owned
        comm.exchange(f, L, self.halo)
        x = st["x"]
        x.copy_(bc)
        r0, r1 = max(1, L.g0) - L.row0, min(L.nz - 1, L.g1) - L.row0
        comm.exchange(x, L, self.halo)  # serves the residual (1 row) and the next V-cycle (all rows)
        # convergence norm: with the peer-memory transports every rank publishes its max |r| to all ranks from the
        # stream and the host polls pinned memory (no NCCL launch, no stream synchronisation in the loop)
        use_norm = bool(st["native"]) and getattr(comm, "norm", None) is not None and getattr(x, "is_cuda", False)

        def norm_issue():
            ops.residual_linf_async(L, x, f, r0, r1, comm.norm["local"])
            comm.norm_exchange_issue()

        def norm_wait():
            comm.norm["issued"] += 1
            return comm.norm_exchange_wait(comm.norm["issued"])

        if use_norm:
            norm_issue()
            residual = norm_wait()
        else:
            residual = comm.max(ops.residual_linf(L, x, f, r0, r1))
        cycles = 0
        def cycle(xc):
            if st["native"]:
                self._vcycle_native()  # works in place on the persistent level-0 buffers (xc is st["x"])
            else:
                xc = self._vcycle(0, xc, f)
            # Dirichlet ring (multigrid_solve.py:437-441,458)
            xc[own, 0] = bc[own, 0]
            xc[own, -1] = bc[own, -1]
            if L.g0 == 0:
                xc[L.h_top] = bc[L.h_top]
            if L.g1 == L.nz:
                xc[L.h_top + n_own - 1] = bc[L.h_top + n_own - 1]
            comm.exchange(xc, L, self.halo)
            if use_norm and xc.data_ptr() == x.data_ptr():
                norm_issue()  # part of the (captured) cycle: residual of the new iterate + its exchange
            return xc

        # CUDA graph: the first cycle runs eagerly (it also creates every context and buffer); from the
        # second cycle on one graph replay issues the whole V-cycle - kernels, halo pushes over peer memory,
        # the coarse gather - without any host work in between (the host only reads the residual).
        graph = st["graph"]
        # multi-rank graphs (NCCL gather + spin-wait halo kernels inside one captured graph) hung on the
        # 2/4-GPU box in round 1 and are therefore opt-in (GSB_SLAB_MULTI_RANK_GRAPH=1) until understood
        import os as _os
        # ... with the peer-memory gather the cycle holds no NCCL node any more and replays on every rank
        all_own = bool(st["native"]) and getattr(comm, "gather", None) is not None and getattr(comm, "peer", None) is not None
        multi_ok = comm.world == 1 or _os.environ.get("GSB_SLAB_MULTI_RANK_GRAPH", "1" if all_own else "0") == "1"
        want_graph = self.use_graph and multi_ok and getattr(x, "is_cuda", False) and graph is None
        import time as _time
        trace = _os.environ.get("GSB_SLAB_TRACE") == "1"
        t_launch = t_wait = 0.0
        while not residual < tol and cycles < max_cycles:
            issued_in_cycle = True
            _t0 = _time.perf_counter()
            if graph is not None:
                graph.replay()
            else:
                x2 = cycle(x)
                if x2.data_ptr() != x.data_ptr() if hasattr(x2, "data_ptr") else x2 is not x:
                    x.copy_(x2)  # multi-launch smoothing phases may end in a pooled buffer
                    issued_in_cycle = False
                if want_graph and cycles == 0:
                    import torch
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    try:
                        # thread_local: the NCCL watchdog thread may query events while this thread captures
                        with torch.cuda.graph(g, capture_error_mode="thread_local" if comm.world > 1 else "global"):
                            x2 = cycle(x)
                        if x2.data_ptr() == x.data_ptr():
                            graph = g
                    except Exception:
                        if self.strict_graph:
                            raise
                        graph = None
                    if comm.world > 1:
                        # all ranks replay or none does: a rank whose capture failed would otherwise issue the
                        # cycle eagerly against peers replaying it
                        if comm.max(0.0 if graph is not None else 1.0) > 0.0:
                            graph = None
                    st["graph"] = graph
                    want_graph = graph is not None
            _t1 = _time.perf_counter()
            if use_norm:
                if not issued_in_cycle:
                    norm_issue()
                residual = norm_wait()
            else:
                residual = comm.max(ops.residual_linf(L, x, f, r0, r1))
            t_launch += _t1 - _t0
            t_wait += _time.perf_counter() - _t1
            cycles += 1
        if trace and comm.rank == 0:
            print(f"[slab trace] cycles {cycles}: host time issuing cycles {t_launch * 1e3:.2f} ms, waiting for the norm "
                  f"{t_wait * 1e3:.2f} ms (graph={graph is not None}, polled norm={use_norm})", flush=True)
        self.used_graph = graph is not None
        return x[own].clone(), residual, cycles, bool(residual < tol)
