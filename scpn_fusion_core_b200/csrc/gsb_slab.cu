// gsb_slab.cu - operators on a Z-row slab of one large grid (single-grid domain
// decomposition, one process per GPU).  A slab is a local array [rows_loc][nr] = halo rows + owned
// rows + halo rows of a level whose global shape is 2^k+1 in Z; the host side
// (scpn_fusion_core_b200/slab.py) exchanges halo rows between neighbouring ranks with NCCL and
// calls these entry points on its own rows.  Every kernel uses the same point functions as the
// single-GPU path, so a slab solve is bit-identical to the single-GPU solve:
//   * smoothing: the temporally blocked sweep of gsb_sweep.cu on the local array with the global
//     row offset as colour-parity offset; with h >= 2*sweeps halo rows the owned rows are exact;
//   * residual + full weighting, prolongation + add: one thread per point with a row offset between
//     the local fine and coarse indices (fine local row = 2 * coarse local row + roff).
#include "gsb_internal.cuh"

#include <cstring>

namespace gsb {

// coarse local rows [ci0, ci1), all coarse columns; column walls get 0 (multigrid_solve.py:303-306)
__global__ void __launch_bounds__(256)
k_slab_residual_restrict(LevelGeom g, const double *__restrict__ psi, const double *__restrict__ src,
                         double *__restrict__ dc, int nrc, int roff, int ci0, int ci1) {
  const int J = blockIdx.x * blockDim.x + threadIdx.x;
  const int I = ci0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (I >= ci1 || J >= nrc) return;
  double v = 0.0;
  if (J > 0 && J < nrc - 1) {
    double d[3][3];
#pragma unroll
    for (int a = -1; a <= 1; ++a)
#pragma unroll
      for (int c = -1; c <= 1; ++c) {
        const int iz = 2 * I + roff + a, ir = 2 * J + c;
        const double *q = psi + (size_t)iz * g.nr + ir;
        const double r = dsub(gs_apply(g, ir, q[0], q[1], q[-1], q[-g.nr], q[g.nr]), src[(size_t)iz * g.nr + ir]);
        d[a + 1][c + 1] = -r;
      }
    v = fw9(d[1][1], d[0][1], d[2][1], d[1][0], d[1][2], d[0][0], d[0][2], d[2][0], d[2][2]);
  }
  dc[(size_t)I * nrc + J] = v;
}

// psi[fine local rows fi0..fi1) interior columns] += P e   (prolongate_bilinear, odd sizes)
__global__ void __launch_bounds__(256)
k_slab_prolong_add(const double *__restrict__ ec, int nrc, double *__restrict__ psi, int nrf, int roff, int fi0,
                   int fi1) {
  const int ir = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int lf = fi0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (lf >= fi1 || ir >= nrf - 1) return;
  const int v = lf - roff;  // = 2 * coarse local row (+1 for rows between two coarse rows)
  const int I = v >> 1, J = ir >> 1;
  const bool ze = (v & 1) == 0, re = (ir & 1) == 0;
  const double *p = ec + (size_t)I * nrc + J;
  double out;
  if (ze && re)
    out = p[0];
  else if (ze)
    out = dmul(0.5, dadd(p[0], p[1]));
  else if (re)
    out = dmul(0.5, dadd(p[0], p[nrc]));
  else
    out = dmul(0.25, dadd(dadd(dadd(p[0], p[nrc]), p[1]), p[nrc + 1]));
  double *q = psi + (size_t)lf * nrf + ir;
  q[0] = dadd(q[0], out);
}

// max |L psi - src| over local rows [row0,row1), interior columns; out accumulates with atomicMax on
// the bit pattern (non-negative doubles order like unsigned integers) - order independent
__global__ void __launch_bounds__(128)
k_slab_residual_linf(LevelGeom g, const double *__restrict__ psi, const double *__restrict__ src, int row0, int row1,
                     unsigned long long *__restrict__ out) {
  __shared__ double sh[32];
  double m = 0.0;
  // grid (column chunks, row bands): a thread keeps its column and walks down the rows of its band with a
  // three-row register window (one new psi row per step, coefficient lookups once per thread, no integer division)
  const int ir = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int nb = gridDim.y, band = blockIdx.y;
  const int z0 = row0 + (int)((long long)(row1 - row0) * band / nb), z1 = row0 + (int)((long long)(row1 - row0) * (band + 1) / nb);
  if (ir <= g.nr - 2 && z0 < z1) {
    const double rs = g.r_safe[ir], irs = g.inv_r_safe[ir];
    const double *q = psi + (size_t)z0 * g.nr + ir;
    const double *f = src + (size_t)z0 * g.nr + ir;
    double sm = q[-g.nr], cw = q[-1], c0 = q[0], ce = q[1];
    for (int iz = z0; iz < z1; ++iz) {
      const double nn = q[g.nr];
      const double r = dsub(gs_apply_v(g, rs, irs, c0, ce, cw, sm, nn), f[0]);
      const double ar = fabs(r);
      if (ar > m || isnan(ar)) m = ar;  // a NaN residual must surface (np.max propagates NaN)
      q += g.nr;
      f += g.nr;
      sm = c0;
      c0 = nn;
      cw = q[-1];
      ce = q[1];
    }
  }
  // block_max uses fmax (drops NaN): carry a NaN flag alongside
  const int anynan = __syncthreads_or(isnan(m) ? 1 : 0);
  m = block_max(isnan(m) ? 0.0 : m, sh);
  if (threadIdx.x == 0) {
    if (anynan)
      atomicMax(out, 0x7ff8000000000000ULL);
    else
      atomicMax(out, (unsigned long long)__double_as_longlong(m));
  }
}

// ------------------------------------------------------------------------------------------
// Halo exchange over NVLink peer memory (no NCCL launch): every rank exposes two inboxes and a
// flag block through CUDA IPC.  push: wait until the neighbour has consumed the previous message,
// store my boundary rows straight into the neighbour's inbox (posted remote writes), fence, then
// raise the neighbour's ready flag.  recv: spin on my own (local) ready flag, copy inbox -> halo
// rows, then tell the sender it may overwrite the inbox.  Epochs count exchanges; both sides issue
// the same sequence, so the flags are monotone.  Multi-CTA: the last CTA to finish (atomic counter)
// publishes the flag.
// flags (int64, in the OWNER's memory): [0] ready_from_up  [1] ready_from_dn   (written by neighbours)
//                                        [2] consumed_by_up [3] consumed_by_dn  (written by neighbours)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ld_sys(const long long *p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(long long *p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct HaloDir {
  const double *src;       // push: my boundary rows            recv: my inbox
  double *dst;             // push: the neighbour's inbox       recv: my halo rows
  const long long *wait;   // flag to wait on (local memory)
  long long *signal;       // flag to raise afterwards (push: neighbour's ready; recv: neighbour's consumed)
  long long *epoch;        // local device counter of completed exchanges in this direction (graph-replay safe:
                           // the epoch is read on the device, not baked into the launch)
  int wait_lag;            // wait until *wait >= epoch - wait_lag  (push: 1 = previous message consumed; recv: 0)
  int *counter;            // local CTA-arrival counter of this direction
};
struct HaloArgs {
  HaloDir d[2];
  long long n;      // doubles per direction
};

__global__ void __launch_bounds__(512) k_halo_copy(const HaloArgs a) {
  const HaloDir &D = a.d[blockIdx.y];
  if (!D.src) return;
  __shared__ long long s_epoch;
  if (threadIdx.x == 0) {
    const long long e = ld_sys(D.epoch) + 1;  // this exchange's number; bumped by the last CTA below
    s_epoch = e;
    while (ld_sys(D.wait) < e - D.wait_lag) __nanosleep(64);
  }
  __syncthreads();
  const long long epoch = s_epoch;
  const long long per = (a.n + gridDim.x - 1) / gridDim.x;
  const long long i0 = per * blockIdx.x, i1 = min(a.n, i0 + per);
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) D.dst[i] = __ldcv(D.src + i);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(D.counter, 1);
    if (arrived == (int)gridDim.x - 1) {  // last CTA of this direction: every CTA has read the epoch by now
      *D.counter = 0;
      *D.epoch = epoch;
      __threadfence_system();
      st_sys(D.signal, epoch);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Coarse-level gather over NVLink peer memory: every rank stores its owned rows of the coarsest distributed
// right-hand side straight into EVERY rank's copy of the full level (all-to-all posted writes, one flag per
// source), replacing the NCCL all-gather of the V-cycle - the cycle then consists of this library's kernels only
// and can be replayed from a CUDA graph on every rank.  Buffers are double-buffered by exchange parity; the
// per-cycle residual all-reduce keeps ranks within one cycle of each other, so no consumed-acknowledgement is
// needed.  The exchange number lives on the device (graph-replay safe), bumped by k_gather_wait.
// ------------------------------------------------------------------------------------------
constexpr int kGatherMaxWorld = 16;
struct GatherPushArgs {
  const double *rows;            // my owned rows (contiguous)
  long long n, off, buf_doubles; // doubles to send, offset inside the level, doubles per buffer half
  double *peer_buf[kGatherMaxWorld];
  long long *peer_flags[kGatherMaxWorld];
  int world, rank;
  int *counters;                 // [world] CTA-arrival counters (zeroed once)
  const long long *epoch;        // completed gathers
};
__global__ void __launch_bounds__(512) k_gather_push(const GatherPushArgs a) {
  const int p = blockIdx.y;
  const long long e = ld_sys(a.epoch) + 1;
  double *dst = a.peer_buf[p] + (e & 1) * a.buf_doubles + a.off;
  const long long per = (a.n + gridDim.x - 1) / gridDim.x;
  const long long i0 = per * blockIdx.x, i1 = min(a.n, i0 + per);
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = a.rows[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(a.counters + p, 1);
    if (arrived == (int)gridDim.x - 1) {
      a.counters[p] = 0;
      __threadfence_system();
      st_sys(a.peer_flags[p] + a.rank, e);
    }
  }
}
// wait for all `world` sources of this gather, copy the assembled level into `out`, bump the epoch
__global__ void __launch_bounds__(512) k_gather_wait(const double *buf, long long buf_doubles, const long long *flags,
                                                     int world, long long n_total, double *out, long long *epoch,
                                                     volatile long long *host_flag) {
  __shared__ long long s_e;
  if (threadIdx.x == 0) s_e = ld_sys(epoch) + 1;
  __syncthreads();
  const long long e = s_e;
  if ((int)threadIdx.x < world)
    while (ld_sys(flags + threadIdx.x) < e) __nanosleep(64);
  __syncthreads();
  const double *src = buf + (e & 1) * buf_doubles;
  for (long long i = threadIdx.x; i < n_total; i += blockDim.x) out[i] = __ldcv(src + i);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *epoch = e;
    if (host_flag) *host_flag = e;  // `out` (and this flag) may be mapped pinned host memory: the host polls, no stream sync
  }
}

static int slab_plan(gsb_ctx *ctx) {
  // a slab context has exactly one level: the local array itself
  return ensure_plan(ctx, 1 << 30);
}

}  // namespace gsb

using namespace gsb;

extern "C" {

// Let kernels running on `device` dereference memory that lives on `peer_device` (needed before IPC-opened
// neighbour buffers can be used by gsb_halo_push / gsb_halo_recv).  Returns GSB_OK if access is (already) enabled.
int gsb_enable_peer_access(int device, int peer_device) {
  if (device == peer_device) return GSB_OK;
  int can = 0;
  GSB_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
  GSB_REQUIRE(can, "gsb_enable_peer_access: the two devices have no peer-to-peer path");
  GSB_CUDA(cudaSetDevice(device));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  GSB_CUDA(e);
  return GSB_OK;
}

// One cudaMalloc'ed, zero-filled block that other processes can map (CUDA IPC): returns the local pointer
// and the 64-byte cudaIpcMemHandle_t to hand to the neighbours.
int gsb_ipc_alloc(int device, long long bytes, void **ptr_out, unsigned char *handle_out64) {
  GSB_REQUIRE(ptr_out && handle_out64 && bytes > 0, "gsb_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  GSB_CUDA(cudaSetDevice(device));
  void *p = nullptr;
  GSB_CUDA(cudaMalloc(&p, (size_t)bytes));
  GSB_CUDA(cudaMemset(p, 0, (size_t)bytes));
  GSB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    GSB_CUDA(e);
  }
  memcpy(handle_out64, &h, 64);
  *ptr_out = p;
  return GSB_OK;
}
// Map a neighbour's block into this process with `device` current (peer access is enabled lazily by the driver).
int gsb_ipc_open(int device, const unsigned char *handle64, void **ptr_out) {
  GSB_REQUIRE(ptr_out && handle64, "gsb_ipc_open: bad argument");
  GSB_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  GSB_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return GSB_OK;
}
int gsb_ipc_close(void *ptr) {
  if (ptr) GSB_CUDA(cudaIpcCloseMemHandle(ptr));
  return GSB_OK;
}
int gsb_ipc_free(void *ptr) {
  if (ptr) GSB_CUDA(cudaFree(ptr));
  return GSB_OK;
}

// Push `n` doubles per direction into the neighbours' inboxes (NULL row pointer = no neighbour on that
// side).  flags_local: this rank's flag block; flags_up/flags_dn: the neighbours' flag blocks and
// inbox_*: their inboxes (peer pointers opened through CUDA IPC); counters: 4 ints and epochs: 4 int64 of local
// scratch (zeroed once): {push up, push dn, recv up, recv dn}.  The exchange number lives on the device, so a
// captured CUDA graph of a V-cycle can be replayed.
int gsb_halo_push(const double *rows_up, const double *rows_dn, long long n, double *up_inbox_dn, double *dn_inbox_up,
                  long long *flags_local, long long *flags_up, long long *flags_dn, int *counters, long long *epochs,
                  void *stream) {
  GSB_REQUIRE(flags_local && counters && epochs && n > 0, "gsb_halo_push: bad argument");
  HaloArgs a{};
  a.n = n;
  if (rows_up) {
    GSB_REQUIRE(up_inbox_dn && flags_up, "gsb_halo_push: upper neighbour buffers missing");
    a.d[0] = HaloDir{rows_up, up_inbox_dn, flags_local + 2, flags_up + 1, epochs + 0, 1, counters + 0};
  }
  if (rows_dn) {
    GSB_REQUIRE(dn_inbox_up && flags_dn, "gsb_halo_push: lower neighbour buffers missing");
    a.d[1] = HaloDir{rows_dn, dn_inbox_up, flags_local + 3, flags_dn + 0, epochs + 1, 1, counters + 1};
  }
  const int nblk = (int)std::min<long long>(16, (n + 8191) / 8192);
  k_halo_copy<<<dim3(nblk, 2), 512, 0, (cudaStream_t)stream>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

// Wait for the neighbours' pushes of `epoch`, copy my inboxes into my halo rows, release the inboxes.
int gsb_halo_recv(double *halo_up, double *halo_dn, long long n, const double *inbox_up, const double *inbox_dn,
                  long long *flags_local, long long *flags_up, long long *flags_dn, int *counters, long long *epochs,
                  void *stream) {
  GSB_REQUIRE(flags_local && counters && epochs && n > 0, "gsb_halo_recv: bad argument");
  HaloArgs a{};
  a.n = n;
  if (halo_up) {
    GSB_REQUIRE(inbox_up && flags_up, "gsb_halo_recv: upper neighbour buffers missing");
    a.d[0] = HaloDir{inbox_up, halo_up, flags_local + 0, flags_up + 3, epochs + 2, 0, counters + 2};
  }
  if (halo_dn) {
    GSB_REQUIRE(inbox_dn && flags_dn, "gsb_halo_recv: lower neighbour buffers missing");
    a.d[1] = HaloDir{inbox_dn, halo_dn, flags_local + 1, flags_dn + 2, epochs + 3, 0, counters + 3};
  }
  const int nblk = (int)std::min<long long>(16, (n + 8191) / 8192);
  k_halo_copy<<<dim3(nblk, 2), 512, 0, (cudaStream_t)stream>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

int gsb_gather_push(const double *rows_dev, long long n, long long off, long long buf_doubles, void *const *peer_bufs,
                    void *const *peer_flags, int world, int rank, int *counters, const long long *epoch, void *stream) {
  GSB_REQUIRE(rows_dev && peer_bufs && peer_flags && counters && epoch, "gsb_gather_push: NULL argument");
  GSB_REQUIRE(world >= 1 && world <= kGatherMaxWorld && rank >= 0 && rank < world, "gsb_gather_push: bad world / rank");
  GSB_REQUIRE(n > 0 && off >= 0 && off + n <= buf_doubles, "gsb_gather_push: rows outside the level");
  GatherPushArgs a{};
  a.rows = rows_dev;
  a.n = n, a.off = off, a.buf_doubles = buf_doubles;
  for (int p = 0; p < world; ++p) {
    GSB_REQUIRE(peer_bufs[p] && peer_flags[p], "gsb_gather_push: missing peer buffer");
    a.peer_buf[p] = static_cast<double *>(peer_bufs[p]);
    a.peer_flags[p] = static_cast<long long *>(peer_flags[p]);
  }
  a.world = world, a.rank = rank, a.counters = counters, a.epoch = epoch;
  const int nblk = (int)std::min<long long>(8, (n + 4095) / 4096);
  k_gather_push<<<dim3(nblk, world), 512, 0, (cudaStream_t)stream>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

int gsb_gather_wait(const double *buf_local, long long buf_doubles, const long long *flags_local, int world,
                    long long n_total, double *out_dev, long long *epoch, long long *host_flag, void *stream) {
  GSB_REQUIRE(buf_local && flags_local && out_dev && epoch, "gsb_gather_wait: NULL argument");
  GSB_REQUIRE(world >= 1 && world <= kGatherMaxWorld && n_total > 0 && n_total <= buf_doubles, "gsb_gather_wait: bad sizes");
  k_gather_wait<<<1, 512, 0, (cudaStream_t)stream>>>(buf_local, buf_doubles, flags_local, world, n_total, out_dev, epoch,
                                                     host_flag);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

int gsb_slab_single_tile(gsb_ctx *ctx, int sweeps) {
  if (!ctx || sweeps < 1 || sweeps > 3) return 0;
  int sc, br, ns, nb;
  sweep_fused_plan(ctx->nz, ctx->nr, 1, 2 * sweeps, ctx->num_sms, &sc, &br, &ns, &nb);
  return (ns == 1 && nb == 1) ? 1 : 0;
}

int gsb_slab_smooth(gsb_ctx *ctx, const double *in_dev, double *out_dev, const double *src_dev, double omega,
                    int sweeps, int par_off, void *stream) {
  GSB_REQUIRE(ctx && in_dev && out_dev && src_dev, "gsb_slab_smooth: NULL argument");
  GSB_REQUIRE(sweeps >= 1 && sweeps <= 3, "gsb_slab_smooth: 1..3 sweeps per call");
  GSB_REQUIRE(std::isfinite(omega) && omega >= 1.0 && omega < 2.0, "omega must be finite and satisfy 1.0 <= omega < 2.0");
  GSB_CUDA(cudaSetDevice(ctx->device));
  int rc = slab_plan(ctx);
  if (rc) return rc;
  return sweep_fused_launch(ctx->levels[0].g, in_dev, ctx->n, out_dev, ctx->n, src_dev, ctx->n, 1, omega, sweeps,
                            par_off & 1, ctx->num_sms, nullptr, (cudaStream_t)stream);
}

int gsb_slab_residual_restrict(gsb_ctx *fine, const double *x_dev, const double *src_dev, double *dc_dev,
                               int nzc_loc, int nrc, int roff, int ci0, int ci1, void *stream) {
  GSB_REQUIRE(fine && x_dev && src_dev && dc_dev, "gsb_slab_residual_restrict: NULL argument");
  GSB_REQUIRE(nrc == (fine->nr + 1) / 2 && (fine->nr & 1), "gsb_slab_residual_restrict: needs an odd fine width");
  GSB_REQUIRE(ci0 >= 0 && ci1 <= nzc_loc && ci0 <= ci1, "gsb_slab_residual_restrict: bad coarse row range");
  if (ci0 < ci1)
    GSB_REQUIRE(2 * ci0 + roff - 2 >= 0 && 2 * (ci1 - 1) + roff + 2 <= fine->nz - 1,
                "gsb_slab_residual_restrict: fine halo rows missing");
  GSB_CUDA(cudaSetDevice(fine->device));
  int rc = slab_plan(fine);
  if (rc) return rc;
  if (ci0 == ci1) return GSB_OK;
  return residual_restrict_tiled_launch(fine->levels[0].g, x_dev, fine->n, src_dev, fine->n, dc_dev, (size_t)nzc_loc * nrc,
                                        nzc_loc, nrc, roff, ci0, ci1, 0, 0, 1, nullptr, (cudaStream_t)stream);
}

int gsb_slab_prolong_add(gsb_ctx *fine, const double *ec_dev, int nzc_loc, int nrc, double *x_dev, int roff,
                         int fi0, int fi1, void *stream) {
  GSB_REQUIRE(fine && ec_dev && x_dev, "gsb_slab_prolong_add: NULL argument");
  GSB_REQUIRE(nrc == (fine->nr + 1) / 2 && (fine->nr & 1), "gsb_slab_prolong_add: needs an odd fine width");
  GSB_REQUIRE(fi0 >= 0 && fi1 <= fine->nz && fi0 <= fi1, "gsb_slab_prolong_add: bad fine row range");
  if (fi0 < fi1)
    GSB_REQUIRE(fi0 - roff >= 0 && ((fi1 - 1 - roff + 1) >> 1) <= nzc_loc - 1, "gsb_slab_prolong_add: coarse halo rows missing");
  GSB_CUDA(cudaSetDevice(fine->device));
  if (fi0 == fi1 || fine->nr < 3) return GSB_OK;
  const dim3 blk(32, 8, 1), grd((fine->nr - 2 + 31) / 32, (fi1 - fi0 + 7) / 8, 1);
  k_slab_prolong_add<<<grd, blk, 0, (cudaStream_t)stream>>>(ec_dev, nrc, x_dev, fine->nr, roff, fi0, fi1);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

int gsb_slab_residual_linf(gsb_ctx *ctx, const double *x_dev, const double *src_dev, int row0, int row1,
                           double *out_dev, void *stream) {
  GSB_REQUIRE(ctx && x_dev && src_dev && out_dev, "gsb_slab_residual_linf: NULL argument");
  GSB_REQUIRE(row0 >= 1 && row1 <= ctx->nz - 1 && row0 <= row1, "gsb_slab_residual_linf: rows must be interior to the local array");
  GSB_CUDA(cudaSetDevice(ctx->device));
  int rc = slab_plan(ctx);
  if (rc) return rc;
  if (row0 == row1 || ctx->nr < 3) return GSB_OK;
  const int cchunks = (ctx->nr - 2 + 127) / 128;
  const int bands = std::max(1, std::min(row1 - row0, (8 * ctx->num_sms + cchunks - 1) / cchunks));
  k_slab_residual_linf<<<dim3(cchunks, bands), 128, 0, (cudaStream_t)stream>>>(ctx->levels[0].g, x_dev, src_dev, row0, row1,
                                                                              reinterpret_cast<unsigned long long *>(out_dev));
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}


// ------------------------------------------------------------------------------------------
// Native V-cycle driver for the distributed levels: every launch of the descent (pre-smoothing,
// residual + restriction, halo exchange of the coarse right-hand side) and of the ascent (halo exchange
// of the coarse correction, prolongation, post-smoothing) is issued from ONE call each, instead of ~25
// host round trips per level from Python.  The gather + replicated coarse solve happens between the two
// calls on the host side.  Halos travel through gsb_halo_push/recv (halo == NULL: single rank).
// ------------------------------------------------------------------------------------------
static int slab_exchange(const gsb_slab_halo_desc *h, double *a, const gsb_slab_level_desc &L, int k, void *stream) {
  if (!h || k <= 0 || (!L.has_up && !L.has_dn)) return GSB_OK;
  const long long n = (long long)k * L.nr;
  GSB_REQUIRE(n <= h->cap, "slab exchange: inbox too small");
  GSB_REQUIRE(L.own1 - L.own0 >= k, "slab exchange: fewer owned rows than halo rows");
  int rc = gsb_halo_push(L.has_up ? a + (size_t)L.own0 * L.nr : nullptr, L.has_dn ? a + (size_t)(L.own1 - k) * L.nr : nullptr, n,
                         L.has_up ? h->up_inbox_dn : nullptr, L.has_dn ? h->dn_inbox_up : nullptr, h->flags_local,
                         L.has_up ? h->flags_up : nullptr, L.has_dn ? h->flags_dn : nullptr, h->counters, h->epochs, stream);
  if (rc) return rc;
  return gsb_halo_recv(L.has_up ? a + (size_t)(L.own0 - k) * L.nr : nullptr, L.has_dn ? a + (size_t)L.own1 * L.nr : nullptr, n,
                       L.has_up ? h->inbox_up : nullptr, L.has_dn ? h->inbox_dn : nullptr, h->flags_local,
                       L.has_up ? h->flags_up : nullptr, L.has_dn ? h->flags_dn : nullptr, h->counters, h->epochs, stream);
}

// `sweeps` sweeps from *cur (ping-pong between x and alt when the level is multi-tile); *cur follows the data
static int slab_smooth_pp(gsb_slab_level_desc &L, double **cur, double omega, int sweeps, void *stream) {
  int left = sweeps;
  while (left > 0) {
    const int s = left < 3 ? left : 3;
    double *dst = *cur;
    if (!gsb_slab_single_tile(L.ctx, s)) {
      dst = (*cur == L.x) ? L.alt : L.x;
      GSB_REQUIRE(dst != nullptr, "slab level needs an alternate buffer (multi-tile sweep)");
    }
    int rc = gsb_slab_smooth(L.ctx, *cur, dst, L.f, omega, s, L.row0, stream);
    if (rc) return rc;
    *cur = dst;
    left -= s;
  }
  return GSB_OK;
}

int gsb_slab_down(gsb_slab_level_desc *lev, int nlev, double *d_last, const gsb_slab_halo_desc *halo, int halo_rows,
                  double omega, int pre, void *stream) {
  GSB_REQUIRE(lev && nlev >= 1 && d_last, "gsb_slab_down: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int l = 0; l < nlev; ++l) {
    gsb_slab_level_desc &L = lev[l];
    double *cur = L.x;
    int rc = slab_smooth_pp(L, &cur, omega, pre, stream);
    if (rc) return rc;
    L.cur = cur;
    double *d = (l + 1 < nlev) ? lev[l + 1].f : d_last;
    GSB_CUDA(cudaMemsetAsync(d, 0, (size_t)L.nzc_loc * L.nrc * sizeof(double), st));
    rc = gsb_slab_residual_restrict(L.ctx, cur, L.f, d, L.nzc_loc, L.nrc, L.roff, L.ci0, L.ci1, stream);
    if (rc) return rc;
    if (l + 1 < nlev) {
      gsb_slab_level_desc &C = lev[l + 1];
      rc = slab_exchange(halo, C.f, C, halo_rows, stream);  // the halo rows are re-smoothed redundantly
      if (rc) return rc;
      GSB_CUDA(cudaMemsetAsync(C.x, 0, (size_t)C.rows_loc * C.nr * sizeof(double), st));
    }
  }
  return GSB_OK;
}

int gsb_slab_up(gsb_slab_level_desc *lev, int nlev, const double *e_last, int nze_last_loc, int roff_last,
                const gsb_slab_halo_desc *halo, int e_rows, double omega, int post, void *stream) {
  GSB_REQUIRE(lev && nlev >= 1 && e_last, "gsb_slab_up: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int l = nlev - 1; l >= 0; --l) {
    gsb_slab_level_desc &L = lev[l];
    double *cur = L.cur ? L.cur : L.x;
    const double *e = e_last;
    int roff = roff_last, nze = nze_last_loc;
    if (l + 1 < nlev) {
      gsb_slab_level_desc &C = lev[l + 1];
      int rc = slab_exchange(halo, C.x, C, e_rows, stream);  // C.cur == C.x after its post-smoothing
      if (rc) return rc;
      e = C.x;
      roff = L.roff;
      nze = L.nzc_loc;
    }
    int rc = gsb_slab_prolong_add(L.ctx, e, nze, L.nrc, cur, roff, L.fi0, L.fi1, stream);
    if (rc) return rc;
    rc = slab_smooth_pp(L, &cur, omega, post, stream);
    if (rc) return rc;
    if (cur != L.x) {  // odd number of out-of-place passes: bring the level home
      GSB_CUDA(cudaMemcpyAsync(L.x, cur, (size_t)L.rows_loc * L.nr * sizeof(double), cudaMemcpyDeviceToDevice, st));
      cur = L.x;
    }
    L.cur = L.x;
  }
  return GSB_OK;
}

}  // extern "C"
