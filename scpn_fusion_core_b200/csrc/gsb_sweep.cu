// gsb_sweep.cu - temporally blocked streaming RB-SOR: S full sweeps (2S colour passes) of mg_smooth
// (multigrid_solve.py:148-208) in ONE pass over HBM.
//
// The per-colour streaming kernel (k_smooth_colour) moves psi twice and the source once per
// colour pass: 48 B of DRAM traffic per point per sweep for 24 algorithmic bytes.  Here every WARP
// owns a 64-column strip of one row band of one equilibrium and marches down the rows on its own:
// no CTA barrier anywhere, only __syncwarp.  The warp keeps a private ring of rows in shared memory
// (filled by cp.async three rows ahead); lane k holds the column pair (2k, 2k+1).  Colour pass t
// (= stage t) works two rows behind pass t-1, so the 2S stages of a step are mutually independent:
//   stage t at row r = i - 2t reads rows r-1, r, r+1: pass t-1 finished r+1 one step earlier, and
//   pass t+1 (row r-2) touches nothing stage t reads;
// a lane therefore loads the operands of all 2S stages, computes 2S independent updates (ILP = 2S
// hides the FP64 latency) and stores them, once per row step.
// Strips/bands overlap by 2S columns/rows; the overlap is recomputed (an update is valid t+1 points
// inside the tile edge after pass t) and only tile interiors are written, so with more than one
// tile per equilibrium the sweep is out of place (neighbours read each other's interiors).
// The ring keeps each row as two half rows (even / odd columns) so operands are unit-stride across
// the lanes.  Arithmetic: the same operand order as sor_point (bit-identical results).
#include "gsb_internal.cuh"

namespace gsb {

constexpr int kSwCols = 64;  // columns per warp strip (2 per lane)
constexpr int kSwWPC = 4;    // warps per CTA (independent of each other)

struct SweepArgs {
  int nz, nr;               // array shape (local rows incl. halo rows in slab mode)
  int par_off;              // added to the local row index for the colour parity (global row offset)
  int band_rows, n_strips, n_bands;  // tile plan
  const double *in;
  double *out;
  const double *src;
  size_t istride, ostride, sstride;
  const double *a_e, *a_w;  // [nr] column tables
  double a_ns, a_c, inv_a_c, omega, omw;
  const int *active;
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// P0 = (zl + par_off + cl) & 1 of the warp's tile, hoisted into a template parameter by the
// dispatching kernel below: with the step loop unrolled over one ring period every ring slot and
// every colour parity is a compile-time constant (no address arithmetic in the loop body).
template <int NST, int P0>
__device__ __forceinline__ void sweep_warp_body(const SweepArgs &a, double *ring_p, double *ring_s, int lane, int zl,
                                                int zh, int z0, int z1, int cl, int wb, int c0, int c1,
                                                const double *gin, const double *gsrc, double *gout) {
  constexpr int NRING = 16;            // rows qi-2(NST-1)-1 .. qi+PF live at step qi
  constexpr int PF = NRING - 2 * NST;  // prefetch distance in rows (deeper for the shorter steps of small NST)
  const int nr = a.nr;
  const int xe = 2 * lane, xo = 2 * lane + 1;  // the lane's columns inside the strip
  const bool have_e = xe < wb, have_o = xo < wb;
  const bool upd_e = xe >= 1 && xe <= wb - 2, upd_o = xo <= wb - 2;
  const bool wr_e = have_e && cl + xe >= c0 && cl + xe < c1, wr_o = have_o && cl + xo >= c0 && cl + xo < c1;
  const double ae_e = have_e ? a.a_e[cl + xe] : 0.0, aw_e = have_e ? a.a_w[cl + xe] : 0.0;
  const double ae_o = have_o ? a.a_e[cl + xo] : 0.0, aw_o = have_o ? a.a_w[cl + xo] : 0.0;
  const double a_ns = a.a_ns, a_c = a.a_c, inv_a_c = a.inv_a_c, omega = a.omega, omw = a.omw;
  double *rp = ring_p + lane;  // ring slot of row zl + q is q & (NRING-1); element [slot][half][lane]
  double *rs = ring_s + lane;
  const int nrows = zh - zl;   // loaded rows, relative index q = r - zl in [0, nrows)

  auto load_row = [&](int q, int slot) {  // asynchronous copy of row zl+q (psi and source) into `slot`
    if (q < nrows) {
      const double *pr = gin + (size_t)(zl + q) * nr + cl;
      const double *sr = gsrc + (size_t)(zl + q) * nr + cl;
      if (have_e) {
        cp_async8(rp + (slot * 2) * 32, pr + xe);
        cp_async8(rs + (slot * 2) * 32, sr + xe);
      }
      if (have_o) {
        cp_async8(rp + (slot * 2 + 1) * 32, pr + xo);
        cp_async8(rs + (slot * 2 + 1) * 32, sr + xo);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int q = 0; q <= PF; ++q) load_row(q, q);
  cp_async_wait<PF - 2>();  // rows 0 .. 2 have landed
  __syncwarp();

  const int q_end = (nrows - 2) + 2 * (NST - 1);  // last step (relative row of stage 0)
  for (int base = 0; base <= q_end; base += NRING) {
#pragma unroll
    for (int u = 0; u < NRING; ++u) {
      const int qi = base + u;  // stage 0 is at relative row qi (row 0 is never updated)
      if (qi >= 1 && qi <= q_end) {
        double W[NST], E[NST], S[NST], N[NST], O[NST], F[NST], V[NST];
        bool on[NST];
        // ---- operands of every stage (stage t = colour pass t at relative row qi - 2t)
#pragma unroll
        for (int t = 0; t < NST; ++t) {
          constexpr int M = NRING - 1;
          const int q = qi - 2 * t;
          const int hx = (t + u + P0) & 1;           // column parity of this pass's points in that row
          const int slot = (u - 2 * t) & M, sl_s = (u - 2 * t - 1) & M, sl_n = (u - 2 * t + 1) & M;  // compile time
          on[t] = (q >= 1 && q <= nrows - 2) && (hx ? upd_o : upd_e);
          const double *me = rp + (slot * 2 + hx) * 32;        // own half row
          const double *ot = rp + (slot * 2 + (1 - hx)) * 32;  // the other colour's half row
          // unconditional loads (always inside the padded ring): no branches, so the 2S stages overlap
          W[t] = hx ? ot[0] : ot[-1];
          E[t] = hx ? ot[1] : ot[0];
          S[t] = rp[(sl_s * 2 + hx) * 32];
          N[t] = rp[(sl_n * 2 + hx) * 32];
          O[t] = me[0];
          F[t] = rs[(slot * 2 + hx) * 32];
        }
        // ---- 2S independent updates
#pragma unroll
        for (int t = 0; t < NST; ++t) {
          const int hx = (t + u + P0) & 1;
          double acc = dadd(dmul(hx ? ae_o : ae_e, E[t]), dmul(hx ? aw_o : aw_e, W[t]));
          acc = dadd(acc, dmul(a_ns, S[t]));
          acc = dadd(acc, dmul(a_ns, N[t]));
          acc = dsub(acc, F[t]);
          const double qq = __dmul_rn(acc, inv_a_c);
          const double gs = __fma_rn(__fma_rn(-a_c, qq, acc), inv_a_c, qq);  // acc / a_c (Markstein sequence, gsb_internal.cuh)
          V[t] = dadd(dmul(omw, O[t]), dmul(omega, gs));
        }
#pragma unroll
        for (int t = 0; t < NST; ++t) {
          const int hx = (t + u + P0) & 1;
          const int slot = (u - 2 * t) & (NRING - 1);
          if (on[t]) rp[(slot * 2 + hx) * 32] = V[t];
        }
        load_row(qi + PF, (u + PF) & (NRING - 1));
        cp_async_wait<PF - 2>();  // row qi+2 has landed (stage 0 of the next step reads it)
        __syncwarp();
        // ---- row w is final once the last pass has processed it: write the tile interior back
        const int w = zl + qi - 2 * (NST - 1);
        if (w >= z0 && w < z1) {
          const int slot = (u - 2 * (NST - 1)) & (NRING - 1);
          double *orow = gout + (size_t)w * nr + cl;
          if (wr_e) orow[xe] = rp[(slot * 2) * 32];
          if (wr_o) orow[xo] = rp[(slot * 2 + 1) * 32];
        }
      }
    }
  }
  // rows that no stage ever processes (next to the array edge: walls / halo rows)
  if (gout != gin) {
    for (int w = z0; w < z1; ++w) {
      if (w >= zl + 1 && w <= zh - 2) continue;
      const double *irow = gin + (size_t)w * nr + cl;
      double *orow = gout + (size_t)w * nr + cl;
      if (wr_e) orow[xe] = irow[xe];
      if (wr_o) orow[xo] = irow[xo];
    }
  }
}

template <int NST>
__global__ void __launch_bounds__(32 * kSwWPC) k_sweep_warp(const SweepArgs a) {
  constexpr int NRING = 16;
  constexpr int STEP = kSwCols - 2 * NST;  // interior columns per strip (even)
  extern __shared__ double sw_pool[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *ring_p = sw_pool + 8 + wid * (2 * NRING * kSwCols);  // [NRING][2][32]; +8: lane 0 may read element -1
  double *ring_s = ring_p + NRING * kSwCols;                   // [NRING][2][32]
  const int b = blockIdx.z;
  if (a.active && !a.active[b]) return;
  // (strip, band) tiles are numbered linearly and dealt four to a CTA, so CTAs are full whatever the strip count
  // (257 columns = 5 strips used to leave 3 of every 8 warp slots empty)
  const int tile = blockIdx.x * kSwWPC + wid;
  if (tile >= a.n_strips * a.n_bands) return;  // warps never synchronise with each other
  const int band = tile / a.n_strips, strip = tile - band * a.n_strips;
  const int nz = a.nz, nr = a.nr;
  // tile: interior rows [z0,z1) x cols [c0,c1); loaded rows [zl,zh) x cols [cl,ch)
  const int z0 = band * a.band_rows, z1 = min(nz, z0 + a.band_rows);
  const int c0 = strip * STEP, c1 = (a.n_strips == 1) ? nr : min(nr, c0 + STEP);  // one strip: nr <= 64
  const int zl = max(0, z0 - NST), zh = min(nz, z1 + NST);
  const int cl = max(0, c0 - NST), ch = min(nr, cl + kSwCols);  // cl is even (STEP and NST are)
  const int wb = ch - cl;
  const double *gin = a.in + (size_t)b * a.istride;
  const double *gsrc = a.src + (size_t)b * a.sstride;
  double *gout = a.out + (size_t)b * a.ostride;
  if ((zl + a.par_off + cl) & 1)
    sweep_warp_body<NST, 1>(a, ring_p, ring_s, lane, zl, zh, z0, z1, cl, wb, c0, c1, gin, gsrc, gout);
  else
    sweep_warp_body<NST, 0>(a, ring_p, ring_s, lane, zl, zh, z0, z1, cl, wb, c0, c1, gin, gsrc, gout);
}

static size_t sweep_smem_bytes(int nst) {
  const int nring = 16;
  (void)nst;
  return (size_t)(kSwWPC * 2 * nring * kSwCols + 8) * sizeof(double);
}

// Tile plan: strips of (64 - 2*NST) interior columns (one strip when the row fits 64 columns); row
// bands sized so that the grid has enough warps to fill the GPU when the batch alone does not.
void sweep_fused_plan(int nz, int nr, int batch, int nst, int num_sms, int *strip_cols, int *band_rows,
                      int *n_strips, int *n_bands) {
  const int step = kSwCols - 2 * nst;
  if (nr <= kSwCols) {
    *strip_cols = nr;
    *n_strips = 1;
  } else {
    *strip_cols = step;
    *n_strips = (nr + step - 1) / step;
  }
  const long long warps = (long long)*n_strips * batch;
  const long long want = 24LL * num_sms;
  int bands = 1;
  // a warp marches down its band row by row, ~1 us per row step when it is alone on its SM (measured on
  // 257^2..1025^2 single grids: 45 us per launch with 32-row bands, whatever the level): a level that cannot fill
  // the GPU is latency bound, so its bands shrink down to 8 rows - the 2*nst halo rows recomputed per band are
  // free there, and the results do not depend on the tiling.  (A cost model that also avoided band counts just
  // past a wave boundary was measured and made things worse: full SMs are throughput bound, r2_streaming_picard.md.)
  if (warps < want) bands = (int)std::min<long long>((want + warps - 1) / warps, std::max(1, nz / 8));
  const int br = (nz + bands - 1) / bands;
  *band_rows = br;
  *n_bands = (nz + br - 1) / br;
}

// S = sweeps (1..3) fused sweeps; in == out is allowed only for a single tile per equilibrium.
int sweep_fused_launch(const LevelGeom &g, const double *in, size_t istride, double *out, size_t ostride,
                       const double *src, size_t sstride, int batch, double omega, int sweeps, int par_off,
                       int num_sms, const int *active, cudaStream_t st) {
  if (g.nz < 3 || g.nr < 3 || sweeps <= 0 || batch <= 0) return GSB_OK;
  GSB_REQUIRE(sweeps <= 3, "sweep_fused_launch: at most 3 fused sweeps");
  const int nst = 2 * sweeps;
  SweepArgs a{};
  a.nz = g.nz;
  a.nr = g.nr;
  a.par_off = par_off;
  int ns, nb;
  int sc;
  sweep_fused_plan(g.nz, g.nr, batch, nst, num_sms, &sc, &a.band_rows, &ns, &nb);
  a.n_strips = ns;
  a.n_bands = nb;
  GSB_REQUIRE(in != out || (ns == 1 && nb == 1), "sweep_fused_launch: in-place needs a single tile per equilibrium");
  a.in = in;
  a.out = out;
  a.src = src;
  a.istride = istride;
  a.ostride = ostride;
  a.sstride = sstride;
  a.a_e = g.a_e;
  a.a_w = g.a_w;
  a.a_ns = g.a_ns;
  a.a_c = g.a_c;
  a.inv_a_c = g.inv_a_c;
  a.omega = omega;
  a.omw = 1.0 - omega;
  a.active = active;
  const size_t smem = sweep_smem_bytes(nst);
  const dim3 grd((ns * nb + kSwWPC - 1) / kSwWPC, 1, batch), blk(32 * kSwWPC, 1, 1);
  GSB_SMEM_OPT_IN(k_sweep_warp<2>, sweep_smem_bytes(2));
  GSB_SMEM_OPT_IN(k_sweep_warp<4>, sweep_smem_bytes(4));
  GSB_SMEM_OPT_IN(k_sweep_warp<6>, sweep_smem_bytes(6));
  if (nst == 2)
    k_sweep_warp<2><<<grd, blk, smem, st>>>(a);
  else if (nst == 4)
    k_sweep_warp<4><<<grd, blk, smem, st>>>(a);
  else
    k_sweep_warp<6><<<grd, blk, smem, st>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

}  // namespace gsb
