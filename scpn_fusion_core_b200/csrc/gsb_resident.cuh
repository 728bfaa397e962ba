// gsb_resident.cuh - shared-memory-resident multigrid: one CTA owns one equilibrium and runs a
// whole V-cycle (all levels) without touching HBM.
//
// Layout ("colour-split"): a level field x[nz][nr] is kept as two planes, one per red/black
// colour c = (iz+ir)&1, each [nz][hw] with hw = (nr+1)/2 and element (iz, k) <-> ir = 2k + s,
// s = ir&1.  With it every operand of a colour pass is unit-stride across the lanes of a warp
// (no bank conflicts on 64-bit accesses):
//   point (iz, k) of colour p, s = (p+iz)&1:   S,N = Q[iz-1][k], Q[iz+1][k]   (Q = other colour)
//                                              s=0: W = Q[iz][k-1], E = Q[iz][k]
//                                              s=1: W = Q[iz][k],   E = Q[iz][k+1]
// Every operator maps a thread to one k slot (one column pair) and a run of consecutive rows and
// slides a register window down the run: a smoother update costs 3 shared loads + 1 store (+ the
// right-hand side), residuals are evaluated once per fine row and column triple, prolongation
// loads two coarse values per coarse row.  Arithmetic is the same point functions as the
// streaming kernels (bit-identical results).
#pragma once

#include "gsb_internal.cuh"

namespace gsb {

constexpr int kResThreads = 512;
constexpr int kResMaxLevels = 12;
constexpr int kResSmemMax = 232448;  // 227 KB opt-in dynamic shared memory per CTA on sm_100

struct RLevel {
  int nz, nr, hw, nk;  // hw = (nr+1)/2 ; nk = nr/2 interior k slots
  int x_off;           // doubles offset of the solution planes [2][nz][hw] in the pool
  int d_off;           // doubles offset of the right-hand-side planes (levels after the first)
  int t_off;           // doubles offset of a shared-memory copy of a_e|a_w (2*nr) or -1
  // thread -> (k slot, row chunk) maps, precomputed on the host (no integer division on device)
  int lk;              // log2 of the padded slot count for k in [0, nk)
  int nch, rpc;        // row chunks over the nz-2 interior rows, rows per chunk
  int lj;              // log2 of the padded slot count for coarse columns J in [1, nrc-2] (restriction INTO this level)
  int cch, crpc;       // chunks over this level's interior rows when it is the coarse side
  LevelGeom g;
};
struct RPlan {
  int nlev;
  int pool_doubles;  // planes + tables; the staged level descriptors follow
  RLevel lev[kResMaxLevels];
};
static_assert(sizeof(RLevel) % 8 == 0, "RLevel is staged to shared memory in 8-byte words");
inline int res_stage_doubles(int nlev) { return nlev * (int)(sizeof(RLevel) / 8); }

// The one dynamic shared-memory array of the resident kernels.  Device functions index it through
// integer offsets (never through pointers handed across a call) so that every access compiles to
// LDS/STS: a generic-address load from shared memory costs ~150 cycles instead of ~30 (measured).
extern __shared__ double res_pool[];

__device__ __forceinline__ const RLevel &res_level(int lev_off, int l) {
  return reinterpret_cast<const RLevel *>(res_pool + lev_off)[l];
}

// Optional in-kernel phase timing (build with -DGSB_PHASE_TIMING): CTA 0 accumulates clock64()
// deltas per V-cycle phase into g_phase[]; read back with gsb_debug_phase_cycles().
#ifdef GSB_PHASE_TIMING
__device__ long long g_phase[64];
#define GSB_PHASE_BEGIN() long long _pt = clock64()
#define GSB_PHASE(idx)                                      \
  do {                                                      \
    if (blockIdx.x == 0 && threadIdx.x == 0) {              \
      const long long _now = clock64();                     \
      g_phase[(idx)] += _now - _pt;                         \
      _pt = _now;                                           \
    }                                                       \
  } while (0)
#define GSB_PHASE_COUNT(idx) g_phase[(idx)] += 1
#else
#define GSB_PHASE_BEGIN() do {} while (0)
#define GSB_PHASE(idx) do {} while (0)
#define GSB_PHASE_COUNT(idx) do {} while (0)
#endif

__device__ __forceinline__ int split_index(int nz, int hw, int iz, int ir) {
  return (((iz + ir) & 1) * nz + iz) * hw + (ir >> 1);
}

// dense global [nz][nr] -> planes (warp per row segment, no integer division)
__device__ __forceinline__ void res_load_dense(const double *__restrict__ g, int xo, int nz, int nr,
                                               int hw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int iz = warp; iz < nz; iz += nw) {
    const double *row = g + (size_t)iz * nr;
    for (int ir = lane; ir < nr; ir += 32) res_pool[xo + (((iz + ir) & 1) * nz + iz) * hw + (ir >> 1)] = row[ir];
  }
}
__device__ __forceinline__ void res_store_dense(double *__restrict__ g, int xo, int nz, int nr, int hw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int iz = warp; iz < nz; iz += nw) {
    double *row = g + (size_t)iz * nr;
    for (int ir = lane; ir < nr; ir += 32) row[ir] = res_pool[xo + (((iz + ir) & 1) * nz + iz) * hw + (ir >> 1)];
  }
}
// One row update of a colour pass.  S = ir&1 of the updated point (compile time): selects which of
// (c_cur, side) is the east / west neighbour and which coefficient pair applies.
template <int S>
__device__ __forceinline__ void res_row(const LevelGeom &g, int po, int qo, double c_prev, double c_cur,
                                        double c_next, double rhs, double ae, double aw, bool ok, double omega,
                                        double omw) {
  const double side = res_pool[qo + (S ? 1 : -1)];
  const double v = sor_point(g, ae, aw, S ? side : c_cur, S ? c_cur : side, c_prev, c_next, rhs, res_pool[po],
                             omega, omw);
  if (ok) res_pool[po] = v;
}

// Rows [z0, z1) of one k slot; S0 = ir&1 of the colour-`parity` point in row z0.
// po/qo: pool offsets of P[z0][k] / Q[z0][k]; rhs: global pointer to R[z0][k] or pool offset ro.
template <bool GLOBAL_RHS, int S0>
__device__ __forceinline__ void res_smooth_rows(const LevelGeom &g, int po, int qo,
                                                const double *__restrict__ r, int ro, int hw, int z0, int z1,
                                                double aeA, double awA, bool okA, double aeB, double awB,
                                                bool okB, double omega, double omw) {
  double c_prev = res_pool[qo - hw], c_cur = res_pool[qo];
  int iz = z0;
  if (GLOBAL_RHS) {
    double rn[4];
    if (iz + 4 <= z1) {
#pragma unroll
      for (int t = 0; t < 4; ++t) rn[t] = r[t * hw];
    }
    for (; iz + 4 <= z1; iz += 4) {
      double rc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) rc[t] = rn[t];
      if (iz + 8 <= z1) {
#pragma unroll
        for (int t = 0; t < 4; ++t) rn[t] = r[(4 + t) * hw];
      }
      const double c1 = res_pool[qo + hw], c2 = res_pool[qo + 2 * hw], c3 = res_pool[qo + 3 * hw],
                   c4 = res_pool[qo + 4 * hw];
      res_row<S0>(g, po, qo, c_prev, c_cur, c1, rc[0], aeA, awA, okA, omega, omw);
      res_row<1 - S0>(g, po + hw, qo + hw, c_cur, c1, c2, rc[1], aeB, awB, okB, omega, omw);
      res_row<S0>(g, po + 2 * hw, qo + 2 * hw, c1, c2, c3, rc[2], aeA, awA, okA, omega, omw);
      res_row<1 - S0>(g, po + 3 * hw, qo + 3 * hw, c2, c3, c4, rc[3], aeB, awB, okB, omega, omw);
      c_prev = c3;
      c_cur = c4;
      po += 4 * hw;
      qo += 4 * hw;
      r += 4 * hw;
    }
  } else {
    for (; iz + 4 <= z1; iz += 4) {
      const double c1 = res_pool[qo + hw], c2 = res_pool[qo + 2 * hw], c3 = res_pool[qo + 3 * hw],
                   c4 = res_pool[qo + 4 * hw];
      const double r0 = res_pool[ro], r1 = res_pool[ro + hw], r2 = res_pool[ro + 2 * hw], r3 = res_pool[ro + 3 * hw];
      res_row<S0>(g, po, qo, c_prev, c_cur, c1, r0, aeA, awA, okA, omega, omw);
      res_row<1 - S0>(g, po + hw, qo + hw, c_cur, c1, c2, r1, aeB, awB, okB, omega, omw);
      res_row<S0>(g, po + 2 * hw, qo + 2 * hw, c1, c2, c3, r2, aeA, awA, okA, omega, omw);
      res_row<1 - S0>(g, po + 3 * hw, qo + 3 * hw, c2, c3, c4, r3, aeB, awB, okB, omega, omw);
      c_prev = c3;
      c_cur = c4;
      po += 4 * hw;
      qo += 4 * hw;
      ro += 4 * hw;
    }
  }
  // tail (< 4 rows): alternate A, B
  bool a_row = true;
  for (; iz < z1; ++iz) {
    const double c1 = res_pool[qo + hw];
    const double rv = GLOBAL_RHS ? r[0] : res_pool[ro];
    if (a_row)
      res_row<S0>(g, po, qo, c_prev, c_cur, c1, rv, aeA, awA, okA, omega, omw);
    else
      res_row<1 - S0>(g, po, qo, c_prev, c_cur, c1, rv, aeB, awB, okB, omega, omw);
    a_row = !a_row;
    c_prev = c_cur;
    c_cur = c1;
    po += hw;
    qo += hw;
    ro += hw;
    r += hw;
  }
}

// One colour pass of RB-SOR on the resident planes of level l; rhs in the same split layout.
// GLOBAL_RHS: rhs (and the coefficient tables) live in global memory (the finest resident level)
// -> software prefetch one 4-row group ahead; otherwise both are in the pool.
template <bool GLOBAL_RHS>
__device__ __noinline__ void res_smooth_pass(int lev_off, int l, const double *__restrict__ rhs_g, int parity,
                                             double omega, double omw) {
  const RLevel &L = res_level(lev_off, l);
  const int nz = L.nz, nr = L.nr, hw = L.hw;
  const int P = L.x_off + parity * nz * hw;
  const int Q = L.x_off + (1 - parity) * nz * hw;
  const int Rp = parity * nz * hw;  // offset of the colour plane inside the rhs planes
  const int nslots = L.nch << L.lk;
  for (int w = threadIdx.x; w < nslots; w += blockDim.x) {
    const int k = w & ((1 << L.lk) - 1), ch = w >> L.lk;
    const int z0 = 1 + ch * L.rpc;
    const int z1 = min(z0 + L.rpc, nz - 1);
    if (k >= L.nk || z0 >= z1) continue;
    const int ir0 = 2 * k, ir1 = 2 * k + 1;
    double ae0, aw0, ae1, aw1;
    if (GLOBAL_RHS) {
      const double *tab = L.g.a_e;  // a_e | a_w, one allocation
      ae0 = tab[ir0];
      aw0 = tab[nr + ir0];
      ae1 = ir1 < nr ? tab[ir1] : 0.0;
      aw1 = ir1 < nr ? tab[nr + ir1] : 0.0;
    } else {
      const int t = L.t_off;
      ae0 = res_pool[t + ir0];
      aw0 = res_pool[t + nr + ir0];
      ae1 = ir1 < nr ? res_pool[t + ir1] : 0.0;
      aw1 = ir1 < nr ? res_pool[t + nr + ir1] : 0.0;
    }
    // k = 0 with S = 0 is the wall column: its `side` read at k-1 stays inside the pool (row >= 1)
    const bool ok0 = ir0 >= 1 && ir0 <= nr - 2, ok1 = ir1 <= nr - 2;
    const int o = z0 * hw + k;
    const double *rg = GLOBAL_RHS ? rhs_g + Rp + o : nullptr;
    const int ro = GLOBAL_RHS ? 0 : L.d_off + Rp + o;
    if (L.rpc == 1) {  // coarse levels: one row per thread, S varies across the warp -> no templated branch
      const int s = (parity + z0) & 1;
      const double cc = res_pool[Q + o], side = res_pool[Q + o - 1 + 2 * s];
      const double v = sor_point(L.g, s ? ae1 : ae0, s ? aw1 : aw0, s ? side : cc, s ? cc : side, res_pool[Q + o - hw],
                                 res_pool[Q + o + hw], GLOBAL_RHS ? rg[0] : res_pool[ro], res_pool[P + o], omega, omw);
      if (s ? ok1 : ok0) res_pool[P + o] = v;
      continue;
    }
    if ((parity + z0) & 1)
      res_smooth_rows<GLOBAL_RHS, 1>(L.g, P + o, Q + o, rg, ro, hw, z0, z1, ae1, aw1, ok1, ae0, aw0, ok0, omega, omw);
    else
      res_smooth_rows<GLOBAL_RHS, 0>(L.g, P + o, Q + o, rg, ro, hw, z0, z1, ae0, aw0, ok0, ae1, aw1, ok1, omega, omw);
  }
  __syncthreads();
}

// d_coarse = restrict_full_weight( -(L x - rhs) ) on the coarse interior (coarse wall unused).
// Thread <-> coarse column J and a chunk of coarse rows; marches over fine rows keeping the three
// residuals of each of the last fine rows (columns 2J-1, 2J, 2J+1) in registers.
template <bool GLOBAL_RHS>
__device__ __noinline__ void res_residual_restrict(int lev_off, int l, const double *__restrict__ rhs_g) {
  const RLevel &L = res_level(lev_off, l);
  const RLevel &C = res_level(lev_off, l + 1);
  const int nz = L.nz, hw = L.hw;
  const int xo = L.x_off;
  const int nslots = C.cch << C.lj;
  for (int w = threadIdx.x; w < nslots; w += blockDim.x) {
    const int J = 1 + (w & ((1 << C.lj) - 1)), ch = w >> C.lj;
    const int I0 = 1 + ch * C.crpc;
    const int I1 = min(I0 + C.crpc, C.nz - 1);
    if (J > C.nr - 2 || I0 >= I1) continue;
    // fine columns 2J-2 .. 2J+2 : even ones (k = J-1, J, J+1) live in plane (iz&1), odd ones
    // (k = J-1, J) in plane ((iz+1)&1)
    const int irm = 2 * J - 1, ir0 = 2 * J, irp = 2 * J + 1;
    const double *ts = L.g.r_safe, *ti = L.g.inv_r_safe;
    const double rsm = ts[irm], rs0 = ts[ir0], rsp = ts[irp], ism = ti[irm], is0 = ti[ir0], isp = ti[irp];
    auto load_row = [&](int iz, double *v) {  // v[0..4] = x(iz, 2J-2 .. 2J+2)
      const int pe = xo + ((iz & 1) * nz + iz) * hw + J;
      const int po = xo + (((iz + 1) & 1) * nz + iz) * hw + J;
      v[0] = res_pool[pe - 1];
      v[1] = res_pool[po - 1];
      v[2] = res_pool[pe];
      v[3] = res_pool[po];
      v[4] = res_pool[pe + 1];
    };
    auto rhs_row = [&](int iz, double *v) {  // v[0..2] = rhs(iz, 2J-1 .. 2J+1)
      const int pe = ((iz & 1) * nz + iz) * hw + J;
      const int po = (((iz + 1) & 1) * nz + iz) * hw + J;
      if (GLOBAL_RHS) {
        v[0] = rhs_g[po - 1];
        v[1] = rhs_g[pe];
        v[2] = rhs_g[po];
      } else {
        v[0] = res_pool[L.d_off + po - 1];
        v[1] = res_pool[L.d_off + pe];
        v[2] = res_pool[L.d_off + po];
      }
    };
    double xm[5], x0[5], xp[5], rA[3], rB[3], rC[3], f[3];
    int iz = 2 * I0 - 1;
    load_row(iz - 1, xm);
    load_row(iz, x0);
    for (int I = I0; I < I1; ++I) {
#pragma unroll
      for (int step = 0; step < 3; ++step) {
        // the first fine row of a coarse row was the last one of the previous coarse row
        if (step == 0 && I > I0) continue;
        load_row(iz + 1, xp);
        rhs_row(iz, f);
        double *rr = step == 0 ? rA : (step == 1 ? rB : rC);
        rr[0] = -dsub(gs_apply_v(L.g, rsm, ism, x0[1], x0[2], x0[0], xm[1], xp[1]), f[0]);
        rr[1] = -dsub(gs_apply_v(L.g, rs0, is0, x0[2], x0[3], x0[1], xm[2], xp[2]), f[1]);
        rr[2] = -dsub(gs_apply_v(L.g, rsp, isp, x0[3], x0[4], x0[2], xm[3], xp[3]), f[2]);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          xm[q] = x0[q];
          x0[q] = xp[q];
        }
        ++iz;
      }
      res_pool[C.d_off + split_index(C.nz, C.hw, I, J)] =
          fw9(rB[1], rA[1], rC[1], rB[0], rB[2], rA[0], rA[2], rC[0], rC[2]);
#pragma unroll
      for (int q = 0; q < 3; ++q) rA[q] = rC[q];
    }
  }
}

// x_fine += P e_coarse on the fine interior.  Thread <-> fine k slot (columns 2k, 2k+1 <-> coarse
// columns k, k+1) and a chunk of fine rows.
static __device__ __noinline__ void res_prolong_add(int lev_off, int l) {
  const RLevel &L = res_level(lev_off, l);
  const RLevel &C = res_level(lev_off, l + 1);
  const int nz = L.nz, nr = L.nr, hw = L.hw;
  const int nzc = C.nz, nrc = C.nr, hwc = C.hw;
  const int xo = L.x_off, eo = C.x_off;
  const int kz = min(nzc, (nz + 1) / 2), kr = min(nrc, (nr + 1) / 2);
  const int hend = min(2 * (nrc - 1), nr - 1), vend = min(2 * (nzc - 1), nz - 1);
  const int nslots = L.nch << L.lk;
  for (int w = threadIdx.x; w < nslots; w += blockDim.x) {
    const int k = w & ((1 << L.lk) - 1), ch = w >> L.lk;
    const int z0 = 1 + ch * L.rpc;
    const int z1 = min(z0 + L.rpc, nz - 1);
    if (k >= L.nk || z0 >= z1) continue;
    const int ire = 2 * k, iro = 2 * k + 1;
    const bool e_ok = ire >= 1 && ire <= nr - 2 && k < kr;
    const bool o_ok = iro <= nr - 2 && iro < hend;
    for (int iz = z0; iz < z1; ++iz) {
      const bool ze = (iz & 1) == 0;
      if (!(ze ? (iz / 2 < kz) : (iz < vend))) continue;
      const int I = iz >> 1;
      const double a = res_pool[eo + split_index(nzc, hwc, I, k)];
      const double b = (k + 1 < nrc) ? res_pool[eo + split_index(nzc, hwc, I, k + 1)] : 0.0;
      double ve, vo;
      if (ze) {
        ve = a;
        vo = dmul(0.5, dadd(a, b));
      } else {
        const double c = res_pool[eo + split_index(nzc, hwc, I + 1, k)];
        const double d = (k + 1 < nrc) ? res_pool[eo + split_index(nzc, hwc, I + 1, k + 1)] : 0.0;
        ve = dmul(0.5, dadd(a, c));
        vo = dmul(0.25, dadd(dadd(dadd(a, c), b), d));
      }
      if (e_ok) {
        const int p = xo + ((iz & 1) * nz + iz) * hw + k;  // even column: colour = iz&1
        res_pool[p] = dadd(res_pool[p], ve);
      }
      if (o_ok) {
        const int p = xo + (((iz + 1) & 1) * nz + iz) * hw + k;
        res_pool[p] = dadd(res_pool[p], vo);
      }
    }
  }
}

// Smoother column tables a_e | a_w (nr each): the coarse resident levels keep a copy in the pool so
// per-pass coefficient fetches cost a shared-memory load, not an L2 round trip.  The level
// descriptors themselves are copied to shared memory as well: with ~2.4 KB of kernel parameters the
// per-pass constant-bank reads miss the constant cache.  Returns the pool offset of the descriptors.
__device__ __forceinline__ int res_stage(const RPlan &plan) {
  const int lev_off = plan.pool_doubles;
  const int words = plan.nlev * (int)(sizeof(RLevel) / 8);
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&plan.lev[0]);
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(res_pool + lev_off);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  for (int l = 0; l < plan.nlev; ++l) {
    const RLevel &R = plan.lev[l];
    if (R.t_off < 0) continue;
    for (int i = threadIdx.x; i < 2 * R.nr; i += blockDim.x) res_pool[R.t_off + i] = R.g.a_e[i];
  }
  __syncthreads();
  return lev_off;
}

// Base level with at most 3x3 interior points and a zero wall: 50 sweeps by ONE warp, one lane per
// point, values in registers, neighbours through warp shuffles (no barriers, no shared-memory round
// trips).  Absent points of smaller grids behave as the zero wall.  Same point function.
static __device__ __noinline__ void res_base_small(int lev_off, int l, double omega, double omw, int sweeps) {
  if (threadIdx.x < 32) {
    const RLevel &B = res_level(lev_off, l);
    const int lane = threadIdx.x;
    const int nz = B.nz, nr = B.nr, hw = B.hw;
    const int i = lane / 3, j = lane - 3 * i;  // interior point (1+i, 1+j)
    const bool on = lane < 9 && (1 + i) <= nz - 2 && (1 + j) <= nr - 2;
    const double ae = on ? res_pool[B.t_off + 1 + j] : 0.0, aw = on ? res_pool[B.t_off + nr + 1 + j] : 0.0;
    const double f = on ? res_pool[B.d_off + split_index(nz, hw, 1 + i, 1 + j)] : 0.0;
    const int par = (i + j) & 1;
    const bool hasE = on && j < 2, hasW = on && j > 0, hasN = on && i < 2, hasS = on && i > 0;
    double v = 0.0;
    for (int s = 0; s < sweeps; ++s) {
#pragma unroll
      for (int parity = 0; parity < 2; ++parity) {
        const double e = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
        const double w = __shfl_sync(0xffffffffu, v, (lane + 31) & 31);
        const double n = __shfl_sync(0xffffffffu, v, (lane + 3) & 31);
        const double so = __shfl_sync(0xffffffffu, v, (lane + 29) & 31);
        const double nv = sor_point(B.g, ae, aw, hasE ? e : 0.0, hasW ? w : 0.0, hasS ? so : 0.0, hasN ? n : 0.0, f,
                                    v, omega, omw);
        if (on && par == parity) v = nv;
      }
    }
    if (on) res_pool[B.x_off + split_index(nz, hw, 1 + i, 1 + j)] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ void res_zero_off(int off, int count) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) res_pool[off + i] = 0.0;
}

// One V-cycle over the `nlev` resident levels staged at `lev_off` (multigrid_solve.py:252-335).
// The finest resident level's planes must already be in the pool; its right-hand side `rhs0` is in
// global memory (split layout).
__device__ __forceinline__ void res_vcycle(int lev_off, int nlev, const double *__restrict__ rhs0, double omega,
                                           int pre, int post) {
  const double omw = 1.0 - omega;
  const int L = nlev;
  GSB_PHASE_BEGIN();
  for (int l = 0; l < L - 1; ++l) {
    const RLevel &F = res_level(lev_off, l);
    const RLevel &C = res_level(lev_off, l + 1);
    if (F.nz > 2 && F.nr > 2)
      for (int s = 0; s < 2 * pre; ++s) {
        if (l == 0)
          res_smooth_pass<true>(lev_off, l, rhs0, s & 1, omega, omw);
        else
          res_smooth_pass<false>(lev_off, l, nullptr, s & 1, omega, omw);
      }
    GSB_PHASE(4 * l + 0);  // pre-smooth
    res_zero_off(C.x_off, 2 * C.nz * C.hw);
    if (C.nz > 2 && C.nr > 2) {
      if (l == 0)
        res_residual_restrict<true>(lev_off, l, rhs0);
      else
        res_residual_restrict<false>(lev_off, l, nullptr);
    }
    __syncthreads();
    GSB_PHASE(4 * l + 1);  // residual + restriction
  }
  {
    const RLevel &B = res_level(lev_off, L - 1);
    if (B.nz > 2 && B.nr > 2) {
      if (L > 1 && B.nz <= 5 && B.nr <= 5)
        res_base_small(lev_off, L - 1, omega, omw, 50);
      else
        for (int s = 0; s < 100; ++s) {
          if (L == 1)
            res_smooth_pass<true>(lev_off, 0, rhs0, s & 1, omega, omw);
          else
            res_smooth_pass<false>(lev_off, L - 1, nullptr, s & 1, omega, omw);
        }
    }
    GSB_PHASE(4 * (L - 1) + 0);  // base solve
  }
  for (int l = L - 2; l >= 0; --l) {
    const RLevel &F = res_level(lev_off, l);
    if (F.nz > 2 && F.nr > 2) {
      res_prolong_add(lev_off, l);
      __syncthreads();
      GSB_PHASE(4 * l + 2);  // prolongation
      for (int s = 0; s < 2 * post; ++s) {
        if (l == 0)
          res_smooth_pass<true>(lev_off, l, rhs0, s & 1, omega, omw);
        else
          res_smooth_pass<false>(lev_off, l, nullptr, s & 1, omega, omw);
      }
      GSB_PHASE(4 * l + 3);  // post-smooth
    }
  }
}

// host: resident plan for levels [l0, end) of ctx; returns false if it does not fit one SM
bool build_rplan(const gsb_ctx *ctx, int l0, int extra_doubles, RPlan *out);

}  // namespace gsb
