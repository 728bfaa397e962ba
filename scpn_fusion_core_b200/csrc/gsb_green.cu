// gsb_green.cu - Green's-function boundary pieces (a15, a16, a18).
//
//  * Cephes ellpk/ellpe in device FP64 (what scipy.special.ellipk/ellipe wrap)
//  * coil -> grid unit tables, batched coil-flux accumulation
//  * coil -> points mutual matrix
//  * lane-C plasma -> wall response matrix and its batched contraction on the FP64 tensor pipe
#include "gsb_internal.cuh"

namespace gsb {

__constant__ double c_ellpk_P[11] = {
    1.37982864606273237150e-4, 2.28025724005875567385e-3, 7.97404013220415179367e-3,
    9.85821379021226008714e-3, 6.87489687449949877925e-3, 6.18901033637687613229e-3,
    8.79078273952743772254e-3, 1.49380448916805252718e-2, 3.08851465246711995998e-2,
    9.65735902811690126535e-2, 1.38629436111989062502e0};
__constant__ double c_ellpk_Q[11] = {
    2.94078955048598507511e-5, 9.14184723865917226571e-4, 5.94058303753167793257e-3,
    1.54850516649762399335e-2, 2.39089602715924892727e-2, 3.01204715227604046988e-2,
    3.73774314173823228969e-2, 4.88280347570998239232e-2, 7.03124996963957469739e-2,
    1.24999999999870820058e-1, 4.99999999999999999821e-1};
__constant__ double c_ellpe_P[11] = {
    1.53552577301013293365e-4, 2.50888492163602060990e-3, 8.68786816565889628429e-3,
    1.07350949056076193403e-2, 7.77395492516787092951e-3, 7.58395289413514708519e-3,
    1.15688436810574127319e-2, 2.18317996015557253103e-2, 5.68051945617860553470e-2,
    4.43147180560990850618e-1, 1.00000000000000000299e0};
__constant__ double c_ellpe_Q[10] = {
    3.27954898576485872656e-5, 1.00962792679356715133e-3, 6.50609489976927491433e-3,
    1.68862163993311317300e-2, 2.61769742454493659583e-2, 3.34833904888224918614e-2,
    4.27180926518931511717e-2, 5.85936634471101055642e-2, 9.37499997197644278445e-2,
    2.49999999999888314361e-1};

template <int N>
__device__ __forceinline__ double polevl(double x, const double *c) {
  double a = c[0];
#pragma unroll
  for (int i = 1; i < N; ++i) a = dadd(dmul(a, x), c[i]);
  return a;
}
// Cephes ellpk(x), x = 1-m in (MACHEP, 1]:  P(x) - log(x) Q(x)
__device__ __forceinline__ double ellpk_x(double x) {
  return dsub(polevl<11>(x, c_ellpk_P), dmul(log(x), polevl<11>(x, c_ellpk_Q)));
}
// Cephes ellpe(m): x = 1-m;  P(x) - log(x) (x Q(x))
__device__ __forceinline__ double ellpe_x(double x) {
  return dsub(polevl<11>(x, c_ellpe_P), dmul(log(x), dmul(x, polevl<10>(x, c_ellpe_Q))));
}

__device__ __forceinline__ double clipd(double v, double lo, double hi) {
  return fmin(fmax(v, lo), hi);
}

// fusion_kernel.py:236-249: sqrt(R Rc) and ((2-k2)K - 2E)/k with k2 clipped to [1e-12, 1-1e-12]
__device__ __forceinline__ void green_lane_a(double R, double Z, double Rc, double Zc, double &sq,
                                             double &term) {
  const double dZ = dsub(Z, Zc);
  const double rp = dadd(R, Rc);
  const double den = dadd(dmul(rp, rp), dmul(dZ, dZ));
  double k2 = __ddiv_rn(dmul(dmul(4.0, R), Rc), den);
  k2 = clipd(k2, 1e-12, 1.0 - 1e-12);
  const double x = dsub(1.0, k2);
  const double K = ellpk_x(x), E = ellpe_x(x);
  sq = __dsqrt_rn(dmul(R, Rc));
  const double k = __dsqrt_rn(k2);
  term = __ddiv_rn(dsub(dmul(dsub(2.0, k2), K), dmul(2.0, E)), k);
}

// fusion_kernel_free_boundary.py:58-80 (SI, self point -> 0)
__device__ __forceinline__ double green_si(double Ro, double Zo, double Rs, double Zs) {
  const double dr = dsub(Ro, Rs), dz = dsub(Zo, Zs);
  const bool self = dadd(dmul(dr, dr), dmul(dz, dz)) < 1e-24;
  const double rp = dadd(Ro, Rs);
  const double den = dadd(dmul(rp, rp), dmul(dz, dz));
  double k2 = den > 1e-30 ? __ddiv_rn(dmul(dmul(4.0, Ro), Rs), fmax(den, 1e-30)) : 0.0;
  k2 = clipd(k2, 1e-12, 1.0 - 1e-12);
  const double k = __dsqrt_rn(k2);
  const double x = dsub(1.0, k2);
  const double K = ellpk_x(x), E = ellpe_x(x);
  // _MU0 / (2 pi) * sqrt(R_obs R_src), _MU0 = 4e-7*pi
  const double mu0 = 4e-7 * 3.141592653589793;
  const double pref = dmul(__ddiv_rn(mu0, 2.0 * 3.141592653589793), __dsqrt_rn(dmul(Ro, Rs)));
  const double flux = __ddiv_rn(dmul(pref, dsub(dmul(dsub(2.0, k2), K), dmul(2.0, E))), k);
  return self ? 0.0 : flux;
}

// jax_free_boundary_gs.py:70-86 with jax_equilibrium_solver.py:50-123 (lane C)
__device__ __forceinline__ double green_lane_c(double R, double Z, double Rc, double Zc, double pref) {
  const double Rs = fmax(R, 1e-6);
  const double rp = dadd(Rs, Rc), dz = dsub(Z, Zc);
  const double den = dadd(dmul(rp, rp), dmul(dz, dz));
  const double k2 = clipd(__ddiv_rn(dmul(dmul(4.0, Rs), Rc), fmax(den, 1e-30)), 1e-9, 0.999999);
  const double k = __dsqrt_rn(k2);
  const double x = clipd(dsub(1.0, k2), 1.0e-16, 1.0);
  const double lx = log(x);
  const double K = dsub(polevl<11>(x, c_ellpk_P), dmul(lx, polevl<11>(x, c_ellpk_Q)));
  const double E = dsub(polevl<11>(x, c_ellpe_P), dmul(dmul(x, lx), polevl<10>(x, c_ellpe_Q)));
  const double two_k = __ddiv_rn(2.0, k);
  const double v = dmul(dmul(pref, __dsqrt_rn(dmul(Rs, Rc))),
                        dsub(dmul(dsub(two_k, k), K), dmul(two_k, E)));
  return isfinite(v) ? v : 0.0;
}

__global__ void __launch_bounds__(256)
k_green_table(const double *__restrict__ rrow, const double *__restrict__ zax, int nz, int nr,
              const double *__restrict__ coil_rz, int si, double *__restrict__ g) {
  const int c = blockIdx.y;
  const size_t n = (size_t)nz * nr;
  const double Rc = coil_rz[2 * c], Zc = coil_rz[2 * c + 1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int iz = (int)(i / nr), ir = (int)(i - (size_t)iz * nr);
    if (si) {
      g[(size_t)c * n + i] = green_si(rrow[ir], zax[iz], Rc, Zc);
    } else {
      double sq, term;
      green_lane_a(rrow[ir], zax[iz], Rc, Zc, sq, term);
      g[((size_t)c * 2) * n + i] = sq;
      g[((size_t)c * 2 + 1) * n + i] = term;
    }
  }
}

// psi[b] = sum_c ... accumulated in coil order starting from 0.0 (Psi_vac += ...)
__global__ void __launch_bounds__(256)
k_coil_flux(const double *__restrict__ g, const double *__restrict__ w, int n_coils, size_t n, int si,
            double *__restrict__ psi) {
  const int b = blockIdx.y;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < n_coils; ++c) {
      const double wc = w[(size_t)b * n_coils + c];
      if (si)
        acc = dadd(acc, dmul(wc, g[(size_t)c * n + i]));
      else
        acc = dadd(acc, dmul(dmul(wc, g[((size_t)c * 2) * n + i]), g[((size_t)c * 2 + 1) * n + i]));
    }
    psi[(size_t)b * n + i] = acc;
  }
}

__global__ void k_mutual(const double *__restrict__ coil_rz, const int *__restrict__ turns,
                         const double *__restrict__ obs, int n_pts, double *__restrict__ m) {
  const int c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const double gval = green_si(obs[2 * i], obs[2 * i + 1], coil_rz[2 * c], coil_rz[2 * c + 1]);
  m[(size_t)c * n_pts + i] = dmul((double)turns[c], gval);
}

// wall ring / interior enumerations in C order (jax_free_boundary_predictive.py:166-180)
__device__ __forceinline__ void wall_coord(int w, int nz, int nr, int &iz, int &ir) {
  if (w < nr) {
    iz = 0;
    ir = w;
  } else if (w >= nr + 2 * (nz - 2)) {
    iz = nz - 1;
    ir = w - (nr + 2 * (nz - 2));
  } else {
    const int t = w - nr;
    iz = 1 + t / 2;
    ir = (t & 1) ? nr - 1 : 0;
  }
}

__global__ void __launch_bounds__(256)
k_wall_matrix(const double *__restrict__ rrow, const double *__restrict__ zax, int nz, int nr,
              double pref, double *__restrict__ m) {
  const int nint = (nz - 2) * (nr - 2);
  const int w = blockIdx.y;
  int wz, wr;
  wall_coord(w, nz, nr, wz, wr);
  const double Rw = rrow[wr], Zw = zax[wz];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nint; s += gridDim.x * blockDim.x) {
    const int iz = 1 + s / (nr - 2), ir = 1 + s % (nr - 2);
    m[(size_t)w * nint + s] = green_lane_c(Rw, Zw, rrow[ir], zax[iz], pref);
  }
}

// ------------------------------------------------------------------------------------------
// a18  wall[b][w] = sum_s M[w][s] * (J[b][interior s] * dA)
// C[B x Nw] = X[B x K] * M^T[K x Nw] on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA).
// CTA tile 64(b) x 64(w), K step 16, 256 threads = 8 warps laid out 2(b) x 4(w); each warp owns a
// 32 x 16 sub-tile = 4 x 2 m8n8 accumulators.  The interior gather and the *dA scaling are fused
// into the X tile load.  Split-K over gridDim.z with a deterministic second-pass reduction.
// ------------------------------------------------------------------------------------------
constexpr int GB = 64, GW = 64, GK = 16;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
k_wall_gemm(const double *__restrict__ M, const double *__restrict__ J, int nz, int nr, int batch,
            int nwall, int nint, double dA, int k_per_split, double *__restrict__ out /*[split][B][Nw]*/) {
  __shared__ double sX[GK][GB + 4];  // [k][b]
  __shared__ double sM[GK][GW + 4];  // [k][w]
  const int b0 = blockIdx.x * GB, w0 = blockIdx.y * GW;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(nint, kbeg + k_per_split);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wb = (warp >> 2) * 32, ww = (warp & 3) * 16;
  const int gid = lane >> 2, tig = lane & 3;
  const size_t n = (size_t)nz * nr;
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    // X tile: 64 b x 16 k ; thread -> (b = tid/4 , 4 consecutive k)
    {
      const int bb = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int s = k0 + kk + q;
        double v = 0.0;
        if (s < kend && b0 + bb < batch) {
          const int iz = 1 + s / (nr - 2), ir = 1 + s % (nr - 2);
          v = J[(size_t)(b0 + bb) * n + (size_t)iz * nr + ir] * dA;
        }
        sX[kk + q][bb] = v;
      }
    }
    // M tile: 64 w x 16 k ; M is [w][s] row-major so k is contiguous
    {
      const int wwl = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int s = k0 + kk + q;
        double v = 0.0;
        if (s < kend && w0 + wwl < nwall) v = M[(size_t)(w0 + wwl) * nint + s];
        sM[kk + q][wwl] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < GK; ks += 4) {
      double a[4], bfr[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sX[ks + tig][wb + i * 8 + gid];    // A[row=gid][k=tig]
#pragma unroll
      for (int j = 0; j < 2; ++j) bfr[j] = sM[ks + tig][ww + j * 8 + gid];  // B[k=tig][col=gid]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bfr[j]);
    }
    __syncthreads();
  }
  // C fragment: row = gid, cols = 2*tig, 2*tig+1
  double *o = out + (size_t)blockIdx.z * batch * nwall;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = b0 + wb + i * 8 + gid;
      const int w = w0 + ww + j * 8 + 2 * tig;
      if (b < batch) {
        if (w < nwall) o[(size_t)b * nwall + w] = acc[i][j][0];
        if (w + 1 < nwall) o[(size_t)b * nwall + w + 1] = acc[i][j][1];
      }
    }
}

// Large-batch variant: CTA tile 128(b) x 64(w), K step 16, 8 warps laid out 4(b) x 2(w), each warp a
// 32 x 32 sub-tile = 4 x 4 m8n8k4 accumulators (two shared loads per DMMA instead of three quarters
// of one per DMMA... 8 LDS.64 feed 16 DMMA).  The K axis is walked as (interior row, 16-column
// chunk) so a K tile never straddles a grid row: no integer division in the gather, and the
// 16-column chunks of a row are contiguous in J and in M.  The next tile is fetched into registers
// while the current one is multiplied (one barrier pair per K step, global latency hidden).
constexpr int HB = 128, HW = 64, HK = 16;
__global__ void __launch_bounds__(256, 2)
k_wall_gemm_big(const double *__restrict__ M, const double *__restrict__ J, int nz, int nr, int batch, int nwall,
                int nint, double dA, int tiles_per_split, double *__restrict__ out /*[split][B][Nw]*/) {
  extern __shared__ double gsm[];  // two stages of { sX[HK][HB+4] ([k][b]), sM[HK][HW+4] ([k][w]) }
  constexpr int LDX = HB + 4, LDM = HW + 4, STAGE = HK * (LDX + LDM);
  const int b0 = blockIdx.x * HB, w0 = blockIdx.y * HW;
  const int ncol = nr - 2;                      // interior columns per grid row
  const int cpr = (ncol + HK - 1) / HK;         // 16-column chunks per row
  const int ntiles = (nz - 2) * cpr;
  const int t_beg = blockIdx.z * tiles_per_split, t_end = min(ntiles, t_beg + tiles_per_split);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wb = (warp >> 1) * 32, ww = (warp & 1) * 32;
  const int gid = lane >> 2, tig = lane & 3;
  const size_t n = (size_t)nz * nr;
  // loader roles: X tile 128 b x 16 k -> thread (b = tid/2, 8 consecutive k); M tile 64 w x 16 k -> (w = tid/4, 4 k)
  const int xb = tid >> 1, xk = (tid & 1) * 8;
  const int mw = tid >> 2, mk = (tid & 3) * 4;
  const bool xb_ok = b0 + xb < batch, mw_ok = w0 + mw < nwall;
  const double *jrow = J + (size_t)min(b0 + xb, batch - 1) * n;
  const double *mrow = M + (size_t)min(w0 + mw, nwall - 1) * nint;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double px[8], pm[4];
  auto fetch = [&](int t) {
    const int row = t / cpr, c0 = (t - row * cpr) * HK;  // once per K tile, not per element
    const double *jp = jrow + (size_t)(row + 1) * nr + 1 + c0 + xk;
    const double *mp = mrow + (size_t)row * ncol + c0 + mk;
#pragma unroll
    for (int q = 0; q < 8; ++q) px[q] = (xb_ok && c0 + xk + q < ncol) ? jp[q] * dA : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) pm[q] = (mw_ok && c0 + mk + q < ncol) ? mp[q] : 0.0;
  };
  auto stash = [&](int stage) {
    double *sX = gsm + stage * STAGE, *sM = sX + HK * LDX;
#pragma unroll
    for (int q = 0; q < 8; ++q) sX[(xk + q) * LDX + xb] = px[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) sM[(mk + q) * LDM + mw] = pm[q];
  };
  if (t_beg < t_end) {
    fetch(t_beg);
    stash(0);
  }
  __syncthreads();
  int cur = 0;
  for (int t = t_beg; t < t_end; ++t) {
    if (t + 1 < t_end) fetch(t + 1);  // global loads of the next tile fly during the multiply
    const double *sX = gsm + cur * STAGE, *sM = sX + HK * LDX;
#pragma unroll
    for (int ks = 0; ks < HK; ks += 4) {
      double a[4], bfr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sX[(ks + tig) * LDX + wb + i * 8 + gid];    // A[row=gid][k=tig]
#pragma unroll
      for (int j = 0; j < 4; ++j) bfr[j] = sM[(ks + tig) * LDM + ww + j * 8 + gid];  // B[k=tig][col=gid]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bfr[j]);
    }
    if (t + 1 < t_end) stash(1 - cur);  // the other stage was last read before the previous barrier
    __syncthreads();
    cur ^= 1;
  }
  double *o = out + (size_t)blockIdx.z * batch * nwall;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + wb + i * 8 + gid;
      const int w = w0 + ww + j * 8 + 2 * tig;
      if (b < batch) {
        if (w < nwall) o[(size_t)b * nwall + w] = acc[i][j][0];
        if (w + 1 < nwall) o[(size_t)b * nwall + w + 1] = acc[i][j][1];
      }
    }
}

__global__ void k_splitk_reduce(const double *__restrict__ part, int splits, size_t total,
                                double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double a = 0.0;
  for (int s = 0; s < splits; ++s) a += part[(size_t)s * total + i];
  out[i] = a;
}

__global__ void k_wall_scatter(const double *__restrict__ wall, double *__restrict__ bc, int nz, int nr,
                               int nwall, int accumulate) {
  const int b = blockIdx.y;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwall) return;
  int iz, ir;
  wall_coord(w, nz, nr, iz, ir);
  double *p = bc + (size_t)b * nz * nr + (size_t)iz * nr + ir;
  const double v = wall[(size_t)b * nwall + w];
  p[0] = accumulate ? p[0] + v : v;
}

}  // namespace gsb

using namespace gsb;

extern "C" {

int gsb_green_table(gsb_ctx *ctx, const double *coil_rz, int n_coils, int si, double *g_dev, void *stream) {
  GSB_REQUIRE(ctx && coil_rz && g_dev, "gsb_green_table: NULL argument");
  GSB_REQUIRE(n_coils >= 1 && n_coils <= 65535, "gsb_green_table: bad coil count");
  GSB_REQUIRE((int)ctx->z_axis.size() == ctx->nz, "gsb_green_table: context was created without a z axis");
  GSB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  double *crz = nullptr;
  GSB_CUDA(cudaMalloc(&crz, 2 * n_coils * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(crz, coil_rz, 2 * n_coils * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    const int blocks = (int)std::min<size_t>((ctx->n + 255) / 256, 1024);
    k_green_table<<<dim3(blocks, n_coils), 256, 0, st>>>(ctx->r_dev, ctx->z_dev, ctx->nz, ctx->nr, crz, si, g_dev);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(crz);
  if (e != cudaSuccess) {
    set_error(std::string("gsb_green_table: ") + cudaGetErrorString(e));
    return GSB_ECUDA;
  }
  return GSB_OK;
}

int gsb_coil_flux(gsb_ctx *ctx, const double *g_dev, const double *w_dev, int n_coils, int si,
                  double *psi_dev, int batch, void *stream) {
  GSB_REQUIRE(ctx && g_dev && w_dev && psi_dev, "gsb_coil_flux: NULL argument");
  GSB_REQUIRE(batch >= 1 && batch <= 65535 && n_coils >= 1, "gsb_coil_flux: bad sizes");
  GSB_CUDA(cudaSetDevice(ctx->device));
  const int blocks = (int)std::min<size_t>((ctx->n + 255) / 256, 256);
  k_coil_flux<<<dim3(blocks, batch), 256, 0, (cudaStream_t)stream>>>(g_dev, w_dev, n_coils, ctx->n, si, psi_dev);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

int gsb_mutual_matrix(const double *coil_rz, const int *turns, int n_coils, const double *obs_rz, int n_pts,
                      double *m_dev, void *stream) {
  GSB_REQUIRE(coil_rz && turns && obs_rz && m_dev, "gsb_mutual_matrix: NULL argument");
  GSB_REQUIRE(n_coils >= 1 && n_coils <= 65535 && n_pts >= 1, "gsb_mutual_matrix: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  double *crz = nullptr, *obs = nullptr;
  int *tr = nullptr;
  cudaError_t e = cudaMalloc(&crz, 2 * n_coils * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&obs, 2 * (size_t)n_pts * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&tr, n_coils * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(crz, coil_rz, 2 * n_coils * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(obs, obs_rz, 2 * (size_t)n_pts * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tr, turns, n_coils * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    k_mutual<<<dim3((n_pts + 127) / 128, n_coils), 128, 0, st>>>(crz, tr, obs, n_pts, m_dev);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(crz);
  cudaFree(obs);
  cudaFree(tr);
  if (e != cudaSuccess) {
    set_error(std::string("gsb_mutual_matrix: ") + cudaGetErrorString(e));
    return GSB_ECUDA;
  }
  return GSB_OK;
}

int gsb_wall_matrix(gsb_ctx *ctx, double mu0, double *m_dev, void *stream) {
  GSB_REQUIRE(ctx && m_dev, "gsb_wall_matrix: NULL argument");
  GSB_REQUIRE((int)ctx->z_axis.size() == ctx->nz, "gsb_wall_matrix: context was created without a z axis");
  GSB_REQUIRE(ctx->nz >= 3 && ctx->nr >= 3, "gsb_wall_matrix: grid has no interior");
  GSB_CUDA(cudaSetDevice(ctx->device));
  volatile double pref = mu0 * 1.0;
  volatile double pref2 = pref / (2.0 * 3.141592653589793);
  const int blocks = std::min((ctx->n_int + 255) / 256, 64);
  k_wall_matrix<<<dim3(blocks, ctx->n_wall), 256, 0, (cudaStream_t)stream>>>(ctx->r_dev, ctx->z_dev, ctx->nz, ctx->nr, pref2, m_dev);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

int gsb_wall_flux(gsb_ctx *ctx, const double *m_dev, const double *jphi_dev, double dA, double *wall_dev,
                  int batch, void *stream) {
  GSB_REQUIRE(ctx && m_dev && jphi_dev && wall_dev, "gsb_wall_flux: NULL argument");
  GSB_REQUIRE(batch >= 1, "gsb_wall_flux: bad batch");
  GSB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nw = ctx->n_wall, ni = ctx->n_int;
  if (batch >= HB) {  // large batches: 128 x 64 tiles, register double buffering
    const int cpr = (ctx->nr - 2 + HK - 1) / HK, ntiles = (ctx->nz - 2) * cpr;
    const int ctas = ((batch + HB - 1) / HB) * ((nw + HW - 1) / HW);
    int splits = 1;
    while (ctas * splits < 2 * ctx->num_sms && splits < 16 && ntiles / (splits * 2) >= 64) splits *= 2;
    const int tps = (ntiles + splits - 1) / splits;
    splits = (ntiles + tps - 1) / tps;
    double *part = wall_dev;
    if (splits > 1) GSB_CUDA(cudaMallocAsync(&part, (size_t)splits * batch * nw * sizeof(double), st));
    constexpr int kBigSmem = 2 * HK * (HB + 4 + HW + 4) * (int)sizeof(double);
    GSB_SMEM_OPT_IN(k_wall_gemm_big, kBigSmem);
    k_wall_gemm_big<<<dim3((batch + HB - 1) / HB, (nw + HW - 1) / HW, splits), 256, kBigSmem, st>>>(
        m_dev, jphi_dev, ctx->nz, ctx->nr, batch, nw, ni, dA, tps, part);
    GSB_LAUNCH_CHECK();
    if (splits > 1) {
      const size_t total = (size_t)batch * nw;
      k_splitk_reduce<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, splits, total, wall_dev);
      GSB_LAUNCH_CHECK();
      GSB_CUDA(cudaFreeAsync(part, st));
    }
    return GSB_OK;
  }
  const int tiles = ((batch + GB - 1) / GB) * ((nw + GW - 1) / GW);
  int splits = 1;
  while (tiles * splits < 148 * 2 && splits < 32 && ni / (splits * 2) >= 512) splits *= 2;
  int kps = (ni + splits - 1) / splits;
  kps = ((kps + GK - 1) / GK) * GK;
  splits = (ni + kps - 1) / kps;
  double *part = wall_dev;
  if (splits > 1) GSB_CUDA(cudaMallocAsync(&part, (size_t)splits * batch * nw * sizeof(double), st));
  k_wall_gemm<<<dim3((batch + GB - 1) / GB, (nw + GW - 1) / GW, splits), 256, 0, st>>>(
      m_dev, jphi_dev, ctx->nz, ctx->nr, batch, nw, ni, dA, kps, part);
  GSB_LAUNCH_CHECK();
  if (splits > 1) {
    const size_t total = (size_t)batch * nw;
    k_splitk_reduce<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, splits, total, wall_dev);
    GSB_LAUNCH_CHECK();
    GSB_CUDA(cudaFreeAsync(part, st));
  }
  return GSB_OK;
}

int gsb_wall_scatter(gsb_ctx *ctx, const double *wall_dev, double *bc_dev, int accumulate, int batch, void *stream) {
  GSB_REQUIRE(ctx && wall_dev && bc_dev, "gsb_wall_scatter: NULL argument");
  GSB_REQUIRE(batch >= 1 && batch <= 65535, "gsb_wall_scatter: bad batch");
  GSB_CUDA(cudaSetDevice(ctx->device));
  k_wall_scatter<<<dim3((ctx->n_wall + 127) / 128, batch), 128, 0, (cudaStream_t)stream>>>(
      wall_dev, bc_dev, ctx->nz, ctx->nr, ctx->n_wall, accumulate);
  GSB_LAUNCH_CHECK();
  return GSB_OK;
}

}  // extern "C"
