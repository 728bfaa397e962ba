// gsb_hpc.cu - the reference's native solver ABI on the GPU.
//
// Same six symbols, argument meaning and error behaviour as src/scpn_fusion/hpc/solver.cpp:200-335
// (HOST buffers in/out, psi persists in the handle, constant Dirichlet wall, omega 1.8 in run_step).
// Arithmetic follows solver.cpp:164-188 (update_point) operand order:
//   p_gs = (source + c_z*(up+down) + c_r+*right + c_r-*left) / center,  source = -1.0*R*j
// There is no CPU fallback: without a CUDA device create_solver() returns NULL, which the
// reference's HPCBridge already treats as "native solver unavailable" (hpc_bridge.py:252-285).
#include "gsb_internal.cuh"

#include <algorithm>

namespace gsb {

struct HpcSolver {
  int nr, nz;
  double rmin, rmax, zmin, zmax, dr, dz;
  double c_z, center, inv_center;
  double boundary = 0.0;
  double *psi = nullptr, *j = nullptr;
  double *tab = nullptr;  // negR | c_plus | c_minus  (nr each)
  unsigned long long *delta = nullptr;  // [0] running max bits, [1] last delta bits
  int *flags = nullptr;                 // [0] done, [1] performed
  cudaStream_t st = nullptr;
  double *h_pin = nullptr;  // pinned staging, 2*n doubles
};

__global__ void k_hpc_boundary(double *__restrict__ psi, int nz, int nr, double v, const int *__restrict__ done) {
  if (done && done[0]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nr) {
    psi[i] = v;
    psi[(size_t)(nz - 1) * nr + i] = v;
  }
  if (i < nz) {
    psi[(size_t)i * nr] = v;
    psi[(size_t)i * nr + nr - 1] = v;
  }
}

__global__ void __launch_bounds__(256)
k_hpc_colour(double *__restrict__ psi, const double *__restrict__ j, const double *__restrict__ tab, int nz,
             int nr, double c_z, double center, double inv_center, double omega, double omw, int parity,
             unsigned long long *__restrict__ delta, const int *__restrict__ done) {
  __shared__ double sh[32];
  if (done && done[0]) return;
  const int iz = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int j0 = (((iz + 1) & 1) == parity) ? 1 : 2;
  const int ir = j0 + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  double d = 0.0;
  if (iz < nz - 1 && ir < nr - 1) {
    double *p = psi + (size_t)iz * nr + ir;
    const double source = dmul(tab[ir], j[(size_t)iz * nr + ir]);  // (-1.0*R)*j
    double acc = dadd(source, dmul(c_z, dadd(p[nr], p[-nr])));
    acc = dadd(acc, dmul(tab[nr + ir], p[1]));
    acc = dadd(acc, dmul(tab[2 * nr + ir], p[-1]));
    const double gs = ddiv_y(acc, center, inv_center);
    const double old = p[0];
    const double nw = dadd(dmul(omw, old), dmul(omega, gs));
    p[0] = nw;
    d = fabs(dsub(nw, old));
  }
  if (delta) {
    // block max then one atomic per CTA; max is order independent -> deterministic
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    double m = warp_max(d);
    if ((t & 31) == 0) sh[t >> 5] = m;
    __syncthreads();
    if (t < 32) {
      m = t < (int)((blockDim.x * blockDim.y + 31) >> 5) ? sh[t] : 0.0;
      m = warp_max(m);
      if (t == 0 && m > 0.0) atomicMax(delta, (unsigned long long)__double_as_longlong(m));
    }
  }
}

__global__ void k_hpc_check(unsigned long long *delta, int *flags, double tol, int sweep_index) {
  if (flags[0]) return;
  const double d = __longlong_as_double((long long)delta[0]);
  delta[1] = delta[0];
  delta[0] = 0ull;
  flags[1] = sweep_index + 1;
  if (d <= tol) flags[0] = 1;
}

static void hpc_sweep(HpcSolver *s, double omega, bool track, int sweep_index, double tol) {
  const int half = (s->nr - 2 + 1) / 2;
  const dim3 blk(32, 8, 1);
  const dim3 grd((std::max(half, 1) + 31) / 32, (std::max(s->nz - 2, 1) + 7) / 8, 1);
  const int nb = (std::max(s->nr, s->nz) + 255) / 256;
  const int *done = track ? s->flags : nullptr;
  k_hpc_boundary<<<nb, 256, 0, s->st>>>(s->psi, s->nz, s->nr, s->boundary, done);
  for (int parity = 0; parity < 2; ++parity)
    k_hpc_colour<<<grd, blk, 0, s->st>>>(s->psi, s->j, s->tab, s->nz, s->nr, s->c_z, s->center, s->inv_center, omega,
                                         1.0 - omega, parity, track ? s->delta : nullptr, done);
  g_launches.fetch_add(3);
  if (track) {
    k_hpc_check<<<1, 1, 0, s->st>>>(s->delta, s->flags, tol, sweep_index);
    g_launches.fetch_add(1);
  }
}

static void hpc_free(HpcSolver *s) {
  if (!s) return;
  if (s->psi) cudaFree(s->psi);
  if (s->j) cudaFree(s->j);
  if (s->tab) cudaFree(s->tab);
  if (s->delta) cudaFree(s->delta);
  if (s->flags) cudaFree(s->flags);
  if (s->h_pin) cudaFreeHost(s->h_pin);
  if (s->st) cudaStreamDestroy(s->st);
  delete s;
}

}  // namespace gsb

using namespace gsb;

extern "C" {

void *create_solver(int nr, int nz, double rmin, double rmax, double zmin, double zmax) {
  if (nr < 2 || nz < 2) return nullptr;
  if (!(rmin < rmax) || !(zmin < zmax)) return nullptr;
  if (gsb_device_count() <= 0) {
    set_error("create_solver: no CUDA device (libgsb200 has no CPU fallback)");
    return nullptr;
  }
  HpcSolver *s = new HpcSolver();
  s->nr = nr;
  s->nz = nz;
  s->rmin = rmin;
  s->rmax = rmax;
  s->zmin = zmin;
  s->zmax = zmax;
  volatile double dr = (rmax - rmin) / (nr - 1), dz = (zmax - zmin) / (nz - 1);
  volatile double dr_sq = dr * dr, dz_sq = dz * dz;
  s->dr = dr;
  s->dz = dz;
  s->c_z = 1.0 / dz_sq;
  volatile double t1 = 2.0 / dr_sq, t2 = 2.0 / dz_sq;
  s->center = t1 + t2;
  s->inv_center = 1.0 / s->center;
  const size_t n = (size_t)nr * nz;
  std::vector<double> tab(3 * (size_t)nr);
  for (int r = 0; r < nr; ++r) {
    volatile double rd = r * dr;
    volatile double R = rmin + rd;
    volatile double inv = 1.0 / dr_sq;
    volatile double den = 2.0 * R;
    volatile double den2 = den * dr;
    volatile double q = 1.0 / den2;
    tab[r] = -1.0 * R;
    tab[nr + r] = inv - q;
    tab[2 * (size_t)nr + r] = inv + q;
  }
  cudaError_t e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&s->psi, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->j, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->tab, tab.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->delta, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&s->flags, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMallocHost(&s->h_pin, 2 * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(s->psi, 0, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(s->j, 0, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(s->tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error(std::string("create_solver: ") + cudaGetErrorString(e));
    hpc_free(s);
    return nullptr;
  }
  return s;
}

void set_boundary_dirichlet(void *solver_ptr, double boundary_value) {
  if (!solver_ptr) return;
  HpcSolver *s = static_cast<HpcSolver *>(solver_ptr);
  s->boundary = boundary_value;
  const int nb = (std::max(s->nr, s->nz) + 255) / 256;
  k_hpc_boundary<<<nb, 256, 0, s->st>>>(s->psi, s->nz, s->nr, s->boundary, nullptr);
  g_launches.fetch_add(1);
  cudaStreamSynchronize(s->st);
}

static bool hpc_upload(HpcSolver *s, const double *j_array, size_t n) {
  memcpy(s->h_pin, j_array, n * sizeof(double));
  return cudaMemcpyAsync(s->j, s->h_pin, n * sizeof(double), cudaMemcpyHostToDevice, s->st) == cudaSuccess;
}
static bool hpc_download(HpcSolver *s, double *psi_array, size_t n) {
  if (cudaMemcpyAsync(s->h_pin + n, s->psi, n * sizeof(double), cudaMemcpyDeviceToHost, s->st) != cudaSuccess) return false;
  if (cudaStreamSynchronize(s->st) != cudaSuccess) return false;
  memcpy(psi_array, s->h_pin + n, n * sizeof(double));
  return true;
}

void run_step(void *solver_ptr, const double *j_array, double *psi_array, int size, int iterations) {
  if (!solver_ptr || !j_array || !psi_array || size <= 0) return;
  HpcSolver *s = static_cast<HpcSolver *>(solver_ptr);
  const size_t n = (size_t)size;
  if (n != (size_t)s->nr * s->nz) return;
  if (!hpc_upload(s, j_array, n)) return;
  const int n_iter = std::max(iterations, 1);
  for (int i = 0; i < n_iter; ++i) hpc_sweep(s, 1.8, false, i, 0.0);
  hpc_download(s, psi_array, n);
}

int run_step_converged(void *solver_ptr, const double *j_array, double *psi_array, int size, int max_iterations,
                       double omega, double tolerance, double *final_delta_out) {
  if (final_delta_out) *final_delta_out = 0.0;
  if (!solver_ptr || !j_array || !psi_array || size <= 0) return 0;
  HpcSolver *s = static_cast<HpcSolver *>(solver_ptr);
  const size_t n = (size_t)size;
  if (n != (size_t)s->nr * s->nz) return 0;
  if (!hpc_upload(s, j_array, n)) return 0;
  const int n_iter = std::max(max_iterations, 1);
  const double omega_safe = std::isfinite(omega) ? omega : 1.8;
  const double tol_safe = std::isfinite(tolerance) ? tolerance : 0.0;
  const double w = std::min(std::max(omega_safe, 0.1), 1.99);
  const double tol = std::max(tol_safe, 0.0);
  cudaMemsetAsync(s->delta, 0, 2 * sizeof(unsigned long long), s->st);
  cudaMemsetAsync(s->flags, 0, 2 * sizeof(int), s->st);
  int h_flags[2] = {0, 0};
  int launched = 0;
  while (launched < n_iter) {
    const int chunk = std::min(32, n_iter - launched);
    for (int i = 0; i < chunk; ++i) hpc_sweep(s, w, true, launched + i, tol);
    launched += chunk;
    cudaMemcpyAsync(h_flags, s->flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->st);
    cudaStreamSynchronize(s->st);
    if (h_flags[0]) break;
  }
  unsigned long long bits[2] = {0, 0};
  cudaMemcpyAsync(bits, s->delta, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st);
  hpc_download(s, psi_array, n);
  double last;
  memcpy(&last, &bits[1], sizeof(double));
  if (final_delta_out) *final_delta_out = last;
  return h_flags[1];
}

void destroy_solver(void *solver_ptr) { hpc_free(static_cast<HpcSolver *>(solver_ptr)); }
void delete_solver(void *solver_ptr) { destroy_solver(solver_ptr); }

}  // extern "C"
