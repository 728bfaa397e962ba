// DMMA.8x8x4 issue microbenchmark (sm_100a): cycles per DMMA per SM sub-partition for several instruction patterns.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/dmma_pipe tools/ubench/dmma_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: u outer, v inner (b changes fastest)   MODE 1: v outer, u inner   MODE 2: like 0 with 8 LDS.64 per 16 DMMA
// MODE 3: 8 independent accumulators only (2x4)  MODE 4: serpentine
// second family: MODE 2 pattern (fragments from shared memory) plus, per K tile of 4 k-steps (= 64 DMMA per warp):
//   F & 1: 8 cp.async 8-byte copies global -> shared (another stage) per thread
//   F & 2: one __syncthreads()
//   F & 4: ~40 integer instructions (address arithmetic stand-in)
template <int F>
__global__ void __launch_bounds__(512, 1) kf(double *out, const double *in, const double *big, int tiles, long long *cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = in[i & 4095];
  __syncthreads();
  double acc[4][4][2];
  for (int u = 0; u < 4; ++u) for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3, warp = threadIdx.x >> 5;
  const int wb = (warp >> 2) * 32, ww = (warp & 3) * 32;
  const int lk = threadIdx.x & 15, lr = threadIdx.x >> 4;
  unsigned junk = threadIdx.x;
  long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    if (F & 2) __syncthreads();
    if (F & 1) {
      double *dst = sm + 8192 + ((t & 1) * 5120) + lr * 20 + lk;
      const double *src = big + ((size_t)blockIdx.x * 997 + t) * 4096 % (1 << 24) + lr * 129 + lk;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + q * 640);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, 8;" ::"r"(d), "l"(src + q * 4128) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    if (F & 4) {
#pragma unroll
      for (int q = 0; q < 40; ++q) junk = junk * 1664525u + 1013904223u + (junk >> 7);
    }
#pragma unroll
    for (int ks = 0; ks < 16; ks += 4) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = sm[(wb + u * 8 + gid) * 20 + ks + tig];
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = sm[2560 + (ww + u * 8 + gid) * 20 + ks + tig];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    }
  }
  long long t1 = clock64();
  double s = (double)junk * 1e-300;
  for (int u = 0; u < 4; ++u) for (int v = 0; v < 4; ++v) s += acc[u][v][0] + acc[u][v][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int F> void runf(const char *name, double *out, double *in, double *big, long long *cyc) {
  const int tiles = 4000, threads = 512, smem = 180 * 1024;
  cudaFuncSetAttribute(kf<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  kf<F><<<148, threads, smem>>>(out, in, big, 50, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kf<F><<<148, threads, smem>>>(out, in, big, tiles, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double dmma_per_smsp = (double)tiles * 64 * 4;
  printf("%-46s %6.2f cycles/DMMA/SMSP  %6.2f TFLOP/s  (%s)\n", name, c / dmma_per_smsp, 148.0 * 4 * dmma_per_smsp * 512 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
__global__ void k(double *out, const double *in, int iters, long long *cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  double acc[4][4][2];
  for (int u = 0; u < 4; ++u) for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
  double a[4], b[4];
  for (int u = 0; u < 4; ++u) a[u] = in[threadIdx.x + 32 * u], b[u] = in[threadIdx.x + 32 * u + 128];
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
      const int ks = (it & 3) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = sm[(u * 8 + gid) * 20 + ks + tig];
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = sm[2560 + (u * 8 + gid) * 20 + ks + tig];
    }
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    } else if (MODE == 1) {
#pragma unroll
      for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int u = 0; u < 4; ++u) dmma(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    } else if (MODE == 3) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) dmma(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) { const int vv = (u & 1) ? 3 - v : v; dmma(acc[u][vv][0], acc[u][vv][1], a[u], b[vv]); }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int u = 0; u < 4; ++u) for (int v = 0; v < 4; ++v) s += acc[u][v][0] + acc[u][v][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name, int threads, double *out, double *in, long long *cyc) {
  const int iters = 20000;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<148, threads, 65536>>>(out, in, 100, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads, 65536>>>(out, in, iters, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = threads / 32 / 4.0;
  const double dmma_per_smsp = (double)iters * 16 * warps_per_smsp;
  const double tf = 148.0 * 4 * dmma_per_smsp * 512 / (ms * 1e-3) / 1e12;
  printf("%-34s threads %4d: %6.2f cycles/DMMA/SMSP  %6.2f TFLOP/s  (%.3f ms)\n", name, threads, c / dmma_per_smsp, tf, ms);
}
int main() {
  double *out, *in; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 4096 * 8); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 4096 * 8);
  for (int threads : {128, 256, 512}) {
    if (threads == 128) { run<0>("u outer, v inner", 128, out, in, cyc); run<1>("v outer, u inner", 128, out, in, cyc); run<2>("u outer + 8 LDS.64 / 16 DMMA", 128, out, in, cyc); run<3>("8 accumulators", 128, out, in, cyc); run<4>("serpentine", 128, out, in, cyc); }
    if (threads == 256) { run<0>("u outer, v inner", 256, out, in, cyc); run<2>("u outer + 8 LDS.64 / 16 DMMA", 256, out, in, cyc); }
    if (threads == 512) { run<0>("u outer, v inner", 512, out, in, cyc); run<1>("v outer, u inner", 512, out, in, cyc); run<2>("u outer + 8 LDS.64 / 16 DMMA", 512, out, in, cyc); run<4>("serpentine", 512, out, in, cyc); }
    
  }
  double *big; cudaMalloc(&big, ((size_t)(1 << 24) + (1 << 20)) * 8); cudaMemset(big, 0, ((size_t)(1 << 24) + (1 << 20)) * 8);
  runf<0>("tile loop: fragments from smem only", out, in, big, cyc);
  runf<1>("+ 8 cp.async/thread/tile", out, in, big, cyc);
  runf<2>("+ __syncthreads/tile", out, in, big, cyc);
  runf<4>("+ 40 integer instr/tile", out, in, big, cyc);
  runf<3>("+ cp.async + barrier", out, in, big, cyc);
  runf<7>("+ cp.async + barrier + integer", out, in, big, cyc);
  return 0;
}
