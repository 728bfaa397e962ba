// FP64 pipe microbenchmark: latency / throughput of DADD, DMUL, DFMA per SM sub-partition on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double *out, double a, double b, int iters, long long *cyc) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) v[i] = __dadd_rn(v[i], b);
      if (OP == 1) v[i] = __dmul_rn(v[i], b);
      if (OP == 2) v[i] = __fma_rn(v[i], b, a);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int OP>
void run(int threads, double *out, long long *cyc) {
  const int iters = 4096;
  k<ILP, OP><<<148, threads>>>(out, 1.0, 1.0000001, iters, cyc);
  cudaDeviceSynchronize();
  k<ILP, OP><<<148, threads>>>(out, 1.0, 1.0000001, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = threads / 32 / 4.0;
  const double inst_per_smsp = (double)iters * ILP * (warps_per_smsp < 1 ? 1 : warps_per_smsp);
  printf("op=%s threads=%4d ILP=%d: %8lld cycles, %.2f cycles per dependent step, %.2f cycles per warp-instr per SMSP\n",
         OP == 0 ? "DADD" : OP == 1 ? "DMUL" : "DFMA", threads, ILP, h, (double)h / iters, (double)h / inst_per_smsp);
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  run<1, 2>(32, out, cyc);  run<1, 0>(32, out, cyc); run<1, 1>(32, out, cyc);
  run<2, 2>(32, out, cyc);  run<4, 2>(32, out, cyc);  run<8, 2>(32, out, cyc);
  run<1, 2>(128, out, cyc); run<2, 2>(128, out, cyc); run<4, 2>(128, out, cyc); run<8, 2>(128, out, cyc);
  run<1, 2>(512, out, cyc); run<2, 2>(512, out, cyc); run<4, 2>(512, out, cyc); run<8, 2>(512, out, cyc);
  run<4, 0>(512, out, cyc); run<4, 1>(512, out, cyc);
  run<4, 2>(1024, out, cyc);
  return 0;
}
