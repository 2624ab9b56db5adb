// dfma_lat.cu — box probe: DFMA dependent-issue latency and per-SM throughput as a function of the
// number of independent chains per thread (ILP) and resident warps per SM sub-partition.
// Not part of the product; numbers feed DESIGN.md §5.1 (occupancy / ILP needed by the fast kernel).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, int iters, long long* cycles) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
  const double b = 1.0000001, c = 0.9999999;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 2000;
  chain<ILP><<<148, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  chain<ILP><<<148, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per_inst = (double)c / (iters * 8.0 * ILP);
  // per SMSP: warps_per_sm/4 warps, each issuing ILP*8*iters DFMA in c cycles
  double rate = (warps_per_sm / 4.0) * iters * 8.0 * ILP / c;  // warp-DFMA per cycle per SMSP
  printf("ILP %d warps/SM %2d : %.2f cycles per DFMA per warp, %.3f warp-DFMA/clk/SMSP (peak 0.5)\n", ILP,
         warps_per_sm, per_inst, rate);
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 12, 16, 32}) {
    run<1>(w, out, cyc);
    run<2>(w, out, cyc);
    run<3>(w, out, cyc);
    run<4>(w, out, cyc);
    run<6>(w, out, cyc);
    run<8>(w, out, cyc);
  }
  return 0;
}
