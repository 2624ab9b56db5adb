// fp64_peak.cu — box probe: fp64 FMA peak, DMMA (mma.sync m8n8k4 f64) peak,
// both concurrently, and a plain copy bandwidth, measured with CUDA events.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const double b = 1.0000001, c = 0.9999999;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
  double a = 1.0000001, b = 0.9999999;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void mixed_kernel(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  const double b = 1.0000001, cc = 0.9999999;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) dmma(c[j][0], c[j][1], b, cc);
    a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
    a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, s = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += s) out[i] = in[i];
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters); }, 5);
  printf("DFMA : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 8 * iters * blocks * threads / ms * 1e-9, ms);
  ms = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters); }, 5);
  // one m8n8k4 = 8*8*4 MAC per warp
  printf("DMMA : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 256 * 8.0 * iters * blocks * (threads / 32) / ms * 1e-9, ms);
  ms = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, iters); }, 5);
  double fl = 2.0 * 8 * iters * blocks * threads + 2.0 * 256 * 4.0 * iters * blocks * (threads / 32);
  printf("MIXED: %.2f TFLOP/s total (%.3f ms): DFMA part %.2f, DMMA part %.2f\n", fl / ms * 1e-9, ms,
         2.0 * 8 * iters * blocks * threads / ms * 1e-9, 2.0 * 256 * 4.0 * iters * blocks * (threads / 32) / ms * 1e-9);
  size_t n = (size_t)1 << 27;  // 2 GiB each
  double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
  ms = time_ms([&] { copy_kernel<<<p.multiProcessorCount * 16, 512>>>(a, b, n); }, 5);
  printf("COPY : %.1f GB/s (read+write)\n", 2.0 * n * 16 / ms * 1e-6);
  return 0;
}
