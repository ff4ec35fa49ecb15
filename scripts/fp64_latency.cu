// fp64_latency.cu -- dependent-issue latency of the fp64 operations the association gate is made of (one warp alone on an
// SM; clock64 around a chain of N dependent operations).  Why: the Monte-Carlo kernel and the line loop are chains of
// ~10^2 dependent fp64 operations per line; their floor is (chain length) x (latency), not the fp64 pipe's throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/fp64_latency scripts/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 2048
template <int OP>
__global__ void k_chain(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = __fma_rn(x, b, a);
    else if (OP == 1) x = __dadd_rn(x, b);
    else if (OP == 2) x = __dmul_rn(x, b);
    else if (OP == 3) x = __ddiv_rn(a, x + b);
    else if (OP == 4) x = sqrt(x + b);
    else if (OP == 5) x = (x != 0.0) ? __dadd_rn(x, b) : x;      /* the zero-skip pattern: DSETP + select */
    else if (OP == 6) x = sin(x) + cos(x);
    else if (OP == 7) { float f = (float)x; f = __fmaf_rn(f, 1.0001f, 0.5f); x = (double)f; }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { *cyc = t1 - t0; }
  out[threadIdx.x] = x;
}
template <int OP>
void run(const char* name, double a, double b, double* out, long long* cyc) {
  k_chain<OP><<<1, 32>>>(out, cyc, a, b);
  k_chain<OP><<<1, 32>>>(out, cyc, a, b);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %7.1f cycles per dependent step\n", name, (double)c / N);
}
__global__ void k_lds(double* out, long long* cyc) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = idx[p];
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  out[threadIdx.x] = p;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  run<0>("DFMA  x = fma(x, b, a)", 1.0, 0.999999, out, cyc);
  run<1>("DADD  x = x + b", 1.0, 1e-9, out, cyc);
  run<2>("DMUL  x = x * b", 1.0, 1.0000001, out, cyc);
  run<5>("zero skip: x = (x != 0) ? x + b : x", 1.0, 1e-9, out, cyc);
  run<3>("x = a / (x + b)   (IEEE double division + add)", 1.3, 0.7, out, cyc);
  run<4>("x = sqrt(x + b)   (+ add)", 2.0, 1.1, out, cyc);
  run<6>("x = sin(x) + cos(x)", 0.3, 0.0, out, cyc);
  run<7>("fp32 FFMA via two conversions", 1.0, 0.0, out, cyc);
  k_lds<<<1, 32>>>(out, cyc); k_lds<<<1, 32>>>(out, cyc); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %7.1f cycles per dependent step\n", "LDS   p = idx[p] (shared-memory pointer chase)", (double)c / N);
  return 0;
}
