// Micro-benchmark: in-place read-modify-write stream over a row-major fp64 matrix in tiles of R rows x W columns
// (each warp-row access = W*8 contiguous bytes), to see how DRAM efficiency depends on the segment width.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/segbench.cu -o /tmp/segbench
#include <cstdio>
#include <cuda_runtime.h>
template <int W>   // W columns per tile row; 256 threads; each thread 2 columns (double2) => W/2 lanes per row
__global__ void __launch_bounds__(256, 2) rmw(double* P, int ld, int n, int R) {
  const int tiles_x = n / W, tiles_y = n / R;
  const int lanes_per_row = W / 2;
  const int rows_per_pass = 256 / lanes_per_row;
  const int lr = threadIdx.x / lanes_per_row, lc = threadIdx.x % lanes_per_row;
  for (long long t = blockIdx.x; t < (long long)tiles_x * tiles_y; t += gridDim.x) {
    const int ty = t / tiles_x, tx = t % tiles_x;
    double* base = P + (size_t)(ty * R) * ld + (size_t)tx * W + 2 * lc;
    double2 v[8];
    const int per = R / rows_per_pass;   // rows handled by this thread (<= 8)
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < per) v[i] = __ldcs((const double2*)(base + (size_t)(lr + i * rows_per_pass) * ld));
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < per) { v[i].x -= 1.0; v[i].y -= 1.0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < per) __stcs((double2*)(base + (size_t)(lr + i * rows_per_pass) * ld), v[i]);
  }
}
__global__ void lin(double2* P, size_t n2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = __ldcs(P + i); v.x -= 1.0; v.y -= 1.0; __stcs(P + i, v);
  }
}
template <int W> void run(double* P, int ld, int n, int R, const char* name) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * 2 * 8;
  rmw<W><<<grid, 256>>>(P, ld, n, R);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) rmw<W><<<grid, 256>>>(P, ld, n, R);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-28s %.3f ms  %.1f GB/s\n", name, ms, 16.0 * n * (double)n / ms / 1e6);
}
int main() {
  const int n = 20480, ld = 20480;   // multiple of every tile shape
  double* P; cudaMalloc(&P, (size_t)ld * n * 8); cudaMemset(P, 0, (size_t)ld * n * 8);
  run<64>(P, ld, n, 64, "tile 64 rows x 64 cols (512B)");
  run<128>(P, ld, n, 32, "tile 32 x 128 (1KB)");
  run<256>(P, ld, n, 16, "tile 16 x 256 (2KB)");
  run<512>(P, ld, n, 8, "tile 8 x 512 (4KB)");
  run<128>(P, ld, n, 16, "tile 16 x 128 (1KB)");
  run<64>(P, ld, n, 32, "tile 32 x 64 (512B)");
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  lin<<<148 * 16, 256>>>((double2*)P, (size_t)ld * n / 2);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) lin<<<148 * 16, 256>>>((double2*)P, (size_t)ld * n / 2);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-28s %.3f ms  %.1f GB/s\n", "linear in-place", ms, 16.0 * n * (double)n / ms / 1e6);
  double* Q; cudaMalloc(&Q, (size_t)ld * n * 8);
  cudaMemcpy(Q, P, (size_t)ld * n * 8, cudaMemcpyDeviceToDevice);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) cudaMemcpyAsync(Q, P, (size_t)ld * n * 8, cudaMemcpyDeviceToDevice);
  cudaEventRecord(b); cudaEventSynchronize(b);
  cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-28s %.3f ms  %.1f GB/s\n", "cudaMemcpy D2D (r+w)", ms, 16.0 * n * (double)n / ms / 1e6);
  return 0;
}
