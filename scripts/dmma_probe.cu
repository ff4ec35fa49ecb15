// dmma_probe.cu -- what is an fp64 tensor-core MMA on B200 worth for the rank-2m covariance sweep?
//   1. rounding: is D = A*B + C of mma.sync.m8n8k4 / m16n8k8 / m16n8k16 (.f64) bit-identical to the chain
//      d = fma(a_k, b_k, d), k = 0 .. K-1 (one rounding per product-add, in k order)?  If it is, a DMMA sweep gives the
//      same bits as sub_rank2() in ekf_device.cuh applied term after term.
//   2. rate: register-resident DMMA vs DFMA, TFLOP/s over the whole GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dmma_probe.bin scripts/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// A: M x K row-major, B: K x 8 (B[k][n]), C/D: M x 8.  One warp per problem.
template <int SHAPE>
__global__ void k_check(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ C, double* __restrict__ D) {
  constexpr int Mr = SHAPE == 0 ? 8 : 16, K = SHAPE == 0 ? 4 : (SHAPE == 1 ? 8 : 16);
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int prob = blockIdx.x;
  const double* a = A + (size_t)prob * Mr * K;
  const double* b = B + (size_t)prob * K * 8;
  const double* c = C + (size_t)prob * Mr * 8;
  double* d = D + (size_t)prob * Mr * 8;
  if (SHAPE == 0) {
    double d0 = c[g * 8 + 2 * t], d1 = c[g * 8 + 2 * t + 1];
    mma884(d0, d1, a[g * K + t], b[t * 8 + g]);
    d[g * 8 + 2 * t] = d0; d[g * 8 + 2 * t + 1] = d1;
  } else if (SHAPE == 1) {
    double dd[4] = {c[g * 8 + 2 * t], c[g * 8 + 2 * t + 1], c[(g + 8) * 8 + 2 * t], c[(g + 8) * 8 + 2 * t + 1]};
    double aa[4], bb[2];
    for (int i = 0; i < 4; ++i) aa[i] = a[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
    for (int i = 0; i < 2; ++i) bb[i] = b[(t + 4 * i) * 8 + g];
    mma1688(dd, aa, bb);
    d[g * 8 + 2 * t] = dd[0]; d[g * 8 + 2 * t + 1] = dd[1]; d[(g + 8) * 8 + 2 * t] = dd[2]; d[(g + 8) * 8 + 2 * t + 1] = dd[3];
  } else {
    double dd[4] = {c[g * 8 + 2 * t], c[g * 8 + 2 * t + 1], c[(g + 8) * 8 + 2 * t], c[(g + 8) * 8 + 2 * t + 1]};
    double aa[8], bb[4];
    for (int i = 0; i < 8; ++i) aa[i] = a[(g + 8 * (i & 1)) * K + t + 4 * (i >> 1)];
    for (int i = 0; i < 4; ++i) bb[i] = b[(t + 4 * i) * 8 + g];
    mma16816(dd, aa, bb);
    d[g * 8 + 2 * t] = dd[0]; d[g * 8 + 2 * t + 1] = dd[1]; d[(g + 8) * 8 + 2 * t] = dd[2]; d[(g + 8) * 8 + 2 * t + 1] = dd[3];
  }
}

// throughput: every warp keeps ACC independent accumulator tiles and issues `iters` rounds of MMAs on them
template <int SHAPE, int ACC>
__global__ void __launch_bounds__(256) k_rate_mma(double* out, int iters, double seed) {
  const int lane = threadIdx.x & 31;
  double acc[ACC][4];
  for (int i = 0; i < ACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = seed * (i + j + lane);
  double a8[8], b4[4];
  for (int i = 0; i < 8; ++i) a8[i] = 1.0 + 1e-9 * (lane + i) * seed;
  for (int i = 0; i < 4; ++i) b4[i] = 1.0 - 1e-9 * (lane + i) * seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      if (SHAPE == 0) { mma884(acc[i][0], acc[i][1], a8[i & 7], b4[i & 3]); mma884(acc[i][2], acc[i][3], a8[(i + 1) & 7], b4[(i + 1) & 3]); }
      else if (SHAPE == 1) { double a4[4] = {a8[0], a8[1], a8[2], a8[3]}; double b2[2] = {b4[i & 1], b4[2 + (i & 1)]}; mma1688(acc[i], a4, b2); }
      else mma16816(acc[i], a8, b4);
    }
  }
  double s = 0.0;
  for (int i = 0; i < ACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  if (s == 123.456) out[0] = s;
}
template <int ACC>
__global__ void __launch_bounds__(256) k_rate_fma(double* out, int iters, double seed) {
  const int lane = threadIdx.x & 31;
  double acc[ACC];
  for (int i = 0; i < ACC; ++i) acc[i] = seed * (i + lane);
  const double a = 1.0 + 1e-9 * lane * seed, b = 1.0 - 1e-9 * lane * seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = fma(a, acc[i], b);
  }
  double s = 0.0;
  for (int i = 0; i < ACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

static double rnd() { return (double)rand() / RAND_MAX; }
static double rnd_wide() { const double m = rnd() * 2.0 - 1.0; return ldexp(m, (rand() % 40) - 20); }

template <int SHAPE>
void check(const char* name) {
  constexpr int Mr = SHAPE == 0 ? 8 : 16, K = SHAPE == 0 ? 4 : (SHAPE == 1 ? 8 : 16);
  const int P = 4096;
  std::vector<double> A((size_t)P * Mr * K), B((size_t)P * K * 8), C((size_t)P * Mr * 8), D((size_t)P * Mr * 8);
  for (auto& v : A) v = rnd_wide();
  for (auto& v : B) v = rnd_wide();
  for (auto& v : C) v = rnd_wide();
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
  k_check<SHAPE><<<P, 32>>>(dA, dB, dC, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); return; }
  cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
  long long n = 0, seq = 0, rev = 0, close = 0, unfused = 0;
  for (int p = 0; p < P; ++p)
    for (int r = 0; r < Mr; ++r)
      for (int c = 0; c < 8; ++c) {
        const double* a = &A[((size_t)p * Mr + r) * K];
        const double* b = &B[(size_t)p * K * 8 + c];
        const double c0 = C[((size_t)p * Mr + r) * 8 + c], got = D[((size_t)p * Mr + r) * 8 + c];
        double s = c0; for (int k = 0; k < K; ++k) s = fma(a[k], b[k * 8], s);
        double q = c0; for (int k = K - 1; k >= 0; --k) q = fma(a[k], b[k * 8], q);
        volatile double u = c0; for (int k = 0; k < K; ++k) { volatile double pr = a[k] * b[k * 8]; u = u + pr; }
        ++n;
        if (memcmp(&s, &got, 8) == 0) ++seq;
        if (memcmp(&q, &got, 8) == 0) ++rev;
        if (u == got) ++unfused;
        double mag = fabs(c0); for (int k = 0; k < K; ++k) mag += fabs(a[k] * b[k * 8]);
        if (fabs(s - got) <= 1e-12 * mag) ++close;
      }
  printf("%s: %lld results; bit-equal to the k-ordered FMA chain: %lld (%.4f%%); reverse chain: %lld; unfused mul+add: %lld; within 1e-12 of it (layout check): %lld\n",
         name, n, seq, 100.0 * seq / n, rev, unfused, close);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dD);
}

template <typename F>
double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < reps; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float t; cudaEventElapsedTime(&t, e0, e1); return t / reps;
}

int main() {
  srand(12345);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("device: %s, %d SMs\n", pr.name, pr.multiProcessorCount);
  check<0>("mma.m8n8k4.f64  ");
  check<1>("mma.m16n8k8.f64 ");
  check<2>("mma.m16n8k16.f64");
  double* out; cudaMalloc(&out, 8);
  const int blocks = pr.multiProcessorCount * 4, iters = 4000;
  const double warps = (double)blocks * 8;
  {
    const double ms = time_ms([&] { k_rate_fma<16><<<blocks, 256>>>(out, iters, 1.0); }, 5);
    printf("DFMA  (16 independent chains / thread): %.2f TFLOP/s\n", warps * 32 * 16.0 * iters * 2 / (ms * 1e-3) / 1e12);
  }
  {
    const double ms = time_ms([&] { k_rate_mma<0, 8><<<blocks, 256>>>(out, iters, 1.0); }, 5);
    printf("DMMA m8n8k4  (8 acc tiles x 2 / warp): %.2f TFLOP/s\n", warps * 8 * 2 * (8.0 * 8 * 4) * iters * 2 / (ms * 1e-3) / 1e12);
  }
  {
    const double ms = time_ms([&] { k_rate_mma<1, 8><<<blocks, 256>>>(out, iters, 1.0); }, 5);
    printf("DMMA m16n8k8 (8 acc tiles / warp):     %.2f TFLOP/s\n", warps * 8 * (16.0 * 8 * 8) * iters * 2 / (ms * 1e-3) / 1e12);
  }
  {
    const double ms = time_ms([&] { k_rate_mma<2, 8><<<blocks, 256>>>(out, iters, 1.0); }, 5);
    printf("DMMA m16n8k16 (8 acc tiles / warp):    %.2f TFLOP/s\n", warps * 8 * (16.0 * 8 * 16) * iters * 2 / (ms * 1e-3) / 1e12);
  }
  {
    const double ms = time_ms([&] { k_rate_mma<2, 16><<<blocks, 256>>>(out, iters, 1.0); }, 5);
    printf("DMMA m16n8k16 (16 acc tiles / warp):   %.2f TFLOP/s\n", warps * 16 * (16.0 * 8 * 16) * iters * 2 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
