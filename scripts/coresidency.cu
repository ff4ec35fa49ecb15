// Can a small (cluster) kernel start on SMs that already host a big-shared-memory persistent CTA?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/coresidency.cu -o scripts/coresidency.bin
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(288, 1) bigA(unsigned long long* ts, unsigned long long dur_ns, double* sink) {
  extern __shared__ unsigned char sm[];
  if (threadIdx.x == 0) sm[0] = 1;
  const unsigned long long t0 = gtimer();
  if (blockIdx.x == 0 && threadIdx.x == 0) ts[0] = t0;
  double acc[48];                       // ~108 registers per thread, like the sweep kernel
  for (int i = 0; i < 48; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  while (gtimer() - t0 < dur_ns) {
    for (int i = 0; i < 48; ++i) acc[i] = acc[i] * 1.0000001 + acc[(i + 7) % 48];
  }
  double t = 0; for (int i = 0; i < 48; ++i) t += acc[i];
  if (t == 1.2345) sink[0] = t;
  if (blockIdx.x == 0 && threadIdx.x == 0) ts[1] = gtimer();
}
template <int REGS_DUMMY>
__global__ void __launch_bounds__(256, REGS_DUMMY >= 60 ? 1 : 2) smallB(unsigned long long* ts, double* sink) {
  __shared__ double s[512];
  double acc[REGS_DUMMY];
  for (int i = 0; i < REGS_DUMMY; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int k = 0; k < 100; ++k)
    for (int i = 0; i < REGS_DUMMY; ++i) acc[i] = acc[i] * 1.0000001 + acc[(i + 1) % REGS_DUMMY];
  double t = 0; for (int i = 0; i < REGS_DUMMY; ++i) t += acc[i];
  s[threadIdx.x] = t;
  if (blockIdx.x == 0 && threadIdx.x == 0) ts[2] = gtimer();
  if (t == 12345.678) sink[0] = s[threadIdx.x ^ 1];
}
template <int RD>
static void run(const char* name, bool cluster, int carve, size_t a_smem, int b_threads) {
  unsigned long long* ts; cudaMalloc(&ts, 64); cudaMemset(ts, 0, 64);
  double* sink; cudaMalloc(&sink, 8);
  cudaStream_t sa, sb; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
  cudaFuncSetAttribute(bigA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a_smem);
  if (carve >= 0) cudaFuncSetAttribute(smallB<RD>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(smallB<RD>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  bigA<<<148, 288, a_smem, sa>>>(ts, 600000ull, sink);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(16); cfg.blockDim = dim3(b_threads); cfg.stream = sb;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cluster ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, smallB<RD>, ts, sink);
  cudaDeviceSynchronize();
  unsigned long long h[3]; cudaMemcpy(h, ts, 24, cudaMemcpyDeviceToHost);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, smallB<RD>);
  printf("%-44s launch=%s  B started %+8.1f us after A start (A ran %.1f us)  => %s   [B regs %d]\n", name, cudaGetErrorString(e),
         ((double)h[2] - (double)h[0]) / 1e3, ((double)h[1] - (double)h[0]) / 1e3, h[2] < h[1] ? "CONCURRENT" : "serialized", fa.numRegs);
}
int main() {
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, bigA); printf("A regs %d\n", fa.numRegs);
  run<40>("cluster B(40), A 148.6KB", true, cudaSharedmemCarveoutMaxShared, 148608, 256);
  run<56>("cluster B(56), A 148.6KB", true, cudaSharedmemCarveoutMaxShared, 148608, 256);
  run<60>("cluster B(60, bounds 1), A 148.6KB", true, cudaSharedmemCarveoutMaxShared, 148608, 256);
  run<40>("cluster B(40), A 165KB", true, cudaSharedmemCarveoutMaxShared, 165000, 256);
  run<40>("cluster B(40), A 180KB", true, cudaSharedmemCarveoutMaxShared, 180000, 256);
  run<40>("cluster B(40), A 190KB", true, cudaSharedmemCarveoutMaxShared, 190000, 256);
  run<20>("cluster B(20), A 148.6KB", true, cudaSharedmemCarveoutMaxShared, 148608, 256);
  return 0;
}
