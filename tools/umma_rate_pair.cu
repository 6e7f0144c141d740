// Microbenchmark: issue rate of tcgen05.mma.cta_group::2 kind::f16 (CTA pair, M = 256 or 128) vs cta_group::1, K-major
// SWIZZLE_NONE operands as the layer GEMMs use them (each CTA holds its own A rows and HALF of the B rows).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/umma_rate_pair.cu -o tools/bin/umma_rate_pair
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

__global__ void __launch_bounds__(128, 1) pair_kernel(int m, int n, int same_acc, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc_pair<512>(&tmem_s);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_f16(m, n, false, false);
    const uint32_t brows = n / 2;
    const uint64_t da = umma_smem_desc(smem_u32(smem), (m / 2) * 16, 128);
    const uint64_t db = umma_smem_desc(smem_u32(smem) + 65536, brows * 16, 128);
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_pair(tmem + (same_acc ? 0 : (k & 1) * 256), da + ((k * (m / 2) * 32) >> 4), db + ((k * brows * 32) >> 4), idesc, 1);
      }
      umma_commit_pair(&bar);
      mbar_wait(&bar, ph); ph ^= 1;
      const long long t1 = clock64();
      if (pass == 1) out[blockIdx.x >> 1] = t1 - t0;
    }
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0); mbar_wait(&bar, 1);
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem);
}

int main(int argc, char** argv) {
  const int pairs = argc > 1 ? atoi(argv[1]) : 1;
  const int reps = 2048;
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 256);
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int cfg[][3] = {{256, 192, 0}, {256, 192, 1}, {256, 256, 0}, {256, 128, 0}, {256, 96, 0}, {128, 192, 0}, {128, 256, 0}};
  for (const auto& c : cfg) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, pair_kernel, c[0], c[1], c[2], reps, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d ERROR %s\n", c[0], c[1], cudaGetErrorString(e)); return 1; }
    long long h[256];
    cudaMemcpy(h, out, sizeof(long long) * pairs, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < pairs; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = static_cast<double>(mx) / reps;
    printf("cta_group::2 M=%d N=%d %s  %8.1f cycles/MMA   %7.0f flop/cycle/SM\n", c[0], c[1], c[2] ? "one acc " : "two accs", cyc,
           2.0 * c[0] * c[1] * 16 / cyc / 2);
  }
  return 0;
}
