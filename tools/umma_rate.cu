// Microbenchmark: issue rate of tcgen05.mma kind::f16 (M=128) for different shared-memory operand layouts.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/umma_rate.cu -o tools/bin/umma_rate
// Only the access pattern matters here (operands are zero-filled); results are cycles per MMA instruction.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

struct Cfg {
  const char* name;
  int n;                 // MMA N
  int a_mn, b_mn;        // 1 = MN-major
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t swz;          // descriptor layout type (0 none, 2 = 128B, 4 = 64B, 6 = 32B)
  uint32_t a_kstep, b_kstep;   // bytes added to the start address per K=16 step
  int ksteps;            // K steps cycled through
  uint32_t b_base;       // byte offset of B in smem
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, c.n, c.a_mn, c.b_mn);
    const uint64_t da = umma_smem_desc(smem_u32(smem), c.a_lbo, c.a_sbo) | (static_cast<uint64_t>(c.swz) << 61);
    const uint64_t db = umma_smem_desc(smem_u32(smem) + c.b_base, c.b_lbo, c.b_sbo) | (static_cast<uint64_t>(c.swz) << 61);
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + (k & 1) * 256, da + ((k * c.a_kstep) >> 4), db + ((k * c.b_kstep) >> 4), idesc, 1);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph); ph ^= 1;
      const long long t1 = clock64();
      if (pass == 1) out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  const int reps = 2048;
  Cfg cfgs[] = {
    // K-major no swizzle, contiguous k-group planes (the lgemm / chain layout): A 128 rows, B n rows
    {"Kmaj  none  N=256 (chain)",      256, 0, 0, 2048, 128, 4096, 128, 0, 4096, 8192, 4, 65536},
    {"Kmaj  none  N=192 (lgemm)",      192, 0, 0, 2048, 128, 3072, 128, 0, 4096, 6144, 4, 65536},
    {"Kmaj  none  N=128",              128, 0, 0, 2048, 128, 2048, 128, 0, 4096, 4096, 4, 65536},
    {"Kmaj  none  N=64",                64, 0, 0, 2048, 128, 1024, 128, 0, 4096, 2048, 4, 65536},
    // transposed chain (chain_t.cu): A = packed weight stage (k-group stride 4096), B = NR-row image, k-group stride NR*16 (+16)
    {"T: A w-stage, B N=80 LB=1296",    80, 0, 0, 4096, 128, 1296, 128, 0, 8192, 2592, 4, 65536},
    {"T: A w-stage, B N=80 LB=1280",    80, 0, 0, 4096, 128, 1280, 128, 0, 8192, 2560, 4, 65536},
    {"T: A w-stage, B N=80 LB=1408",    80, 0, 0, 4096, 128, 1408, 128, 0, 8192, 2816, 4, 65536},
    {"T: A 2048,    B N=80 LB=1280",    80, 0, 0, 2048, 128, 1280, 128, 0, 4096, 2560, 4, 65536},
    {"T: A w-stage, B N=64 LB=1024",    64, 0, 0, 4096, 128, 1024, 128, 0, 8192, 2048, 4, 65536},
    {"T: A w-stage, B N=128 LB=2048",  128, 0, 0, 4096, 128, 2048, 128, 0, 8192, 4096, 4, 65536},
    {"T: A w-stage, B N=32 LB=528",     32, 0, 0, 4096, 128, 528, 128, 0, 8192, 1056, 4, 65536},
    {"T: A w-stage, B N=16 LB=272",     16, 0, 0, 4096, 128, 272, 128, 0, 8192, 544, 4, 65536},
    // MN-major no swizzle over the 128-row image (the wgrad layout): LBO 128 (k groups), SBO 2048 (mn groups)
    {"MNmaj none  N=128 (wgrad)",      128, 1, 1, 128, 2048, 128, 2048, 0, 256, 256, 8, 32768},
    {"MNmaj none  N=256",              256, 1, 1, 128, 2048, 128, 2048, 0, 256, 256, 8, 32768},
    // mixed: A MN-major none, B K-major none
    {"A MN / B K none N=128",          128, 1, 0, 128, 2048, 2048, 128, 0, 256, 4096, 4, 65536},
    {"A K / B MN none N=128",          128, 0, 1, 2048, 128, 128, 2048, 0, 4096, 256, 4, 65536},
    // K-major 128B swizzle: rows of 64 halves, 8-row atoms of 1024 B
    {"Kmaj  sw128 N=256",              256, 0, 0, 16, 1024, 16, 1024, 2, 32, 32, 4, 65536},
    {"Kmaj  sw128 N=128",              128, 0, 0, 16, 1024, 16, 1024, 2, 32, 32, 4, 65536},
    // MN-major 128B swizzle: atoms of 64 mn x 8 k (1024 B); [k/8][mn/64] -> LBO 1024 (mn atoms), SBO 2048/4096 (k groups)
    {"MNmaj sw128 N=128",              128, 1, 1, 1024, 2048, 1024, 2048, 2, 4096, 4096, 8, 65536},
    {"MNmaj sw128 N=256",              256, 1, 1, 1024, 2048, 1024, 4096, 2, 4096, 8192, 8, 65536},
    // MN-major no swizzle with mn groups contiguous: [k/8][mn/8][8 k][16 B]: LBO = (M/8)*128, SBO = 128
    {"MNmaj none contig-mn N=128",     128, 1, 1, 2048, 128, 2048, 128, 0, 4096, 4096, 8, 65536},
  };
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 256);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (const Cfg& c : cfgs) {
    rate_kernel<<<grid, 128, 200 * 1024>>>(c, reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-32s ERROR %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[256];
    cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = static_cast<double>(mx) / reps;
    printf("%-32s %8.1f cycles/MMA   %7.0f flop/cycle/SM\n", c.name, cyc, 2.0 * 128 * c.n * 16 / cyc);
  }
  return 0;
}
