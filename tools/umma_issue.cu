// Microbenchmark: cost of ISSUING tcgen05.mma from different code shapes (N=64 so the tensor pipe needs only 32 cycles
// per instruction and the issue path is what is measured).
//   style 0: `if (threadIdx.x == 0)` around the whole loop, descriptors computed by that thread (vector registers)
//   style 1: whole warp runs the loop, descriptors warp-uniform, only the mma under `if (elect_one())`
//   style 2: whole warp runs the loop, `if (lane == 0)` around the mma only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/umma_issue.cu -o tools/bin/umma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

template <int STYLE>
__global__ void __launch_bounds__(128, 1) k(int n, int reps, uint32_t slot_bytes, int n_slots, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = umma_idesc_f16(128, n, false, false);
  const uint64_t da0 = umma_smem_desc(smem_u32(smem), 2048, 128);
  const uint64_t db0 = umma_smem_desc(smem_u32(smem) + 16384, n * 16, 128);
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    if (STYLE == 0) {
      if (threadIdx.x == 0) {
        uint32_t slot = 0;
        for (int r = 0; r < reps; ++r) {
          const uint32_t so = (slot * slot_bytes) >> 4;
#pragma unroll
          for (int kk = 0; kk < 6; ++kk) umma_f16(tmem, da0 + so + (kk >> 1) * 256, db0 + so + (kk & 1) * 64, idesc, (r | kk) != 0);
          if (++slot == static_cast<uint32_t>(n_slots)) slot = 0;
        }
        umma_commit(&bar);
      }
    } else {
      uint32_t slot = 0;
      const bool leader = STYLE == 1 ? elect_one() : (threadIdx.x == 0);
      for (int r = 0; r < reps; ++r) {
        const uint32_t so = (slot * slot_bytes) >> 4;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 6; ++kk) umma_f16(tmem, da0 + so + (kk >> 1) * 256, db0 + so + (kk & 1) * 64, idesc, (r | kk) != 0);
        }
        __syncwarp();
        if (++slot == static_cast<uint32_t>(n_slots)) slot = 0;
      }
      if (leader) umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int STYLE>
void run(const char* name, long long* out) {
  cudaFuncSetAttribute(k<STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int reps = 512;
  for (int rep = 0; rep < 2; ++rep) k<STYLE><<<1, 128, 160 * 1024>>>(64, reps, 20480, 5, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s ERROR %s\n", name, cudaGetErrorString(e)); return; }
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-60s %6.1f cycles per MMA (tensor pipe needs 32)\n", name, static_cast<double>(h) / (reps * 6));
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  run<0>("single thread owns the loop (current kernels)", out);
  run<1>("convergent warp, mma under elect_one()", out);
  run<2>("convergent warp, mma under lane == 0", out);
  return 0;
}
