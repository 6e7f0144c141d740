// Microbenchmark 2: what costs ~270 cycles per slot in { n x tcgen05.mma ; commit }?  Variants: unrolled inner loop (template),
// accumulator alternating per slot or fixed, single divergent thread vs convergent warp + elect_one.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/umma_commit2.cu -o tools/bin/umma_commit2
#include <cstdio>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

template <int NMMA, int ALT, int WARP>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8], fin;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&fin, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const int n = 192;
  const uint32_t idesc = umma_idesc_f16(128, n, false, false);
  const uint64_t da0 = umma_smem_desc(smem_u32(smem), 2048, 128);
  const uint64_t db0 = umma_smem_desc(smem_u32(smem) + 65536, n * 16, 128);
  if (threadIdx.x < 32 && (WARP || threadIdx.x == 0)) {
    const bool leader = WARP ? elect_one() : true;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int slot = r & 7;
      if (r >= 8) mbar_wait(&bar[slot], ((r >> 3) - 1) & 1);
      const uint32_t acc = tmem + (ALT ? (r & 1) * 256 : 0);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < NMMA; ++kk)
          umma_f16(acc, da0 + (kk & 3) * 256, db0 + (kk & 3) * ((n * 32) >> 4), idesc, 1);
        umma_commit(&bar[slot]);
      }
      if (WARP) __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) umma_commit(&fin);
    mbar_wait(&fin, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int NMMA, int ALT, int WARP>
void run(long long* out) {
  cudaFuncSetAttribute(k<NMMA, ALT, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int reps = 512;
  for (int rep = 0; rep < 2; ++rep) k<NMMA, ALT, WARP><<<1, 128, 160 * 1024>>>(reps, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return; }
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("%2d MMAs/slot, acc %s, %s: %7.1f cycles/slot (tensor pipe needs %d)\n", NMMA, ALT ? "alternating" : "fixed      ",
         WARP ? "convergent warp + elect" : "single divergent thread", static_cast<double>(h[1]) / reps, NMMA * 96);
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  run<2, 1, 0>(out); run<4, 1, 0>(out); run<6, 1, 0>(out); run<12, 1, 0>(out);
  run<2, 0, 0>(out); run<4, 0, 0>(out); run<6, 0, 0>(out); run<12, 0, 0>(out);
  run<2, 1, 1>(out); run<4, 1, 1>(out); run<6, 1, 1>(out); run<12, 1, 1>(out);
  run<6, 0, 1>(out);
  return 0;
}
