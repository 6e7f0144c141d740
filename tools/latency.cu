// Microbenchmark: latencies that make up the producer <-> MMA ring round trip.
//  1. mbarrier ping-pong between two warps (arrive -> other thread's try_wait returns), per one-way hop
//  2. tcgen05.commit (nothing outstanding) -> waiting thread wakes
//  3. tcgen05.mma (N=192) x4 + commit -> waiting thread wakes
//  4. one bulk copy of S bytes, issue -> own try_wait returns (idle chip / 148 CTAs doing the same)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/latency.cu -o tools/bin/latency
#include <cstdio>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, uint32_t copy_bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t ping, pong, cbar, fbar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&ping, 1); mbar_init(&pong, 1); mbar_init(&cbar, 1); mbar_init(&fbar, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int R = 256;
  // 1. ping-pong
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) { mbar_arrive(&ping); mbar_wait(&pong, r & 1); }
    out[blockIdx.x * 8 + 0] = (clock64() - t0) / (2 * R);
  } else if (threadIdx.x == 32) {
    for (int r = 0; r < R; ++r) { mbar_wait(&ping, r & 1); mbar_arrive(&pong); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // 2. empty commit -> own wait
    long long t0 = clock64();
    for (int r = 0; r < R; ++r) { umma_commit(&cbar); mbar_wait(&cbar, r & 1); }
    out[blockIdx.x * 8 + 1] = (clock64() - t0) / R;
    // 3. 4 MMAs + commit -> own wait
    const uint32_t idesc = umma_idesc_f16(128, 192, false, false);
    const uint64_t da = umma_smem_desc(smem_u32(smem), 2048, 128), db = umma_smem_desc(smem_u32(smem) + 16384, 3072, 128);
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_s, da + kk * 256, db + kk * 384, idesc, 1);
      umma_commit(&cbar); mbar_wait(&cbar, r & 1);
    }
    out[blockIdx.x * 8 + 2] = (clock64() - t0) / R;
    // 4. single copy latency
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
      mbar_arrive_expect_tx(&fbar, copy_bytes);
      bulk_g2s(smem, src + (static_cast<size_t>(blockIdx.x) * R + r) % 1024 * 32768, copy_bytes, &fbar);
      mbar_wait(&fbar, r & 1);
    }
    out[blockIdx.x * 8 + 3] = (clock64() - t0) / R;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_s);
}

int main() {
  uint8_t* src; cudaMalloc(&src, 32u << 20); cudaMemset(src, 0, 32u << 20);
  long long* out; cudaMalloc(&out, 8 * 8 * 256);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65 * 1024);
  for (int g : {1, 148})
    for (uint32_t cb : {2048u, 16384u, 32768u}) {
      for (int rep = 0; rep < 2; ++rep) k<<<g, 64, 65 * 1024>>>(src, cb, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[8 * 256]; cudaMemcpy(h, out, 8 * 8 * g, cudaMemcpyDeviceToHost);
      long long m[4] = {0, 0, 0, 0};
      for (int i = 0; i < g; ++i) for (int j = 0; j < 4; ++j) m[j] = h[i * 8 + j] > m[j] ? h[i * 8 + j] : m[j];
      printf("CTAs %3d: mbarrier hop %4lld cyc | empty commit->wake %4lld | 4 MMA(N=192)+commit->wake %4lld (tensor 384) | %5u B copy issue->wake %5lld\n",
             g, m[0], m[1], m[2], cb, m[3]);
    }
  return 0;
}
