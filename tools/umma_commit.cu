// Microbenchmark: cost of the MMA thread's per-slot loop { wait(bar[slot]) ; n x tcgen05.mma (N = 192) ; tcgen05.commit -> bar[slot] }
// as a function of the MMAs per commit, for single CTAs (cta_group::1) and CTA pairs (cta_group::2, multicast commit).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/umma_commit.cu -o tools/bin/umma_commit
#include <cstdio>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

template <int PAIR>
__global__ void __launch_bounds__(128, 1) k(int n_mma, int reps, int do_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8], fin;
  __shared__ uint32_t tmem_s;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_init(&fin, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { if (PAIR) tmem_alloc_pair<512>(&tmem_s); else tmem_alloc<512>(&tmem_s); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const int n = 192;
  const uint32_t idesc = umma_idesc_f16(PAIR ? 256 : 128, n, false, false);
  const uint32_t brows = PAIR ? n / 2 : n;
  const uint64_t da0 = umma_smem_desc(smem_u32(smem), 2048, 128);
  const uint64_t db0 = umma_smem_desc(smem_u32(smem) + 65536, brows * 16, 128);
  if (threadIdx.x == 0 && rank == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int slot = r & 7;
      if (do_commit && r >= 8) mbar_wait(&bar[slot], ((r >> 3) - 1) & 1);
      for (int kk = 0; kk < n_mma; ++kk) {
        const uint64_t da = da0 + (kk & 3) * 256, db = db0 + (kk & 3) * ((brows * 32) >> 4);
        if (PAIR) umma_f16_pair(tmem + (r & 1) * 256, da, db, idesc, 1); else umma_f16(tmem + (r & 1) * 256, da, db, idesc, 1);
      }
      if (do_commit) { if (PAIR) umma_commit_pair(&bar[slot]); else umma_commit(&bar[slot]); }
    }
    const long long t1 = clock64();
    if (PAIR) umma_commit_pair(&fin); else umma_commit(&fin);
    mbar_wait(&fin, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(&fin, 0);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { if (PAIR) tmem_dealloc_pair<512>(tmem); else tmem_dealloc<512>(tmem); }
}

template <int PAIR>
void run(long long* out) {
  cudaFuncSetAttribute(k<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int reps = 512;
  for (int dc = 1; dc >= 0; --dc)
    for (int n_mma : {2, 4, 6, 8, 12}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(PAIR ? 2 : 1); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 160 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        cudaLaunchKernelEx(&lc, k<PAIR>, n_mma, reps, dc, out);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return; }
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      printf("%s %s %2d MMAs per slot: issue loop %7.1f cycles/slot, until done %7.1f cycles/slot (tensor pipe needs %d)\n", PAIR ? "pair  " : "single",
             dc ? "commit   " : "no commit", n_mma, static_cast<double>(h[0]) / reps, static_cast<double>(h[1]) / reps, n_mma * 96);
    }
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  run<0>(out);
  run<1>(out);
  return 0;
}
