// Microbenchmark: what in the layer-GEMM's MMA-issue loop costs more than the tensor pipe?  N=192, 6 MMAs per "slot"
// (3 passes x 2 K-steps), all into one accumulator, optionally followed by tcgen05.commit to a ring of mbarriers and
// a wait on an mbarrier that a second thread arrives on (the producer handshake without any copies).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/umma_loop.cu -o tools/bin/umma_loop
#include <cstdio>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

constexpr int kSlots = 5;
// mode bit 0: commit per slot;  bit 1: producer handshake (wait full / producer waits empty);  bit 2: distinct smem per slot
__global__ void __launch_bounds__(128, 1) k(int mode, int n_slots_total, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots], empty[kSlots], done;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&done, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  constexpr uint32_t slot_bytes = 40960;
  if (threadIdx.x == 0 && (mode & 2)) {            // producer stand-in
    uint32_t slot = 0, ph = 0;
    for (int s = 0; s < n_slots_total; ++s) {
      mbar_wait(&empty[slot], ph ^ 1);
      mbar_arrive(&full[slot]);
      if (++slot == kSlots) { slot = 0; ph ^= 1; }
    }
  }
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_f16(128, 192, false, false);
    const uint64_t da0 = umma_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t db0 = umma_smem_desc(smem_u32(smem), 3072, 128);
    const long long t0 = clock64();
    uint32_t slot = 0, ph = 0;
    for (int s = 0; s < n_slots_total; ++s) {
      if (mode & 2) { mbar_wait(&full[slot], ph); tc_fence_after(); }
      const uint32_t so = (mode & 4) ? (slot * slot_bytes) >> 4 : 0;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint64_t dah = da0 + so + kk * 256, dbh = db0 + so + (16384 >> 4) + kk * 384;
        umma_f16(tmem, dah, dbh, idesc, (s | kk) != 0);
        umma_f16(tmem, dah + (8192 >> 4), dbh, idesc, 1);
        umma_f16(tmem, dah, dbh + (12288 >> 4), idesc, 1);
      }
      if (mode & 1) umma_commit(&empty[slot]);
      if (++slot == kSlots) { slot = 0; ph ^= 1; }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out; cudaMalloc(&out, 8 * 256);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  const char* names[] = {"MMAs only, same smem", "+ commit per slot", "+ handshake (no commit -> n/a)", "+ commit + producer handshake",
                         "MMAs only, smem per slot", "+ commit per slot", "n/a", "+ commit + producer handshake (lgemm shape)"};
  for (int g : {1, 148})
    for (int mode : {0, 1, 3, 4, 5, 7}) {
      const int n = 240;
      for (int rep = 0; rep < 2; ++rep) k<<<g, 128, 201 * 1024>>>(mode, n, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[256]; cudaMemcpy(h, out, 8 * g, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("CTAs %3d  %-46s %6.1f cycles per MMA (tensor pipe needs 96)\n", g, names[mode], static_cast<double>(mx) / (n * 6));
    }
  return 0;
}
