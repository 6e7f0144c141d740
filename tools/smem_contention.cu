// Microbenchmark: do bulk-TMA writes into shared memory and tcgen05.mma operand reads from shared memory contend?
// One CTA per SM; thread 0 streams L2-resident data into a 4-slot ring, thread 32 issues back-to-back MMAs
// (M=128, K=16, K-major, no swizzle) on a separate operand region.  Each is timed alone and together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/smem_contention.cu -o tools/bin/smem_contention
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

constexpr int kSlots = 4, kSlotBytes = 32768, kOperandBytes = 65536;

__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, size_t src_bytes, uint32_t copy_bytes, int n_copies, int n, int reps,
                                          int cta_pairs_note, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots], done;
  __shared__ uint32_t tmem_s;
  uint8_t* ring = smem + kOperandBytes;
  for (int i = threadIdx.x; i < kOperandBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1); mbar_init(&done, 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0 && n_copies > 0) {
    const int per_slot = kSlotBytes / copy_bytes;
    const size_t span = src_bytes / gridDim.x & ~static_cast<size_t>(kSlotBytes - 1);
    const uint8_t* base = src + blockIdx.x * span;
    const long long t0 = clock64();
    uint32_t ph = 0; int slot = 0; size_t off = 0;
    for (int it = 0; it < n_copies + kSlots; ++it) {
      if (it >= kSlots) mbar_wait(&full[slot], ph);
      if (it < n_copies) {
        mbar_arrive_expect_tx(&full[slot], per_slot * copy_bytes);
        for (int j = 0; j < per_slot; ++j) bulk_g2s(ring + slot * kSlotBytes + j * copy_bytes, base + off + j * copy_bytes, copy_bytes, &full[slot]);
        off += kSlotBytes; if (off + kSlotBytes > span) off = 0;
      }
      if (++slot == kSlots) { slot = 0; if (it >= kSlots) ph ^= 1; }
    }
    out[2 * blockIdx.x] = clock64() - t0;
  }
  if (threadIdx.x == 32 && reps > 0) {
    const uint32_t tmem = tmem_s;
    const uint32_t idesc = umma_idesc_f16(128, n, false, false);
    const uint64_t da = umma_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t db = umma_smem_desc(smem_u32(smem) + 16384, n * 16, 128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem + (kk & 1) * 256, da + kk * 256, db + ((kk * n * 32) >> 4), idesc, 1);
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    out[2 * blockIdx.x + 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_s);
}

int main() {
  const size_t bytes = 48u << 20;
  uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 0, bytes);
  long long* out; cudaMalloc(&out, 16 * 256);
  const int smem = kOperandBytes + kSlots * kSlotBytes + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int g = 148;
  for (int n : {128, 192, 256})
    for (uint32_t cb : {32768u, 8192u}) {
      for (int mode = 0; mode < 3; ++mode) {   // 0 TMA alone, 1 MMA alone, 2 both
        const int n_copies = mode == 1 ? 0 : 256;          // 8 MiB per CTA
        // size the MMA loop to last about as long as the copy stream
        const int reps = mode == 0 ? 0 : 256 * 32768 / 60 / (n / 2) / 4 * 4;
        cudaMemset(out, 0, 16 * 256);
        for (int rep = 0; rep < 2; ++rep) k<<<g, 64, smem>>>(src, bytes, cb, n_copies, n, reps, 0, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
        long long h[512]; cudaMemcpy(h, out, 16 * g, cudaMemcpyDeviceToHost);
        long long mt = 0, mm = 0; for (int i = 0; i < g; ++i) { mt = h[2 * i] > mt ? h[2 * i] : mt; mm = h[2 * i + 1] > mm ? h[2 * i + 1] : mm; }
        printf("N=%3d copy %5u B  %-9s :", n, cb, mode == 0 ? "TMA alone" : (mode == 1 ? "MMA alone" : "both"));
        if (n_copies) printf("  TMA %6.1f B/cycle/SM", 256.0 * 32768 / mt);
        if (reps) printf("  MMA %6.1f cycles/instr (ideal %d; operand reads %5.1f B/cycle)", static_cast<double>(mm) / reps, n / 2,
                         (4096.0 + n * 32.0) / (static_cast<double>(mm) / reps));
        printf("\n");
      }
    }
  return 0;
}
