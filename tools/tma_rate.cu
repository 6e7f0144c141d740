// Microbenchmark: sustained L2 -> shared-memory ingest rate of 1-D bulk TMA copies per SM, as a function of the number
// of CTAs pulling at once and of the copy size.  The source (48 MiB) is L2 resident after the warm-up pass.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/tma_rate.cu -o tools/bin/tma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

constexpr int kSlots = 6, kSlotBytes = 32768;

__global__ void __launch_bounds__(64, 1) pull_kernel(const uint8_t* src, size_t src_bytes, uint32_t copy_bytes, int n_copies, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots];
  if (threadIdx.x == 0) { for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int per_slot = kSlotBytes / copy_bytes;
    const size_t span = src_bytes / gridDim.x & ~static_cast<size_t>(kSlotBytes - 1);
    const uint8_t* base = src + blockIdx.x * span;
    const long long t0 = clock64();
    uint32_t ph = 0; int slot = 0; size_t off = 0;
    // keep kSlots slots in flight: issue slot i, wait for the oldest
    for (int it = 0; it < n_copies + kSlots; ++it) {
      if (it >= kSlots) { mbar_wait(&full[slot], ph); }      // oldest copy of this slot landed
      if (it < n_copies) {
        mbar_arrive_expect_tx(&full[slot], per_slot * copy_bytes);
        for (int j = 0; j < per_slot; ++j) bulk_g2s(smem + slot * kSlotBytes + j * copy_bytes, base + off + j * copy_bytes, copy_bytes, &full[slot]);
        off += kSlotBytes; if (off + kSlotBytes > span) off = 0;
      }
      if (++slot == kSlots) { slot = 0; if (it >= kSlots) ph ^= 1; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t bytes = 48u << 20;
  uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
  long long* out; cudaMalloc(&out, 8 * 256);
  cudaFuncSetAttribute(pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kSlotBytes + 1024);
  const int grids[] = {1, 8, 37, 74, 148};
  const uint32_t sizes[] = {32768, 8192, 2048};
  for (uint32_t cb : sizes)
    for (int g : grids) {
      const int n = 512;   // 16 MiB per CTA
      for (int rep = 0; rep < 2; ++rep) pull_kernel<<<g, 64, kSlots * kSlotBytes + 1024>>>(src, bytes, cb, n, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[256]; cudaMemcpy(h, out, 8 * g, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bpc = static_cast<double>(n) * kSlotBytes / mx;
      printf("copy %6u B  CTAs %3d : %6.1f B/cycle/SM   %7.0f B/cycle chip\n", cb, g, bpc, bpc * g);
    }
  return 0;
}
