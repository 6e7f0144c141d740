// Microbenchmark: the layer-GEMM producer's exact copy pattern (per ring slot: A_hi 8 KB + A_lo 8 KB from a private
// stream, B_hi 12 KB + B_lo 12 KB from a 576 KB region that EVERY CTA reads), no MMAs.  Variants isolate the cost of
// the shared-B hot spot and of the copy granularity.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/tma_pattern.cu -o tools/bin/tma_pattern
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;

constexpr int kSlots = 5, kSlotBytes = 40960;

// variant 0: as lgemm (B shared by all CTAs)   1: B private per CTA   2: one 40 KB private copy per slot
// variant 3: as 0 but every CTA starts at a different B stage (staggered)
__global__ void __launch_bounds__(64, 1) k(const uint8_t* a_src, size_t a_bytes, const uint8_t* b_src, int variant, int n_slots, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots];
  if (threadIdx.x == 0) { for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t span = (a_bytes / gridDim.x) & ~static_cast<size_t>(65535);
    const uint8_t* a0 = a_src + blockIdx.x * span;
    const uint8_t* b0 = (variant == 1) ? a0 + span / 2 : b_src;
    const size_t b_span = (variant == 1) ? (span / 2 & ~static_cast<size_t>(32767)) : 589824;    // 576 KB = 24 stages x 24 KB
    size_t ao = 0, bo = (variant == 3) ? (blockIdx.x % 24) * 24576 : 0;
    const long long t0 = clock64();
    uint32_t ph = 0; int slot = 0;
    for (int it = 0; it < n_slots + kSlots; ++it) {
      if (it >= kSlots) mbar_wait(&full[slot], ph);
      if (it < n_slots) {
        uint8_t* dst = smem + slot * kSlotBytes;
        mbar_arrive_expect_tx(&full[slot], kSlotBytes);
        if (variant == 2) {
          bulk_g2s(dst, a0 + ao, kSlotBytes, &full[slot]);
          ao += kSlotBytes; if (ao + kSlotBytes > span) ao = 0;
        } else {
          bulk_g2s(dst, a0 + ao, 8192, &full[slot]);
          bulk_g2s(dst + 16384, b0 + bo, 12288, &full[slot]);
          bulk_g2s(dst + 8192, a0 + ao + 8192, 8192, &full[slot]);
          bulk_g2s(dst + 28672, b0 + bo + 12288, 12288, &full[slot]);
          ao += 16384; if (ao + 16384 > ((variant == 1) ? span / 2 : span)) ao = 0;
          bo += 24576; if (bo + 24576 > b_span) bo = 0;
        }
      }
      if (++slot == kSlots) { slot = 0; if (it >= kSlots) ph ^= 1; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t bytes = 64u << 20;
  uint8_t *a, *b; cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes); cudaMalloc(&b, 1 << 20); cudaMemset(b, 0, 1 << 20);
  long long* out; cudaMalloc(&out, 8 * 256);
  const int smem = kSlots * kSlotBytes + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"lgemm pattern, B shared by all CTAs", "same, B private per CTA", "one private 40 KB copy per slot", "B shared, CTAs staggered over B stages"};
  for (int g : {148, 74})
    for (int v = 0; v < 4; ++v) {
      const int n = 288;   // 11.5 MB per CTA
      for (int rep = 0; rep < 2; ++rep) k<<<g, 64, smem>>>(a, bytes, b, v, n, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[256]; cudaMemcpy(h, out, 8 * g, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("CTAs %3d  %-42s : %6.1f B/cycle/SM\n", g, names[v], static_cast<double>(n) * kSlotBytes / mx);
    }
  return 0;
}
