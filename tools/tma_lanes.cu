// Microbenchmark: is the per-copy cost of small bulk TMA copies paid by the issuing thread or by the copy engine?
// A 32 KB slot is filled by `per_slot` copies, issued either all by lane 0 or one per lane by `per_slot` lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mri_implicit_neural_representations_b200/csrc tools/tma_lanes.cu -o tools/bin/tma_lanes
#include <cstdio>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
using namespace inr;
constexpr int kSlots = 6, kSlotBytes = 32768;

__global__ void __launch_bounds__(32, 1) k(const uint8_t* src, size_t src_bytes, int per_slot, int multi_lane, int n_slots, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots];
  const int lane = threadIdx.x;
  if (lane == 0) { for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
  __syncwarp();
  const uint32_t cb = kSlotBytes / per_slot;
  const size_t span = (src_bytes / gridDim.x) & ~static_cast<size_t>(kSlotBytes - 1);
  const uint8_t* base = src + blockIdx.x * span;
  const long long t0 = clock64();
  uint32_t ph = 0; int slot = 0; size_t off = 0;
  for (int it = 0; it < n_slots + kSlots; ++it) {
    if (it >= kSlots) mbar_wait(&full[slot], ph);          // all lanes poll
    if (it < n_slots) {
      if (lane == 0) mbar_arrive_expect_tx(&full[slot], kSlotBytes);
      __syncwarp();
      if (multi_lane) {
        if (lane < per_slot) bulk_g2s(smem + slot * kSlotBytes + lane * cb, base + off + lane * cb, cb, &full[slot]);
      } else if (lane == 0) {
        for (int j = 0; j < per_slot; ++j) bulk_g2s(smem + slot * kSlotBytes + j * cb, base + off + j * cb, cb, &full[slot]);
      }
      off += kSlotBytes; if (off + kSlotBytes > span) off = 0;
    }
    if (++slot == kSlots) { slot = 0; if (it >= kSlots) ph ^= 1; }
  }
  if (lane == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const size_t bytes = 48u << 20;
  uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 0, bytes);
  long long* out; cudaMalloc(&out, 8 * 256);
  const int smem = kSlots * kSlotBytes + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int g : {1, 148})
    for (int per : {1, 4, 16})
      for (int ml : {0, 1}) {
        if (per == 1 && ml) continue;
        for (int rep = 0; rep < 2; ++rep) k<<<g, 32, smem>>>(src, bytes, per, ml, 512, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
        long long h[256]; cudaMemcpy(h, out, 8 * g, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("CTAs %3d  %2d copies of %5d B per slot, issued by %-9s : %6.1f B/cycle/SM\n", g, per, kSlotBytes / per, ml ? "one lane each" : "lane 0", 512.0 * kSlotBytes / mx);
      }
  return 0;
}
