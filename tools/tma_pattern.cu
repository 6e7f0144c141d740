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
// variant 4: as 0, plus 512 threads per CTA storing 16 KB per slot to global memory (the forward epilogue's write traffic)
// variant 6: as 4 but the stores cycle over 29 MB (L2 resident) instead of 155 MB (thrashes the 126 MB L2)
// variant 5: as 0, plus 512 threads per CTA loading 16 KB per slot with ld.global (the dgrad epilogue's read traffic)
__global__ void __launch_bounds__(576, 1) k(const uint8_t* a_src, size_t a_bytes, const uint8_t* b_src, int variant, int n_slots, long long* out, uint8_t* wdst) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots];
  __shared__ volatile int progress;
  if (threadIdx.x == 0) { progress = 0; for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x >= 64 && variant >= 4) {
    // side traffic per slot period: 16 KB per CTA, paced by the producer's progress counter in shared memory.
    // Every warp instruction touches 512 contiguous bytes (lane * 16), as the kernels' epilogues do.
    const int t = threadIdx.x - 64, wid = t >> 5, lane = t & 31;
    uint8_t* mine = wdst + static_cast<size_t>(blockIdx.x) * 16384 + wid * 1024 + lane * 16;
    const size_t stride = static_cast<size_t>(gridDim.x) * 16384;
    uint8_t* stage = smem + kSlots * kSlotBytes;                 // 16 KB staging for the bulk-store variant
    uint4 acc = make_uint4(t, 1, 2, 3);
    for (int it = 0; it < n_slots; ++it) {
      uint8_t* p = mine + (static_cast<size_t>(it) % (variant == 6 ? 12 : 64)) * stride;
      if (variant == 4 || variant == 6) { st_global_v4(p, acc); st_global_v4(p + 512, acc); }
      else if (variant == 5) { const uint4 v = ld_global_nc_v4(p), w = ld_global_nc_v4(p + 512); acc.x += v.x + w.y; }
      else if (variant == 7) {       // stage in shared memory, one 2 KB bulk store per pair of warps
        *reinterpret_cast<uint4*>(stage + wid * 1024 + lane * 16) = acc;
        *reinterpret_cast<uint4*>(stage + wid * 1024 + 512 + lane * 16) = acc;
        fence_proxy_async_smem();
        named_bar_sync(1 + (wid >> 1), 64);
        if ((t & 63) == 0) {
          bulk_s2g(wdst + static_cast<size_t>(blockIdx.x) * 16384 + (wid >> 1) * 2048 + (static_cast<size_t>(it) % 12) * stride,
                   stage + (wid >> 1) * 2048, 2048);
          bulk_commit();
          bulk_wait_read0();
        }
        named_bar_sync(1 + (wid >> 1), 64);
      }
      while (progress < it - 2) {}
    }
    if (acc.x == 0x12345678u) out[200] = acc.x;
  }
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
      progress = it;
    }
    progress = 1 << 30;
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t bytes = 64u << 20;
  uint8_t *a, *b; cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes); cudaMalloc(&b, 1 << 20); cudaMemset(b, 0, 1 << 20);
  long long* out; cudaMalloc(&out, 8 * 256);
  uint8_t* wdst; cudaMalloc(&wdst, static_cast<size_t>(64) * 148 * 512 * 32); cudaMemset(wdst, 0, static_cast<size_t>(64) * 148 * 512 * 32);
  const int smem = kSlots * kSlotBytes + 16384 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"lgemm pattern, B shared by all CTAs", "same, B private per CTA", "one private 40 KB copy per slot", "B shared, CTAs staggered over B stages",
                         "lgemm pattern + 16 KB of st.global per slot", "lgemm pattern + 16 KB of ld.global per slot",
                         "lgemm pattern + 16 KB of st.global per slot, 29 MB target",
                         "lgemm pattern + 16 KB per slot via smem + 2 KB bulk stores, 29 MB target"};
  for (int g : {148, 74})
    for (int v = 0; v < 8; ++v) {
      const int n = 288;   // 11.5 MB per CTA
      for (int rep = 0; rep < 2; ++rep) k<<<g, 576, smem>>>(a, bytes, b, v, n, out, wdst);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[256]; cudaMemcpy(h, out, 8 * g, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("CTAs %3d  %-42s : %6.1f B/cycle/SM\n", g, names[v], static_cast<double>(n) * kSlotBytes / mx);
    }
  return 0;
}
