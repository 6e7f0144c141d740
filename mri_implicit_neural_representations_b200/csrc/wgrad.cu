// Split-K weight/bias gradient GEMMs on tcgen05:  dW_l[out,in] = sum_rows dZ_l[row,out] * H_l[row,in].
//
// K is the batch dimension, so both operands are the activation / gradient images the forward and
// backward kernels saved, consumed as MN-major UMMA operands with no transposition pass.  A work
// item is (unit, split): a 128 x N block of one weight matrix accumulated over a contiguous range
// of row tiles; partial sums go to gpart[split][param] (fixed layout, no atomics) and are added in
// a fixed order by the optimiser kernel, so the whole step is bit-reproducible.
// Bias gradients ride along as one extra N=16 MMA against an all-ones operand.
// The last (out_f-wide) layer is computed transposed: D[in_chunk, 16] = H^T dZ_last.
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "inr_kernels.cuh"

namespace inr {

constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;            // A sub-image (32 KB) + B sub-image (<= 32 KB)
constexpr int kWgOnesBytes = 4096;         // all-ones operand: any descriptor stride lands on 1.0
constexpr int kWgThreads = 192;                 // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int kWgSmem = kWgStages * kWgStageBytes + kWgOnesBytes + 4 * 32 * 33 * 4 + 1024;

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  __half* ones = reinterpret_cast<__half*>(smem + kWgStages * kWgStageBytes);
  float* tr = reinterpret_cast<float*>(smem + kWgStages * kWgStageBytes + kWgOnesBytes);   // 4 x [32][33]
  __shared__ uint64_t full[kWgStages], empty[kWgStages], acc_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = blockIdx.x % a.n_units, split = blockIdx.x / a.n_units;
  const WgradUnit& U = a.u[unit];
  const int t0 = static_cast<int>((static_cast<long long>(a.n_tiles) * split) / a.n_split);
  const int t1 = static_cast<int>((static_cast<long long>(a.n_tiles) * (split + 1)) / a.n_split);

  if (tid == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < kWgOnesBytes / 2; i += kWgThreads) ones[i] = __float2half(1.0f);
  fence_proxy_async_smem();
  if (warp == 2) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t slot = it % kWgStages, ph = (it / kWgStages) & 1;
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], U.a_bytes + U.b_bytes);
        uint8_t* dst = stages + slot * kWgStageBytes;
        bulk_g2s(dst, a.ws + U.a_off + static_cast<size_t>(t) * U.a_tile_stride + U.a_sub, U.a_bytes, &full[slot]);
        bulk_g2s(dst + 32768, a.ws + U.b_off + static_cast<size_t>(t) * U.b_tile_stride + U.b_sub, U.b_bytes, &full[slot]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kTileM, U.n, true, true);
      constexpr uint32_t idesc_bias = umma_idesc_f16(kTileM, 16, true, true);
      const uint32_t ones_s = smem_u32(ones);
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t slot = it % kWgStages;
        mbar_wait(&full[slot], (it / kWgStages) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(stages + slot * kWgStageBytes), b_base = a_base + 32768;
#pragma unroll
        for (int k = 0; k < kTileM / 16; ++k) {
          const uint64_t da = umma_smem_desc(a_base + k * 256, 128, 2048);
          const uint64_t db = umma_smem_desc(b_base + k * 256, 128, 2048);
          umma_f16(tmem, da, db, idesc, (it | k) != 0);
        }
        if (U.bias_off >= 0) {
          // normal unit: bias[o] = sum_rows dZ[row,o] * 1  -> A = dZ image, B = ones, accumulator columns 128..143
          // transposed unit: bias[o] = sum_rows 1 * dZ_last[row,o] -> A = ones, B = dZ_last image
#pragma unroll
          for (int k = 0; k < kTileM / 16; ++k) {
            const uint64_t d_ones = umma_smem_desc(ones_s, 128, 128);
            const uint64_t da = U.transposed ? d_ones : umma_smem_desc(a_base + k * 256, 128, 2048);
            const uint64_t db = U.transposed ? umma_smem_desc(b_base + k * 256, 128, 2048) : d_ones;
            umma_f16(tmem + 128, da, db, idesc_bias, (it | k) != 0);
          }
        }
        umma_commit(&empty[slot]);
      }
      umma_commit(&acc_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> gpart[split]
    const int q = warp & 3;                       // warps 2,3,4,5 -> lane quarters 2,3,0,1
    const int r_local = q * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    float* gp = reinterpret_cast<float*>(a.ws + a.gpart_off) + static_cast<size_t>(split) * a.n_params;
    float* my_tr = tr + (warp - 2) * 32 * 33;
    const bool have = t1 > t0;
    if (have) { mbar_wait(&acc_full, 0); tc_fence_after(); }
    if (!U.transposed) {
      for (int c0 = 0; c0 < U.n; c0 += 32) {
        float v[32];
        if (have) { tmem_ld32(tmem + t_lane + c0, v); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) my_tr[lane * 33 + i] = v[i];
        __syncwarp();
        // 32 consecutive columns of one weight row per store instruction (128 B, coalesced)
        int col = U.col0 + c0 + lane;
        if (U.perm_e > 0) {     // sin/cos-interleaved chunk order -> reference column order
          const int kp = col, c64 = kp >> 6, kk = kp & 63;
          col = (kk < 32) ? (32 * c64 + kk) : (U.perm_e + 32 * c64 + kk - 32);
        }
        const bool col_ok = (c0 + lane) < U.cols_valid;
        for (int rr = 0; rr < 32; ++rr) {
          const int r = q * 32 + rr;
          if (r < U.rows_valid && col_ok)
            gp[U.out_off + static_cast<size_t>(U.row0 + r) * U.out_ld + col] = my_tr[rr * 33 + lane];
        }
        __syncwarp();
      }
      if (U.bias_off >= 0) {
        float v[16];
        if (have) { tmem_ld16(tmem + t_lane + 128, v); tmem_ld_wait(); } else v[0] = 0.f;
        if (r_local < U.rows_valid) gp[U.bias_off + U.row0 + r_local] = v[0];
      }
    } else {
      // D[i_local, o] -> dW[o][col0 + i_local]; lanes write consecutive i (coalesced)
      float v[16];
      if (have) { tmem_ld16(tmem + t_lane, v); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < U.rows_valid && r_local < U.cols_valid) gp[U.out_off + o * U.out_ld + U.col0 + r_local] = v[o];
      if (U.bias_off >= 0) {
        float b[16];
        if (have) { tmem_ld16(tmem + t_lane + 128, b); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) b[i] = 0.f;
        }
        if (r_local == 0) {
#pragma unroll
          for (int o = 0; o < kMaxOut; ++o)
            if (o < U.rows_valid) gp[U.bias_off + o] = b[o];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem);
}

cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t stream) {
  const int grid = a.n_units * a.n_split;
  if (grid <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  wgrad_kernel<<<grid, kWgThreads, kWgSmem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace inr
