// Split-K weight/bias gradient GEMMs on tcgen05:  dW_l[out,in] = sum_rows dZ_l[row,out] * H_l[row,in].
//
// K is the batch dimension, so both operands are the activation / gradient images the forward and
// backward kernels saved, consumed as MN-major UMMA operands with no transposition pass.  A work
// item is (unit, split): a 128 x N block of one weight matrix accumulated over a contiguous range
// of row tiles; partial sums go to gpart[split][param] (fixed layout, no atomics) and are added in
// a fixed order by the optimiser kernel, so the whole step is bit-reproducible.  A unit keeps one A sub-image (128
// output features) per row tile and sweeps up to three 128-feature B chunks against it (N up to 384), which cuts the
// L2->SM operand traffic that bounds this kernel by a third compared with square 128 x 128 units.
// Bias gradients ride along as one extra N=16 MMA against an all-ones operand.
// The last (out_f-wide) layer is computed transposed: D[in_chunk, 16] = H^T dZ_last.
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "inr_kernels.cuh"

namespace inr {

#define WG_TRACE(slot) do { if (a.trace) a.trace[64 + blockIdx.x * 8 + (slot)] = global_ns(); } while (0)

constexpr int kWgSlots = 6;                     // ring of 32 KB slots: an A sub-image or one B chunk each
constexpr int kWgSlotBytes = 32768;
constexpr int kWgOnesBytes = 4096;              // all-ones operand: any descriptor stride lands on 1.0
constexpr int kWgThreads = 192;                 // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int kWgSmem = kWgSlots * kWgSlotBytes + kWgOnesBytes + 4 * 32 * 33 * 4 + 1024;
static_assert(128 * (384 * 4 + 16) <= kWgSmem - 1024, "epilogue staging (128 rows x 384 fp32 + 16 B pitch pad) must fit the dead ring");

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  __half* ones = reinterpret_cast<__half*>(smem + kWgSlots * kWgSlotBytes);
  float* tr = reinterpret_cast<float*>(smem + kWgSlots * kWgSlotBytes + kWgOnesBytes);   // 4 x [32][33]
  __shared__ uint64_t full[kWgSlots], empty[kWgSlots], acc_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // items of this CTA: one heavy (unit, split) pair, or up to `group` light ones worked off one after the other
  const int n_hi = a.n_heavy * a.n_split, n_lo = a.n_light * a.n_split;
  int item0, n_my;
  if (static_cast<int>(blockIdx.x) < n_hi) { item0 = blockIdx.x; n_my = 1; }
  else {
    item0 = (blockIdx.x - n_hi) * a.group;
    n_my = n_lo - item0 < a.group ? n_lo - item0 : a.group;
    item0 += n_hi;
  }

  griddep_launch_dependents();      // nothing here is launched as a programmatic dependent of wgrad; harmless
  for (int i = tid; i < kWgOnesBytes / 2; i += kWgThreads) ones[i] = __float2half(1.0f);
  fence_proxy_async_smem();
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) WG_TRACE(0);
  griddep_wait();                   // the last dgrad GEMM (and every kernel before it) complete: the images are final

  for (int my = 0; my < n_my; ++my) {
  const int item = item0 + my;
  const int unit = item < n_hi ? a.order[item % a.n_heavy] : a.order[a.n_heavy + (item - n_hi) % a.n_light];
  const int split = item < n_hi ? item / a.n_heavy : (item - n_hi) / a.n_light;
  const WgradUnit& U = a.u[unit];
  const int nch = U.n_chunks > 1 ? U.n_chunks : 1;
  const int bias_col = 128 * nch;               // accumulator columns of the bias MMA
  const uint32_t nA = nch == 1 ? 3 : 2, nB = kWgSlots - nA;
  const int t0 = static_cast<int>((static_cast<long long>(a.n_tiles) * split) / a.n_split);
  const int t1 = static_cast<int>((static_cast<long long>(a.n_tiles) * (split + 1)) / a.n_split);

  // every item starts from freshly initialised barriers: the previous item's MMAs, commits and bulk copies have all
  // completed (its epilogue waited on acc_full, which is committed last), and the A / B sub-ring partition may change
  if (tid == 0) {
    for (int i = 0; i < kWgSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0 && my == 0) WG_TRACE(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: per tile  A, B_0 .. B_{nch-1}
    // A sub-images and B chunks live in separate sub-rings (slots [0,nA) and [nA,6)) so that the A slot a unit holds
    // for all of its chunks never blocks the in-order prefetch of the following B chunks.
    if (lane == 0) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      const uint64_t pol_first = l2_policy_evict_first();
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&empty[as], aph ^ 1);
        mbar_arrive_expect_tx(&full[as], U.a_bytes);
        if (a.l2_hints & 1) bulk_g2s_hint(ring + as * kWgSlotBytes, a.ws + U.a_off + static_cast<size_t>(t) * U.a_tile_stride + U.a_sub, U.a_bytes, &full[as], pol_first);
        else bulk_g2s(ring + as * kWgSlotBytes, a.ws + U.a_off + static_cast<size_t>(t) * U.a_tile_stride + U.a_sub, U.a_bytes, &full[as]);
        if (++as == nA) { as = 0; aph ^= 1; }
        for (int c = 0; c < nch; ++c) {
          const uint32_t slot = nA + bs;
          mbar_wait(&empty[slot], bph ^ 1);
          mbar_arrive_expect_tx(&full[slot], U.b_bytes);
          const uint8_t* bsrc = a.ws + U.b_off + static_cast<size_t>(t) * U.b_tile_stride + U.b_sub + static_cast<size_t>(c) * U.b_bytes;
          if (a.l2_hints & 2) bulk_g2s_hint(ring + slot * kWgSlotBytes, bsrc, U.b_bytes, &full[slot], pol_first);
          else bulk_g2s(ring + slot * kWgSlotBytes, bsrc, U.b_bytes, &full[slot]);
          if (++bs == nB) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the ring and one elected lane issues: the descriptors stay warp-uniform (uniform registers).  A
    // single divergent thread paid ~280 cycles per slot + ~50 per MMA (tools/umma_commit2.cu), more than the 512 tensor
    // cycles of the eight N = 128 MMAs of a chunk.
    {
      const uint32_t idesc = umma_idesc_f16(kTileM, U.n, true, true);
      constexpr uint32_t idesc_bias = umma_idesc_f16(kTileM, 16, true, true);
      const uint64_t d_ones = umma_smem_desc(smem_u32(ones), 128, 128);
      const int ksteps = a.tile_rows ? a.tile_rows / 16 : kTileM / 16;   // K = 16 batch rows per MMA
      const uint64_t d0 = umma_smem_desc(smem_u32(ring), 128, a.lb ? a.lb : 2048);   // MN-major image: LBO 128 (K groups), SBO = k-group stride
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, first = 1;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&full[as], aph);
        tc_fence_after();
        if (t == t0 && lane == 0) WG_TRACE(2);
        const uint64_t da0 = d0 + ((as * kWgSlotBytes) >> 4);
        for (int c = 0; c < nch; ++c) {
          const uint32_t slot = nA + bs;
          mbar_wait(&full[slot], bph);
          tc_fence_after();
          __syncwarp();
          const uint64_t db0 = d0 + ((slot * kWgSlotBytes) >> 4);
          if (elect_one()) {
            if (ksteps == kTileM / 16) {
#pragma unroll
              for (int k = 0; k < kTileM / 16; ++k)        // K = 16 batch rows per MMA: 256 B of every feature group
                umma_f16(tmem + 128 * c, da0 + k * 16, db0 + k * 16, idesc, (first ^ 1) | (k != 0));
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_f16(tmem + 128 * c, da0 + k * 16, db0 + k * 16, idesc, (first ^ 1) | (k != 0));
            }
            if (U.bias_off >= 0 && c == nch - 1) {
              // normal unit: bias[o] = sum_rows dZ[row,o] * 1  -> A = dZ image, B = ones
              // transposed unit: bias[o] = sum_rows 1 * dZ_last[row,o] -> A = ones, B = dZ_last image
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t da = U.transposed ? d_ones : da0 + k * 16;
                const uint64_t db = U.transposed ? db0 + k * 16 : d_ones;
                umma_f16(tmem + bias_col, da, db, idesc_bias, (first ^ 1) | (k != 0));
              }
            }
            umma_commit(&empty[slot]);                       // commits cover every MMA issued so far
            if (c == nch - 1) umma_commit(&empty[as]);
          }
          __syncwarp();
          if (++bs == nB) { bs = 0; bph ^= 1; }
        }
        if (++as == nA) { as = 0; aph ^= 1; }
        first = 0;
      }
      __syncwarp();
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
      if (lane == 0) WG_TRACE(3);
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
    if (tid == 64) WG_TRACE(4);
    if (!U.transposed) {
      const int n_cols = nch > 1 ? 128 * nch : U.n;
      // Fast path: the operand ring is dead once acc_full fired, so the whole 128 x n_cols fp32 block is staged there
      // (row pitch +16 B: conflict-free 16 B stores) and every thread bulk-stores its own row -- no transposition pass.
      const uint32_t row_bytes = static_cast<uint32_t>(n_cols) * 4, pitch = row_bytes + 16;
      float* dst_row = gp + U.out_off + static_cast<size_t>(U.row0 + r_local) * U.out_ld;
      const int seg_cols = U.perm_e > 0 ? n_cols / 2 : n_cols;               // gauss first layer: two destination runs
      const int dcol0 = U.perm_e > 0 ? 32 * (U.col0 >> 6) : U.col0;
      const bool fast = have && n_cols >= 128 && U.rows_valid == 128 && U.cols_valid == n_cols && (U.out_ld & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(dst_row + dcol0) | reinterpret_cast<uintptr_t>(dst_row + U.perm_e + dcol0)) & 15) == 0 &&
                        (U.perm_e == 0 || (U.col0 & 63) == 0);
      if (fast) {
        uint8_t* srow = smem + static_cast<size_t>(r_local) * pitch;
        for (int c0 = 0; c0 < n_cols; c0 += 32) {
          float v[32];
          tmem_ld32(tmem + t_lane + c0, v);
          tmem_ld_wait();
          int sc = c0;                                                      // staging column of this 32-column chunk
          if (U.perm_e > 0) sc = ((c0 & 32) ? seg_cols : 0) + 32 * (c0 >> 6);
          float4* d = reinterpret_cast<float4*>(srow + sc * 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        fence_proxy_async_smem();
        bulk_s2g(dst_row + dcol0, srow, seg_cols * 4);
        if (U.perm_e > 0) bulk_s2g(dst_row + U.perm_e + dcol0, srow + seg_cols * 4, seg_cols * 4);
        bulk_commit();
        if (U.bias_off >= 0) {
          float v[16];
          tmem_ld16(tmem + t_lane + bias_col, v); tmem_ld_wait();
          gp[U.bias_off + U.row0 + r_local] = v[0];
        }
        bulk_wait_read0();
      } else {
      for (int c0 = 0; c0 < n_cols; c0 += 32) {
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
        if (have) { tmem_ld16(tmem + t_lane + bias_col, v); tmem_ld_wait(); } else v[0] = 0.f;
        if (r_local < U.rows_valid) gp[U.bias_off + U.row0 + r_local] = v[0];
      }
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
        if (have) { tmem_ld16(tmem + t_lane + bias_col, b); tmem_ld_wait(); }
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
  if (tid == 64) WG_TRACE(5);
  tc_fence_before();
  __syncthreads();            // item boundary: ring, barriers and accumulator are free again
  tc_fence_after();
  }  // items
  if (warp == 2) tmem_dealloc<512>(tmem_base_s);
  if (tid == 0) WG_TRACE(6);
}

cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t stream) {
  if (a.n_heavy + a.n_light != a.n_units || a.group < 1) return cudaErrorInvalidValue;
  const int grid = a.n_heavy * a.n_split + (a.n_light * a.n_split + a.group - 1) / a.group;
  if (grid <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  WgradArgs b = a;
  const char* env = std::getenv("INR_WGRAD_L2");       // overrides the caller's choice per launch (A/B runs)
  if (env) b.l2_hints = std::atoi(env);
  return launch_dependent(wgrad_kernel, dim3(grid), dim3(kWgThreads), kWgSmem, stream, b);
}

}  // namespace inr
