// Streaming layer GEMM with fused epilogues ("lgemm"): D[128 x 192] = sum_passes A[128 x 384] * B[192 x 384]^T per work
// item (row tile, N-block), A and B both streamed from HBM operand images through a 5-stage bulk-TMA ring,
// tcgen05.mma kind::f16 with fp32 accumulation in TMEM, two accumulators ping-ponged between work items so the
// epilogue of item i overlaps the MMAs of item i+1.  Persistent CTAs (one per SM) walk the items.
//
// Epilogues:
//   LG_WIRE_FWD   : complex Gabor wavelet  y = exp(j w z - |s z|^2),  z = a + jb = acc + bias
//                   (reference src/models/networks.py:199-204), written as the fp16 hi/lo operand images of the next
//                   layer plus the fp16 (a|b) image the backward pass needs.  3-pass split GEMM.
//   LG_WIRE_DGRAD : dL/d(a,b) of the previous layer from dL/dh = acc:
//                   P = Re(conj(g) y), Q = Im(conj(g) y);  dza = -2 s^2 a P - w Q;  dzb = -(w + 2 s^2 b) P.  1-pass GEMM.
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "wire.cuh"

namespace inr {

constexpr int kLgStages = 5;
constexpr int kLgStageBytes = 2 * kWStageABytes + 2 * kWStageBBytes;   // 40960
constexpr int kLgComputeThreads = 512;
constexpr int kLgThreads = 128 + kLgComputeThreads;
constexpr int kLgSmem = kLgStages * kLgStageBytes + 1024;
constexpr int kLgKStages = kW2 / kStageK;                              // 12

__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

__global__ void __launch_bounds__(kLgThreads, 1) lgemm_kernel(const __grid_constant__ LGemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kLgStages], empty[kLgStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_ba[kWP], s_bb[kWP];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_items = a.n_tiles * a.n_nblocks;

  if (tid == 0) {
    for (int i = 0; i < kLgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kLgComputeThreads); }
    mbar_fence_init();
  }
  if (a.mode == LG_WIRE_FWD) {
    for (int j = tid; j < kWP; j += kLgThreads) {
      s_ba[j] = j < a.c_valid ? a.bias[2 * j] : 0.f;
      s_bb[j] = j < a.c_valid ? a.bias[2 * j + 1] : 0.f;
    }
  }
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t stage_tx = a.passes == 3 ? kLgStageBytes : (kWStageABytes + kWStageBBytes);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / a.n_nblocks, nb = item % a.n_nblocks;
        const size_t a_off = static_cast<size_t>(tile) * kWTileBytes;
        const size_t b_off = static_cast<size_t>(nb) * (kLgKStages * kWStageBBytes);
        for (int s = 0; s < kLgKStages; ++s, ++it) {
          const uint32_t slot = it % kLgStages, ph = (it / kLgStages) & 1;
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&full[slot], stage_tx);
          uint8_t* dst = smem + slot * kLgStageBytes;
          bulk_g2s(dst, a.a_hi + a_off + static_cast<size_t>(s) * kWStageABytes, kWStageABytes, &full[slot]);
          bulk_g2s(dst + 2 * kWStageABytes, a.b_hi + b_off + static_cast<size_t>(s) * kWStageBBytes, kWStageBBytes, &full[slot]);
          if (a.passes == 3) {
            bulk_g2s(dst + kWStageABytes, a.a_lo + a_off + static_cast<size_t>(s) * kWStageABytes, kWStageABytes, &full[slot]);
            bulk_g2s(dst + 2 * kWStageABytes + kWStageBBytes, a.b_lo + b_off + static_cast<size_t>(s) * kWStageBBytes,
                     kWStageBBytes, &full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, kWNT, false, false);
      uint32_t it = 0, n_done = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const uint32_t ab = n_done & 1, use = n_done >> 1;
        mbar_wait(&acc_empty[ab], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem + ab * 256;
        for (int s = 0; s < kLgKStages; ++s, ++it) {
          const uint32_t slot = it % kLgStages;
          mbar_wait(&full[slot], (it / kLgStages) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + slot * kLgStageBytes);
          const uint32_t a_hi = base, a_lo = base + kWStageABytes;
          const uint32_t b_hi = base + 2 * kWStageABytes, b_lo = b_hi + kWStageBBytes;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint64_t dah = umma_smem_desc(a_hi + kk * 4096, 2048, 128);
            const uint64_t dbh = umma_smem_desc(b_hi + kk * 6144, 3072, 128);
            umma_f16(acc, dah, dbh, idesc, (s | kk) != 0);
            if (a.passes == 3) {
              const uint64_t dal = umma_smem_desc(a_lo + kk * 4096, 2048, 128);
              const uint64_t dbl = umma_smem_desc(b_lo + kk * 6144, 3072, 128);
              umma_f16(acc, dal, dbh, idesc, 1);
              umma_f16(acc, dah, dbl, idesc, 1);
            }
          }
          umma_commit(&empty[slot]);
        }
        umma_commit(&acc_full[ab]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps (16)
    const int q = warp & 3, sub = (warp - 4) >> 2, row = q * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    const float w = a.omega, s2 = a.sigma * a.sigma;
    // per-layer power-of-two gradient scales: WIRE's gradient norm grows ~10x per layer towards the input, one global
    // loss scale would saturate the fp16 dZ images of the lower layers
    float ratio = 1.f, amax = 0.f;
    if (a.mode == LG_WIRE_DGRAD) ratio = a.scal[SC_LAYER_SCALE + a.dst_layer] / a.scal[SC_LAYER_SCALE + a.src_layer];
    uint32_t n_done = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
      const int tile = item / a.n_nblocks, nb = item % a.n_nblocks;
      const uint32_t ab = n_done & 1, use = n_done >> 1;
      const size_t img = static_cast<size_t>(tile) * kWTileBytes + row * 16;
      mbar_wait(&acc_full[ab], use & 1);
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        const int c0 = 24 * sub + 8 * i;                 // feature inside the N-block
        const int f0 = kWFeatPerBlock * nb + c0;         // complex feature index (multiple of 8)
        const size_t off_r = img + static_cast<size_t>(f0 >> 3) * 2048;               // real-part k-group
        const size_t off_i = img + static_cast<size_t>((kWP + f0) >> 3) * 2048;       // imaginary-part k-group
        float va[8], vb[8];
        if (a.mode == LG_WIRE_FWD) {
          tmem_ld8(tmem + t_lane + ab * 256 + c0, va);
          tmem_ld8(tmem + t_lane + ab * 256 + kWFeatPerBlock + c0, vb);
          tmem_ld_wait();
          float yr[8], yi[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float za = va[e] + s_ba[f0 + e], zb = vb[e] + s_bb[f0 + e];
            va[e] = za; vb[e] = zb;
            const float mag = __expf(-w * zb - s2 * (za * za + zb * zb));
            const float ang = w * za;
            const bool live = (f0 + e) < a.c_valid;
            yr[e] = live ? mag * fast_cos(ang) : 0.f;
            yi[e] = live ? mag * fast_sin(ang) : 0.f;
          }
          uint4 rh, rl, ih, il;
          split_h2(yr[0], yr[1], rh.x, rl.x); split_h2(yr[2], yr[3], rh.y, rl.y);
          split_h2(yr[4], yr[5], rh.z, rl.z); split_h2(yr[6], yr[7], rh.w, rl.w);
          split_h2(yi[0], yi[1], ih.x, il.x); split_h2(yi[2], yi[3], ih.y, il.y);
          split_h2(yi[4], yi[5], ih.z, il.z); split_h2(yi[6], yi[7], ih.w, il.w);
          st_global_v4(a.out_hi + off_r, rh); st_global_v4(a.out_hi + off_i, ih);
          st_global_v4(a.out_lo + off_r, rl); st_global_v4(a.out_lo + off_i, il);
          if (a.train) {
            st_global_v4(a.out_ab + off_r, make_uint4(pack_h2(va[0], va[1]), pack_h2(va[2], va[3]), pack_h2(va[4], va[5]), pack_h2(va[6], va[7])));
            st_global_v4(a.out_ab + off_i, make_uint4(pack_h2(vb[0], vb[1]), pack_h2(vb[2], vb[3]), pack_h2(vb[4], vb[5]), pack_h2(vb[6], vb[7])));
          }
        } else {
          const uint4 yr4 = ld_global_nc_v4(a.in_y + off_r), yi4 = ld_global_nc_v4(a.in_y + off_i);
          const uint4 a4 = ld_global_nc_v4(a.in_ab + off_r);
          uint4 b4 = make_uint4(0u, 0u, 0u, 0u);
          if (!a.real_first) b4 = ld_global_nc_v4(a.in_ab + off_i);
          tmem_ld8(tmem + t_lane + ab * 256 + c0, va);                       // dL/d Re(h)
          tmem_ld8(tmem + t_lane + ab * 256 + kWFeatPerBlock + c0, vb);      // dL/d Im(h)
          tmem_ld_wait();
          float yr[8], yi[8], za[8], zb[8], da[8], db[8];
          unpack8(yr4, yr); unpack8(yi4, yi); unpack8(a4, za); unpack8(b4, zb);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float P = va[e] * yr[e] + vb[e] * yi[e];
            const float Q = va[e] * yi[e] - vb[e] * yr[e];
            da[e] = ratio * (-2.f * s2 * za[e] * P - w * Q);
            db[e] = a.real_first ? 0.f : ratio * (-(w + 2.f * s2 * zb[e]) * P);
            amax = fmaxf(amax, fmaxf(fabsf(da[e]), fabsf(db[e])));
          }
          st_global_v4(a.out_dz + off_r, make_uint4(pack_h2(da[0], da[1]), pack_h2(da[2], da[3]), pack_h2(da[4], da[5]), pack_h2(da[6], da[7])));
          st_global_v4(a.out_dz + off_i, make_uint4(pack_h2(db[0], db[1]), pack_h2(db[2], db[3]), pack_h2(db[4], db[5]), pack_h2(db[6], db[7])));
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
    if (a.mode == LG_WIRE_DGRAD) {      // amax of the stored (scaled) values -> next step's scale; order-independent
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
      if (lane == 0 && amax > 0.f && isfinite(amax))
        atomicMax(reinterpret_cast<unsigned int*>(const_cast<float*>(a.scal)) + SC_LAYER_AMAX + a.dst_layer, __float_as_uint(amax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

cudaError_t launch_lgemm(const LGemmArgs& a, int n_sm, cudaStream_t stream) {
  const int items = a.n_tiles * a.n_nblocks;
  const int grid = items < n_sm ? items : n_sm;
  if (grid <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(lgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLgSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  lgemm_kernel<<<grid, kLgThreads, kLgSmem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace inr
