// Fused whole-MLP backward (dgrad chain) for width-256 real-valued chains.
//
// Prologue (every CTA, fixed reduction order => bit-reproducible): reduce the forward kernel's
// per-tile loss partials into the step scalars: masked row count m, HDR filter mean A, the loss
// value, the normalisation factors cA / cB and a power-of-two gradient scale S chosen from the
// amax of dL/dz_last so that every dZ fits fp16 (hard part 1).
//
// Per 128-row tile:  dZ_last = S (cA gA + cB gB)                       (CUDA cores, fp32)
//                    dZ_{L-1} = (dZ_last W_last) * act'(z_{L-1})        (CUDA cores, K = out_f <= 4)
//                    dZ_{l-1} = (dZ_l W_l) * act'(z_{l-1})              (tcgen05, fp32 accumulate in TMEM)
// Every dZ_l is written as an fp16 operand image for the split-K wgrad kernel; the image of the
// current layer also stays in shared memory (in place) as the A operand of the next dgrad GEMM.
// act'(z) images come from the forward kernel (cos(w0 z) for SIREN -- the w0 factor rides in the packed dgrad
// weights and in the staged W_last --, 1[z>0] for ReLU).
#include <cstdlib>
#include <cuda_runtime.h>
#include <cmath>
#include "inr_ptx.cuh"
#include "inr_kernels.cuh"
#include "inr_loss.cuh"

namespace inr {

constexpr int kBwdStages = 5;
constexpr int kBwdComputeThreads = 512;                   // warps 4..19
constexpr int kBwdThreads = 128 + kBwdComputeThreads;
constexpr int kBwdSmem = 2 * kActBytes + kBwdStages * kStageBytes + kMaxOut * kWidth * 4 + 1024;
static_assert(kBwdSmem <= 227 * 1024, "backward kernel shared memory budget");

__global__ void __launch_bounds__(kBwdThreads, 1) chain_bwd_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;                       // dZ_l image: A operand, rewritten in place per layer
  uint8_t* dimg = smem + kActBytes;          // act'(z) image of the layer being produced
  uint8_t* wring = smem + 2 * kActBytes;
  float* c_wlast = reinterpret_cast<float*>(smem + 2 * kActBytes + kBwdStages * kStageBytes);   // [kMaxOut][256] * zscale
  __shared__ uint64_t w_full[kBwdStages], w_empty[kBwdStages], d_full[4], d_empty, act_full[4], acc_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sc[kScalars];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ChainModel& M = a.m;
  const int n_tiles = a.w.n_tiles;
  const float zscale = (M.act == ACT_SIN) ? M.w0 : 1.f;   // act' images hold cos(w0 z); w0 rides in the weights

  griddep_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&d_full[i], 1); mbar_init(&act_full[i], kBwdComputeThreads); }
    mbar_init(&d_empty, kBwdComputeThreads);
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < M.out_f * kWidth; i += kBwdThreads) c_wlast[i] = zscale * a.params[M.w_off[M.n_gemm] + i];
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  griddep_wait();            // forward kernel complete: tile partials, loss pieces and act' images are final
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  reduce_step_scalars(reinterpret_cast<const float*>(a.ws + a.w.part_off), a.w.n_tiles, a.loss, a.m.out_f, a.bs_k, a.hyper, a.step,
                      blockIdx.x == 0 ? reinterpret_cast<float*>(a.ws + a.w.scal_off) : nullptr, sc);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: act' images + dgrad weight stages
    if (lane == 0) {
      uint32_t it = 0, dq = 0;
      const uint64_t pol_first = l2_policy_evict_first();
      auto load_dimg = [&](int l, int tile) {
        mbar_wait(&d_empty, (dq & 1) ^ 1);
        ++dq;
        const uint8_t* src = a.ws + a.w.d_off[l] + static_cast<size_t>(tile) * kActBytes;
        for (int c = 0; c < 4; ++c) {
          mbar_arrive_expect_tx(&d_full[c], kChunkBytes);
          if (a.l2_hints & 1) bulk_g2s_hint(dimg + c * kChunkBytes, src + c * kChunkBytes, kChunkBytes, &d_full[c], pol_first);
          else bulk_g2s(dimg + c * kChunkBytes, src + c * kChunkBytes, kChunkBytes, &d_full[c]);
        }
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        load_dimg(M.n_gemm - 1, tile);
        for (int l = M.n_gemm - 1; l >= 1; --l) {
          const uint8_t* src = a.wpack + M.wd_off[l];
          for (int s = 0; s < kWidth / kStageK; ++s, ++it) {
            const uint32_t slot = it % kBwdStages, ph = (it / kBwdStages) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&w_full[slot], kStageBytes);
            bulk_g2s(wring + slot * kStageBytes, src + static_cast<size_t>(s) * kStageBytes, kStageBytes, &w_full[slot]);
          }
          load_dimg(l - 1, tile);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: dH_{l-1} = dZ_l * W_l
    // convergent warp, one elected lane issues (warp-uniform descriptors; see chain_fwd.cu)
    {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, kWidth, false, false);
      uint32_t it = 0, act_use = 0;
      const uint32_t act_s = smem_u32(act), wring_s = smem_u32(wring);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = M.n_gemm - 1; l >= 1; --l, ++act_use) {
          const uint32_t acc = tmem + (l & 1) * kWidth;
          for (int c = 0; c < 4; ++c) {
            mbar_wait(&act_full[c], act_use & 1);
            tc_fence_after();
            const uint32_t a_base = act_s + c * kChunkBytes;
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2, ++it) {
              const uint32_t slot = it % kBwdStages;
              mbar_wait(&w_full[slot], (it / kBwdStages) & 1);
              tc_fence_after();
              __syncwarp();
              const uint32_t b_base = wring_s + slot * kStageBytes;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t da = umma_smem_desc(a_base + (s2 * 2 + kk) * 4096, 2048, 128);
                  const uint64_t db = umma_smem_desc(b_base + kk * 8192, 4096, 128);
                  umma_f16(acc, da, db, idesc, (c | s2 | kk) != 0);
                }
                umma_commit(&w_empty[slot]);
                if (c == 3 && s2 == 1) umma_commit(&acc_full[l & 1]);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps (16)
    const int q = warp & 3, sub = (warp - 4) >> 2, row = q * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    const float S = sc[SC_SCALE], cA = sc[SC_CA], cB = sc[SC_CB];
    uint32_t dq = 0, acc_ph[2] = {0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int grow = tile * kTileM + row;
      const bool valid = grow < a.bs;
      // ---- dZ_last (fp32) and its padded fp16 image
      float dz[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        if (a.dout) {
          for (int o = 0; o < M.out_f; ++o) dz[o] = S * a.dout[static_cast<size_t>(grow) * M.out_f + o];
        } else {
          const float4 g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.ws + a.w.g_off) +
                                                             (static_cast<size_t>(tile) * kTileM + row) * 4);
          dz[0] = S * (cA * g.x + cB * g.z);
          dz[1] = S * (cA * g.y + cB * g.w);
        }
      }
      if (sub == 0) {
        uint8_t* zl = a.ws + a.w.dzlast_off + static_cast<size_t>(tile) * kDzLastBytes;
        st_global_v4(zl + row * 16, make_uint4(pack_h2(dz[0], dz[1]), pack_h2(dz[2], dz[3]), 0u, 0u));
        st_global_v4(zl + 2048 + row * 16, make_uint4(0u, 0u, 0u, 0u));
      }
      // ---- layers, top down.  l == n_gemm-1: CUDA-core product with W_last; below: TMEM accumulators
      for (int l = M.n_gemm - 1; l >= 0; --l) {
        const bool from_last = (l == M.n_gemm - 1);
        uint8_t* dz_img = a.ws + a.w.dz_off[l] + static_cast<size_t>(tile) * kActBytes;
        if (!from_last) {
          mbar_wait(&acc_full[(l + 1) & 1], acc_ph[(l + 1) & 1]);
          acc_ph[(l + 1) & 1] ^= 1;
          tc_fence_after();
        }
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          const int col0 = g * kChunkCols + sub * 16;
          mbar_wait(&d_full[g], dq & 1);
          float v[16];
          if (from_last) {
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o)
                if (o < M.out_f) {
                  const float4 w = *reinterpret_cast<const float4*>(c_wlast + o * kWidth + col0 + 4 * i4);
                  acc.x = fmaf(dz[o], w.x, acc.x); acc.y = fmaf(dz[o], w.y, acc.y);
                  acc.z = fmaf(dz[o], w.z, acc.z); acc.w = fmaf(dz[o], w.w, acc.w);
                }
              v[4 * i4] = acc.x; v[4 * i4 + 1] = acc.y; v[4 * i4 + 2] = acc.z; v[4 * i4 + 3] = acc.w;
            }
          } else {
            tmem_ld16(tmem + t_lane + ((l + 1) & 1) * kWidth + col0, v);
            tmem_ld_wait();
          }
          const int kg0 = col0 >> 3;
          uint4 zv[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint4 dr = *reinterpret_cast<const uint4*>(dimg + (kg0 + j) * 2048 + row * 16);
            const __half2* dh = reinterpret_cast<const __half2*>(&dr);
            float z[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 d2 = __half22float2(dh[i]);
              z[2 * i] = v[8 * j + 2 * i] * d2.x;
              z[2 * i + 1] = v[8 * j + 2 * i + 1] * d2.y;
            }
            zv[j] = make_uint4(pack_h2(z[0], z[1]), pack_h2(z[2], z[3]), pack_h2(z[4], z[5]), pack_h2(z[6], z[7]));
          }
          if (l >= 1) {
            *reinterpret_cast<uint4*>(act + kg0 * 2048 + row * 16) = zv[0];
            *reinterpret_cast<uint4*>(act + (kg0 + 1) * 2048 + row * 16) = zv[1];
          }
          st_global_v4(dz_img + kg0 * 2048 + row * 16, zv[0]);
          st_global_v4(dz_img + (kg0 + 1) * 2048 + row * 16, zv[1]);
          if (l >= 1) {
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&act_full[g]);
          }
        }
        ++dq;
        mbar_arrive(&d_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

cudaError_t launch_chain_bwd(const BwdArgs& a, int n_sm, cudaStream_t stream) {
  const int grid = a.w.n_tiles < n_sm ? a.w.n_tiles : n_sm;
  if (grid <= 0) return cudaSuccess;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  BwdArgs b = a;
  // the act' images are dead once this kernel has read them: marking their lines evict_first leaves more of the dZ images
  // in L2 for wgrad (bs 100 000: wgrad 82.6 -> 80.3 us in-process, neutral at 300 000); INR_BWD_L2=0 switches it off per launch
  const char* env = std::getenv("INR_BWD_L2");
  b.l2_hints = env ? std::atoi(env) : 1;
  return launch_dependent(chain_bwd_kernel, dim3(grid), dim3(kBwdThreads), kBwdSmem, stream, b);
}

}  // namespace inr
