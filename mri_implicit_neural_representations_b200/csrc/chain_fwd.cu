// Fused whole-MLP forward for width-256 real-valued chains (SIREN, FFN):
//   positional encoding -> [tcgen05 GEMM -> bias + activation epilogue] x n_gemm -> final small linear
//   -> last activation -> loss pieces, one CTA per 128-coordinate tile, persistent over tiles.
//
// Warp roles (640 threads): warp 0 = weight-stage producer (1-D bulk TMA copies into a ring),
// warp 1 = MMA issuer (one thread, tcgen05.mma kind::f16, fp32 accumulators in TMEM, two
// accumulators ping-ponged by layer), warp 2 = TMEM allocator, warps 4..19 = compute (4 per TMEM lane
// quarter): they generate the encoding chunks that feed layer 0 and run every epilogue.  Activations never leave the SM
// between layers: the epilogue writes the fp16 A-operand image of the next layer in place, 64
// columns at a time, and the next layer's MMAs trail it chunk by chunk.
//
// Mirrors (results, not code): src/models/networks.py:23-35 (encoder), :74-124 (SIREN), :48-69 (FFN);
// loss pieces follow src/metrics/losses.py and src/train.py:178-182 (see loss_row below).
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "inr_kernels.cuh"
#include "inr_loss.cuh"

namespace inr {

// debug trace: slot i of the buffer <- %globaltimer (ns); only CTA 0, only when a buffer is supplied
#define INR_TRACE(args, slot) do { if ((args).trace && blockIdx.x == 0) (args).trace[(slot)] = global_ns(); } while (0)

constexpr int kFwdStages = 8;
constexpr int kFwdComputeThreads = 512;                   // warps 4..19
constexpr int kFwdThreads = 128 + kFwdComputeThreads;
constexpr int kFwdMaxEnc = 512;                           // encoder features staged in shared memory
constexpr int kFwdConstBytes = ((kMaxLayers - 1) * kWidth + kMaxOut * kWidth + kFwdMaxEnc * 3) * 4;
constexpr int kFwdSmem = kActBytes + kFwdStages * kStageBytes + kFwdConstBytes + 1024;
static_assert(kFwdSmem <= 227 * 1024, "forward kernel shared memory budget");

template <int ACT>
__global__ void __launch_bounds__(kFwdThreads, 1) chain_fwd_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* act = smem;
  uint8_t* wring = smem + kActBytes;
  float* c_bias = reinterpret_cast<float*>(smem + kActBytes + kFwdStages * kStageBytes);   // [n_gemm][256], pre-scaled
  float* c_wlast = c_bias + (kMaxLayers - 1) * kWidth;                                      // [kMaxOut][256]
  float* c_encB = c_wlast + kMaxOut * kWidth;                                               // [E][3]
  __shared__ uint64_t w_full[kFwdStages], w_empty[kFwdStages], in_full[4], in_empty[4], act_full[4], acc_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float out_part[3][kTileM][kMaxOut];
  __shared__ float red[4][8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ChainModel& M = a.m;
  const int n_tiles = a.w.n_tiles;
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const float zscale = (ACT == ACT_SIN) ? M.w0 : 1.f;     // packed weights carry the same factor (optim.cu)
  if (tid == 0) INR_TRACE(a, 0);
  griddep_launch_dependents();      // the backward kernel may set up (barriers, TMEM, constants) on SMs this grid leaves idle

  if (tid == 0) {
    for (int i = 0; i < kFwdStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&in_full[i], kFwdComputeThreads); mbar_init(&in_empty[i], 1); mbar_init(&act_full[i], kFwdComputeThreads);
    }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_fence_init();
    if (blockIdx.x == 0 && a.step_counter) *a.step_counter += 1;
  }
  // stage the small per-model constants once per CTA (broadcast LDS.128 in the epilogues instead of uniform LDGs)
  for (int i = tid; i < M.n_gemm * kWidth; i += kFwdThreads)
    c_bias[i] = zscale * a.params[M.b_off[i / kWidth] + (i % kWidth)];
  for (int i = tid; i < M.out_f * kWidth; i += kFwdThreads) c_wlast[i] = a.params[M.w_off[M.n_gemm] + i];
  if (M.input_kind == INPUT_GAUSS)
    for (int i = tid; i < M.enc_size * 3; i += kFwdThreads) c_encB[i] = 6.283185307179586f * a.encB[i];   // radians per unit coordinate
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) INR_TRACE(a, 1);

  if (warp == 0) {
    // ------------------------------------------------------------------ weight-stage producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < M.n_gemm; ++l) {
          const int nst = (l == 0 ? M.k0 : kWidth) / kStageK;
          const uint8_t* src = a.wpack + M.wf_off[l];
          for (int s = 0; s < nst; ++s, ++it) {
            const uint32_t slot = it % kFwdStages, ph = (it / kFwdStages) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&w_full[slot], kStageBytes);
            bulk_g2s(wring + slot * kStageBytes, src + static_cast<size_t>(s) * kStageBytes, kStageBytes, &w_full[slot]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the pipeline (all lanes poll the barriers) and one elected lane issues, so the descriptors stay
    // warp-uniform: a single divergent thread paid ~280 cycles per stage + ~50 per MMA (tools/umma_commit2.cu), more than
    // the 256 tensor cycles of a K = 32 weight stage.
    {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, kWidth, false, false);
      uint32_t it = 0, inq = 0, act_use = 0;
      const uint32_t act_s = smem_u32(act), wring_s = smem_u32(wring);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < M.n_gemm; ++l) {
          const uint32_t acc = tmem + (l & 1) * kWidth;
          const int nchunks = (l == 0 ? M.k0 : kWidth) / kChunkCols;
          for (int c = 0; c < nchunks; ++c) {
            uint32_t a_base, in_slot = 0;
            if (l == 0) {
              in_slot = inq & 3;
              mbar_wait(&in_full[in_slot], (inq >> 2) & 1);
              a_base = act_s + in_slot * kChunkBytes;
            } else {
              mbar_wait(&act_full[c], act_use & 1);
              a_base = act_s + c * kChunkBytes;
            }
            tc_fence_after();
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2, ++it) {
              const uint32_t slot = it % kFwdStages;
              mbar_wait(&w_full[slot], (it / kFwdStages) & 1);
              if (it < 8 && lane == 0) INR_TRACE(a, 48 + it);         // MMA warp: first eight weight stages landed
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
                if (l == 0 && s2 == 1) umma_commit(&in_empty[in_slot]);
                if (s2 == 1 && c == nchunks - 1) umma_commit(&acc_full[l & 1]);
              }
              __syncwarp();
            }
            if (l == 0) ++inq;
          }
          if (tile == 0 && lane == 0) INR_TRACE(a, 40 + l);          // MMA warp: layer l fully issued
          if (l > 0) ++act_use;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps (16): 4 per TMEM lane quarter
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int sub = (warp - 4) >> 2;        // which 16 columns of every 64-column chunk this warp owns
    const int row = q * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    uint32_t inq = 0, acc_ph[2] = {0, 0};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int grow = tile * kTileM + row;            // row inside the batch
      const bool valid = grow < a.bs;
      const size_t srow = static_cast<size_t>(row_base) + grow;   // row inside the resident arrays
      // ---------------- layer-0 input chunks (64 K-columns each) into the 4-slot ring
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (M.input_kind == INPUT_GAUSS && valid) {
        cx = a.coords[srow * 3 + 0]; cy = a.coords[srow * 3 + 1]; cz = a.coords[srow * 3 + 2];
      }
      float t_pref[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
      bool in_loss_pref = false;
      if (sub == 0 && a.train && valid && a.gt && a.loss.kind != LOSS_NONE) {   // loss inputs: fetched now, used at the tile's end
        in_loss_pref = a.mask ? (a.mask[srow] != 0) : true;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) if (o < M.out_f) t_pref[o] = a.gt[srow * M.out_f + o];
      }
      uint8_t* h0_img = a.ws + a.w.h_off[0] + static_cast<size_t>(tile) * (kTileM * M.k0 * 2);
      const int n_in_chunks = M.k0 / kChunkCols;
      for (int c = 0; c < n_in_chunks; ++c, ++inq) {
        const uint32_t slot = inq & 3;
        mbar_wait(&in_empty[slot], ((inq >> 2) & 1) ^ 1);
        uint8_t* dst = act + slot * kChunkBytes;
        uint4 v0, v1;
        int g0, g1;   // the two 16-byte k-groups (within the chunk) this thread fills
        if (M.input_kind == INPUT_GAUSS) {
          // chunk c holds [sin f | cos f] for the 32 encoder features f = 32c .. 32c+31; this thread does 8 of them
          const float* b = c_encB + (c * 32 + sub * 8) * 3;
          float s[8], co[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // 2*pi*(x . B_f); sin.approx's own range reduction (x * 1/2pi -> MUFU) is accurate to ~1e-5 rad
            // at |arg| ~ 100, far below the fp16 rounding of the stored feature
            const float ang = fmaf(cx, b[3 * i], fmaf(cy, b[3 * i + 1], cz * b[3 * i + 2]));
            s[i] = fast_sin(ang); co[i] = fast_cos(ang);
          }
          v0 = make_uint4(pack_h2(s[0], s[1]), pack_h2(s[2], s[3]), pack_h2(s[4], s[5]), pack_h2(s[6], s[7]));
          v1 = make_uint4(pack_h2(co[0], co[1]), pack_h2(co[2], co[3]), pack_h2(co[4], co[5]), pack_h2(co[6], co[7]));
          g0 = sub; g1 = 4 + sub;
        } else {
          // dense fp32 input: natural column order, this thread converts 16 of the chunk's 64 columns
          const float* xr = a.x + srow * M.k0 + c * kChunkCols + sub * 16;
          float4 p[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) p[j] = valid ? *reinterpret_cast<const float4*>(xr + 4 * j) : make_float4(0, 0, 0, 0);
          v0 = make_uint4(pack_h2(p[0].x, p[0].y), pack_h2(p[0].z, p[0].w), pack_h2(p[1].x, p[1].y), pack_h2(p[1].z, p[1].w));
          v1 = make_uint4(pack_h2(p[2].x, p[2].y), pack_h2(p[2].z, p[2].w), pack_h2(p[3].x, p[3].y), pack_h2(p[3].z, p[3].w));
          g0 = sub * 2; g1 = sub * 2 + 1;
        }
        *reinterpret_cast<uint4*>(dst + g0 * 2048 + row * 16) = v0;
        *reinterpret_cast<uint4*>(dst + g1 * 2048 + row * 16) = v1;
        if (a.train) {
          uint8_t* gdst = h0_img + static_cast<size_t>(c) * kChunkBytes;
          st_global_v4(gdst + g0 * 2048 + row * 16, v0);
          st_global_v4(gdst + g1 * 2048 + row * 16, v1);
        }
        fence_proxy_async_smem();
        mbar_arrive(&in_full[slot]);
        if (tid == 128 && tile == 0) INR_TRACE(a, 2 + c);      // compute: encoding chunk c written
      }
      // ---------------- epilogues
      float po[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
      for (int l = 0; l < M.n_gemm; ++l) {
        const bool last_gemm = (l == M.n_gemm - 1);
        const float* bias = c_bias + l * kWidth;
        uint8_t* h_img = a.ws + a.w.h_off[l + 1] + static_cast<size_t>(tile) * kActBytes;
        uint8_t* d_img = a.ws + a.w.d_off[l] + static_cast<size_t>(tile) * kActBytes;
        mbar_wait(&acc_full[l & 1], acc_ph[l & 1]);
        acc_ph[l & 1] ^= 1;
        tc_fence_after();
        if (tid == 128 && tile == 0) INR_TRACE(a, 12 + 5 * l);  // compute: accumulator of layer l ready
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          const int col0 = g * kChunkCols + sub * 16;   // chunk g completes after step g: the next layer's MMAs trail in order
          float v[16];
          tmem_ld16(tmem + t_lane + (l & 1) * kWidth + col0, v);
          tmem_ld_wait();
          uint4 hv[2], dv[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float h[8], d[8];
            const float4 b0 = *reinterpret_cast<const float4*>(bias + col0 + 8 * j);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + col0 + 8 * j + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float z = v[8 * j + i] + bb[i];          // SIREN: already w0 * (x W^T + b)
              if (ACT == ACT_SIN) { h[i] = fast_sin(z); d[i] = fast_cos(z); }
              else { h[i] = fmaxf(z, 0.f); d[i] = z > 0.f ? 1.f : 0.f; }
            }
            hv[j] = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            dv[j] = make_uint4(pack_h2(d[0], d[1]), pack_h2(d[2], d[3]), pack_h2(d[4], d[5]), pack_h2(d[6], d[7]));
            if (last_gemm) {
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o)
                if (o < M.out_f) {
                  const float4 w0v = *reinterpret_cast<const float4*>(c_wlast + o * kWidth + col0 + 8 * j);
                  const float4 w1v = *reinterpret_cast<const float4*>(c_wlast + o * kWidth + col0 + 8 * j + 4);
                  po[o] = fmaf(h[0], w0v.x, fmaf(h[1], w0v.y, fmaf(h[2], w0v.z, fmaf(h[3], w0v.w, po[o]))));
                  po[o] = fmaf(h[4], w1v.x, fmaf(h[5], w1v.y, fmaf(h[6], w1v.z, fmaf(h[7], w1v.w, po[o]))));
                }
            }
          }
          const int kg0 = col0 >> 3;
          if (!last_gemm) {
            *reinterpret_cast<uint4*>(act + kg0 * 2048 + row * 16) = hv[0];
            *reinterpret_cast<uint4*>(act + (kg0 + 1) * 2048 + row * 16) = hv[1];
          }
          if (a.train) {
            st_global_v4(h_img + kg0 * 2048 + row * 16, hv[0]);
            st_global_v4(h_img + (kg0 + 1) * 2048 + row * 16, hv[1]);
            st_global_v4(d_img + kg0 * 2048 + row * 16, dv[0]);
            st_global_v4(d_img + (kg0 + 1) * 2048 + row * 16, dv[1]);
          }
          if (!last_gemm) {
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&act_full[g]);
          }
          if (tid == 128 && tile == 0) INR_TRACE(a, 13 + 5 * l + g);   // compute: epilogue step g of layer l done
        }
      }
      // ---------------- final linear (CUDA cores) + last activation + loss pieces
      if (sub > 0) {
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) out_part[sub - 1][row][o] = po[o];
      }
      named_bar_sync(1, kFwdComputeThreads);
      if (sub == 0) {
        float y[kMaxOut], t[kMaxOut], dact[kMaxOut];
        const float* bl = a.params + M.b_off[M.n_gemm];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          y[o] = 0.f; t[o] = 0.f; dact[o] = 1.f;
          if (o < M.out_f) {
            const float z = ((po[o] + out_part[0][row][o]) + (out_part[1][row][o] + out_part[2][row][o])) + __ldg(bl + o);
            if (M.last_act == LAST_TANH) { y[o] = tanh_acc(z); dact[o] = 1.f - y[o] * y[o]; }
            else if (M.last_act == LAST_SIGMOID) { y[o] = 1.f / (1.f + expf(-z)); dact[o] = y[o] * (1.f - y[o]); }
            else y[o] = z;
          }
        }
        if (valid && a.out) {
          for (int o = 0; o < M.out_f; ++o) a.out[(static_cast<size_t>(grow)) * M.out_f + o] = y[o];
        }
        if (a.train) {
          float lA = 0.f, lB = 0.f, fs = 0.f, cnt = 0.f, amA = 0.f, amB = 0.f;
          float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && a.gt && a.loss.kind != LOSS_NONE) {
            const bool in_loss = in_loss_pref;
            if (a.loss.kind == LOSS_HDR) {   // filter term runs over ALL batch rows (unmasked kcoords)
              const float kx = a.coords[srow * 3 + 1], ky = a.coords[srow * 3 + 2];
              const float f = expf(-(kx * kx + ky * ky) / (2.f * a.loss.sigma * a.loss.sigma));
              fs = (1.f - f) * (1.f - f);
            }
            if (in_loss) {
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o) t[o] = t_pref[o];
              RowLoss r = loss_row(a.loss, M.out_f, y, t);
              lA = r.lossA; lB = r.lossB; cnt = 1.f;
              float ga[kMaxOut], gb[kMaxOut];
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o) {
                ga[o] = r.gA[o] * dact[o]; gb[o] = r.gB[o] * dact[o];
                amA = fmaxf(amA, fabsf(ga[o])); amB = fmaxf(amB, fabsf(gb[o]));
              }
              gq = make_float4(ga[0], ga[1], gb[0], gb[1]);
            }
          }
          float* gdst = reinterpret_cast<float*>(a.ws + a.w.g_off) + (static_cast<size_t>(tile) * kTileM + row) * 4;
          *reinterpret_cast<float4*>(gdst) = gq;
          // deterministic tile partials: warp shuffle tree, then 4 warps in fixed order
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            lA += __shfl_xor_sync(0xffffffffu, lA, off); lB += __shfl_xor_sync(0xffffffffu, lB, off);
            fs += __shfl_xor_sync(0xffffffffu, fs, off); cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
            amA = fmaxf(amA, __shfl_xor_sync(0xffffffffu, amA, off));
            amB = fmaxf(amB, __shfl_xor_sync(0xffffffffu, amB, off));
          }
          if (lane == 0) { red[q][0] = lA; red[q][1] = lB; red[q][2] = fs; red[q][3] = cnt; red[q][4] = amA; red[q][5] = amB; }
          named_bar_sync(2, 128);
          if (q == 0 && lane == 0) {
            float* pdst = reinterpret_cast<float*>(a.ws + a.w.part_off) + static_cast<size_t>(tile) * kPartialsPerTile;
            pdst[0] = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
            pdst[1] = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
            pdst[2] = (red[0][2] + red[1][2]) + (red[2][2] + red[3][2]);
            pdst[3] = (red[0][3] + red[1][3]) + (red[2][3] + red[3][3]);
            pdst[4] = fmaxf(fmaxf(red[0][4], red[1][4]), fmaxf(red[2][4], red[3][4]));
            pdst[5] = fmaxf(fmaxf(red[0][5], red[1][5]), fmaxf(red[2][5], red[3][5]));
            pdst[6] = 0.f; pdst[7] = 0.f;
          }
          named_bar_sync(2, 128);
        }
      }
    }
  }
  if (tid == 128) INR_TRACE(a, 38);                      // compute thread: all tiles done
  tc_fence_before();
  __syncthreads();
  if (tid == 0) INR_TRACE(a, 39);
  if (warp == 2) tmem_dealloc<512>(tmem);
}

cudaError_t launch_chain_fwd(const FwdArgs& a, int n_sm, cudaStream_t stream) {
  const int grid = a.w.n_tiles < n_sm ? a.w.n_tiles : n_sm;
  if (grid <= 0) return cudaSuccess;
  static bool attr_done = false;     // once per process: keeps attribute calls out of graph capture
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(chain_fwd_kernel<ACT_SIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chain_fwd_kernel<ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (a.m.act == ACT_SIN) chain_fwd_kernel<ACT_SIN><<<grid, kFwdThreads, kFwdSmem, stream>>>(a);
  else chain_fwd_kernel<ACT_RELU><<<grid, kFwdThreads, kFwdSmem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace inr
