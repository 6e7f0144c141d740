// Transposed ("swap-AB") whole-MLP forward and backward for width-256 real-valued chains (SIREN, FFN) at SMALL batches.
//
// The row-tile kernels (chain_fwd.cu / chain_bwd.cu) put 128 batch rows on the TMEM lanes; a batch of 10 000 rows (BASELINE
// configs[0]) is then 79 tiles on 148 SMs, and a tile's epilogue is bound per SM sub-partition (the warps of TMEM lane
// quarter q live on sub-partition q: 32 rows x 256 columns x (sin, cos) whatever the tile height).  Here the product is
// transposed: D^T[256 features x NR rows] = W[256 x K] * H^T, i.e. the packed weights are the A operand (two M = 128 MMAs
// per K step, the same packed stages the row-tile kernels read as B), the activation image [NR rows x K] is the B operand,
// the output FEATURES sit on the TMEM lanes and the batch rows on the columns.  NR is any multiple of 16 up to 80, chosen
// so that ceil(bs / NR) <= #SMs: 10 000 rows = 125 tiles of 80 rows, every tile does 5/8 of the transcendental work per
// sub-partition and all of them run in one wave.
//
// Images (shared memory and HBM alike) are NR-row tiles:  elem(row r, feature f) at (f/8) * LB + r*16 + (f%8)*2 with the
// k-group stride LB = NR*16 + 16 (Workspace::lb): a K-major UMMA B operand with LBO = LB, SBO = 128 -- and the MN-major operand
// of the split-K wgrad kernel with SBO = LB (WgradArgs::lb).  The 16 padding bytes per k-group keep the epilogue's 2-byte
// stores (a warp = 32 consecutive features of one row = 4 k-groups) free of bank conflicts: NR*16 alone is a multiple of 256.
// A whole image of a tile is one contiguous run, so it leaves and enters by ONE bulk copy.
//
// Weight stages are shared inside a thread-block cluster: every tile needs all of the layer's weights, and 125 CTAs pulling
// 640 KB each out of L2 is 80 MB per forward pass -- more than 10 us of L2 -> SM bandwidth on its own.  The CTAs of a cluster
// (4, or 2 where the device cannot co-schedule 4) walk the stages in lockstep; CTA c fetches 1/C of every stage and multicasts
// it into the same ring slot of all C CTAs, and a slot is refilled once the MMAs of all C CTAs have released it (commit
// multicast to every CTA's `empty` barrier).  A grid rounded up to whole clusters recomputes the last tile in its spare CTAs
// (identical bytes to identical addresses) instead of special-casing the ring protocol.
//
// Status: opt-in (INR_CHAIN_T=1), parity-tested on both tilings (tests/test_gpu_chain.py, tests/test_gpu_chain_t.py).  Measured on
// B200 at 10 000 rows the epilogue of a layer drops from 3.7 to 1.8 us as intended, but K runs over FEATURES here, so layer l + 1
// cannot start before the whole image of layer l exists: the MMA phases (paced by the weight ring, 3.3 us per hidden layer)
// are exposed instead of trailing the epilogue chunk by chunk, and the step takes 71.9 us against 59.7 us.
//
// Mirrors (results, not code): src/models/networks.py:23-35 (encoder), :74-124 (SIREN), :48-69 (FFN).
#include <cuda_runtime.h>
#include <cmath>
#include <cstdlib>
#include "inr_ptx.cuh"
#include "inr_kernels.cuh"
#include "inr_loss.cuh"

namespace inr {

constexpr int kTMaxRows = 80;                               // NR <= 80: 512 features x 80 rows x 2 B = 80 KB input image
constexpr int kTAccStride = 96;                             // TMEM columns between accumulators (>= NR): 4 x 96 <= 512
constexpr int kTThreads = 640, kTCompute = 512;
constexpr int kTFwdStages = 4;
constexpr int kTMaxLB = kTMaxRows * 16 + 16;                // k-group stride at NR = 80
constexpr int kTXBytes = 64 * kTMaxLB;                      // 512-feature image: 82944
constexpr int kTYBytes = 32 * kTMaxLB;                      // 256-feature image: 41472
constexpr int kTFwdConstBytes = ((kMaxLayers - 1) * kWidth + kMaxOut * kWidth + 512 * 3) * 4;
constexpr int kTFwdSmem = kTXBytes + kTYBytes + kTFwdStages * kStageBytes + kTFwdConstBytes + 1024;
static_assert(kTFwdSmem <= 227 * 1024, "transposed forward kernel shared memory budget");

#define INR_TRACE(args, slot) do { if ((args).trace && blockIdx.x == 0) (args).trace[(slot)] = global_ns(); } while (0)

__device__ __forceinline__ void st_shared_h(uint8_t* p, float v) { *reinterpret_cast<__half*>(p) = __float2half_rn(v); }

template <int ACT>
__global__ void __launch_bounds__(kTThreads, 1) chain_fwd_t_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* X = smem;                                        // layer-0 input image (K0 features); later H images / act' staging
  uint8_t* Y = smem + kTXBytes;
  uint8_t* wring = Y + kTYBytes;
  float* c_bias = reinterpret_cast<float*>(wring + kTFwdStages * kStageBytes);      // [n_gemm][256], pre-scaled by w0
  float* c_wlast = c_bias + (kMaxLayers - 1) * kWidth;                              // [kMaxOut][256]
  float* c_encB = c_wlast + kMaxOut * kWidth;                                       // [E][3], radians per unit coordinate
  __shared__ uint64_t w_full[kTFwdStages], w_empty[kTFwdStages], img_full, acc_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[16][8];
  __shared__ float s_xyz[kTMaxRows * 3];                    // coordinates of the tile's rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ChainModel& M = a.m;
  const int NR = a.w.tile_rows, LB = a.w.lb;
  const int n_tiles = a.w.n_tiles;
  const int n_rounds = (n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctas();
  const uint16_t cmask = static_cast<uint16_t>((1u << csize) - 1u);
  const uint32_t slice = kStageBytes / csize;               // this CTA's share of every weight stage
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const float zscale = (ACT == ACT_SIN) ? M.w0 : 1.f;
  if (tid == 0) INR_TRACE(a, 0);
  griddep_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < kTFwdStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], csize); }
    mbar_init(&img_full, kTCompute);
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_fence_init();
    if (blockIdx.x == 0 && a.step_counter) *a.step_counter += 1;
  }
  for (int i = tid; i < M.n_gemm * kWidth; i += kTThreads) c_bias[i] = zscale * a.params[M.b_off[i / kWidth] + (i % kWidth)];
  for (int i = tid; i < M.out_f * kWidth; i += kTThreads) c_wlast[i] = a.params[M.w_off[M.n_gemm] + i];
  if (M.input_kind == INPUT_GAUSS)
    for (int i = tid; i < M.enc_size * 3; i += kTThreads) c_encB[i] = 6.283185307179586f * a.encB[i];
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  cluster_sync_all();        // (also the CTA barrier) nobody multicasts into a peer whose barriers are not initialised yet
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) INR_TRACE(a, 1);

  if (warp == 0) {
    // ------------------------------------------------------------------ weight-stage producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int rd = 0; rd < n_rounds; ++rd)
        for (int l = 0; l < M.n_gemm; ++l) {
          const int nst = (l == 0 ? M.k0 : kWidth) / kStageK;
          const uint8_t* src = a.wpack + M.wf_off[l] + crank * slice;
          for (int s = 0; s < nst; ++s, ++it) {
            const uint32_t slot = it % kTFwdStages, ph = (it / kTFwdStages) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1);              // released by the MMAs of every CTA of the cluster
            mbar_arrive_expect_tx(&w_full[slot], kStageBytes);
            bulk_g2s_mc(wring + slot * kStageBytes + crank * slice, src + static_cast<size_t>(s) * kStageBytes, slice, &w_full[slot], cmask);
          }
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane)
    const uint32_t idesc = umma_idesc_f16(kTileM, NR, false, false);
    const uint32_t x_s = smem_u32(X), y_s = smem_u32(Y), wring_s = smem_u32(wring);
    uint32_t it = 0, iq = 0;
    for (int rd = 0; rd < n_rounds; ++rd)
      for (int l = 0; l < M.n_gemm; ++l) {
        mbar_wait(&img_full, iq & 1); ++iq;                 // input image of layer l complete
        tc_fence_after();
        const uint32_t img = (l & 1) ? y_s : x_s;           // layer 0 reads X, writes Y; layer 1 reads Y, writes X; ...
        const int nst = (l == 0 ? M.k0 : kWidth) / kStageK;
        for (int s = 0; s < nst; ++s, ++it) {
          const uint32_t slot = it % kTFwdStages;
          mbar_wait(&w_full[slot], (it / kTFwdStages) & 1);
          if (it < 8 && lane == 0) INR_TRACE(a, 48 + it);
          tc_fence_after();
          __syncwarp();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                // A: output features 128h .. 128h+127 of the packed stage (256 rows x 32 K: k-group stride 4096 B)
                const uint64_t da = umma_smem_desc(wring_s + slot * kStageBytes + h * 2048 + kk * 8192, 4096, 128);
                const uint64_t db = umma_smem_desc(img + (s * 4 + kk * 2) * LB, LB, 128);
                umma_f16(tmem + ((l & 1) * 2 + h) * kTAccStride, da, db, idesc, (s | kk) != 0);
              }
            umma_commit_mc(&w_empty[slot], cmask);
            if (s == nst - 1) umma_commit(&acc_full[l & 1]);
          }
          __syncwarp();
        }
        if (rd == 0 && lane == 0) INR_TRACE(a, 40 + l);
      }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps (16)
    const int ct = tid - 128;                               // 0 .. 511
    const int q = warp & 3, j = (warp - 4) >> 2, h = j & 1, ch = j >> 1;
    const int f = 128 * h + 32 * q + lane;                  // output feature of this thread (a TMEM lane of accumulator h)
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    const int c_lo = ch * (NR >> 1), c_hi = c_lo + (NR >> 1);
    uint8_t* DST = X + 32 * LB;                     // act' staging: upper half of X, free once layer 0's MMAs are done
    uint32_t acc_ph[2] = {0, 0};
    for (int rd = 0; rd < n_rounds; ++rd) {
      const int tile = min(static_cast<int>(blockIdx.x) + rd * static_cast<int>(gridDim.x), n_tiles - 1);
      const int r0g = tile * NR;                            // first batch row of the tile
      if (rd > 0) named_bar_sync(1, kTCompute);             // the previous tile's last readers of s_xyz are done
      // everything the tile reads from global memory row by row is fetched now, in one round trip: coordinates into shared
      // memory (the encoder reads a different row per item), targets and mask into the registers of the row's loss thread
      if (a.coords && ct < 3 * NR) {
        const int idx = r0g * 3 + ct;
        s_xyz[ct] = idx < a.bs * 3 ? a.coords[static_cast<size_t>(row_base) * 3 + idx] : 0.f;
      }
      float t_pref[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
      bool in_loss_pref = false;
      if (a.train && (ct & 3) == 0 && ct < 4 * NR && r0g + (ct >> 2) < a.bs && a.gt && a.loss.kind != LOSS_NONE) {
        const size_t srow = static_cast<size_t>(row_base) + r0g + (ct >> 2);
        in_loss_pref = a.mask ? (a.mask[srow] != 0) : true;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) if (o < M.out_f) t_pref[o] = a.gt[srow * M.out_f + o];
      }
      // every bulk store issued so far has finished READING shared memory before anybody overwrites an image
      if (ct == 0) bulk_wait_read0();
      named_bar_sync(1, kTCompute);
      // ---------------- layer-0 input image
      if (M.input_kind == INPUT_GAUSS) {
        // [sin 32 | cos 32] per 64 K columns, as the packed first-layer weights expect (perm_e)
        const int n_fg = M.enc_size >> 3;
        for (int item = ct; item < NR * n_fg; item += kTCompute) {
          const int r = item % NR, fg = item / NR;
          const int grow = r0g + r;
          float s[8], co[8];
          if (grow < a.bs) {
            const float cx = s_xyz[3 * r], cy = s_xyz[3 * r + 1], cz = s_xyz[3 * r + 2];
            const float* b = c_encB + fg * 24;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float ang = fmaf(cx, b[3 * i], fmaf(cy, b[3 * i + 1], cz * b[3 * i + 2]));
              s[i] = fast_sin(ang); co[i] = fast_cos(ang);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] = 0.f; co[i] = 0.f; }
          }
          const int kg_s = 8 * (fg >> 2) + (fg & 3);        // k-group of the sin part inside its 64-column chunk; cos: + 4
          *reinterpret_cast<uint4*>(X + kg_s * LB + r * 16) =
              make_uint4(pack_h2(s[0], s[1]), pack_h2(s[2], s[3]), pack_h2(s[4], s[5]), pack_h2(s[6], s[7]));
          *reinterpret_cast<uint4*>(X + (kg_s + 4) * LB + r * 16) =
              make_uint4(pack_h2(co[0], co[1]), pack_h2(co[2], co[3]), pack_h2(co[4], co[5]), pack_h2(co[6], co[7]));
        }
      } else {
        const int n_kg = M.k0 >> 3;
        for (int item = ct; item < NR * n_kg; item += kTCompute) {
          const int r = item % NR, kg = item / NR;
          const int grow = r0g + r;
          float v[8];
          const float* xr = a.x + (static_cast<size_t>(row_base) + grow) * M.k0 + kg * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = grow < a.bs ? xr[i] : 0.f;
          *reinterpret_cast<uint4*>(X + kg * LB + r * 16) =
              make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&img_full);
      if (ct == 0 && rd == 0) INR_TRACE(a, 2);
      named_bar_sync(1, kTCompute);
      if (ct == 0 && a.train) {
        bulk_s2g(a.ws + a.w.h_off[0] + static_cast<size_t>(tile) * ((M.k0 >> 3) * LB), X, (M.k0 >> 3) * LB);
        bulk_commit();
      }
      // ---------------- epilogues: bias + activation, next input image + act' image
      for (int l = 0; l < M.n_gemm; ++l) {
        mbar_wait(&acc_full[l & 1], acc_ph[l & 1]);
        acc_ph[l & 1] ^= 1;
        tc_fence_after();
        if (ct == 0 && rd == 0) INR_TRACE(a, 12 + 5 * l);
        if (ct == 0) bulk_wait_read0();
        named_bar_sync(1, kTCompute);
        if (ct == 0 && rd == 0) INR_TRACE(a, 13 + 5 * l);
        uint8_t* OUT = (l & 1) ? X : Y;
        const float bias = c_bias[l * kWidth + f];
        const uint32_t acc = tmem + t_lane + ((l & 1) * 2 + h) * kTAccStride;
        uint8_t* o_f = OUT + (f >> 3) * LB + (f & 7) * 2;
        uint8_t* d_f = DST + (f >> 3) * LB + (f & 7) * 2;
        for (int c8 = c_lo; c8 < c_hi; c8 += 8) {
          float v[8];
          tmem_ld8(acc + c8, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float z = v[i] + bias;                    // SIREN: already w0 * (W x + b)
            float hv, dv;
            if (ACT == ACT_SIN) { hv = fast_sin(z); dv = fast_cos(z); }
            else { hv = fmaxf(z, 0.f); dv = z > 0.f ? 1.f : 0.f; }
            st_shared_h(o_f + (c8 + i) * 16, hv);
            if (a.train) st_shared_h(d_f + (c8 + i) * 16, dv);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        if (l < M.n_gemm - 1) mbar_arrive(&img_full);
        if (ct == 0 && rd == 0) INR_TRACE(a, 14 + 5 * l);
        named_bar_sync(1, kTCompute);
        if (ct == 0 && rd == 0) INR_TRACE(a, 15 + 5 * l);
        if (ct == 0 && a.train) {
          bulk_s2g(a.ws + a.w.h_off[l + 1] + static_cast<size_t>(tile) * (32 * LB), OUT, 32 * LB);
          bulk_s2g(a.ws + a.w.d_off[l] + static_cast<size_t>(tile) * (32 * LB), DST, 32 * LB);
          bulk_commit();
        }
        if (ct == 0 && rd == 0 && l == M.n_gemm - 1) INR_TRACE(a, 30);
      }
      // ---------------- final linear (CUDA cores, four threads per row) + last activation + loss pieces
      const uint8_t* HL = ((M.n_gemm - 1) & 1) ? X : Y;
      const int nw = NR >> 3;                               // warps that hold rows: 4 * NR threads
      if (ct < 4 * NR) {
        const int r = ct >> 2, part = ct & 3, grow = r0g + r;
        const bool valid = part == 0 && grow < a.bs;        // lane `part == 0` of a row's quad owns the row's output and loss
        float po[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
        // 64 of the 256 hidden features: k-groups part, part + 4, ... (neighbouring lanes 1296 B apart: no bank conflict).
        // Not unrolled on purpose: code that runs once per CTA is paced by cold instruction fetches, not by its arithmetic
        // (the unrolled loop took 3.7 us here).
#pragma unroll 1
        for (int kg = part; kg < kWidth / 8; kg += 4) {
          const uint4 hv = *reinterpret_cast<const uint4*>(HL + kg * LB + r * 16);
          const __half2* hh = reinterpret_cast<const __half2*>(&hv);
          float hf[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 t2 = __half22float2(hh[i]); hf[2 * i] = t2.x; hf[2 * i + 1] = t2.y; }
#pragma unroll
          for (int o = 0; o < kMaxOut; ++o)
            if (o < M.out_f) {
              const float4 w0v = *reinterpret_cast<const float4*>(c_wlast + o * kWidth + kg * 8);
              const float4 w1v = *reinterpret_cast<const float4*>(c_wlast + o * kWidth + kg * 8 + 4);
              const float s0 = fmaf(hf[0], w0v.x, fmaf(hf[1], w0v.y, fmaf(hf[2], w0v.z, hf[3] * w0v.w)));
              const float s1 = fmaf(hf[4], w1v.x, fmaf(hf[5], w1v.y, fmaf(hf[6], w1v.z, hf[7] * w1v.w)));
              po[o] += s0 + s1;                             // short chains, then a tree over the quad: fp32 error stays ~1e-7
            }
        }
        if (ct == 0 && rd == 0) INR_TRACE(a, 31);
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          po[o] += __shfl_xor_sync(0xffffffffu, po[o], 1);
          po[o] += __shfl_xor_sync(0xffffffffu, po[o], 2);
        }
        float y[kMaxOut], t[kMaxOut], dact[kMaxOut];
        const float* bl = a.params + M.b_off[M.n_gemm];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          y[o] = 0.f; t[o] = 0.f; dact[o] = 1.f;
          if (o < M.out_f) {
            const float z = po[o] + __ldg(bl + o);
            if (M.last_act == LAST_TANH) { y[o] = tanh_acc(z); dact[o] = 1.f - y[o] * y[o]; }
            else if (M.last_act == LAST_SIGMOID) { y[o] = 1.f / (1.f + expf(-z)); dact[o] = y[o] * (1.f - y[o]); }
            else y[o] = z;
          }
        }
        if (valid && a.out)
          for (int o = 0; o < M.out_f; ++o) a.out[static_cast<size_t>(grow) * M.out_f + o] = y[o];
        if (a.train) {
          float lA = 0.f, lB = 0.f, fs = 0.f, cnt = 0.f, amA = 0.f, amB = 0.f;
          float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && a.gt && a.loss.kind != LOSS_NONE) {
            const bool in_loss = in_loss_pref;
            if (a.loss.kind == LOSS_HDR) {   // filter term runs over ALL batch rows (unmasked kcoords)
              const float kx = s_xyz[3 * r + 1], ky = s_xyz[3 * r + 2];
              const float ff = expf(-(kx * kx + ky * ky) / (2.f * a.loss.sigma * a.loss.sigma));
              fs = (1.f - ff) * (1.f - ff);
            }
            if (in_loss) {
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o) t[o] = t_pref[o];
              RowLoss rl = loss_row(a.loss, M.out_f, y, t);
              lA = rl.lossA; lB = rl.lossB; cnt = 1.f;
              float ga[kMaxOut], gb[kMaxOut];
#pragma unroll
              for (int o = 0; o < kMaxOut; ++o) {
                ga[o] = rl.gA[o] * dact[o]; gb[o] = rl.gB[o] * dact[o];
                amA = fmaxf(amA, fabsf(ga[o])); amB = fmaxf(amB, fabsf(gb[o]));
              }
              gq = make_float4(ga[0], ga[1], gb[0], gb[1]);
            }
          }
          if (part == 0)
            reinterpret_cast<float4*>(a.ws + a.w.g_off)[static_cast<size_t>(tile) * NR + r] = gq;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            lA += __shfl_xor_sync(0xffffffffu, lA, off); lB += __shfl_xor_sync(0xffffffffu, lB, off);
            fs += __shfl_xor_sync(0xffffffffu, fs, off); cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
            amA = fmaxf(amA, __shfl_xor_sync(0xffffffffu, amA, off));
            amB = fmaxf(amB, __shfl_xor_sync(0xffffffffu, amB, off));
          }
          if (ct == 0 && rd == 0) INR_TRACE(a, 32);
          const int wq = ct >> 5;
          if (lane == 0) { red[wq][0] = lA; red[wq][1] = lB; red[wq][2] = fs; red[wq][3] = cnt; red[wq][4] = amA; red[wq][5] = amB; }
          named_bar_sync(2, 32 * nw);
          if (ct == 0 && rd == 0) INR_TRACE(a, 33);
          if (ct == 0) {
            float* pdst = reinterpret_cast<float*>(a.ws + a.w.part_off) + static_cast<size_t>(tile) * kPartialsPerTile;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, m4 = 0.f, m5 = 0.f;
            for (int w = 0; w < nw; ++w) {                  // fixed order
              s0 += red[w][0]; s1 += red[w][1]; s2 += red[w][2]; s3 += red[w][3];
              m4 = fmaxf(m4, red[w][4]); m5 = fmaxf(m5, red[w][5]);
            }
            pdst[0] = s0; pdst[1] = s1; pdst[2] = s2; pdst[3] = s3; pdst[4] = m4; pdst[5] = m5; pdst[6] = 0.f; pdst[7] = 0.f;
          }
          named_bar_sync(2, 32 * nw);
        }
      }
    }
    if (ct == 0) INR_TRACE(a, 37);
    if (ct == 0) bulk_wait0();                              // the last images have left shared memory before the CTA exits
    if (ct == 0) INR_TRACE(a, 38);
  }
  tc_fence_before();
  cluster_sync_all();        // no CTA leaves while a peer may still multicast into its ring or arrive on its barriers
  if (tid == 0) INR_TRACE(a, 39);
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// =====================================================================================================================
// backward
// =====================================================================================================================
constexpr int kTBwdStages = 3;
constexpr int kTBwdSmem = 4 * kTYBytes + kTBwdStages * kStageBytes + kMaxOut * kWidth * 4 + kTMaxRows * 16 + 1024;
static_assert(kTBwdSmem <= 227 * 1024, "transposed backward kernel shared memory budget");

__global__ void __launch_bounds__(kTThreads, 1) chain_bwd_t_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Zb = smem;                                       // dZ images, ping-pong (x kTYBytes): B operand of the dgrad GEMMs
  uint8_t* Db = smem + 2 * kTYBytes;                        // act' images, double-buffered (x kTYBytes)
  uint8_t* wring = smem + 4 * kTYBytes;
  float* c_wlast = reinterpret_cast<float*>(wring + kTBwdStages * kStageBytes);     // [kMaxOut][256] * zscale
  float* dzl = c_wlast + kMaxOut * kWidth;                                          // [NR][4] dZ_last of the tile, fp32
  __shared__ uint64_t w_full[kTBwdStages], w_empty[kTBwdStages], d_full[2], d_empty[2], img_full, acc_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sc[kScalars];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ChainModel& M = a.m;
  const int NR = a.w.tile_rows, LB = a.w.lb;
  const int n_tiles = a.w.n_tiles;
  const uint32_t img_bytes = 32u * LB;
  const int n_rounds = (n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctas();
  const uint16_t cmask = static_cast<uint16_t>((1u << csize) - 1u);
  const uint32_t slice = kStageBytes / csize;
  const float zscale = (M.act == ACT_SIN) ? M.w0 : 1.f;

  griddep_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < kTBwdStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], csize); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], kTCompute); mbar_init(&acc_full[i], 1); }
    mbar_init(&img_full, kTCompute);
    mbar_fence_init();
  }
  for (int i = tid; i < M.out_f * kWidth; i += kTThreads) c_wlast[i] = zscale * a.params[M.w_off[M.n_gemm] + i];
  if (warp == 2) tmem_alloc<512>(&tmem_base_s);
  griddep_wait();            // forward kernel complete: tile partials, loss pieces and act' images are final
  tc_fence_before();
  cluster_sync_all();        // (also the CTA barrier) peers' barriers are initialised before anybody multicasts
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  reduce_step_scalars(reinterpret_cast<const float*>(a.ws + a.w.part_off), a.w.n_tiles, a.loss, a.m.out_f, a.bs_k, a.hyper, a.step,
                      blockIdx.x == 0 ? reinterpret_cast<float*>(a.ws + a.w.scal_off) : nullptr, sc);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: act' images + dgrad weight stages
    if (lane == 0) {
      uint32_t it = 0, dq = 0;
      auto load_d = [&](int l, int tile) {
        const uint32_t slot = dq & 1;
        mbar_wait(&d_empty[slot], ((dq >> 1) & 1) ^ 1);
        ++dq;
        mbar_arrive_expect_tx(&d_full[slot], img_bytes);
        bulk_g2s(Db + slot * kTYBytes, a.ws + a.w.d_off[l] + static_cast<size_t>(tile) * img_bytes, img_bytes, &d_full[slot]);
      };
      for (int rd = 0; rd < n_rounds; ++rd) {
        const int tile = min(static_cast<int>(blockIdx.x) + rd * static_cast<int>(gridDim.x), n_tiles - 1);
        load_d(M.n_gemm - 1, tile);
        for (int l = M.n_gemm - 1; l >= 1; --l) {
          load_d(l - 1, tile);                              // ahead of the weight stages: ready when the epilogue needs it
          const uint8_t* src = a.wpack + M.wd_off[l] + crank * slice;
          for (int s = 0; s < kWidth / kStageK; ++s, ++it) {
            const uint32_t slot = it % kTBwdStages, ph = (it / kTBwdStages) & 1;
            mbar_wait(&w_empty[slot], ph ^ 1);              // released by the MMAs of every CTA of the cluster
            mbar_arrive_expect_tx(&w_full[slot], kStageBytes);
            bulk_g2s_mc(wring + slot * kStageBytes + crank * slice, src + static_cast<size_t>(s) * kStageBytes, slice, &w_full[slot], cmask);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: dH_{l-1}^T = W_l^T(packed) * dZ_l^T
    const uint32_t idesc = umma_idesc_f16(kTileM, NR, false, false);
    const uint32_t z_s = smem_u32(Zb), wring_s = smem_u32(wring);
    uint32_t it = 0, iq = 0;
    for (int rd = 0; rd < n_rounds; ++rd) {
      uint32_t zi = 0;                                      // dZ of the top layer lives in Z[0]
      for (int l = M.n_gemm - 1; l >= 1; --l, zi ^= 1) {
        mbar_wait(&img_full, iq & 1); ++iq;
        tc_fence_after();
        for (int s = 0; s < kWidth / kStageK; ++s, ++it) {
          const uint32_t slot = it % kTBwdStages;
          mbar_wait(&w_full[slot], (it / kTBwdStages) & 1);
          tc_fence_after();
          __syncwarp();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t da = umma_smem_desc(wring_s + slot * kStageBytes + h * 2048 + kk * 8192, 4096, 128);
                const uint64_t db = umma_smem_desc(z_s + zi * kTYBytes + (s * 4 + kk * 2) * LB, LB, 128);
                umma_f16(tmem + ((l & 1) * 2 + h) * kTAccStride, da, db, idesc, (s | kk) != 0);
              }
            umma_commit_mc(&w_empty[slot], cmask);
            if (s == kWidth / kStageK - 1) umma_commit(&acc_full[l & 1]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps (16)
    const int ct = tid - 128;
    const int q = warp & 3, j = (warp - 4) >> 2, h = j & 1, ch = j >> 1;
    const int f = 128 * h + 32 * q + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    const int c_lo = ch * (NR >> 1), c_hi = c_lo + (NR >> 1);
    const float S = sc[SC_SCALE], cA = sc[SC_CA], cB = sc[SC_CB];
    const size_t f_off = static_cast<size_t>(f >> 3) * LB + (f & 7) * 2;
    uint32_t dq = 0, acc_ph[2] = {0, 0};
    for (int rd = 0; rd < n_rounds; ++rd) {
      const int tile = min(static_cast<int>(blockIdx.x) + rd * static_cast<int>(gridDim.x), n_tiles - 1);
      const int r0g = tile * NR;
      if (ct == 0) bulk_wait_read0();
      named_bar_sync(1, kTCompute);
      // ---- dZ_last (fp32) and its padded fp16 image for the last layer's wgrad unit
      if (ct < NR) {
        const int grow = r0g + ct;
        float dz[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
        if (grow < a.bs) {
          if (a.dout) {
            for (int o = 0; o < M.out_f; ++o) dz[o] = S * a.dout[static_cast<size_t>(grow) * M.out_f + o];
          } else {
            const float4 g = reinterpret_cast<const float4*>(a.ws + a.w.g_off)[static_cast<size_t>(tile) * NR + ct];
            dz[0] = S * (cA * g.x + cB * g.z);
            dz[1] = S * (cA * g.y + cB * g.w);
          }
        }
        *reinterpret_cast<float4*>(dzl + 4 * ct) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        uint8_t* zl = a.ws + a.w.dzlast_off + static_cast<size_t>(tile) * (2 * LB);
        st_global_v4(zl + ct * 16, make_uint4(pack_h2(dz[0], dz[1]), pack_h2(dz[2], dz[3]), 0u, 0u));
        st_global_v4(zl + LB + ct * 16, make_uint4(0u, 0u, 0u, 0u));
      }
      named_bar_sync(1, kTCompute);
      // ---- layers, top down.  l == n_gemm-1: CUDA-core product with W_last; below: TMEM accumulators
      uint32_t zi = 0;
      for (int l = M.n_gemm - 1; l >= 0; --l) {
        const bool from_last = (l == M.n_gemm - 1);
        if (!from_last) {
          mbar_wait(&acc_full[(l + 1) & 1], acc_ph[(l + 1) & 1]);
          acc_ph[(l + 1) & 1] ^= 1;
          tc_fence_after();
          zi ^= 1;
          if (ct == 0) bulk_wait_read0();                   // the image about to be overwritten has left shared memory
          named_bar_sync(1, kTCompute);
        }
        const uint32_t dslot = dq & 1;
        mbar_wait(&d_full[dslot], (dq >> 1) & 1);
        ++dq;
        uint8_t* zo = Zb + zi * kTYBytes + f_off;
        const uint8_t* dd = Db + dslot * kTYBytes + f_off;
        const uint32_t acc = tmem + t_lane + (((l + 1) & 1) * 2 + h) * kTAccStride;
        float wl[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
        if (from_last)
          for (int o = 0; o < M.out_f; ++o) wl[o] = c_wlast[o * kWidth + f];
        for (int c8 = c_lo; c8 < c_hi; c8 += 8) {
          float v[8];
          if (from_last) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 dz = *reinterpret_cast<const float4*>(dzl + 4 * (c8 + i));
              v[i] = dz.x * wl[0] + dz.y * wl[1] + dz.z * wl[2] + dz.w * wl[3];
            }
          } else {
            tmem_ld8(acc + c8, v);
            tmem_ld_wait();
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dv = __half2float(*reinterpret_cast<const __half*>(dd + (c8 + i) * 16));
            st_shared_h(zo + (c8 + i) * 16, v[i] * dv);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        if (l >= 1) mbar_arrive(&img_full);
        mbar_arrive(&d_empty[dslot]);
        named_bar_sync(1, kTCompute);
        if (ct == 0) {
          bulk_s2g(a.ws + a.w.dz_off[l] + static_cast<size_t>(tile) * img_bytes, Zb + zi * kTYBytes, img_bytes);
          bulk_commit();
        }
      }
    }
    if (ct == 0) bulk_wait0();
  }
  tc_fence_before();
  cluster_sync_all();        // no CTA leaves while a peer may still multicast into its ring or arrive on its barriers
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// Cluster size of the weight multicast: 4 where the device can co-schedule the whole grid as clusters of 4 (148 SMs in GPCs of
// uneven size: some cannot), else 2, else single CTAs.  INR_CHAIN_T_CLUSTER overrides.  Cached per grid size: the occupancy
// query stays out of the step (and out of graph capture).
template <typename Kern>
static int pick_cluster(Kern kern, int smem, int ctas, int* cache) {
  if (ctas < 1 || ctas > 1023) return 1;
  if (cache[ctas]) return cache[ctas];
  int pick = 1;
  const char* e = std::getenv("INR_CHAIN_T_CLUSTER");
  if (e) {
    const int v = std::atoi(e);
    pick = (v == 4 || v == 2) ? v : 1;
  } else {
    for (int c = 4; c >= 2 && pick == 1; c >>= 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((ctas + c - 1) / c * c); cfg.blockDim = dim3(kTThreads); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n * c >= static_cast<int>(cfg.gridDim.x)) pick = c;
      else cudaGetLastError();
    }
  }
  cache[ctas] = pick;
  return pick;
}

cudaError_t launch_chain_fwd_t(const FwdArgs& a, int n_sm, cudaStream_t stream) {
  const int ctas = a.w.n_tiles < n_sm ? a.w.n_tiles : n_sm;
  if (ctas <= 0) return cudaSuccess;
  static bool attr_done = false;
  static int cache[2][1024];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(chain_fwd_t_kernel<ACT_SIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTFwdSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chain_fwd_t_kernel<ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTFwdSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const bool sine = a.m.act == ACT_SIN;
  auto kern = sine ? chain_fwd_t_kernel<ACT_SIN> : chain_fwd_t_kernel<ACT_RELU>;
  const int c = pick_cluster(kern, kTFwdSmem, ctas, cache[sine ? 0 : 1]);
  const int grid = (ctas + c - 1) / c * c;
  return launch_clustered(kern, dim3(grid), dim3(kTThreads), kTFwdSmem, stream, a, c, false);
}

cudaError_t launch_chain_bwd_t(const BwdArgs& a, int n_sm, cudaStream_t stream) {
  const int ctas = a.w.n_tiles < n_sm ? a.w.n_tiles : n_sm;
  if (ctas <= 0) return cudaSuccess;
  static bool attr_done = false;
  static int cache[1024];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(chain_bwd_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTBwdSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int c = pick_cluster(chain_bwd_t_kernel, kTBwdSmem, ctas, cache);
  const int grid = (ctas + c - 1) / c * c;
  return launch_clustered(chain_bwd_t_kernel, dim3(grid), dim3(kTThreads), kTBwdSmem, stream, a, c, true);
}

}  // namespace inr
