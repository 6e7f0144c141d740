// Streaming layer GEMM with fused epilogues ("lgemm").  A work item is (row tile, N-block); it accumulates one or two
// GEMM segments  acc[:, col_s : col_s + nt] = sum_passes A_s[128 x K_s] * B_s[nt x K_s]^T  with A and B streamed from
// HBM operand images through a bulk-TMA ring, tcgen05.mma kind::f16 with fp32 accumulation in TMEM.  Two 256-column
// accumulators ping-pong between work items so the epilogue of item i overlaps the MMAs of item i+1.  Persistent
// CTAs (one per SM) walk the items.
//
// The WIRE modes run the GEMMs of all hidden layers as one launch (LGemmArgs::chain): items are ordered layer-major,
// dealt round-robin to the CTAs, and item (layer, tile, *) starts once both N-blocks of (layer - 1, tile) have been
// stored (per-(layer, tile) counters in global memory).
//
// Epilogues:
//   LG_WIRE_FWD   : complex Gabor wavelet  y = exp(j w z - |s z|^2),  z = a + jb = acc + bias
//                   (reference src/models/networks.py:199-204) -> fp16 hi/lo operand images of the next layer plus
//                   the fp16 (a|b) image the backward pass needs.  3-pass split GEMM, nt = 192 (96 features x (a,b)).
//   LG_WIRE_DGRAD : dL/d(a,b) of the previous layer from dL/dh = acc:
//                   P = Re(conj(g) y), Q = Im(conj(g) y);  dza = -2 s^2 a P - w Q;  dzb = -(w + 2 s^2 b) P.  1 pass.
//   LG_MFN_FWD    : multiplicative filter stage  z_i = sin(x Om_i^T + phi_i) * (z_{i-1} W_i^T + b_i)
//                   (reference src/models/mfn.py:34-38,57-58; BoundedLinear :281-286 as a row mask on the linear term).
//                   Two segments (filter GEMM, linear GEMM), nt = 128; stage 0 has the filter segment only.
//   LG_MFN_DGRAD  : dz_{i-1} = dh_i W_i (+ head gradient), then dh_{i-1} = dz g, dp_{i-1} = dz h cos(p).
//   LG_W2D_FWD / LG_W2D_DGRAD : WIRE2D layers (reference wire2d.py:38-60), two linears per layer: forward item = 64
//                   features x (a, b, c, d) columns, y = exp(j w (a+jb)) exp(-s^2 (a^2+b^2+c^2+d^2)); dgrad item = 128 input
//                   features x (re, im), the epilogue turns dL/dh into the four pre-activation gradients of the layer below.
//   LG_GABOR_E    : Gabor envelope E = exp(-gamma/2 (|x|^2 + |mu|^2 - 2 x mu^T)) (reference mfn.py:117-131) as an fp16
//                   image; MFN_FWD then stores f = sin(p) E and cos(p) E in place of sin / cos, which makes the
//                   backward epilogue of a Gabor stage identical to the Fourier one plus the q = dL/df * f image.
#include <cstdlib>
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "wire.cuh"

namespace inr {

#define LG_TRACE(slot) do { if (a.trace && (slot) < 16) a.trace[64 + blockIdx.x * 32 + (slot)] = global_ns(); } while (0)

constexpr int kLgMaxSlots = 12;
constexpr int kLgRingBytes = 5 * 40960;                                // 204800
constexpr int kLgComputeThreads = 512;
constexpr int kLgThreads = 128 + kLgComputeThreads;
constexpr int kLgSmem = kLgRingBytes + 1024;
// BRES kernels (weights resident, see lgemm_kernel): 208 KB, as much as fits next to the static bias / barrier arrays
constexpr int kLgRingBytesBres = 212992;
constexpr int kLgSmemBres = kLgRingBytesBres + 1024;
// (A fifth K = 48 slot for the forward chain -- 210 KB of ring, all an SM has next to the static arrays -- was measured and
// changes nothing: 78.9 us against 78.6 us.)
template <int MODE, int PAIR, int BRES> struct LgRing {
  static constexpr int bytes = BRES ? kLgRingBytesBres : kLgRingBytes;
  static constexpr int smem = BRES ? kLgSmemBres : kLgSmem;
};


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
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
}

// PAIR = 1: the CTAs of a cluster of two work on two neighbouring row tiles of the same (layer, N-block) with ONE
// tcgen05.mma.cta_group::2 stream (M = 256): each CTA streams its own A tile but only HALF of the weight rows, which cuts the
// operand bytes an SM ingests per flop by 30 % (WIRE: 28 KB instead of 40 KB per ring slot, 7 slots deep instead of 5).
//
// BRES = 1 (WIRE chains on CTA pairs): the weight block of the (layer, N-block) a pair is working on stays RESIDENT in shared
// memory and only the A tiles stream through the ring.  The chains are bound by the chip's L2 -> SM bandwidth (measured: MMA
// warp waiting for operands 50 % of the time at 8.2 TB/s of bulk-copy traffic), and a pair's items of one layer all use
// the same N-block (items are dealt with an even stride), so the 147 KB (forward, hi + lo) / 74 KB (dgrad) of weights
// that every item used to pull again are now fetched once per layer: L2 -> SM bytes per item 343 -> 196 KB forward, 170 -> 96 KB
// dgrad.  b_full / b_empty hand the resident block between producer and MMA warp exactly like a ring slot.
// BRES = 2 (3-pass forward): only the hi image is resident, the lo image keeps streaming with the A tiles -- 270 KB per item, but
// 139 KB of ring instead of 65 KB (the ring depth is what covers the L2 round trip of the A tiles).
// BULK = 1 (WIRE forward chain, opt-in: see launch_lgemm for the measurement): the epilogue leaves its H_hi / H_lo images (last hidden layer: H_hi and the (a | b) image)
// through a 32 KB staging block behind the ring and 8 KB bulk stores issued by the otherwise idle TMEM-allocator warp, instead
// of 16-byte st.global from 512 threads: global stores from the SM's threads throttle the SM's own bulk loads
// (profiles/r01_microbench_summary.md: 67.9 -> 47.8 B/cycle of TMA ingest next to st.global, 59.1 next to bulk stores).
// The thread -> feature mapping becomes c0 = 32 i + 8 sub, so that the four column groups of an iteration form one
// contiguous 8 KB run per image.
constexpr int kLgStageBytes = 32768;
template <int PASSES, int MODE, int KSTEPS, int PAIR, int BRES = 0, int FIRST = 0, int BULK = 0>
__global__ void __launch_bounds__(kLgThreads, 1) lgemm_kernel(const __grid_constant__ LGemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kLgMaxSlots], empty[kLgMaxSlots], acc_full[2], acc_empty[2], stored[2], b_full, b_empty, stg_full, stg_empty;
  __shared__ uint32_t tmem_base_s;
  constexpr bool kChain = MODE == LG_WIRE_FWD || MODE == LG_WIRE_DGRAD || MODE == LG_W2D_FWD || MODE == LG_W2D_DGRAD;
  constexpr int kBiasFloats = MODE == LG_WIRE_FWD ? kWMaxDepth * kWP : 512;
  __shared__ __align__(16) float s_ba[kBiasFloats], s_bb[kBiasFloats];   // WIRE_FWD: bias re / im per chain layer;  MFN: b_i / phi_i (width <= 512)
  // WIRE_FWD / WIRE_DGRAD (top items): final-layer weights (Wr[0], Wi[0], Wr[1], Wi[1]) per feature
  __shared__ float4 s_lw[(MODE == LG_WIRE_FWD || MODE == LG_WIRE_DGRAD) ? kWP : 1];
  const bool has_top = MODE == LG_WIRE_DGRAD && a.top_w != nullptr;      // chain[0] = backward of the final linear, no GEMM
  __shared__ float s_w0[(FIRST && MODE == LG_WIRE_FWD) ? kWP * 3 : 1];    // WIRE_FWD with the first layer folded in: W0 [c][3]
  const bool has_first = FIRST && MODE == LG_WIRE_FWD && a.first_w != nullptr;   // chain[0] = real first layer, no GEMM
  const bool has_nogemm = has_top || has_first;                          // items of chain[0] use no operands and issue no MMAs

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs of the pair)
  const int cta0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);     // work-item walker of this CTA (pair)
  const int n_walk = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int n_rowgroups = PAIR ? (a.n_tiles + 1) >> 1 : a.n_tiles;    // a pair item covers tiles 2g and 2g + 1
  const int per_layer = n_rowgroups * a.n_nblocks;         // items of one layer
  const int n_layers = kChain ? a.chain_len : 1;
  const int n_items = per_layer * n_layers;
  // A ring slot holds KSTEPS consecutive K=16 steps of every operand (a K=32 stage image is two contiguous halves):
  // few steps per slot = deeper ring of smaller copies, many = fewer barrier round trips for the single producer / MMA threads
  const uint32_t b_rows = static_cast<uint32_t>(a.nt) >> PAIR;           // weight rows this CTA holds (pair: half of them)
  const uint32_t b_bytes = b_rows * 32 * KSTEPS;                         // B part of a slot: KSTEPS x (rows x 16 K x 2 B)
  constexpr uint32_t a_bytes = (kWStageABytes / 2) * KSTEPS;
  const uint32_t a_lo_off = a_bytes;
  const uint32_t b_hi_off = PASSES == 3 ? 2 * a_bytes : a_bytes;
  const uint32_t b_lo_off = b_hi_off + b_bytes;
  constexpr bool kLoStreams = BRES == 2 && PASSES == 3;     // B lo image in the ring slots
  const uint32_t slot_bytes = BRES ? (PASSES == 3 ? 2 * a_bytes : a_bytes) + (kLoStreams ? b_bytes : 0u)
                                   : (PASSES == 3 ? 2 * (a_bytes + b_bytes) : (a_bytes + b_bytes));
  const uint32_t b_lo_ring = 2 * a_bytes;                    // kLoStreams: B lo part of a slot behind A hi / A lo
  // resident weights: this CTA's half of the N-block over the whole K, hi image then (3-pass forward) lo image; ring behind it
  const uint32_t bres_half = BRES ? b_rows * 64 * static_cast<uint32_t>(a.seg[0].k_stages) : 0u;
  const uint32_t bres_bytes = BRES ? ((PASSES == 3 && !kLoStreams) ? 2 : 1) * bres_half : 0u;
  uint8_t* ring = smem + bres_bytes;
  int n_slots = (LgRing<MODE, PAIR, BRES>::bytes - bres_bytes) / slot_bytes;
  if (n_slots > kLgMaxSlots) n_slots = kLgMaxSlots;

  // The next layer GEMM of the stream may be scheduled onto SMs this grid has already left; what runs before
  // griddep_wait() (barrier init, TMEM allocation) touches no global memory, everything after it sees the previous
  // kernel -- and with it every earlier kernel of the step -- complete.
  griddep_launch_dependents();
  if (tid == 0) {
    LG_TRACE(0);
    // pair: the leader's full barrier also takes the peer's "my half of the slot has landed" arrival, its acc_empty the
    // arrivals of both CTAs' epilogue threads
    for (int i = 0; i < kLgMaxSlots; ++i) { mbar_init(&full[i], (PAIR && rank == 0) ? 2 : 1); mbar_init(&empty[i], 1); }
    mbar_init(&b_full, (PAIR && rank == 0) ? 2 : 1); mbar_init(&b_empty, 1);
    // epilogue hand-backs: ONE arrival per warp (lane 0 after __syncwarp) instead of 512 arrivals on one barrier word per item --
    // across the cluster for the peer CTA's acc_empty.  (Timing-neutral on B200: 76.3 against 76.2 us for the forward chain.)
    const uint32_t n_arr = kLgComputeThreads / 32;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], n_arr << PAIR); mbar_init(&stored[i], n_arr + (BULK ? 1 : 0));
    }
    mbar_init(&stg_full, n_arr); mbar_init(&stg_empty, 1);
    mbar_fence_init();
  }
  if (warp == 2) { if (PAIR) tmem_alloc_pair<512>(&tmem_base_s); else tmem_alloc<512>(&tmem_base_s); }
  griddep_wait();
  if (MODE == LG_WIRE_FWD) {
    for (int i = tid; i < n_layers * kWP; i += kLgThreads) {
      const int j = i % kWP;
      const float* bias = a.chain[i / kWP].bias;
      if (has_first && i < kWP) {      // real first layer: real bias
        s_ba[i] = j < a.c_valid ? bias[j] : 0.f;
        s_bb[i] = 0.f;
      } else {
        s_ba[i] = j < a.c_valid ? bias[2 * j] : 0.f;
        s_bb[i] = j < a.c_valid ? bias[2 * j + 1] : 0.f;
      }
    }
    if (has_first) {
      for (int i = tid; i < kWP * 3; i += kLgThreads) s_w0[i] = (i / 3) < a.c_valid ? a.first_w[i] : 0.f;
      if (blockIdx.x == 0 && tid == 0 && a.step_counter) *a.step_counter += 1;
    }
    const float* last_w = a.chain[n_layers - 1].last_w;
    for (int j = tid; j < kWP; j += kLgThreads) {
      float4 lw = make_float4(0.f, 0.f, 0.f, 0.f);
      if (last_w && j < a.c_valid) {
        lw.x = last_w[2 * j]; lw.y = last_w[2 * j + 1];
        if (a.out_f > 1) { lw.z = last_w[2 * (a.c_valid + j)]; lw.w = last_w[2 * (a.c_valid + j) + 1]; }
      }
      s_lw[j] = lw;
    }
  } else if (MODE == LG_WIRE_DGRAD) {
    if (has_top)
      for (int j = tid; j < kWP; j += kLgThreads) {
        float4 lw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < a.c_valid) {
          lw.x = a.top_w[2 * j]; lw.y = a.top_w[2 * j + 1];
          if (a.out_f > 1) { lw.z = a.top_w[2 * (a.c_valid + j)]; lw.w = a.top_w[2 * (a.c_valid + j) + 1]; }
        }
        s_lw[j] = lw;
      }
  } else if (MODE == LG_MFN_FWD) {
    const int width = a.n_nblocks * a.nt;
    for (int j = tid; j < width; j += kLgThreads) {
      s_ba[j] = a.bias ? a.bias[j] : 0.f;
      s_bb[j] = a.act_w0 != 0.f ? a.act_w0 * a.phi[j] : a.phi[j];
    }
  } else if (MODE == LG_GABOR_E) {
    const int width = a.n_nblocks * a.nt;
    for (int j = tid; j < width; j += kLgThreads) {
      s_ba[j] = a.gamma[j];
      s_bb[j] = a.mn[j];
    }
  } else if (MODE == LG_MFN_DGRAD) {
    // head-gradient injection: the head's weight rows (out_f <= 2, width <= 512) staged once per launch instead of two
    // scalar global loads per feature and thread in the epilogue (ncu: 116 M against 47 M warp instructions per launch)
    if (a.head_dout) {
      const int width = a.n_nblocks * a.nt;
      for (int j = tid; j < width; j += kLgThreads) {
        s_ba[j] = a.head_w[j];
        s_bb[j] = a.out_f > 1 ? a.head_w[width + j] : 0.f;
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: the peer's barriers must be initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) LG_TRACE(1);
  uint8_t* const stage = ring + n_slots * slot_bytes;      // BULK: 32 KB behind the last ring slot (checked at launch)

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // The whole warp walks the ring (all lanes poll the barriers, so the warp stays convergent); lane 0 arms the slot's
    // barrier, lanes 0 .. 3 issue one bulk copy each.  A single divergent thread spent ~110 cycles per copy and, with four
    // copies and the waits, more than the 576 tensor cycles of a slot (tools/umma_commit2.cu, in-kernel cycle counters).
    {
      uint32_t slot = 0, ph = 0;
      int cur_b = -1;                  // BRES: (layer, N-block) of the resident weights
      uint32_t bph = 0;
      const bool tr = a.trace != nullptr;
      long long c_flag = 0, c_empty = 0, c_issue = 0, tq = 0;
      const long long tr_c0 = tr ? clock64() : 0;
      const unsigned long long tr_n0 = tr ? global_ns() : 0;
      const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();     // LGemmArgs::l2_policy
      for (int item = cta0; item < n_items; item += n_walk) {
        const int layer = item / per_layer, rem = item - layer * per_layer;
        const int tile = PAIR ? 2 * (rem / a.n_nblocks) + static_cast<int>(rank) : rem / a.n_nblocks, nb = rem % a.n_nblocks;
        const bool phantom = PAIR && tile >= a.n_tiles;       // odd tile count: the last pair's second CTA only lends its B half
        if (kChain && layer > 0 && !phantom) {
          // the A images of this tile are what the previous layer's items (tile, 0 .. n_nblocks-1) stored; every lane that
          // issues copies acquires the counter itself and orders its own async-proxy reads behind it
          if (tr) tq = clock64();
          if (!(a.dbg & 8)) flag_wait_ge(a.chain_flags + static_cast<size_t>(layer - 1) * a.n_tiles + tile, static_cast<unsigned int>(a.n_nblocks));
          if (!(a.dbg & 4)) fence_proxy_async_global();
          if (tr) c_flag += clock64() - tq;
        }
        if (MODE == LG_WIRE_DGRAD && !phantom && lane < 4 && !(a.dbg & 16)) {
          // the epilogue of this item reads 96 KB the forward pass wrote long ago (activation y and pre-activation (a, b) of
          // the layer below: four contiguous 24 KB runs, HBM by now): pull them into L2 while the item's MMAs run, so the
          // epilogue's loads see L2 latency instead of an exposed DRAM round trip per item
          const LGemmLayer& Lq = a.chain[layer];
          const uint8_t* base = (lane < 2 ? Lq.in_y : Lq.in_ab) + static_cast<size_t>(tile) * kWTileBytes +
                                static_cast<size_t>(((lane & 1) ? kWP / 8 : 0) + (kWFeatPerBlock / 8) * nb) * 2048;
          if (!(lane == 3 && Lq.real_first)) {
            if ((a.l2_policy & 2) && lane >= 2) bulk_prefetch_l2_hint(base, (kWFeatPerBlock / 8) * 2048, pol_first);   // (a, b): dead after this item
            else bulk_prefetch_l2(base, (kWFeatPerBlock / 8) * 2048);
          }
        }
        if (has_nogemm && layer == 0) continue;   // top / first-layer items have no operands: nothing to stream, no ring slots used
        if (BRES && layer * a.n_nblocks + nb != cur_b) {
          // new weight block: wait until the MMAs that read the old one are done (commit of the last item that used it), then
          // fetch this CTA's half -- lanes 0-3 a quarter of the hi image each, lanes 4-7 of the lo image
          if (cur_b >= 0) { mbar_wait(&b_empty, bph); bph ^= 1; }
          cur_b = layer * a.n_nblocks + nb;
          if (lane == 0) mbar_arrive_expect_tx(&b_full, bres_bytes);
          __syncwarp();
          if (lane < ((PASSES == 3 && !kLoStreams) ? 8 : 4)) {
            const uint8_t* bsrc = ((lane >> 2) ? a.chain[layer].b_lo : a.chain[layer].b_hi) + (static_cast<size_t>(nb << PAIR) + rank) * bres_half;
            const uint32_t q = bres_half >> 2, off = (lane & 3) * q;
            bulk_g2s(smem + (lane >> 2) * bres_half + off, bsrc + off, q, &b_full);
          }
          __syncwarp();
        }
        for (int sg = 0; sg < a.n_seg; ++sg) {
          const LGemmSeg& S = a.seg[sg];
          const int n_it = S.k_stages * 2 / KSTEPS;
          const uint8_t* a_hi = kChain ? a.chain[layer].a_hi : S.a_hi;
          const uint8_t* a_lo = kChain ? a.chain[layer].a_lo : S.a_lo;
          const uint8_t* b_hi = kChain ? a.chain[layer].b_hi : S.b_hi;
          const uint8_t* b_lo = kChain ? a.chain[layer].b_lo : S.b_lo;
          // this lane's copy of every slot: lane 0 A_hi, 1 B_hi, 2 A_lo, 3 B_lo  (pair layout of a packed N-block:
          // [half][K16 step][k-group][nt / 2 rows][8 K] -- each CTA's half is contiguous)
          const bool is_b = lane & 1, is_lo = (lane & 2) != 0;
          const uint8_t* src = nullptr;
          uint32_t bytes = 0, dst_off = 0;
          if (lane < (PASSES == 3 ? 4 : 2)) {
            if (is_b) {
              if (!BRES || (kLoStreams && is_lo)) {
                src = (is_lo ? b_lo : b_hi) + (static_cast<size_t>(nb << PAIR) + rank) * n_it * b_bytes;
                bytes = b_bytes; dst_off = BRES ? b_lo_ring : (is_lo ? b_lo_off : b_hi_off);
              }
            } else if (!phantom) {
              src = (is_lo ? a_lo : a_hi) + static_cast<size_t>(tile) * S.a_tile_bytes;
              bytes = a_bytes; dst_off = is_lo ? a_lo_off : 0;
            }
          }
          // L2 eviction priority of the lines this lane streams (LGemmArgs::l2_policy bit 0): the hi image of an A tile is read
          // again later in the step -- the forward's H_hi by the dgrad epilogues and by wgrad, the dgrad chain's dZ by wgrad --
          // so its lines are marked evict_last as they pass; the step's working set (~400 MB of images) is three times the L2
          const bool hint = (a.l2_policy & 1) && !is_b && !is_lo && src;
          for (int s = 0; s < n_it; ++s) {
            if (tr) tq = clock64();
            mbar_wait(&empty[slot], ph ^ 1);
            if (tr) { const long long t1 = clock64(); c_empty += t1 - tq; tq = t1; }
            if (a.dbg & 2) {
              if (lane == 0) mbar_arrive(&full[slot]);
            } else {
              if (lane == 0) mbar_arrive_expect_tx(&full[slot], phantom ? slot_bytes - (PASSES == 3 ? 2 : 1) * a_bytes : slot_bytes);
              __syncwarp();
              if (src) {
                if (hint) bulk_g2s_hint(ring + slot * slot_bytes + dst_off, src, bytes, &full[slot], pol_last);
                else bulk_g2s(ring + slot * slot_bytes + dst_off, src, bytes, &full[slot]);
                src += bytes;
              }
            }
            __syncwarp();
            if (++slot == static_cast<uint32_t>(n_slots)) { slot = 0; ph ^= 1; }
            if (tr) c_issue += clock64() - tq;
          }
        }
      }
      if (tr && lane == 0) {   // cycles of the producer warp: waiting for tiles of the previous layer | for free slots | issuing copies
        a.trace[64 + blockIdx.x * 32 + 16] = c_flag; a.trace[64 + blockIdx.x * 32 + 17] = c_empty; a.trace[64 + blockIdx.x * 32 + 18] = c_issue;
        a.trace[64 + blockIdx.x * 32 + 24] = clock64() - tr_c0; a.trace[64 + blockIdx.x * 32 + 25] = global_ns() - tr_n0;   // SM clock = ratio
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (PAIR && rank != 0) {
      // ---------------------------------------------------------------- peer CTA: no MMAs to issue; relay "slot landed" to the leader
      if (lane == 0) {
        uint32_t slot = 0, ph = 0, bph = 0;
        int cur_b = -1;
        for (int item = cta0; item < n_items; item += n_walk) {
          if (has_nogemm && item < per_layer) continue;
          if (BRES) {
            const int layer = item / per_layer, nb = (item - layer * per_layer) % a.n_nblocks;
            if (layer * a.n_nblocks + nb != cur_b) {        // my half of the new weight block has landed: tell the leader
              cur_b = layer * a.n_nblocks + nb;
              mbar_wait(&b_full, bph); bph ^= 1;
              mbar_arrive_cluster(&b_full, 0);
            }
          }
          for (int sg = 0; sg < a.n_seg; ++sg) {
            const int n_it = a.seg[sg].k_stages * 2 / KSTEPS;
            for (int s = 0; s < n_it; ++s) {
              mbar_wait(&full[slot], ph);
              mbar_arrive_cluster(&full[slot], 0);
              if (++slot == static_cast<uint32_t>(n_slots)) { slot = 0; ph ^= 1; }
            }
          }
        }
      }
    } else {
      // the whole warp walks the ring and one elected lane issues: descriptors stay warp-uniform (uniform registers), which
      // takes the per-slot issue cost from ~280 + 50 per MMA cycles (single divergent thread) to ~40 (tools/umma_commit2.cu)
      const uint32_t idesc = umma_idesc_f16(kTileM << PAIR, a.nt, false, false);
      const uint32_t b_lbo = b_rows * 16;
      // descriptors differ only in their 14-bit start-address field (bytes >> 4): build once, add offsets
      const uint64_t da0 = umma_smem_desc(smem_u32(ring), 2048, 128);
      const uint64_t db0 = umma_smem_desc(BRES ? smem_u32(smem) : smem_u32(ring), b_lbo, 128);
      const uint64_t db0r = umma_smem_desc(smem_u32(ring), b_lbo, 128);      // kLoStreams: B lo inside the ring slot
      const uint32_t bk16 = (2 * b_lbo) >> 4;                            // one K=16 step of B: two k-groups of nt x 16 B
      uint32_t slot = 0, ph = 0, n_done = 0, bph = 0;
      int cur_b = -1;
      const bool tr = a.trace != nullptr;
      long long c_acc = 0, c_full = 0, c_mma = 0, tq = 0;
      for (int item = cta0; item < n_items; item += n_walk, ++n_done) {
        const uint32_t ab = n_done & 1, use = n_done >> 1;
        if (tr) tq = clock64();
        mbar_wait(&acc_empty[ab], (use & 1) ^ 1);
        if (tr) c_acc += clock64() - tq;
        if (has_nogemm && item < per_layer) {
          // top / first-layer item: no MMAs; the commit (nothing outstanding for this accumulator) keeps the accumulator barriers of both
          // CTAs in step with the epilogue warps
          tc_fence_after();
          __syncwarp();
          if (elect_one()) { if (PAIR) umma_commit_pair(&acc_full[ab]); else umma_commit(&acc_full[ab]); }
          __syncwarp();
          continue;
        }
        int key = 0, next_key = -1;
        if (BRES) {
          const int layer = item / per_layer, nb = (item - layer * per_layer) % a.n_nblocks;
          key = layer * a.n_nblocks + nb;
          if (item + n_walk < n_items) {
            const int nl = (item + n_walk) / per_layer;
            next_key = nl * a.n_nblocks + ((item + n_walk) - nl * per_layer) % a.n_nblocks;
          }
          if (key != cur_b) {                               // both halves of the new weight block have landed
            cur_b = key;
            if (tr) tq = clock64();
            mbar_wait(&b_full, bph); bph ^= 1;
            if (tr) c_full += clock64() - tq;
          }
        }
        tc_fence_after();
        for (int sg = 0; sg < a.n_seg; ++sg) {
          const LGemmSeg& S = a.seg[sg];
          const uint32_t acc = tmem + ab * 256 + S.acc_col;
          const int n_it = S.k_stages * 2 / KSTEPS;
          for (int s = 0; s < n_it; ++s) {
            if (tr) tq = clock64();
            mbar_wait(&full[slot], ph);
            if (tr) { const long long t1 = clock64(); c_full += t1 - tq; tq = t1; }
            tc_fence_after();
            __syncwarp();
            const uint32_t so = (slot * slot_bytes) >> 4;
            if (elect_one()) {
            if (!(a.dbg & 1))
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {          // K = 16 steps inside the slot
              const uint64_t dah = da0 + so + k * 256;                                            // 4096 B per step
              const uint64_t dbh = BRES ? db0 + static_cast<uint32_t>(s * KSTEPS + k) * bk16 : db0 + so + (b_hi_off >> 4) + k * bk16;
              if (PAIR) umma_f16_pair(acc, dah, dbh, idesc, (s | k) != 0); else umma_f16(acc, dah, dbh, idesc, (s | k) != 0);
              if (PASSES == 3) {
                const uint64_t dal = dah + (a_lo_off >> 4);
                const uint64_t dbl = kLoStreams ? db0r + so + (b_lo_ring >> 4) + k * bk16 : dbh + ((BRES ? bres_half : b_bytes) >> 4);
                if (PAIR) { umma_f16_pair(acc, dal, dbh, idesc, 1); umma_f16_pair(acc, dah, dbl, idesc, 1); }
                else { umma_f16(acc, dal, dbh, idesc, 1); umma_f16(acc, dah, dbl, idesc, 1); }
              }
            }
            if (PAIR) umma_commit_pair(&empty[slot]); else umma_commit(&empty[slot]);
            }
            __syncwarp();
            if (++slot == static_cast<uint32_t>(n_slots)) { slot = 0; ph ^= 1; }
            if (tr) c_mma += clock64() - tq;
          }
        }
        __syncwarp();
        if (elect_one()) {
          if (PAIR) umma_commit_pair(&acc_full[ab]); else umma_commit(&acc_full[ab]);
          // last item on this weight block: its MMAs release the resident weights to both producers
          if (BRES && next_key >= 0 && next_key != key) { if (PAIR) umma_commit_pair(&b_empty); else umma_commit(&b_empty); }
        }
        __syncwarp();
        if (lane == 0) LG_TRACE(2 + 3 * n_done);
      }
      if (tr && lane == 0) {      // cycles of the MMA thread: waiting for a drained accumulator | for landed slots | issuing MMAs + commits
        a.trace[64 + blockIdx.x * 32 + 20] = c_acc; a.trace[64 + blockIdx.x * 32 + 21] = c_full; a.trace[64 + blockIdx.x * 32 + 22] = c_mma;
        a.trace[64 + blockIdx.x * 32 + 23] = n_done;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ BULK: store thread (the TMEM allocator warp is idle here)
    if (BULK && MODE == LG_WIRE_FWD && lane == 0) {
      uint32_t it = 0, n_done = 0;
      int pending = -1;                  // accumulator index of the item whose stores are issued but not yet known complete
      const bool publish = kChain && n_layers > 1;
      for (int item = cta0; item < n_items; item += n_walk, ++n_done) {
        const int layer = item / per_layer, rem = item - layer * per_layer;
        const int tile = PAIR ? 2 * (rem / a.n_nblocks) + static_cast<int>(rank) : rem / a.n_nblocks, nb = rem % a.n_nblocks;
        const int ab = static_cast<int>(n_done & 1);
        if (PAIR && tile >= a.n_tiles) {               // phantom tile: nothing staged
          if (pending >= 0) { bulk_wait0(); if (publish) mbar_arrive(&stored[pending]); pending = -1; }
          if (publish) mbar_arrive(&stored[ab]);
          continue;
        }
        const LGemmLayer& Ly = a.chain[layer];
        const size_t base = static_cast<size_t>(tile) * kWTileBytes + static_cast<size_t>(12 * nb) * 2048;
        constexpr size_t kIm = static_cast<size_t>(kWP / 8) * 2048;
        uint8_t* const d0 = Ly.out_hi + base;
        uint8_t* const d1 = Ly.out_lo ? Ly.out_lo + base : ((a.train && Ly.out_ab) ? Ly.out_ab + base : nullptr);
        for (int i = 0; i < 3; ++i) {
          mbar_wait(&stg_full, it & 1);
          bulk_s2g(d0 + i * 8192, stage, 8192);
          bulk_s2g(d0 + kIm + i * 8192, stage + 8192, 8192);
          if (d1) { bulk_s2g(d1 + i * 8192, stage + 16384, 8192); bulk_s2g(d1 + kIm + i * 8192, stage + 24576, 8192); }
          bulk_commit();
          if (i == 0 && pending >= 0) {                 // every group but the one just committed is complete: the previous item is out
            bulk_wait1();
            if (publish) mbar_arrive(&stored[pending]);
            pending = -1;
          }
          bulk_wait_read0();                            // the staging block has been read
          mbar_arrive(&stg_empty);
          ++it;
        }
        pending = ab;
      }
      bulk_wait0();
      if (pending >= 0 && publish) mbar_arrive(&stored[pending]);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ publisher (layer chains): hands finished tiles to the next layer
    // The epilogue threads arrive on stored[ab] once their stores of an item are issued (release at CTA scope); this one
    // thread then makes them visible device-wide (fence, cumulative over what it acquired) and bumps the (layer, tile)
    // counter the next layer's producers poll -- the store drain is waited out here, off the epilogue's critical path.
    if (kChain && n_layers > 1 && lane == 0) {
      uint32_t n_done = 0;
      for (int item = cta0; item < n_items; item += n_walk, ++n_done) {
        const int layer = item / per_layer, rem = item - layer * per_layer;
        const int tile = PAIR ? 2 * (rem / a.n_nblocks) + static_cast<int>(rank) : rem / a.n_nblocks;
        mbar_wait(&stored[n_done & 1], (n_done >> 1) & 1);
        if (layer + 1 < n_layers && tile < a.n_tiles) {
          __threadfence();
          atomicAdd(a.chain_flags + static_cast<size_t>(layer) * a.n_tiles + tile, 1u);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps (16)
    const int q = warp & 3, sub = (warp - 4) >> 2, row = q * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    const float s2 = a.sigma * a.sigma;
    const bool dgrad = MODE == LG_WIRE_DGRAD || MODE == LG_MFN_DGRAD || MODE == LG_W2D_DGRAD;
    // per-layer power-of-two gradient scales (WIRE's gradient norm grows ~10x per layer towards the input; one global
    // loss scale would saturate the fp16 images of the lower layers)
    float ratio = 1.f, amax = 0.f, s_dst = 1.f;
    if (dgrad && !kChain) {
      s_dst = a.scal[SC_LAYER_SCALE + a.dst_layer];
      ratio = s_dst / a.scal[SC_LAYER_SCALE + a.src_layer];
    }
    // amax of the stored (scaled) values -> next step's scale of that layer; order-independent
    auto flush_amax = [&](int dst_layer) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
      if (lane == 0 && amax > 0.f && isfinite(amax))
        atomicMax(reinterpret_cast<unsigned int*>(const_cast<float*>(a.scal)) + SC_LAYER_AMAX + dst_layer, __float_as_uint(amax));
      amax = 0.f;
    };
    int cur_layer = -1;
    uint32_t n_done = 0, stg_it = 0;
    for (int item = cta0; item < n_items; item += n_walk, ++n_done) {
      const int layer = item / per_layer, rem = item - layer * per_layer;
      const int tile = PAIR ? 2 * (rem / a.n_nblocks) + static_cast<int>(rank) : rem / a.n_nblocks, nb = rem % a.n_nblocks;
      const LGemmLayer& Ly = a.chain[kChain ? layer : 0];
      if (PAIR && tile >= a.n_tiles) {      // phantom tile of an odd tile count: nothing to store, just hand the accumulator back
        mbar_wait(&acc_full[n_done & 1], (n_done >> 1) & 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(&acc_empty[n_done & 1], 0);
          if (kChain && n_layers > 1) mbar_arrive(&stored[n_done & 1]);
        }
        continue;
      }
      if (MODE == LG_W2D_FWD && layer != cur_layer) {
        // WIRE2D: both complex biases of a layer fill the staging arrays, so they are re-staged when the CTA moves on to
        // the next layer (its items are layer-major); the barriers keep slower epilogue warps off the old values
        named_bar_sync(2, kLgComputeThreads);
        for (int j = tid - 128; j < kW2dMaxP; j += kLgComputeThreads) {
          const bool ok = j < a.c_valid;
          s_ba[j] = ok ? Ly.bias[2 * j] : 0.f;             s_bb[j] = ok ? Ly.bias[2 * j + 1] : 0.f;
          s_ba[kW2dMaxP + j] = ok ? Ly.bias2[2 * j] : 0.f; s_bb[kW2dMaxP + j] = ok ? Ly.bias2[2 * j + 1] : 0.f;
        }
        named_bar_sync(2, kLgComputeThreads);
        cur_layer = layer;
      }
      if ((MODE == LG_WIRE_DGRAD || MODE == LG_W2D_DGRAD) && layer != cur_layer) {     // a CTA's items are layer-major: one flush per layer
        if (cur_layer >= 0) flush_amax(a.chain[cur_layer].dst_layer);
        cur_layer = layer;
        s_dst = a.scal[SC_LAYER_SCALE + Ly.dst_layer];
        // top items: dL/dh is built from dz_last, which carries the global scale S
        ratio = s_dst / a.scal[(has_top && layer == 0) ? SC_SCALE : SC_LAYER_SCALE + Ly.src_layer];
      }
      const float w = kChain ? Ly.omega : a.omega;
      const uint32_t ab = n_done & 1, use = n_done >> 1;
      const uint32_t acc = tmem + t_lane + ab * 256;
      if (MODE == LG_W2D_FWD) {
        // ---------------- WIRE2D forward: 64 features per item, accumulator columns [a | b | c | d] (64 each)
        const int P2 = a.p2, pg = P2 >> 3;
        const size_t himg = static_cast<size_t>(tile) * (static_cast<size_t>(kTileM) * 2 * P2 * 2) + row * 16;
        const size_t zimg = static_cast<size_t>(tile) * (static_cast<size_t>(kTileM) * 4 * P2 * 2) + row * 16;
        mbar_wait(&acc_full[ab], use & 1);
        tc_fence_after();
        if (tid == 128) LG_TRACE(3 + 3 * n_done);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c0 = 16 * sub + 8 * i;               // feature inside the block
          const int f0 = kW2dFwdFeat * nb + c0;          // complex feature index (multiple of 8)
          float va[8], vb[8], vc[8], vd[8], yr[8], yi[8];
          tmem_ld8(acc + c0, va); tmem_ld8(acc + 64 + c0, vb); tmem_ld8(acc + 128 + c0, vc); tmem_ld8(acc + 192 + c0, vd);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float za = va[e] + s_ba[f0 + e], zb = vb[e] + s_bb[f0 + e];
            const float zc = vc[e] + s_ba[kW2dMaxP + f0 + e], zd = vd[e] + s_bb[kW2dMaxP + f0 + e];
            va[e] = za; vb[e] = zb; vc[e] = zc; vd[e] = zd;
            const float mag = __expf(-w * zb - s2 * (za * za + zb * zb + zc * zc + zd * zd));
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
          const size_t off_r = himg + static_cast<size_t>(f0 >> 3) * 2048, off_i = himg + static_cast<size_t>(pg + (f0 >> 3)) * 2048;
          st_global_v4(Ly.out_hi + off_r, rh); st_global_v4(Ly.out_hi + off_i, ih);
          st_global_v4(Ly.out_lo + off_r, rl); st_global_v4(Ly.out_lo + off_i, il);
          if (a.train) {
            const size_t z0 = zimg + static_cast<size_t>(f0 >> 3) * 2048, zs = static_cast<size_t>(pg) * 2048;
            st_global_v4(Ly.out_ab + z0, pack8(va)); st_global_v4(Ly.out_ab + z0 + zs, pack8(vb));
            st_global_v4(Ly.out_ab + z0 + 2 * zs, pack8(vc)); st_global_v4(Ly.out_ab + z0 + 3 * zs, pack8(vd));
          }
        }
      } else if (MODE == LG_W2D_DGRAD) {
        // ---------------- WIRE2D dgrad: 128 input features per item, accumulator columns [dL/dRe h | dL/dIm h] (128 each)
        const int P2 = a.p2, pg = P2 >> 3;
        const size_t himg = static_cast<size_t>(tile) * (static_cast<size_t>(kTileM) * 2 * P2 * 2) + row * 16;
        const size_t zimg = static_cast<size_t>(tile) * (static_cast<size_t>(kTileM) * 4 * P2 * 2) + row * 16;
        const size_t zs = static_cast<size_t>(pg) * 2048;
        mbar_wait(&acc_full[ab], use & 1);
        tc_fence_after();
        if (tid == 128) LG_TRACE(3 + 3 * n_done);
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int c0 = 32 * sub + 8 * i;
          const int f0 = kW2dBwdFeat * nb + c0;
          const size_t z0 = zimg + static_cast<size_t>(f0 >> 3) * 2048;
          const uint4 yr4 = ld_global_nc_v4(Ly.in_y + himg + static_cast<size_t>(f0 >> 3) * 2048);
          const uint4 yi4 = ld_global_nc_v4(Ly.in_y + himg + static_cast<size_t>(pg + (f0 >> 3)) * 2048);
          const uint4 a4 = ld_global_nc_v4(Ly.in_ab + z0), c4 = ld_global_nc_v4(Ly.in_ab + z0 + 2 * zs);
          uint4 b4 = make_uint4(0u, 0u, 0u, 0u), d4 = make_uint4(0u, 0u, 0u, 0u);
          if (!Ly.real_first) { b4 = ld_global_nc_v4(Ly.in_ab + z0 + zs); d4 = ld_global_nc_v4(Ly.in_ab + z0 + 3 * zs); }
          float gr[8], gi[8];
          tmem_ld8(acc + c0, gr);
          tmem_ld8(acc + 128 + c0, gi);
          tmem_ld_wait();
          float yr[8], yi[8], za[8], zb[8], zc[8], zd[8], da[8], db[8], dc[8], dd[8];
          unpack8(yr4, yr); unpack8(yi4, yi); unpack8(a4, za); unpack8(b4, zb); unpack8(c4, zc); unpack8(d4, zd);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float Pp = gr[e] * yr[e] + gi[e] * yi[e];
            const float Q = gr[e] * yi[e] - gi[e] * yr[e];
            da[e] = ratio * (-2.f * s2 * za[e] * Pp - w * Q);
            db[e] = Ly.real_first ? 0.f : ratio * (-(w + 2.f * s2 * zb[e]) * Pp);
            dc[e] = ratio * (-2.f * s2 * zc[e] * Pp);
            dd[e] = Ly.real_first ? 0.f : ratio * (-2.f * s2 * zd[e] * Pp);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(da[e]), fabsf(db[e])), fmaxf(fabsf(dc[e]), fabsf(dd[e]))));
          }
          st_global_v4(Ly.out_dz + z0, pack8(da)); st_global_v4(Ly.out_dz + z0 + zs, pack8(db));
          st_global_v4(Ly.out_dz + z0 + 2 * zs, pack8(dc)); st_global_v4(Ly.out_dz + z0 + 3 * zs, pack8(dd));
        }
      } else if (MODE == LG_WIRE_FWD || MODE == LG_WIRE_DGRAD) {
        // The epilogue is bound by instruction issue (ncu: issue slots busy 57 % of the forward chain's cycles, ~1800 warp
        // instructions per item and warp before this was trimmed), so the per-feature code is kept minimal: image pointers and
        // constants hoisted to the item, biases by 16-byte shared loads, exp as one ex2 with log2(e) folded into the
        // constants, the live-feature mask only in the one group that straddles c_valid.
        // this thread's three k-groups: real part at group 12 nb + 3 sub + i, imaginary part 24 groups further
        // (BULK: k-groups 12 nb + 4 i + sub, features 96 nb + 32 i + 8 sub + e)
        constexpr int kGi = BULK ? 4 : 1, kGs = BULK ? 1 : 3;           // k-group stride of the iteration index / of sub
        const size_t img = static_cast<size_t>(tile) * kWTileBytes + row * 16 + static_cast<size_t>(12 * nb + kGs * sub) * 2048;
        constexpr size_t kImOff = static_cast<size_t>(kWP / 8) * 2048;
        constexpr size_t kIt = static_cast<size_t>(kGi) * 2048;         // image bytes between this thread's iterations
        const int fb = kWFeatPerBlock * nb + 8 * kGs * sub;  // first complex feature of this thread (multiple of 8)
        // dgrad: the saved activations do not depend on the accumulator -- fetch all of them before waiting for the MMAs
        uint4 pre[3][4];
        const bool real_first = MODE == LG_WIRE_DGRAD && Ly.real_first != 0;
        if (MODE == LG_WIRE_DGRAD) {
          const uint8_t* const py = Ly.in_y + img;
          const uint8_t* const pab = Ly.in_ab + img;
#ifdef INR_LGEMM_EXPERIMENTS
          if (a.dbg & 256) {      // timing experiment: no saved-activation loads (results are wrong)
#pragma unroll
            for (int i = 0; i < 3; ++i) pre[i][0] = pre[i][1] = pre[i][2] = pre[i][3] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
          } else
#endif
          if (a.l2_policy & 2) {  // the (a, b) image is dead after this item -> its lines leave L2 first
            const uint64_t pf = l2_policy_evict_first();
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              pre[i][0] = ld_global_nc_v4(py + i * 2048); pre[i][1] = ld_global_nc_v4(py + kImOff + i * 2048);
              pre[i][2] = ld_global_nc_v4_hint(pab + i * 2048, pf);
              pre[i][3] = real_first ? make_uint4(0u, 0u, 0u, 0u) : ld_global_nc_v4_hint(pab + kImOff + i * 2048, pf);
            }
          } else
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            pre[i][0] = ld_global_nc_v4(py + i * 2048); pre[i][1] = ld_global_nc_v4(py + kImOff + i * 2048);
            pre[i][2] = ld_global_nc_v4(pab + i * 2048);
            pre[i][3] = real_first ? make_uint4(0u, 0u, 0u, 0u) : ld_global_nc_v4(pab + kImOff + i * 2048);
          }
        }
        const bool top = MODE == LG_WIRE_DGRAD && has_top && layer == 0;
        float dzo0 = 0.f, dzo1 = 0.f;                       // top items: S * dL/dout of this thread's row
        if (top) {
          const int grow = tile * kTileM + row;
          const float S = a.scal[SC_SCALE];
          if (grow < a.bs) {
            if (a.top_dout) {
              dzo0 = S * a.top_dout[static_cast<size_t>(grow) * a.out_f];
              if (a.out_f > 1) dzo1 = S * a.top_dout[static_cast<size_t>(grow) * a.out_f + 1];
            } else {
              const float cA = a.scal[SC_CA], cB = a.scal[SC_CB];
              const float4 g = *reinterpret_cast<const float4*>(a.top_g + static_cast<size_t>(grow) * 4);
              dzo0 = S * (cA * g.x + cB * g.z);
              dzo1 = S * (cA * g.y + cB * g.w);
            }
          }
          if (nb == 0 && sub == 0) {                        // padded dz_last image for the final layer's wgrad
            uint8_t* zl = a.top_dzlast + static_cast<size_t>(tile) * kDzLastBytes;
            st_global_v4(zl + row * 16, make_uint4(pack_h2(dzo0, dzo1), 0u, 0u, 0u));
            st_global_v4(zl + 2048 + row * 16, make_uint4(0u, 0u, 0u, 0u));
          }
        }
        mbar_wait(&acc_full[ab], use & 1);
        tc_fence_after();
        if (tid == 128) LG_TRACE(3 + 3 * n_done);
        float o0 = 0.f, o1 = 0.f;                           // this thread's share of the final linear (last hidden layer only)
        const bool first = FIRST && MODE == LG_WIRE_FWD && has_first && layer == 0;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;                 // first-layer items: this row's coordinates
        if (FIRST && first) {
          const int grow = tile * kTileM + row;
          if (grow < a.bs) {
            const float* c = a.coords + (static_cast<size_t>(a.row_offset ? *a.row_offset : 0) + grow) * 3;
            x0 = c[0]; x1 = c[1]; x2 = c[2];
          }
          if (a.train && nb == 0 && sub == 0) {             // coordinate image for the first layer's wgrad
            const float xs[3] = {x0, x1, x2};
            float v[8];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float h = __half2float(__float2half_rn(xs[c]));
              v[c] = h; v[4 + c] = xs[c] - h;
            }
            v[3] = grow < a.bs ? 1.f : 0.f; v[7] = 0.f;
            uint8_t* xi = a.ximg + static_cast<size_t>(tile) * kDzLastBytes;
            st_global_v4(xi + row * 16, pack8(v));
            st_global_v4(xi + 2048 + row * 16, make_uint4(0u, 0u, 0u, 0u));
          }
        }
        if (MODE == LG_WIRE_FWD && !FIRST && !BULK) {
          // Default forward epilogue: the plain per-feature form.  A leaner instruction stream (the branch below: ~800 instead
          // of ~1300 warp instructions per item and warp) measures 2.8 us SLOWER on this chain (85.1 / 85.9 / 86.1 us against
          // 88.8 / 88.1 / 88.8 us, libraries alternated on one box): the chain is bound by operand ingest next to the
          // epilogue's stores, and sixteen warps that finish their arithmetic sooner fire those stores in denser bursts.
          const size_t img0 = static_cast<size_t>(tile) * kWTileBytes + row * 16;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c0 = 24 * sub + 8 * i;
            const int f0 = kWFeatPerBlock * nb + c0;
            const size_t off_r = img0 + static_cast<size_t>(f0 >> 3) * 2048;
            const size_t off_i = img0 + static_cast<size_t>((kWP + f0) >> 3) * 2048;
            float va[8], vb[8];
            tmem_ld8(acc + c0, va);
            tmem_ld8(acc + kWFeatPerBlock + c0, vb);
            tmem_ld_wait();
            float yr[8], yi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float za = va[e] + s_ba[layer * kWP + f0 + e], zb = vb[e] + s_bb[layer * kWP + f0 + e];
              va[e] = za; vb[e] = zb;
              const bool live = (f0 + e) < a.c_valid;
              if (a.dbg & 64) { yr[e] = live ? za * zb : 0.f; yi[e] = live ? za + zb : 0.f; continue; }   // timing experiment: no MUFU
              const float mag = __expf(-w * zb - s2 * (za * za + zb * zb));
              const float ang = w * za;
              yr[e] = live ? mag * fast_cos(ang) : 0.f;
              yi[e] = live ? mag * fast_sin(ang) : 0.f;
            }
            if (MODE == LG_WIRE_FWD && Ly.out_part) {       // features in ascending order: fixed summation order
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float4 lw = s_lw[f0 + e];            // same address in every lane: broadcast
                o0 = fmaf(yr[e], lw.x, fmaf(-yi[e], lw.y, o0));
                o1 = fmaf(yr[e], lw.z, fmaf(-yi[e], lw.w, o1));
              }
            }
            uint4 rh, rl, ih, il;
            split_h2(yr[0], yr[1], rh.x, rl.x); split_h2(yr[2], yr[3], rh.y, rl.y);
            split_h2(yr[4], yr[5], rh.z, rl.z); split_h2(yr[6], yr[7], rh.w, rl.w);
            split_h2(yi[0], yi[1], ih.x, il.x); split_h2(yi[2], yi[3], ih.y, il.y);
            split_h2(yi[4], yi[5], ih.z, il.z); split_h2(yi[6], yi[7], ih.w, il.w);
            if (a.dbg & 32) {      // timing experiment: one store instead of six
              rh.x ^= rl.x ^ ih.x ^ il.x; rh.y ^= rl.y ^ ih.y ^ il.y; rh.z ^= rl.z ^ ih.z ^ il.z; rh.w ^= rl.w ^ ih.w ^ il.w;
              const uint4 pa = pack8(va), pb = pack8(vb);
              rh.x ^= pa.x ^ pb.x; rh.y ^= pa.y ^ pb.y; rh.z ^= pa.z ^ pb.z; rh.w ^= pa.w ^ pb.w;
              st_global_v4(Ly.out_hi + off_r, rh);
              continue;
            }
            st_global_v4(Ly.out_hi + off_r, rh); st_global_v4(Ly.out_hi + off_i, ih);
            if (Ly.out_lo) {       // null for the last hidden layer: nothing reads its lo image (the final linear rode along above)
              st_global_v4(Ly.out_lo + off_r, rl); st_global_v4(Ly.out_lo + off_i, il);
            }
            if (a.train) {
              st_global_v4(Ly.out_ab + off_r, pack8(va));
              st_global_v4(Ly.out_ab + off_i, pack8(vb));
            }
          }
          if (Ly.out_part)
            reinterpret_cast<float4*>(Ly.out_part)[(static_cast<size_t>(tile) * kWOutParts + nb * 4 + sub) * kTileM + row] =
                make_float4(o0, o1, 0.f, 0.f);
        } else if (MODE == LG_WIRE_FWD) {
          uint8_t* const p_hi = Ly.out_hi + img;
          uint8_t* const p_lo = Ly.out_lo ? Ly.out_lo + img : nullptr;      // null for the last hidden layer: nothing reads its lo
          uint8_t* const p_ab = a.train ? Ly.out_ab + img : nullptr;        // image (the final linear rides along below)
          uint8_t* const stg = stage + sub * 2048 + row * 16;               // BULK: this thread's 16 bytes in each staged 8 KB run
          const bool ride = Ly.out_part != nullptr;
          const float* const bias_a = s_ba + layer * kWP + fb;
          const float* const bias_b = s_bb + layer * kWP + fb;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c0 = 8 * kGs * sub + 8 * kGi * i;    // feature inside the N-block
            float va[8], vb[8], yr[8], yi[8];
            const int n_live = a.c_valid - (fb + 8 * kGi * i);   // < 8 only in the group that straddles the real width (and beyond it)
            if (FIRST && first) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {                 // z = x W0^T + b0 (real; the bias sits in s_ba of chain layer 0)
                const int f = fb + 8 * kGi * i + e;
                const float za = fmaf(x0, s_w0[3 * f], fmaf(x1, s_w0[3 * f + 1], fmaf(x2, s_w0[3 * f + 2], s_ba[f])));
                va[e] = za; vb[e] = 0.f;
                // the expressions wire_first_kernel evaluates, so that both ways give the same bits
                const float mag = __expf(-s2 * za * za);
                yr[e] = mag * fast_cos(w * za);
                yi[e] = mag * fast_sin(w * za);
              }
            } else {
              tmem_ld8(acc + c0, va);
              tmem_ld8(acc + kWFeatPerBlock + c0, vb);
              const float4 a0 = *reinterpret_cast<const float4*>(bias_a + 8 * kGi * i), a1 = *reinterpret_cast<const float4*>(bias_a + 8 * kGi * i + 4);
              const float4 b0 = *reinterpret_cast<const float4*>(bias_b + 8 * kGi * i), b1 = *reinterpret_cast<const float4*>(bias_b + 8 * kGi * i + 4);
              const float ba[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float za = va[e] + ba[e], zb = vb[e] + bb[e];
                va[e] = za; vb[e] = zb;
#ifdef INR_LGEMM_EXPERIMENTS
                if (a.dbg & 64) { yr[e] = za * zb; yi[e] = za + zb; continue; }   // timing experiment: no MUFU
#endif
                const float mag = __expf(-w * zb - s2 * (za * za + zb * zb));      // the default epilogue's expression: same bits
                const float ang = w * za;
                yr[e] = mag * fast_cos(ang);
                yi[e] = mag * fast_sin(ang);
              }
            }
            if (n_live < 8) {                               // padded features stay exactly zero in every image
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (e >= n_live) { yr[e] = 0.f; yi[e] = 0.f; }
            }
            if (ride) {                                     // features in ascending order: fixed summation order
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float4 lw = s_lw[fb + 8 * kGi * i + e];    // same address in every lane: broadcast
                o0 = fmaf(yr[e], lw.x, fmaf(-yi[e], lw.y, o0));
                o1 = fmaf(yr[e], lw.z, fmaf(-yi[e], lw.w, o1));
              }
            }
            uint4 rh, rl, ih, il;
            split_h2(yr[0], yr[1], rh.x, rl.x); split_h2(yr[2], yr[3], rh.y, rl.y);
            split_h2(yr[4], yr[5], rh.z, rl.z); split_h2(yr[6], yr[7], rh.w, rl.w);
            split_h2(yi[0], yi[1], ih.x, il.x); split_h2(yi[2], yi[3], ih.y, il.y);
            split_h2(yi[4], yi[5], ih.z, il.z); split_h2(yi[6], yi[7], ih.w, il.w);
#ifdef INR_LGEMM_EXPERIMENTS
            if (a.dbg & 32) {      // timing experiment: one store instead of six
              rh.x ^= rl.x ^ ih.x ^ il.x; rh.y ^= rl.y ^ ih.y ^ il.y; rh.z ^= rl.z ^ ih.z ^ il.z; rh.w ^= rl.w ^ ih.w ^ il.w;
              const uint4 pa = pack8(va), pb = pack8(vb);
              rh.x ^= pa.x ^ pb.x; rh.y ^= pa.y ^ pb.y; rh.z ^= pa.z ^ pb.z; rh.w ^= pa.w ^ pb.w;
              st_global_v4(p_hi + i * kIt, rh);
              continue;
            }
#endif
            if (BULK) {
              // staged: [H_hi re | H_hi im | H_lo re | H_lo im], or with no lo image (last hidden layer) [.. | a | b]
              mbar_wait(&stg_empty, (stg_it & 1) ^ 1);
              ++stg_it;
              st_shared_v4(stg, rh); st_shared_v4(stg + 8192, ih);
              if (p_lo) { st_shared_v4(stg + 16384, rl); st_shared_v4(stg + 24576, il); }
              else if (p_ab) { st_shared_v4(stg + 16384, pack8(va)); st_shared_v4(stg + 24576, pack8(vb)); }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(&stg_full);
              if (p_lo && p_ab) { st_global_v4(p_ab + i * kIt, pack8(va)); st_global_v4(p_ab + kImOff + i * kIt, pack8(vb)); }
            } else {
              st_global_v4(p_hi + i * kIt, rh); st_global_v4(p_hi + kImOff + i * kIt, ih);
              if (p_lo) { st_global_v4(p_lo + i * kIt, rl); st_global_v4(p_lo + kImOff + i * kIt, il); }
              if (p_ab) {
                st_global_v4(p_ab + i * kIt, pack8(va));
                if (!(FIRST && first)) st_global_v4(p_ab + kImOff + i * kIt, pack8(vb));   // the real first layer has no b part (dgrad: real_first)
              }
            }
          }
          if (ride)
            reinterpret_cast<float4*>(Ly.out_part)[(static_cast<size_t>(tile) * kWOutParts + nb * 4 + sub) * kTileM + row] =
                make_float4(o0, o1, 0.f, 0.f);
        } else {
          uint8_t* const p_dz = Ly.out_dz + img;
          // da = ratio (-2 s^2 za P - w Q),  db = ratio (-(w + 2 s^2 zb) P)  with P = Re(conj(g) y), Q = Im(conj(g) y)
          const float c1 = -2.f * s2 * ratio, c2 = -w * ratio;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c0 = 24 * sub + 8 * i;
            float va[8], vb[8];
            if (top) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {                 // out = Re(h W^T + b):  dL/d Re(h_j) = dz . Wr_j,  dL/d Im(h_j) = -dz . Wi_j
                const float4 lw = s_lw[fb + 8 * i + e];
                va[e] = fmaf(dzo1, lw.z, fmaf(dzo0, lw.x, 0.f));
                vb[e] = fmaf(-dzo1, lw.w, fmaf(-dzo0, lw.y, 0.f));
              }
            } else {
              tmem_ld8(acc + c0, va);                       // dL/d Re(h)
              tmem_ld8(acc + kWFeatPerBlock + c0, vb);      // dL/d Im(h)
              tmem_ld_wait();
            }
            float yr[8], yi[8], za[8], zb[8], da[8], db[8];
            unpack8(pre[i][0], yr); unpack8(pre[i][1], yi); unpack8(pre[i][2], za); unpack8(pre[i][3], zb);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float P = fmaf(va[e], yr[e], vb[e] * yi[e]);
              const float Q = fmaf(va[e], yi[e], -(vb[e] * yr[e]));
              da[e] = fmaf(c1 * za[e], P, c2 * Q);
              db[e] = fmaf(c1, zb[e], c2) * P;
            }
            if (real_first) {
#pragma unroll
              for (int e = 0; e < 8; ++e) db[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fmaxf(fabsf(da[e]), fabsf(db[e])));
            st_global_v4(p_dz + i * 2048, pack8(da));
            st_global_v4(p_dz + kImOff + i * 2048, pack8(db));
          }
        }
      } else {
        // ---------------- MFN stages: nt = 128 columns per N-block, this warp owns 32 of them (4 steps of 8)
        const size_t img = static_cast<size_t>(tile) * a.feat_tile_bytes + row * 16;
        uint4 pre[4][3];
        if (MODE == LG_MFN_DGRAD) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const size_t off = img + static_cast<size_t>((a.nt * nb + 32 * sub + 8 * i) >> 3) * 2048;
            pre[i][0] = a.in_y ? ld_global_nc_v4(a.in_y + off) : make_uint4(0u, 0u, 0u, 0u);
            pre[i][1] = ld_global_nc_v4(a.in_ab + off);
            pre[i][2] = a.real_first ? make_uint4(0u, 0u, 0u, 0u) : ld_global_nc_v4(a.in_h + off);
          }
        }
        mbar_wait(&acc_full[ab], use & 1);
        tc_fence_after();
        if (tid == 128) LG_TRACE(3 + 3 * n_done);
        const int grow = tile * kTileM + row;
        bool masked = false;            // BoundedLinear: this row's input to the linear was zeroed
        if (a.dist && grow < a.bs) { const float d = a.dist[(a.dist_row_offset ? *a.dist_row_offset : 0) + grow]; masked = (d < a.bound_lo) || (d > a.bound_hi); }
        float hd[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == LG_MFN_DGRAD && a.head_dout && grow < a.bs) {
#pragma unroll
          for (int o = 0; o < kMaxOut; ++o)
            if (o < a.out_f) hd[o] = s_dst * a.head_dout[static_cast<size_t>(grow) * a.head_ld + a.head_col + o];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = 32 * sub + 8 * i;                 // column inside the N-block
          const int f0 = a.nt * nb + c0;                   // feature index
          const size_t off = img + static_cast<size_t>(f0 >> 3) * 2048;
          float vp[8], vh[8];
          if (MODE == LG_MFN_FWD) {
            tmem_ld8(acc + c0, vp);                        // filter pre-activation  x Om^T
            if (a.n_seg == 2) tmem_ld8(acc + a.nt + c0, vh);   // linear  z_{i-1} W^T
            tmem_ld_wait();
            float g[8], c[8], h[8], z[8], env[8];
            if (a.in_e) unpack8(ld_global_nc_v4(a.in_e + off), env);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float p = vp[e] + s_bb[f0 + e];
              if (a.act_w0 != 0.f) {      // wide SIREN / FFN layer: z = act(w0 (W z + b)), c = d z / d(W z + b)
                if (a.act_kind == ACT_RELU) { g[e] = fmaxf(p, 0.f); c[e] = p > 0.f ? 1.f : 0.f; }
                else { g[e] = fast_sin(p); c[e] = a.act_w0 * fast_cos(p); }
              } else {
                g[e] = fast_sin(p); c[e] = fast_cos(p);
              }
              if (a.in_e) { g[e] *= env[e]; c[e] *= env[e]; }     // Gabor: f = sin(p) E, d f / d p = cos(p) E
              h[e] = a.n_seg == 2 ? ((masked ? 0.f : vh[e]) + s_ba[f0 + e]) : 1.f;
              z[e] = g[e] * h[e];
            }
            st_global_v4(a.out_hi + off, pack8(z));
            if (a.train) {
              if (a.out_lo) st_global_v4(a.out_lo + off, pack8(g));
              st_global_v4(a.out_ab + off, pack8(c));
              if (a.n_seg == 2) st_global_v4(a.out_h + off, pack8(h));
            }
          } else if (MODE == LG_GABOR_E) {
            tmem_ld8(acc + c0, vp);                        // x mu^T
            tmem_ld_wait();
            const float xn = a.xn[grow];
            float env[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) env[e] = __expf(-0.5f * s_ba[f0 + e] * (xn + s_bb[f0 + e] - 2.f * vp[e]));
            st_global_v4(a.out_hi + off, pack8(env));
          } else {
            const uint4 g4 = pre[i][0], c4 = pre[i][1], h4 = pre[i][2];
            tmem_ld8(acc + c0, vp);                        // S[src] * dh_i W_i  = S[src] * dz_{i-1} (before heads)
            tmem_ld_wait();
            float g[8], c[8], h[8], dh[8], dp[8], q[8];
            unpack8(g4, g); unpack8(c4, c); unpack8(h4, h);
            float hw0[8], hw1[8];                          // head weights of these 8 features (zeros without a head)
            if (a.head_dout) {
              const float4 w00 = *reinterpret_cast<const float4*>(s_ba + f0), w01 = *reinterpret_cast<const float4*>(s_ba + f0 + 4);
              const float4 w10 = *reinterpret_cast<const float4*>(s_bb + f0), w11 = *reinterpret_cast<const float4*>(s_bb + f0 + 4);
              hw0[0] = w00.x; hw0[1] = w00.y; hw0[2] = w00.z; hw0[3] = w00.w; hw0[4] = w01.x; hw0[5] = w01.y; hw0[6] = w01.z; hw0[7] = w01.w;
              hw1[0] = w10.x; hw1[1] = w10.y; hw1[2] = w10.z; hw1[3] = w10.w; hw1[4] = w11.x; hw1[5] = w11.y; hw1[6] = w11.z; hw1[7] = w11.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float dz = ratio * vp[e];    // rows masked in the source stage arrive as zeros (its stored dh is masked)
              if (a.head_dout) {           // same order as before: output 0, then output 1 (hd[1] = 0 for one output)
                dz = fmaf(hd[0], hw0[e], dz);
                if (a.out_f > 1) dz = fmaf(hd[1], hw1[e], dz);
              }
              if (a.real_first) { dh[e] = 0.f; dp[e] = dz * c[e]; q[e] = dz * g[e]; }
              else { dh[e] = dz * g[e]; dp[e] = dz * h[e] * c[e]; q[e] = dh[e] * h[e]; }
              amax = fmaxf(amax, fmaxf(fabsf(dh[e]), fabsf(dp[e])));
              if (a.out_q) amax = fmaxf(amax, fabsf(q[e]));
            }
            if (a.out_q) st_global_v4(a.out_q + off, pack8(q));      // Gabor: q = dL/df * f (g holds f = sin(p) E)
            if (!a.real_first) {
              if (a.out_dzu) st_global_v4(a.out_dzu + off, pack8(dh));     // unmasked copy: bias gradient
              if (masked) {
#pragma unroll
                for (int e = 0; e < 8; ++e) dh[e] = 0.f;                    // this row fed zeros to the target stage's linear
              }
              st_global_v4(a.out_dz + off, pack8(dh));
            }
            st_global_v4(a.out_dp + off, pack8(dp));
          }
        }
      }
      if (tid == 128) LG_TRACE(4 + 3 * n_done);
      tc_fence_before();
      __syncwarp();              // every lane's TMEM loads have retired and its stores of the item are issued
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(&acc_empty[ab], 0); else mbar_arrive(&acc_empty[ab]);
        if (kChain && n_layers > 1) mbar_arrive(&stored[ab]);     // the warp's stores of the item are issued (publisher warp below)
      }
    }
    if (MODE == LG_WIRE_DGRAD || MODE == LG_W2D_DGRAD) {
      if (cur_layer >= 0) flush_amax(a.chain[cur_layer].dst_layer);
    } else if (dgrad) {
      flush_amax(a.dst_layer);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // pair: the peer's shared memory and barriers stay alive until both are done
  if (warp == 2) { if (PAIR) tmem_dealloc_pair<512>(tmem); else tmem_dealloc<512>(tmem); }
  if (tid == 0) LG_TRACE(15);
}

// INR_PDL=0 launches the GEMM kernels without the programmatic-serialization attribute (plain stream order)
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("INR_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// Chained launches need every CTA resident.  Two fits stepping concurrently on one device (HP search: many small fits per GPU
// on separate streams) must therefore share the SMs explicitly: inr_set_sm_budget(n) caps the grid of every chained launch of
// this process at n SMs, so that the chained kernels of all concurrent fits fit on the chip together (budgets summing to at
// most the SM count); non-chained kernels never wait on other CTAs and simply drain.  0 = the whole chip (one fit at a time).
static int g_sm_budget = 0;
void lgemm_set_sm_budget(int n) { g_sm_budget = n > 0 ? n : 0; }
int lgemm_sm_budget() { return g_sm_budget; }

// WIRE layer chains run on CTA pairs (INR_LGEMM_PAIR=0 keeps single CTAs with the same pair-packed weights: debugging only)
static bool lgemm_pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("INR_LGEMM_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// Which WIRE chains keep the weight block of a pair's current (layer, N-block) resident in shared memory: INR_LGEMM_BRES bit 0 =
// dgrad chain (default on; in-process A/B, tools/ab_variants.py: backward kernels of the bs 25 000 step 117.4 -> 113.5-114.0 us
// with 24 / 32 / 48 KB slots, no gain with 16 KB slots, +15 us with 8 KB slots: fewer, larger copies), bit 1 = forward chain
// (default off: 88.3 us against 79.5 us streaming with everything resident, equal with the hi image resident.  Neither fewer
// L2 -> SM bytes nor a deeper ring -- a fifth slot was tried too -- move the 3-pass forward; what is left as a suspect is the
// shared memory itself: per K = 16 step the MMAs read 21-30 KB of operands in 288 tensor cycles while the bulk copies
// write 14 KB, i.e. 120-150 B per cycle against the 128 B per cycle an SM's shared memory moves).  Read per launch.
static int lgemm_bres_mask() {
  const char* e = std::getenv("INR_LGEMM_BRES");
  return e ? std::atoi(e) : 1;
}

template <int P, int M, int K, int PAIR, int BRES = 0, int FIRST = 0, int BULK = 0>
static cudaError_t lgemm_launch_one(const LGemmArgs& a, int n_sm, cudaStream_t stream) {
  static bool attr = false;
  static int max_clusters = 0;
  auto kern = lgemm_kernel<P, M, K, PAIR, BRES, FIRST, BULK>;
  // BULK: the staging block sits behind the last slot of the WIRE forward ring (K = 48 slots of 43 008 B: four of them leave 32 KB)
  static_assert(!BULK || (M == LG_WIRE_FWD && P == 3 && K == 3 && PAIR == 1 && BRES == 0 &&
                          kLgRingBytes - (kLgRingBytes / 43008) * 43008 >= kLgStageBytes), "staging block does not fit");
  constexpr int kLgSmem = LgRing<M, PAIR, BRES>::smem;
  if (BRES && (a.n_seg != 1 || !PAIR)) return cudaErrorInvalidValue;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kLgSmem);
    if (e != cudaSuccess) return e;
    if (PAIR) {
      // chained items wait on tiles other CTAs produce: every cluster of the grid must be resident at once
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(2 * (n_sm / 2)); q.blockDim = dim3(kLgThreads); q.dynamicSmemBytes = kLgSmem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q);
      if (e != cudaSuccess) return e;
      if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
    }
    attr = true;
  }
  const bool chain = a.mode == LG_WIRE_FWD || a.mode == LG_WIRE_DGRAD || a.mode == LG_W2D_FWD || a.mode == LG_W2D_DGRAD;
  const int rowgroups = PAIR ? (a.n_tiles + 1) / 2 : a.n_tiles;
  const int items = rowgroups * a.n_nblocks * (chain ? a.chain_len : 1);
  if (chain && g_sm_budget > 0 && g_sm_budget < n_sm) n_sm = g_sm_budget < 2 ? 2 : g_sm_budget;
  int walkers = PAIR ? n_sm / 2 : n_sm;                    // never more CTAs than SMs
  if (PAIR && walkers > max_clusters) walkers = max_clusters;
  if (items < walkers) walkers = items;
  if (BRES && walkers > 2) walkers &= ~1;                  // an even item stride keeps a pair on one N-block for a whole layer
  if (walkers <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(walkers << PAIR); cfg.blockDim = dim3(kLgThreads); cfg.dynamicSmemBytes = kLgSmem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n_at = 0;
  if (PAIR) {
    at[n_at].id = cudaLaunchAttributeClusterDimension;
    at[n_at].val.clusterDim.x = 2; at[n_at].val.clusterDim.y = 1; at[n_at].val.clusterDim.z = 1;
    ++n_at;
  }
  if (pdl_enabled()) {
    at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n_at].val.programmaticStreamSerializationAllowed = 1;
    ++n_at;
  }
  cfg.attrs = at; cfg.numAttrs = n_at;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t launch_lgemm(const LGemmArgs& a_in, int n_sm, cudaStream_t stream) {
  LGemmArgs a = a_in;
  {
    // L2 eviction priorities (see the producer): on for the WIRE chains, where they were measured (backward kernels 110.1 ->
    // 106.2 us, forward unchanged); INR_LGEMM_L2 overrides per launch (A/B runs, tools/ab_variants.py)
    const char* env = std::getenv("INR_LGEMM_L2");
    a.l2_policy = env ? std::atoi(env) : ((a.mode == LG_WIRE_FWD || a.mode == LG_WIRE_DGRAD) ? 3 : 0);
  }
  const bool chain = a.mode == LG_WIRE_FWD || a.mode == LG_WIRE_DGRAD || a.mode == LG_W2D_FWD || a.mode == LG_W2D_DGRAD;
  if (chain && (a.chain_len < 1 || a.chain_len > kWMaxDepth || (a.chain_len > 1 && !a.chain_flags))) return cudaErrorInvalidValue;
  if (a.n_tiles <= 0) return cudaSuccess;
  const int expect_passes = (a.mode == LG_WIRE_FWD || a.mode == LG_W2D_FWD) ? 3 : 1;
  if (a.passes != expect_passes) return cudaErrorInvalidValue;
  cudaError_t e;
  switch (a.mode) {
    case LG_WIRE_FWD: {
      const char* ekf = std::getenv("INR_LG_KF");          // K = 16 steps per ring slot (read per launch)
      const int kf = ekf ? std::atoi(ekf) : 3;
      if (!lgemm_pair_enabled()) return cudaErrorNotSupported;
      if (a.first_w) {       // first layer folded in (opt-in): one instantiation, K = 48 slots
        e = lgemm_launch_one<3, LG_WIRE_FWD, 3, 1, 0, 1>(a, n_sm, stream);
        break;
      }
      {                      // opt-in (INR_LG_BULK=1, read per launch): epilogue images through shared memory + bulk stores.
        // Measured on B200, in-process A/B at bs 25 000: 85.2 us against 84.1 us with plain st.global -- the microbenchmark's
        // gain (stores that no longer throttle the bulk loads) is eaten by twelve more copies per item on a copy engine the
        // producer already waits on for 46 % of its time.  Not the default.
        const char* eb = std::getenv("INR_LG_BULK");
        if (kf == 3 && !(lgemm_bres_mask() & 2) && eb && eb[0] == '1') {
          e = lgemm_launch_one<3, LG_WIRE_FWD, 3, 1, 0, 0, 1>(a, n_sm, stream);
          break;
        }
      }
      if (lgemm_bres_mask() & 2) {
        static int kfb = -1;
        if (kfb < 0) { const char* e2 = std::getenv("INR_LG_KFB"); kfb = e2 ? std::atoi(e2) : 2; }
        static int lo = -1;      // INR_LG_LO=1: the weights' lo image streams (BRES = 2)
        if (lo < 0) { const char* e2 = std::getenv("INR_LG_LO"); lo = e2 ? std::atoi(e2) : 1; }
        if (lo) e = kfb == 3 ? lgemm_launch_one<3, LG_WIRE_FWD, 3, 1, 2>(a, n_sm, stream) : lgemm_launch_one<3, LG_WIRE_FWD, 2, 1, 2>(a, n_sm, stream);
        else e = kfb == 1 ? lgemm_launch_one<3, LG_WIRE_FWD, 1, 1, 1>(a, n_sm, stream) : lgemm_launch_one<3, LG_WIRE_FWD, 2, 1, 1>(a, n_sm, stream);
        break;
      }
      e = kf == 6 ? lgemm_launch_one<3, LG_WIRE_FWD, 6, 1>(a, n_sm, stream)
        : kf == 4 ? lgemm_launch_one<3, LG_WIRE_FWD, 4, 1>(a, n_sm, stream)
        : kf == 3 ? lgemm_launch_one<3, LG_WIRE_FWD, 3, 1>(a, n_sm, stream) : lgemm_launch_one<3, LG_WIRE_FWD, 2, 1>(a, n_sm, stream);
      break;
    }
    case LG_WIRE_DGRAD: {
      static int kd = -1;
      if (kd < 0) { const char* e2 = std::getenv("INR_LG_KD"); kd = e2 ? std::atoi(e2) : 4; }
      if (!lgemm_pair_enabled()) return cudaErrorNotSupported;
      if (lgemm_bres_mask() & 1) {
        const char* e2 = std::getenv("INR_LG_KDB");        // K = 16 steps per ring slot (read per launch)
        const int kdb = e2 ? std::atoi(e2) : 8;
        e = kdb == 12 ? lgemm_launch_one<1, LG_WIRE_DGRAD, 12, 1, 1>(a, n_sm, stream)
          : kdb == 6 ? lgemm_launch_one<1, LG_WIRE_DGRAD, 6, 1, 1>(a, n_sm, stream)
          : kdb == 4 ? lgemm_launch_one<1, LG_WIRE_DGRAD, 4, 1, 1>(a, n_sm, stream) : lgemm_launch_one<1, LG_WIRE_DGRAD, 8, 1, 1>(a, n_sm, stream);
        break;
      }
      e = kd == 8 ? lgemm_launch_one<1, LG_WIRE_DGRAD, 8, 1>(a, n_sm, stream)
        : kd == 6 ? lgemm_launch_one<1, LG_WIRE_DGRAD, 6, 1>(a, n_sm, stream) : lgemm_launch_one<1, LG_WIRE_DGRAD, 4, 1>(a, n_sm, stream);
      break;
    }
    case LG_MFN_FWD:    e = lgemm_launch_one<1, LG_MFN_FWD, 4, 0>(a, n_sm, stream); break;
    case LG_MFN_DGRAD:  e = lgemm_launch_one<1, LG_MFN_DGRAD, 4, 0>(a, n_sm, stream); break;
    case LG_GABOR_E:    e = lgemm_launch_one<1, LG_GABOR_E, 4, 0>(a, n_sm, stream); break;
    case LG_W2D_FWD:    e = lgemm_launch_one<3, LG_W2D_FWD, 2, 0>(a, n_sm, stream); break;
    case LG_W2D_DGRAD:  e = lgemm_launch_one<1, LG_W2D_DGRAD, 4, 0>(a, n_sm, stream); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace inr
