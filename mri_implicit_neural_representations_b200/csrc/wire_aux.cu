// CUDA-core kernels around the WIRE layer GEMMs (HBM-bound elementwise / small-K work):
//   wire_first_kernel   : real first layer 3 -> C + Gabor wavelet (reference networks.py:185-204 with is_first), writes the
//                         hi/lo operand images of hidden layer 1, the (a) image for backward and the coordinate image
//                         used by the first layer's wgrad
//   wire_last_kernel    : final complex linear C -> out, real part (networks.py:247-258): adds up the partial sums the last
//                         hidden layer's GEMM epilogue left per row, + per-row loss pieces (+ step scalars in its last CTA)
//   wire_scalars_kernel : step scalars from the tile partials (one block)
//   wire_blast_kernel   : backward of the final layer + Gabor derivative of the last hidden layer
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "wire.cuh"
#include "inr_loss.cuh"

namespace inr {

constexpr int kWAuxSplit = 4;     // CTAs per row tile in the first-layer / backward-entry kernels

__device__ __forceinline__ void wire_split8(const float (&y)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void wire_unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// ------------------------------------------------------------------------------------------------ first layer
__global__ void __launch_bounds__(256) wire_first_kernel(const __grid_constant__ WireAuxArgs a) {
  __shared__ float sW[kWP * 3], sB[kWP];
  const WireModel& M = a.m;
  const int tile = blockIdx.x, tid = threadIdx.x;
  griddep_launch_dependents();       // the first layer GEMM may set up (barriers, TMEM) while this kernel runs
  for (int i = tid; i < kWP * 3; i += 256) sW[i] = (i / 3) < M.c ? a.params[M.w_off[0] + i] : 0.f;
  for (int i = tid; i < kWP; i += 256) sB[i] = i < M.c ? a.params[M.b_off[0] + i] : 0.f;
  if (tile == 0 && tid == 0 && blockIdx.y == 0 && a.step_counter) *a.step_counter += 1;
  if (blockIdx.y == 0 && tid < kWMaxDepth)      // hand-over counters of the chained forward GEMMs (lgemm.cu) start at zero
    reinterpret_cast<unsigned int*>(a.ws + a.w.flags_fwd)[tid * a.w.n_tiles + tile] = 0u;
  __syncthreads();
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const float w = M.omega_first, s2 = M.sigma * M.sigma;
  uint8_t* hhi = a.ws + a.w.hhi[1] + static_cast<size_t>(tile) * kWTileBytes;
  uint8_t* hlo = a.ws + a.w.hlo[1] + static_cast<size_t>(tile) * kWTileBytes;
  uint8_t* ab = a.ws + a.w.ab[0] + static_cast<size_t>(tile) * kWTileBytes;
  // four CTAs per row tile, each one quarter of the feature groups: these kernels are latency-bound at one CTA per tile
  const int kg_per = (kWP / 8) / kWAuxSplit, kg0 = blockIdx.y * kg_per;
  for (int idx = tid; idx < kTileM * kg_per; idx += 256) {
    const int row = idx & (kTileM - 1), kg = kg0 + (idx >> 7);
    const int grow = tile * kTileM + row;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (grow < a.bs) {
      const float* c = a.coords + (static_cast<size_t>(row_base) + grow) * 3;
      x0 = c[0]; x1 = c[1]; x2 = c[2];
    }
    float za[8], yr[8], yi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int f = kg * 8 + e;
      const float z = fmaf(x0, sW[3 * f], fmaf(x1, sW[3 * f + 1], fmaf(x2, sW[3 * f + 2], sB[f])));
      za[e] = z;
      const float mag = __expf(-s2 * z * z);
      const bool live = f < M.c;
      yr[e] = live ? mag * fast_cos(w * z) : 0.f;
      yi[e] = live ? mag * fast_sin(w * z) : 0.f;
    }
    uint4 rh, rl, ih, il;
    wire_split8(yr, rh, rl);
    wire_split8(yi, ih, il);
    const size_t off_r = static_cast<size_t>(kg) * 2048 + row * 16, off_i = static_cast<size_t>(kWP / 8 + kg) * 2048 + row * 16;
    st_global_v4(hhi + off_r, rh); st_global_v4(hhi + off_i, ih);
    st_global_v4(hlo + off_r, rl); st_global_v4(hlo + off_i, il);
    if (a.train) {
      st_global_v4(ab + off_r, make_uint4(pack_h2(za[0], za[1]), pack_h2(za[2], za[3]), pack_h2(za[4], za[5]), pack_h2(za[6], za[7])));
      if (kg == 0) {
        // coordinate image for the first layer's wgrad: [x_hi(3), 1, x_lo(3), 0 | 0 x 8]
        const float xs[3] = {x0, x1, x2};
        float v[8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float h = __half2float(__float2half_rn(xs[c]));
          v[c] = h; v[4 + c] = xs[c] - h;
        }
        v[3] = grow < a.bs ? 1.f : 0.f; v[7] = 0.f;
        uint8_t* xi = a.ws + a.w.ximg + static_cast<size_t>(tile) * kDzLastBytes;
        st_global_v4(xi + row * 16, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
        st_global_v4(xi + 2048 + row * 16, make_uint4(0u, 0u, 0u, 0u));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ step scalars
// One block: tile partials -> loss, normalisers, global gradient scale, Adam bias corrections, per-layer dZ scales.
__device__ inline void wire_step_scalars(const WireAuxArgs& a) {
  __shared__ float sc[kScalars];
  __shared__ float s_old[kWMaxDepth + 1], s_seen[kWMaxDepth + 1];
  float* g = reinterpret_cast<float*>(a.ws + a.w.scal);
  unsigned int* am = reinterpret_cast<unsigned int*>(g) + SC_LAYER_AMAX;
  // one thread per layer fetches last step's scale / amax while the block reduces the tile partials: the recurrence
  // below then runs out of shared memory instead of through 2 x (depth + 1) dependent global loads on one thread
  if (static_cast<int>(threadIdx.x) <= a.m.depth) {
    s_old[threadIdx.x] = g[SC_LAYER_SCALE + threadIdx.x];
    s_seen[threadIdx.x] = __uint_as_float(am[threadIdx.x]);
  }
  reduce_step_scalars(reinterpret_cast<const float*>(a.ws + a.w.part), a.w.n_tiles, a.loss, a.m.out_f, a.bs_k, a.hyper, a.step, g, sc);
  if (threadIdx.x == 0) {
    // per-layer dZ scales for this step from the previous step's amax (lagged dynamic scaling, 2^10 headroom below
    // the fp16 maximum); uncalibrated layers start 16x below the layer above (gradients grow ~10x per layer downwards)
    float above = sc[SC_SCALE];
    for (int l = a.m.depth; l >= 0; --l) {
      const float old = s_old[l];
      const float seen = s_seen[l];
      float S;
      if (old > 0.f && seen > 0.f && isfinite(seen) && isfinite(old)) {
        int e = static_cast<int>(floorf(log2f(64.f * old / seen)));
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        S = exp2f(static_cast<float>(e));
      } else {
        S = above * 0.0625f;
      }
      g[SC_LAYER_SCALE + l] = S;
      am[l] = 0u;
      above = S;
    }
  }
}

__global__ void __launch_bounds__(256) wire_scalars_kernel(const __grid_constant__ WireAuxArgs a) {
  // hand-over counters of the chained dgrad GEMMs start at zero (this kernel precedes every backward that wire_last's folded
  // scalars do not: autograd face, TV pass, calibration re-runs)
  unsigned int* fb = reinterpret_cast<unsigned int*>(a.ws + a.w.flags_bwd);
  for (int i = threadIdx.x; i < kWMaxDepth * a.w.n_tiles; i += blockDim.x) fb[i] = 0u;
  wire_step_scalars(a);
}

// ------------------------------------------------------------------------------------------------ final layer + loss
__global__ void __launch_bounds__(512) wire_last_kernel(const __grid_constant__ WireAuxArgs a) {
  __shared__ float red[4][8];
  __shared__ float s_part[3][kTileM][kMaxOut];
  const WireModel& M = a.m;
  // 4 threads per row, each reduces 6 of the 24 feature groups (independent 16-byte loads in flight), then one combines
  const int tile = blockIdx.x, row = threadIdx.x & (kTileM - 1), part = threadIdx.x >> 7, lane = row & 31, q = row >> 5;
  const int L = M.depth + 1;
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const int grow = tile * kTileM + row;
  const bool valid = grow < a.bs;
  const size_t srow = static_cast<size_t>(row_base) + grow;
  // What the row's loss needs besides the network output -- mask, target, k-space position, output bias -- is input of the
  // step, not a product of the layer chain: fetch it while the chain's last items are still running.
  bool in_loss_pref = false;
  float t_pref[kMaxOut] = {0.f, 0.f, 0.f, 0.f}, b_pref[kMaxOut] = {0.f, 0.f, 0.f, 0.f}, kx_pref = 0.f, ky_pref = 0.f;
  if (part == 0) {
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) if (o < M.out_f) b_pref[o] = a.params[M.b_off[L] + 2 * o];
    if (a.train && valid && a.gt && a.loss.kind != LOSS_NONE) {
      in_loss_pref = a.mask ? (a.mask[srow] != 0) : true;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) if (o < M.out_f) t_pref[o] = a.gt[srow * M.out_f + o];
      if (a.loss.kind == LOSS_HDR) { kx_pref = a.coords[srow * 3 + 1]; ky_pref = a.coords[srow * 3 + 2]; }
    }
  }
  griddep_wait();                    // the last layer GEMM's partial sums are complete
  // hand-over counters of the chained dgrad GEMMs start at zero (the chain may follow this kernel directly: the backward of
  // the final linear rides in it, lgemm.cu "top" items)
  if (a.train && threadIdx.x < kWMaxDepth) reinterpret_cast<unsigned int*>(a.ws + a.w.flags_bwd)[threadIdx.x * a.w.n_tiles + tile] = 0u;
  // ... and the forward chain's own counters go back to zero for the next launch, which may start a step (first layer folded
  // in) and then has no kernel before it that could do this
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kWMaxDepth)
    reinterpret_cast<unsigned int*>(a.ws + a.w.flags_fwd)[(threadIdx.x - 32) * a.w.n_tiles + tile] = 0u;
  // Dependents (the backward entry kernel) are released only now: whatever they read ahead of their own wait -- the saved
  // activations of the last hidden layer -- is final once this kernel is past the chain.
  griddep_launch_dependents();
  float acc[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
  {
    // the last hidden layer's GEMM epilogue left kWOutParts partial sums per row (lgemm.cu, LG_WIRE_FWD): 2 of them per
    // thread here, combined below in a fixed order
    const float4* op = reinterpret_cast<const float4*>(a.ws + a.w.outacc) + static_cast<size_t>(tile) * kWOutParts * kTileM + row;
    const float4 p0 = op[(2 * part) * kTileM], p1 = op[(2 * part + 1) * kTileM];
    acc[0] = p0.x + p1.x; acc[1] = p0.y + p1.y;
  }
  if (part > 0) {
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) s_part[part - 1][row][o] = acc[o];
  }
  __syncthreads();
  if (part == 0) {
  float y[kMaxOut] = {0.f, 0.f, 0.f, 0.f}, t[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o)
    if (o < M.out_f)      // fixed combination order; real part of the complex bias
      y[o] = ((acc[o] + s_part[0][row][o]) + (s_part[1][row][o] + s_part[2][row][o])) + b_pref[o];
  if (valid && a.out)
    for (int o = 0; o < M.out_f; ++o) a.out[static_cast<size_t>(grow) * M.out_f + o] = y[o];
  if (a.train) {
  float lA = 0.f, lB = 0.f, fs = 0.f, cnt = 0.f, amA = 0.f, amB = 0.f;
  float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid && a.gt && a.loss.kind != LOSS_NONE) {
    const bool in_loss = in_loss_pref;
    if (a.loss.kind == LOSS_HDR) {
      const float kx = kx_pref, ky = ky_pref;
      const float f = expf(-(kx * kx + ky * ky) / (2.f * a.loss.sigma * a.loss.sigma));
      fs = (1.f - f) * (1.f - f);
    }
    if (in_loss) {
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) t[o] = t_pref[o];
      RowLoss r = loss_row(a.loss, M.out_f, y, t);
      lA = r.lossA; lB = r.lossB; cnt = 1.f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) { amA = fmaxf(amA, fabsf(r.gA[o])); amB = fmaxf(amB, fabsf(r.gB[o])); }
      gq = make_float4(r.gA[0], r.gA[1], r.gB[0], r.gB[1]);
    }
  }
  float* gdst = reinterpret_cast<float*>(a.ws + a.w.g) + (static_cast<size_t>(tile) * kTileM + row) * 4;
  *reinterpret_cast<float4*>(gdst) = gq;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lA += __shfl_xor_sync(0xffffffffu, lA, off); lB += __shfl_xor_sync(0xffffffffu, lB, off);
    fs += __shfl_xor_sync(0xffffffffu, fs, off); cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    amA = fmaxf(amA, __shfl_xor_sync(0xffffffffu, amA, off)); amB = fmaxf(amB, __shfl_xor_sync(0xffffffffu, amB, off));
  }
  if (lane == 0) { red[q][0] = lA; red[q][1] = lB; red[q][2] = fs; red[q][3] = cnt; red[q][4] = amA; red[q][5] = amB; }
  named_bar_sync(1, 128);
  if (row == 0) {
    float* pdst = reinterpret_cast<float*>(a.ws + a.w.part) + static_cast<size_t>(tile) * kPartialsPerTile;
    pdst[0] = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
    pdst[1] = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
    pdst[2] = (red[0][2] + red[1][2]) + (red[2][2] + red[3][2]);
    pdst[3] = (red[0][3] + red[1][3]) + (red[2][3] + red[3][3]);
    pdst[4] = fmaxf(fmaxf(red[0][4], red[1][4]), fmaxf(red[2][4], red[3][4]));
    pdst[5] = fmaxf(fmaxf(red[0][5], red[1][5]), fmaxf(red[2][5], red[3][5]));
    pdst[6] = 0.f; pdst[7] = 0.f;
  }
  }  // train
  }  // part == 0
  if (!(a.train && a.fold_scalars)) return;
  // The last CTA to finish reduces all tile partials to the step scalars (what wire_scalars_kernel does as a launch
  // of its own): every CTA publishes its partials (fence) before it counts itself in.
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  unsigned int* done = reinterpret_cast<unsigned int*>(a.ws + a.w.scal) + SC_DONE_COUNT;
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  wire_step_scalars(a);
  if (threadIdx.x == 0) *done = 0u;
}

// amax partials for an externally supplied dL/dout (autograd path)
__global__ void __launch_bounds__(128) wire_dout_amax_kernel(const float* dout, int bs, int out_f, float* partials) {
  __shared__ float red[4];
  const int tile = blockIdx.x, row = tile * kTileM + threadIdx.x;
  float am = 0.f;
  if (row < bs)
    for (int o = 0; o < out_f; ++o) am = fmaxf(am, fabsf(dout[static_cast<size_t>(row) * out_f + o]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = am;
  __syncthreads();
  if (threadIdx.x == 0) {
    float* p = partials + static_cast<size_t>(tile) * kPartialsPerTile;
    for (int i = 0; i < 8; ++i) p[i] = 0.f;
    p[4] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  }
}

// ------------------------------------------------------------------------------------------------ backward of the final layer
__global__ void __launch_bounds__(256) wire_blast_kernel(const __grid_constant__ WireAuxArgs a) {
  __shared__ float sWr[kMaxOut][kWP], sWi[kMaxOut][kWP];
  const WireModel& M = a.m;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int L = M.depth + 1;
  griddep_launch_dependents();
  for (int i = tid; i < kMaxOut * kWP; i += 256) {
    const int o = i / kWP, j = i % kWP;
    const bool ok = o < M.out_f && j < M.c;
    sWr[o][j] = ok ? a.params[M.w_off[L] + (o * M.c + j) * 2] : 0.f;
    sWi[o][j] = ok ? a.params[M.w_off[L] + (o * M.c + j) * 2 + 1] : 0.f;
  }
  const float w = M.depth >= 1 ? M.omega_hidden : M.omega_first, s2 = M.sigma * M.sigma;
  const uint8_t* yimg = a.ws + a.w.hhi[L] + static_cast<size_t>(tile) * kWTileBytes;
  const uint8_t* abimg = a.ws + a.w.ab[M.depth] + static_cast<size_t>(tile) * kWTileBytes;
  uint8_t* dzimg = a.ws + a.w.dz[M.depth] + static_cast<size_t>(tile) * kWTileBytes;
  constexpr int kg_per = (kWP / 8) / kWAuxSplit, kIt = kTileM * kg_per / 256;     // 6 feature groups per CTA, 3 per thread
  const int kg0 = blockIdx.y * kg_per;
  const int row = tid & (kTileM - 1), grow = tile * kTileM + row;
  // The saved activations of the last hidden layer are final before this kernel is released (wire_last_kernel triggers its
  // dependents after its own wait on the layer chain): request all of them now, ahead of the loss scalars
  uint4 qy[kIt][2], qab[kIt][2];
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int kg = kg0 + ((tid + it * 256) >> 7);
    const size_t off_r = static_cast<size_t>(kg) * 2048 + row * 16, off_i = static_cast<size_t>(kWP / 8 + kg) * 2048 + row * 16;
    qy[it][0] = ld_global_nc_v4(yimg + off_r); qy[it][1] = ld_global_nc_v4(yimg + off_i);
    qab[it][0] = ld_global_nc_v4(abimg + off_r);
    qab[it][1] = M.depth >= 1 ? ld_global_nc_v4(abimg + off_i) : make_uint4(0u, 0u, 0u, 0u);
  }
  griddep_wait();                    // loss pieces and step scalars of wire_last_kernel are complete
  __syncthreads();
  if (blockIdx.y == 0 && tid < kWMaxDepth)      // hand-over counters of the chained dgrad GEMMs start at zero
    reinterpret_cast<unsigned int*>(a.ws + a.w.flags_bwd)[tid * a.w.n_tiles + tile] = 0u;
  const float* sc = reinterpret_cast<const float*>(a.ws + a.w.scal);
  const float S = sc[SC_SCALE], cA = sc[SC_CA], cB = sc[SC_CB];
  const float ratio = sc[SC_LAYER_SCALE + M.depth] / S;      // dz_last carries S, the stored dZ of the last hidden layer S[depth]
  float amax = 0.f;
  float dz[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
  if (grow < a.bs) {
    if (a.dout) {
      for (int o = 0; o < M.out_f; ++o) dz[o] = S * a.dout[static_cast<size_t>(grow) * M.out_f + o];
    } else {
      const float4 g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.ws + a.w.g) +
                                                         (static_cast<size_t>(tile) * kTileM + row) * 4);
      dz[0] = S * (cA * g.x + cB * g.z);
      dz[1] = S * (cA * g.y + cB * g.w);
    }
  }
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int kg = kg0 + ((tid + it * 256) >> 7);
    if (kg == 0) {
      uint8_t* zl = a.ws + a.w.dzlast + static_cast<size_t>(tile) * kDzLastBytes;
      st_global_v4(zl + row * 16, make_uint4(pack_h2(dz[0], dz[1]), pack_h2(dz[2], dz[3]), 0u, 0u));
      st_global_v4(zl + 2048 + row * 16, make_uint4(0u, 0u, 0u, 0u));
    }
    const size_t off_r = static_cast<size_t>(kg) * 2048 + row * 16, off_i = static_cast<size_t>(kWP / 8 + kg) * 2048 + row * 16;
    float yr[8], yi[8], za[8], zb[8], da[8], db[8];
    wire_unpack8(qy[it][0], yr);
    wire_unpack8(qy[it][1], yi);
    wire_unpack8(qab[it][0], za);
    wire_unpack8(qab[it][1], zb);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = kg * 8 + e;
      float gr = 0.f, gi = 0.f;     // dL/d Re(h_j), dL/d Im(h_j) for out = Re(h W^T + b)
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < M.out_f) { gr = fmaf(dz[o], sWr[o][j], gr); gi = fmaf(-dz[o], sWi[o][j], gi); }
      const float P = gr * yr[e] + gi * yi[e];
      const float Q = gr * yi[e] - gi * yr[e];
      da[e] = ratio * (-2.f * s2 * za[e] * P - w * Q);
      db[e] = M.depth >= 1 ? ratio * (-(w + 2.f * s2 * zb[e]) * P) : 0.f;
      amax = fmaxf(amax, fmaxf(fabsf(da[e]), fabsf(db[e])));
    }
    st_global_v4(dzimg + off_r, make_uint4(pack_h2(da[0], da[1]), pack_h2(da[2], da[3]), pack_h2(da[4], da[5]), pack_h2(da[6], da[7])));
    st_global_v4(dzimg + off_i, make_uint4(pack_h2(db[0], db[1]), pack_h2(db[2], db[3]), pack_h2(db[4], db[5]), pack_h2(db[6], db[7])));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if ((tid & 31) == 0 && amax > 0.f && isfinite(amax))
    atomicMax(reinterpret_cast<unsigned int*>(a.ws + a.w.scal) + SC_LAYER_AMAX + M.depth, __float_as_uint(amax));
}

cudaError_t launch_wire_first(const WireAuxArgs& a, cudaStream_t st) {
  wire_first_kernel<<<dim3(a.w.n_tiles, kWAuxSplit), 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_wire_last(const WireAuxArgs& a, cudaStream_t st) {
  return launch_dependent(wire_last_kernel, dim3(a.w.n_tiles), dim3(512), 0, st, a);
}
cudaError_t launch_wire_scalars(const WireAuxArgs& a, cudaStream_t st) {
  wire_scalars_kernel<<<1, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_wire_dout_amax(const WireAuxArgs& a, cudaStream_t st) {
  wire_dout_amax_kernel<<<a.w.n_tiles, 128, 0, st>>>(a.dout, a.bs, a.m.out_f, reinterpret_cast<float*>(a.ws + a.w.part));
  return cudaGetLastError();
}
cudaError_t launch_wire_blast(const WireAuxArgs& a, cudaStream_t st) {
  return launch_dependent(wire_blast_kernel, dim3(a.w.n_tiles, kWAuxSplit), dim3(256), 0, st, a);
}

}  // namespace inr
