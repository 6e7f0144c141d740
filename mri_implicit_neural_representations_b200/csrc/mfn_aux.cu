// CUDA-core kernels around the MFN stage GEMMs (HBM-bound elementwise / small-N work):
//   mfn_encode_kernel  : gauss positional encoding (reference networks.py:30-33) or dense [bs,in] input -> fp16 X image
//   mfn_head_kernel    : output heads y_k = V_k z_{s_k} + c_k (reference mfn.py:38,92,262-263) (+ fused loss pieces for
//                        single-head models)
//   mfn_scalars_kernel : step scalars + per-stage gradient scales
//   mfn_top_kernel     : backward entry: fp16 dL/dy images per head and (dh, dp) of the top live stage
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "mfn.cuh"
#include "inr_loss.cuh"

namespace inr {

__device__ __forceinline__ void mfn_unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 mfn_pack8(const float (&f)[8]) {
  return make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------ input image
__global__ void __launch_bounds__(256) mfn_encode_kernel(const __grid_constant__ MfnAuxArgs a) {
  __shared__ float s_xn[2][kTileM];     // Gabor: |x|^2 partial sums of the two threads that share a row (fixed order)
  const MfnModel& M = a.m;
  const int tile = blockIdx.x, tid = threadIdx.x;
  float xn_part = 0.f;
  if (tile == 0 && tid == 0 && a.step_counter) *a.step_counter += 1;
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const uint32_t tile_bytes = kTileM * M.in_f * 2;
  uint8_t* ximg = a.ws + a.w.x + static_cast<size_t>(tile) * tile_bytes;
  const int n_kg = M.in_f / 8;
  for (int idx = tid; idx < kTileM * n_kg; idx += 256) {
    const int row = idx & (kTileM - 1), kg = idx >> 7;
    const int grow = tile * kTileM + row;
    const size_t srow = static_cast<size_t>(row_base) + grow;
    float v[8];
    if (M.input_kind == INPUT_GAUSS) {
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (grow < a.bs) { const float* c = a.coords + srow * 3; x0 = c[0]; x1 = c[1]; x2 = c[2]; }
      const bool is_cos = kg * 8 >= M.enc_size;
      const int f0 = kg * 8 - (is_cos ? M.enc_size : 0);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float* b = a.encB + (f0 + e) * 3;
        const float ang = 6.283185307179586f * fmaf(x0, __ldg(b), fmaf(x1, __ldg(b + 1), x2 * __ldg(b + 2)));
        v[e] = is_cos ? fast_cos(ang) : fast_sin(ang);
      }
    } else if (M.input_kind == INPUT_LOGF) {
      // reference networks.py:24-29: per coordinate c the block [sin(2 pi x_c B_0..n-1) | cos(2 pi x_c B_0..n-1)], B = a.encB [n];
      // features beyond 6 n are the zero padding of the operand image
      float xc[3] = {0.f, 0.f, 0.f};
      if (grow < a.bs) { const float* c = a.coords + srow * 3; xc[0] = c[0]; xc[1] = c[1]; xc[2] = c[2]; }
      const int n = M.enc_n;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = kg * 8 + e;
        float val = 0.f;
        if (f < 6 * n) {
          const int c = f / (2 * n), r = f - c * 2 * n;
          const bool is_cos = r >= n;
          // full-precision sine: B reaches 2^scale, so the angle spans thousands of radians where sin.approx loses its digits
          const float ang = (6.283185307179586f * xc[c]) * __ldg(a.encB + (is_cos ? r - n : r));
          val = is_cos ? cosf(ang) : sinf(ang);
        }
        v[e] = val;
      }
    } else {
      const float* xr = a.x + srow * M.in_f + kg * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = grow < a.bs ? xr[e] : 0.f;
    }
    st_global_v4(ximg + static_cast<size_t>(kg) * 2048 + row * 16, mfn_pack8(v));
#pragma unroll
    for (int e = 0; e < 8; ++e) xn_part = fmaf(v[e], v[e], xn_part);     // thread tid always owns row tid & 127
  }
  if (M.gabor) {
    // |x|^2 per row from the fp32 values (reference mfn.py:127) and the [1 | |x|^2] image the (s, u) wgrad units read
    s_xn[tid >> 7][tid & (kTileM - 1)] = xn_part;
    __syncthreads();
    if (tid < kTileM) {
      const float xn = s_xn[0][tid] + s_xn[1][tid];
      reinterpret_cast<float*>(a.ws + a.w.xn)[static_cast<size_t>(tile) * kTileM + tid] = xn;
      const bool valid = tile * kTileM + tid < a.bs;
      uint8_t* xa = a.ws + a.w.xa + static_cast<size_t>(tile) * kDzLastBytes;
      st_global_v4(xa + tid * 16, make_uint4(valid ? pack_h2(1.f, xn) : 0u, 0u, 0u, 0u));
      st_global_v4(xa + 2048 + tid * 16, make_uint4(0u, 0u, 0u, 0u));
    }
  }
}

// ------------------------------------------------------------------------------------------------ Gabor helpers
// |mu_j|^2 for every live stage (one warp per feature row), consumed by the envelope epilogue.
__global__ void __launch_bounds__(256) mfn_gabor_prep_kernel(const __grid_constant__ MfnAuxArgs a) {
  const MfnModel& M = a.m;
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= (M.top + 1) * M.width) return;
  const int i = warp / M.width, j = warp % M.width;
  const float* mu = a.params + M.mu_off[i] + static_cast<size_t>(j) * M.in_f;
  float acc = 0.f;
  for (int k = lane; k < M.in_f; k += 32) acc = fmaf(mu[k], mu[k], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) reinterpret_cast<float*>(a.ws + a.w.mn)[warp] = acc;
}

// Finalise d mu and d gamma from the split-K reductions (still carrying the stage scale S_i; the optimiser divides):
//   d mu_jk = gamma_j (Qx_jk - s_j mu_jk),   d gamma_j = -1/2 (u_j + |mu_j|^2 s_j - 2 sum_k mu_jk Qx_jk)
// (derivation: f_j = sin(p_j) exp(-gamma_j D_j / 2), D_j = |x|^2 + |mu_j|^2 - 2 x.mu_j, q_j = dL/df_j f_j;
//  dL/dmu_j = sum_rows q_j gamma_j (x - mu_j), dL/dgamma_j = -1/2 sum_rows q_j D_j; reference mfn.py:117-131.)
// One warp per (stage, feature); all split partials are added in split order, so the result is run-to-run identical.
__global__ void __launch_bounds__(256) mfn_gabor_grad_kernel(const __grid_constant__ MfnAuxArgs a, int n_split, int gstride) {
  const MfnModel& M = a.m;
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= (M.top + 1) * M.width) return;
  const int i = warp / M.width, j = warp % M.width;
  const float* gp = reinterpret_cast<const float*>(a.ws + a.w.gpart);
  float* gfin = reinterpret_cast<float*>(a.ws + a.w.gfin);
  float sj = 0.f, uj = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float* aux = gp + static_cast<size_t>(s) * gstride + M.aux_off[i] + j * 16;
    sj += aux[0]; uj += aux[1];
  }
  const float gamma = a.params[M.gamma_off[i] + j];
  const float* mu = a.params + M.mu_off[i] + static_cast<size_t>(j) * M.in_f;
  float dot = 0.f, mn = 0.f;
  for (int k = lane; k < M.in_f; k += 32) {
    float qx = 0.f;
    for (int s = 0; s < n_split; ++s) qx += gp[static_cast<size_t>(s) * gstride + M.mu_off[i] + static_cast<size_t>(j) * M.in_f + k];
    const float m = mu[k];
    gfin[M.gfin_mu[i] + static_cast<size_t>(j) * M.in_f + k] = gamma * (qx - sj * m);
    dot = fmaf(m, qx, dot); mn = fmaf(m, m, mn);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, off); mn += __shfl_xor_sync(0xffffffffu, mn, off); }
  if (lane == 0) gfin[M.gfin_gamma[i] + j] = -0.5f * (uj + mn * sj - 2.f * dot);
}

// ------------------------------------------------------------------------------------------------ heads (+ fused loss)
__global__ void __launch_bounds__(128) mfn_head_kernel(const __grid_constant__ MfnAuxArgs a) {
  __shared__ float sW[kMaxOut][512];
  __shared__ float red[4][8];
  const MfnModel& M = a.m;
  const int tile = blockIdx.x, row = threadIdx.x, lane = row & 31, q = row >> 5;
  // live head handled by this block
  int k = -1, seen = 0;
  for (int kk = 0; kk < M.n_heads; ++kk)
    if (M.head_live[kk]) { if (seen == static_cast<int>(blockIdx.y)) { k = kk; break; } ++seen; }
  const int stage = M.head_stage[k];
  for (int i = row; i < kMaxOut * M.width; i += 128) {
    const int o = i / M.width, j = i % M.width;
    sW[o][j] = o < M.out_f ? a.params[M.head_w[k] + o * M.width + j] : 0.f;
  }
  __syncthreads();
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const int grow = tile * kTileM + row;
  const bool valid = grow < a.bs;
  const size_t srow = static_cast<size_t>(row_base) + grow;
  const uint32_t tile_bytes = kTileM * M.width * 2;
  const uint8_t* zimg = a.ws + a.w.z[stage] + static_cast<size_t>(tile) * tile_bytes + row * 16;
  float acc[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
  for (int kg = 0; kg < M.width / 8; ++kg) {
    float z[8];
    mfn_unpack8(ld_global_nc_v4(zimg + static_cast<size_t>(kg) * 2048), z);
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < M.out_f) acc[o] = fmaf(z[e], sW[o][kg * 8 + e], acc[o]);
  }
  float y[kMaxOut] = {0.f, 0.f, 0.f, 0.f}, t[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
  float dact[kMaxOut] = {1.f, 1.f, 1.f, 1.f};      // d y / d(V z + c) of the wide chain's output activation
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o)
    if (o < M.out_f) {
      y[o] = acc[o] + a.params[M.head_b[k] + o];
      if (M.chain) {      // reference networks.py:94-96 (tanh), :63 (sigmoid), :107-117 (sine output layer)
        if (M.last_act == LAST_TANH) { y[o] = tanhf(y[o]); dact[o] = 1.f - y[o] * y[o]; }
        else if (M.last_act == LAST_SIGMOID) { y[o] = 1.f / (1.f + expf(-y[o])); dact[o] = y[o] * (1.f - y[o]); }
        else if (M.last_act == LAST_SIN) { const float pz = M.w0 * y[o]; y[o] = sinf(pz); dact[o] = M.w0 * cosf(pz); }
      }
    }
  const int out_ld = M.n_out * M.out_f, col = static_cast<int>(blockIdx.y) * M.out_f;
  if (valid && a.out)
    for (int o = 0; o < M.out_f; ++o) a.out[static_cast<size_t>(grow) * out_ld + col + o] = y[o];
  if (a.train && M.chain && a.loss.kind == LOSS_NONE) {
    // unfused path of a wide chain: the backward entry multiplies the caller's dL/dy by the output activation's derivative
    float* ddst = reinterpret_cast<float*>(a.ws + a.w.gl) + (static_cast<size_t>(tile) * kTileM + row) * 4;
    *reinterpret_cast<float4*>(ddst) = make_float4(dact[0], dact[1], dact[2], dact[3]);
  }
  if (!a.train || a.loss.kind == LOSS_NONE || M.n_out != 1) return;
  // ---- fused loss pieces (single-head models)
  float lA = 0.f, lB = 0.f, fs = 0.f, cnt = 0.f, amA = 0.f, amB = 0.f;
  float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid && a.gt) {
    const bool in_loss = a.mask ? (a.mask[srow] != 0) : true;
    if (a.loss.kind == LOSS_HDR && a.coords) {
      const float kx = a.coords[srow * 3 + 1], ky = a.coords[srow * 3 + 2];
      const float f = expf(-(kx * kx + ky * ky) / (2.f * a.loss.sigma * a.loss.sigma));
      fs = (1.f - f) * (1.f - f);
    }
    if (in_loss) {
      for (int o = 0; o < M.out_f; ++o) t[o] = a.gt[srow * M.out_f + o];
      RowLoss r = loss_row(a.loss, M.out_f, y, t);
      lA = r.lossA; lB = r.lossB; cnt = 1.f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) {       // the stored pieces are gradients with respect to the head's linear output
        r.gA[o] *= dact[o]; r.gB[o] *= dact[o];
        amA = fmaxf(amA, fabsf(r.gA[o])); amB = fmaxf(amB, fabsf(r.gB[o]));
      }
      gq = make_float4(r.gA[0], r.gA[1], r.gB[0], r.gB[1]);
    }
  }
  float* gdst = reinterpret_cast<float*>(a.ws + a.w.gl) + (static_cast<size_t>(tile) * kTileM + row) * 4;
  *reinterpret_cast<float4*>(gdst) = gq;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lA += __shfl_xor_sync(0xffffffffu, lA, off); lB += __shfl_xor_sync(0xffffffffu, lB, off);
    fs += __shfl_xor_sync(0xffffffffu, fs, off); cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    amA = fmaxf(amA, __shfl_xor_sync(0xffffffffu, amA, off)); amB = fmaxf(amB, __shfl_xor_sync(0xffffffffu, amB, off));
  }
  if (lane == 0) { red[q][0] = lA; red[q][1] = lB; red[q][2] = fs; red[q][3] = cnt; red[q][4] = amA; red[q][5] = amB; }
  __syncthreads();
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
}

// ------------------------------------------------------------------------------------------------ fused multi-head loss
// Reference src/train_kspace_multiscale.py:173-190: loss = sum_k w_kind * loss_kind(out_k[mask], gt[mask]) + cons_weight *
// sum_{k >= 1} mse(out_{k-1}[S_{k-1}].detach(), out_k[S_{k-1}]), S_i = rows with dist outside [cons_lo[i], cons_hi[i]]
// (ConsistencyLoss over ALL rows of the batch, src/metrics/losses.py:317-323).  One block per row tile evaluates every live
// head for its 128 rows, so a thread holds all head outputs of its row: per head it leaves the unnormalised gradient pieces
// (loss part, consistency part) and the tile partials; the normalisers (masked row count, |S_i|) are applied after the
// fixed-order reduction in mfn_ms_scalars_kernel.  out_f == 2, at most 4 live heads.
__global__ void __launch_bounds__(128) mfn_ms_head_kernel(const __grid_constant__ MfnAuxArgs a) {
  __shared__ float sW[2][512];
  __shared__ float red[4][4][8];      // [head][warp][slot]
  const MfnModel& M = a.m;
  const int tile = blockIdx.x, row = threadIdx.x, lane = row & 31, q = row >> 5;
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const int grow = tile * kTileM + row;
  const bool valid = grow < a.bs;
  const size_t srow = static_cast<size_t>(row_base) + grow;
  const uint32_t tile_bytes = kTileM * M.width * 2;
  float y[4][2];
  int hk = 0;
  for (int k = 0; k < M.n_heads; ++k) {
    if (!M.head_live[k]) continue;
    __syncthreads();
    for (int i = row; i < 2 * M.width; i += 128) sW[i / M.width][i % M.width] = a.params[M.head_w[k] + i];
    __syncthreads();
    const uint8_t* zimg = a.ws + a.w.z[M.head_stage[k]] + static_cast<size_t>(tile) * tile_bytes + row * 16;
    float a0 = 0.f, a1 = 0.f;
    for (int kg = 0; kg < M.width / 8; ++kg) {
      float z[8];
      mfn_unpack8(ld_global_nc_v4(zimg + static_cast<size_t>(kg) * 2048), z);
#pragma unroll
      for (int e = 0; e < 8; ++e) { a0 = fmaf(z[e], sW[0][kg * 8 + e], a0); a1 = fmaf(z[e], sW[1][kg * 8 + e], a1); }
    }
    y[hk][0] = a0 + a.params[M.head_b[k]]; y[hk][1] = a1 + a.params[M.head_b[k] + 1];
    if (valid && a.out) {
      float* o = a.out + static_cast<size_t>(grow) * (M.n_out * 2) + hk * 2;
      o[0] = y[hk][0]; o[1] = y[hk][1];
    }
    ++hk;
  }
  float t[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
  bool in_loss = false;
  float dist = 0.f;
  if (valid) {
    in_loss = a.mask ? (a.mask[srow] != 0) : true;
    t[0] = a.gt[srow * 2]; t[1] = a.gt[srow * 2 + 1];
    dist = a.dist ? a.dist[srow] : 0.f;
  }
  for (int h = 0; h < M.n_out; ++h) {
    float lA = 0.f, lC = 0.f, cnt = 0.f, cntS = 0.f, amA = 0.f, amB = 0.f;
    float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid && in_loss) {
      float yy[kMaxOut] = {y[h][0], y[h][1], 0.f, 0.f};
      RowLoss r = loss_row(a.loss, 2, yy, t);
      lA = r.lossA; cnt = 1.f;
      gq.x = r.gA[0]; gq.y = r.gA[1];
      amA = fmaxf(fabsf(r.gA[0]), fabsf(r.gA[1]));
    }
    if (valid && h >= 1 && a.loss.cons_weight != 0.f && (dist < a.loss.cons_lo[h - 1] || dist > a.loss.cons_hi[h - 1])) {
      const float d0 = y[h][0] - y[h - 1][0], d1 = y[h][1] - y[h - 1][1];
      gq.z = d0; gq.w = d1; lC = d0 * d0 + d1 * d1; cntS = 1.f;
      amB = fmaxf(fabsf(d0), fabsf(d1));
    }
    reinterpret_cast<float4*>(a.ws + a.w.msg)[(static_cast<size_t>(h) * a.w.n_tiles + tile) * kTileM + row] = gq;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lA += __shfl_xor_sync(0xffffffffu, lA, off); lC += __shfl_xor_sync(0xffffffffu, lC, off);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, off); cntS += __shfl_xor_sync(0xffffffffu, cntS, off);
      amA = fmaxf(amA, __shfl_xor_sync(0xffffffffu, amA, off)); amB = fmaxf(amB, __shfl_xor_sync(0xffffffffu, amB, off));
    }
    if (lane == 0) { red[h][q][0] = lA; red[h][q][1] = lC; red[h][q][2] = cnt; red[h][q][3] = cntS; red[h][q][4] = amA; red[h][q][5] = amB; }
  }
  __syncthreads();
  if (row < M.n_out) {
    const int h = row;
    float* pdst = reinterpret_cast<float*>(a.ws + a.w.msp) + (static_cast<size_t>(tile) * M.n_out + h) * kPartialsPerTile;
    for (int s = 0; s < 4; ++s) pdst[s] = (red[h][0][s] + red[h][1][s]) + (red[h][2][s] + red[h][3][s]);
    pdst[4] = fmaxf(fmaxf(red[h][0][4], red[h][1][4]), fmaxf(red[h][2][4], red[h][3][4]));
    pdst[5] = fmaxf(fmaxf(red[h][0][5], red[h][1][5]), fmaxf(red[h][2][5], red[h][3][5]));
    pdst[6] = 0.f; pdst[7] = 0.f;
  }
}

// Fixed-order reduction of the (tile, head) partials -> per-head normalisers and the composite loss.
__global__ void __launch_bounds__(256) mfn_ms_scalars_kernel(const __grid_constant__ MfnAuxArgs a) {
  __shared__ float part[8][4];
  const MfnModel& M = a.m;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* g = reinterpret_cast<float*>(a.ws + a.w.scal);
  const float* src = reinterpret_cast<const float*>(a.ws + a.w.msp);
  float total = 0.f;
  for (int h = 0; h < M.n_out; ++h) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int t = tid; t < a.w.n_tiles; t += 256) {
      const float* qd = src + (static_cast<size_t>(t) * M.n_out + h) * kPartialsPerTile;
      s0 += qd[0]; s1 += qd[1]; s2 += qd[2]; s3 += qd[3];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off);
      s2 += __shfl_xor_sync(0xffffffffu, s2, off); s3 += __shfl_xor_sync(0xffffffffu, s3, off);
    }
    __syncthreads();
    if (lane == 0) { part[warp][0] = s0; part[warp][1] = s1; part[warp][2] = s2; part[warp][3] = s3; }
    __syncthreads();
    if (tid == 0) {
      float lA = 0.f, lC = 0.f, cnt = 0.f, cntS = 0.f;
      for (int w = 0; w < 8; ++w) { lA += part[w][0]; lC += part[w][1]; cnt += part[w][2]; cntS += part[w][3]; }
      const float m = fmaxf(cnt, 1.f);
      float cA = 0.f, lv = 0.f;
      switch (a.loss.kind) {       // weights of the training loop, as in reduce_step_scalars
        case LOSS_L2:   cA = 1.f / (m * 2.f);  lv = lA * 0.5f / (m * 2.f); break;
        case LOSS_L1:   cA = 0.5f / (m * 2.f); lv = lA * 0.5f / (m * 2.f); break;
        case LOSS_MSLE: cA = 1.f / (m * 2.f);  lv = lA * 0.5f / (m * 2.f); break;
        case LOSS_LSL:  cA = 1.f / m;          lv = lA * 0.5f / m; break;
        default: break;
      }
      float cB = 0.f;
      if (cntS > 0.f) { cB = a.loss.cons_weight / cntS; lv += a.loss.cons_weight * lC / (2.f * cntS); }
      g[SC_HEAD_NORM + 2 * h] = cA; g[SC_HEAD_NORM + 2 * h + 1] = cB;
      total += lv;
    }
  }
  if (tid == 0) g[SC_MS_LOSS] = total;
}

// dL/dy [rows, n_out * 2] fp32 (normalised, unscaled) for the backward entry, which then proceeds as for an external dL/dout.
__global__ void __launch_bounds__(128) mfn_ms_dout_kernel(const __grid_constant__ MfnAuxArgs a) {
  const MfnModel& M = a.m;
  const int tile = blockIdx.x, row = threadIdx.x;
  const int grow = tile * kTileM + row;
  const float* g = reinterpret_cast<const float*>(a.ws + a.w.scal);
  float* dy = reinterpret_cast<float*>(a.ws + a.w.dyf) + static_cast<size_t>(grow) * (M.n_out * 2);
  for (int h = 0; h < M.n_out; ++h) {
    const float4 gq = reinterpret_cast<const float4*>(a.ws + a.w.msg)[(static_cast<size_t>(h) * a.w.n_tiles + tile) * kTileM + row];
    const float cA = g[SC_HEAD_NORM + 2 * h], cB = g[SC_HEAD_NORM + 2 * h + 1];
    dy[2 * h] = cA * gq.x + cB * gq.z; dy[2 * h + 1] = cA * gq.y + cB * gq.w;
  }
}

// amax partials of an external dL/dout [bs, n_out*out_f]
__global__ void __launch_bounds__(128) mfn_dout_amax_kernel(const float* dout, int bs, int ld, float* partials) {
  __shared__ float red[4];
  const int tile = blockIdx.x, row = tile * kTileM + threadIdx.x;
  float am = 0.f;
  if (row < bs)
    for (int o = 0; o < ld; ++o) am = fmaxf(am, fabsf(dout[static_cast<size_t>(row) * ld + o]));
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

// ------------------------------------------------------------------------------------------------ step scalars
__global__ void __launch_bounds__(256) mfn_scalars_kernel(const __grid_constant__ MfnAuxArgs a) {
  __shared__ float sc[kScalars];
  float* g = reinterpret_cast<float*>(a.ws + a.w.scal);
  reduce_step_scalars(reinterpret_cast<const float*>(a.ws + a.w.part), a.w.n_tiles, a.loss, a.m.out_f, a.bs_k, a.hyper, a.step, g, sc);
  if (threadIdx.x == 0 && a.train == 2) g[SC_LOSS] = g[SC_MS_LOSS];      // fused multi-head step: the loss was reduced by mfn_ms_scalars_kernel
  if (threadIdx.x == 0) {
    // per-stage scales from the previous step's amax (lagged dynamic scaling, 2^10 headroom below the fp16 maximum);
    // uncalibrated stages start at the loss scale divided by 16
    unsigned int* am = reinterpret_cast<unsigned int*>(g) + SC_LAYER_AMAX;
    for (int l = a.m.top; l >= 0; --l) {
      const float old = g[SC_LAYER_SCALE + l];
      const float seen = __uint_as_float(am[l]);
      float S;
      if (old > 0.f && seen > 0.f && isfinite(seen) && isfinite(old)) {
        int e = static_cast<int>(floorf(log2f(64.f * old / seen)));
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        S = exp2f(static_cast<float>(e));
      } else {
        S = sc[SC_SCALE] * 0.0625f;
      }
      g[SC_LAYER_SCALE + l] = S;
      am[l] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward entry
// dL/dy_k images for the head wgrad units, and (dh, dp) of the top live stage from its head gradient.
__global__ void __launch_bounds__(256) mfn_top_kernel(const __grid_constant__ MfnAuxArgs a) {
  const MfnModel& M = a.m;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const float* sc = reinterpret_cast<const float*>(a.ws + a.w.scal);
  const float S = sc[SC_SCALE], cA = sc[SC_CA], cB = sc[SC_CB];
  const int top = M.top;
  const float St = sc[SC_LAYER_SCALE + top];
  const int out_ld = M.n_out * M.out_f;
  const uint32_t tile_bytes = kTileM * M.width * 2;
  const int hk = M.stage_head[top];                   // live head index of the top stage
  int k_top = -1, seen = 0;
  for (int kk = 0; kk < M.n_heads; ++kk)
    if (M.head_live[kk]) { if (seen == hk) { k_top = kk; break; } ++seen; }
  const float* Wt = a.params + M.head_w[k_top];
  const uint8_t* gimg = a.ws + a.w.g[top] + static_cast<size_t>(tile) * tile_bytes;
  const uint8_t* cimg = a.ws + a.w.cp[top] + static_cast<size_t>(tile) * tile_bytes;
  const uint8_t* himg = a.ws + a.w.h[top] + static_cast<size_t>(tile) * tile_bytes;
  uint8_t* dhimg = a.ws + a.w.dh[top] + static_cast<size_t>(tile) * tile_bytes;
  uint8_t* dpimg = a.ws + a.w.dp[top] + static_cast<size_t>(tile) * tile_bytes;
  float amax = 0.f;
  const int n_kg = M.width / 8;
  for (int idx = tid; idx < kTileM * n_kg; idx += 256) {
    const int row = idx & (kTileM - 1), kg = idx >> 7;
    const int grow = tile * kTileM + row;
    // ---- dL/dy of every live head for this row (unscaled, fp32)
    if (kg < M.n_out) {       // one thread per (row, head) writes that head's padded fp16 image
      float d[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
      if (grow < a.bs) {
        if (a.dout) {
          for (int o = 0; o < M.out_f; ++o) d[o] = S * a.dout[static_cast<size_t>(grow) * out_ld + kg * M.out_f + o];
          if (M.chain) {      // output activation of the wide chain: derivative left by mfn_head_kernel
            const float4 da = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.ws + a.w.gl) +
                                                                (static_cast<size_t>(tile) * kTileM + row) * 4);
            d[0] *= da.x; d[1] *= da.y; d[2] *= da.z; d[3] *= da.w;
          }
        } else {
          const float4 g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.ws + a.w.gl) +
                                                             (static_cast<size_t>(tile) * kTileM + row) * 4);
          d[0] = S * (cA * g.x + cB * g.z); d[1] = S * (cA * g.y + cB * g.w);
        }
      }
      uint8_t* zl = a.ws + a.w.dout[kg] + static_cast<size_t>(tile) * kDzLastBytes;
      st_global_v4(zl + row * 16, make_uint4(pack_h2(d[0], d[1]), pack_h2(d[2], d[3]), 0u, 0u));
      st_global_v4(zl + 2048 + row * 16, make_uint4(0u, 0u, 0u, 0u));
    }
    float dy[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
    if (grow < a.bs) {
      if (a.dout) {
        for (int o = 0; o < M.out_f; ++o) dy[o] = St * a.dout[static_cast<size_t>(grow) * out_ld + hk * M.out_f + o];
        if (M.chain) {
          const float4 da = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.ws + a.w.gl) +
                                                              (static_cast<size_t>(tile) * kTileM + row) * 4);
          dy[0] *= da.x; dy[1] *= da.y; dy[2] *= da.z; dy[3] *= da.w;
        }
      } else {
        const float4 g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.ws + a.w.gl) +
                                                           (static_cast<size_t>(tile) * kTileM + row) * 4);
        dy[0] = St * (cA * g.x + cB * g.z); dy[1] = St * (cA * g.y + cB * g.w);
      }
    }
    const size_t off = static_cast<size_t>(kg) * 2048 + row * 16;
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, c[8], h[8], dh[8], dp[8], q[8];
    const bool filt_only = M.chain || top == 0;      // z_top = act(p_top): no multiplicative linear term
    if (!M.chain) mfn_unpack8(ld_global_nc_v4(gimg + off), g);
    mfn_unpack8(ld_global_nc_v4(cimg + off), c);
    if (!filt_only) mfn_unpack8(ld_global_nc_v4(himg + off), h);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float dz = 0.f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < M.out_f) dz = fmaf(dy[o], __ldg(Wt + o * M.width + kg * 8 + e), dz);
      if (!filt_only) { dh[e] = dz * g[e]; dp[e] = dz * h[e] * c[e]; q[e] = dh[e] * h[e]; }
      else { dh[e] = 0.f; dp[e] = dz * c[e]; q[e] = dz * g[e]; }
      amax = fmaxf(amax, fmaxf(fabsf(dh[e]), fabsf(dp[e])));
      if (M.gabor) amax = fmaxf(amax, fabsf(q[e]));
    }
    if (M.gabor) st_global_v4(a.ws + a.w.q[top] + static_cast<size_t>(tile) * tile_bytes + off, mfn_pack8(q));
    if (!filt_only) {
      if (M.bounded) {
        st_global_v4(a.ws + a.w.dhu[top] + static_cast<size_t>(tile) * tile_bytes + off, mfn_pack8(dh));
        bool masked = false;
        if (grow < a.bs) { const float dd = a.dist[(a.row_offset ? *a.row_offset : 0) + grow]; masked = (dd < M.bound_lo[top]) || (dd > M.bound_hi[top]); }
        if (masked) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dh[e] = 0.f;
        }
      }
      st_global_v4(dhimg + off, mfn_pack8(dh));
    }
    st_global_v4(dpimg + off, mfn_pack8(dp));
    if (M.bounded && kg < 2) {        // ones image [128 x 16] for the bias units (identical for every tile)
      const uint32_t one2 = 0x3C003C00u;
      st_global_v4(a.ws + a.w.ones + static_cast<size_t>(kg) * 2048 + row * 16, make_uint4(one2, one2, one2, one2));
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if ((tid & 31) == 0 && amax > 0.f && isfinite(amax))
    atomicMax(reinterpret_cast<unsigned int*>(a.ws + a.w.scal) + SC_LAYER_AMAX + top, __float_as_uint(amax));
}

cudaError_t launch_mfn_encode(const MfnAuxArgs& a, cudaStream_t st) {
  mfn_encode_kernel<<<a.w.n_tiles, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mfn_gabor_prep(const MfnAuxArgs& a, cudaStream_t st) {
  const int warps = (a.m.top + 1) * a.m.width;
  mfn_gabor_prep_kernel<<<(warps + 7) / 8, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mfn_gabor_grad(const MfnAuxArgs& a, cudaStream_t st) {
  const int warps = (a.m.top + 1) * a.m.width;
  const int gstride = (a.m.g_floats + 3) & ~3;
  mfn_gabor_grad_kernel<<<(warps + 7) / 8, 256, 0, st>>>(a, a.w.n_split, gstride);
  return cudaGetLastError();
}
cudaError_t launch_mfn_head(const MfnAuxArgs& a, cudaStream_t st) {
  mfn_head_kernel<<<dim3(a.w.n_tiles, a.m.n_out), 128, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mfn_dout_amax(const MfnAuxArgs& a, cudaStream_t st) {
  mfn_dout_amax_kernel<<<a.w.n_tiles, 128, 0, st>>>(a.dout, a.bs, a.m.n_out * a.m.out_f, reinterpret_cast<float*>(a.ws + a.w.part));
  return cudaGetLastError();
}
cudaError_t launch_mfn_scalars(const MfnAuxArgs& a, cudaStream_t st) {
  mfn_scalars_kernel<<<1, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mfn_ms_loss(const MfnAuxArgs& a, cudaStream_t st) {
  mfn_ms_head_kernel<<<a.w.n_tiles, 128, 0, st>>>(a);
  mfn_ms_scalars_kernel<<<1, 256, 0, st>>>(a);
  mfn_ms_dout_kernel<<<a.w.n_tiles, 128, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_mfn_top(const MfnAuxArgs& a, cudaStream_t st) {
  mfn_top_kernel<<<a.w.n_tiles, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace inr
