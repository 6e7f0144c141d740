// CUDA-core kernels around the WIRE2D layer GEMMs (reference src/models/wire2d.py:4-118).  Same structure as wire_aux.cu;
// a layer evaluates  y = exp(j w lin) * exp(-s^2 (|lin|^2 + |orth|^2)),  lin = a + jb = linear(h),  orth = c + jd =
// scale_orth(h)  (:50-60).  With P = Re(conj(g) y), Q = Im(conj(g) y) for the upstream gradient g = dL/dy:
//   dL/da = -2 s^2 a P - w Q,   dL/db = -(w + 2 s^2 b) P,   dL/dc = -2 s^2 c P,   dL/dd = -2 s^2 d P.
//   w2d_first_kernel : real first layer (both linears real, :27-30) -> H images of layer 1, Z image [a|0|c|0], coordinate image
//   w2d_last_kernel  : final complex linear, real part (:98-117) + per-row loss pieces; with `last_tanh` (:106-107) the output is
//                      Re(tanh(z)), z = x + jy the complex linear:  out = sinh 2x / (cosh 2x + cos 2y), and the row's
//                      d out / dx, d out / dy are left in the workspace for the backward entry kernel
//   w2d_blast_kernel : backward of the final layer + derivative of the last Gabor layer
#include <cuda_runtime.h>
#include "inr_ptx.cuh"
#include "wire.cuh"
#include "inr_loss.cuh"

namespace inr {

constexpr int kW2dAuxSplit = 4;     // CTAs per row tile in the first-layer / backward-entry kernels

__device__ __forceinline__ void w2d_split8(const float (&y)[8], uint4& hi, uint4& lo) {
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
__device__ __forceinline__ void w2d_unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 w2d_pack8(const float (&f)[8]) {
  return make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------ first layer
__global__ void __launch_bounds__(256) w2d_first_kernel(const __grid_constant__ WireAuxArgs a) {
  __shared__ float sW[kW2dMaxP * 3], sB[kW2dMaxP], sV[kW2dMaxP * 3], sVB[kW2dMaxP];
  const WireModel& M = a.m;
  const int P = M.P, tile = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < P * 3; i += 256) {
    sW[i] = (i / 3) < M.c ? a.params[M.w_off[0] + i] : 0.f;
    sV[i] = (i / 3) < M.c ? a.params[M.v_off[0] + i] : 0.f;
  }
  for (int i = tid; i < P; i += 256) {
    sB[i] = i < M.c ? a.params[M.b_off[0] + i] : 0.f;
    sVB[i] = i < M.c ? a.params[M.vb_off[0] + i] : 0.f;
  }
  if (tile == 0 && tid == 0 && blockIdx.y == 0 && a.step_counter) *a.step_counter += 1;
  if (blockIdx.y == 0 && tid < kWMaxDepth)      // hand-over counters of the chained forward GEMMs (lgemm.cu) start at zero
    reinterpret_cast<unsigned int*>(a.ws + a.w.flags_fwd)[tid * a.w.n_tiles + tile] = 0u;
  __syncthreads();
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const float w = M.omega_first, s2 = M.sigma * M.sigma;
  const size_t th = static_cast<size_t>(kTileM) * 2 * P * 2, tz = 2 * th;
  uint8_t* hhi = a.ws + a.w.hhi[1] + tile * th;
  uint8_t* hlo = a.ws + a.w.hlo[1] + tile * th;
  uint8_t* zimg = a.ws + a.w.ab[0] + tile * tz;
  const int pg = P / 8;
  const int kg_per = pg / kW2dAuxSplit, kg0 = blockIdx.y * kg_per;      // four CTAs per row tile (latency-bound otherwise)
  for (int idx = tid; idx < kTileM * kg_per; idx += 256) {
    const int row = idx & (kTileM - 1), kg = kg0 + (idx >> 7);
    const int grow = tile * kTileM + row;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (grow < a.bs) {
      const float* c = a.coords + (static_cast<size_t>(row_base) + grow) * 3;
      x0 = c[0]; x1 = c[1]; x2 = c[2];
    }
    float za[8], zc[8], yr[8], yi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int f = kg * 8 + e;
      const float za_ = fmaf(x0, sW[3 * f], fmaf(x1, sW[3 * f + 1], fmaf(x2, sW[3 * f + 2], sB[f])));
      const float zc_ = fmaf(x0, sV[3 * f], fmaf(x1, sV[3 * f + 1], fmaf(x2, sV[3 * f + 2], sVB[f])));
      za[e] = za_; zc[e] = zc_;
      const float mag = __expf(-s2 * (za_ * za_ + zc_ * zc_));
      const bool live = f < M.c;
      yr[e] = live ? mag * fast_cos(w * za_) : 0.f;
      yi[e] = live ? mag * fast_sin(w * za_) : 0.f;
    }
    uint4 rh, rl, ih, il;
    w2d_split8(yr, rh, rl);
    w2d_split8(yi, ih, il);
    const size_t ro = static_cast<size_t>(row) * 16;
    st_global_v4(hhi + static_cast<size_t>(kg) * 2048 + ro, rh); st_global_v4(hhi + static_cast<size_t>(pg + kg) * 2048 + ro, ih);
    st_global_v4(hlo + static_cast<size_t>(kg) * 2048 + ro, rl); st_global_v4(hlo + static_cast<size_t>(pg + kg) * 2048 + ro, il);
    if (a.train) {
      const uint4 z0 = make_uint4(0u, 0u, 0u, 0u);
      st_global_v4(zimg + static_cast<size_t>(kg) * 2048 + ro, w2d_pack8(za));
      st_global_v4(zimg + static_cast<size_t>(pg + kg) * 2048 + ro, z0);
      st_global_v4(zimg + static_cast<size_t>(2 * pg + kg) * 2048 + ro, w2d_pack8(zc));
      st_global_v4(zimg + static_cast<size_t>(3 * pg + kg) * 2048 + ro, z0);
      if (kg == 0) {      // coordinate image for the first layer's wgrad: [x_hi(3), 1, x_lo(3), 0 | 0 x 8]
        const float xs[3] = {x0, x1, x2};
        float v[8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float h = __half2float(__float2half_rn(xs[c]));
          v[c] = h; v[4 + c] = xs[c] - h;
        }
        v[3] = grow < a.bs ? 1.f : 0.f; v[7] = 0.f;
        uint8_t* xi = a.ws + a.w.ximg + static_cast<size_t>(tile) * kDzLastBytes;
        st_global_v4(xi + ro, w2d_pack8(v));
        st_global_v4(xi + 2048 + ro, z0);
      }
    }
  }
}

// Real part of the complex tanh and its two partial derivatives, written with e = exp(-2|x|) so that nothing overflows:
//   Re tanh(x + jy) = sgn(x) (1 - e^2) / D,  D = 1 + e^2 + 2 e cos 2y
//   d/dx = (8 e^2 + 4 e (1 + e^2) cos 2y) / D^2,   d/dy = sgn(x) 4 e (1 - e^2) sin 2y / D^2
__device__ __forceinline__ void w2d_re_tanh(float x, float y, float& out, float& tx, float& ty) {
  const float e = expf(-2.f * fabsf(x)), e2 = e * e;
  float s, c;
  sincosf(2.f * y, &s, &c);
  const float D = 1.f + e2 + 2.f * e * c, sg = x < 0.f ? -1.f : 1.f;
  out = sg * (1.f - e2) / D;
  tx = (8.f * e2 + 4.f * e * (1.f + e2) * c) / (D * D);
  ty = sg * 4.f * e * (1.f - e2) * s / (D * D);
}

// ------------------------------------------------------------------------------------------------ final layer + loss
__global__ void __launch_bounds__(512) w2d_last_kernel(const __grid_constant__ WireAuxArgs a) {
  __shared__ float sWr[kMaxOut][kW2dMaxP], sWi[kMaxOut][kW2dMaxP];
  __shared__ float red[4][8];
  __shared__ float s_part[3][kTileM][kMaxOut];     // out_f <= 2: columns 0..1 real parts, 2..3 imaginary parts (tanh tail)
  const WireModel& M = a.m;
  const int P = M.P, pg = P / 8;
  const int tile = blockIdx.x, row = threadIdx.x & (kTileM - 1), part = threadIdx.x >> 7, lane = row & 31, q = row >> 5;
  const int L = M.depth + 1;
  const bool ctanh = M.last_tanh != 0;
  for (int i = threadIdx.x; i < kMaxOut * P; i += 512) {
    const int o = i / P, j = i % P;
    const bool ok = o < M.out_f && j < M.c;
    sWr[o][j] = ok ? a.params[M.w_off[L] + (o * M.c + j) * 2] : 0.f;
    sWi[o][j] = ok ? a.params[M.w_off[L] + (o * M.c + j) * 2 + 1] : 0.f;
  }
  __syncthreads();
  const int row_base = a.row_offset ? *a.row_offset : 0;
  const int grow = tile * kTileM + row;
  const bool valid = grow < a.bs;
  const size_t srow = static_cast<size_t>(row_base) + grow;
  const size_t th = static_cast<size_t>(kTileM) * 2 * P * 2;
  const uint8_t* hhi = a.ws + a.w.hhi[L] + tile * th + row * 16;
  const uint8_t* hlo = a.ws + a.w.hlo[L] + tile * th + row * 16;
  float acc[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
  for (int kg = part * (pg / 4); kg < (part + 1) * (pg / 4); ++kg) {
    float rh[8], rl[8], ih[8], il[8];
    w2d_unpack8(ld_global_nc_v4(hhi + static_cast<size_t>(kg) * 2048), rh);
    w2d_unpack8(ld_global_nc_v4(hlo + static_cast<size_t>(kg) * 2048), rl);
    w2d_unpack8(ld_global_nc_v4(hhi + static_cast<size_t>(pg + kg) * 2048), ih);
    w2d_unpack8(ld_global_nc_v4(hlo + static_cast<size_t>(pg + kg) * 2048), il);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float hr = rh[e] + rl[e], hi = ih[e] + il[e];
#pragma unroll
      for (int o = 0; o < 2; ++o)
        if (o < M.out_f) {
          acc[o] = fmaf(hr, sWr[o][kg * 8 + e], fmaf(-hi, sWi[o][kg * 8 + e], acc[o]));
          if (ctanh) acc[2 + o] = fmaf(hr, sWi[o][kg * 8 + e], fmaf(hi, sWr[o][kg * 8 + e], acc[2 + o]));
        }
    }
  }
  if (part > 0) {
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o) s_part[part - 1][row][o] = acc[o];
  }
  __syncthreads();
  if (part > 0) return;
  float y[kMaxOut] = {0.f, 0.f, 0.f, 0.f}, t[kMaxOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 2; ++o)
    if (o < M.out_f)
      y[o] = ((acc[o] + s_part[0][row][o]) + (s_part[1][row][o] + s_part[2][row][o])) + a.params[M.b_off[L] + 2 * o];
  if (ctanh) {
    float fac[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 2; ++o)
      if (o < M.out_f) {
        const float zi = ((acc[2 + o] + s_part[0][row][2 + o]) + (s_part[1][row][2 + o] + s_part[2][row][2 + o])) +
                         a.params[M.b_off[L] + 2 * o + 1];
        w2d_re_tanh(y[o], zi, y[o], fac[2 * o], fac[2 * o + 1]);
      }
    if (a.train)      // (d out_0 / dx, d out_0 / dy, d out_1 / dx, d out_1 / dy) per row; WIRE2D does not use this slot otherwise
      reinterpret_cast<float4*>(a.ws + a.w.outacc)[static_cast<size_t>(tile) * kTileM + row] = make_float4(fac[0], fac[1], fac[2], fac[3]);
  }
  if (valid && a.out)
    for (int o = 0; o < M.out_f; ++o) a.out[static_cast<size_t>(grow) * M.out_f + o] = y[o];
  if (!a.train) return;
  float lA = 0.f, lB = 0.f, fs = 0.f, cnt = 0.f, amA = 0.f, amB = 0.f;
  float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid && a.gt && a.loss.kind != LOSS_NONE) {
    const bool in_loss = a.mask ? (a.mask[srow] != 0) : true;
    if (a.loss.kind == LOSS_HDR) {
      const float kx = a.coords[srow * 3 + 1], ky = a.coords[srow * 3 + 2];
      const float f = expf(-(kx * kx + ky * ky) / (2.f * a.loss.sigma * a.loss.sigma));
      fs = (1.f - f) * (1.f - f);
    }
    if (in_loss) {
      for (int o = 0; o < M.out_f; ++o) t[o] = a.gt[srow * M.out_f + o];
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
}

// ------------------------------------------------------------------------------------------------ backward of the final layer
__global__ void __launch_bounds__(256) w2d_blast_kernel(const __grid_constant__ WireAuxArgs a) {
  __shared__ float sWr[kMaxOut][kW2dMaxP], sWi[kMaxOut][kW2dMaxP];
  const WireModel& M = a.m;
  const int P = M.P, pg = P / 8, tile = blockIdx.x, tid = threadIdx.x;
  const int L = M.depth + 1;
  for (int i = tid; i < kMaxOut * P; i += 256) {
    const int o = i / P, j = i % P;
    const bool ok = o < M.out_f && j < M.c;
    sWr[o][j] = ok ? a.params[M.w_off[L] + (o * M.c + j) * 2] : 0.f;
    sWi[o][j] = ok ? a.params[M.w_off[L] + (o * M.c + j) * 2 + 1] : 0.f;
  }
  __syncthreads();
  if (blockIdx.y == 0 && tid < kWMaxDepth)      // hand-over counters of the chained dgrad GEMMs start at zero
    reinterpret_cast<unsigned int*>(a.ws + a.w.flags_bwd)[tid * a.w.n_tiles + tile] = 0u;
  const float* sc = reinterpret_cast<const float*>(a.ws + a.w.scal);
  const float S = sc[SC_SCALE], cA = sc[SC_CA], cB = sc[SC_CB];
  const float ratio = sc[SC_LAYER_SCALE + M.depth] / S;
  float amax = 0.f;
  const float w = M.depth >= 1 ? M.omega_hidden : M.omega_first, s2 = M.sigma * M.sigma;
  const size_t th = static_cast<size_t>(kTileM) * 2 * P * 2, tz = 2 * th;
  const uint8_t* yimg = a.ws + a.w.hhi[L] + tile * th;
  const uint8_t* zimg = a.ws + a.w.ab[M.depth] + tile * tz;
  uint8_t* dzimg = a.ws + a.w.dz[M.depth] + tile * tz;
  const bool complex_layer = M.depth >= 1;
  const int kg_per = pg / kW2dAuxSplit, kg0 = blockIdx.y * kg_per;
  for (int idx = tid; idx < kTileM * kg_per; idx += 256) {
    const int row = idx & (kTileM - 1), kg = kg0 + (idx >> 7);
    const int grow = tile * kTileM + row;
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
    // gradient of the final linear's pre-activation z = x + jy: (gx, gy) = dL/dout * (d out/dx, d out/dy); without the
    // tanh tail out = x, i.e. (dz, 0).  dz[0..1] = gx, dz[2..3] = gy from here on.
    if (M.last_tanh) {
      const float4 f = reinterpret_cast<const float4*>(a.ws + a.w.outacc)[static_cast<size_t>(tile) * kTileM + row];
      dz[2] = dz[0] * f.y; dz[3] = dz[1] * f.w;
      dz[0] *= f.x; dz[1] *= f.z;
    }
    const size_t ro = static_cast<size_t>(row) * 16;
    if (kg == 0) {
      uint8_t* zl = a.ws + a.w.dzlast + static_cast<size_t>(tile) * kDzLastBytes;
      st_global_v4(zl + ro, make_uint4(pack_h2(dz[0], dz[1]), pack_h2(dz[2], dz[3]), 0u, 0u));
      st_global_v4(zl + 2048 + ro, make_uint4(0u, 0u, 0u, 0u));
    }
    float yr[8], yi[8], za[8], zb[8], zc[8], zd[8], da[8], db[8], dc[8], dd[8];
    w2d_unpack8(ld_global_nc_v4(yimg + static_cast<size_t>(kg) * 2048 + ro), yr);
    w2d_unpack8(ld_global_nc_v4(yimg + static_cast<size_t>(pg + kg) * 2048 + ro), yi);
    w2d_unpack8(ld_global_nc_v4(zimg + static_cast<size_t>(kg) * 2048 + ro), za);
    w2d_unpack8(ld_global_nc_v4(zimg + static_cast<size_t>(pg + kg) * 2048 + ro), zb);
    w2d_unpack8(ld_global_nc_v4(zimg + static_cast<size_t>(2 * pg + kg) * 2048 + ro), zc);
    w2d_unpack8(ld_global_nc_v4(zimg + static_cast<size_t>(3 * pg + kg) * 2048 + ro), zd);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = kg * 8 + e;
      float gr = 0.f, gi = 0.f;     // dL/d Re(h_j), dL/d Im(h_j) for z = h W^T + b, dL/dz = gx + j gy
#pragma unroll
      for (int o = 0; o < 2; ++o)
        if (o < M.out_f) {
          gr = fmaf(dz[o], sWr[o][j], fmaf(dz[2 + o], sWi[o][j], gr));
          gi = fmaf(-dz[o], sWi[o][j], fmaf(dz[2 + o], sWr[o][j], gi));
        }
      const float Pp = gr * yr[e] + gi * yi[e];
      const float Q = gr * yi[e] - gi * yr[e];
      da[e] = ratio * (-2.f * s2 * za[e] * Pp - w * Q);
      db[e] = complex_layer ? ratio * (-(w + 2.f * s2 * zb[e]) * Pp) : 0.f;
      dc[e] = ratio * (-2.f * s2 * zc[e] * Pp);
      dd[e] = complex_layer ? ratio * (-2.f * s2 * zd[e] * Pp) : 0.f;
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(da[e]), fabsf(db[e])), fmaxf(fabsf(dc[e]), fabsf(dd[e]))));
    }
    st_global_v4(dzimg + static_cast<size_t>(kg) * 2048 + ro, w2d_pack8(da));
    st_global_v4(dzimg + static_cast<size_t>(pg + kg) * 2048 + ro, w2d_pack8(db));
    st_global_v4(dzimg + static_cast<size_t>(2 * pg + kg) * 2048 + ro, w2d_pack8(dc));
    st_global_v4(dzimg + static_cast<size_t>(3 * pg + kg) * 2048 + ro, w2d_pack8(dd));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if ((tid & 31) == 0 && amax > 0.f && isfinite(amax))
    atomicMax(reinterpret_cast<unsigned int*>(a.ws + a.w.scal) + SC_LAYER_AMAX + M.depth, __float_as_uint(amax));
}

cudaError_t launch_w2d_first(const WireAuxArgs& a, cudaStream_t st) {
  w2d_first_kernel<<<dim3(a.w.n_tiles, kW2dAuxSplit), 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_w2d_last(const WireAuxArgs& a, cudaStream_t st) {
  w2d_last_kernel<<<a.w.n_tiles, 512, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_w2d_blast(const WireAuxArgs& a, cudaStream_t st) {
  w2d_blast_kernel<<<dim3(a.w.n_tiles, kW2dAuxSplit), 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace inr
