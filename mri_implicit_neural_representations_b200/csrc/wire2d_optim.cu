// WIRE2D optimiser / packer (same role as wire_optim.cu): one thread per real parameter component of the flat buffer.
// Per layer the reference registers omega_0, scale_0 (frozen), linear.{weight,bias}, scale_orth.{weight,bias}
// (wire2d.py:35-46); layer 0 is real, hidden layers complex, the final layer a complex nn.Linear.
//   hidden-layer gradients from the split-K block D = dZ^T [hr|hi]  ([4P][2P] + bias [4P], rows a | b | c | d):
//     dWr[o,i] = D[o,i] + D[P+o,P+i]      dWi[o,i] = D[P+o,i] - D[o,P+i]        (linear)
//     dVr[o,i] = D[2P+o,i] + D[3P+o,P+i]  dVi[o,i] = D[3P+o,i] - D[2P+o,P+i]    (scale_orth)
//   first layer (real): D0 [4P][16] against the coordinate image [x_hi(3), 1, x_lo(3), 0]: dW[o,c] = D0[o,c] + D0[o,4+c],
//     db[o] = D0[o,3]; scale_orth from the rows 2P+o.   final layer: as WIRE (DT [16][2P] + bias); with the complex tanh tail
//     rows 2 + o of DT hold the gradient of Im z_o:  dWr[o,j] = DT[o,j] + DT[2+o,P+j],  dWi[o,j] = -DT[o,P+j] + DT[2+o,j].
//   packed operands (fp16): forward blocks of 64 output features, rows [a | b | c | d] x K = 2P with
//     a_o = [Wr_o | -Wi_o], b_o = [Wi_o | Wr_o], c_o = [Vr_o | -Vi_o], d_o = [Vi_o | Vr_o]  (hi + lo copies);
//   dgrad blocks of 128 input features, rows [dhr | dhi] x K = 4P with
//     dhr_i = [Wr_:i | Wi_:i | Vr_:i | Vi_:i],  dhi_i = [-Wi_:i | Wr_:i | -Vi_:i | Vr_:i].
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "wire.cuh"
#include "inr_loss.cuh"

namespace inr {

// byte offset of element (n, k) inside one N-block image [K/32 stages][256 x 32]
__device__ __forceinline__ uint32_t w2d_blk_off(int n, int k) {
  return static_cast<uint32_t>(k >> 5) * (kW2dNT * 64) + ((k & 31) >> 3) * (kW2dNT * 16) + n * 16 + (k & 7) * 2;
}
__device__ __forceinline__ void w2d_put(uint8_t* hi, uint8_t* lo, uint32_t off, float v) {
  const __half h = __float2half_rn(v);
  *reinterpret_cast<__half*>(hi + off) = h;
  if (lo) *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
}

// lin: 0 linear, 1 scale_orth; comp: 0 real part, 1 imaginary part of the complex weight [o][i]
__device__ void w2d_pack_hidden(const WireModel& M, uint8_t* wpack, int l, int lin, int o, int i, int comp, float v) {
  const int P = M.P;
  uint8_t* fh = wpack + M.wf_hi[l];
  uint8_t* fl = wpack + M.wf_lo[l];
  uint8_t* dh = wpack + M.wd_hi[l];
  const uint32_t fblk = static_cast<uint32_t>(2 * P / 32) * (kW2dNT * 64);     // forward block bytes (K = 2P)
  const uint32_t dblk = static_cast<uint32_t>(4 * P / 32) * (kW2dNT * 64);     // dgrad block bytes (K = 4P)
  const int nbo = o / kW2dFwdFeat, no = o % kW2dFwdFeat;
  const int nbi = i / kW2dBwdFeat, ni = i % kW2dBwdFeat;
  const int ra = (2 * lin) * kW2dFwdFeat + no, rb = (2 * lin + 1) * kW2dFwdFeat + no;     // forward rows a/c and b/d
  const int kr = (2 * lin) * P + o, ki = (2 * lin + 1) * P + o;                          // dgrad K index of d(a|c)_o and d(b|d)_o
  if (comp == 0) {   // real part
    w2d_put(fh, fl, nbo * fblk + w2d_blk_off(ra, i), v);
    w2d_put(fh, fl, nbo * fblk + w2d_blk_off(rb, P + i), v);
    w2d_put(dh, nullptr, nbi * dblk + w2d_blk_off(ni, kr), v);
    w2d_put(dh, nullptr, nbi * dblk + w2d_blk_off(kW2dBwdFeat + ni, ki), v);
  } else {           // imaginary part
    w2d_put(fh, fl, nbo * fblk + w2d_blk_off(ra, P + i), -v);
    w2d_put(fh, fl, nbo * fblk + w2d_blk_off(rb, i), v);
    w2d_put(dh, nullptr, nbi * dblk + w2d_blk_off(ni, ki), v);
    w2d_put(dh, nullptr, nbi * dblk + w2d_blk_off(kW2dBwdFeat + ni, kr), -v);
  }
}

struct W2dLoc { int layer, kind, lin, idx; };   // kind: 0 weight, 1 bias, 2 frozen, -1 none
__device__ __forceinline__ W2dLoc w2d_locate(const WireModel& M, int p) {
  const int L = M.depth + 1;
  for (int l = 0; l <= L; ++l) {
    if (l < L && (p == M.omega_off[l] || p == M.scale_off[l])) return {l, 2, 0, 0};
    const int wn = (l == 0) ? M.c * M.in_f : (l == L ? M.out_f * M.c * 2 : M.c * M.c * 2);
    const int bn = (l == 0) ? M.c : (l == L ? M.out_f * 2 : M.c * 2);
    if (p >= M.w_off[l] && p < M.w_off[l] + wn) return {l, 0, 0, p - M.w_off[l]};
    if (p >= M.b_off[l] && p < M.b_off[l] + bn) return {l, 1, 0, p - M.b_off[l]};
    if (l < L) {
      if (p >= M.v_off[l] && p < M.v_off[l] + wn) return {l, 0, 1, p - M.v_off[l]};
      if (p >= M.vb_off[l] && p < M.vb_off[l] + bn) return {l, 1, 1, p - M.vb_off[l]};
    }
  }
  return {-1, -1, 0, 0};
}

__global__ void __launch_bounds__(256) w2d_adam_kernel(const __grid_constant__ WireAdamArgs a) {
  __shared__ float s_c[4];
  const WireModel& M = a.m;
  if (threadIdx.x == 0) {
    const float* sc = a.scal;
    s_c[2] = sc ? sc[SC_INV_SCALE] : 1.f;
    if (a.do_adam && a.scal_has_bc && sc) {
      s_c[0] = sc[SC_STEP_SIZE]; s_c[1] = sc[SC_BC2_SQRT];
    } else if (a.do_adam) {
      const double t = static_cast<double>(*a.step);
      s_c[0] = static_cast<float>(static_cast<double>(a.hyper[0]) / (1.0 - pow(static_cast<double>(a.hyper[1]), t)));
      s_c[1] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(a.hyper[2]), t)));
    }
    if (blockIdx.x == 0 && !a.pack_only) {
      if (a.loss_out && sc) *a.loss_out = sc[SC_LOSS];
      if (a.row_offset) *a.row_offset += a.row_advance;
    }
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= M.n_params) return;
  const int L = M.depth + 1, P = M.P, K2 = 2 * P;
  const W2dLoc loc = w2d_locate(M, p);
  if (loc.kind == 2 || loc.kind < 0) { if (a.grads && !a.pack_only) a.grads[p] = 0.f; return; }
  const int layer = loc.layer, idx = loc.idx, lin = loc.lin;
  float w = a.params[p];
  int o = 0, i = 0, comp = 0;
  if (layer >= 1 && layer < L && loc.kind == 0) { const int e = idx >> 1; comp = idx & 1; o = e / M.c; i = e % M.c; }
  if (!a.pack_only) {
    float inv_scale = s_c[2];
    if (a.scal && layer < L) inv_scale = 1.f / a.scal[SC_LAYER_SCALE + layer];
    const int r0 = 2 * lin * P;                  // row of d(a|c) in the D blocks; d(b|d) follows at + P
    float g = 0.f;
    for (int s = 0; s < a.n_split; ++s) {
      const float* G = a.gpart + static_cast<size_t>(s) * M.gd_floats;
      if (layer == 0) {
        const float* D0 = G + M.gd_first;
        if (loc.kind == 0) { const int oo = idx / M.in_f, cc = idx % M.in_f; g += D0[(r0 + oo) * 16 + cc] + D0[(r0 + oo) * 16 + 4 + cc]; }
        else g += D0[(r0 + idx) * 16 + 3];
      } else if (layer < L) {
        const float* D = G + M.gd_hidden[layer];
        if (loc.kind == 0) {
          g += comp == 0 ? D[static_cast<size_t>(r0 + o) * K2 + i] + D[static_cast<size_t>(r0 + P + o) * K2 + P + i]
                         : D[static_cast<size_t>(r0 + P + o) * K2 + i] - D[static_cast<size_t>(r0 + o) * K2 + P + i];
        } else {
          const int oo = idx >> 1;
          g += D[static_cast<size_t>(4 * P) * K2 + r0 + ((idx & 1) ? P + oo : oo)];
        }
      } else {
        const float* DT = G + M.gd_final;
        if (loc.kind == 0) {
          const int e = idx >> 1, oo = e / M.c, j = e % M.c;
          g += (idx & 1) ? -DT[oo * K2 + P + j] : DT[oo * K2 + j];
          // tanh tail: rows 2 + o of DT carry gy (gradient of Im z):  dWr += gy . hi,  dWi += gy . hr
          if (M.last_tanh) g += (idx & 1) ? DT[(2 + oo) * K2 + j] : DT[(2 + oo) * K2 + P + j];
        } else {
          g += (idx & 1) ? (M.last_tanh ? DT[16 * K2 + 2 + (idx >> 1)] : 0.f) : DT[16 * K2 + (idx >> 1)];
        }
      }
    }
    g *= inv_scale;
    if (a.grads) a.grads[p] = g;
    if (!a.do_adam) return;
    const float b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], wd = a.hyper[4];
    if (wd != 0.f) g = fmaf(wd, w, g);
    const float m = b1 * a.mom[p] + (1.f - b1) * g;
    const float v = b2 * a.var[p] + (1.f - b2) * g * g;
    a.mom[p] = m; a.var[p] = v;
    w = w - s_c[0] * (m / (sqrtf(v) / s_c[1] + eps));
    a.params[p] = w;
  }
  if (layer >= 1 && layer < L && loc.kind == 0) w2d_pack_hidden(M, a.wpack, layer, lin, o, i, comp, w);
}

// Adam on externally reduced gradients (data-parallel path, optionally with the peer exchange fused in).
__global__ void __launch_bounds__(256) w2d_adam_flat_kernel(const __grid_constant__ WireAdamArgs a) {
  __shared__ float s_c[2];
  const WireModel& M = a.m;
  griddep_wait();            // launched as a programmatic dependent of the kernel that reduced this rank's gradients
  const int t_new = *a.step + ((a.peer.n_ranks > 0 && a.peer.done) ? 1 : 0);
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(t_new);
    s_c[0] = static_cast<float>(static_cast<double>(a.hyper[0]) / (1.0 - pow(static_cast<double>(a.hyper[1]), t)));
    s_c[1] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(a.hyper[2]), t)));
  }
  __syncthreads();
  __shared__ __align__(16) float s_g[4 * 256];
  const bool peer = a.peer.n_ranks > 0;
  const int reps = peer ? 4 : 1;
  const size_t base = static_cast<size_t>(blockIdx.x) * 256 * reps;
  if (peer) {
    peer_barrier(a.peer, static_cast<unsigned int>(t_new));
    peer_gather(a.peer, base, static_cast<size_t>((M.n_params + 3) & ~3), s_g);
  }
  const int L = M.depth + 1;
  for (int j = 0; j < reps; ++j) {
    const int p = static_cast<int>(base) + j * 256 + threadIdx.x;
    if (p >= M.n_params) break;
    const W2dLoc loc = w2d_locate(M, p);
    if (loc.kind == 2 || loc.kind < 0) continue;
    float g = peer ? s_g[j * 256 + threadIdx.x] : a.gpart[p], w = a.params[p];
    const float b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], wd = a.hyper[4];
    if (wd != 0.f) g = fmaf(wd, w, g);
    const float m = b1 * a.mom[p] + (1.f - b1) * g;
    const float v = b2 * a.var[p] + (1.f - b2) * g * g;
    a.mom[p] = m; a.var[p] = v;
    w = w - s_c[0] * (m / (sqrtf(v) / s_c[1] + eps));
    a.params[p] = w;
    if (loc.layer >= 1 && loc.layer < L && loc.kind == 0) {
      const int e = loc.idx >> 1;
      w2d_pack_hidden(M, a.wpack, loc.layer, loc.lin, e / M.c, e % M.c, loc.idx & 1, w);
    }
  }
  if (peer) peer_finish_step(a.peer, a.step, t_new);
}

cudaError_t launch_w2d_adam(const WireAdamArgs& a, cudaStream_t st) {
  w2d_adam_kernel<<<(a.m.n_params + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_w2d_adam_flat(const WireAdamArgs& a, cudaStream_t st) {
  const int per_cta = a.peer.n_ranks > 0 ? 1024 : 256;
  return launch_dependent(w2d_adam_flat_kernel, dim3((a.m.n_params + per_cta - 1) / per_cta), dim3(256), 0, st, a);
}

}  // namespace inr
