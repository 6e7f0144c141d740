// WIRE optimiser / packer: one thread per real parameter component of the flat buffer (complex tensors are interleaved
// (re, im) pairs updated as independent reals, exactly what torch.optim.Adam does through view_as_real).
//   gradient gather from the split-K "virtual" blocks D = dZ^T [hr|hi] (see wire.cuh):
//     dL/dWr[o,i] = D[o,i] + D[P+o,P+i]       dL/dWi[o,i] = D[P+o,i] - D[o,P+i]          (P = 192)
//     final layer (out = Re(h W^T + b)):  dL/dWr[o,j] = DT[o,j],  dL/dWi[o,j] = -DT[o,P+j],  dL/d Im(b) = 0
//     first layer (real):  dL/dW[o,c] = D0[o,c] + D0[o,4+c] (hi + lo coordinate columns),  dL/db[o] = D0[o,3]
//   then torch-semantics Adam and re-packing of the fp16 hi/lo block operands of the hidden layers:
//     forward  rows  a_j = [Wr_j | -Wi_j],  b_j = [Wi_j | Wr_j]            (N-blocks of 96 features: a-rows then b-rows)
//     dgrad    rows  dhr_i = [Wr_:i | Wi_:i],  dhi_i = [-Wi_:i | Wr_:i]
// Frozen parameters (omega_0, scale_0; reference networks.py:191-192) are never touched.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "wire.cuh"
#include "inr_loss.cuh"

namespace inr {

constexpr uint32_t kWBlockBytesC = (kW2 / kStageK) * kWStageBBytes;   // 147456 per N-block
// byte offset of element (n, k) inside one N-block image, pair layout: [half n / 96][k-group k / 8][96 rows][8 K] -- the CTA
// pair of lgemm_kernel<.., PAIR = 1> streams one contiguous half each (rows 0..95 = the "a" rows, 96..191 = the "b" rows)
__device__ __forceinline__ uint32_t wblk_off(int n, int k) {
  const int h = n / kWFeatPerBlock, nl = n - h * kWFeatPerBlock;
  return static_cast<uint32_t>(h) * (kWBlockBytesC / 2) + (k >> 3) * (kWFeatPerBlock * 16) + nl * 16 + (k & 7) * 2;
}
constexpr uint32_t kWBlockBytes = (kW2 / kStageK) * kWStageBBytes;   // 147456 per N-block

__device__ __forceinline__ void put_hi_lo(uint8_t* hi, uint8_t* lo, uint32_t off, float v) {
  const __half h = __float2half_rn(v);
  *reinterpret_cast<__half*>(hi + off) = h;
  if (lo) *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
}

__device__ void wire_pack_hidden(const WireModel& M, uint8_t* wpack, int l, int o, int i, int comp, float v) {
  uint8_t* fh = wpack + M.wf_hi[l];
  uint8_t* fl = wpack + M.wf_lo[l];
  uint8_t* dh = wpack + M.wd_hi[l];
  const int nbo = o / kWFeatPerBlock, no = o % kWFeatPerBlock;     // forward: N over output features
  const int nbi = i / kWFeatPerBlock, ni = i % kWFeatPerBlock;     // dgrad:   N over input features
  if (comp == 0) {   // Wr
    put_hi_lo(fh, fl, nbo * kWBlockBytes + wblk_off(no, i), v);
    put_hi_lo(fh, fl, nbo * kWBlockBytes + wblk_off(kWFeatPerBlock + no, kWP + i), v);
    put_hi_lo(dh, nullptr, nbi * kWBlockBytes + wblk_off(ni, o), v);
    put_hi_lo(dh, nullptr, nbi * kWBlockBytes + wblk_off(kWFeatPerBlock + ni, kWP + o), v);
  } else {           // Wi
    put_hi_lo(fh, fl, nbo * kWBlockBytes + wblk_off(no, kWP + i), -v);
    put_hi_lo(fh, fl, nbo * kWBlockBytes + wblk_off(kWFeatPerBlock + no, i), v);
    put_hi_lo(dh, nullptr, nbi * kWBlockBytes + wblk_off(ni, kWP + o), v);
    put_hi_lo(dh, nullptr, nbi * kWBlockBytes + wblk_off(kWFeatPerBlock + ni, o), -v);
  }
}

__global__ void __launch_bounds__(256) wire_adam_kernel(const __grid_constant__ WireAdamArgs a) {
  __shared__ float s_c[4];
  const WireModel& M = a.m;
  // Ahead of the wait on wgrad: everything that no kernel of this step writes -- the tensor this thread's parameter belongs
  // to, the master weight and the Adam moments (last written by the previous step's optimiser, which completed before this
  // step's first kernel started).  Scalars, step counter and split-K partials are read behind the wait.
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int layer = -1, kind = -1, idx = 0;     // kind: 0 weight, 1 bias, 2 frozen
  const int L = M.depth + 1;
  const bool upd = !a.pack_only && a.do_adam;
  float w = 0.f, m_old = 0.f, v_old = 0.f;
  if (p < M.n_params) {
    for (int l = 0; l <= L; ++l) {
      const int wn = (l == 0) ? M.c * M.in_f : (l == L ? M.out_f * M.c * 2 : M.c * M.c * 2);
      const int bn = (l == 0) ? M.c : (l == L ? M.out_f * 2 : M.c * 2);
      if (l < L && (p == M.omega_off[l] || p == M.scale_off[l])) { layer = l; kind = 2; break; }
      if (p >= M.w_off[l] && p < M.w_off[l] + wn) { layer = l; kind = 0; idx = p - M.w_off[l]; break; }
      if (p >= M.b_off[l] && p < M.b_off[l] + bn) { layer = l; kind = 1; idx = p - M.b_off[l]; break; }
    }
    if (a.params_stable && (kind == 0 || kind == 1)) {
      w = a.params[p];
      if (upd) { m_old = a.mom[p]; v_old = a.var[p]; }
    }
  }
  griddep_wait();                    // split-K partials of wgrad complete
  if (!a.params_stable && p < M.n_params && (kind == 0 || kind == 1)) {     // stand-alone calls: whatever precedes may have written them
    w = a.params[p];
    if (upd) { m_old = a.mom[p]; v_old = a.var[p]; }
  }
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
  if (p >= M.n_params) return;
  if (kind == 2 || kind < 0) { if (a.grads && !a.pack_only) a.grads[p] = 0.f; return; }
  int o = 0, i = 0, comp = 0;
  if (layer >= 1 && layer < L && kind == 0) { const int e = idx >> 1; comp = idx & 1; o = e / M.c; i = e % M.c; }
  if (!a.pack_only) {
    // ---- gradient gather (fixed split order); every layer's block carries that layer's own dZ scale
    float inv_scale = s_c[2];
    if (a.scal && layer < L) inv_scale = 1.f / a.scal[SC_LAYER_SCALE + layer];
    // the two D entries (or one) this parameter gathers are the same for every split: resolve them once, then walk the
    // split copies four at a time so eight independent loads are in flight (summation order unchanged: split order)
    int i0 = 0, i1 = -1;
    float sg1 = 1.f, sg0 = 1.f;
    if (layer == 0) {
      if (kind == 0) { const int oo = idx / M.in_f, cc = idx % M.in_f; i0 = M.gd_first + oo * 16 + cc; i1 = i0 + 4; }
      else i0 = M.gd_first + idx * 16 + 3;
    } else if (layer < L) {
      if (kind == 0) {
        if (comp == 0) { i0 = M.gd_hidden[layer] + o * kW2 + i; i1 = M.gd_hidden[layer] + (kWP + o) * kW2 + kWP + i; }
        else { i0 = M.gd_hidden[layer] + (kWP + o) * kW2 + i; i1 = M.gd_hidden[layer] + o * kW2 + kWP + i; sg1 = -1.f; }
      } else {
        const int oo = idx >> 1;
        i0 = M.gd_hidden[layer] + kW2 * kW2 + ((idx & 1) ? kWP + oo : oo);
      }
    } else {
      if (kind == 0) {
        const int e = idx >> 1, oo = e / M.c, j = e % M.c;
        if (idx & 1) { i0 = M.gd_final + oo * kW2 + kWP + j; sg0 = -1.f; } else i0 = M.gd_final + oo * kW2 + j;
      } else {
        if (idx & 1) sg0 = 0.f;      // Im(b) of the final layer never reaches the real output
        i0 = M.gd_final + 16 * kW2 + (idx >> 1);
      }
    }
    float g = 0.f;
    {
      const float* G = a.gpart;
      const size_t st = M.gd_floats;
      int sp = 0;
      for (; sp + 8 <= a.n_split; sp += 8) {      // 16 independent loads in flight
        float v[8], u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { v[q] = G[(sp + q) * st + i0]; u[q] = i1 >= 0 ? G[(sp + q) * st + i1] : 0.f; }
#pragma unroll
        for (int q = 0; q < 8; ++q) g += sg0 * v[q] + sg1 * u[q];
      }
      for (; sp + 4 <= a.n_split; sp += 4) {
        float v[4], u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = G[(sp + q) * st + i0]; u[q] = i1 >= 0 ? G[(sp + q) * st + i1] : 0.f; }
#pragma unroll
        for (int q = 0; q < 4; ++q) g += sg0 * v[q] + sg1 * u[q];
      }
      for (; sp < a.n_split; ++sp) g += sg0 * G[sp * st + i0] + (i1 >= 0 ? sg1 * G[sp * st + i1] : 0.f);
    }
    g *= inv_scale;
    if (a.grads) a.grads[p] = g;
    if (!a.do_adam) return;
    const float b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], wd = a.hyper[4];
    if (wd != 0.f) g = fmaf(wd, w, g);
    const float m = b1 * m_old + (1.f - b1) * g;
    const float v = b2 * v_old + (1.f - b2) * g * g;
    a.mom[p] = m; a.var[p] = v;
    w = w - s_c[0] * (m / (sqrtf(v) / s_c[1] + eps));
    a.params[p] = w;
  }
  if (layer >= 1 && layer < L && kind == 0) wire_pack_hidden(M, a.wpack, layer, o, i, comp, w);
}

// Adam on externally reduced gradients (data-parallel path): gpart = plain flat gradients in parameter order.
__global__ void __launch_bounds__(256) wire_adam_flat_kernel(const __grid_constant__ WireAdamArgs a) {
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
    int layer = -1, idx = 0;
    bool frozen = false;
    for (int l = 0; l <= L; ++l) {
      if (l < L && (p == M.omega_off[l] || p == M.scale_off[l])) { frozen = true; break; }
      if (l >= 1 && l < L && p >= M.w_off[l] && p < M.w_off[l] + M.c * M.c * 2) { layer = l; idx = p - M.w_off[l]; break; }
    }
    if (frozen) continue;
    float g = peer ? s_g[j * 256 + threadIdx.x] : a.gpart[p], w = a.params[p];
    const float b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], wd = a.hyper[4];
    if (wd != 0.f) g = fmaf(wd, w, g);
    const float m = b1 * a.mom[p] + (1.f - b1) * g;
    const float v = b2 * a.var[p] + (1.f - b2) * g * g;
    a.mom[p] = m; a.var[p] = v;
    w = w - s_c[0] * (m / (sqrtf(v) / s_c[1] + eps));
    a.params[p] = w;
    if (layer >= 1) { const int e = idx >> 1; wire_pack_hidden(M, a.wpack, layer, e / M.c, e % M.c, idx & 1, w); }
  }
  if (peer) peer_finish_step(a.peer, a.step, t_new);
}

cudaError_t launch_wire_adam(const WireAdamArgs& a, cudaStream_t st) {
  return launch_dependent(wire_adam_kernel, dim3((a.m.n_params + 255) / 256), dim3(256), 0, st, a);
}
cudaError_t launch_wire_adam_flat(const WireAdamArgs& a, cudaStream_t st) {
  const int per_cta = a.peer.n_ranks > 0 ? 1024 : 256;
  return launch_dependent(wire_adam_flat_kernel, dim3((a.m.n_params + per_cta - 1) / per_cta), dim3(256), 0, st, a);
}

}  // namespace inr
