// Fused optimiser step over the flat parameter buffer (torch.optim.Adam semantics, src/train.py:76,190):
//   g = (sum over split-K partials, fixed order) / S  [+ L1/L2 regulariser gradient, src/models/regularization.py]
//   Adam update in fp32, then the fp16 operand copies of every tensor-core weight are re-packed in the
//   same pass (forward stages, dgrad stages), so the next step's kernels never touch fp32 weights.
// Also here: the standalone packer (after load_state_dict) and the amax pre-pass for external dL/dout.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "inr_kernels.cuh"
#include "inr_loss.cuh"

namespace inr {

__device__ __forceinline__ int find_seg(const SegDesc* seg, int n_seg, int p) {
  int lo = 0, hi = n_seg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg[mid].off <= p) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ void pack_store(const SegDesc& s, uint8_t* wpack, int local, float val) {
  const int o = local / s.cols, i = local - o * s.cols;
  if (s.layout == 1) {      // lgemm operands: N-blocks of nt rows, K stages of 32 columns
    const uint32_t stage_b = static_cast<uint32_t>(s.nt) * 64;
    if (s.pack_fwd) {       // B[n = out, k = in]
      const int nb = o / s.nt, n = o % s.nt, st = i >> 5, kk = i & 31;
      *reinterpret_cast<__half*>(wpack + s.wf_off + (static_cast<size_t>(nb) * ((s.kpad ? s.kpad : s.cols) >> 5) + st) * stage_b + (kk >> 3) * (s.nt * 16) +
                                 n * 16 + (kk & 7) * 2) = __float2half_rn(val * s.fwd_scale);
    }
    if (s.pack_bwd) {       // B^T[n = in, k = out]
      const int nb = i / s.nt, n = i % s.nt, st = o >> 5, kk = o & 31;
      *reinterpret_cast<__half*>(wpack + s.wd_off + (static_cast<size_t>(nb) * (s.rows >> 5) + st) * stage_b + (kk >> 3) * (s.nt * 16) +
                                 n * 16 + (kk & 7) * 2) = __float2half_rn(val * s.bwd_scale);
    }
    return;
  }
  if (s.pack_fwd) {
    const __half h = __float2half_rn(val * s.fwd_scale);
    int kp = i;
    if (s.perm_e > 0) {
      if (i < s.perm_e) kp = 64 * (i >> 5) + (i & 31);
      else { const int j = i - s.perm_e; kp = 64 * (j >> 5) + 32 + (j & 31); }
    }
    const int st = kp >> 5, kk = kp & 31;
    *reinterpret_cast<__half*>(wpack + s.wf_off + static_cast<size_t>(st) * kStageBytes + (kk >> 3) * 4096 + o * 16 + (kk & 7) * 2) = h;
  }
  if (s.pack_bwd) {
    const __half h = __float2half_rn(val * s.bwd_scale);
    const int st = o >> 5, kk = o & 31;
    *reinterpret_cast<__half*>(wpack + s.wd_off + static_cast<size_t>(st) * kStageBytes + (kk >> 3) * 4096 + i * 16 + (kk & 7) * 2) = h;
  }
}

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a) {
  __shared__ float s_c[4];   // step_size, sqrt(bc2), inv_scale, -
  griddep_wait();            // split-K partials of wgrad (or the exchanged gradients) complete
  if (threadIdx.x == 0) {
    const float* sc = a.scal;
    s_c[2] = sc ? sc[SC_INV_SCALE] : 1.f;
    if (a.do_adam && a.scal_has_bc && sc) {
      s_c[0] = sc[SC_STEP_SIZE]; s_c[1] = sc[SC_BC2_SQRT];
    } else if (a.do_adam) {
      const int t = *a.step + (a.peer.n_ranks > 0 && a.peer.done ? 1 : 0);
      const double b1 = a.hyper[1], b2 = a.hyper[2];
      const double bc1 = 1.0 - pow(b1, static_cast<double>(t));
      const double bc2 = 1.0 - pow(b2, static_cast<double>(t));
      s_c[0] = static_cast<float>(static_cast<double>(a.hyper[0]) / bc1);
      s_c[1] = static_cast<float>(sqrt(bc2));
    }
    if (blockIdx.x == 0) {
      if (a.loss_out && sc) *a.loss_out = sc[SC_LOSS];
      if (a.row_offset) *a.row_offset += a.row_advance;
    }
  }
  __syncthreads();
  // peer mode (data-parallel exchange fused in): one CTA owns 1024 consecutive parameters -- gather, then 4 passes
  __shared__ __align__(16) float s_g[4 * 256];
  const bool peer = a.peer.n_ranks > 0;
  const int reps = peer ? 4 : 1;
  const size_t base = static_cast<size_t>(blockIdx.x) * 256 * reps;
  const int t_new = peer ? *a.step + (a.peer.done ? 1 : 0) : 0;
  if (peer) {
    peer_barrier(a.peer, static_cast<unsigned int>(t_new));
    peer_gather(a.peer, base, static_cast<size_t>((a.n_params + 3) & ~3), s_g);
  }
  for (int j = 0; j < reps; ++j) {
    const int p = static_cast<int>(base) + j * 256 + threadIdx.x;
    if (p >= a.n_params) break;
    const int si = find_seg(a.seg, a.n_seg, p);
    const SegDesc& sg = a.seg[si];
    if (sg.frozen) { if (a.grads) a.grads[p] = 0.f; continue; }
    const float* gp = a.gpart;
    // everything this parameter needs is requested up front, so the latencies of the master weight, the two moments,
    // the layer scale and the split partials overlap instead of queueing behind each other
    float w = a.params[p];
    const float m_old = a.do_adam ? a.m[p] : 0.f, v_old = a.do_adam ? a.v[p] : 0.f;
    const float gscale = (sg.scale_slot >= 0 && a.scal) ? 1.f / a.scal[sg.scale_slot] : s_c[2];
    float g = 0.f;
    if (peer) g = s_g[j * 256 + threadIdx.x];
    else if (sg.gfin_off >= 0 && a.gfin) g = a.gfin[sg.gfin_off + (p - sg.off)];
    else
    {   // split partials in split order (bit-reproducible); eight loads in flight per thread
      int sp = 0;
      for (; sp + 8 <= a.n_split; sp += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = gp[static_cast<size_t>(sp + q) * a.gstride + p];
#pragma unroll
        for (int q = 0; q < 8; ++q) g += v[q];
      }
      for (; sp + 4 <= a.n_split; sp += 4) {
        const float v0 = gp[static_cast<size_t>(sp) * a.gstride + p], v1 = gp[static_cast<size_t>(sp + 1) * a.gstride + p];
        const float v2 = gp[static_cast<size_t>(sp + 2) * a.gstride + p], v3 = gp[static_cast<size_t>(sp + 3) * a.gstride + p];
        g += v0; g += v1; g += v2; g += v3;
      }
      for (; sp < a.n_split; ++sp) g += gp[static_cast<size_t>(sp) * a.gstride + p];
    }
    g *= gscale;
    if (a.do_adam) {
      const float l1 = a.hyper[5], l2 = a.hyper[6];
      if (l1 != 0.f) g += l1 * (w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f));
      if (l2 != 0.f) g += 2.f * l2 * w;
    }
    if (a.grads) a.grads[p] = g;
    if (!a.do_adam) continue;
    const float b1 = a.hyper[1], b2 = a.hyper[2], eps = a.hyper[3], wd = a.hyper[4];
    if (wd != 0.f) g = fmaf(wd, w, g);
    const float m = b1 * m_old + (1.f - b1) * g;
    const float v = b2 * v_old + (1.f - b2) * g * g;
    a.m[p] = m; a.v[p] = v;
    const float denom = sqrtf(v) / s_c[1] + eps;
    w = w - s_c[0] * (m / denom);
    a.params[p] = w;
    if (sg.pack_fwd | sg.pack_bwd) pack_store(sg, a.wpack, p - sg.off, w);
  }
  if (peer) peer_finish_step(a.peer, a.step, t_new);
}

__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ AdamArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.n_params) return;
  const int si = find_seg(a.seg, a.n_seg, p);
  const SegDesc& sg = a.seg[si];
  if (sg.pack_fwd | sg.pack_bwd) pack_store(sg, a.wpack, p - sg.off, a.params[p]);
}

// Tile partials for an externally supplied dL/dout (autograd path): amax only.
__global__ void __launch_bounds__(128) dout_amax_kernel(const float* dout, int bs, int out_f, float* partials) {
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
    p[0] = 0.f; p[1] = 0.f; p[2] = 0.f; p[3] = 0.f;
    p[4] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    p[5] = 0.f; p[6] = 0.f; p[7] = 0.f;
  }
}

// Total-variation term of the per-coil training loop (reference src/train.py:173-174, losses.py:326-343):
//   tv = weight * (L1mean(img[:, :-1] - img[:, 1:]) + L1mean(img[:-1] - img[1:])),  img = out.view(H, W, out_f)
// evaluated on ALL rows of the batch (before the undersampling mask selects rows for the main loss).  One thread per
// row writes d tv / d out into the B pieces of the row's loss record; every pair difference is counted by its
// upper-left element, tile sums in fixed order (bit-reproducible).
__global__ void __launch_bounds__(kTileM) tv_kernel(const TvArgs a) {
  __shared__ float red[4][2];
  const int tid = threadIdx.x, row = blockIdx.x * kTileM + tid;
  float lossn = 0.f, am = 0.f, t[2] = {0.f, 0.f};
  if (row < a.bs) {
    const int h = row / a.w, w = row % a.w;
    const float cw = a.w > 1 ? a.weight / (static_cast<float>(a.h) * (a.w - 1) * a.out_f) : 0.f;
    const float ch = a.h > 1 ? a.weight / (static_cast<float>(a.h - 1) * a.w * a.out_f) : 0.f;
    auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
    for (int o = 0; o < a.out_f && o < 2; ++o) {
      const float x = a.out[static_cast<size_t>(row) * a.out_f + o];
      if (w < a.w - 1) { const float d = x - a.out[static_cast<size_t>(row + 1) * a.out_f + o]; lossn += cw * fabsf(d); t[o] += cw * sgn(d); }
      if (w > 0) { const float d = a.out[static_cast<size_t>(row - 1) * a.out_f + o] - x; t[o] -= cw * sgn(d); }
      if (h < a.h - 1) { const float d = x - a.out[static_cast<size_t>(row + a.w) * a.out_f + o]; lossn += ch * fabsf(d); t[o] += ch * sgn(d); }
      if (h > 0) { const float d = a.out[static_cast<size_t>(row - a.w) * a.out_f + o] - x; t[o] -= ch * sgn(d); }
    }
    am = fmaxf(fabsf(t[0]), fabsf(t[1]));
  }
  float4* gr = reinterpret_cast<float4*>(a.g) + row;
  float4 v = *gr; v.z = t[0]; v.w = t[1]; *gr = v;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { lossn += __shfl_xor_sync(0xffffffffu, lossn, off); am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off)); }
  if ((tid & 31) == 0) { red[tid >> 5][0] = lossn; red[tid >> 5][1] = am; }
  __syncthreads();
  if (tid == 0) {
    float* q = a.part + static_cast<size_t>(blockIdx.x) * kPartialsPerTile;
    q[1] = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
    q[5] = fmaxf(fmaxf(red[0][1], red[1][1]), fmaxf(red[2][1], red[3][1]));
  }
}

cudaError_t launch_tv(const TvArgs& a, int n_tiles, cudaStream_t st) {
  tv_kernel<<<n_tiles, kTileM, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_adam(const AdamArgs& a, cudaStream_t stream) {
  const int per_cta = a.peer.n_ranks > 0 ? 1024 : 256;
  const int grid = (a.n_params + per_cta - 1) / per_cta;
  return launch_dependent(adam_kernel, dim3(grid), dim3(256), 0, stream, a);
}
cudaError_t launch_pack(const AdamArgs& a, cudaStream_t stream) {
  const int grid = (a.n_params + 255) / 256;
  pack_kernel<<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_dout_amax(const float* dout, int bs, int out_f, float* partials, int n_tiles, cudaStream_t stream) {
  dout_amax_kernel<<<n_tiles, 128, 0, stream>>>(dout, bs, out_f, partials);
  return cudaGetLastError();
}

}  // namespace inr
