// Loss pieces shared by every model family: per-row loss numerators / unnormalised gradients (forward side) and the
// fixed-order reduction of the per-tile partials into the step scalars (backward side).
#pragma once
#include <cmath>
#include "inr_kernels.cuh"
#include "inr_ptx.cuh"

namespace inr {

__device__ __forceinline__ float tanh_acc(float x) { return tanhf(x); }

// Per-row loss pieces.  y = network output, t = target.  Returns unnormalised dL/dy parts gA, gB and
// loss numerators; the normalisation by the (masked) row count happens in the backward prologue.
struct RowLoss {
  float lossA, lossB, gA[kMaxOut], gB[kMaxOut];
};
__device__ __forceinline__ RowLoss loss_row(const LossDesc& L, int out_f, const float* y, const float* t) {
  RowLoss r;
  r.lossA = 0.f; r.lossB = 0.f;
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) { r.gA[o] = 0.f; r.gB[o] = 0.f; }
  switch (L.kind) {
    case LOSS_L2:
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) if (o < out_f) { float e = y[o] - t[o]; r.lossA += e * e; r.gA[o] = e; }
      break;
    case LOSS_L1:
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) if (o < out_f) {
        float e = y[o] - t[o]; r.lossA += fabsf(e); r.gA[o] = (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f);
      }
      break;
    case LOSS_MSLE:   // src/metrics/losses.py:26 (NaN for arguments <= 0, as the reference)
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) if (o < out_f) {
        float ax = y[o] + 1.f + 1e-9f;
        float d = logf(ax) - logf(t[o] + 1.f + 1e-9f);
        r.lossA += d * d; r.gA[o] = d / ax;
      }
      break;
    case LOSS_TANH:   // src/metrics/losses.py:131
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o) if (o < out_f) {
        float tx = tanh_acc(y[o]), ty = tanh_acc(t[o]);
        float d = tx - ty; r.lossA += d * d; r.gA[o] = d * (1.f - tx * tx);
      }
      break;
    case LOSS_LSL: {  // src/metrics/losses.py:221-223, complex pairs, |x| detached
      float e0 = y[0] - t[0], e1 = y[1] - t[1];
      float d = sqrtf(y[0] * y[0] + y[1] * y[1]) + L.eps;
      float inv = 1.f / (d * d);
      r.lossA = (e0 * e0 + e1 * e1) * inv; r.gA[0] = e0 * inv; r.gA[1] = e1 * inv;
    } break;
    case LOSS_HDR: {  // src/metrics/losses.py:250-259 in separable form
      float e0 = y[0] - t[0], e1 = y[1] - t[1];
      float ae2 = e0 * e0 + e1 * e1;
      float ax2 = y[0] * y[0] + y[1] * y[1];
      float d = sqrtf(ax2) + L.eps;
      float lg = logf(sqrtf(ae2) / d);
      float inv = 1.f / (d * d);
      r.lossA = lg * lg;
      r.lossB = ax2 * inv;
      float c = 2.f * lg / ae2;
      r.gA[0] = c * e0; r.gA[1] = c * e1;
      r.gB[0] = 2.f * y[0] * inv; r.gB[1] = 2.f * y[1] * inv;
    } break;
    default: break;
  }
  return r;
}


// Entry barrier of the peer-memory gradient exchange (see PeerArgs): CTA 0 publishes `epoch` to every rank, every CTA
// waits until all ranks have published it.  Bounded wait (PeerArgs::timeout_ns, 120 s unless INR_PEER_TIMEOUT_S says
// otherwise: long enough for a straggler that validates or saves a checkpoint) -> trap instead of a hung GPU.
__device__ __forceinline__ void peer_barrier(const PeerArgs& P, unsigned int epoch) {
  if (static_cast<int>(threadIdx.x) < P.n_ranks) {
    if (blockIdx.x == 0) { __threadfence_system(); st_release_sys(P.flags[threadIdx.x] + P.rank, epoch); }
    const unsigned int* mine = P.flags[P.rank] + threadIdx.x;
    const uint64_t t0 = global_ns();
    unsigned int spins = 0;
    while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 0x3FF) == 0 && global_ns() - t0 > (P.timeout_ns ? P.timeout_ns : 120000000000ull)) { asm volatile("trap;"); }
    }
  }
  __syncthreads();
}
// Exit of a peer-mode optimiser kernel: the last CTA to get here publishes the new step count (all CTAs worked with
// t = old + 1, which they read before any CTA could have stored it: the store happens after EVERY CTA has arrived).
__device__ __forceinline__ void peer_finish_step(const PeerArgs& P, const int* step, int t) {
  __syncthreads();
  if (threadIdx.x == 0 && P.done) {
    __threadfence();
    if (atomicAdd(P.done, 1u) == gridDim.x - 1) { *const_cast<int*>(step) = t; *P.done = 0u; __threadfence(); }
  }
}
// Gather phase: the CTA owns parameters [base, base + 4 * blockDim.x); thread t pulls one float4 from every rank (all
// loads in flight together: the exchange is NVLink-latency bound), adds them in rank order and leaves the mean in
// shared memory for the per-parameter optimiser code that follows.  `n` is a multiple of 4 (buffers are padded).
__device__ __forceinline__ float4 ld_relaxed_sys_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_gather(const PeerArgs& P, size_t base, size_t n, float* s_g /*[4 * blockDim.x]*/) {
  const size_t p4 = base + 4 * static_cast<size_t>(threadIdx.x);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p4 < n) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q)
      if (q < P.n_ranks) v[q] = ld_relaxed_sys_f32x4(P.grads[q] + p4);
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q)
      if (q < P.n_ranks) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
    const float inv = 1.f / static_cast<float>(P.n_ranks);
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  }
  reinterpret_cast<float4*>(s_g)[threadIdx.x] = acc;
  __syncthreads();
}

// Fixed-order block reduction of the tile partials -> step scalars sc[kScalars] in shared memory (all threads of the
// block must call this; works for any blockDim.x that is a multiple of 32 from 64 up to 1024).  Bit-reproducible.
//   loss value, masked row count m, HDR filter mean, normalisers cA / cB (training-loop weights of
//   src/train.py:178-182 folded in), power-of-two gradient scale S = 2^floor(log2(256 / amax)), Adam bias corrections.
__device__ inline void reduce_step_scalars(const float* part_g, int n_tiles, const LossDesc& loss, int out_f, int bs_k,
                                           const float* hyper, const int* step, float* scal_global, float* sc) {
  __shared__ float part[32][6];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, m4 = 0.f, m5 = 0.f;
  for (int t = tid; t < n_tiles; t += blockDim.x) {
    const float* q = part_g + static_cast<size_t>(t) * kPartialsPerTile;
    s0 += q[0]; s1 += q[1]; s2 += q[2]; s3 += q[3];
    m4 = fmaxf(m4, q[4]); m5 = fmaxf(m5, q[5]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off); s3 += __shfl_xor_sync(0xffffffffu, s3, off);
    m4 = fmaxf(m4, __shfl_xor_sync(0xffffffffu, m4, off)); m5 = fmaxf(m5, __shfl_xor_sync(0xffffffffu, m5, off));
  }
  if (lane == 0) { part[warp][0] = s0; part[warp][1] = s1; part[warp][2] = s2; part[warp][3] = s3; part[warp][4] = m4; part[warp][5] = m5; }
  if (tid == 32) {
    // Adam bias corrections (fp64 pow, a microsecond on this machine): nobody in this kernel needs them, so a thread of
    // its own computes them next to the reduction instead of at the end of thread 0's serial path (blockDim.x >= 64)
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (scal_global && hyper && step) {     // torch.optim.Adam: step_size = lr / (1 - b1^t), denom uses sqrt(1 - b2^t)
      const double t = static_cast<double>(*step);
      step_size = static_cast<float>(static_cast<double>(hyper[0]) / (1.0 - pow(static_cast<double>(hyper[1]), t)));
      bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(hyper[2]), t)));
    }
    sc[SC_STEP_SIZE] = step_size; sc[SC_BC2_SQRT] = bc2_sqrt;
    if (scal_global) { scal_global[SC_STEP_SIZE] = step_size; scal_global[SC_BC2_SQRT] = bc2_sqrt; }
  }
  __syncthreads();
  if (tid == 0) {
    float lA = 0.f, lB = 0.f, fs = 0.f, cnt = 0.f, amA = 0.f, amB = 0.f;
    for (int w = 0; w < nwarps; ++w) {
      lA += part[w][0]; lB += part[w][1]; fs += part[w][2]; cnt += part[w][3];
      amA = fmaxf(amA, part[w][4]); amB = fmaxf(amB, part[w][5]);
    }
    float m = fmaxf(cnt, 1.f);
    const float of = static_cast<float>(out_f);
    float cA = 0.f, cB = 0.f, lossv = 0.f, fmean = 0.f, reg = 0.f;
    const float* dpn = nullptr;       // data-parallel shard: normalisers of the GLOBAL batch (see LossDesc)
    if (loss.dp_norm) {
      dpn = loss.dp_norm + 2 * static_cast<size_t>(loss.dp_cursor && loss.dp_rows > 0 ? *loss.dp_cursor / loss.dp_rows : 0);
      m = dpn[0] > 0.f ? dpn[0] : 1.f;
    }
    switch (loss.kind) {
      case LOSS_L2:   cA = 1.f / (m * of);  lossv = lA * 0.5f / (m * of); break;
      case LOSS_L1:   cA = 0.5f / (m * of); lossv = lA * 0.5f / (m * of); break;
      case LOSS_MSLE: cA = 1.f / (m * of);  lossv = lA * 0.5f / (m * of); break;
      case LOSS_TANH: cA = 2.f / (m * of);  lossv = lA / (m * of); break;
      case LOSS_LSL:  cA = 1.f / m;         lossv = lA * 0.5f / m; break;
      case LOSS_HDR:
        fmean = dpn ? dpn[1] : fs / static_cast<float>(bs_k > 0 ? bs_k : 1);
        cA = 1.f / m; cB = loss.factor * fmean / m;
        reg = loss.factor * fmean * lB / m;
        lossv = lA / m + reg;
        break;
      default: cA = 1.f; break;   // external dout: gA holds dL/dz_last already
    }
    if (loss.tv_weight > 0.f && loss.kind != LOSS_HDR) { cB = 1.f; lossv += lB; }   // TV pieces arrive fully normalised
    const float amax = cA * amA + fabsf(cB) * amB;
    float S = 1.f;
    if (amax > 0.f && isfinite(amax)) {
      int e = static_cast<int>(floorf(log2f(256.f / amax)));
      e = e < -60 ? -60 : (e > 60 ? 60 : e);
      S = exp2f(static_cast<float>(e));
    }
    sc[SC_LOSS] = lossv; sc[SC_SCALE] = S; sc[SC_CA] = cA; sc[SC_CB] = cB; sc[SC_COUNT] = cnt; sc[SC_FMEAN] = fmean;
    sc[SC_REG] = reg; sc[SC_INV_SCALE] = 1.f / S;
    if (scal_global)
      for (int i = 0; i < 8; ++i) scal_global[i] = sc[i];
  }
  __syncthreads();
}

}  // namespace inr
