// Descriptors of the multiplicative-filter-network path (FourierNet, MultiscaleKFourier, MultiscaleBoundedFourier;
// reference src/models/mfn.py).  z_0 = sin(Om_0 x + phi_0);  z_i = sin(Om_i x + phi_i) * (W_i z_{i-1} + b_i), i = 1..L,
// heads y_k = V_k z_{s_k} + c_k at stages s_k.  Every stage is one lgemm launch with two GEMM segments.
//
// Images per 128-row tile, `width` features, fp16, layout elem(r,f) at (f/8)*2048 + r*16 + (f%8)*2:
//   X (encoded input), Z[i], G[i] = sin p_i, CP[i] = cos p_i (i = 0..L), H[i] = W_i z_{i-1} + b_i (i = 1..L),
//   DH[i] = S_i dL/dh_i (i = 1..L), DP[i] = S_i dL/dp_i (i = 0..L), DOUT[k] = S dL/dy_k padded to 16 columns.
#pragma once
#include <cstdint>
#include "inr_kernels.cuh"

namespace inr {

constexpr int kMfnMaxStages = 10;     // L + 1 <= 10
constexpr int kMfnNT = 128;

struct MfnModel {
  int L;                 // linear layers (reference network_depth); stages 0..L
  int width, in_f, out_f;
  int n_heads;           // FourierNet: 1 (output_linear at stage L); multiscale: L + 1 (output_linear.i), few are live
  int head_stage[kMfnMaxStages];      // stage of head k
  int head_live[kMfnMaxStages];       // 1: head k is returned by forward (reference output_layers)
  int stage_head[kMfnMaxStages];      // live head index attached to stage i, or -1
  int top;               // highest stage with a live head: stages above it are dead (reference computes and discards them)
  int n_out;             // number of live heads (columns of `out` = n_out * out_f, in stage order)
  int bounded;           // MultiscaleBoundedFourier: rows with dist outside [lo, hi] are zeroed before linear i
  float bound_lo[kMfnMaxStages], bound_hi[kMfnMaxStages];   // indexed by stage (stage i uses reference linear[i-1])
  int input_kind;        // INPUT_GAUSS / INPUT_DENSE / INPUT_LOGF
  int enc_n;             // LogF: frequencies per coordinate (the X image holds 6 n features, zero-padded to in_f)
  int enc_size;
  // float offsets in the flat parameter buffer
  int lin_w[kMfnMaxStages], lin_b[kMfnMaxStages];       // stage i = 1..L
  int head_w[kMfnMaxStages], head_b[kMfnMaxStages];
  int filt_w[kMfnMaxStages], filt_b[kMfnMaxStages];     // stage i = 0..L
  int n_params;
  // "wide chain": SIREN / FFN at widths the on-chip chain kernels are not built for (reference networks.py:48-124 takes any
  // width; 8 shipped configs use 512) run on the same stage GEMMs as plain layers  z_s = act(w0 (W_s z_{s-1} + b_s)):
  // stage 0 reads the X image through filt_w[0] / filt_b[0], stage s >= 1 reads Z[s-1] through lin_w[s] / lin_b[s]; there
  // are no filters above stage 0, CP[s] holds d act / d(pre-activation) = w0 cos(p) (or the ReLU mask), DP[s] the gradient
  // with respect to the pre-activation W z + b, and the single head applies `last_act`.
  int chain;             // 1: wide chain
  int act;               // chain: ACT_SIN / ACT_RELU
  int last_act;          // chain: LAST_LINEAR / LAST_TANH / LAST_SIGMOID / LAST_SIN (sine output layer, network_last_linear False)
  float w0;              // chain: 30 for SIREN, 1 for FFN
  // Gabor filters (GaborNet / KGaborNet, reference mfn.py:96-204): per stage mu [width, in_f] and gamma [width] precede
  // linear.{weight,bias} in the state_dict.  Their gradients come from three split-K reductions of q = dL/df * f:
  //   Qx = q^T x (written at mu's own offset), s = sum_rows q, u = sum_rows q |x|^2 (aux block [width][16], cols 0 / 1)
  //   d mu_j = gamma_j (Qx_j - s_j mu_j),   d gamma_j = -1/2 (u_j + |mu_j|^2 s_j - 2 mu_j . Qx_j)
  int gabor;
  int mu_off[kMfnMaxStages], gamma_off[kMfnMaxStages];
  int aux_off[kMfnMaxStages];        // float offset of the stage's aux block inside one gpart split copy (>= n_params)
  int gfin_mu[kMfnMaxStages], gfin_gamma[kMfnMaxStages];   // float offsets in the finalised-gradient buffer
  int g_floats;                      // floats of one gpart split copy: n_params (+ aux blocks)
  int gfin_floats;
  // packed fp16 operands (bytes in wpack)
  uint32_t pk_filt[kMfnMaxStages], pk_lin[kMfnMaxStages], pk_lin_t[kMfnMaxStages], pk_mu[kMfnMaxStages];
  uint32_t wpack_bytes;
};

struct MfnWorkspace {
  uint64_t x, z[kMfnMaxStages], g[kMfnMaxStages], cp[kMfnMaxStages], h[kMfnMaxStages], dh[kMfnMaxStages], dp[kMfnMaxStages];
  uint64_t dout[kMfnMaxStages];      // per live head
  uint64_t dhu[kMfnMaxStages];       // bounded: dh with NO row mask (bias gradient); dh[] then holds the masked rows
  uint64_t ones;                     // bounded: [128 x 16] fp16 ones, B operand of the bias units
  uint64_t q[kMfnMaxStages];         // Gabor: q_i = S_i dL/df_i * f_i images
  uint64_t e;                        // Gabor: envelope image of the stage being evaluated (transient)
  uint64_t xa;                       // Gabor: per tile [128 x 16] fp16 image, column 0 = 1, column 1 = |x|^2
  uint64_t xn;                       // Gabor: |x|^2 per row, fp32
  uint64_t mn;                       // Gabor: |mu_j|^2 per stage and feature, fp32 [top + 1][width]
  uint64_t gfin;                     // Gabor: finalised (still scaled) d mu / d gamma, fp32
  uint64_t gl, part, scal, gpart, total;
  uint64_t msg, msp, dyf;            // fused multi-head loss: per-head float4 loss pieces, per (tile, head) partials, fp32 dL/dy
  int n_tiles, n_split;
};

struct MfnAuxArgs {
  MfnModel m;
  MfnWorkspace w;
  LossDesc loss;
  const float* params;
  const float* coords; const float* x; const float* encB;
  const float* gt; const uint8_t* mask; const float* dist;
  float* out;
  const float* dout;       // external dL/dout [bs, n_out*out_f] (autograd path) or null (fused single-head loss)
  uint8_t* ws;
  const int* row_offset; int* step_counter;
  const float* hyper; const int* step;
  int bs, train, bs_k;
};

}  // namespace inr
