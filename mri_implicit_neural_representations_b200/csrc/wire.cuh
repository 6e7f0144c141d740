// Descriptors of the WIRE (complex Gabor) path.
//
// Complex layers run as real block GEMMs: a complex feature vector h (C features, padded to P = 192) is the real
// row [hr(0..P) | hi(0..P)] (K2 = 2P = 384); z = h W^T + b becomes  a = [hr|hi].[Wr|-Wi]^T,  b = [hr|hi].[Wi|Wr]^T.
// Forward GEMMs use a 3-pass fp16 split (A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation): one fp16 pass
// misses the 1e-3 per-layer bar because the Gabor wavelet multiplies pre-activation errors by ~(omega + 2 sigma^2|z|)
// (hard part 1).  dgrad / wgrad run one fp16 pass on loss-scaled gradients.
//
// Images (same byte layout as the chain kernels: elem(r,f) at (f/8)*2048 + r*16 + (f%8)*2 per 128-row tile):
//   H_hi[l], H_lo[l] : fp16 hi / lo parts of the complex input of layer l (l = 1 .. depth+1), 384 features
//   AB[l]            : fp16 pre-activation (a | b) of layer l (l = 0 .. depth)
//   dZ[l]            : fp16 S * dL/d(a | b) of layer l
#pragma once
#include <cstdint>
#include "inr_kernels.cuh"

namespace inr {

constexpr int kWP = 192;                 // padded complex width
constexpr int kW2 = 2 * kWP;             // real block width (K and N of hidden layers)
constexpr int kWNT = 192;                // accumulator columns per work item: 96 features x (a, b)
constexpr int kWFeatPerBlock = 96;
constexpr int kWTileBytes = kTileM * kW2 * 2;         // 98304: one 128-row image of 384 fp16 features
constexpr int kWStageBBytes = kWNT * kStageK * 2;     // 12288: B stage (192 rows x 32 K)
constexpr int kWStageABytes = kTileM * kStageK * 2;   // 8192:  A stage (128 rows x 32 K)
constexpr int kWMaxDepth = 8;
constexpr int kWFlagTiles = 4096;        // forward hand-over counters live at a fixed workspace offset for up to this many row tiles

// WIRE2D (reference src/models/wire2d.py): every layer has TWO linears (`linear`, `scale_orth`); the hidden width is not
// reduced.  Complex features are padded to P2 (multiple of 64, <= 256); pre-activation / gradient images hold 4 P2 real
// features [a | b | c | d] (a + jb = linear, c + jd = scale_orth), H images 2 P2 = [hr | hi].
constexpr int kW2dMaxP = 256;
constexpr int kW2dNT = 256;              // accumulator columns per item: forward 64 features x (a,b,c,d); dgrad 128 features x (re,im)
constexpr int kW2dFwdFeat = 64, kW2dBwdFeat = 128;

enum LGemmMode { LG_WIRE_FWD = 1, LG_WIRE_DGRAD = 2, LG_MFN_FWD = 3, LG_MFN_DGRAD = 4, LG_GABOR_E = 5, LG_W2D_FWD = 6, LG_W2D_DGRAD = 7 };

// One GEMM segment of a work item: acc[:, acc_col : acc_col + nt] = A[128 x K] * B_block[nt x K]^T
struct LGemmSeg {
  const uint8_t* a_hi;      // A images, tile stride a_tile_bytes = 128 * K * 2
  const uint8_t* a_lo;      // null for 1-pass
  const uint8_t* b_hi;      // packed B, per N-block: [K/32 stages][nt x 32]
  const uint8_t* b_lo;
  uint32_t a_tile_bytes;
  int k_stages;             // K / 32
  int acc_col;              // column offset inside the item's accumulator
};

// WIRE layer chains: the forward (or dgrad) GEMMs of ALL hidden layers run as ONE persistent launch.  Items are ordered
// layer-major and dealt round-robin to the CTAs; item (layer, tile, *) needs the images both N-blocks of (layer - 1, tile)
// wrote, which is tracked per (layer, tile) by a counter in global memory (release by the writing epilogue, acquire by
// the reading producer thread).  A layer's tail no longer idles a third of the SMs: they move on to the next layer.
struct LGemmLayer {
  const uint8_t* a_hi; const uint8_t* a_lo;     // A operand images of this layer (forward: H_hi / H_lo; dgrad: dZ)
  const uint8_t* b_hi; const uint8_t* b_lo;     // packed weights
  const float* bias;                            // forward: complex bias, interleaved
  const float* bias2;                           // WIRE2D forward: complex bias of scale_orth, interleaved
  uint8_t* out_hi; uint8_t* out_lo; uint8_t* out_ab;   // forward outputs (out_lo null: not stored)
  const uint8_t* in_y; const uint8_t* in_ab;    // dgrad epilogue inputs
  uint8_t* out_dz;                              // dgrad output
  float omega;                                  // Gabor constant of the layer whose activation / derivative is evaluated
  int real_first;                               // dgrad: the target layer is the real first layer
  int src_layer, dst_layer;                     // dgrad: scale indices
  // forward, last layer of the chain: the final complex linear (reference networks.py:247-258, real part) rides along --
  // every epilogue thread adds its 24 features' share of out[row][o] = sum_f Re(y_f W[o][f]) into a fixed slot
  const float* last_w;      // final-layer weight [out_f][c] complex, interleaved (re, im), or null
  float* out_part;          // [tile][kWOutParts][128 rows] float4 partial outputs (no atomics), or null
};

struct LGemmArgs {
  LGemmLayer chain[kWMaxDepth];   // WIRE / WIRE2D forward and dgrad only
  int chain_len;                  // layers in the chain (>= 1 for the WIRE modes)
  unsigned int* chain_flags;      // [chain_len][n_tiles] finished N-blocks per (layer, tile), zeroed before the launch
  LGemmSeg seg[2];
  int n_seg;                // 1 (WIRE) or 2 (MFN stage: filter GEMM + linear GEMM)
  int nt;                   // UMMA N of every segment; n_seg * nt <= 256 (two 256-column accumulators ping-pong)
  int n_tiles, n_nblocks, passes, mode;
  // ---- epilogue operands
  const float* bias;        // WIRE_FWD: complex bias, interleaved (re, im); MFN_FWD: linear bias b_i [width]
  const float* phi;         // MFN_FWD: filter bias phi_i [width]
  float act_w0;             // MFN_FWD, != 0: plain layer of a wide SIREN / FFN chain: z = act(acc + w0 * phi) from ONE segment
                            // (B packed with w0 folded in), CP image = d act / d(W z + b) = w0 cos(p) or the ReLU mask
  int act_kind;             // wide chain: ACT_SIN / ACT_RELU
  float omega, sigma;       // WIRE: Gabor constants of the layer whose activation / derivative is evaluated
  int c_valid;              // WIRE: real complex width (181)
  int p2;                   // WIRE2D: padded complex width P2
  const float* bias2;       // WIRE2D_FWD: complex bias of scale_orth, interleaved (re, im)
  int train;                // FWD: also store what backward needs
  int real_first;           // WIRE_DGRAD: target layer is the real first layer;  MFN: target stage is stage 0 (z_0 = sin p_0)
  uint8_t* out_hi;          // WIRE_FWD: H_hi / H_lo of the next layer, AB of this layer.  MFN_FWD: z_i image
  uint8_t* out_lo;          //                                                             MFN_FWD: sin(p_i) image
  uint8_t* out_ab;          //                                                             MFN_FWD: cos(p_i) image
  uint8_t* out_h;           // MFN_FWD: linear output h_i image
  const uint8_t* in_y;      // WIRE_DGRAD: H_hi of the layer's input, AB of the target layer.  MFN_DGRAD: sin(p) of the target stage
  const uint8_t* in_ab;     //                                                                  MFN_DGRAD: cos(p) of the target stage
  const uint8_t* in_h;      // MFN_DGRAD: h of the target stage
  uint8_t* out_dz;          // WIRE_DGRAD: dZ image of the target layer.  MFN_DGRAD: dh image of the target stage
  uint8_t* out_dp;          // MFN_DGRAD: dp image of the target stage
  uint8_t* out_dzu;         // MFN_DGRAD, bounded: dh of the target stage without its row mask (bias gradient) or null
  const float* scal;        // DGRAD: step scalars (per-layer scales at SC_LAYER_SCALE, amax at SC_LAYER_AMAX)
  int src_layer, dst_layer; // DGRAD: A holds S[src] * grad_src, the epilogue stores S[dst] * grad_dst
  // MFN_DGRAD head-gradient injection and BoundedLinear masking
  const float* head_dout;   // [rows, head_ld] fp32 dL/dout (unscaled) or null
  const float* head_w;      // [out_f, width] weight of the head attached to the target stage
  int head_col, head_ld, out_f, bs;
  const float* dist;        // BoundedLinear: per-row distance [rows] or null (indexed from *dist_row_offset when that is given)
  const int* dist_row_offset;
  float bound_lo, bound_hi; // FWD: rows with dist outside [lo, hi] are zeroed before THIS stage's linear;
                            // DGRAD: the same mask of the TARGET stage, applied to the stored dh (W-wgrad and dgrad operand)
  // Gabor filters (reference mfn.py:96-131): envelope E = exp(-gamma/2 (|x|^2 + |mu|^2 - 2 x.mu)) as its own GEMM pass
  const uint8_t* in_e;      // MFN_FWD: envelope image of this stage (null for Fourier filters)
  uint8_t* out_q;           // MFN_DGRAD: q = dL/df * f image of the target stage (drives d mu, d gamma) or null
  const float* xn;          // GABOR_E: |x|^2 per row, fp32
  const float* gamma;       // GABOR_E: gamma_j [width]
  const float* mn;          // GABOR_E: |mu_j|^2 [width]
  uint32_t feat_tile_bytes; // bytes of one 128-row image of the epilogue's feature images (MFN: 128 * width * 2)
  // WIRE_DGRAD chain with the backward of the final complex linear folded in (what wire_blast_kernel does as a launch of its
  // own): top_w != null makes chain[0] a "top" pseudo-layer without a GEMM -- its items' epilogues take dL/dh from
  // dL/dout * W_last instead of an accumulator (out = Re(h W^T + b), reference networks.py:247-258) and also write the
  // padded dz_last image the final layer's wgrad reads.  Uses bs / out_f above.
  const float* top_w;       // final-layer weight [out_f][c] complex, interleaved (re, im)
  const float* top_g;       // per-row loss-gradient pieces [rows][4] fp32 (gA0, gA1, gB0, gB1) of wire_last_kernel
  const float* top_dout;    // external dL/dout [bs][out_f] (autograd face) or null
  uint8_t* top_dzlast;      // dz_last image [tile][2][128 rows][8] fp16
  // WIRE_FWD chain with the real first layer folded in (what wire_first_kernel does as a launch of its own): first_w != null
  // makes chain[0] the first layer 3 -> C (reference networks.py:185-204 with is_first) -- items without a GEMM whose
  // epilogues evaluate z = x W0^T + b0 from the coordinates and write the same images as every other layer (H_hi / H_lo of
  // layer 1, the (a) image, and the coordinate image the first layer's wgrad reads).  Uses bs above; chain[0].bias is the REAL bias.
  const float* first_w;     // first-layer weight [c][3] fp32
  const float* coords;      // [rows][3] fp32, indexed from *row_offset when that is given
  const int* row_offset;
  int* step_counter;        // incremented once per launch (by one thread) or null
  uint8_t* ximg;            // coordinate image [tile][2][128 rows][8] fp16: [x_hi(3), 1, x_lo(3), 0 | 0 x 8]
  int l2_policy;            // set by launch_lgemm: bit 0 hi A images pass as evict_last, bit 1 the dgrad epilogue's (a, b) loads as evict_first
  int dbg;                  // debug (INR_LGEMM_DBG), timing experiments only, results are wrong: bit 0 skip MMAs, bit 1 skip operand copies,
                            // bit 2 skip the proxy fence and bit 3 the hand-over wait of chained layers
  unsigned long long* trace; // debug: 16 %globaltimer stamps per CTA from slot 64 (null in production)
};

struct WireModel {
  int depth;                // hidden complex layers
  int c;                    // complex width (181)
  int in_f, out_f;          // 3, 2
  int nlin;                 // 1: WIRE, 2: WIRE2D (second linear `scale_orth` per layer)
  int last_tanh;            // WIRE2D: complex tanh after the final linear (reference wire2d.py:106-107), output = Re(tanh(z))
  int P;                    // padded complex width: 192 (WIRE) or a multiple of 64 up to 256 (WIRE2D)
  int v_off[kWMaxDepth + 1], vb_off[kWMaxDepth + 1];          // WIRE2D: scale_orth weight / bias offsets, layers 0..depth
  float omega_first, omega_hidden, sigma;
  // float offsets in the flat parameter buffer (reference state_dict order)
  int omega_off[kWMaxDepth + 1], scale_off[kWMaxDepth + 1];   // frozen scalars
  int w_off[kWMaxDepth + 2], b_off[kWMaxDepth + 2];            // layer 0 real [c,3]; 1..depth complex [c,c]; depth+1 complex [out,c]
  int n_params;
  // packed operand copies (bytes in wpack): forward hi/lo and dgrad per hidden layer
  uint32_t wf_hi[kWMaxDepth + 1], wf_lo[kWMaxDepth + 1], wd_hi[kWMaxDepth + 1];
  uint32_t wpack_bytes;
  // virtual gradient blocks (float offsets inside one split of gpart)
  int gd_hidden[kWMaxDepth + 1];      // [384][384] + bias [384]
  int gd_final, gd_first;             // [16][384] + bias[16];  [256][16]
  int gd_floats;
};

struct WireWorkspace {
  uint64_t hhi[kWMaxDepth + 2], hlo[kWMaxDepth + 2], ab[kWMaxDepth + 1], dz[kWMaxDepth + 1];
  uint64_t dzlast, ximg, outacc, g, part, scal, gpart, total;
  uint64_t flags_fwd, flags_bwd;      // uint32 [kWMaxDepth][n_tiles] each: layer-chain hand-over counters (flags_fwd at a FIXED
                                      // offset with room for kWFlagTiles tiles, whatever the batch: zeroed after use)
  int n_tiles, n_split;
};

constexpr int kWOutParts = 8;   // partial sums of the final linear per row: 2 N-blocks x 4 epilogue column groups

struct WireAuxArgs {
  WireModel m;
  WireWorkspace w;
  LossDesc loss;
  const float* params;
  const float* coords; const float* gt; const uint8_t* mask;
  float* out;
  uint8_t* ws;
  const int* row_offset; int* step_counter;
  const float* hyper; const int* step;
  const float* dout;
  int bs, train, bs_k;
  int fold_scalars;         // wire_last: the last CTA to finish also computes the step scalars (no wire_scalars launch)
};

struct WireAdamArgs {
  WireModel m;
  int n_split;
  float* params; float* mom; float* var; float* grads;
  uint8_t* wpack;
  const float* gpart; const float* scal; const float* hyper; const int* step;
  float* loss_out; int* row_offset; int row_advance;
  int do_adam, scal_has_bc, pack_only;
  int params_stable;        // fused step only: nothing since the step's first kernel wrote params / moments, so the optimiser
                            // kernel may fetch them ahead of its programmatic-dependent wait on wgrad
  PeerArgs peer;            // flat kernel only: n_ranks > 0 -> gradients = mean over ranks of peer.grads[q]
};

}  // namespace inr
