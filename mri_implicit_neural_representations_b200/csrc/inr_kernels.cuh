// Shared host/device descriptors for the fused coordinate-MLP ("chain") training kernels.
//
// Data layout in HBM (all fp16 "images" are exact shared-memory pictures of UMMA operands):
//   activation / gradient image of one 128-row tile, F features:
//       elem(row r, feature f) at byte (f/8)*2048 + r*16 + (f%8)*2          (size 256*F bytes)
//     - read as a K-major A operand (K = features) by forward / dgrad   (LBO 2048, SBO 128)
//     - read as an MN-major operand (K = rows)      by wgrad            (LBO 128,  SBO 2048)
//   packed weight stage (256 outputs x 32 K):  elem(n, kk) at (kk/8)*4096 + n*16 + (kk%8)*2  (16 KB)
#pragma once
#include <cstdint>

namespace inr {

constexpr int kTileM = 128;          // coordinates per tile (UMMA M)
constexpr int kWidth = 256;          // hidden width handled by one UMMA N
constexpr int kStageK = 32;          // K columns per packed weight stage
constexpr int kStageBytes = kWidth * kStageK * 2;   // 16384
constexpr int kChunkCols = 64;       // activation chunk that gates the trailing MMA
constexpr int kChunkBytes = kTileM * kChunkCols * 2;  // 16384
constexpr int kActBytes = kTileM * kWidth * 2;        // 65536
constexpr int kMaxLayers = 12;
constexpr int kMaxOut = 4;
constexpr int kDzLastCols = 16;      // last-layer dZ image is padded to one UMMA N=16
constexpr int kDzLastBytes = kTileM * kDzLastCols * 2;  // 4096
constexpr int kPartialsPerTile = 8;  // lossA, lossB, fsum, count, amaxA, amaxB, -, -
constexpr int kScalars = 64;          // step scalars + per-layer gradient scales + amax accumulators (WIRE)

enum Act { ACT_SIN = 0, ACT_RELU = 1 };
enum LastAct { LAST_LINEAR = 0, LAST_TANH = 1, LAST_SIGMOID = 2, LAST_SIN = 3 };
enum InputKind { INPUT_GAUSS = 0, INPUT_DENSE = 1, INPUT_LOGF = 2 };
enum LossKind { LOSS_NONE = 0, LOSS_L2 = 1, LOSS_L1 = 2, LOSS_MSLE = 3, LOSS_TANH = 4, LOSS_LSL = 5, LOSS_HDR = 6 };
// scalar slots written by the backward prologue (device memory, fp32)
enum Scalar { SC_LOSS = 0, SC_SCALE = 1, SC_CA = 2, SC_CB = 3, SC_COUNT = 4, SC_FMEAN = 5, SC_REG = 6, SC_INV_SCALE = 7,
              SC_STEP_SIZE = 8, SC_BC2_SQRT = 9,
              SC_LAYER_SCALE = 16,   // [16 .. 16+depth]: power-of-two scale of the dZ image of layer l (WIRE)
              SC_LAYER_AMAX = 40,
              SC_HEAD_NORM = 50,     // [50 .. 57]: (cA_k, cB_k) of head k, k < 4 (fused multi-head loss)
              SC_MS_LOSS = 58,       // composite loss of the fused multi-head step

              SC_DONE_COUNT = 63 };  // uint32 count of finished CTAs (last-CTA reduction of wire_last), self-resetting  // [40 .. 40+depth]: amax (float bits, atomicMax) of the scaled dZ of layer l, this step   // Adam bias corrections, computed once per step (fp64) by the backward prologue

struct ChainModel {
  int n_gemm;        // tensor-core layers (reference depth - 1)
  int k0;            // input features of layer 0 (multiple of 64)
  int out_f;         // outputs of the final CUDA-core layer (<= kMaxOut)
  int act;           // Act
  int last_act;      // LastAct
  int input_kind;    // InputKind
  int enc_size;      // E for the gauss encoder (k0 == 2E)
  float w0;          // 30 for SIREN
  int w_off[kMaxLayers];     // float offsets of weight [out,in] in the flat parameter buffer; [n_gemm] = last layer
  int b_off[kMaxLayers];
  uint32_t wf_off[kMaxLayers];   // byte offsets of forward-packed stages in wpack
  uint32_t wd_off[kMaxLayers];   // byte offsets of dgrad-packed stages (layers >= 1)
  int n_params;
  uint32_t wpack_bytes;
};

struct LossDesc {
  int kind;
  float eps, sigma, factor;   // HDR / LSL options
  float tv_weight;            // > 0: total-variation term of reference losses.py:326-343 on the batch viewed as [tv_h, tv_w, out]
  int tv_h, tv_w;
  // data-parallel fits: this rank holds a shard of every GLOBAL grid-order batch; the loss means run over the global batch
  // (reference src/train.py:172-182), so the normalisers come from a table that depends on the inputs only:
  // dp_norm[2 * b] = (rows of global batch b that enter the loss) / world, dp_norm[2 * b + 1] = HDR filter mean of global
  // batch b; b = *dp_cursor / dp_rows (0 without a cursor).  The mean over ranks of the per-rank gradients / loss values
  // is then exactly the global-batch gradient / loss.  Null: the batch's own count / mean (single-process semantics).
  const float* dp_norm;
  int dp_rows;
  const int* dp_cursor;
  // multi-head (multi-scale) fits, reference src/train_kspace_multiscale.py:173-190: every head k gets the per-head loss
  // `kind` on the full target plus cons_weight * ConsistencyLoss (src/metrics/losses.py:292-324): for k >= 1 the rows with
  // dist < cons_lo[k-1] or dist > cons_hi[k-1] pull head k towards the (detached) output of head k-1
  float cons_weight;
  float cons_lo[8], cons_hi[8];
};

struct TvArgs {               // tv_kernel (optim.cu): adds the TV gradient / loss to the per-row loss pieces of one batch
  const float* out;           // [bs, out_f] network output of the batch (all rows, before the row mask)
  float* g;                   // [rows_pad] float4 loss pieces; .zw receive the TV gradient (normaliser cB = 1)
  float* part;                // [n_tiles][kPartialsPerTile] tile partials; slots 1 (loss B) and 5 (amax B) are written
  int bs, h, w, out_f;
  float weight;
};

// Workspace byte offsets for a batch of n_tiles tiles (computed on the host by plan_workspace()).
struct Workspace {
  uint64_t h_off[kMaxLayers];    // H image of the input of layer l, l = 0..n_gemm
  uint64_t d_off[kMaxLayers];    // activation-derivative image of layer l, l = 0..n_gemm-1
  uint64_t dz_off[kMaxLayers];   // dZ image of layer l, l = 0..n_gemm-1
  uint64_t dzlast_off;
  uint64_t g_off;                // [rows_pad][4] fp32: unnormalised dL/dz_last parts A and B
  uint64_t part_off;             // [n_tiles][8] fp32
  uint64_t scal_off;             // [16] fp32
  uint64_t gpart_off;            // [n_split][n_params rounded up to 4] fp32 split-K gradient partials (scaled)
  uint64_t total;
  int n_tiles, n_split;
  // Row-tile geometry of every image: 128 rows with k-group stride 2048 B (chain_fwd.cu / chain_bwd.cu), or the transposed
  // small-batch kernels' NR <= 80 rows with the padded stride NR*16 + 16 (chain_t.cu).  elem(r, f) at (f/8)*lb + r*16 + (f%8)*2.
  int tile_rows, lb;
};

struct FwdArgs {
  ChainModel m;
  Workspace w;
  LossDesc loss;
  const float* params;
  const uint8_t* wpack;
  const float* coords;   // [bs,3]
  const float* x;        // [bs,k0] when input_kind == INPUT_DENSE
  const float* encB;     // [E,3]
  const float* gt;       // [bs,out_f] (training) or null
  const uint8_t* mask;   // [bs] 0/1 or null
  float* out;            // [bs,out_f] or null
  uint8_t* ws;
  const int* row_offset; // optional device scalar added to the batch start row (graph replay)
  int* step_counter;     // optional: incremented once per launch by block 0 (Adam bias correction)
  int bs;
  int train;             // 1: store images + loss pieces
  unsigned long long* trace;   // debug: per-phase %globaltimer stamps of CTA 0 (null in production)
};

struct BwdArgs {
  ChainModel m;
  Workspace w;
  LossDesc loss;
  const float* params;
  const uint8_t* wpack;
  const float* dout;     // optional external dL/dout [bs,out_f] (autograd path); null -> use loss pieces
  uint8_t* ws;
  int bs;
  int bs_k;              // rows of (unmasked) kcoords for HDR's filter mean
  const float* hyper;    // optional (fused step): Adam hyper-parameters, to pre-compute the bias corrections
  const int* step;       // optional: 1-based step count (already incremented by the forward kernel)
  unsigned long long* trace;
  int l2_hints;          // set by launch_chain_bwd: bit 0 the act' images (read here last) leave L2 first
};

struct WgradUnit {
  uint64_t a_off, b_off;        // byte offsets of the operand image families inside ws
  uint32_t a_tile_stride, b_tile_stride, a_sub, b_sub;   // bytes
  uint32_t a_bytes, b_bytes;    // bytes copied per tile
  int n;                        // UMMA N (128 or 16)
  int transposed;               // 1: D[i][o] -> dW[o][i]  (last layer)
  int out_off, out_ld;          // float offset / leading dim of the weight gradient
  int row0, col0;               // sub-block origin inside the weight
  int rows_valid, cols_valid;
  int bias_off;                 // float offset of the bias gradient or -1
  int perm_e;                   // >0: columns are in sin/cos-interleaved order with E = perm_e
  int n_chunks;                 // B chunks of 128 features swept against one resident A sub-image (0/1: single chunk)
};

constexpr int kMaxUnits = 320;      // by-value kernel parameter: 320 x (88 + 2) B < 32 KB
// Static two-class schedule of the (unit, split) items over CTAs: a "heavy" item (three-chunk units of the hidden layers)
// gets a CTA of its own, `group` "light" items (first / last layer, single-chunk units) share one, so every CTA streams
// about the same number of operand bytes.  order[] lists the heavy units first, then the light ones.
struct WgradArgs {
  WgradUnit u[kMaxUnits];
  uint16_t order[kMaxUnits];
  int n_heavy, n_light, group;
  int n_units, n_split, n_tiles, n_params;
  int tile_rows, lb;            // rows per tile / k-group stride of the operand images (0: 128 rows, 2048 B)
  uint8_t* ws;
  uint64_t gpart_off;
  unsigned long long* trace;   // debug: 8 %globaltimer stamps per CTA starting at slot 64 (null in production)
  int l2_hints;                // bit 0: A sub-images (read once) leave L2 first, bit 1: B chunks too
};

// Data-parallel gradient exchange over NVLink peer memory, fused into the optimiser kernel: every rank's
// reduced fp32 gradients live in symmetric memory; the optimiser of rank r waits until all ranks have published step
// `epoch` (one flag per peer, release / acquire at system scope), then sums the n_ranks copies with plain peer loads in
// rank order (bit-identical on every rank) -- no separate all-reduce kernel.  The gradient buffers are double-buffered
// by step parity on the host side, which makes the single barrier per step sufficient.
constexpr int kMaxRanks = 8;
struct PeerArgs {
  const float* grads[kMaxRanks];    // rank q's gradient buffer of this step (peer-mapped), [n_params]
  unsigned int* flags[kMaxRanks];   // rank q's flag array uint32[kMaxRanks] (peer-mapped); flags[q][r] = last epoch r published to q
  int n_ranks, rank;                // n_ranks == 0: no exchange
  unsigned long long timeout_ns;    // entry barrier: trap after this long without the other ranks (0: 120 s)
  // the optimiser kernel itself advances the device step counter (no host-enqueued increment between the kernels of a step):
  // every CTA uses t = *step + 1; the LAST CTA to finish (counter `done`, self-resetting) stores t
  unsigned int* done;
};

struct SegDesc {          // one parameter tensor for the optimiser / packer
  int off, rows, cols;    // flat float offset; weight [rows, cols] or bias (rows = n, cols = 1)
  int layer;              // chain layer index or -1 (not packed)
  int pack_fwd, pack_bwd; // 1: write fp16 copies
  int perm_e;             // forward K permutation (gauss input layer)
  int kpad;               // layout 1: K the packed operand is padded to (0: cols) -- LogF input layer, 6 n features padded to 128s
  uint32_t wf_off, wd_off;
  float fwd_scale, bwd_scale;   // factor folded into the fp16 copies (SIREN: w0, so the accumulators hold w0*z)
  int layout;             // 0: chain stages (256 rows x 32 K); 1: lgemm N-blocks of `nt` rows x 32 K
  int nt;
  int scale_slot;         // >= 0: this tensor's gradient partials carry scal[scale_slot] instead of the global loss scale
  int gfin_off;           // >= 0: the (scaled) gradient is read from gfin[gfin_off + index] instead of the split partials
  int frozen;             // no gradient ever reaches this tensor (dead stage / unused head): skipped like grad None in torch
};
constexpr int kMaxSegs = 64;
struct AdamArgs {
  SegDesc seg[kMaxSegs];
  int n_seg, n_params, n_split, n_tiles;
  int gstride;              // floats between consecutive split copies of gpart (n_params rounded up to 4)
  float* params; float* m; float* v;
  float* grads;             // optional: unscaled fp32 gradients are written here when non-null
  uint8_t* wpack;
  PeerArgs peer;            // n_ranks > 0: gradients = mean over ranks of peer.grads[q] (gpart ignored)
  const float* gfin;        // finalised gradients of the tensors with gfin_off >= 0 (Gabor mu / gamma) or null
  const float* gpart;       // [n_split][n_params] gradient partials (scaled by S), or plain gradients (n_split 1)
  const float* scal;        // step scalars written by the backward prologue, or null (scale 1, no loss)
  const float* hyper;       // device: lr, beta1, beta2, eps, weight_decay, reg_l1, reg_l2
  const int* step;          // device: 1-based step count for bias correction
  float* loss_out;          // optional device scalar: loss of this step
  int* row_offset; int row_advance;   // optional: advance the device-side batch cursor
  int do_adam;              // 0: only reduce partials into grads
  int scal_has_bc;          // 1: scal[SC_STEP_SIZE], scal[SC_BC2_SQRT] are valid for this step
};

}  // namespace inr
