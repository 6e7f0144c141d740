/* inr_b200.h -- C ABI of the B200-native INR fitting engine (libinr_b200.so).
 *
 * Drop-in boundary for ONE hot path of luisdavid64/MRI-Implicit-Neural-Representations: the per-batch
 * body of the training loops (reference src/train.py:158-192 and src/train_kspace_multiscale.py:164-201):
 * encoder.embedding -> model(x) -> loss -> backward -> Adam.step.  The reference has no native layer
 * (it is pure PyTorch), so each entry point cites the Python it replaces.
 *
 * Conventions
 *  - every function returns 0 on success or a negative INR_E* code; inr_last_error() gives the message;
 *    no C++ exception crosses the boundary;
 *  - all buffers are caller-owned DEVICE pointers (plain pointers + sizes, no torch types); the library
 *    owns only the immutable plan; nothing allocates or synchronises inside a call except
 *    inr_plan_create / inr_selftest_umma;
 *  - every call is asynchronous on the cudaStream_t passed as `void* stream` (0 = default stream);
 *  - one in-flight step per (plan, workspace);
 *  - WIRE / WIRE2D: the layer GEMMs of a step run as chained persistent launches whose CTAs hand tiles to each other
 *    and therefore must all be resident (grid <= SM count).  Steps of DIFFERENT engines must not run concurrently on
 *    one device (issue them on one stream, or on streams that are ordered against each other): two chained launches
 *    sharing the SMs could starve each other's unscheduled CTAs.  A dependency that is not met within ~4 s traps
 *    (CUDA launch error) instead of hanging.
 *
 * Flat parameter buffer: fp32, tensors in reference state_dict order (e.g. model.0.linear.weight [256,512]
 * row-major, model.0.linear.bias [256], ...), see inr_plan_tensor().
 */
#ifndef INR_B200_H
#define INR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define INR_OK 0
#define INR_EINVAL (-1)      /* bad argument / unsupported shape */
#define INR_ECUDA (-2)       /* CUDA runtime error (message has the cudaError string) */
#define INR_EUNSUPPORTED (-3)

/* model kinds (reference src/train.py:55-70) */
#define INR_MODEL_SIREN 1    /* src/models/networks.py:99-124 */
#define INR_MODEL_FFN 2      /* src/models/networks.py:48-69  */
#define INR_MODEL_WIRE 3     /* src/models/networks.py:160-260 (complex Gabor; in 3, hidden int(width/sqrt 2) complex) */
#define INR_MODEL_FOURIER 4  /* FourierNet, src/models/mfn.py:61-94 (one head after stage `depth`) */
#define INR_MODEL_MS_FOURIER 5          /* MultiscaleKFourier, src/models/mfn.py:206-267 (heads at output_layers) */
#define INR_MODEL_MS_BOUNDED_FOURIER 6  /* MultiscaleBoundedFourier, src/models/mfn.py:288-356 (BoundedLinear row masks) */
#define INR_MODEL_WIRE2D 8              /* WIRE2D, src/models/wire2d.py:63-118 (linear + scale_orth per layer, width not reduced, in 3) */
#define INR_MODEL_GABOR 7               /* GaborNet / KGaborNet, src/models/mfn.py:96-204 (mu, gamma, linear per filter; one head) */
/* encoders (src/models/networks.py:7-35) */
#define INR_ENC_NONE 0       /* x is the dense [bs, in] fp32 network input */
#define INR_ENC_GAUSS 1      /* gamma(x) = [sin(2 pi x B^T), cos(2 pi x B^T)] computed in-kernel from coords */
#define INR_ENC_LOGF 2       /* per coordinate c: [sin(2 pi x_c B), cos(2 pi x_c B)], B = 2^linspace(0, scale, n), n = embedding_size / 6
                              * (src/models/networks.py:14-16,24-29), computed in-kernel from coords; encB = B [n] fp32;
                              * network_input_size = 6 n.  SIREN / FFN only (they then run on the streaming layer GEMMs) */
/* last-layer activation */
#define INR_LAST_LINEAR 0
#define INR_LAST_TANH 1      /* SIREN last_tanh, src/models/networks.py:94-95 */
#define INR_LAST_SIGMOID 2   /* FFN, src/models/networks.py:63 */
#define INR_LAST_SIN 3       /* SIREN with network_last_linear False: sine output layer, src/models/networks.py:107-117 */
/* losses as weighted by the training loop (src/train.py:81-98,178-182) */
#define INR_LOSS_NONE 0
#define INR_LOSS_L2 1        /* 0.5 * torch.nn.MSELoss */
#define INR_LOSS_L1 2        /* 0.5 * torch.nn.L1Loss */
#define INR_LOSS_MSLE 3      /* 0.5 * MSLELoss, src/metrics/losses.py:18-27 */
#define INR_LOSS_TANH 4      /* TanhL2Loss, src/metrics/losses.py:121-139 */
#define INR_LOSS_LSL 5       /* 0.5 * LogSpaceLoss, src/metrics/losses.py:204-223 */
#define INR_LOSS_HDR 6       /* HDRLoss_FF, src/metrics/losses.py:226-264 */

typedef struct inr_model_desc {
  int32_t model;            /* INR_MODEL_* */
  int32_t in_features;      /* net.network_input_size */
  int32_t out_features;     /* net.network_output_size */
  int32_t depth;            /* net.network_depth */
  int32_t width;            /* net.network_width */
  int32_t last_act;         /* INR_LAST_* */
  int32_t encoder;          /* INR_ENC_* */
  int32_t enc_size;         /* encoder.embedding_size (gauss) */
  float w0;                 /* SIREN: 30 (hard-wired in the reference, src/models/networks.py:75); WIRE: first_omega_0 */
  float hidden_omega_0;     /* WIRE: net.hidden_omega_0 */
  float sigma0;             /* WIRE: net.scale */
  int32_t head_mask;        /* multiscale MFN: bit i set = output_linear[i] is returned (reference output_layers, default [1,3,5,7]) */
  float bounds[20];         /* MS_BOUNDED_FOURIER: (lo, hi) of BoundedLinear i at [2i], [2i+1] (reference `boundaries`) */
} inr_model_desc;

typedef struct inr_loss_desc {
  int32_t kind;             /* INR_LOSS_* */
  float hdr_eps, hdr_sigma, hdr_factor;   /* loss_opts of src/metrics/losses.py:230-234 */
  /* total-variation term of the per-coil loop (src/train.py:173-174, tv_loss of src/metrics/losses.py:326-343):
   * tv_weight > 0 adds tv_weight * (L1mean(d/dw) + L1mean(d/dh)) of out.view(tv_h, tv_w, out) over ALL rows of the
   * batch (bs must equal tv_h * tv_w, `out` must be given; not combinable with INR_LOSS_HDR in the fused step) */
  float tv_weight;
  int32_t tv_h, tv_w;
  /* data-parallel fits (one process per GPU, every rank a contiguous shard of each global grid-order batch): the
   * reference's loss means run over the WHOLE batch (src/train.py:172-182; HDR: src/metrics/losses.py:241-262), so the
   * normalisers are read from a DEVICE table built from the inputs alone: dp_norm[2 * b] = rows of global batch b that
   * enter the loss / world size, dp_norm[2 * b + 1] = HDR filter mean of global batch b; b = *row_cursor_dev / dp_rows
   * (b = 0 without a cursor).  The mean over ranks of the per-rank gradients is then the global-batch gradient.
   * NULL: single-process semantics (the batch's own count and filter mean). */
  const float* dp_norm;
  int32_t dp_rows;
  /* multi-scale fits (src/train_kspace_multiscale.py:173-190): every head k is supervised with `kind` on the full target
   * and, for k >= 1, pulled towards the detached output of head k-1 on the rows whose dist_to_center lies outside
   * [cons_bounds[2(k-1)], cons_bounds[2(k-1)+1]] -- cons_weight * ConsistencyLoss(bounds), src/metrics/losses.py:292-324
   * (weight 0.1 in the reference loop).  Used by inr_train_step_dist only. */
  float cons_weight;
  float cons_bounds[16];
} inr_loss_desc;

typedef struct inr_tensor_info {
  int64_t offset;           /* float offset inside the flat parameter buffer */
  int32_t rows, cols;       /* weight [rows, cols]; bias: rows = n, cols = 1 */
  int32_t layer, is_bias;
  int32_t is_complex;       /* complex64 tensor: rows*cols interleaved (re, im) float pairs */
  int32_t frozen;           /* requires_grad False in the reference (WIRE omega_0 / scale_0): never updated */
} inr_tensor_info;

/* hyper-parameter block read from DEVICE memory each step (so a captured CUDA graph sees updates):
 * float[8] = { lr, beta1, beta2, eps, weight_decay, reg_l1, reg_l2, unused }
 * (torch.optim.Adam, src/train.py:76; Regularization_L1/L2, src/models/regularization.py:21-36). */
#define INR_HYPER_FLOATS 8
/* step scalars the backward pass leaves at the head of the workspace scalar block, readable with
 * inr_scalars_offset(): loss, grad scale S, cA, cB, masked row count, HDR filter mean, HDR reg term, 1/S */
#define INR_SCALAR_FLOATS 64

typedef struct inr_plan inr_plan;

const char* inr_last_error(void);

/* replaces: model construction dispatch, src/train.py:55-70 */
int inr_plan_create(const inr_model_desc* desc, inr_plan** out);
int inr_plan_destroy(inr_plan* plan);
int inr_plan_param_count(const inr_plan* plan, int64_t* n_params);
int inr_plan_tensor_count(const inr_plan* plan, int32_t* n);
int inr_plan_tensor(const inr_plan* plan, int32_t index, inr_tensor_info* out);
/* bytes of the fp16 operand copies of the weights (caller allocates, 1024-aligned) */
int inr_wpack_bytes(const inr_plan* plan, size_t* bytes);
/* bytes of activation/gradient workspace for batches of up to `bs` coordinates.  The first INR_WS_ZERO_BYTES bytes of
 * the workspace must be ZERO before the first call and belong to the library afterwards: the scalar block
 * (INR_SCALAR_FLOATS floats) carries state across steps (WIRE / MFN per-layer gradient scales and amax accumulators, and
 * the finished-CTA counter of the last-CTA reduction in wire_last_kernel, which resets itself); behind it sit the tile
 * hand-over counters of the WIRE forward layer chain, which every forward pass leaves at zero again (the chain is the first
 * kernel of a step, so no kernel before it could clear them).  The rest of the workspace needs no initialisation. */
#define INR_WS_ZERO_BYTES (1024 + 8 * 4096 * 4)
int inr_workspace_bytes(const inr_plan* plan, int64_t bs, size_t* bytes);
/* byte offset of the INR_SCALAR_FLOATS step scalars inside a workspace sized for `bs` */
int inr_scalars_offset(const inr_plan* plan, int64_t bs, size_t* offset);

/* debugging / tests: byte offsets of the workspace regions for batches of `bs`:
 * out[0..11] = H image of layer l input, out[12..23] = act' images, out[24..35] = dZ images,
 * out[36] = dZ_last, [37] = loss gradient pieces, [38] = tile partials, [39] = scalars, [40] = split-K partials,
 * [41] = n_tiles, [42] = n_split, [43] = total bytes. `n` must be >= 44. */
int inr_workspace_layout(const inr_plan* plan, int64_t bs, uint64_t* out, int32_t n);

/* (re)build the fp16 operand copies from the fp32 parameters: after init / load_state_dict
 * (src/train.py:117-121) or any external edit of the parameters */
int inr_pack_weights(const inr_plan* plan, const float* params, void* wpack, void* stream);

/* replaces: encoder.embedding + model(coords), src/train.py:163-169 / :206-213.
 * input: coords [bs,3] fp32 when the plan's encoder is GAUSS (encB = [enc_size,3] fp32),
 *        x [bs,in_features] fp32 when the encoder is NONE.
 * train != 0 saves the activation images a later inr_backward needs. */
int inr_forward(const inr_plan* plan, const float* params, const void* wpack, const float* input,
                const float* encB, int64_t bs, void* workspace, float* out, int32_t train, void* stream);

/* replaces: loss.backward() for an externally computed dL/dout [bs,out] (arbitrary PyTorch loss on top of
 * the fused model, e.g. CenterLoss), src/train.py:189.  grads = flat fp32 buffer like params. */
int inr_backward(const inr_plan* plan, const float* params, const void* wpack, const float* dout,
                 int64_t bs, void* workspace, float* grads, void* stream);

/* multiscale MFN variants of inr_forward / inr_backward with the per-row distance to the k-space centre
 * (reference MultiscaleBoundedFourier.forward(coords, dist_to_center), src/models/mfn.py:344-356; BoundedLinear :281-286
 * with a 1-D dist: whole rows are zeroed).  `out` / `dout` have n_heads * out_features columns, heads in stage order. */
int inr_forward_dist(const inr_plan* plan, const float* params, const void* wpack, const float* input,
                     const float* encB, const float* dist, int64_t bs, void* workspace, float* out, int32_t train,
                     void* stream);
int inr_backward_dist(const inr_plan* plan, const float* params, const void* wpack, const float* dout,
                      const float* dist, int64_t bs, void* workspace, float* grads, void* stream);

/* replaces: optim.step() (+ regulariser gradient), src/train.py:185-190; re-packs the fp16 copies. */
int inr_adam_step(const inr_plan* plan, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  void* wpack, const float* hyper_dev, const int32_t* step_dev, void* stream);

/* inr_adam_step with the data-parallel gradient exchange fused into the optimiser kernel (replaces the
 * all-reduce a DistributedDataParallel wrapper of src/train.py:189-190 would issue).  peer_grads / peer_flags are HOST
 * arrays of n_ranks DEVICE pointers into symmetric (peer-mapped, NVLink) memory: rank q's reduced fp32 gradients of this
 * step [param_count] (as written by inr_grad_step) and rank q's flag block uint32[64] (zero-initialised, same block every
 * step; words 0..7 = last step each rank published, word 32 = finished-CTA counter).  Like inr_train_step, the kernel advances
 * the device step counter itself: every CTA works with t = *step_dev + 1, publishes t to every rank's flags, waits (bounded,
 * INR_PEER_TIMEOUT_S) until every rank has published it, uses mean_q(peer_grads[q]) in rank order as the gradient, and the
 * last CTA to finish stores t -- no host-enqueued increment sits between the kernels of a step.  The caller alternates two
 * gradient buffers by step parity. */
int inr_adam_step_peers(const inr_plan* plan, float* params, const float* const* peer_grads, uint32_t* const* peer_flags,
                        int32_t n_ranks, int32_t rank, float* exp_avg, float* exp_avg_sq, void* wpack,
                        const float* hyper_dev, int32_t* step_dev, void* stream);

/* replaces: the whole loop body src/train.py:160-192 for fusable model+loss combinations:
 * forward, loss (+ row mask, src/train.py:172-177), backward, Adam, fp16 re-pack.
 *  coords/gt/mask are the RESIDENT arrays of the slice; the batch is rows
 *  [*row_cursor_dev, *row_cursor_dev + bs) when row_cursor_dev != NULL (then advanced by `bs` on the
 *  device at the end of the step, so a captured graph walks the slice in grid order), else [0, bs).
 *  step_dev is incremented on the device before use; loss_out_dev (optional) receives the step's loss. */
int inr_train_step(const inr_plan* plan, const inr_loss_desc* loss, float* params, float* exp_avg,
                   float* exp_avg_sq, void* wpack, const float* hyper_dev, int32_t* step_dev,
                   const float* coords, const float* input_x, const float* encB, const float* gt,
                   const uint8_t* mask, int64_t bs, int32_t* row_cursor_dev, void* workspace,
                   float* out, float* loss_out_dev, void* stream);

/* inr_train_step for the multi-scale models (MultiscaleKFourier / MultiscaleBoundedFourier): replaces the loop body of
 * src/train_kspace_multiscale.py:164-192 -- model(coords, dist_to_center) -> per-head loss + 0.1 * ConsistencyLoss ->
 * backward -> Adam.  dist: fp32 [rows] distance of every row to the k-space centre, indexed like coords / gt (from
 * *row_cursor_dev when a cursor is given).  The row mask applies to the per-head loss, not to the consistency term (:176-183).
 * L2 / L1 / MSLE / LSL, up to 4 heads of 2 outputs, no TV term (those configurations use the autograd face). */
int inr_train_step_dist(const inr_plan* plan, const inr_loss_desc* loss, float* params, float* exp_avg,
                        float* exp_avg_sq, void* wpack, const float* hyper_dev, int32_t* step_dev,
                        const float* coords, const float* input_x, const float* encB, const float* gt,
                        const uint8_t* mask, const float* dist, int64_t bs, int32_t* row_cursor_dev, void* workspace,
                        float* out, float* loss_out_dev, void* stream);

/* inr_train_step_dist without the optimiser (gradients only; see inr_grad_step) */
int inr_grad_step_dist(const inr_plan* plan, const inr_loss_desc* loss, const float* params, const void* wpack,
                       const float* coords, const float* input_x, const float* encB, const float* gt,
                       const uint8_t* mask, const float* dist, int64_t bs, int32_t* row_cursor_dev, void* workspace,
                       float* out, float* grads, float* loss_out_dev, void* stream);

/* inr_train_step without the optimiser: forward + loss + backward, gradients (unscaled fp32, flat, reference
 * parameter order) into `grads`.  For data-parallel training: all-reduce `grads` across ranks, then
 * inr_adam_step.  Replaces src/train.py:160-189 (everything up to optim.step()). */
int inr_grad_step(const inr_plan* plan, const inr_loss_desc* loss, const float* params, const void* wpack,
                  const float* coords, const float* input_x, const float* encB, const float* gt,
                  const uint8_t* mask, int64_t bs, int32_t* row_cursor_dev, void* workspace, float* out,
                  float* grads, float* loss_out_dev, void* stream);

/* measurement only: runs inr_train_step `reps` times with CUDA events between its four kernels and returns the
 * average duration in ms of {forward, dgrad, wgrad, optimiser, forward layer-GEMM launches (WIRE only)} in
 * ms_out4, a host array of at least 5 floats.  WIRE reports backward (dgrad + wgrad) in the third slot.  Synchronises. */
int inr_profile_step(const inr_plan* plan, const inr_loss_desc* loss, float* params, float* exp_avg,
                     float* exp_avg_sq, void* wpack, const float* hyper_dev, int32_t* step_dev,
                     const float* coords, const float* input_x, const float* encB, const float* gt,
                     const uint8_t* mask, int64_t bs, void* workspace, float* out, int32_t reps, float* ms_out4,
                     void* stream);

/* Concurrent fits on ONE device (HP search, src/parameter_search: many small fits per GPU on separate streams).  The chained
 * layer-GEMM launches of WIRE / WIRE2D hand tiles from CTA to CTA and need all their CTAs resident; two such launches of
 * different fits that each want the whole chip can starve each other (bounded wait -> trap).  inr_set_sm_budget(n) caps every
 * chained launch of this process at n SMs: give concurrent fits budgets that sum to at most the SM count (e.g. 74 + 74) and
 * they run side by side.  0 (default) = the whole chip, one fit per device at a time.  Process-global. */
int inr_set_sm_budget(int32_t n_sm);

/* debugging: CTA 0 of the forward kernel writes %globaltimer stamps of its phases into this device buffer of
 * 64 uint64 (NULL switches tracing off; off by default).  Process-global, not thread-safe. */
int inr_debug_set_trace(void* dev_u64_buffer_64);

/* one tcgen05 GEMM per operand-layout family against a host loop (allocates + synchronises; test only) */
int inr_selftest_umma(int mode, int variant, float* max_abs_err, float* ref_absmax);

#ifdef __cplusplus
}
#endif
#endif
