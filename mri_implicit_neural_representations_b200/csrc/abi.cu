// extern "C" boundary of libinr_b200.so (declared in include/inr_b200.h): plan construction, workspace
// layout, and the launch sequences of the fused kernels.  No exceptions leave this file.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <new>
#include <cuda_runtime.h>
#include "../../include/inr_b200.h"
#include "inr_kernels.cuh"

namespace inr {
cudaError_t launch_chain_fwd(const FwdArgs& a, int n_sm, cudaStream_t stream);
cudaError_t launch_chain_bwd(const BwdArgs& a, int n_sm, cudaStream_t stream);
cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t stream);
cudaError_t launch_adam(const AdamArgs& a, cudaStream_t stream);
cudaError_t launch_pack(const AdamArgs& a, cudaStream_t stream);
cudaError_t launch_dout_amax(const float* dout, int bs, int out_f, float* partials, int n_tiles, cudaStream_t stream);
}  // namespace inr

using namespace inr;

struct inr_plan {
  inr_model_desc desc;
  ChainModel model;
  std::vector<inr_tensor_info> tensors;
  std::vector<SegDesc> segs;
  std::vector<WgradUnit> units;   // offsets inside the workspace are filled per call (they depend on bs)
  std::vector<int> unit_layer;    // chain layer of each unit
  int n_sm;
};

static thread_local std::string g_err;
static unsigned long long* g_trace = nullptr;   // debug: device buffer for in-kernel phase stamps (inr_debug_set_trace)
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(INR_ECUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
static uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

extern "C" const char* inr_last_error(void) { return g_err.c_str(); }
extern "C" int inr_debug_set_trace(void* dev_u64_buffer_64) { g_trace = static_cast<unsigned long long*>(dev_u64_buffer_64); return INR_OK; }

extern "C" int inr_plan_create(const inr_model_desc* d, inr_plan** out) {
  if (!d || !out) return fail(INR_EINVAL, "null argument");
  if (d->model != INR_MODEL_SIREN && d->model != INR_MODEL_FFN) return fail(INR_EUNSUPPORTED, "model kind not built yet");
  if (d->width != kWidth) return fail(INR_EUNSUPPORTED, "tensor-core chain kernels are built for network_width 256");
  if (d->depth < 2 || d->depth - 1 > kMaxLayers - 1) return fail(INR_EINVAL, "network_depth out of range");
  if (d->out_features < 1 || d->out_features > kMaxOut) return fail(INR_EUNSUPPORTED, "network_output_size must be 1..4");
  if (d->encoder == INR_ENC_GAUSS) {
    if (d->in_features != 2 * d->enc_size || d->enc_size % 32 != 0)
      return fail(INR_EINVAL, "gauss encoder needs network_input_size == 2*embedding_size, embedding_size % 32 == 0");
  } else if (d->encoder != INR_ENC_NONE) {
    return fail(INR_EUNSUPPORTED, "encoder kind not built yet");
  }
  if (d->encoder == INR_ENC_GAUSS && d->enc_size > 512) return fail(INR_EUNSUPPORTED, "embedding_size above 512 is not staged in shared memory");
  if (d->in_features % 64 != 0 || d->in_features > 2048) return fail(INR_EUNSUPPORTED, "network_input_size must be a multiple of 64");
  inr_plan* p = new (std::nothrow) inr_plan();
  if (!p) return fail(INR_EINVAL, "out of host memory");
  p->desc = *d;
  ChainModel& M = p->model;
  std::memset(&M, 0, sizeof(M));
  M.n_gemm = d->depth - 1;
  M.k0 = d->in_features;
  M.out_f = d->out_features;
  M.act = d->model == INR_MODEL_SIREN ? ACT_SIN : ACT_RELU;
  M.last_act = d->last_act;
  M.input_kind = d->encoder == INR_ENC_GAUSS ? INPUT_GAUSS : INPUT_DENSE;
  M.enc_size = d->enc_size;
  M.w0 = d->w0;
  int off = 0;
  uint32_t woff = 0;
  for (int l = 0; l <= M.n_gemm; ++l) {
    const int rows = l == M.n_gemm ? M.out_f : kWidth;
    const int cols = l == 0 ? M.k0 : kWidth;
    M.w_off[l] = off;
    p->tensors.push_back({off, rows, cols, l, 0});
    SegDesc sw{};
    sw.off = off; sw.rows = rows; sw.cols = cols; sw.layer = l;
    sw.fwd_scale = sw.bwd_scale = (M.act == ACT_SIN) ? M.w0 : 1.f;
    if (l < M.n_gemm) {
      sw.pack_fwd = 1;
      sw.perm_e = (l == 0 && M.input_kind == INPUT_GAUSS) ? M.enc_size : 0;
      sw.wf_off = woff; M.wf_off[l] = woff; woff += static_cast<uint32_t>(rows) * cols * 2;
      if (l >= 1) { sw.pack_bwd = 1; sw.wd_off = woff; M.wd_off[l] = woff; woff += static_cast<uint32_t>(rows) * cols * 2; }
    }
    p->segs.push_back(sw);
    off += rows * cols;
    M.b_off[l] = off;
    p->tensors.push_back({off, rows, 1, l, 1});
    SegDesc sb{};
    sb.off = off; sb.rows = rows; sb.cols = 1; sb.layer = l;
    p->segs.push_back(sb);
    off += rows;
  }
  M.n_params = off;
  M.wpack_bytes = woff;
  // wgrad work units
  for (int l = 0; l < M.n_gemm; ++l) {
    const int K = l == 0 ? M.k0 : kWidth;
    for (int mh = 0; mh < kWidth / 128; ++mh)
      for (int nc = 0; nc < K / 128; ++nc) {
        WgradUnit u{};
        u.a_tile_stride = kActBytes; u.a_sub = mh * 32768; u.a_bytes = 32768;
        u.b_tile_stride = kTileM * K * 2; u.b_sub = nc * 32768; u.b_bytes = 32768;
        u.n = 128; u.transposed = 0;
        u.out_off = M.w_off[l]; u.out_ld = K; u.row0 = mh * 128; u.col0 = nc * 128;
        u.rows_valid = 128; u.cols_valid = 128;
        u.bias_off = nc == 0 ? M.b_off[l] : -1;
        u.perm_e = (l == 0 && M.input_kind == INPUT_GAUSS) ? M.enc_size : 0;
        p->units.push_back(u); p->unit_layer.push_back(l);
      }
  }
  for (int ic = 0; ic < kWidth / 128; ++ic) {
    WgradUnit u{};
    u.a_tile_stride = kActBytes; u.a_sub = ic * 32768; u.a_bytes = 32768;
    u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
    u.n = kDzLastCols; u.transposed = 1;
    u.out_off = M.w_off[M.n_gemm]; u.out_ld = kWidth; u.row0 = 0; u.col0 = ic * 128;
    u.rows_valid = M.out_f; u.cols_valid = 128;
    u.bias_off = ic == 0 ? M.b_off[M.n_gemm] : -1;
    u.perm_e = 0;
    p->units.push_back(u); p->unit_layer.push_back(M.n_gemm);
  }
  if (static_cast<int>(p->units.size()) > kMaxUnits || static_cast<int>(p->segs.size()) > kMaxSegs) {
    delete p;
    return fail(INR_EUNSUPPORTED, "model too deep for the static unit tables");
  }
  p->n_sm = 148;
  int dev = 0, n = 0;
  if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0 && cudaGetDevice(&dev) == cudaSuccess) {
    int sm = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sm > 0) p->n_sm = sm;
  } else {
    cudaGetLastError();
  }
  *out = p;
  return INR_OK;
}

extern "C" int inr_plan_destroy(inr_plan* p) { delete p; return INR_OK; }
extern "C" int inr_plan_param_count(const inr_plan* p, int64_t* n) {
  if (!p || !n) return fail(INR_EINVAL, "null argument");
  *n = p->model.n_params; return INR_OK;
}
extern "C" int inr_plan_tensor_count(const inr_plan* p, int32_t* n) {
  if (!p || !n) return fail(INR_EINVAL, "null argument");
  *n = static_cast<int32_t>(p->tensors.size()); return INR_OK;
}
extern "C" int inr_plan_tensor(const inr_plan* p, int32_t i, inr_tensor_info* out) {
  if (!p || !out || i < 0 || i >= static_cast<int32_t>(p->tensors.size())) return fail(INR_EINVAL, "bad tensor index");
  *out = p->tensors[i]; return INR_OK;
}
extern "C" int inr_wpack_bytes(const inr_plan* p, size_t* b) {
  if (!p || !b) return fail(INR_EINVAL, "null argument");
  *b = p->model.wpack_bytes; return INR_OK;
}

static Workspace plan_workspace(const inr_plan* p, int64_t bs) {
  const ChainModel& M = p->model;
  Workspace w{};
  const int T = static_cast<int>((bs + kTileM - 1) / kTileM);
  w.n_tiles = T;
  int ns = p->n_sm / static_cast<int>(p->units.size());
  if (ns < 1) ns = 1;
  if (ns > T) ns = T > 0 ? T : 1;
  w.n_split = ns;
  uint64_t o = 0;
  w.scal_off = o; o += align_up(kScalars * 4, 1024);
  w.part_off = o; o += align_up(static_cast<uint64_t>(T) * kPartialsPerTile * 4, 1024);
  w.g_off = o; o += align_up(static_cast<uint64_t>(T) * kTileM * 16, 1024);
  for (int l = 0; l <= M.n_gemm; ++l) {
    const int K = l == 0 ? M.k0 : kWidth;
    w.h_off[l] = o; o += static_cast<uint64_t>(T) * kTileM * K * 2;
  }
  for (int l = 0; l < M.n_gemm; ++l) { w.d_off[l] = o; o += static_cast<uint64_t>(T) * kActBytes; }
  for (int l = 0; l < M.n_gemm; ++l) { w.dz_off[l] = o; o += static_cast<uint64_t>(T) * kActBytes; }
  w.dzlast_off = o; o += align_up(static_cast<uint64_t>(T) * kDzLastBytes, 1024);
  w.gpart_off = o; o += align_up(static_cast<uint64_t>(ns) * M.n_params * 4, 1024);
  w.total = o;
  return w;
}

extern "C" int inr_workspace_bytes(const inr_plan* p, int64_t bs, size_t* bytes) {
  if (!p || !bytes || bs <= 0) return fail(INR_EINVAL, "bad argument");
  *bytes = plan_workspace(p, bs).total; return INR_OK;
}
extern "C" int inr_scalars_offset(const inr_plan* p, int64_t bs, size_t* off) {
  if (!p || !off || bs <= 0) return fail(INR_EINVAL, "bad argument");
  *off = plan_workspace(p, bs).scal_off; return INR_OK;
}

extern "C" int inr_workspace_layout(const inr_plan* p, int64_t bs, uint64_t* out, int32_t n) {
  if (!p || !out || bs <= 0 || n < 44) return fail(INR_EINVAL, "bad argument");
  const Workspace w = plan_workspace(p, bs);
  for (int l = 0; l < kMaxLayers; ++l) { out[l] = w.h_off[l]; out[12 + l] = w.d_off[l]; out[24 + l] = w.dz_off[l]; }
  out[36] = w.dzlast_off; out[37] = w.g_off; out[38] = w.part_off; out[39] = w.scal_off; out[40] = w.gpart_off;
  out[41] = static_cast<uint64_t>(w.n_tiles); out[42] = static_cast<uint64_t>(w.n_split); out[43] = w.total;
  return INR_OK;
}

static void fill_adam(const inr_plan* p, AdamArgs& a) {
  std::memset(&a, 0, sizeof(a));
  a.n_seg = static_cast<int>(p->segs.size());
  for (int i = 0; i < a.n_seg; ++i) a.seg[i] = p->segs[i];
  a.n_params = p->model.n_params;
}

static void fill_wgrad(const inr_plan* p, const Workspace& w, uint8_t* ws, WgradArgs& g) {
  const ChainModel& M = p->model;
  std::memset(&g, 0, sizeof(g));
  g.n_units = static_cast<int>(p->units.size());
  for (int i = 0; i < g.n_units; ++i) {
    WgradUnit u = p->units[i];
    const int l = p->unit_layer[i];
    if (l < M.n_gemm) { u.a_off = w.dz_off[l]; u.b_off = w.h_off[l]; }
    else { u.a_off = w.h_off[M.n_gemm]; u.b_off = w.dzlast_off; }
    g.u[i] = u;
  }
  g.n_split = w.n_split; g.n_tiles = w.n_tiles; g.n_params = M.n_params;
  g.ws = ws; g.gpart_off = w.gpart_off;
}

extern "C" int inr_pack_weights(const inr_plan* p, const float* params, void* wpack, void* stream) {
  if (!p || !params || !wpack) return fail(INR_EINVAL, "null argument");
  AdamArgs a; fill_adam(p, a);
  a.params = const_cast<float*>(params); a.wpack = static_cast<uint8_t*>(wpack);
  cudaError_t e = launch_pack(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "pack_kernel");
}

static int run_forward(const inr_plan* p, const Workspace& w, const LossDesc& loss, const float* params, const void* wpack,
                       const float* coords, const float* x, const float* encB, const float* gt, const uint8_t* mask,
                       int64_t bs, void* ws, float* out, int train, const int* row_off, int* step, cudaStream_t st) {
  FwdArgs f{};
  f.m = p->model; f.w = w; f.loss = loss;
  f.params = params; f.wpack = static_cast<const uint8_t*>(wpack);
  f.coords = coords; f.x = x; f.encB = encB; f.gt = gt; f.mask = mask; f.out = out;
  f.ws = static_cast<uint8_t*>(ws); f.row_offset = row_off; f.step_counter = step;
  f.bs = static_cast<int>(bs); f.train = train; f.trace = g_trace;
  cudaError_t e = launch_chain_fwd(f, p->n_sm, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "chain_fwd_kernel");
}

extern "C" int inr_forward(const inr_plan* p, const float* params, const void* wpack, const float* input, const float* encB,
                           int64_t bs, void* workspace, float* out, int32_t train, void* stream) {
  if (!p || !params || !wpack || !input || !out || bs <= 0) return fail(INR_EINVAL, "bad argument");
  if (train && !workspace) return fail(INR_EINVAL, "training forward needs a workspace");
  const bool gauss = p->model.input_kind == INPUT_GAUSS;
  if (gauss && !encB) return fail(INR_EINVAL, "gauss encoder needs encB");
  Workspace w = plan_workspace(p, bs);
  LossDesc none{LOSS_NONE, 0.f, 0.f, 0.f};
  return run_forward(p, w, none, params, wpack, gauss ? input : nullptr, gauss ? nullptr : input, encB, nullptr, nullptr, bs,
                     workspace, out, train ? 1 : 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

static int run_backward(const inr_plan* p, const Workspace& w, const LossDesc& loss, const float* params, const void* wpack,
                        const float* dout, int64_t bs, void* ws, cudaStream_t st, cudaEvent_t mid = nullptr,
                        const float* hyper = nullptr, const int* step = nullptr) {
  BwdArgs b{};
  b.hyper = hyper; b.step = step;
  b.m = p->model; b.w = w; b.loss = loss;
  b.params = params; b.wpack = static_cast<const uint8_t*>(wpack); b.dout = dout;
  b.ws = static_cast<uint8_t*>(ws); b.bs = static_cast<int>(bs); b.bs_k = static_cast<int>(bs);
  cudaError_t e = launch_chain_bwd(b, p->n_sm, st);
  if (e != cudaSuccess) return cuda_fail(e, "chain_bwd_kernel");
  if (mid) cudaEventRecord(mid, st);
  WgradArgs g; fill_wgrad(p, w, static_cast<uint8_t*>(ws), g);
  e = launch_wgrad(g, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "wgrad_kernel");
}

extern "C" int inr_backward(const inr_plan* p, const float* params, const void* wpack, const float* dout, int64_t bs,
                            void* workspace, float* grads, void* stream) {
  if (!p || !params || !wpack || !dout || !workspace || !grads || bs <= 0) return fail(INR_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace w = plan_workspace(p, bs);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  cudaError_t e = launch_dout_amax(dout, static_cast<int>(bs), p->model.out_f, reinterpret_cast<float*>(ws + w.part_off), w.n_tiles, st);
  if (e != cudaSuccess) return cuda_fail(e, "dout_amax_kernel");
  LossDesc none{LOSS_NONE, 0.f, 0.f, 0.f};
  int rc = run_backward(p, w, none, params, wpack, dout, bs, workspace, st);
  if (rc) return rc;
  AdamArgs a; fill_adam(p, a);
  a.n_split = w.n_split; a.n_tiles = w.n_tiles;
  a.params = const_cast<float*>(params); a.grads = grads;
  a.gpart = reinterpret_cast<const float*>(ws + w.gpart_off);
  a.scal = reinterpret_cast<const float*>(ws + w.scal_off);
  a.do_adam = 0;
  e = launch_adam(a, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "adam_kernel(reduce)");
}

extern "C" int inr_adam_step(const inr_plan* p, float* params, const float* grads, float* m, float* v, void* wpack,
                             const float* hyper_dev, const int32_t* step_dev, void* stream) {
  if (!p || !params || !grads || !m || !v || !wpack || !hyper_dev || !step_dev) return fail(INR_EINVAL, "null argument");
  AdamArgs a; fill_adam(p, a);
  a.n_split = 1; a.params = params; a.m = m; a.v = v; a.wpack = static_cast<uint8_t*>(wpack);
  a.gpart = grads; a.scal = nullptr; a.hyper = hyper_dev; a.step = step_dev; a.do_adam = 1;
  cudaError_t e = launch_adam(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "adam_kernel");
}

static int train_step_impl(const inr_plan* p, const inr_loss_desc* loss, float* params, float* m, float* v, void* wpack,
                           const float* hyper_dev, int32_t* step_dev, const float* coords, const float* input_x,
                           const float* encB, const float* gt, const uint8_t* mask, int64_t bs, int32_t* row_cursor_dev,
                           void* workspace, float* out, float* loss_out_dev, cudaStream_t st, cudaEvent_t* ev,
                           float* grads_only = nullptr) {
  const bool no_adam = grads_only != nullptr;
  if (!p || !loss || !params || !wpack || !gt || !workspace || bs <= 0) return fail(INR_EINVAL, "bad argument");
  if (!no_adam && (!m || !v || !hyper_dev || !step_dev)) return fail(INR_EINVAL, "bad argument");
  const bool gauss = p->model.input_kind == INPUT_GAUSS;
  if (gauss && (!coords || !encB)) return fail(INR_EINVAL, "gauss encoder needs coords and encB");
  if (!gauss && !input_x) return fail(INR_EINVAL, "dense input needs input_x");
  if (loss->kind < INR_LOSS_L2 || loss->kind > INR_LOSS_HDR) return fail(INR_EINVAL, "unknown loss kind");
  if ((loss->kind == INR_LOSS_HDR || loss->kind == INR_LOSS_LSL) && p->model.out_f != 2)
    return fail(INR_EINVAL, "complex-valued losses need network_output_size == 2");
  if (loss->kind == INR_LOSS_HDR && !coords) return fail(INR_EINVAL, "HDR loss needs kcoords");
  if (p->model.out_f > 2) return fail(INR_EUNSUPPORTED, "fused loss path packs at most 2 outputs");
  Workspace w = plan_workspace(p, bs);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  LossDesc L{loss->kind, loss->hdr_eps, loss->hdr_sigma, loss->hdr_factor};
  if (ev) cudaEventRecord(ev[0], st);
  int rc = run_forward(p, w, L, params, wpack, coords, input_x, encB, gt, mask, bs, workspace, out, 1, row_cursor_dev,
                       no_adam ? nullptr : step_dev, st);
  if (rc) return rc;
  if (ev) cudaEventRecord(ev[1], st);
  rc = run_backward(p, w, L, params, wpack, nullptr, bs, workspace, st, ev ? ev[2] : nullptr,
                    no_adam ? nullptr : hyper_dev, no_adam ? nullptr : step_dev);
  if (rc) return rc;
  if (ev) cudaEventRecord(ev[3], st);
  AdamArgs a; fill_adam(p, a);
  a.n_split = w.n_split; a.n_tiles = w.n_tiles;
  a.params = params; a.m = m; a.v = v; a.wpack = static_cast<uint8_t*>(wpack);
  a.gpart = reinterpret_cast<const float*>(ws + w.gpart_off);
  a.scal = reinterpret_cast<const float*>(ws + w.scal_off);
  a.hyper = hyper_dev; a.step = step_dev; a.loss_out = loss_out_dev;
  a.row_offset = row_cursor_dev; a.row_advance = static_cast<int>(bs);
  a.do_adam = no_adam ? 0 : 1; a.scal_has_bc = no_adam ? 0 : 1; a.grads = grads_only;
  cudaError_t e = launch_adam(a, st);
  if (ev) cudaEventRecord(ev[4], st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "adam_kernel");
}

extern "C" int inr_train_step(const inr_plan* p, const inr_loss_desc* loss, float* params, float* m, float* v, void* wpack,
                              const float* hyper_dev, int32_t* step_dev, const float* coords, const float* input_x,
                              const float* encB, const float* gt, const uint8_t* mask, int64_t bs, int32_t* row_cursor_dev,
                              void* workspace, float* out, float* loss_out_dev, void* stream) {
  return train_step_impl(p, loss, params, m, v, wpack, hyper_dev, step_dev, coords, input_x, encB, gt, mask, bs,
                         row_cursor_dev, workspace, out, loss_out_dev, static_cast<cudaStream_t>(stream), nullptr);
}

extern "C" int inr_grad_step(const inr_plan* p, const inr_loss_desc* loss, const float* params, const void* wpack,
                             const float* coords, const float* input_x, const float* encB, const float* gt,
                             const uint8_t* mask, int64_t bs, int32_t* row_cursor_dev, void* workspace, float* out,
                             float* grads, float* loss_out_dev, void* stream) {
  if (!grads) return fail(INR_EINVAL, "grads must not be null");
  return train_step_impl(p, loss, const_cast<float*>(params), nullptr, nullptr, const_cast<void*>(wpack), nullptr, nullptr,
                         coords, input_x, encB, gt, mask, bs, row_cursor_dev, workspace, out, loss_out_dev,
                         static_cast<cudaStream_t>(stream), nullptr, grads);
}

extern "C" int inr_profile_step(const inr_plan* p, const inr_loss_desc* loss, float* params, float* m, float* v, void* wpack,
                                const float* hyper_dev, int32_t* step_dev, const float* coords, const float* input_x,
                                const float* encB, const float* gt, const uint8_t* mask, int64_t bs, void* workspace,
                                int32_t reps, float* ms_out4, void* stream) {
  if (!ms_out4 || reps <= 0) return fail(INR_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t ev[5];
  for (auto& e : ev) if (cudaEventCreate(&e) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaEventCreate");
  double acc[4] = {0, 0, 0, 0};
  int rc = INR_OK;
  for (int r = 0; r < reps && rc == INR_OK; ++r) {
    rc = train_step_impl(p, loss, params, m, v, wpack, hyper_dev, step_dev, coords, input_x, encB, gt, mask, bs, nullptr,
                         workspace, nullptr, nullptr, st, ev);
    if (rc) break;
    cudaError_t e = cudaEventSynchronize(ev[4]);
    if (e != cudaSuccess) { rc = cuda_fail(e, "profile step"); break; }
    for (int k = 0; k < 4; ++k) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); acc[k] += ms; }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int k = 0; k < 4; ++k) ms_out4[k] = static_cast<float>(acc[k] / reps);
  return rc;
}
