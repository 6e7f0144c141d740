// extern "C" boundary of libinr_b200.so (declared in include/inr_b200.h): plan construction, workspace
// layout, and the launch sequences of the fused kernels.  No exceptions leave this file.
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <new>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/inr_b200.h"
#include "inr_kernels.cuh"
#include "wire.cuh"
#include "mfn.cuh"

namespace inr {
cudaError_t launch_chain_fwd(const FwdArgs& a, int n_sm, cudaStream_t stream);
cudaError_t launch_chain_bwd(const BwdArgs& a, int n_sm, cudaStream_t stream);
cudaError_t launch_chain_fwd_t(const FwdArgs& a, int n_sm, cudaStream_t stream);
cudaError_t launch_chain_bwd_t(const BwdArgs& a, int n_sm, cudaStream_t stream);
cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t stream);
cudaError_t launch_adam(const AdamArgs& a, cudaStream_t stream);
cudaError_t launch_pack(const AdamArgs& a, cudaStream_t stream);
cudaError_t launch_dout_amax(const float* dout, int bs, int out_f, float* partials, int n_tiles, cudaStream_t stream);
cudaError_t launch_lgemm(const LGemmArgs& a, int n_sm, cudaStream_t stream);
void lgemm_set_sm_budget(int n);
int lgemm_sm_budget();
cudaError_t launch_wire_first(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_wire_last(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_wire_scalars(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_wire_dout_amax(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_wire_blast(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_wire_adam(const WireAdamArgs& a, cudaStream_t st);
cudaError_t launch_wire_adam_flat(const WireAdamArgs& a, cudaStream_t st);
cudaError_t launch_mfn_encode(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_head(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_dout_amax(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_scalars(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_top(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_ms_loss(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_gabor_prep(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_mfn_gabor_grad(const MfnAuxArgs& a, cudaStream_t st);
cudaError_t launch_tv(const TvArgs& a, int n_tiles, cudaStream_t st);
cudaError_t launch_w2d_first(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_w2d_last(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_w2d_blast(const WireAuxArgs& a, cudaStream_t st);
cudaError_t launch_w2d_adam(const WireAdamArgs& a, cudaStream_t st);
cudaError_t launch_w2d_adam_flat(const WireAdamArgs& a, cudaStream_t st);
}  // namespace inr

using namespace inr;

struct inr_plan {
  inr_model_desc desc;
  bool is_wire = false;
  WireModel wm;
  bool is_mfn = false;
  MfnModel mm;
  ChainModel model;
  std::vector<inr_tensor_info> tensors;
  std::vector<SegDesc> segs;
  std::vector<WgradUnit> units;   // offsets inside the workspace are filled per call (they depend on bs)
  std::vector<int> unit_layer;    // chain layer of each unit
  int n_sm;
  // wgrad schedule (WgradArgs): heavy units first in `order`, `group` light (unit, split) items per CTA
  std::vector<uint16_t> order;
  int n_heavy = 0, n_light = 0, group = 1;
};

static thread_local std::string g_err;

// NVTX ranges around the kernel groups of a step (header-only nvtx3: a no-op unless a profiler injects itself); INR_NVTX=0
// switches them off
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("INR_NVTX"); v = (e && e[0] == '0') ? 0 : 1; }
    on = v != 0;
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};
// wgrad units sweep up to three 128-feature B chunks per resident A sub-image: split `n` chunks into equal groups
static int chunk_group(int n) { const int groups = (n + 2) / 3; return (n + groups - 1) / groups; }

// What one row tile costs a wgrad unit, in operand bytes (what bounds the kernel): its A sub-image plus every B chunk --
// but never less than 48 KB: a unit with tiny operands is bound by the ring's round trip (about 2.2 us over the three
// 32 KB slots its A sub-images get), not by its bytes.  With this floor two light items share a CTA on every model
// measured (WIRE: 54.5 -> 47 us against 57 us with three per CTA; SIREN: 20.7 us either way).
static uint32_t unit_cost(const WgradUnit& u) {
  return std::max<uint32_t>(49152u, u.a_bytes + static_cast<uint32_t>(u.n_chunks > 1 ? u.n_chunks : 1) * u.b_bytes);
}

// Two-class static schedule: units costing more than half of the dearest one are "heavy" (one CTA per (unit, split)
// item), the others are "light" and `group` of their items share a CTA, group = dearest / dearest light cost.
static void build_wgrad_sched(inr_plan* p) {
  uint32_t cmax = 0, lmax = 0;
  for (const WgradUnit& u : p->units) cmax = std::max(cmax, unit_cost(u));
  p->order.clear();
  std::vector<uint16_t> light;
  for (size_t i = 0; i < p->units.size(); ++i) {
    const uint32_t c = unit_cost(p->units[i]);
    if (2 * c > cmax) p->order.push_back(static_cast<uint16_t>(i));
    else { light.push_back(static_cast<uint16_t>(i)); lmax = std::max(lmax, c); }
  }
  p->n_heavy = static_cast<int>(p->order.size());
  p->n_light = static_cast<int>(light.size());
  p->group = lmax ? std::min<int>(8, std::max<int>(1, static_cast<int>(cmax / lmax))) : 1;
  // More units than SMs (the MFN models: split factor 1, several waves of CTAs): the hardware's CTA scheduler already
  // balances the waves, and a CTA that works off several whole-batch items in turn would only lengthen the tail
  // (GaborNet config 3: 5.42 -> 5.90 ms per step with grouping).  Group only when everything fits one wave.
  if (static_cast<int>(p->units.size()) > p->n_sm) p->group = 1;
  p->order.insert(p->order.end(), light.begin(), light.end());
}

// split-K factor: as many row-tile ranges as fit one wave of CTAs (heavy items 1 per CTA, light items `group` per CTA)
static int wgrad_splits(const inr_plan* p, int T) {
  int ns = 1;
  while (p->n_heavy * (ns + 1) + (p->n_light * (ns + 1) + p->group - 1) / p->group <= p->n_sm) ++ns;
  if (ns > T) ns = T > 0 ? T : 1;
  return ns;
}

static void fill_wgrad_sched(const inr_plan* p, WgradArgs& g) {
  g.n_heavy = p->n_heavy; g.n_light = p->n_light; g.group = p->group;
  for (size_t i = 0; i < p->order.size(); ++i) g.order[i] = p->order[i];
}

// float stride between split-K partial copies: a multiple of 4 so every row the wgrad epilogue bulk-stores stays 16 B aligned
static int gpart_stride(int n_params) { return (n_params + 3) & ~3; }

static unsigned long long* g_trace = nullptr;   // debug: device buffer for in-kernel phase stamps (inr_debug_set_trace)
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(INR_ECUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
static uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

extern "C" const char* inr_last_error(void) { return g_err.c_str(); }
extern "C" int inr_set_sm_budget(int32_t n_sm) {
  if (n_sm < 0) return fail(INR_EINVAL, "SM budget must be >= 0 (0 = whole chip)");
  inr::lgemm_set_sm_budget(n_sm);
  return INR_OK;
}
static int g_trace_lgemm_sel = -1, g_trace_lgemm_count = 0;   // debug: which layer-GEMM launch after set_trace gets the buffer
extern "C" int inr_debug_set_trace(void* dev_u64_buffer_64) {
  g_trace = static_cast<unsigned long long*>(dev_u64_buffer_64);
  const char* sel = std::getenv("INR_TRACE_LGEMM");
  g_trace_lgemm_sel = sel ? std::atoi(sel) : -1;
  g_trace_lgemm_count = 0;
  return INR_OK;
}
// TV term between the forward (which left the main loss pieces and `out`) and the backward (which reduces the partials)
static int run_tv(const inr_loss_desc* loss, const float* out, int out_f, int64_t bs, uint8_t* ws, uint64_t g_off, uint64_t part_off,
                  int n_tiles, cudaStream_t st) {
  if (!(loss->tv_weight > 0.f)) return INR_OK;
  if (loss->kind == INR_LOSS_HDR) return fail(INR_EUNSUPPORTED, "TV + HDR is not fused (use the autograd face)");
  if (!out) return fail(INR_EINVAL, "the TV term needs the `out` buffer");
  if (loss->tv_h < 1 || loss->tv_w < 1 || static_cast<int64_t>(loss->tv_h) * loss->tv_w != bs)
    return fail(INR_EINVAL, "TV needs bs == tv_h * tv_w (one coil per batch, src/train.py:173-174)");
  TvArgs t{out, reinterpret_cast<float*>(ws + g_off), reinterpret_cast<float*>(ws + part_off), static_cast<int>(bs), loss->tv_h,
           loss->tv_w, out_f, loss->tv_weight};
  cudaError_t e = launch_tv(t, n_tiles, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "tv_kernel");
}

// read per launch (tools/ab_variants.py switches it between launches of one process)
// (default off: measured on B200 at bs 25 000, the chain grows by 9.6 us -- 2.65 rounds of store-bound items during which the
// tensor pipe idles -- while wire_first_kernel takes 13 us next to it: forward phase 117.2 -> 113.3 us with eager launches, but
// 228.1 -> 230.4 us per step in the CUDA graph, where the separate kernel's launch is already hidden.  INR_WIRE_FOLD_FIRST=1.)
static bool wire_fold_first() { const char* e = std::getenv("INR_WIRE_FOLD_FIRST"); return e && e[0] == '1'; }
static bool wire_fold_blast() { const char* e = std::getenv("INR_WIRE_FOLD_BLAST"); return !(e && e[0] == '0'); }
static int lgemm_dbg() { const char* e = std::getenv("INR_LGEMM_DBG"); return e ? std::atoi(e) : 0; }     // timing experiments; read per launch
static unsigned long long* lgemm_trace_ptr() { return (g_trace && g_trace_lgemm_count++ == g_trace_lgemm_sel) ? g_trace : nullptr; }

static int wire_plan_create(const inr_model_desc* d, inr_plan** out);
static int wire2d_plan_create(const inr_model_desc* d, inr_plan** out);
static WireWorkspace wire_workspace(const inr_plan* p, int64_t bs);
static int mfn_plan_create(const inr_model_desc* d, inr_plan** out);
static int wide_plan_create(const inr_model_desc* d, inr_plan** out);
static MfnWorkspace mfn_workspace(const inr_plan* p, int64_t bs);

extern "C" int inr_plan_create(const inr_model_desc* d, inr_plan** out) {
  if (!d || !out) return fail(INR_EINVAL, "null argument");
  if (d->model == INR_MODEL_WIRE) return wire_plan_create(d, out);
  if (d->model == INR_MODEL_WIRE2D) return wire2d_plan_create(d, out);
  if (d->model == INR_MODEL_FOURIER || d->model == INR_MODEL_MS_FOURIER || d->model == INR_MODEL_MS_BOUNDED_FOURIER ||
      d->model == INR_MODEL_GABOR)
    return mfn_plan_create(d, out);
  if (d->model != INR_MODEL_SIREN && d->model != INR_MODEL_FFN) return fail(INR_EUNSUPPORTED, "model kind not built yet");
  // the on-chip chain kernels hold one 128 x 256 activation tile per SM; other widths (8 shipped SIREN configs use 512), a
  // single sine layer (network_depth 1 or 2) and the sine output layer run layer by layer on the streaming stage GEMMs
  if (d->width != kWidth || d->depth < 3 || d->last_act == INR_LAST_SIN || d->encoder == INR_ENC_LOGF) return wide_plan_create(d, out);
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
    p->tensors.push_back({off, rows, cols, l, 0, 0, 0});
    SegDesc sw{};
    sw.off = off; sw.rows = rows; sw.cols = cols; sw.layer = l;
    sw.fwd_scale = sw.bwd_scale = (M.act == ACT_SIN) ? M.w0 : 1.f;
    sw.scale_slot = -1; sw.gfin_off = -1;
    if (l < M.n_gemm) {
      sw.pack_fwd = 1;
      sw.perm_e = (l == 0 && M.input_kind == INPUT_GAUSS) ? M.enc_size : 0;
      sw.wf_off = woff; M.wf_off[l] = woff; woff += static_cast<uint32_t>(rows) * cols * 2;
      if (l >= 1) { sw.pack_bwd = 1; sw.wd_off = woff; M.wd_off[l] = woff; woff += static_cast<uint32_t>(rows) * cols * 2; }
    }
    p->segs.push_back(sw);
    off += rows * cols;
    M.b_off[l] = off;
    p->tensors.push_back({off, rows, 1, l, 1, 0, 0});
    SegDesc sb{};
    sb.off = off; sb.rows = rows; sb.cols = 1; sb.layer = l; sb.scale_slot = -1; sb.gfin_off = -1;
    p->segs.push_back(sb);
    off += rows;
  }
  M.n_params = off;
  M.wpack_bytes = woff;
  // wgrad work units: 128 output features x up to 3 chunks of 128 input features per unit
  for (int l = 0; l < M.n_gemm; ++l) {
    const int K = l == 0 ? M.k0 : kWidth;
    for (int mh = 0; mh < kWidth / 128; ++mh)
      for (int c0 = 0, per = chunk_group(K / 128); c0 < K / 128; c0 += per) {
        const int nch = (K / 128 - c0) < per ? (K / 128 - c0) : per;
        WgradUnit u{};
        u.a_tile_stride = kActBytes; u.a_sub = mh * 32768; u.a_bytes = 32768;
        u.b_tile_stride = kTileM * K * 2; u.b_sub = c0 * 32768; u.b_bytes = 32768;
        u.n = 128; u.n_chunks = nch; u.transposed = 0;
        u.out_off = M.w_off[l]; u.out_ld = K; u.row0 = mh * 128; u.col0 = c0 * 128;
        u.rows_valid = 128; u.cols_valid = 128 * nch;
        u.bias_off = c0 == 0 ? M.b_off[l] : -1;
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
  build_wgrad_sched(p);
  *out = p;
  return INR_OK;
}

extern "C" int inr_plan_destroy(inr_plan* p) { delete p; return INR_OK; }
extern "C" int inr_plan_param_count(const inr_plan* p, int64_t* n) {
  if (!p || !n) return fail(INR_EINVAL, "null argument");
  *n = p->is_wire ? p->wm.n_params : (p->is_mfn ? p->mm.n_params : p->model.n_params); return INR_OK;
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
  *b = p->is_wire ? p->wm.wpack_bytes : (p->is_mfn ? p->mm.wpack_bytes : p->model.wpack_bytes); return INR_OK;
}

// Rows per tile of the width-256 chain: 128 (chain_fwd.cu / chain_bwd.cu).  With INR_CHAIN_T=1 a batch below one wave of
// 80-row tiles (BASELINE configs[0]: 10 000 rows = 79 128-row tiles on 148 SMs) is cut into one tile of NR <= 80 rows per SM
// for the transposed kernels of chain_t.cu instead.  Opt-in: measured on B200 they halve the epilogue but expose the MMA phases
// (71.9 us against 59.7 us per step at 10 000 rows).
static int chain_tile_rows(const inr_plan* p, int64_t bs) {
  const char* e = std::getenv("INR_CHAIN_T");          // read per call: the parity tests run both tilings in one process
  const int on = e ? std::atoi(e) : 0;
  const int kTMaxRows = 80, kTMaxK0 = 512;
  if (!on || p->model.k0 > kTMaxK0 || bs > static_cast<int64_t>(p->n_sm) * kTMaxRows) return kTileM;
  const int nr = 16 * static_cast<int>((bs + 16ll * p->n_sm - 1) / (16ll * p->n_sm));
  return nr < 16 ? 16 : nr;
}

static Workspace plan_workspace(const inr_plan* p, int64_t bs) {
  const ChainModel& M = p->model;
  Workspace w{};
  w.tile_rows = chain_tile_rows(p, bs);
  w.lb = w.tile_rows == kTileM ? 2048 : w.tile_rows * 16 + 16;
  const int T = static_cast<int>((bs + w.tile_rows - 1) / w.tile_rows);
  w.n_tiles = T;
  const int ns = wgrad_splits(p, T);
  w.n_split = ns;
  const uint64_t lb = static_cast<uint64_t>(w.lb);
  uint64_t o = 0;
  w.scal_off = o; o += align_up(kScalars * 4, 1024);
  w.part_off = o; o += align_up(static_cast<uint64_t>(T) * kPartialsPerTile * 4, 1024);
  // loss records, indexed by batch row; the TV kernel walks them in 128-row blocks whatever the tile height
  const uint64_t g_rows = std::max<uint64_t>(static_cast<uint64_t>(T) * w.tile_rows, (static_cast<uint64_t>(bs) + 127) / 128 * 128);
  w.g_off = o; o += align_up(g_rows * 16, 1024);
  for (int l = 0; l <= M.n_gemm; ++l) {
    const int K = l == 0 ? M.k0 : kWidth;
    w.h_off[l] = o; o += align_up(static_cast<uint64_t>(T) * (K / 8) * lb, 1024);
  }
  for (int l = 0; l < M.n_gemm; ++l) { w.d_off[l] = o; o += align_up(static_cast<uint64_t>(T) * (kWidth / 8) * lb, 1024); }
  for (int l = 0; l < M.n_gemm; ++l) { w.dz_off[l] = o; o += align_up(static_cast<uint64_t>(T) * (kWidth / 8) * lb, 1024); }
  w.dzlast_off = o; o += align_up(static_cast<uint64_t>(T) * 2 * lb, 1024);
  w.gpart_off = o; o += align_up(static_cast<uint64_t>(ns) * gpart_stride(M.n_params) * 4, 1024);
  w.total = o;
  return w;
}

// A workspace sized for `bs` rows serves every smaller batch too.  The chain's tile height steps with the batch (short tiles
// carry 16 padding bytes per k-group), so its footprint is not monotone across a step: take the largest one up to bs.
static uint64_t chain_workspace_bytes(const inr_plan* p, int64_t bs) {
  uint64_t t = plan_workspace(p, bs).total;
  for (int k = 1; k <= 5; ++k) {
    const int64_t edge = static_cast<int64_t>(p->n_sm) * 16 * k;
    if (edge < bs) t = std::max<uint64_t>(t, plan_workspace(p, edge).total);
  }
  return t;
}

extern "C" int inr_workspace_bytes(const inr_plan* p, int64_t bs, size_t* bytes) {
  if (!p || !bytes || bs <= 0) return fail(INR_EINVAL, "bad argument");
  *bytes = p->is_wire ? wire_workspace(p, bs).total : (p->is_mfn ? mfn_workspace(p, bs).total : chain_workspace_bytes(p, bs)); return INR_OK;
}
extern "C" int inr_scalars_offset(const inr_plan* p, int64_t bs, size_t* off) {
  if (!p || !off || bs <= 0) return fail(INR_EINVAL, "bad argument");
  *off = p->is_wire ? wire_workspace(p, bs).scal : (p->is_mfn ? mfn_workspace(p, bs).scal : plan_workspace(p, bs).scal_off); return INR_OK;
}

extern "C" int inr_workspace_layout(const inr_plan* p, int64_t bs, uint64_t* out, int32_t n) {
  if (!p || !out || bs <= 0 || n < 44) return fail(INR_EINVAL, "bad argument");
  if (p->is_wire) {   // H_hi at [l], H_lo at [12+l] is not representable in 12 slots each: report hi / ab / dz families
    const WireWorkspace w = wire_workspace(p, bs);
    for (int l = 0; l < kMaxLayers; ++l) {
      out[l] = l <= p->wm.depth + 1 ? w.hhi[l] : 0; out[12 + l] = l <= p->wm.depth ? w.ab[l] : 0;
      out[24 + l] = l <= p->wm.depth ? w.dz[l] : 0;
    }
    out[36] = w.dzlast; out[37] = w.g; out[38] = w.part; out[39] = w.scal; out[40] = w.gpart;
    out[41] = static_cast<uint64_t>(w.n_tiles); out[42] = static_cast<uint64_t>(w.n_split); out[43] = w.total;
    return INR_OK;
  }
  if (p->is_mfn) {     // z images at [l], sin(p) at [12+l], dp at [24+l]
    const MfnWorkspace w = mfn_workspace(p, bs);
    for (int l = 0; l < kMaxLayers; ++l) {
      out[l] = l <= p->mm.top ? w.z[l] : 0; out[12 + l] = l <= p->mm.top ? w.g[l] : 0; out[24 + l] = l <= p->mm.top ? w.dp[l] : 0;
    }
    out[36] = w.x; out[37] = w.gl; out[38] = w.part; out[39] = w.scal; out[40] = w.gpart;
    out[41] = static_cast<uint64_t>(w.n_tiles); out[42] = static_cast<uint64_t>(w.n_split); out[43] = w.total;
    if (n >= 60) {     // Gabor extras: q images at [44+l], then gfin, gpart float stride, aux offset of stage 0, envelope image
      for (int l = 0; l < kMaxLayers; ++l) out[44 + l] = (p->mm.gabor && l <= p->mm.top) ? w.q[l] : 0;
      out[56] = w.gfin; out[57] = static_cast<uint64_t>(gpart_stride(p->mm.g_floats));
      out[58] = static_cast<uint64_t>(p->mm.gabor ? p->mm.aux_off[0] : 0); out[59] = w.e;
    }
    return INR_OK;
  }
  const Workspace w = plan_workspace(p, bs);
  for (int l = 0; l < kMaxLayers; ++l) { out[l] = w.h_off[l]; out[12 + l] = w.d_off[l]; out[24 + l] = w.dz_off[l]; }
  out[36] = w.dzlast_off; out[37] = w.g_off; out[38] = w.part_off; out[39] = w.scal_off; out[40] = w.gpart_off;
  out[41] = static_cast<uint64_t>(w.n_tiles); out[42] = static_cast<uint64_t>(w.n_split); out[43] = w.total;
  if (n >= 46) { out[44] = static_cast<uint64_t>(w.tile_rows); out[45] = static_cast<uint64_t>(w.lb); }
  return INR_OK;
}

static void fill_adam(const inr_plan* p, AdamArgs& a) {
  std::memset(&a, 0, sizeof(a));
  a.n_seg = static_cast<int>(p->segs.size());
  for (int i = 0; i < a.n_seg; ++i) a.seg[i] = p->segs[i];
  a.n_params = p->is_mfn ? p->mm.n_params : p->model.n_params;
  a.gstride = gpart_stride(p->is_mfn ? p->mm.g_floats : a.n_params);
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
    if (w.lb != 2048) {      // the plan's units are written for 2048-byte k-groups (128-row tiles): rescale to this batch's tiles
      const uint32_t lb = static_cast<uint32_t>(w.lb);
      u.a_tile_stride = u.a_tile_stride / 2048 * lb; u.b_tile_stride = u.b_tile_stride / 2048 * lb;
      u.a_sub = u.a_sub / 2048 * lb; u.b_sub = u.b_sub / 2048 * lb;
      u.a_bytes = u.a_bytes / 2048 * lb; u.b_bytes = u.b_bytes / 2048 * lb;
    }
    g.u[i] = u;
  }
  fill_wgrad_sched(p, g);
  g.tile_rows = w.tile_rows; g.lb = w.lb;
  g.n_split = w.n_split; g.n_tiles = w.n_tiles; g.n_params = gpart_stride(M.n_params);
  g.ws = ws; g.gpart_off = w.gpart_off;
}


// =====================================================================================================================
// WIRE (complex Gabor) path
// =====================================================================================================================
static int query_sm_count() {
  int dev = 0, n = 0, sm = 148;
  if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0 && cudaGetDevice(&dev) == cudaSuccess) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sm = v;
  } else {
    cudaGetLastError();
  }
  return sm;
}

static int wire_plan_create(const inr_model_desc* d, inr_plan** out) {
  const int c = static_cast<int>(static_cast<double>(d->width) / 1.4142135623730951);   // int(width / sqrt(2)), networks.py:228
  if (d->in_features != 3) return fail(INR_EUNSUPPORTED, "WIRE kernels take raw (coil, kx, ky) coordinates: network_input_size 3");
  if (d->encoder != INR_ENC_NONE) return fail(INR_EUNSUPPORTED, "WIRE is used without positional encoding");
  if (c < 8 || c > kWP) return fail(INR_EUNSUPPORTED, "WIRE kernels are built for complex width <= 192 (network_width <= 271)");
  if (d->depth < 1 || d->depth > kWMaxDepth) return fail(INR_EINVAL, "network_depth out of range");
  if (d->out_features < 1 || d->out_features > 2) return fail(INR_EUNSUPPORTED, "network_output_size must be 1 or 2");
  inr_plan* p = new (std::nothrow) inr_plan();
  if (!p) return fail(INR_EINVAL, "out of host memory");
  p->desc = *d;
  p->is_wire = true;
  WireModel& M = p->wm;
  std::memset(&M, 0, sizeof(M));
  M.depth = d->depth; M.c = c; M.in_f = 3; M.out_f = d->out_features; M.nlin = 1; M.P = kWP;
  M.omega_first = d->w0; M.omega_hidden = d->hidden_omega_0; M.sigma = d->sigma0;
  int off = 0;
  const int L = M.depth + 1;
  for (int l = 0; l <= L; ++l) {
    if (l < L) {
      M.omega_off[l] = off; p->tensors.push_back({off, 1, 1, l, 0, 0, 1}); off += 1;
      M.scale_off[l] = off; p->tensors.push_back({off, 1, 1, l, 0, 0, 1}); off += 1;
    }
    if (l == 0) {
      M.w_off[l] = off; p->tensors.push_back({off, c, 3, l, 0, 0, 0}); off += c * 3;
      M.b_off[l] = off; p->tensors.push_back({off, c, 1, l, 1, 0, 0}); off += c;
    } else {
      const int rows = l == L ? M.out_f : c;
      M.w_off[l] = off; p->tensors.push_back({off, rows, c, l, 0, 1, 0}); off += rows * c * 2;
      M.b_off[l] = off; p->tensors.push_back({off, rows, 1, l, 1, 1, 0}); off += rows * 2;
    }
  }
  M.n_params = off;
  const uint32_t blk = 2u * (kW2 / kStageK) * kWStageBBytes;      // two N-blocks of one packed operand
  uint32_t wo = 0;
  for (int l = 1; l <= M.depth; ++l) {
    M.wf_hi[l] = wo; wo += blk;
    M.wf_lo[l] = wo; wo += blk;
    M.wd_hi[l] = wo; wo += blk;
  }
  M.wpack_bytes = wo;
  int go = 0;
  for (int l = 1; l <= M.depth; ++l) { M.gd_hidden[l] = go; go += kW2 * kW2 + kW2; }
  M.gd_final = go; go += 16 * kW2 + 16;
  M.gd_first = go; go += 256 * 16;
  M.gd_floats = go;
  // wgrad units (offsets into the workspace are filled per call)
  for (int l = 1; l <= M.depth; ++l)
    for (int mc = 0; mc < 3; ++mc) {         // 128 rows of D = dZ^T [hr|hi] against all 384 columns (3 chunks)
      WgradUnit u{};
      u.a_tile_stride = kWTileBytes; u.a_sub = mc * 32768; u.a_bytes = 32768;
      u.b_tile_stride = kWTileBytes; u.b_sub = 0; u.b_bytes = 32768;
      u.n = 128; u.n_chunks = 3; u.transposed = 0;
      u.out_off = M.gd_hidden[l]; u.out_ld = kW2; u.row0 = mc * 128; u.col0 = 0;
      u.rows_valid = 128; u.cols_valid = kW2;
      u.bias_off = M.gd_hidden[l] + kW2 * kW2;
      p->units.push_back(u); p->unit_layer.push_back(l);
    }
  for (int mc = 0; mc < 3; ++mc) {          // final layer, transposed: D^T[o][f] = sum_rows dz_last[o] * [hr|hi][f]
    WgradUnit u{};
    u.a_tile_stride = kWTileBytes; u.a_sub = mc * 32768; u.a_bytes = 32768;
    u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
    u.n = kDzLastCols; u.transposed = 1;
    u.out_off = M.gd_final; u.out_ld = kW2; u.row0 = 0; u.col0 = mc * 128;
    u.rows_valid = M.out_f; u.cols_valid = 128;
    u.bias_off = mc == 0 ? M.gd_final + 16 * kW2 : -1;
    p->units.push_back(u); p->unit_layer.push_back(L);
  }
  for (int mc = 0; mc < 2; ++mc) {          // first layer: D0[o][col] = sum_rows dza0[o] * ximg[col]
    WgradUnit u{};
    u.a_tile_stride = kWTileBytes; u.a_sub = mc * 32768; u.a_bytes = 32768;
    u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
    u.n = kDzLastCols; u.transposed = 0;
    u.out_off = M.gd_first; u.out_ld = 16; u.row0 = mc * 128; u.col0 = 0;
    u.rows_valid = 128; u.cols_valid = 16;
    u.bias_off = -1;
    p->units.push_back(u); p->unit_layer.push_back(0);
  }
  if (static_cast<int>(p->units.size()) > kMaxUnits) { delete p; return fail(INR_EUNSUPPORTED, "WIRE model too deep for the static unit table"); }
  p->n_sm = query_sm_count();
  build_wgrad_sched(p);
  *out = p;
  return INR_OK;
}

// WIRE2D (reference src/models/wire2d.py): two linears per layer, hidden width = network_width (not reduced)
static int wire2d_plan_create(const inr_model_desc* d, inr_plan** out) {
  const int c = d->width;
  if (d->in_features != 3) return fail(INR_EUNSUPPORTED, "WIRE2D kernels take raw (coil, kx, ky) coordinates: network_input_size 3");
  if (d->encoder != INR_ENC_NONE) return fail(INR_EUNSUPPORTED, "WIRE2D is used without positional encoding");
  if (c < 8 || c > kW2dMaxP) return fail(INR_EUNSUPPORTED, "WIRE2D kernels are built for network_width <= 256");
  if (d->depth < 1 || d->depth > kWMaxDepth) return fail(INR_EINVAL, "network_depth out of range");
  if (d->out_features < 1 || d->out_features > 2) return fail(INR_EUNSUPPORTED, "network_output_size must be 1 or 2");
  if (d->last_act != INR_LAST_LINEAR && d->last_act != INR_LAST_TANH)
    return fail(INR_EUNSUPPORTED, "WIRE2D output: linear (real part) or last_tanh (real part of the complex tanh)");
  inr_plan* p = new (std::nothrow) inr_plan();
  if (!p) return fail(INR_EINVAL, "out of host memory");
  p->desc = *d;
  p->is_wire = true;
  WireModel& M = p->wm;
  std::memset(&M, 0, sizeof(M));
  const int P = (c + 63) / 64 * 64;
  M.depth = d->depth; M.c = c; M.in_f = 3; M.out_f = d->out_features; M.nlin = 2; M.P = P;
  M.last_tanh = d->last_act == INR_LAST_TANH ? 1 : 0;
  M.omega_first = d->w0; M.omega_hidden = d->hidden_omega_0; M.sigma = d->sigma0;
  int off = 0;
  const int L = M.depth + 1;
  for (int l = 0; l <= L; ++l) {      // state_dict order: omega_0, scale_0, linear.{weight,bias}, scale_orth.{weight,bias}
    if (l < L) {
      M.omega_off[l] = off; p->tensors.push_back({off, 1, 1, l, 0, 0, 1}); off += 1;
      M.scale_off[l] = off; p->tensors.push_back({off, 1, 1, l, 0, 0, 1}); off += 1;
    }
    if (l == 0) {
      M.w_off[l] = off; p->tensors.push_back({off, c, 3, l, 0, 0, 0}); off += c * 3;
      M.b_off[l] = off; p->tensors.push_back({off, c, 1, l, 1, 0, 0}); off += c;
      M.v_off[l] = off; p->tensors.push_back({off, c, 3, l, 0, 0, 0}); off += c * 3;
      M.vb_off[l] = off; p->tensors.push_back({off, c, 1, l, 1, 0, 0}); off += c;
    } else {
      const int rows = l == L ? M.out_f : c;
      M.w_off[l] = off; p->tensors.push_back({off, rows, c, l, 0, 1, 0}); off += rows * c * 2;
      M.b_off[l] = off; p->tensors.push_back({off, rows, 1, l, 1, 1, 0}); off += rows * 2;
      if (l < L) {
        M.v_off[l] = off; p->tensors.push_back({off, c, c, l, 0, 1, 0}); off += c * c * 2;
        M.vb_off[l] = off; p->tensors.push_back({off, c, 1, l, 1, 1, 0}); off += c * 2;
      }
    }
  }
  M.n_params = off;
  const uint32_t fwd_bytes = static_cast<uint32_t>(P / kW2dFwdFeat) * (2 * P / 32) * (kW2dNT * 64);   // all forward N-blocks, K = 2P
  const uint32_t bwd_bytes = static_cast<uint32_t>(P / kW2dBwdFeat) * (4 * P / 32) * (kW2dNT * 64);   // all dgrad N-blocks, K = 4P
  uint32_t wo = 0;
  for (int l = 1; l <= M.depth; ++l) {
    M.wf_hi[l] = wo; wo += fwd_bytes;
    M.wf_lo[l] = wo; wo += fwd_bytes;
    M.wd_hi[l] = wo; wo += bwd_bytes;
  }
  M.wpack_bytes = wo;
  const int K2 = 2 * P, Z4 = 4 * P;
  int go = 0;
  for (int l = 1; l <= M.depth; ++l) { M.gd_hidden[l] = go; go += Z4 * K2 + Z4; }
  M.gd_final = go; go += 16 * K2 + 16;
  M.gd_first = go; go += Z4 * 16;
  M.gd_floats = go;
  const uint32_t th = static_cast<uint32_t>(kTileM) * K2 * 2, tz = static_cast<uint32_t>(kTileM) * Z4 * 2;
  const int hc = K2 / 128, zc = Z4 / 128;
  for (int l = 1; l <= M.depth; ++l)
    for (int mc = 0; mc < zc; ++mc)            // 128 rows of D = dZ^T [hr|hi]
      for (int c0 = 0, per = chunk_group(hc); c0 < hc; c0 += per) {
        const int nch = (hc - c0) < per ? (hc - c0) : per;
        WgradUnit u{};
        u.a_tile_stride = tz; u.a_sub = mc * 32768; u.a_bytes = 32768;
        u.b_tile_stride = th; u.b_sub = c0 * 32768; u.b_bytes = 32768;
        u.n = 128; u.n_chunks = nch; u.transposed = 0;
        u.out_off = M.gd_hidden[l]; u.out_ld = K2; u.row0 = mc * 128; u.col0 = c0 * 128;
        u.rows_valid = 128; u.cols_valid = 128 * nch;
        u.bias_off = c0 == 0 ? M.gd_hidden[l] + Z4 * K2 : -1;
        p->units.push_back(u); p->unit_layer.push_back(l);
      }
  for (int mc = 0; mc < hc; ++mc) {            // final layer, transposed
    WgradUnit u{};
    u.a_tile_stride = th; u.a_sub = mc * 32768; u.a_bytes = 32768;
    u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
    u.n = kDzLastCols; u.transposed = 1;
    u.out_off = M.gd_final; u.out_ld = K2; u.row0 = 0; u.col0 = mc * 128;
    u.rows_valid = M.last_tanh ? 4 : M.out_f; u.cols_valid = 128;     // tanh tail: dz_last = [gx0 gx1 | gy0 gy1] (w2d_blast_kernel)
    u.bias_off = mc == 0 ? M.gd_final + 16 * K2 : -1;
    p->units.push_back(u); p->unit_layer.push_back(L);
  }
  for (int mc = 0; mc < zc; ++mc) {            // first layer: D0[row][col] = sum_rows dZ0[row] * ximg[col] (b / d rows stay zero)
    WgradUnit u{};
    u.a_tile_stride = tz; u.a_sub = mc * 32768; u.a_bytes = 32768;
    u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
    u.n = kDzLastCols; u.transposed = 0;
    u.out_off = M.gd_first; u.out_ld = 16; u.row0 = mc * 128; u.col0 = 0;
    u.rows_valid = 128; u.cols_valid = 16;
    u.bias_off = -1;
    p->units.push_back(u); p->unit_layer.push_back(0);
  }
  if (static_cast<int>(p->units.size()) > kMaxUnits) { delete p; return fail(INR_EUNSUPPORTED, "WIRE2D model too deep for the static unit table"); }
  p->n_sm = query_sm_count();
  build_wgrad_sched(p);
  *out = p;
  return INR_OK;
}

static WireWorkspace wire_workspace(const inr_plan* p, int64_t bs) {
  const WireModel& M = p->wm;
  WireWorkspace w{};
  const int T = static_cast<int>((bs + kTileM - 1) / kTileM);
  w.n_tiles = T;
  const int ns = wgrad_splits(p, T);
  w.n_split = ns;
  uint64_t o = 0;
  w.scal = o; o += align_up(kScalars * 4, 1024);
  // forward hand-over counters: same place for every batch size (a chain that starts the step cannot have them zeroed by a
  // kernel before it: wire_last zeroes what the chain used, and the block starts out zero -- include/inr_b200.h)
  // (batches above kWFlagTiles row tiles get a larger block; the first-layer fold is off for them and wire_first zeroes it)
  w.flags_fwd = o; o += align_up(static_cast<uint64_t>(kWMaxDepth) * (T > kWFlagTiles ? T : kWFlagTiles) * 4, 1024);
  w.part = o; o += align_up(static_cast<uint64_t>(T) * kPartialsPerTile * 4, 1024);
  w.g = o; o += align_up(static_cast<uint64_t>(T) * kTileM * 16, 1024);
  w.outacc = o; o += align_up(static_cast<uint64_t>(T) * kWOutParts * kTileM * 16, 1024);   // partial outputs of the final linear
  w.flags_bwd = o; o += align_up(static_cast<uint64_t>(kWMaxDepth) * T * 4, 1024);          // dgrad hand-over counters
  const uint64_t img = static_cast<uint64_t>(T) * kTileM * 2 * M.P * 2;            // H images: [hr | hi]
  const uint64_t zimg = img * M.nlin;                                              // pre-activation / gradient images
  for (int l = 1; l <= M.depth + 1; ++l) { w.hhi[l] = o; o += img; w.hlo[l] = o; o += img; }
  for (int l = 0; l <= M.depth; ++l) { w.ab[l] = o; o += zimg; }
  for (int l = 0; l <= M.depth; ++l) { w.dz[l] = o; o += zimg; }
  w.dzlast = o; o += align_up(static_cast<uint64_t>(T) * kDzLastBytes, 1024);
  w.ximg = o; o += align_up(static_cast<uint64_t>(T) * kDzLastBytes, 1024);
  w.gpart = o; o += align_up(static_cast<uint64_t>(ns) * M.gd_floats * 4, 1024);
  w.total = o;
  return w;
}

static void wire_aux_fill(const inr_plan* p, const WireWorkspace& w, WireAuxArgs& x, const float* params, void* ws, int64_t bs) {
  std::memset(&x, 0, sizeof(x));
  x.m = p->wm; x.w = w; x.params = params; x.ws = static_cast<uint8_t*>(ws);
  x.bs = static_cast<int>(bs); x.bs_k = static_cast<int>(bs);
  x.loss = LossDesc{LOSS_NONE, 0.f, 0.f, 0.f};
}

static int wire_forward_impl(const inr_plan* p, const WireWorkspace& w, const LossDesc& loss, const float* params, const void* wpack,
                             const float* coords, const float* gt, const uint8_t* mask, int64_t bs, void* ws, float* out, int train,
                             const int* row_off, int* step, cudaStream_t st, cudaEvent_t* gemm_ev = nullptr,
                             bool fold_scalars = false, const float* hyper = nullptr) {
  NvtxRange nv("inr.wire.forward+loss");
  const WireModel& M = p->wm;
  uint8_t* W = static_cast<uint8_t*>(ws);
  const uint8_t* wp = static_cast<const uint8_t*>(wpack);
  WireAuxArgs x; wire_aux_fill(p, w, x, params, ws, bs);
  x.loss = loss; x.coords = coords; x.gt = gt; x.mask = mask; x.out = out; x.train = train;
  x.row_offset = row_off; x.step_counter = step;
  // WIRE, opt-in (INR_WIRE_FOLD_FIRST=1): the real first layer rides in the forward chain as its first items -- one launch less
  const bool fold_first = M.nlin == 1 && M.depth + 1 <= kWMaxDepth && w.n_tiles <= kWFlagTiles && wire_fold_first();
  cudaError_t e = cudaSuccess;
  if (!fold_first) {
    e = M.nlin == 2 ? launch_w2d_first(x, st) : launch_wire_first(x, st);
    if (e != cudaSuccess) return cuda_fail(e, "wire_first_kernel");
  }
  if (gemm_ev) cudaEventRecord(gemm_ev[0], st);
  if (M.nlin == 2) {       // WIRE2D: all hidden layers as ONE chained launch (hand-over through w.flags_fwd, zeroed by w2d_first)
    LGemmArgs g{};
    g.seg[0].a_tile_bytes = static_cast<uint32_t>(kTileM) * 2 * M.P * 2; g.seg[0].k_stages = 2 * M.P / kStageK; g.seg[0].acc_col = 0;
    g.n_seg = 1; g.nt = kW2dNT; g.n_tiles = w.n_tiles; g.n_nblocks = M.P / kW2dFwdFeat; g.passes = 3; g.mode = LG_W2D_FWD;
    g.sigma = M.sigma; g.c_valid = M.c; g.p2 = M.P; g.train = train;
    g.chain_len = M.depth; g.chain_flags = reinterpret_cast<unsigned int*>(W + w.flags_fwd);
    for (int l = 1; l <= M.depth; ++l) {
      LGemmLayer& c = g.chain[l - 1];
      c.a_hi = W + w.hhi[l]; c.a_lo = W + w.hlo[l];
      c.b_hi = wp + M.wf_hi[l]; c.b_lo = wp + M.wf_lo[l];
      c.bias = params + M.b_off[l]; c.bias2 = params + M.vb_off[l]; c.omega = M.omega_hidden;
      c.out_hi = W + w.hhi[l + 1]; c.out_lo = W + w.hlo[l + 1]; c.out_ab = W + w.ab[l];
    }
    e = launch_lgemm((g.trace = lgemm_trace_ptr(), g.dbg = lgemm_dbg(), g), p->n_sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(wire2d fwd)");
  }
  if (M.nlin == 1) {       // all hidden layers as ONE chained launch (tile hand-over through w.flags_fwd, zeroed by wire_first)
    LGemmArgs g{};
    g.seg[0].a_tile_bytes = kWTileBytes; g.seg[0].k_stages = kW2 / kStageK; g.seg[0].acc_col = 0; g.n_seg = 1; g.nt = kWNT;
    g.n_tiles = w.n_tiles; g.n_nblocks = 2; g.passes = 3; g.mode = LG_WIRE_FWD;
    g.sigma = M.sigma; g.c_valid = M.c; g.train = train; g.out_f = M.out_f;
    const int f0 = fold_first ? 1 : 0;
    g.chain_len = M.depth + f0; g.chain_flags = reinterpret_cast<unsigned int*>(W + w.flags_fwd);
    if (fold_first) {       // chain[0]: z = x W0^T + b0 from the coordinates, Gabor wavelet, H images of layer 1 + (a) image + coordinate image
      LGemmLayer& c = g.chain[0];
      c.bias = params + M.b_off[0]; c.omega = M.omega_first;
      c.out_hi = W + w.hhi[1]; c.out_lo = W + w.hlo[1]; c.out_ab = W + w.ab[0];
      g.first_w = params + M.w_off[0]; g.coords = coords; g.row_offset = row_off; g.step_counter = step;
      g.ximg = W + w.ximg; g.bs = static_cast<int>(bs);
    }
    for (int l = 1; l <= M.depth; ++l) {
      LGemmLayer& c = g.chain[l - 1 + f0];
      c.a_hi = W + w.hhi[l]; c.a_lo = W + w.hlo[l];
      c.b_hi = wp + M.wf_hi[l]; c.b_lo = wp + M.wf_lo[l];
      c.bias = params + M.b_off[l]; c.omega = M.omega_hidden;
      c.out_hi = W + w.hhi[l + 1]; c.out_lo = W + w.hlo[l + 1]; c.out_ab = W + w.ab[l];
      if (l == M.depth) {     // the final linear rides in this layer's epilogue (partial sums per row; out_f <= 2, depth >= 1)
        c.last_w = params + M.w_off[M.depth + 1]; c.out_part = reinterpret_cast<float*>(W + w.outacc);
        c.out_lo = nullptr;   // H_lo of the last hidden layer had one reader, the final linear
      }
    }
    e = launch_lgemm((g.trace = lgemm_trace_ptr(), g.dbg = lgemm_dbg(), g), p->n_sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(fwd)");
  }
  if (gemm_ev) cudaEventRecord(gemm_ev[1], st);
  x.step_counter = nullptr;
  if (fold_scalars && M.nlin == 1 && train) { x.fold_scalars = 1; x.hyper = hyper; x.step = step; }
  e = M.nlin == 2 ? launch_w2d_last(x, st) : launch_wire_last(x, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "wire_last_kernel");
}

static int wire_backward_impl(const inr_plan* p, const WireWorkspace& w, const LossDesc& loss, const float* params, const void* wpack,
                              const float* dout, int64_t bs, void* ws, const float* hyper, const int* step, cudaStream_t st,
                              bool scalars_done = false) {
  NvtxRange nv("inr.wire.backward(dgrad+wgrad)");
  const WireModel& M = p->wm;
  uint8_t* W = static_cast<uint8_t*>(ws);
  const uint8_t* wp = static_cast<const uint8_t*>(wpack);
  WireAuxArgs x; wire_aux_fill(p, w, x, params, ws, bs);
  x.loss = loss; x.dout = dout; x.hyper = hyper; x.step = step;
  cudaError_t e;
  if (dout) { e = launch_wire_dout_amax(x, st); if (e != cudaSuccess) return cuda_fail(e, "wire_dout_amax_kernel"); }
  if (!scalars_done) {      // the fused step computed them in wire_last's last CTA
    e = launch_wire_scalars(x, st);
    if (e != cudaSuccess) return cuda_fail(e, "wire_scalars_kernel");
  }
  // WIRE: the backward of the final linear rides in the dgrad chain as its first ("top") items -- one launch less, and the
  // chain's CTAs start on it while wire_last's tail is still draining (INR_WIRE_FOLD_BLAST=0: separate wire_blast_kernel)
  const bool fold_blast = M.nlin == 1 && M.depth + 1 <= kWMaxDepth && wire_fold_blast();
  if (!fold_blast) {
    e = M.nlin == 2 ? launch_w2d_blast(x, st) : launch_wire_blast(x, st);
    if (e != cudaSuccess) return cuda_fail(e, "wire_blast_kernel");
  }
  if (M.nlin == 2) {      // WIRE2D dgrad, layers depth .. 1 as ONE chained launch (hand-over through w.flags_bwd, zeroed by w2d_blast)
    LGemmArgs g{};
    g.seg[0].a_tile_bytes = static_cast<uint32_t>(kTileM) * 4 * M.P * 2; g.seg[0].k_stages = 4 * M.P / kStageK; g.seg[0].acc_col = 0;
    g.n_seg = 1; g.nt = kW2dNT; g.n_tiles = w.n_tiles; g.n_nblocks = M.P / kW2dBwdFeat; g.passes = 1; g.mode = LG_W2D_DGRAD;
    g.sigma = M.sigma; g.c_valid = M.c; g.p2 = M.P;
    g.scal = reinterpret_cast<const float*>(W + w.scal);
    g.chain_len = M.depth; g.chain_flags = reinterpret_cast<unsigned int*>(W + w.flags_bwd);
    for (int l = M.depth; l >= 1; --l) {
      LGemmLayer& c = g.chain[M.depth - l];
      c.a_hi = W + w.dz[l]; c.b_hi = wp + M.wd_hi[l];
      c.omega = (l - 1 == 0) ? M.omega_first : M.omega_hidden;
      c.real_first = (l - 1 == 0) ? 1 : 0;
      c.in_y = W + w.hhi[l]; c.in_ab = W + w.ab[l - 1]; c.out_dz = W + w.dz[l - 1];
      c.src_layer = l; c.dst_layer = l - 1;
    }
    e = launch_lgemm((g.trace = lgemm_trace_ptr(), g.dbg = lgemm_dbg(), g), p->n_sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(wire2d dgrad)");
  }
  if (M.nlin == 1) {      // dL/dh_{l-1} = dZ_l * conj-block(W_l), then the Gabor derivative of layer l-1: layers depth .. 1 as ONE
    LGemmArgs g{};        // chained launch (tile hand-over through w.flags_bwd, zeroed by wire_blast)
    g.seg[0].a_tile_bytes = kWTileBytes; g.seg[0].k_stages = kW2 / kStageK; g.seg[0].acc_col = 0; g.n_seg = 1; g.nt = kWNT;
    g.n_tiles = w.n_tiles; g.n_nblocks = 2; g.passes = 1; g.mode = LG_WIRE_DGRAD;
    g.sigma = M.sigma; g.c_valid = M.c;
    g.scal = reinterpret_cast<const float*>(W + w.scal);
    const int top = fold_blast ? 1 : 0;
    g.chain_len = M.depth + top; g.chain_flags = reinterpret_cast<unsigned int*>(W + w.flags_bwd);
    if (fold_blast) {       // chain[0]: dL/dh_L from dL/dout * W_last, Gabor derivative of the last hidden layer, dz_last image
      LGemmLayer& c = g.chain[0];
      c.omega = M.omega_hidden; c.real_first = 0;
      c.in_y = W + w.hhi[M.depth + 1]; c.in_ab = W + w.ab[M.depth]; c.out_dz = W + w.dz[M.depth];
      c.src_layer = -1; c.dst_layer = M.depth;
      g.top_w = params + M.w_off[M.depth + 1];
      g.top_g = reinterpret_cast<const float*>(W + w.g);
      g.top_dout = dout;
      g.top_dzlast = W + w.dzlast;
      g.bs = static_cast<int>(bs); g.out_f = M.out_f;
    }
    for (int l = M.depth; l >= 1; --l) {
      LGemmLayer& c = g.chain[M.depth - l + top];
      c.a_hi = W + w.dz[l]; c.b_hi = wp + M.wd_hi[l];
      c.omega = (l - 1 == 0) ? M.omega_first : M.omega_hidden;
      c.real_first = (l - 1 == 0) ? 1 : 0;
      c.in_y = W + w.hhi[l]; c.in_ab = W + w.ab[l - 1]; c.out_dz = W + w.dz[l - 1];
      c.src_layer = l; c.dst_layer = l - 1;
    }
    e = launch_lgemm((g.trace = lgemm_trace_ptr(), g.dbg = lgemm_dbg(), g), p->n_sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(dgrad)");
  }
  WgradArgs wg; std::memset(&wg, 0, sizeof(wg)); wg.trace = g_trace;
  wg.n_units = static_cast<int>(p->units.size());
  const int L = M.depth + 1;
  for (int i = 0; i < wg.n_units; ++i) {
    WgradUnit u = p->units[i];
    const int l = p->unit_layer[i];
    if (l == 0) { u.a_off = w.dz[0]; u.b_off = w.ximg; }
    else if (l == L) { u.a_off = w.hhi[L]; u.b_off = w.dzlast; }
    else { u.a_off = w.dz[l]; u.b_off = w.hhi[l]; }
    wg.u[i] = u;
  }
  fill_wgrad_sched(p, wg);
  wg.n_split = w.n_split; wg.n_tiles = w.n_tiles; wg.n_params = M.gd_floats; wg.ws = W; wg.gpart_off = w.gpart;
  wg.l2_hints = 1;       // the A sub-images (dZ, read once, by this kernel last) leave L2 first: measured with the chains' policy, lgemm.cu
  e = launch_wgrad(wg, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "wgrad_kernel(wire)");
}

static void wire_adam_fill(const inr_plan* p, WireAdamArgs& a) {
  std::memset(&a, 0, sizeof(a));
  a.m = p->wm;
}

// =====================================================================================================================
// MFN (multiplicative filter network) path: FourierNet, MultiscaleKFourier, MultiscaleBoundedFourier
// =====================================================================================================================
static int mfn_plan_create(const inr_model_desc* d, inr_plan** out) {
  if (d->width % kMfnNT != 0 || d->width > 512 || d->width < 128) return fail(INR_EUNSUPPORTED, "MFN kernels are built for network_width 128..512, multiple of 128");
  if (d->in_features % 128 != 0 || d->in_features > 1024) return fail(INR_EUNSUPPORTED, "MFN network_input_size must be a multiple of 128");
  if (d->depth < 1 || d->depth + 1 > kMfnMaxStages) return fail(INR_EINVAL, "network_depth out of range");
  if (d->out_features < 1 || d->out_features > 2) return fail(INR_EUNSUPPORTED, "network_output_size must be 1 or 2");
  if (d->encoder == INR_ENC_GAUSS && d->in_features != 2 * d->enc_size) return fail(INR_EINVAL, "gauss encoder needs network_input_size == 2*embedding_size");
  if (d->encoder != INR_ENC_GAUSS && d->encoder != INR_ENC_NONE) return fail(INR_EUNSUPPORTED, "the MFN kernels take the gauss encoder or dense input");
  inr_plan* p = new (std::nothrow) inr_plan();
  if (!p) return fail(INR_EINVAL, "out of host memory");
  p->desc = *d;
  p->is_mfn = true;
  MfnModel& M = p->mm;
  std::memset(&M, 0, sizeof(M));
  const int L = d->depth, W = d->width, IN = d->in_features, OF = d->out_features;
  M.L = L; M.width = W; M.in_f = IN; M.out_f = OF;
  M.input_kind = d->encoder == INR_ENC_GAUSS ? INPUT_GAUSS : INPUT_DENSE;
  M.enc_size = d->enc_size;
  const bool multi = d->model == INR_MODEL_MS_FOURIER || d->model == INR_MODEL_MS_BOUNDED_FOURIER;
  M.gabor = d->model == INR_MODEL_GABOR ? 1 : 0;
  M.bounded = d->model == INR_MODEL_MS_BOUNDED_FOURIER ? 1 : 0;
  for (int i = 0; i < kMfnMaxStages; ++i) { M.stage_head[i] = -1; M.bound_lo[i] = 0.f; M.bound_hi[i] = 3.0e38f; }
  if (multi) {
    M.n_heads = L + 1;
    const int mask = d->head_mask ? d->head_mask : 0xAA;      // reference default output_layers = [1,3,5,7]
    int live = 0;
    for (int k = 0; k <= L; ++k) {
      M.head_stage[k] = k;
      M.head_live[k] = (k >= 1 && ((mask >> k) & 1)) ? 1 : 0;
      if (M.head_live[k]) { M.stage_head[k] = live++; M.top = k; }
    }
    M.n_out = live;
    if (live == 0) { delete p; return fail(INR_EINVAL, "no output layer selected"); }
  } else {
    M.n_heads = 1; M.head_stage[0] = L; M.head_live[0] = 1; M.stage_head[L] = 0; M.top = L; M.n_out = 1;
  }
  if (M.bounded)
    for (int i = 1; i <= L; ++i) { M.bound_lo[i] = d->bounds[2 * (i - 1)]; M.bound_hi[i] = d->bounds[2 * (i - 1) + 1]; }
  // ---- parameter layout = reference state_dict order: linear.*, output_linear(.*), filters.*
  int off = 0;
  uint32_t wo = 0;
  auto add_seg = [&](int o, int rows, int cols, int stage, bool is_bias, bool dead, int scale_slot, int pf, int pb,
                     uint32_t wf, uint32_t wd, int gfin_off = -1) {
    p->tensors.push_back({o, rows, cols, stage, is_bias ? 1 : 0, 0, dead ? 1 : 0});
    SegDesc s{};
    s.off = o; s.rows = rows; s.cols = cols; s.layer = stage; s.pack_fwd = pf; s.pack_bwd = pb; s.wf_off = wf; s.wd_off = wd;
    s.fwd_scale = 1.f; s.bwd_scale = 1.f; s.layout = 1; s.nt = kMfnNT; s.scale_slot = scale_slot; s.frozen = dead ? 1 : 0; s.gfin_off = gfin_off;
    p->segs.push_back(s);
  };
  for (int i = 1; i <= L; ++i) {
    const bool dead = i > M.top;
    M.lin_w[i] = off; M.pk_lin[i] = wo; wo += static_cast<uint32_t>(W) * W * 2; M.pk_lin_t[i] = wo; wo += static_cast<uint32_t>(W) * W * 2;
    add_seg(off, W, W, i, false, dead, SC_LAYER_SCALE + i, 1, 1, M.pk_lin[i], M.pk_lin_t[i]); off += W * W;
    M.lin_b[i] = off; add_seg(off, W, 1, i, true, dead, SC_LAYER_SCALE + i, 0, 0, 0, 0); off += W;
  }
  for (int k = 0; k < M.n_heads; ++k) {
    const bool dead = !M.head_live[k];
    M.head_w[k] = off; add_seg(off, OF, W, M.head_stage[k], false, dead, -1, 0, 0, 0, 0); off += OF * W;
    M.head_b[k] = off; add_seg(off, OF, 1, M.head_stage[k], true, dead, -1, 0, 0, 0, 0); off += OF;
  }
  int gfin = 0;
  for (int i = 0; i <= L; ++i) {
    const bool dead = i > M.top;
    if (M.gabor) {      // filters.i.mu, filters.i.gamma come first (module parameters before the `linear` child)
      M.mu_off[i] = off; M.pk_mu[i] = wo; wo += static_cast<uint32_t>(W) * IN * 2; M.gfin_mu[i] = gfin; gfin += W * IN;
      add_seg(off, W, IN, i, false, dead, SC_LAYER_SCALE + i, 1, 0, M.pk_mu[i], 0, M.gfin_mu[i]); off += W * IN;
      M.gamma_off[i] = off; M.gfin_gamma[i] = gfin; gfin += W;
      add_seg(off, W, 1, i, true, dead, SC_LAYER_SCALE + i, 0, 0, 0, 0, M.gfin_gamma[i]); off += W;
    }
    M.filt_w[i] = off; M.pk_filt[i] = wo; wo += static_cast<uint32_t>(W) * IN * 2;
    add_seg(off, W, IN, i, false, dead, SC_LAYER_SCALE + i, 1, 0, M.pk_filt[i], 0); off += W * IN;
    M.filt_b[i] = off; add_seg(off, W, 1, i, true, dead, SC_LAYER_SCALE + i, 0, 0, 0, 0); off += W;
  }
  M.n_params = off;
  M.wpack_bytes = wo;
  M.gfin_floats = gfin;
  int goff = gpart_stride(off);
  if (M.gabor)
    for (int i = 0; i <= M.top; ++i) { M.aux_off[i] = goff; goff += W * 16; }
  M.g_floats = goff;
  // ---- wgrad units
  const int wc = W / 128, ic = IN / 128;
  const uint32_t wtile = static_cast<uint32_t>(kTileM) * W * 2, xtile = static_cast<uint32_t>(kTileM) * IN * 2;
  for (int i = 1; i <= M.top; ++i)          // dW_i = DH[i]^T Z[i-1], db_i = sum DH[i]
    for (int mc = 0; mc < wc; ++mc)
      for (int c0 = 0, per = chunk_group(wc); c0 < wc; c0 += per) {
        const int nch = (wc - c0) < per ? (wc - c0) : per;
        WgradUnit u{};
        u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
        u.b_tile_stride = wtile; u.b_sub = c0 * 32768; u.b_bytes = 32768;
        u.n = 128; u.n_chunks = nch; u.out_off = M.lin_w[i]; u.out_ld = W; u.row0 = mc * 128; u.col0 = c0 * 128;
        u.rows_valid = 128; u.cols_valid = 128 * nch; u.bias_off = (c0 == 0 && !M.bounded) ? M.lin_b[i] : -1;
        p->units.push_back(u); p->unit_layer.push_back(100 + i);
      }
  if (M.bounded)                           // db_i = sum_rows of the UNMASKED dh_i: D[o][0] against a ones operand
    for (int i = 1; i <= M.top; ++i)
      for (int mc = 0; mc < wc; ++mc) {
        WgradUnit u{};
        u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
        u.b_tile_stride = 0; u.b_sub = 0; u.b_bytes = kDzLastBytes;
        u.n = kDzLastCols; u.out_off = M.lin_b[i]; u.out_ld = 1; u.row0 = mc * 128; u.col0 = 0;
        u.rows_valid = 128; u.cols_valid = 1; u.bias_off = -1;
        p->units.push_back(u); p->unit_layer.push_back(400 + i);
      }
  for (int i = 0; i <= M.top; ++i)          // dOm_i = DP[i]^T X, dphi_i = sum DP[i]
    for (int mc = 0; mc < wc; ++mc)
      for (int c0 = 0, per = chunk_group(ic); c0 < ic; c0 += per) {
        const int nch = (ic - c0) < per ? (ic - c0) : per;
        WgradUnit u{};
        u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
        u.b_tile_stride = xtile; u.b_sub = c0 * 32768; u.b_bytes = 32768;
        u.n = 128; u.n_chunks = nch; u.out_off = M.filt_w[i]; u.out_ld = IN; u.row0 = mc * 128; u.col0 = c0 * 128;
        u.rows_valid = 128; u.cols_valid = 128 * nch; u.bias_off = c0 == 0 ? M.filt_b[i] : -1;
        p->units.push_back(u); p->unit_layer.push_back(200 + i);
      }
  if (M.gabor)
    for (int i = 0; i <= M.top; ++i)
      for (int mc = 0; mc < wc; ++mc) {
        for (int c0 = 0, per = chunk_group(ic); c0 < ic; c0 += per) {      // Qx_i = Q[i]^T X  (at mu_i's offset)
          const int nch = (ic - c0) < per ? (ic - c0) : per;
          WgradUnit u{};
          u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
          u.b_tile_stride = xtile; u.b_sub = c0 * 32768; u.b_bytes = 32768;
          u.n = 128; u.n_chunks = nch; u.out_off = M.mu_off[i]; u.out_ld = IN; u.row0 = mc * 128; u.col0 = c0 * 128;
          u.rows_valid = 128; u.cols_valid = 128 * nch; u.bias_off = -1;
          p->units.push_back(u); p->unit_layer.push_back(500 + i);
        }
        WgradUnit u{};                                                      // (s, u)_i = Q[i]^T [1 | |x|^2]
        u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
        u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
        u.n = kDzLastCols; u.out_off = M.aux_off[i]; u.out_ld = 16; u.row0 = mc * 128; u.col0 = 0;
        u.rows_valid = 128; u.cols_valid = 16; u.bias_off = -1;
        p->units.push_back(u); p->unit_layer.push_back(600 + i);
      }
  for (int k = 0; k < M.n_heads; ++k) {     // dV_k^T[o][f] = sum_rows dy_k[o] z[f]
    if (!M.head_live[k]) continue;
    for (int mc = 0; mc < wc; ++mc) {
      WgradUnit u{};
      u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
      u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
      u.n = kDzLastCols; u.transposed = 1; u.out_off = M.head_w[k]; u.out_ld = W; u.row0 = 0; u.col0 = mc * 128;
      u.rows_valid = OF; u.cols_valid = 128; u.bias_off = mc == 0 ? M.head_b[k] : -1;
      p->units.push_back(u); p->unit_layer.push_back(300 + k);
    }
  }
  if (static_cast<int>(p->units.size()) > kMaxUnits || static_cast<int>(p->segs.size()) > kMaxSegs) {
    delete p;
    return fail(INR_EUNSUPPORTED, "MFN model too large for the static unit / segment tables");
  }
  p->n_sm = query_sm_count();
  build_wgrad_sched(p);
  *out = p;
  return INR_OK;
}

// SIREN / FFN as a chain of plain layers on the MFN stage GEMMs (MfnModel::chain): reference src/models/networks.py:48-124.
// Sine layers: first (in -> W) + max(depth - 2, 0) hidden (W -> W); then the output layer (W -> out) with its activation.
static int wide_plan_create(const inr_model_desc* d, inr_plan** out) {
  const int W = d->width, OF = d->out_features;
  // LogF encoder (reference networks.py:14-16,24-29): n = int(embedding_size / 6) frequencies per coordinate, 6 n input
  // features; the operand image (and the packed first-layer weights) are zero-padded to the next multiple of 128
  const bool logf = d->encoder == INR_ENC_LOGF;
  const int enc_n = logf ? d->enc_size / 6 : 0;
  const int IN_LD = d->in_features;                          // columns of the first-layer weight in the parameter buffer
  const int IN = logf ? (IN_LD + 127) / 128 * 128 : IN_LD;   // K of the first-layer GEMM
  const int n_sine = d->depth >= 2 ? d->depth - 1 : 1;      // reference: depth 1 and depth 2 both build [first, last]
  if (W % kMfnNT != 0 || W < kMfnNT || W > 512) return fail(INR_EUNSUPPORTED, "SIREN / FFN kernels need network_width in {128, 256, 384, 512}");
  if (logf && (enc_n < 1 || IN_LD != 6 * enc_n)) return fail(INR_EINVAL, "LogF encoder needs network_input_size == 6 * int(embedding_size / 6)");
  if (IN % 128 != 0 || IN < 128 || IN > 2048) return fail(INR_EUNSUPPORTED, "this network_width runs on the streaming layer GEMMs: network_input_size must be a multiple of 128");
  if (d->depth < 1 || n_sine > kMfnMaxStages) return fail(INR_EINVAL, "network_depth out of range");
  if (OF < 1 || OF > kMaxOut) return fail(INR_EUNSUPPORTED, "network_output_size must be 1..4");
  if (d->encoder == INR_ENC_GAUSS) {
    if (IN != 2 * d->enc_size) return fail(INR_EINVAL, "gauss encoder needs network_input_size == 2*embedding_size");
  } else if (d->encoder != INR_ENC_NONE && !logf) {
    return fail(INR_EUNSUPPORTED, "encoder kind not built yet");
  }
  inr_plan* p = new (std::nothrow) inr_plan();
  if (!p) return fail(INR_EINVAL, "out of host memory");
  p->desc = *d;
  p->is_mfn = true;
  MfnModel& M = p->mm;
  std::memset(&M, 0, sizeof(M));
  const int L = n_sine - 1;
  M.L = L; M.width = W; M.in_f = IN; M.out_f = OF;
  M.input_kind = d->encoder == INR_ENC_GAUSS ? INPUT_GAUSS : (logf ? INPUT_LOGF : INPUT_DENSE);
  M.enc_size = d->enc_size; M.enc_n = enc_n;
  M.chain = 1; M.act = d->model == INR_MODEL_SIREN ? ACT_SIN : ACT_RELU; M.last_act = d->last_act;
  M.w0 = d->model == INR_MODEL_SIREN ? d->w0 : 1.f;
  for (int i = 0; i < kMfnMaxStages; ++i) { M.stage_head[i] = -1; M.bound_lo[i] = 0.f; M.bound_hi[i] = 3.0e38f; }
  M.n_heads = 1; M.head_stage[0] = L; M.head_live[0] = 1; M.stage_head[L] = 0; M.top = L; M.n_out = 1;
  // parameter layout = reference state_dict order: model.<l>.linear.{weight,bias}, l = 0 .. n_sine (FFN: model.<2l>.*)
  int off = 0;
  uint32_t wo = 0;
  auto add_seg = [&](int o, int rows, int cols, int stage, bool is_bias, int scale_slot, int pf, int pb, uint32_t wf, uint32_t wd) {
    p->tensors.push_back({o, rows, cols, stage, is_bias ? 1 : 0, 0, 0});
    SegDesc s{};
    s.off = o; s.rows = rows; s.cols = cols; s.layer = stage; s.pack_fwd = pf; s.pack_bwd = pb; s.wf_off = wf; s.wd_off = wd;
    s.fwd_scale = M.w0; s.bwd_scale = 1.f;      // forward operand carries w0 (accumulator = w0 W z); dgrad uses W itself
    s.layout = 1; s.nt = kMfnNT; s.scale_slot = scale_slot; s.frozen = 0; s.gfin_off = -1;
    p->segs.push_back(s);
  };
  for (int i = 0; i <= L; ++i) {
    const int cols = i == 0 ? IN : W;
    const uint32_t bytes = static_cast<uint32_t>(W) * cols * 2;
    if (i == 0) {
      M.filt_w[0] = off; M.pk_filt[0] = wo; wo += bytes;
      add_seg(off, W, IN_LD, 0, false, SC_LAYER_SCALE + 0, 1, 0, M.pk_filt[0], 0); off += W * IN_LD;
      p->segs.back().kpad = IN;      // LogF: 6 n columns packed into K = IN (the padding stays zero: the buffer starts out zero)
      M.filt_b[0] = off; add_seg(off, W, 1, 0, true, SC_LAYER_SCALE + 0, 0, 0, 0, 0); off += W;
    } else {
      M.lin_w[i] = off; M.pk_lin[i] = wo; wo += bytes; M.pk_lin_t[i] = wo; wo += bytes;
      add_seg(off, W, W, i, false, SC_LAYER_SCALE + i, 1, 1, M.pk_lin[i], M.pk_lin_t[i]); off += W * W;
      M.lin_b[i] = off; add_seg(off, W, 1, i, true, SC_LAYER_SCALE + i, 0, 0, 0, 0); off += W;
    }
  }
  M.head_w[0] = off; add_seg(off, OF, W, L, false, -1, 0, 0, 0, 0); off += OF * W;
  M.head_b[0] = off; add_seg(off, OF, 1, L, true, -1, 0, 0, 0, 0); off += OF;
  M.n_params = off;
  M.wpack_bytes = wo;
  M.gfin_floats = 0;
  M.g_floats = gpart_stride(off);
  // ---- wgrad units: dW_s = DP[s]^T (X | Z[s-1]), db_s = sum DP[s]; head as in the MFNs
  const int wc = W / 128;
  const uint32_t wtile = static_cast<uint32_t>(kTileM) * W * 2;
  for (int i = 0; i <= L; ++i) {
    const int K = i == 0 ? IN : W, kc = K / 128;
    const uint32_t btile = static_cast<uint32_t>(kTileM) * K * 2;
    for (int mc = 0; mc < wc; ++mc)
      for (int c0 = 0, per = chunk_group(kc); c0 < kc; c0 += per) {
        const int nch = (kc - c0) < per ? (kc - c0) : per;
        WgradUnit u{};
        u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
        u.b_tile_stride = btile; u.b_sub = c0 * 32768; u.b_bytes = 32768;
        u.n = 128; u.n_chunks = nch; u.out_off = i == 0 ? M.filt_w[0] : M.lin_w[i]; u.out_ld = K; u.row0 = mc * 128; u.col0 = c0 * 128;
        u.rows_valid = 128; u.cols_valid = 128 * nch; u.bias_off = c0 == 0 ? (i == 0 ? M.filt_b[0] : M.lin_b[i]) : -1;
        if (i == 0 && IN_LD != IN) {       // LogF: the parameter tensor has 6 n columns, the padded ones are dropped on store
          u.out_ld = IN_LD;
          const int left = IN_LD - c0 * 128;
          u.cols_valid = left < 128 * nch ? (left > 0 ? left : 0) : 128 * nch;
        }
        p->units.push_back(u); p->unit_layer.push_back(i == 0 ? 200 : 700 + i);
      }
  }
  for (int mc = 0; mc < wc; ++mc) {          // dV^T[o][f] = sum_rows dy[o] z_L[f]
    WgradUnit u{};
    u.a_tile_stride = wtile; u.a_sub = mc * 32768; u.a_bytes = 32768;
    u.b_tile_stride = kDzLastBytes; u.b_sub = 0; u.b_bytes = kDzLastBytes;
    u.n = kDzLastCols; u.transposed = 1; u.out_off = M.head_w[0]; u.out_ld = W; u.row0 = 0; u.col0 = mc * 128;
    u.rows_valid = OF; u.cols_valid = 128; u.bias_off = mc == 0 ? M.head_b[0] : -1;
    p->units.push_back(u); p->unit_layer.push_back(300);
  }
  if (static_cast<int>(p->units.size()) > kMaxUnits || static_cast<int>(p->segs.size()) > kMaxSegs) {
    delete p;
    return fail(INR_EUNSUPPORTED, "model too large for the static unit / segment tables");
  }
  p->n_sm = query_sm_count();
  build_wgrad_sched(p);
  *out = p;
  return INR_OK;
}

static MfnWorkspace mfn_workspace(const inr_plan* p, int64_t bs) {
  const MfnModel& M = p->mm;
  MfnWorkspace w{};
  const int T = static_cast<int>((bs + kTileM - 1) / kTileM);
  w.n_tiles = T;
  const int ns = wgrad_splits(p, T);
  w.n_split = ns;
  uint64_t o = 0;
  w.scal = o; o += align_up(kScalars * 4, 1024);
  w.part = o; o += align_up(static_cast<uint64_t>(T) * kPartialsPerTile * 4, 1024);
  w.gl = o; o += align_up(static_cast<uint64_t>(T) * kTileM * 16, 1024);
  const uint64_t wimg = static_cast<uint64_t>(T) * kTileM * M.width * 2, ximg = static_cast<uint64_t>(T) * kTileM * M.in_f * 2;
  w.x = o; o += ximg;
  for (int i = 0; i <= M.top; ++i) { w.z[i] = o; o += wimg; w.g[i] = o; o += wimg; w.cp[i] = o; o += wimg; w.dp[i] = o; o += wimg; }
  for (int i = 1; i <= M.top; ++i) { w.h[i] = o; o += wimg; w.dh[i] = o; o += wimg; }
  for (int k = 0; k < M.n_out; ++k) { w.dout[k] = o; o += align_up(static_cast<uint64_t>(T) * kDzLastBytes, 1024); }
  if (M.bounded) {
    for (int i = 1; i <= M.top; ++i) { w.dhu[i] = o; o += wimg; }
    w.ones = o; o += 4096;
  }
  if (M.gabor) {
    for (int i = 0; i <= M.top; ++i) { w.q[i] = o; o += wimg; }
    w.e = o; o += wimg;
    w.xa = o; o += align_up(static_cast<uint64_t>(T) * kDzLastBytes, 1024);
    w.xn = o; o += align_up(static_cast<uint64_t>(T) * kTileM * 4, 1024);
    w.mn = o; o += align_up(static_cast<uint64_t>(M.top + 1) * M.width * 4, 1024);
    w.gfin = o; o += align_up(static_cast<uint64_t>(M.gfin_floats) * 4, 1024);
  }
  if (M.n_out > 1) {       // fused multi-head loss (inr_train_step_dist)
    w.msg = o; o += align_up(static_cast<uint64_t>(M.n_out) * T * kTileM * 16, 1024);
    w.msp = o; o += align_up(static_cast<uint64_t>(T) * M.n_out * kPartialsPerTile * 4, 1024);
    w.dyf = o; o += align_up(static_cast<uint64_t>(T) * kTileM * M.n_out * M.out_f * 4, 1024);
  }
  w.gpart = o; o += align_up(static_cast<uint64_t>(ns) * gpart_stride(M.g_floats) * 4, 1024);
  w.total = o;
  return w;
}

static void mfn_aux_fill(const inr_plan* p, const MfnWorkspace& w, MfnAuxArgs& x, const float* params, void* ws, int64_t bs) {
  std::memset(&x, 0, sizeof(x));
  x.m = p->mm; x.w = w; x.params = params; x.ws = static_cast<uint8_t*>(ws);
  x.bs = static_cast<int>(bs); x.bs_k = static_cast<int>(bs);
  x.loss = LossDesc{LOSS_NONE, 0.f, 0.f, 0.f};
}

static int mfn_forward_impl(const inr_plan* p, const MfnWorkspace& w, const LossDesc& loss, const float* params, const void* wpack,
                            const float* coords, const float* xin, const float* encB, const float* gt, const uint8_t* mask,
                            const float* dist, int64_t bs, void* ws, float* out, int train, const int* row_off, int* step,
                            cudaStream_t st) {
  const MfnModel& M = p->mm;
  uint8_t* W = static_cast<uint8_t*>(ws);
  const uint8_t* wp = static_cast<const uint8_t*>(wpack);
  if (M.bounded && !dist) return fail(INR_EINVAL, "BoundedFourier needs dist_to_center");
  NvtxRange nv("inr.mfn.forward+loss");
  MfnAuxArgs x; mfn_aux_fill(p, w, x, params, ws, bs);
  x.loss = loss; x.coords = coords; x.x = xin; x.encB = encB; x.gt = gt; x.mask = mask; x.dist = dist; x.out = out; x.train = train;
  x.row_offset = row_off; x.step_counter = step;
  cudaError_t e = launch_mfn_encode(x, st);
  if (e != cudaSuccess) return cuda_fail(e, "mfn_encode_kernel");
  const uint32_t wtile = static_cast<uint32_t>(kTileM) * M.width * 2, xtile = static_cast<uint32_t>(kTileM) * M.in_f * 2;
  if (M.gabor) {
    e = launch_mfn_gabor_prep(x, st);
    if (e != cudaSuccess) return cuda_fail(e, "mfn_gabor_prep_kernel");
  }
  for (int i = 0; i <= M.top; ++i) {
    if (M.gabor) {      // envelope of stage i: E = exp(-gamma/2 (|x|^2 + |mu|^2 - 2 x mu^T))
      LGemmArgs ge{};
      ge.seg[0].a_hi = W + w.x; ge.seg[0].b_hi = wp + M.pk_mu[i]; ge.seg[0].a_tile_bytes = xtile; ge.seg[0].k_stages = M.in_f / 32;
      ge.seg[0].acc_col = 0; ge.n_seg = 1;
      ge.nt = kMfnNT; ge.n_tiles = w.n_tiles; ge.n_nblocks = M.width / kMfnNT; ge.passes = 1; ge.mode = LG_GABOR_E;
      ge.xn = reinterpret_cast<const float*>(W + w.xn); ge.gamma = params + M.gamma_off[i];
      ge.mn = reinterpret_cast<const float*>(W + w.mn) + static_cast<size_t>(i) * M.width;
      ge.out_hi = W + w.e; ge.feat_tile_bytes = wtile; ge.bs = static_cast<int>(bs);
      e = launch_lgemm((ge.trace = lgemm_trace_ptr(), ge.dbg = lgemm_dbg(), ge), p->n_sm, st);
      if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(gabor envelope)");
    }
    LGemmArgs g{};
    if (M.gabor) g.in_e = W + w.e;
    g.seg[0].a_hi = W + w.x; g.seg[0].b_hi = wp + M.pk_filt[i]; g.seg[0].a_tile_bytes = xtile; g.seg[0].k_stages = M.in_f / 32; g.seg[0].acc_col = 0;
    g.n_seg = 1;
    if (M.chain) {          // plain layer of a wide SIREN / FFN: ONE segment, z_i = act(w0 (W_i z_{i-1} + b_i))
      if (i >= 1) { g.seg[0].a_hi = W + w.z[i - 1]; g.seg[0].b_hi = wp + M.pk_lin[i]; g.seg[0].a_tile_bytes = wtile; g.seg[0].k_stages = M.width / 32; }
      g.act_w0 = M.w0; g.act_kind = M.act;
    } else if (i >= 1) {
      g.seg[1].a_hi = W + w.z[i - 1]; g.seg[1].b_hi = wp + M.pk_lin[i]; g.seg[1].a_tile_bytes = wtile; g.seg[1].k_stages = M.width / 32;
      g.seg[1].acc_col = kMfnNT; g.n_seg = 2;
      g.bias = params + M.lin_b[i];
    }
    g.nt = kMfnNT; g.n_tiles = w.n_tiles; g.n_nblocks = M.width / kMfnNT; g.passes = 1; g.mode = LG_MFN_FWD;
    g.phi = params + ((M.chain && i >= 1) ? M.lin_b[i] : M.filt_b[i]); g.train = train;
    g.out_hi = W + w.z[i]; g.out_lo = M.chain ? nullptr : W + w.g[i]; g.out_ab = W + w.cp[i]; g.out_h = (i >= 1 && !M.chain) ? W + w.h[i] : nullptr;
    g.feat_tile_bytes = wtile; g.bs = static_cast<int>(bs);
    if (M.bounded && i >= 1) { g.dist = dist; g.dist_row_offset = row_off; g.bound_lo = M.bound_lo[i]; g.bound_hi = M.bound_hi[i]; }
    e = launch_lgemm((g.trace = lgemm_trace_ptr(), g.dbg = lgemm_dbg(), g), p->n_sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(mfn fwd)");
  }
  x.step_counter = nullptr;
  if (train && M.n_out > 1 && loss.kind != LOSS_NONE) {     // fused multi-head step: heads + per-head loss + consistency term
    e = launch_mfn_ms_loss(x, st);
    return e == cudaSuccess ? INR_OK : cuda_fail(e, "mfn_ms_head_kernel");
  }
  e = launch_mfn_head(x, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "mfn_head_kernel");
}

static int mfn_backward_impl(const inr_plan* p, const MfnWorkspace& w, const LossDesc& loss, const float* params, const void* wpack,
                             const float* dout, const float* dist, int64_t bs, void* ws, const float* hyper, const int* step,
                             cudaStream_t st, const int* row_off = nullptr, bool keep_ms_loss = false) {
  const MfnModel& M = p->mm;
  uint8_t* W = static_cast<uint8_t*>(ws);
  const uint8_t* wp = static_cast<const uint8_t*>(wpack);
  if (M.bounded && !dist) return fail(INR_EINVAL, "BoundedFourier needs dist_to_center");
  NvtxRange nv("inr.mfn.backward(dgrad+wgrad)");
  MfnAuxArgs x; mfn_aux_fill(p, w, x, params, ws, bs);
  x.loss = loss; x.dout = dout; x.dist = dist; x.hyper = hyper; x.step = step; x.row_offset = row_off;
  if (keep_ms_loss) x.train = 2;       // the composite loss value was reduced by mfn_ms_scalars_kernel
  cudaError_t e;
  if (dout) { e = launch_mfn_dout_amax(x, st); if (e != cudaSuccess) return cuda_fail(e, "mfn_dout_amax_kernel"); }
  e = launch_mfn_scalars(x, st);
  if (e != cudaSuccess) return cuda_fail(e, "mfn_scalars_kernel");
  e = launch_mfn_top(x, st);
  if (e != cudaSuccess) return cuda_fail(e, "mfn_top_kernel");
  const uint32_t wtile = static_cast<uint32_t>(kTileM) * M.width * 2;
  for (int i = M.top; i >= 1; --i) {        // dz_{i-1} = dh_i W_i (+ head gradient of stage i-1), then (dh, dp) of stage i-1
    LGemmArgs g{};
    // wide chain: dL/d(pre-activation of layer i-1) = (DP[i] W_i) * CP[i-1] -- the stage-0 epilogue for every layer
    g.seg[0].a_hi = W + (M.chain ? w.dp[i] : w.dh[i]); g.seg[0].b_hi = wp + M.pk_lin_t[i]; g.seg[0].a_tile_bytes = wtile; g.seg[0].k_stages = M.width / 32;
    g.seg[0].acc_col = 0; g.n_seg = 1;
    g.nt = kMfnNT; g.n_tiles = w.n_tiles; g.n_nblocks = M.width / kMfnNT; g.passes = 1; g.mode = LG_MFN_DGRAD;
    g.real_first = (i - 1 == 0 || M.chain) ? 1 : 0;
    g.in_y = M.chain ? nullptr : W + w.g[i - 1]; g.in_ab = W + w.cp[i - 1]; g.in_h = (i - 1 >= 1 && !M.chain) ? W + w.h[i - 1] : nullptr;
    g.out_dz = (i - 1 >= 1 && !M.chain) ? W + w.dh[i - 1] : nullptr; g.out_dp = W + w.dp[i - 1];
    g.out_dzu = (M.bounded && i - 1 >= 1) ? W + w.dhu[i - 1] : nullptr;
    g.out_q = M.gabor ? W + w.q[i - 1] : nullptr;
    g.scal = reinterpret_cast<const float*>(W + w.scal); g.src_layer = i; g.dst_layer = i - 1;
    g.feat_tile_bytes = wtile; g.bs = static_cast<int>(bs); g.out_f = M.out_f;
    if (dout && M.stage_head[i - 1] >= 0) {
      int kk = -1, seen = 0;
      for (int k = 0; k < M.n_heads; ++k) if (M.head_live[k]) { if (seen == M.stage_head[i - 1]) { kk = k; break; } ++seen; }
      g.head_dout = dout; g.head_w = params + M.head_w[kk]; g.head_col = M.stage_head[i - 1] * M.out_f; g.head_ld = M.n_out * M.out_f;
    }
    if (M.bounded && i - 1 >= 1) { g.dist = dist; g.dist_row_offset = row_off; g.bound_lo = M.bound_lo[i - 1]; g.bound_hi = M.bound_hi[i - 1]; }
    e = launch_lgemm((g.trace = lgemm_trace_ptr(), g.dbg = lgemm_dbg(), g), p->n_sm, st);
    if (e != cudaSuccess) return cuda_fail(e, "lgemm_kernel(mfn dgrad)");
  }
  WgradArgs wg; std::memset(&wg, 0, sizeof(wg)); wg.trace = g_trace;
  wg.n_units = static_cast<int>(p->units.size());
  for (int i = 0; i < wg.n_units; ++i) {
    WgradUnit u = p->units[i];
    const int code = p->unit_layer[i];
    if (code >= 700) { u.a_off = w.dp[code - 700]; u.b_off = w.z[code - 700 - 1]; }      // wide chain: dW_s = DP[s]^T Z[s-1]
    else if (code >= 600) { u.a_off = w.q[code - 600]; u.b_off = w.xa; }
    else if (code >= 500) { u.a_off = w.q[code - 500]; u.b_off = w.x; }
    else if (code >= 400) { u.a_off = w.dhu[code - 400]; u.b_off = w.ones; }
    else if (code >= 300) { const int k = code - 300; u.a_off = w.z[M.head_stage[k]]; u.b_off = w.dout[M.stage_head[M.head_stage[k]]]; }
    else if (code >= 200) { u.a_off = w.dp[code - 200]; u.b_off = w.x; }
    else { const int s = code - 100; u.a_off = w.dh[s]; u.b_off = w.z[s - 1]; }
    wg.u[i] = u;
  }
  fill_wgrad_sched(p, wg);
  wg.n_split = w.n_split; wg.n_tiles = w.n_tiles; wg.n_params = gpart_stride(M.g_floats); wg.ws = W; wg.gpart_off = w.gpart;
  e = launch_wgrad(wg, st);
  if (e != cudaSuccess) return cuda_fail(e, "wgrad_kernel(mfn)");
  if (M.gabor) {
    e = launch_mfn_gabor_grad(x, st);
    if (e != cudaSuccess) return cuda_fail(e, "mfn_gabor_grad_kernel");
  }
  return INR_OK;
}

extern "C" int inr_pack_weights(const inr_plan* p, const float* params, void* wpack, void* stream) {
  if (!p || !params || !wpack) return fail(INR_EINVAL, "null argument");
  if (p->is_wire) {
    WireAdamArgs wa; wire_adam_fill(p, wa);
    wa.params = const_cast<float*>(params); wa.wpack = static_cast<uint8_t*>(wpack); wa.pack_only = 1;
    cudaError_t we = (p->wm.nlin == 2 ? launch_w2d_adam(wa, static_cast<cudaStream_t>(stream)) : launch_wire_adam(wa, static_cast<cudaStream_t>(stream)));
    return we == cudaSuccess ? INR_OK : cuda_fail(we, "wire_adam_kernel(pack)");
  }
  AdamArgs a; fill_adam(p, a);
  a.params = const_cast<float*>(params); a.wpack = static_cast<uint8_t*>(wpack);
  cudaError_t e = launch_pack(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "pack_kernel");
}

static int run_forward(const inr_plan* p, const Workspace& w, const LossDesc& loss, const float* params, const void* wpack,
                       const float* coords, const float* x, const float* encB, const float* gt, const uint8_t* mask,
                       int64_t bs, void* ws, float* out, int train, const int* row_off, int* step, cudaStream_t st) {
  NvtxRange nv("inr.chain.forward+loss");
  FwdArgs f{};
  f.m = p->model; f.w = w; f.loss = loss;
  f.params = params; f.wpack = static_cast<const uint8_t*>(wpack);
  f.coords = coords; f.x = x; f.encB = encB; f.gt = gt; f.mask = mask; f.out = out;
  f.ws = static_cast<uint8_t*>(ws); f.row_offset = row_off; f.step_counter = step;
  f.bs = static_cast<int>(bs); f.train = train; f.trace = g_trace;
  cudaError_t e = w.tile_rows == kTileM ? launch_chain_fwd(f, p->n_sm, st) : launch_chain_fwd_t(f, p->n_sm, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "chain_fwd_kernel");
}

extern "C" int inr_forward(const inr_plan* p, const float* params, const void* wpack, const float* input, const float* encB,
                           int64_t bs, void* workspace, float* out, int32_t train, void* stream) {
  if (!p || !params || !wpack || !input || !out || bs <= 0) return fail(INR_EINVAL, "bad argument");
  if (train && !workspace) return fail(INR_EINVAL, "training forward needs a workspace");
  if (p->is_mfn) {
    if (!workspace) return fail(INR_EINVAL, "MFN forward streams its activations through the workspace");
    const bool mg = p->mm.input_kind != INPUT_DENSE;       // gauss / LogF: `input` = coords, encoded in-kernel
    if (mg && !encB) return fail(INR_EINVAL, "gauss / LogF encoder needs encB");
    const MfnWorkspace mw = mfn_workspace(p, bs);
    return mfn_forward_impl(p, mw, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, mg ? input : nullptr, mg ? nullptr : input, encB,
                            nullptr, nullptr, nullptr, bs, workspace, out, train ? 1 : 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
  }
  if (p->is_wire) {
    if (!workspace) return fail(INR_EINVAL, "WIRE forward streams its activations through the workspace");
    const WireWorkspace ww = wire_workspace(p, bs);
    return wire_forward_impl(p, ww, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, input, nullptr, nullptr, bs, workspace, out,
                             train ? 1 : 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
  }
  const bool gauss = p->model.input_kind == INPUT_GAUSS;
  if (gauss && !encB) return fail(INR_EINVAL, "gauss encoder needs encB");
  Workspace w = plan_workspace(p, bs);
  LossDesc none{LOSS_NONE, 0.f, 0.f, 0.f};
  return run_forward(p, w, none, params, wpack, gauss ? input : nullptr, gauss ? nullptr : input, encB, nullptr, nullptr, bs,
                     workspace, out, train ? 1 : 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int inr_forward_dist(const inr_plan* p, const float* params, const void* wpack, const float* input, const float* encB,
                                const float* dist, int64_t bs, void* workspace, float* out, int32_t train, void* stream) {
  if (!p || !params || !wpack || !input || !out || !workspace || bs <= 0) return fail(INR_EINVAL, "bad argument");
  if (!p->is_mfn) return fail(INR_EINVAL, "inr_forward_dist is for the multiscale MFN models");
  const bool mg = p->mm.input_kind != INPUT_DENSE;
  if (mg && !encB) return fail(INR_EINVAL, "gauss encoder needs encB");
  const MfnWorkspace mw = mfn_workspace(p, bs);
  return mfn_forward_impl(p, mw, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, mg ? input : nullptr, mg ? nullptr : input, encB,
                          nullptr, nullptr, dist, bs, workspace, out, train ? 1 : 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int inr_backward_dist(const inr_plan* p, const float* params, const void* wpack, const float* dout, const float* dist,
                                 int64_t bs, void* workspace, float* grads, void* stream) {
  if (!p || !params || !wpack || !dout || !workspace || !grads || bs <= 0) return fail(INR_EINVAL, "bad argument");
  if (!p->is_mfn) return fail(INR_EINVAL, "inr_backward_dist is for the multiscale MFN models");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const MfnWorkspace mw = mfn_workspace(p, bs);
  int rcm = mfn_backward_impl(p, mw, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, dout, dist, bs, workspace, nullptr, nullptr, st);
  if (rcm) return rcm;
  AdamArgs ma; fill_adam(p, ma);
  ma.n_split = mw.n_split; ma.n_tiles = mw.n_tiles;
  ma.params = const_cast<float*>(params); ma.grads = grads;
  ma.gpart = reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.gpart);
  ma.gfin = p->mm.gabor ? reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.gfin) : nullptr;
  ma.scal = reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.scal);
  ma.do_adam = 0;
  cudaError_t me = launch_adam(ma, st);
  return me == cudaSuccess ? INR_OK : cuda_fail(me, "adam_kernel(reduce, mfn)");
}

static int run_backward(const inr_plan* p, const Workspace& w, const LossDesc& loss, const float* params, const void* wpack,
                        const float* dout, int64_t bs, void* ws, cudaStream_t st, cudaEvent_t mid = nullptr,
                        const float* hyper = nullptr, const int* step = nullptr) {
  NvtxRange nv("inr.chain.backward(dgrad+wgrad)");
  BwdArgs b{};
  b.hyper = hyper; b.step = step;
  b.m = p->model; b.w = w; b.loss = loss;
  b.params = params; b.wpack = static_cast<const uint8_t*>(wpack); b.dout = dout;
  b.ws = static_cast<uint8_t*>(ws); b.bs = static_cast<int>(bs); b.bs_k = static_cast<int>(bs);
  cudaError_t e = w.tile_rows == kTileM ? launch_chain_bwd(b, p->n_sm, st) : launch_chain_bwd_t(b, p->n_sm, st);
  if (e != cudaSuccess) return cuda_fail(e, "chain_bwd_kernel");
  if (mid) cudaEventRecord(mid, st);
  WgradArgs g; fill_wgrad(p, w, static_cast<uint8_t*>(ws), g); g.trace = g_trace;
  g.l2_hints = 3;        // dZ and H images are read by this kernel last: their lines leave L2 first (bs 100 000: wgrad 88.2 -> 82.6 us; no effect at 10 000, -0.8 % at 300 000)
  e = launch_wgrad(g, st);
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "wgrad_kernel");
}

extern "C" int inr_backward(const inr_plan* p, const float* params, const void* wpack, const float* dout, int64_t bs,
                            void* workspace, float* grads, void* stream) {
  if (!p || !params || !wpack || !dout || !workspace || !grads || bs <= 0) return fail(INR_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->is_mfn) {
    const MfnWorkspace mw = mfn_workspace(p, bs);
    int rcm = mfn_backward_impl(p, mw, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, dout, nullptr, bs, workspace, nullptr, nullptr, st);
    if (rcm) return rcm;
    AdamArgs ma; fill_adam(p, ma);
    ma.n_split = mw.n_split; ma.n_tiles = mw.n_tiles;
    ma.params = const_cast<float*>(params); ma.grads = grads;
    ma.gpart = reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.gpart);
  ma.gfin = p->mm.gabor ? reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.gfin) : nullptr;
    ma.gfin = p->mm.gabor ? reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.gfin) : nullptr;
    ma.scal = reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + mw.scal);
    ma.do_adam = 0;
    cudaError_t me = launch_adam(ma, st);
    return me == cudaSuccess ? INR_OK : cuda_fail(me, "adam_kernel(reduce, mfn)");
  }
  if (p->is_wire) {
    const WireWorkspace ww = wire_workspace(p, bs);
    int rcw = wire_backward_impl(p, ww, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, dout, bs, workspace, nullptr, nullptr, st);
    if (rcw) return rcw;
    WireAdamArgs wa; wire_adam_fill(p, wa);
    wa.n_split = ww.n_split; wa.params = const_cast<float*>(params); wa.grads = grads;
    wa.gpart = reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + ww.gpart);
    wa.scal = reinterpret_cast<const float*>(static_cast<uint8_t*>(workspace) + ww.scal);
    cudaError_t we = p->wm.nlin == 2 ? launch_w2d_adam(wa, st) : launch_wire_adam(wa, st);
    return we == cudaSuccess ? INR_OK : cuda_fail(we, "wire_adam_kernel(reduce)");
  }
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
  if (p->is_wire) {
    WireAdamArgs wa; wire_adam_fill(p, wa);
    wa.params = params; wa.mom = m; wa.var = v; wa.wpack = static_cast<uint8_t*>(wpack); wa.gpart = grads;
    wa.hyper = hyper_dev; wa.step = step_dev; wa.do_adam = 1;
    cudaError_t we = p->wm.nlin == 2 ? launch_w2d_adam_flat(wa, static_cast<cudaStream_t>(stream)) : launch_wire_adam_flat(wa, static_cast<cudaStream_t>(stream));
    return we == cudaSuccess ? INR_OK : cuda_fail(we, "wire_adam_flat_kernel");
  }
  AdamArgs a; fill_adam(p, a);
  a.n_split = 1; a.params = params; a.m = m; a.v = v; a.wpack = static_cast<uint8_t*>(wpack);
  a.gpart = grads; a.scal = nullptr; a.hyper = hyper_dev; a.step = step_dev; a.do_adam = 1;
  cudaError_t e = launch_adam(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "adam_kernel");
}

extern "C" int inr_adam_step_peers(const inr_plan* p, float* params, const float* const* peer_grads, uint32_t* const* peer_flags,
                                   int32_t n_ranks, int32_t rank, float* m, float* v, void* wpack, const float* hyper_dev,
                                   int32_t* step_dev, void* stream) {
  if (!p || !params || !peer_grads || !peer_flags || !m || !v || !wpack || !hyper_dev || !step_dev) return fail(INR_EINVAL, "null argument");
  if (n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks) return fail(INR_EINVAL, "bad rank / world size (at most 8 ranks)");
  PeerArgs P{};
  P.n_ranks = n_ranks; P.rank = rank;
  {
    static unsigned long long tmo = 0;
    if (!tmo) { const char* e = std::getenv("INR_PEER_TIMEOUT_S"); const double s = e ? std::atof(e) : 120.0; tmo = static_cast<unsigned long long>((s > 0 ? s : 120.0) * 1e9); }
    P.timeout_ns = tmo;
  }
  for (int q = 0; q < n_ranks; ++q) {
    if (!peer_grads[q] || !peer_flags[q]) return fail(INR_EINVAL, "null peer pointer");
    P.grads[q] = peer_grads[q]; P.flags[q] = peer_flags[q];
  }
  P.done = peer_flags[rank] + 32;      // word 32 of this rank's own flag block: finished-CTA counter of the optimiser kernel
  if (p->is_wire) {
    WireAdamArgs wa; wire_adam_fill(p, wa);
    wa.params = params; wa.mom = m; wa.var = v; wa.wpack = static_cast<uint8_t*>(wpack); wa.gpart = peer_grads[rank];
    wa.hyper = hyper_dev; wa.step = step_dev; wa.do_adam = 1; wa.peer = P;
    cudaError_t we = p->wm.nlin == 2 ? launch_w2d_adam_flat(wa, static_cast<cudaStream_t>(stream)) : launch_wire_adam_flat(wa, static_cast<cudaStream_t>(stream));
    return we == cudaSuccess ? INR_OK : cuda_fail(we, "wire_adam_flat_kernel(peers)");
  }
  AdamArgs a; fill_adam(p, a);
  a.n_split = 1; a.params = params; a.m = m; a.v = v; a.wpack = static_cast<uint8_t*>(wpack);
  a.gpart = peer_grads[rank]; a.scal = nullptr; a.hyper = hyper_dev; a.step = step_dev; a.do_adam = 1; a.peer = P;
  cudaError_t e = launch_adam(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? INR_OK : cuda_fail(e, "adam_kernel(peers)");
}

static int train_step_impl(const inr_plan* p, const inr_loss_desc* loss, float* params, float* m, float* v, void* wpack,
                           const float* hyper_dev, int32_t* step_dev, const float* coords, const float* input_x,
                           const float* encB, const float* gt, const uint8_t* mask, int64_t bs, int32_t* row_cursor_dev,
                           void* workspace, float* out, float* loss_out_dev, cudaStream_t st, cudaEvent_t* ev,
                           float* grads_only = nullptr, const float* dist = nullptr) {
  const bool no_adam = grads_only != nullptr;
  NvtxRange nv_step(no_adam ? "inr.grad_step" : "inr.train_step");
  if (!p || !loss || !params || !wpack || !gt || !workspace || bs <= 0) return fail(INR_EINVAL, "bad argument");
  if (!no_adam && (!m || !v || !hyper_dev || !step_dev)) return fail(INR_EINVAL, "bad argument");
  if (p->is_mfn) {
    const bool mg = p->mm.input_kind != INPUT_DENSE;
    if (mg && (!coords || !encB)) return fail(INR_EINVAL, "gauss / LogF encoder needs coords and encB");
    if (!mg && !input_x) return fail(INR_EINVAL, "dense input needs input_x");
    const bool multi = p->mm.n_out != 1;
    if (multi) {      // fused multi-scale step (reference src/train_kspace_multiscale.py:173-190)
      if (p->mm.n_out > 4 || p->mm.out_f != 2) return fail(INR_EUNSUPPORTED, "the fused multi-head loss handles up to 4 heads of 2 outputs");
      if (loss->kind != INR_LOSS_L2 && loss->kind != INR_LOSS_L1 && loss->kind != INR_LOSS_MSLE && loss->kind != INR_LOSS_LSL)
        return fail(INR_EUNSUPPORTED, "fused multi-head losses: L2, L1, MSLE, LSL (+ consistency); others run through the autograd face");
      if (loss->tv_weight > 0.f) return fail(INR_EUNSUPPORTED, "the TV term of the multi-scale loop runs through the autograd face");
      if (loss->dp_norm) return fail(INR_EUNSUPPORTED, "data-parallel normalisers are not wired into the multi-head loss");
      if ((loss->cons_weight != 0.f || p->mm.bounded) && !dist) return fail(INR_EINVAL, "the multi-scale step needs dist_to_center");
    }
    if (loss->kind < INR_LOSS_L2 || loss->kind > INR_LOSS_HDR) return fail(INR_EINVAL, "unknown loss kind");
    if ((loss->kind == INR_LOSS_HDR || loss->kind == INR_LOSS_LSL) && p->mm.out_f != 2)
      return fail(INR_EINVAL, "complex-valued losses need network_output_size == 2");
    const MfnWorkspace mw = mfn_workspace(p, bs);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    LossDesc ML{loss->kind, loss->hdr_eps, loss->hdr_sigma, loss->hdr_factor, loss->tv_weight, loss->tv_h, loss->tv_w,
                loss->dp_norm, loss->dp_rows, row_cursor_dev};
    if (multi) {
      ML.cons_weight = loss->cons_weight;
      for (int i = 0; i < 8; ++i) { ML.cons_lo[i] = loss->cons_bounds[2 * i]; ML.cons_hi[i] = loss->cons_bounds[2 * i + 1]; }
    }
    if (ev) cudaEventRecord(ev[0], st);
    int rcm = mfn_forward_impl(p, mw, ML, params, wpack, coords, input_x, encB, gt, mask, dist, bs, workspace, out, 1,
                               row_cursor_dev, no_adam ? nullptr : step_dev, st);
    if (rcm) return rcm;
    rcm = run_tv(loss, out, p->mm.out_f, bs, wsb, mw.gl, mw.part, mw.n_tiles, st);
    if (rcm) return rcm;
    if (ev) { cudaEventRecord(ev[1], st); cudaEventRecord(ev[2], st); }
    if (multi)        // the fused loss left fp32 dL/dy of every head: from here on the path of an external dL/dout
      rcm = mfn_backward_impl(p, mw, LossDesc{LOSS_NONE, 0.f, 0.f, 0.f}, params, wpack, reinterpret_cast<const float*>(wsb + mw.dyf), dist, bs,
                              workspace, no_adam ? nullptr : hyper_dev, no_adam ? nullptr : step_dev, st, row_cursor_dev, true);
    else
      rcm = mfn_backward_impl(p, mw, ML, params, wpack, nullptr, dist, bs, workspace, no_adam ? nullptr : hyper_dev,
                              no_adam ? nullptr : step_dev, st, row_cursor_dev);
    if (rcm) return rcm;
    if (ev) cudaEventRecord(ev[3], st);
    AdamArgs ma; fill_adam(p, ma);
    ma.n_split = mw.n_split; ma.n_tiles = mw.n_tiles;
    ma.params = params; ma.m = m; ma.v = v; ma.wpack = static_cast<uint8_t*>(wpack);
    ma.gpart = reinterpret_cast<const float*>(wsb + mw.gpart); ma.scal = reinterpret_cast<const float*>(wsb + mw.scal);
    ma.gfin = p->mm.gabor ? reinterpret_cast<const float*>(wsb + mw.gfin) : nullptr;
    ma.hyper = hyper_dev; ma.step = step_dev; ma.loss_out = loss_out_dev;
    ma.row_offset = row_cursor_dev; ma.row_advance = static_cast<int>(bs);
    ma.do_adam = no_adam ? 0 : 1; ma.scal_has_bc = no_adam ? 0 : 1; ma.grads = grads_only;
    cudaError_t me = launch_adam(ma, st);
    if (ev) { cudaEventRecord(ev[4], st); cudaEventRecord(ev[5], st); cudaEventRecord(ev[6], st); }
    return me == cudaSuccess ? INR_OK : cuda_fail(me, "adam_kernel(mfn)");
  }
  if (p->is_wire) {
    if (!coords) return fail(INR_EINVAL, "WIRE needs coords");
    if (loss->kind < INR_LOSS_L2 || loss->kind > INR_LOSS_HDR) return fail(INR_EINVAL, "unknown loss kind");
    if ((loss->kind == INR_LOSS_HDR || loss->kind == INR_LOSS_LSL) && p->wm.out_f != 2)
      return fail(INR_EINVAL, "complex-valued losses need network_output_size == 2");
    const WireWorkspace ww = wire_workspace(p, bs);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    const LossDesc WL{loss->kind, loss->hdr_eps, loss->hdr_sigma, loss->hdr_factor, loss->tv_weight, loss->tv_h, loss->tv_w,
                      loss->dp_norm, loss->dp_rows, row_cursor_dev};
    if (ev) cudaEventRecord(ev[0], st);
    // without a TV pass between forward and backward the step scalars are reduced by the last CTA of wire_last
    const bool fold = p->wm.nlin == 1 && !(loss->tv_weight > 0.f);
    int rcw = wire_forward_impl(p, ww, WL, params, wpack, coords, gt, mask, bs, workspace, out, 1, row_cursor_dev,
                                no_adam ? nullptr : step_dev, st, ev ? ev + 5 : nullptr, fold, no_adam ? nullptr : hyper_dev);
    if (rcw) return rcw;
    rcw = run_tv(loss, out, p->wm.out_f, bs, wsb, ww.g, ww.part, ww.n_tiles, st);
    if (rcw) return rcw;
    if (ev) { cudaEventRecord(ev[1], st); cudaEventRecord(ev[2], st); }
    rcw = wire_backward_impl(p, ww, WL, params, wpack, nullptr, bs, workspace, no_adam ? nullptr : hyper_dev,
                             no_adam ? nullptr : step_dev, st, fold);
    if (rcw) return rcw;
    if (ev) cudaEventRecord(ev[3], st);
    WireAdamArgs wa; wire_adam_fill(p, wa);
    wa.n_split = ww.n_split; wa.params = params; wa.mom = m; wa.var = v; wa.wpack = static_cast<uint8_t*>(wpack);
    wa.gpart = reinterpret_cast<const float*>(wsb + ww.gpart); wa.scal = reinterpret_cast<const float*>(wsb + ww.scal);
    wa.hyper = hyper_dev; wa.step = step_dev; wa.loss_out = loss_out_dev;
    wa.row_offset = row_cursor_dev; wa.row_advance = static_cast<int>(bs);
    wa.do_adam = no_adam ? 0 : 1; wa.scal_has_bc = no_adam ? 0 : 1; wa.grads = grads_only;
    wa.params_stable = 1;     // the forward pass of this very call read them: complete since before the step's first kernel
    cudaError_t we = p->wm.nlin == 2 ? launch_w2d_adam(wa, st) : launch_wire_adam(wa, st);
    if (ev) cudaEventRecord(ev[4], st);
    return we == cudaSuccess ? INR_OK : cuda_fail(we, "wire_adam_kernel");
  }
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
  LossDesc L{loss->kind, loss->hdr_eps, loss->hdr_sigma, loss->hdr_factor, loss->tv_weight, loss->tv_h, loss->tv_w,
             loss->dp_norm, loss->dp_rows, row_cursor_dev};
  if (ev) cudaEventRecord(ev[0], st);
  int rc = run_forward(p, w, L, params, wpack, coords, input_x, encB, gt, mask, bs, workspace, out, 1, row_cursor_dev,
                       no_adam ? nullptr : step_dev, st);
  if (rc) return rc;
  rc = run_tv(loss, out, p->model.out_f, bs, ws, w.g_off, w.part_off, static_cast<int>((bs + kTileM - 1) / kTileM), st);
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

extern "C" int inr_train_step_dist(const inr_plan* p, const inr_loss_desc* loss, float* params, float* m, float* v, void* wpack,
                                   const float* hyper_dev, int32_t* step_dev, const float* coords, const float* input_x,
                                   const float* encB, const float* gt, const uint8_t* mask, const float* dist, int64_t bs,
                                   int32_t* row_cursor_dev, void* workspace, float* out, float* loss_out_dev, void* stream) {
  if (p && !p->is_mfn) return fail(INR_EINVAL, "inr_train_step_dist is for the multiscale MFN models");
  return train_step_impl(p, loss, params, m, v, wpack, hyper_dev, step_dev, coords, input_x, encB, gt, mask, bs,
                         row_cursor_dev, workspace, out, loss_out_dev, static_cast<cudaStream_t>(stream), nullptr, nullptr, dist);
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

extern "C" int inr_grad_step_dist(const inr_plan* p, const inr_loss_desc* loss, const float* params, const void* wpack,
                                  const float* coords, const float* input_x, const float* encB, const float* gt,
                                  const uint8_t* mask, const float* dist, int64_t bs, int32_t* row_cursor_dev, void* workspace,
                                  float* out, float* grads, float* loss_out_dev, void* stream) {
  if (!grads) return fail(INR_EINVAL, "grads must not be null");
  if (p && !p->is_mfn) return fail(INR_EINVAL, "inr_grad_step_dist is for the multiscale MFN models");
  return train_step_impl(p, loss, const_cast<float*>(params), nullptr, nullptr, const_cast<void*>(wpack), nullptr, nullptr,
                         coords, input_x, encB, gt, mask, bs, row_cursor_dev, workspace, out, loss_out_dev,
                         static_cast<cudaStream_t>(stream), nullptr, grads, dist);
}

extern "C" int inr_profile_step(const inr_plan* p, const inr_loss_desc* loss, float* params, float* m, float* v, void* wpack,
                                const float* hyper_dev, int32_t* step_dev, const float* coords, const float* input_x,
                                const float* encB, const float* gt, const uint8_t* mask, int64_t bs, void* workspace,
                                float* out, int32_t reps, float* ms_out4, void* stream) {
  if (!ms_out4 || reps <= 0) return fail(INR_EINVAL, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t ev[7];
  for (auto& e : ev) if (cudaEventCreate(&e) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaEventCreate");
  double acc[5] = {0, 0, 0, 0, 0};
  int rc = INR_OK;
  for (int r = 0; r < reps && rc == INR_OK; ++r) {
    rc = train_step_impl(p, loss, params, m, v, wpack, hyper_dev, step_dev, coords, input_x, encB, gt, mask, bs, nullptr,
                         workspace, out, nullptr, st, ev);
    if (rc) break;
    cudaError_t e = cudaEventSynchronize(ev[4]);
    if (e != cudaSuccess) { rc = cuda_fail(e, "profile step"); break; }
    for (int k = 0; k < 4; ++k) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); acc[k] += ms; }
    if (p->is_wire) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[5], ev[6]); acc[4] += ms; }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int k = 0; k < 4; ++k) ms_out4[k] = static_cast<float>(acc[k] / reps);
  ms_out4[4] = static_cast<float>(acc[4] / reps);      // WIRE: the `depth` forward layer-GEMM launches together
  return rc;
}
