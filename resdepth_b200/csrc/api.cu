// C ABI of the B200-native ResDepth hot path (see include/resdepth_b200.h): the layer plan of
// lib.UNet.UNet (reference lib/UNet.py:104-246), workspace management, and the forward / backward schedules.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/resdepth_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rd {

thread_local std::string g_last_error;
long long g_launch_count = 0;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return 1;
}

struct ParamInfo {
  std::string name;
  long long numel;
  long long offset;   // in floats, 16-byte aligned
};

struct ConvBlock {     // conv3x3 (+BN) + activation (+pool): conv_block / bottleneck / decoder conv (lib/UNet.py:36-93)
  int Cin = 0, Cout = 0;
  int act = RD_ACT_RELU;
  bool pool = false;
  long long w = -1, bias = -1, gamma = -1, beta = -1, slope = -1;   // param arena offsets
  long long rm = -1, rv = -1;                                       // buffer arena offsets
  // workspace
  float *z = nullptr, *a = nullptr, *p = nullptr;
  float *mean = nullptr, *invstd = nullptr, *scale = nullptr, *shift = nullptr;
  float *w_kn = nullptr, *w_nk = nullptr, *wd_kn = nullptr, *wd_nk = nullptr;
  bool tc = false;     // GEMMs of this block run on tcgen05
  bool bb = false;     // backward GEMMs of this block take bf16 operands (dz, block input, dgrad weights)
  int wide = 0;        // weight gradient through the wide-N reduce plan: 1 = rows are ci, 2 = rows are co
  void* gyb = nullptr; // bf16 dz buffer of this block (gy_b or gy_b2)
  int par = 0;         // its parity
  TcRowsPlan tc_fwd, tc_dgrad;
  TcReducePlan tc_wgrad;
  // bf16 shadows (uint16 storage): a_b / p_b copies of a / p for consumers' bf16 GEMMs, dgrad weights
  void *a_b = nullptr, *p_b = nullptr, *wd_nk_b = nullptr;
};

struct UpConv {        // ConvTranspose2d(C, C, 2, 2), or Upsample(bilinear, x2) + Conv2d 1x1 (lib/UNet.py:17-24)
  int C = 0;
  bool bilinear = false;
  long long w = -1, bias = -1;
  float *u = nullptr;
  float *t = nullptr;      // bilinear: 1x1-conv output at the low resolution
  // transposed: w_kn = [ci][(ab,co)], w_nk = its transpose;  bilinear: w_kn = W^T = [ci][co], w_nk = W = [co][ci]
  float *w_kn = nullptr, *w_nk = nullptr;
  bool tc = false;
  bool bb = false;     // bf16 backward GEMMs
  TcRowsPlan tc_fwd, tc_dgrad;
  TcReducePlan tc_wgrad;
  void *u_b = nullptr, *w_kn_b = nullptr;   // bf16 copy of u (next conv's wgrad operand), bf16 dgrad weights
};

}  // namespace rd

using namespace rd;

namespace rd {
// Everything that depends on the (batch, tile, with_backward) layout of the workspace: the slab, the pointers carved
// out of it (inside the per-layer structs too) and the TMA descriptors / tile plans built on those pointers.  A handle
// keeps the layouts it has used (rd_reserve switches between them without synchronising or rebuilding anything).
struct Workspace {
  std::vector<ConvBlock> enc, dec;   // dec has depth-1 entries
  ConvBlock bott;
  std::vector<UpConv> ups;           // depth entries
  void* slab = nullptr;
  size_t slab_bytes = 0;
  int res_batch = 0, res_tile = 0, res_bwd = 0;
  long long ws_id = 0, ws_last_use = 0;
  float *partials = nullptr, *scratch = nullptr, *part = nullptr, *coef = nullptr, *consts = nullptr;
  size_t partials_floats = 0, scratch_floats = 0, part_floats = 0;
  std::vector<float*> g_skip;
  float *gy = nullptr, *gh = nullptr, *gp = nullptr;
  void* gp_b = nullptr;        // bf16 gradient at a pooled tensor (bf16 backward)
  void* gh_b = nullptr;        // bf16 gradient at an up-conv input (= output of the block before it), bf16 backward
  float* xcol = nullptr;       // im2col expansion of the input for the first layer's tensor-core wgrad
  int xcol_k = 0;
  float* gt = nullptr;         // bilinear up-mode: gradient at the low-resolution 1x1-conv output
  void* gy_b = nullptr;
  void* gy_b2 = nullptr;       // second dz buffer: consecutive blocks alternate so a weight gradient on the side
                               // stream can still read block k's dz while block k+1's BatchNorm backward writes its own
  std::vector<void*> gs_b;
  // weight gradients overlap the rest of the backward pass on a side stream (tensor-bound GEMMs next to the
  // HBM-bound BatchNorm backward kernels); only when every GEMM of the backward runs the bf16 tcgen05 path
  bool overlap = false;
  void* xcol_b = nullptr;
  int xcol_b_k = 0;            // operand width of the expansion in the reduce GEMMs (64 / 128)
  int xcol_b_pitch = 0;        // stored columns per pixel: 32 when the live columns fit (the boxes zero-fill the rest)
  float *ob_mean = nullptr, *ob_invstd = nullptr, *ob_affine = nullptr;
  float* taps = nullptr;       // last-conv forward: 9 tap partial sums per pixel
  float* gram = nullptr;       // first encoder block: Gram matrix of the bf16 im2col expansion [Kc][Kc]
  TcReducePlan tc_gram;        // reduce GEMM (xcol, xcol) that produces it
  bool gram_fused = false;     // the first layer's weight-gradient GEMM also produces the Gram matrix (one pass over xcol)
  long long eval_pack_gen = 0; // rd_freeze_params generation whose weights / BatchNorm vectors this layout's packs hold
};
}  // namespace rd

struct rd_handle : rd::Workspace {
  rd_config cfg;
  int device = 0;
  int depth = 0;
  std::vector<int> widths;
  long long last_w = -1, last_b = -1;
  long long freeze_gen = 0;    // rd_freeze_params: generation counter, bumped by every switch-on and by rd_bind
  bool frozen = false;
  std::vector<ParamInfo> params, buffers;
  long long param_floats = 0, buffer_floats = 0;
  float *P = nullptr, *G = nullptr, *BUF = nullptr;   // bound arenas

  std::vector<rd::Workspace> ws_cache;   // layouts not in use right now (their slabs stay allocated)
  long long ws_next_id = 1, ws_clock = 0;
  // bf16 backward (rd_config.bwd_mode; default in TF32 mode): dz and the skip gradients as bf16 GEMM operands
  bool bwd_bf16 = false;
  bool overlap_allowed = true; // rd_set_overlap
  cudaStream_t side = nullptr;
  cudaEvent_t ev_main = nullptr, ev_join = nullptr, ev_wg[2] = {nullptr, nullptr};
  bool wg_pending[2] = {false, false};
  // outer_skip_BN: BatchNorm2d(1) on input channel 0
  long long ob_gamma = -1, ob_beta = -1, ob_rm = -1, ob_rv = -1;
  // state of the last forward
  int fwd_batch = 0, fwd_tile = 0, fwd_mode = -1;
  // staged backward (rd_backward_stage): next expected stage, and whether the pooled-tensor gradient is in gp_b
  int bw_next_stage = 0;
  bool bw_gp_bf16 = false, bw_gh_bf16 = false;
  bool xcol_early = false;     // the first layer's im2col expansion was enqueued at the start of this backward pass
  bool tf32() const { return cfg.math_mode == RD_MATH_TF32; }

  // per-category CUDA-event timing (rd_profile_*): off by default
  bool profiling = false;
  struct ProfRec { int cat; cudaEvent_t e0, e1; double flops, bytes; int launches; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[RD_PROF_NUM] = {0}, prof_flops[RD_PROF_NUM] = {0}, prof_bytes[RD_PROF_NUM] = {0};
  long long prof_launches[RD_PROF_NUM] = {0}, prof_calls[RD_PROF_NUM] = {0};
};

namespace {
const char* const kProfNames[RD_PROF_NUM] = {
    "conv3x3_fwd", "conv3x3_dgrad", "conv3x3_wgrad", "convT_fwd", "convT_dgrad", "convT_wgrad",
    "first_conv_fwd", "first_conv_wgrad", "last_conv_fwd", "last_conv_bwd", "bn_finalize", "bn_act_pool",
    "bn_bwd_reduce", "bn_bwd_apply", "pack_weights", "unpack_grads", "bias_grad", "loss"};

// brackets the launches of one category with events on the launching stream when profiling is on
struct ProfScope {
  rd_handle* h;
  cudaStream_t s;
  size_t idx = (size_t)-1;
  long long launches0 = 0;
  ProfScope(rd_handle* h_, int cat, double flops, double bytes, cudaStream_t s_) : h(h_), s(s_) {
    if (!h->profiling) return;
    cudaEvent_t ev[2];
    for (auto& e : ev) {
      if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
      else if (cudaEventCreate(&e) != cudaSuccess) return;
    }
    cudaEventRecord(ev[0], s);
    launches0 = rd::g_launch_count;
    h->prof.push_back({cat, ev[0], ev[1], flops, bytes, 0});
    idx = h->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx == (size_t)-1) return;
    cudaEventRecord(h->prof[idx].e1, s);
    h->prof[idx].launches = (int)(rd::g_launch_count - launches0);
  }
};
}  // namespace


namespace {

long long align4(long long v) { return (v + 3) & ~3LL; }

void add_param(rd_handle* h, const std::string& name, long long numel, long long* off_out) {
  ParamInfo p{name, numel, h->param_floats};
  if (off_out) *off_out = p.offset;
  h->param_floats = align4(h->param_floats + numel);
  h->params.push_back(p);
}
void add_buffer(rd_handle* h, const std::string& name, long long numel, long long* off_out) {
  ParamInfo p{name, numel, h->buffer_floats};
  if (off_out) *off_out = p.offset;
  h->buffer_floats = align4(h->buffer_floats + numel);
  h->buffers.push_back(p);
}

// registers the parameters of one conv block under `prefix` in nn.Module registration order
void plan_block(rd_handle* h, ConvBlock& b, const std::string& prefix, int Cin, int Cout, int act, bool pool) {
  b.Cin = Cin; b.Cout = Cout; b.act = act; b.pool = pool;
  add_param(h, prefix + ".0.weight", (long long)Cout * Cin * 9, &b.w);
  int idx = 1;
  if (h->cfg.do_bn) {
    add_param(h, prefix + ".1.weight", Cout, &b.gamma);
    add_param(h, prefix + ".1.bias", Cout, &b.beta);
    add_buffer(h, prefix + ".1.running_mean", Cout, &b.rm);
    add_buffer(h, prefix + ".1.running_var", Cout, &b.rv);
    idx = 2;
  } else {
    add_param(h, prefix + ".0.bias", Cout, &b.bias);
  }
  if (act == RD_ACT_PRELU) add_param(h, prefix + "." + std::to_string(idx) + ".weight", 1, &b.slope);
}

Gather gather_conv3x3(int H, int W, int C) {
  Gather g{};
  g.ntaps = 9; g.ups = 1; g.Hs = g.Ho = H; g.Ws = g.Wo = W; g.C = C;
  for (int t = 0; t < 9; ++t) { g.dh[t] = (signed char)(t / 3 - 1); g.dw[t] = (signed char)(t % 3 - 1); }
  return g;
}
Gather gather_plain(int H, int W, int C) {
  Gather g{};
  g.ntaps = 1; g.ups = 1; g.Hs = g.Ho = H; g.Ws = g.Wo = W; g.C = C;
  return g;
}
Gather gather_up2(int Hin, int Win, int C) {      // rows = input pixels, 4 taps into the 2x-upsampled tensor
  Gather g{};
  g.ntaps = 4; g.ups = 2; g.Hs = 2 * Hin; g.Ws = 2 * Win; g.Ho = Hin; g.Wo = Win; g.C = C;
  for (int t = 0; t < 4; ++t) { g.dh[t] = (signed char)(t >> 1); g.dw[t] = (signed char)(t & 1); }
  return g;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<char*>(b)) {}
  float* take(size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += (floats * sizeof(float) + 255) & ~size_t(255);
    return p;
  }
};

// lays out the workspace; with base == nullptr only measures it
size_t carve(rd_handle* h, void* base, int B, int T, int bwd) {
  Carver c(base);
  const int D = h->depth;
  const bool tf = h->tf32();
  size_t max_out = 0, max_pool = 0;
  const bool bf = bwd && h->bwd_bf16;
  auto block = [&](ConvBlock& b, int H, bool first) {
    const size_t n = (size_t)B * H * H * b.Cout;
    b.z = c.take(n);
    b.a = c.take(n);
    b.p = b.pool ? c.take(n / 4) : nullptr;
    b.a_b = (bf && !b.pool) ? c.take(n / 2 + 64) : nullptr;
    b.p_b = (bf && b.pool) ? c.take(n / 8 + 64) : nullptr;
    b.mean = c.take(b.Cout); b.invstd = c.take(b.Cout); b.scale = c.take(b.Cout); b.shift = c.take(b.Cout);
    if (!first) {
      const size_t wn = (size_t)9 * b.Cin * b.Cout;
      b.w_kn = c.take(wn);
      b.wd_kn = c.take(wn);
      b.w_nk = tf ? c.take(wn) : nullptr;
      b.wd_nk = tf ? c.take(wn) : nullptr;
      b.wd_nk_b = bf ? c.take(wn / 2 + 64) : nullptr;
    }
    if (n > max_out) max_out = n;
    if (b.pool && n / 4 > max_pool) max_pool = n / 4;
  };
  for (int i = 0; i < D; ++i) block(h->enc[i], T >> i, i == 0);
  block(h->bott, T >> D, false);
  for (int j = 0; j < D; ++j) {
    UpConv& u = h->ups[j];
    const int Hout = T >> (D - 1 - j);
    u.u = c.take((size_t)B * Hout * Hout * u.C);
    u.w_kn = c.take((size_t)4 * u.C * u.C);
    u.w_nk = c.take((size_t)4 * u.C * u.C);
    u.t = u.bilinear ? c.take((size_t)B * (Hout / 2) * (Hout / 2) * u.C) : nullptr;
    u.u_b = (bf && j < D - 1) ? c.take((size_t)B * Hout * Hout * u.C / 2 + 64) : nullptr;
    u.w_kn_b = bf ? c.take((size_t)4 * u.C * u.C / 2 + 64) : nullptr;
    if (j < D - 1) block(h->dec[j], Hout, false);
  }
  h->partials_floats = (size_t)2 << 20;
  h->partials = c.take(h->partials_floats);
  h->scratch_floats = (size_t)1 << 20;
  h->scratch = c.take(h->scratch_floats);
  h->coef = c.take(4 * 2048);
  h->consts = c.take(64);
  h->ob_mean = c.take(4); h->ob_invstd = c.take(4); h->ob_affine = c.take(4);
  h->taps = c.take((size_t)9 * B * T * T);               // tap partials of the last conv's forward (planar, 36 B / pixel)
  h->gram = c.take(128 * 128);
  if (bwd) {
    h->part_floats = (size_t)8 << 20;
    h->part = c.take(h->part_floats);
    h->g_skip.resize(D);
    for (int i = 0; i < D; ++i) h->g_skip[i] = c.take((size_t)B * (T >> i) * (T >> i) * h->enc[i].Cout);
    h->gy = c.take(max_out);
    h->gh = c.take(max_out);
    h->gp = c.take(max_pool);
    h->gt = h->cfg.up_mode == RD_UP_BILINEAR ? c.take(max_out / 4 + 64) : nullptr;
    h->gy_b = bf ? c.take(max_out / 2 + 64) : nullptr;
    h->gy_b2 = bf ? c.take(max_out / 2 + 64) : nullptr;
    h->gp_b = bf ? c.take(max_pool / 2 + 64) : nullptr;
    h->gh_b = bf ? c.take(max_out / 2 + 64) : nullptr;
    h->gs_b.assign(D, nullptr);
    if (bf)
      for (int i = 0; i < D; ++i) h->gs_b[i] = c.take((size_t)B * (T >> i) * (T >> i) * h->enc[i].Cout / 2 + 64);
    h->xcol_b_k = h->enc[0].Cin * 9 <= 64 ? 64 : 128;
    {
      static const bool xcol32 = getenv("RESDEPTH_XCOL64") == nullptr;
      h->xcol_b_pitch = (xcol32 && h->enc[0].Cin * 9 + 1 <= 32 && h->enc[0].Cin <= 3) ? 32 : h->xcol_b_k;
    }
    h->xcol_b = bf ? c.take((size_t)B * T * T * h->xcol_b_pitch / 2 + 64) : nullptr;
    h->xcol_k = (h->enc[0].Cin * 9 + 31) / 32 * 32;
    h->xcol = (tf && !bf) ? c.take((size_t)B * T * T * h->xcol_k) : nullptr;
  } else {
    h->xcol = nullptr;
    h->gt = nullptr;
    h->gy_b = nullptr; h->gy_b2 = nullptr; h->xcol_b = nullptr; h->gp_b = nullptr; h->gh_b = nullptr;
    h->gs_b.assign(D, nullptr);
    h->part = nullptr; h->part_floats = 0;
    h->g_skip.assign(D, nullptr);
    h->gy = h->gh = h->gp = nullptr;
  }
  return c.off;
}

// TMA descriptors + tile plans of every tcgen05 layer for the current workspace layout
int build_tc_plans(rd_handle* h, int B, int T, int bwd) {
  const int D = h->depth;
  auto off = [&](ConvBlock& b) { b.tc = b.bb = false; b.wide = 0; b.tc_fwd.valid = b.tc_dgrad.valid = b.tc_wgrad.valid = false; };
  for (auto& b : h->enc) off(b);
  off(h->bott);
  for (auto& b : h->dec) off(b);
  for (auto& u : h->ups) { u.tc = u.bb = false; u.tc_fwd.valid = u.tc_dgrad.valid = u.tc_wgrad.valid = false; }
  if (!h->tf32() || !tc_available()) return 0;
  const bool bf = bwd && h->gy_b != nullptr;            // bf16 operands for the backward GEMMs
  h->overlap = false;
  {
    // dz buffers alternate in the order the backward pass visits the blocks: dec[D-2..0], bottleneck, enc[D-1..0]
    int ord = 0;
    auto assign = [&](ConvBlock& b) { b.par = ord & 1; b.gyb = b.par ? h->gy_b2 : h->gy_b; ++ord; };
    for (int j = D - 2; j >= 0; --j) assign(h->dec[j]);
    assign(h->bott);
    for (int i = D - 1; i >= 0; --i) assign(h->enc[i]);
  }
  auto block = [&](ConvBlock& b, const float* src, const void* src_b, int H) -> int {
    Gather gf = gather_conv3x3(H, H, b.Cin);
    Gather gd = gather_conv3x3(H, H, b.Cout);
    if (!tc_rows_eligible(gf, b.Cout) || !tc_rows_eligible(gd, b.Cin)) return 0;
    RD_TRY(tc_make_rows_plan(&b.tc_fwd, src, gf, B, b.w_nk, b.Cout));
    if (bwd) {
      if (bf && src_b && b.wd_nk_b && tc_rows_eligible(gd, b.Cin, 1) && tc_reduce_eligible(gf, b.Cout, 1)) {
        RD_TRY(tc_make_rows_plan(&b.tc_dgrad, b.gyb, gd, B, b.wd_nk_b, b.Cin, 1));
        // narrow output tiles (Cout or Cin of 64 / 128) are L2-bound: use the wide-N formulation where it applies
        static const bool no_wide = getenv("RESDEPTH_NO_WIDE_WGRAD") != nullptr;
        b.wide = 0;
        if (!no_wide && b.Cout <= 128 && tc_reduce_wide_eligible(b.Cin, b.Cout, H, H)) {
          // D[ci][(t,co)] = sum_q x[q][ci] * dz[q - off(t)][co]
          if (tc_make_reduce_plan_wide(&b.tc_wgrad, src_b, b.Cin, b.gyb, b.Cout, -1, B, H, H, h->part, h->part_floats) == 0)
            b.wide = 1;
        } else if (!no_wide && b.Cin == 64 && tc_reduce_wide_eligible(b.Cout, b.Cin, H, H)) {
          // D[co][(t,ci)] = sum_p dz[p][co] * x[p + off(t)][ci]
          if (tc_make_reduce_plan_wide(&b.tc_wgrad, b.gyb, b.Cout, src_b, b.Cin, +1, B, H, H, h->part, h->part_floats) == 0)
            b.wide = 2;
        }
        if (!b.wide)
          RD_TRY(tc_make_reduce_plan(&b.tc_wgrad, src_b, gf, B, b.gyb, b.Cout, h->part, h->part_floats, 1));
        b.bb = true;
      } else {
        RD_TRY(tc_make_rows_plan(&b.tc_dgrad, h->gy, gd, B, b.wd_nk, b.Cin));
        if (tc_reduce_eligible(gf, b.Cout))
          RD_TRY(tc_make_reduce_plan(&b.tc_wgrad, src, gf, B, h->gy, b.Cout, h->part, h->part_floats));
      }
    }
    b.tc = true;
    return 0;
  };
  if (bwd && h->enc[0].Cout % 32 == 0) {
    ConvBlock& b0 = h->enc[0];
    Gather gb = gather_plain(T, T, h->xcol_b_k);
    h->tc_gram.valid = false;
    h->gram_fused = false;
    if (bf && h->xcol_b && tc_reduce_eligible(gb, b0.Cout, 1)) {
      const int xp = h->xcol_b_pitch != h->xcol_b_k ? h->xcol_b_pitch : 0;
      RD_TRY(tc_make_reduce_plan(&b0.tc_wgrad, h->xcol_b, gb, B, b0.gyb, b0.Cout, h->part, h->part_floats, 1, xp, 0));
      b0.bb = true;
      // Gram matrix of the expansion (same GEMM with the expansion as both operands): lets the first block's weight
      // gradient be formed from gY, without the apply pass of its BatchNorm backward (block_backward)
      static const bool no_gram = getenv("RESDEPTH_NO_GRAM") != nullptr;
      // ... and since both GEMMs have the expansion as their M operand, ONE GEMM with the N tile [xcol | gY] (the xcol box
      // in shared memory is both operands) yields G and X^T gY from a single pass over xcol: tc_reduce_plan_add_gram
      static const bool gram_separate = getenv("RESDEPTH_GRAM_SEPARATE") != nullptr;
      if (!no_gram && h->cfg.do_bn && b0.Cin * 9 < h->xcol_b_k && tc_reduce_eligible(gb, h->xcol_b_k, 1)) {
        if (!gram_separate && h->xcol_b_k == 64 && b0.Cout == 64 && b0.tc_wgrad.BN == 64 &&
            (size_t)b0.tc_wgrad.splits * 64 * 128 <= h->part_floats) {
          RD_TRY(tc_reduce_plan_add_gram(&b0.tc_wgrad, h->part_floats));
          h->gram_fused = true;
        } else {
          RD_TRY(tc_make_reduce_plan(&h->tc_gram, h->xcol_b, gb, B, h->xcol_b, h->xcol_b_k, h->part, h->part_floats, 1, xp, xp));
        }
      }
    } else if (h->xcol) {
      Gather g0 = gather_plain(T, T, h->xcol_k);
      if (tc_reduce_eligible(g0, b0.Cout))
        RD_TRY(tc_make_reduce_plan(&b0.tc_wgrad, h->xcol, g0, B, h->gy, b0.Cout, h->part, h->part_floats));
    }
  }
  for (int i = 1; i < D; ++i) RD_TRY(block(h->enc[i], h->enc[i - 1].p, h->enc[i - 1].p_b, T >> i));
  RD_TRY(block(h->bott, h->enc[D - 1].p, h->enc[D - 1].p_b, T >> D));
  for (int j = 0; j < D; ++j) {
    UpConv& u = h->ups[j];
    const int Hin = T >> (D - j);
    const float* src = j == 0 ? h->bott.a : h->dec[j - 1].a;
    Gather gf = gather_plain(Hin, Hin, u.C);
    Gather gd = gather_up2(Hin, Hin, u.C);
    if (u.bilinear) {
      if (tc_rows_eligible(gf, u.C)) {
        RD_TRY(tc_make_rows_plan(&u.tc_fwd, src, gf, B, u.w_nk, u.C));
        if (bwd) {
          RD_TRY(tc_make_rows_plan(&u.tc_dgrad, h->gt, gf, B, u.w_kn, u.C));
          if (tc_reduce_eligible(gf, u.C))
            RD_TRY(tc_make_reduce_plan(&u.tc_wgrad, src, gf, B, h->gt, u.C, h->part, h->part_floats));
        }
        u.tc = true;
      }
    } else if (tc_rows_eligible(gf, 4 * u.C) && tc_rows_eligible(gd, u.C)) {
      RD_TRY(tc_make_rows_plan(&u.tc_fwd, src, gf, B, u.w_nk, 4 * u.C));
      if (bwd) {
        const void* src_b = j == 0 ? h->bott.a_b : h->dec[j - 1].a_b;
        if (bf && src_b && u.w_kn_b && h->gs_b[D - 1 - j] && tc_rows_eligible(gd, u.C, 1) && tc_reduce_eligible(gd, u.C, 1)) {
          RD_TRY(tc_make_rows_plan(&u.tc_dgrad, h->gs_b[D - 1 - j], gd, B, u.w_kn_b, u.C, 1));
          RD_TRY(tc_make_reduce_plan(&u.tc_wgrad, h->gs_b[D - 1 - j], gd, B, src_b, u.C, h->part, h->part_floats, 1));
          u.bb = true;
        } else {
          RD_TRY(tc_make_rows_plan(&u.tc_dgrad, h->g_skip[D - 1 - j], gd, B, u.w_kn, u.C));
          if (tc_reduce_eligible(gd, u.C))
            RD_TRY(tc_make_reduce_plan(&u.tc_wgrad, h->g_skip[D - 1 - j], gd, B, src, u.C, h->part, h->part_floats));
        }
      }
      u.tc = true;
    }
    if (j < D - 1) RD_TRY(block(h->dec[j], u.u, u.tc && !u.bilinear ? u.u_b : nullptr, 2 * Hin));
  }
  // the bf16 copy of an up-conv's output gradient is written by the tcgen05 dgrad of the decoder conv after it
  // (level 0: by the last-conv backward); without that producer the up-conv falls back to TF32 operands
  for (int j = 0; j < D - 1; ++j) {
    UpConv& u = h->ups[j];
    if (u.bb && !h->dec[j].tc) {
      const int Hin = T >> (D - j);
      const float* src = j == 0 ? h->bott.a : h->dec[j - 1].a;
      Gather gd = gather_up2(Hin, Hin, u.C);
      u.bb = false;
      u.tc_wgrad.valid = false;
      RD_TRY(tc_make_rows_plan(&u.tc_dgrad, h->g_skip[D - 1 - j], gd, B, u.w_kn, u.C));
      if (tc_reduce_eligible(gd, u.C))
        RD_TRY(tc_make_reduce_plan(&u.tc_wgrad, h->g_skip[D - 1 - j], gd, B, src, u.C, h->part, h->part_floats));
    }
  }
  // side-stream weight gradients: every reduce GEMM must be a bf16 tcgen05 plan (they share the split buffer, which
  // then belongs to the side stream alone, and read only bf16 tensors that the main stream no longer rewrites)
  static const bool no_overlap = getenv("RESDEPTH_NO_OVERLAP") != nullptr;
  if (bwd && bf && !no_overlap && h->side) {
    bool all = true;
    auto ok = [&](const ConvBlock& b) { return b.bb && b.tc_wgrad.valid && b.tc_wgrad.p.bf16; };
    for (auto& b : h->enc) all = all && ok(b);
    for (auto& b : h->dec) all = all && ok(b);
    all = all && ok(h->bott);
    for (auto& u : h->ups) all = all && u.bb && u.tc_wgrad.valid && !u.bilinear;
    h->overlap = all;
  }
  return 0;
}

int check_shape(const rd_handle* h, int B, int T) {
  if (B < 1) return fail("batch must be >= 1 (got %d)", B);
  const int D = h->depth;
  if (T < (1 << D) || T % (1 << D)) return fail("tile size %d is not a positive multiple of 2^depth = %d", T, 1 << D);
  return 0;
}

BnLayer bn_view(const rd_handle* h, const ConvBlock& b) {
  BnLayer L{};
  L.C = b.Cout;
  L.gamma = b.gamma >= 0 ? h->P + b.gamma : nullptr;
  L.beta = b.beta >= 0 ? h->P + b.beta : nullptr;
  L.conv_bias = b.bias >= 0 ? h->P + b.bias : nullptr;
  L.running_mean = b.rm >= 0 ? h->BUF + b.rm : nullptr;
  L.running_var = b.rv >= 0 ? h->BUF + b.rv : nullptr;
  L.mean = b.mean; L.invstd = b.invstd; L.scale = b.scale; L.shift = b.shift;
  return L;
}

Act act_view(const rd_handle* h, const ConvBlock& b) {
  Act a{};
  a.kind = b.act;
  if (b.act == RD_ACT_PRELU) a.slope = h->P + b.slope;
  else a.slope = h->consts + (b.act == RD_ACT_LRELU ? 1 : 0);
  return a;
}

int pack_weights(rd_handle* h, bool for_backward, cudaStream_t s) {
  const int rnd = h->tf32() ? 1 : 0;
  ProfScope ps(h, RD_PROF_PACK, 0.0, 4.0 * 3.0 * (double)h->param_floats, s);
  PackJobs J;
  auto block = [&](ConvBlock& b) -> int {
    if (!b.w_kn) return 0;
    // tcgen05 layers read the [N][K] copies, CUDA-core layers the [K][N] copies: pack only what is used
    return pack_jobs_add(J, PACK_CONV3X3, h->P + b.w, b.tc ? nullptr : b.w_kn, b.tc ? b.w_nk : nullptr,
                         for_backward && !b.tc ? b.wd_kn : nullptr, for_backward && b.tc && !b.bb ? b.wd_nk : nullptr,
                         for_backward && b.bb ? b.wd_nk_b : nullptr, b.Cout, b.Cin, rnd && b.tc);
  };
  for (auto& b : h->enc) RD_TRY(block(b));
  RD_TRY(block(h->bott));
  for (auto& b : h->dec) RD_TRY(block(b));
  for (auto& u : h->ups) {
    if (u.bilinear) RD_TRY(pack_jobs_add(J, PACK_CONV1X1, h->P + u.w, u.w_nk, u.w_kn, nullptr, nullptr, nullptr, u.C, u.C, rnd && u.tc));
    else RD_TRY(pack_jobs_add(J, PACK_CONVT, h->P + u.w, u.w_kn, u.w_nk, nullptr, nullptr, for_backward && u.bb ? u.w_kn_b : nullptr,
                              u.C, u.C, rnd && u.tc));
  }
  return launch_pack_batched(J, s);      // one launch for all layers
}

// conv3x3 of one block: src NHWC [B,H,W,Cin] -> z (+ statistics partials)
int conv_block_forward(rd_handle* h, ConvBlock& b, const float* src, int B, int H, bool batch_stats, int* np,
                       cudaStream_t s) {
  Gather g = gather_conv3x3(H, H, b.Cin);
  Epilogue e{};
  e.mode = batch_stats ? EPI_STATS : EPI_PLAIN;
  e.out = b.z;
  e.partials = h->partials;
  *np = 0;
  const double px = (double)B * H * H;
  ProfScope ps(h, RD_PROF_CONV_FWD, 2.0 * 9.0 * b.Cin * b.Cout * px, 4.0 * px * (b.Cin + b.Cout), s);
  if (b.tc) return launch_gemm_rows_tc(b.tc_fwd, e, np, s);
  return launch_gemm_rows_simt(src, g, b.w_kn, B, b.Cout, e, np, s);
}

// Inference (RD_FWD_EVAL) on a tcgen05 block: BatchNorm(running stats) + activation (+ 2x2 max-pool) folded into the
// conv epilogue -- the raw conv output z is never written
int conv_block_forward_fused_eval(rd_handle* h, ConvBlock& b, int B, int H, int round_a, int round_p, cudaStream_t s) {
  Epilogue e{};                                          // scale / shift: bn_eval_all at the start of rd_forward
  e.mode = EPI_BNACT;
  e.out = b.a;
  e.round_tf32 = round_a;
  e.scale = b.scale;
  e.shift = b.shift;
  e.slope = act_view(h, b).slope;
  e.pool_out = b.pool ? b.p : nullptr;
  e.round_pool = round_p;
  const double px = (double)B * H * H;
  ProfScope ps(h, RD_PROF_CONV_FWD, 2.0 * 9.0 * b.Cin * b.Cout * px, 4.0 * px * (b.Cin + b.Cout * (b.pool ? 1.25 : 1.0)), s);
  return launch_gemm_rows_tc(b.tc_fwd, e, nullptr, s);
}

// eval-mode BatchNorm (running statistics) of every block: one launch per forward pass
int bn_eval_all(rd_handle* h, cudaStream_t s) {
  BnEvalJobs J;
  for (auto& b : h->enc) RD_TRY(bn_eval_jobs_add(J, bn_view(h, b), h->cfg.do_bn));
  RD_TRY(bn_eval_jobs_add(J, bn_view(h, h->bott), h->cfg.do_bn));
  for (auto& b : h->dec) RD_TRY(bn_eval_jobs_add(J, bn_view(h, b), h->cfg.do_bn));
  ProfScope ps(h, RD_PROF_BN_FINALIZE, 0.0, 0.0, s);
  return launch_bn_eval_batched(J, s);
}

// encoder level i keeps only z and the pooled tensor: its up-conv (tcgen05, transposed) applies BN + activation to z
// when it adds the skip.  Only in the saving forward modes (the fused inference path never writes z).
bool skip_from_z(const rd_handle* h, int i, int save) {
  const UpConv& u = h->ups[h->depth - 1 - i];
  return save && h->enc[i].pool && u.tc && !u.bilinear && (h->enc[i].Cout % 32 == 0);
}

// BN statistics finalize + fused normalise/activation(/pool) pass of one block
int bn_act_forward(rd_handle* h, ConvBlock& b, int np, int B, int H, bool train, int round_a, int round_p,
                   bool shadows, cudaStream_t s, bool write_a = true) {
  BnLayer L = bn_view(h, b);
  {
    // running-statistics modes were finalized for all layers at once (bn_eval_all); batch statistics need this layer's sums
    if (train) {
      ProfScope ps(h, RD_PROF_BN_FINALIZE, 0.0, 0.0, s);
      RD_TRY(launch_bn_finalize(L, h->partials, np, (long long)B * H * H, train, h->cfg.do_bn, s));
    }
  }
  const double n = (double)B * H * H * b.Cout;
  ProfScope ps(h, RD_PROF_BN_ACT_POOL, 0.0, 4.0 * n * (b.pool ? (write_a ? 2.25 : 1.25) : 2.0), s);
  return launch_bn_act_pool(b.z, b.scale, b.shift, act_view(h, b), write_a ? b.a : nullptr, b.pool ? b.p : nullptr, B, H, H, b.Cout,
                            round_a, round_p, shadows ? b.a_b : nullptr, shadows ? b.p_b : nullptr, s);
}

}  // namespace

extern "C" {

int rd_abi_version(void) { return RD_ABI_VERSION; }
const char* rd_last_error(void) { return g_last_error.c_str(); }

int rd_create(const rd_config* cfg, int device, rd_handle** out) {
  if (!cfg || !out) return fail("rd_create: null argument");
  *out = nullptr;
  if (cfg->n_input_channels < 1 || cfg->n_input_channels > 8)
    return fail("n_input_channels=%d outside the supported range 1..8", cfg->n_input_channels);
  if (cfg->depth < 1 || cfg->depth > 8) return fail("depth=%d outside the supported range 1..8", cfg->depth);
  if (cfg->start_kernel < 32 || cfg->start_kernel % 32)
    return fail("start_kernel=%d must be a positive multiple of 32", cfg->start_kernel);
  if (cfg->start_kernel > 128) return fail("start_kernel=%d > 128 is not supported by the full-resolution kernels", cfg->start_kernel);
  if (cfg->max_filter_depth % 32 || cfg->max_filter_depth < cfg->start_kernel || cfg->max_filter_depth > 1024)
    return fail("max_filter_depth=%d must be a multiple of 32 in [start_kernel, 1024]", cfg->max_filter_depth);
  if (cfg->up_mode != RD_UP_TRANSPOSE && cfg->up_mode != RD_UP_BILINEAR) return fail("unknown up_mode %d", cfg->up_mode);
  for (int a : {cfg->act_encoder, cfg->act_decoder, cfg->act_bottleneck})
    if (a < RD_ACT_RELU || a > RD_ACT_PRELU) return fail("unknown activation id %d", a);
  if (cfg->math_mode != RD_MATH_FP32 && cfg->math_mode != RD_MATH_TF32) return fail("unknown math mode %d", cfg->math_mode);
  if (cfg->bwd_mode < RD_BWD_AUTO || cfg->bwd_mode > RD_BWD_BF16) return fail("unknown bwd_mode %d", cfg->bwd_mode);

  rd_handle* h = new rd_handle();
  h->cfg = *cfg;
  h->device = device;
  const int D = h->depth = cfg->depth;
  for (int i = 0; i < D; ++i) {
    long long w = (long long)cfg->start_kernel << i;
    h->widths.push_back((int)(w > cfg->max_filter_depth ? cfg->max_filter_depth : w));   // lib/UNet.py:152-155
  }
  h->enc.resize(D);
  h->dec.resize(D - 1);
  h->ups.resize(D);
  for (int i = 0; i < D; ++i)
    plan_block(h, h->enc[i], "encoder." + std::to_string(i) + ".0", i == 0 ? cfg->n_input_channels : h->widths[i - 1],
               h->widths[i], cfg->act_encoder, true);
  plan_block(h, h->bott, "bottleneck", h->widths[D - 1], h->widths[D - 1], cfg->act_bottleneck, false);
  for (int j = 0; j < D; ++j) {
    const int C = h->widths[D - 1 - j];
    h->ups[j].C = C;
    const bool bil = cfg->up_mode == RD_UP_BILINEAR;
    h->ups[j].bilinear = bil;
    // transposed: the ConvTranspose2d itself; bilinear: Sequential(Upsample, Conv2d 1x1) -> ".1" is the conv
    const std::string p = "decoder." + std::to_string(j) + (j < D - 1 ? ".0" : "") + (bil ? ".1" : "");
    add_param(h, p + ".weight", (long long)C * C * (bil ? 1 : 4), &h->ups[j].w);
    add_param(h, p + ".bias", C, &h->ups[j].bias);
    if (j < D - 1)
      plan_block(h, h->dec[j], "decoder." + std::to_string(j) + ".1", C, h->widths[D - 2 - j], cfg->act_decoder, false);
  }
  {
    // operand type of the backward GEMMs: explicit in the config, or (RD_BWD_AUTO) bf16 unless RESDEPTH_BWD=tf32
    const char* env = getenv("RESDEPTH_BWD");
    bool bf = !(env && std::string(env) == "tf32");
    if (cfg->bwd_mode == RD_BWD_TF32) bf = false;
    if (cfg->bwd_mode == RD_BWD_BF16) bf = true;
    h->bwd_bf16 = cfg->math_mode == RD_MATH_TF32 && bf;
  }
  add_param(h, "last_layer.weight", (long long)cfg->start_kernel * 9, &h->last_w);
  if (cfg->bias_conv_layer) add_param(h, "last_layer.bias", 1, &h->last_b);
  if (cfg->outer_skip && cfg->outer_skip_bn) {                       // lib/UNet.py:189-194
    add_param(h, "layer_outer_skip.0.weight", 1, &h->ob_gamma);
    add_param(h, "layer_outer_skip.0.bias", 1, &h->ob_beta);
    add_buffer(h, "layer_outer_skip.0.running_mean", 1, &h->ob_rm);
    add_buffer(h, "layer_outer_skip.0.running_var", 1, &h->ob_rv);
  }
  *out = h;
  return 0;
}

static void destroy_side(rd_handle* h) {
  if (h->side) cudaStreamDestroy(h->side);
  for (cudaEvent_t e : {h->ev_main, h->ev_join, h->ev_wg[0], h->ev_wg[1]})
    if (e) cudaEventDestroy(e);
  h->side = nullptr;
  h->ev_main = h->ev_join = h->ev_wg[0] = h->ev_wg[1] = nullptr;
}

int rd_destroy(rd_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->slab) cudaFree(h->slab);
  for (auto& w : h->ws_cache)
    if (w.slab) cudaFree(w.slab);
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto& e : h->event_pool) cudaEventDestroy(e);
  destroy_side(h);
  delete h;
  return 0;
}

int rd_num_params(const rd_handle* h) { return (int)h->params.size(); }
int rd_num_buffers(const rd_handle* h) { return (int)h->buffers.size(); }
int64_t rd_param_arena_size(const rd_handle* h) { return h->param_floats; }
int64_t rd_buffer_arena_size(const rd_handle* h) { return h->buffer_floats; }

static int info(const std::vector<ParamInfo>& v, int index, char* name64, int64_t* numel, int64_t* offset) {
  if (index < 0 || index >= (int)v.size()) return fail("index %d out of range", index);
  if (name64) {
    std::strncpy(name64, v[index].name.c_str(), 63);
    name64[63] = 0;
  }
  if (numel) *numel = v[index].numel;
  if (offset) *offset = v[index].offset;
  return 0;
}
int rd_param_info(const rd_handle* h, int index, char* name64, int64_t* numel, int64_t* offset) {
  return info(h->params, index, name64, numel, offset);
}
int rd_buffer_info(const rd_handle* h, int index, char* name64, int64_t* numel, int64_t* offset) {
  return info(h->buffers, index, name64, numel, offset);
}

int rd_bind(rd_handle* h, float* params, float* grads, float* bn_buffers) {
  if (!h) return fail("rd_bind: null handle");
  if (!params) return fail("rd_bind: null parameter arena");
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)bn_buffers) & 15) return fail("rd_bind: arenas must be 16-byte aligned");
  h->P = params; h->G = grads; h->BUF = bn_buffers;
  ++h->freeze_gen;                               // packs built from the previous arenas are stale
  return 0;
}

static const int kMaxLayouts = 4;     // the current one + three cached (train batch, validation batch, partial last batch)

static bool layout_serves(const rd::Workspace& w, int batch, int tile, int with_backward) {
  return w.slab && batch == w.res_batch && tile == w.res_tile && (w.res_bwd || !with_backward);
}

static int evict_oldest_layout(rd_handle* h) {
  size_t k = 0;
  for (size_t i = 1; i < h->ws_cache.size(); ++i)
    if (h->ws_cache[i].ws_last_use < h->ws_cache[k].ws_last_use) k = i;
  RD_CUDA(cudaDeviceSynchronize());              // work on the evicted slab may still be in flight
  RD_CUDA(cudaFree(h->ws_cache[k].slab));
  h->ws_cache.erase(h->ws_cache.begin() + k);
  return 0;
}

int rd_reserve(rd_handle* h, int batch, int tile, int with_backward) {
  if (!h) return fail("rd_reserve: null handle");
  RD_TRY(check_shape(h, batch, tile));
  RD_CUDA(cudaSetDevice(h->device));
  rd::Workspace& cur = *h;
  if (layout_serves(cur, batch, tile, with_backward)) return 0;
  // another shape: park the current layout (slab, carved pointers, TMA descriptors stay valid) and look for one that
  // was built for this shape before -- switching costs neither a synchronisation nor a plan rebuild
  if (cur.slab) {
    cur.ws_last_use = ++h->ws_clock;
    h->ws_cache.push_back(cur);
    cur.slab = nullptr; cur.slab_bytes = 0; cur.ws_id = 0;
    cur.res_batch = cur.res_tile = cur.res_bwd = 0;
  }
  h->fwd_mode = -1;
  for (size_t i = 0; i < h->ws_cache.size(); ++i) {
    if (layout_serves(h->ws_cache[i], batch, tile, with_backward)) {
      cur = h->ws_cache[i];
      h->ws_cache.erase(h->ws_cache.begin() + i);
      return 0;
    }
  }
  while ((int)h->ws_cache.size() >= kMaxLayouts - 1) RD_TRY(evict_oldest_layout(h));
  const size_t need = carve(h, nullptr, batch, tile, with_backward);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, need);
  while (e != cudaSuccess && !h->ws_cache.empty()) {       // out of memory: give the parked layouts back first
    cudaGetLastError();
    RD_TRY(evict_oldest_layout(h));
    e = cudaMalloc(&p, need);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("workspace allocation of %.1f MB failed: %s", need / 1048576.0, cudaGetErrorString(e));
  }
  cur.slab = p; cur.slab_bytes = need;
  carve(h, cur.slab, batch, tile, with_backward);
  if (with_backward && !h->side) {                          // side stream + events of the overlapped weight gradients
    // High priority: a weight-gradient CTA (one per SM, tensor-bound, ~200 KB of shared memory) should take the first
    // SM slot an HBM-bound BatchNorm CTA of the main stream frees, so that the two kinds of work really run side by
    // side; at default priority the block scheduler drains the older kernel's grid first and the GEMM starts in its tail.
    static const bool flat_prio = getenv("RESDEPTH_SIDE_PRIO0") != nullptr;
    int prio_lo = 0, prio_hi = 0;
    RD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    RD_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, flat_prio ? prio_lo : prio_hi));
    for (cudaEvent_t* ev : {&h->ev_main, &h->ev_join, &h->ev_wg[0], &h->ev_wg[1]})
      RD_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  }
  RD_TRY(build_tc_plans(h, batch, tile, with_backward));
  const float consts[4] = {0.f, 0.01f, 1.f, 0.f};          // relu slope, LeakyReLU default slope (lib/UNet.py:30)
  RD_CUDA(cudaMemcpy(h->consts, consts, sizeof(consts), cudaMemcpyHostToDevice));
  if (h->xcol_b) {                                         // row padding of the bf16 im2col expansion stays zero
    RD_TRY(launch_im2col_first_bf16_clear(h->xcol_b, (size_t)batch * tile * tile * h->xcol_b_pitch * 2, nullptr));
    RD_CUDA(cudaDeviceSynchronize());
  }
  cur.res_batch = batch; cur.res_tile = tile; cur.res_bwd = with_backward;
  cur.eval_pack_gen = 0;
  cur.ws_id = h->ws_next_id++;
  return 0;
}

int64_t rd_workspace_id(const rd_handle* h) { return h ? h->ws_id : 0; }

int rd_workspace_alive(const rd_handle* h, int64_t id) {
  if (!h || id <= 0) return 0;
  if (h->slab && h->ws_id == id) return 1;
  for (auto& w : h->ws_cache)
    if (w.ws_id == id) return 1;
  return 0;
}

int64_t rd_workspace_bytes(const rd_handle* h) { return h ? (int64_t)h->slab_bytes : 0; }

int rd_forward(rd_handle* h, const float* x, float* y, int batch, int tile, int mode, void* stream) {
  if (!h || !x || !y) return fail("rd_forward: null argument");
  if (!h->P) return fail("rd_forward: parameters not bound (rd_bind)");
  if (mode < RD_FWD_EVAL || mode > RD_FWD_EVAL_SAVE) return fail("rd_forward: bad mode %d", mode);
  if ((h->cfg.do_bn || h->ob_gamma >= 0) && !h->BUF) return fail("rd_forward: BatchNorm buffers not bound");
  RD_CUDA(cudaSetDevice(h->device));
  const int save = mode != RD_FWD_EVAL;
  RD_TRY(rd_reserve(h, batch, tile, save));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = batch, T = tile, D = h->depth;
  const bool train = mode == RD_FWD_TRAIN;
  const bool stats = train && h->cfg.do_bn;
  const int tf = h->tf32() ? 1 : 0;
  const bool fuse_eval = mode == RD_FWD_EVAL;         // nothing is kept for a backward pass: fold BN into the convs
  h->fwd_mode = -1;
  // constant-weight inference (rd_freeze_params): this layout's packs were built from the frozen arenas already
  const bool reuse_packs = fuse_eval && h->frozen && h->eval_pack_gen == h->freeze_gen;
  if (!reuse_packs) {
    RD_TRY(pack_weights(h, save, s));
    if (!train) RD_TRY(bn_eval_all(h, s));
  }
  h->eval_pack_gen = (fuse_eval && h->frozen) ? h->freeze_gen : 0;

  int np = 0;
  for (int i = 0; i < D; ++i) {
    ConvBlock& b = h->enc[i];
    const int H = T >> i;
    if (i == 0) {
      const double px = (double)B * H * H;
      ProfScope ps(h, RD_PROF_FIRST_FWD, 2.0 * 9.0 * b.Cin * b.Cout * px, 4.0 * px * (b.Cin + b.Cout), s);
      static const bool no_first_tc = getenv("RESDEPTH_NO_FIRST_TC") != nullptr;
      static const bool no_first_fuse = getenv("RESDEPTH_NO_FIRST_FUSE") != nullptr;
      if (fuse_eval && h->tf32() && !no_first_tc && !no_first_fuse && conv_first_tc_fuse_ok(b.Cin, b.Cout, H, H)) {
        // inference: BatchNorm(running statistics) + activation + max-pool folded into the conv (z0 is never written)
        RD_TRY(launch_conv_first_tc(x, h->P + b.w, nullptr, nullptr, &np, B, b.Cin, H, H, b.Cout, s, b.scale, b.shift,
                                    act_view(h, b).slope, b.a, b.p, tf));
        continue;
      }
      if (h->tf32() && !no_first_tc && conv_first_tc_shape_ok(b.Cin, b.Cout, H, H))
        RD_TRY(launch_conv_first_tc(x, h->P + b.w, b.z, stats ? h->partials : nullptr, &np, B, b.Cin, H, H, b.Cout, s));
      else
        RD_TRY(launch_conv_first_fwd(x, h->P + b.w, b.z, stats ? h->partials : nullptr, &np, B, b.Cin, H, H, b.Cout, s));
    } else if (fuse_eval && b.tc) {
      RD_TRY(conv_block_forward_fused_eval(h, b, B, H, 0, tf, s));
      continue;
    } else {
      RD_TRY(conv_block_forward(h, b, h->enc[i - 1].p, B, H, stats, &np, s));
    }
    // training / saving forward: the full-resolution activated tensor of an encoder level is only ever read as the
    // additive skip of its up-conv; a tcgen05 transposed conv re-derives it from z in its epilogue, so it is not written
    RD_TRY(bn_act_forward(h, b, np, B, H, train, 0, tf, save != 0, s, !skip_from_z(h, i, save)));
  }
  {
    ConvBlock& b = h->bott;
    const int H = T >> D;
    if (fuse_eval && b.tc) {
      RD_TRY(conv_block_forward_fused_eval(h, b, B, H, tf, 0, s));
    } else {
      RD_TRY(conv_block_forward(h, b, h->enc[D - 1].p, B, H, stats, &np, s));
      RD_TRY(bn_act_forward(h, b, np, B, H, train, tf, 0, save != 0, s));
    }
  }
  const float* cur = h->bott.a;
  int Hc = T >> D;
  for (int j = 0; j < D; ++j) {
    UpConv& u = h->ups[j];
    Gather g = gather_plain(Hc, Hc, u.C);
    Epilogue e{};
    e.mode = EPI_CONVT;
    e.out = u.u;
    e.bias = h->P + u.bias;
    e.skip = h->enc[D - 1 - j].a;                       // additive skip, lib/UNet.py:96-101,220,224
    if (skip_from_z(h, D - 1 - j, save)) {
      ConvBlock& eb = h->enc[D - 1 - j];
      e.skip = eb.z;
      e.skip_scale = eb.scale;
      e.skip_shift = eb.shift;
      e.skip_slope = act_view(h, eb).slope;
    }
    e.round_tf32 = (j < D - 1) ? tf : 0;
    e.out_b = (save && u.tc && !u.bilinear) ? u.u_b : nullptr;       // bf16 copy: wgrad operand of the next conv
    if (u.bilinear) {
      // conv1x1(upsample(h)) == upsample(conv1x1(h)): 1-tap GEMM at the low resolution, then interpolate + bias + skip
      const double px = (double)B * Hc * Hc;
      ProfScope ps(h, RD_PROF_CONVT_FWD, 2.0 * u.C * u.C * px, 4.0 * px * u.C * (1.0 + 4.0 + 4.0), s);
      Epilogue ep{};
      ep.mode = EPI_PLAIN;
      ep.out = u.t;
      if (u.tc) RD_TRY(launch_gemm_rows_tc(u.tc_fwd, ep, nullptr, s));
      else RD_TRY(launch_gemm_rows_simt(cur, g, u.w_kn, B, u.C, ep, nullptr, s));
      RD_TRY(launch_bilinear_up_add(u.t, e.bias, e.skip, u.u, B, Hc, Hc, u.C, e.round_tf32, s));
    } else {
      const double px = (double)B * Hc * Hc;
      ProfScope ps(h, RD_PROF_CONVT_FWD, 2.0 * 4.0 * u.C * u.C * px, 4.0 * px * u.C * (1.0 + 4.0 + 4.0), s);
      if (u.tc) RD_TRY(launch_gemm_rows_tc(u.tc_fwd, e, nullptr, s));
      else RD_TRY(launch_gemm_rows_simt(cur, g, u.w_kn, B, 4 * u.C, e, nullptr, s));
    }
    Hc *= 2;
    if (j < D - 1) {
      ConvBlock& b = h->dec[j];
      if (fuse_eval && b.tc) {
        RD_TRY(conv_block_forward_fused_eval(h, b, B, Hc, tf, 0, s));
      } else {
        RD_TRY(conv_block_forward(h, b, u.u, B, Hc, stats, &np, s));
        RD_TRY(bn_act_forward(h, b, np, B, Hc, train, tf, 0, save != 0, s));
      }
      cur = b.a;
    }
  }
  const int C0 = h->cfg.start_kernel;
  {
    const double px = (double)B * T * T;
    ProfScope ps(h, RD_PROF_LAST_FWD, 2.0 * 9.0 * C0 * px, 4.0 * px * (C0 + 2.0), s);
    const float* x_affine = nullptr;
    if (h->ob_gamma >= 0) {                                   // outer_skip_BN: BatchNorm2d(1) on channel 0 of x
      int nparts = 0;
      if (train) RD_TRY(launch_outer_bn_reduce(x, nullptr, nullptr, h->partials, &nparts, B, h->cfg.n_input_channels, T * T, s));
      BnLayer L{};
      L.C = 1;
      L.gamma = h->P + h->ob_gamma; L.beta = h->P + h->ob_beta;
      L.running_mean = h->BUF + h->ob_rm; L.running_var = h->BUF + h->ob_rv;
      L.mean = h->ob_mean; L.invstd = h->ob_invstd; L.scale = h->ob_affine; L.shift = h->ob_affine + 1;
      RD_TRY(launch_bn_finalize(L, h->partials, nparts, (long long)B * T * T, train, 1, s));
      x_affine = h->ob_affine;
    }
    RD_TRY(launch_conv_last_fwd(h->ups[D - 1].u, h->P + h->last_w, h->last_b >= 0 ? h->P + h->last_b : nullptr,
                                h->cfg.outer_skip ? x : nullptr, h->cfg.n_input_channels * T * T, x_affine, y, B, T, T,
                                C0, h->taps, s));
  }
  if (save) { h->fwd_batch = B; h->fwd_tile = T; h->fwd_mode = mode; h->bw_next_stage = 0; }
  return 0;
}

int rd_loss(rd_handle* h, const float* y_pred, const float* target, const uint8_t* mask, const float* mean,
            const float* std, float* loss_out, float* dy_out, int batch, int tile, void* stream) {
  if (!h || !y_pred || !target || !mask || !mean || !std || !loss_out) return fail("rd_loss: null argument");
  RD_CUDA(cudaSetDevice(h->device));
  if (!h->slab) RD_TRY(rd_reserve(h, batch, tile, 0));
  ProfScope ps(h, RD_PROF_LOSS, 0.0, (double)batch * tile * tile * (4.0 + 4.0 + 1.0 + (dy_out ? 13.0 : 0.0)),
               reinterpret_cast<cudaStream_t>(stream));
  return launch_loss(y_pred, target, mask, mean, std, loss_out, dy_out, h->scratch, batch, tile * tile,
                     reinterpret_cast<cudaStream_t>(stream));
}

namespace {

// the gradient at u (= skip gradient of encoder level i) lives only in its bf16 copy gs_b[i]: the producer (last-conv
// backward for level 0, the tcgen05 dgrad of the decoder conv above for the others) and both readers (the up-conv's
// bf16 GEMMs and the BatchNorm backward of the encoder level) agree on this predicate
bool skip_grad_bf16(const rd_handle* h, int i) {
  const int D = h->depth, j = D - 1 - i;               // up-conv whose output gradient this is
  if (!h->gs_b[i] || !h->ups[j].bb) return false;
  return i == 0 ? true : h->dec[j].tc;
}

// backward of one conv block.  g_full: gradient at the (un-pooled) block output, g_pool: gradient at the pooled
// output; src_in: the block's input (NHWC) or, for the first encoder block, the network input x (NCHW).
struct GradRef {          // a gradient tensor stored as fp32 or as bf16 (the bf16 backward keeps some only in bf16)
  const void* p = nullptr;
  int bf16 = 0;
  GradRef() {}
  GradRef(const float* f) : p(f), bf16(0) {}
  GradRef(const void* q, int is_bf16) : p(q), bf16(is_bf16) {}
};

// Gram matrix of the first layer's im2col expansion: reduce GEMM (xcol, xcol) -> split partials -> h->gram
int first_layer_gram(rd_handle* h, cudaStream_t s) {
  RD_TRY(launch_gemm_reduce_tc(h->tc_gram, s));
  const int kc = h->xcol_b_k;
  return launch_sum_partials(h->part, h->tc_gram.splits, kc * kc, kc * kc, 1, h->gram, s);
}

int block_backward(rd_handle* h, ConvBlock& b, GradRef g_full, GradRef g_pool, int B, int H,
                   const float* src_in, bool first, float* dgrad_out, int round_dgrad, float* dgrad_colsum,
                   void* dgrad_out_b, cudaStream_t s) {
  BnLayer L = bn_view(h, b);
  Act act = act_view(h, b);
  const int do_bn = h->cfg.do_bn;
  int np = 0;
  const double px = (double)B * H * H, n = px * b.Cout;
  const double gin = (g_full.p ? (g_full.bf16 ? 0.5 : 1.0) : 0.0) + (g_pool.p ? (g_pool.bf16 ? 0.125 : 0.25) : 0.0);
  // First encoder block on the bf16 tcgen05 path: its dz feeds only the weight gradient, which is rebuilt from gY, the
  // Gram matrix of the input expansion and the BatchNorm coefficients (launch_first_grad_correct) -- the reduce pass
  // stores gY and the apply pass (a second sweep over z, the largest tensor of the network) is skipped.
  const bool gy_path = first && b.bb && b.tc_wgrad.valid && (h->tc_gram.valid || h->gram_fused) && g_pool.p && do_bn;
  const bool ov0 = h->overlap && h->overlap_allowed;
  if (gy_path && ov0 && h->wg_pending[b.par]) {           // gY goes into this block's dz buffer: wait for its last reader
    RD_CUDA(cudaStreamWaitEvent(s, h->ev_wg[b.par], 0));
    h->wg_pending[b.par] = false;
  }
  {
    ProfScope ps(h, RD_PROF_BN_BWD_REDUCE, 0.0, 4.0 * n * (1.0 + gin + (gy_path ? 0.5 : 0.0)), s);
    RD_TRY(launch_bn_bwd_reduce(g_full.p, g_pool.p, g_full.bf16, g_pool.bf16, b.z, L, act, h->partials, &np, B, H, H, s,
                                gy_path ? b.gyb : nullptr));
    RD_TRY(launch_bn_bwd_finalize(L, h->partials, np, (long long)B * H * H, do_bn, h->fwd_mode == RD_FWD_TRAIN,
                                  do_bn ? h->G + b.gamma : nullptr, h->G + (do_bn ? b.beta : b.bias),
                                  b.slope >= 0 ? h->G + b.slope : nullptr, h->scratch, h->coef, s));
  }
  // weight gradients go to the side stream (h->overlap): fork after dz is written, join at the end of rd_backward;
  // the dz buffer of this parity may still be read by the weight gradient launched two blocks ago
  const bool ov = h->overlap && h->overlap_allowed;
  cudaStream_t ws = ov ? h->side : s;
  if (ov && h->wg_pending[b.par]) {
    RD_CUDA(cudaStreamWaitEvent(s, h->ev_wg[b.par], 0));
    h->wg_pending[b.par] = false;
  }
  if (!gy_path) {
    ProfScope ps(h, RD_PROF_BN_BWD_APPLY, 0.0, 4.0 * n * (1.0 + (b.bb ? 0.5 : 1.0) + gin), s);
    RD_TRY(launch_bn_bwd_apply(g_full.p, g_pool.p, g_full.bf16, g_pool.bf16, b.z, L, act, h->coef, b.bb ? nullptr : h->gy, b.bb ? b.gyb : nullptr, B,
                               H, H, h->tf32() && (b.tc || b.tc_wgrad.valid), s));
  }
  // Order of the two GEMMs that read dz.  Both need a whole SM (shared memory), so they never run side by side; what can
  // share an SM with the weight gradient is the HBM-bound BatchNorm backward of the NEXT block.  With the side stream on,
  // the data gradient (critical path) is therefore enqueued first and the weight gradient is released when it completes:
  // it then overlaps the main stream's next elementwise kernels instead of delaying the data gradient.
  static const bool wg_first = getenv("RESDEPTH_WG_FIRST") != nullptr;
  const bool dgrad_first = ov && !wg_first && !first && (dgrad_out || dgrad_out_b);
  auto do_wgrad = [&]() -> int {
  if (ov) {
    RD_CUDA(cudaEventRecord(h->ev_main, s));
    RD_CUDA(cudaStreamWaitEvent(ws, h->ev_main, 0));
  }
  if (first) {
    ProfScope ps(h, RD_PROF_FIRST_WGRAD, 2.0 * 9.0 * b.Cin * b.Cout * px, 4.0 * px * (b.Cin + b.Cout), ws);
    if (b.tc_wgrad.valid) {
      const int kc = b.bb ? h->xcol_b_k : h->xcol_k;
      if (h->xcol_early) h->xcol_early = false;           // expansion already enqueued at the start of the backward pass
      else if (b.bb) {
        RD_TRY(launch_im2col_first_bf16(src_in, h->xcol_b, B, b.Cin, H, H, h->xcol_b_pitch, ws));
        if (gy_path && !h->gram_fused) RD_TRY(first_layer_gram(h, ws));
      } else RD_TRY(launch_im2col_first(src_in, h->xcol, B, b.Cin, H, H, kc, 1, ws));
      RD_TRY(launch_gemm_reduce_tc(b.tc_wgrad, ws));
      if (h->gram_fused && b.bb) {
        // partials [split][kc][2 kc] -> h->gram [kc][2 kc]: columns 0..kc-1 = Gram matrix, kc..2kc-1 = X^T gY
        RD_TRY(launch_sum_partials(h->part, b.tc_wgrad.splits, kc * 2 * kc, kc * 2 * kc, 1, h->gram, ws));
        RD_TRY(launch_unpack_first_grad(h->gram + kc, 1, h->G + b.w, b.Cout, b.Cin * 9, kc, ws, 2 * kc));
        if (gy_path) RD_TRY(launch_first_grad_correct(h->G + b.w, h->P + b.w, h->gram, h->coef, b.Cout, b.Cin * 9, 2 * kc, ws));
      } else {
        RD_TRY(launch_unpack_first_grad(h->part, b.tc_wgrad.splits, h->G + b.w, b.Cout, b.Cin * 9, kc, ws));
        if (gy_path) RD_TRY(launch_first_grad_correct(h->G + b.w, h->P + b.w, h->gram, h->coef, b.Cout, b.Cin * 9, kc, ws));
      }
    } else {
      RD_TRY(launch_conv_first_wgrad(src_in, h->gy, h->G + b.w, h->scratch, h->scratch_floats, B, b.Cin, H, H, b.Cout, s));
    }
  } else {
    Gather g = gather_conv3x3(H, H, b.Cin);
    int S = 0;
    {
      ProfScope ps(h, RD_PROF_CONV_WGRAD, 2.0 * 9.0 * b.Cin * b.Cout * px, 4.0 * px * (b.Cin + b.Cout), ws);
      if (b.tc_wgrad.valid) {
        RD_TRY(launch_gemm_reduce_tc(b.tc_wgrad, ws));
        S = b.tc_wgrad.splits;
      } else {
        RD_TRY(launch_gemm_reduce_simt(src_in, g, h->gy, B, b.Cout, h->part, h->part_floats, &S, ws));
      }
    }
    ProfScope ps(h, RD_PROF_UNPACK, 0.0, 4.0 * 9.0 * b.Cin * b.Cout * (S + 1.0), ws);
    if (b.wide && b.tc_wgrad.valid)
      RD_TRY(launch_unpack_conv_grad_wide(h->part, S, b.tc_wgrad.p.N, h->G + b.w, b.Cout, b.Cin, b.wide == 1, ws));
    else
      RD_TRY(launch_unpack_conv_grad(h->part, S, h->G + b.w, b.Cout, b.Cin, 9, ws));
  }
  if (ov) {
    RD_CUDA(cudaEventRecord(h->ev_wg[b.par], ws));
    h->wg_pending[b.par] = true;
  }
  return 0;
  };
  auto do_dgrad = [&]() -> int {
  if (dgrad_out || dgrad_out_b) {
    Gather g = gather_conv3x3(H, H, b.Cout);
    Epilogue e{};
    e.mode = dgrad_colsum ? EPI_STATS : EPI_PLAIN;       // column sums of dX = bias gradient of the up-conv before it
    e.out = dgrad_out;
    e.partials = h->partials;
    e.round_tf32 = round_dgrad;
    e.out_b = dgrad_out_b;
    int npart = 0;
    {
      ProfScope ps(h, RD_PROF_CONV_DGRAD, 2.0 * 9.0 * b.Cin * b.Cout * px,
                   px * (b.Cin * (dgrad_out ? 4.0 : 2.0) + b.Cout * (b.bb ? 2.0 : 4.0)), s);
      if (!dgrad_out && !b.tc) return fail("block_backward: bf16-only gradient output needs the tcgen05 dgrad");
      if (b.tc) RD_TRY(launch_gemm_rows_tc(b.tc_dgrad, e, &npart, s));
      else RD_TRY(launch_gemm_rows_simt(h->gy, g, b.wd_kn, B, b.Cin, e, &npart, s));
    }
    if (dgrad_colsum) {
      ProfScope ps(h, RD_PROF_BIAS_GRAD, 0.0, 0.0, s);
      RD_TRY(launch_sum_partials(h->partials, npart, b.Cin, 2 * b.Cin, 2, dgrad_colsum, s));
    }
  }
  return 0;
  };
  if (dgrad_first) {
    RD_TRY(do_dgrad());
    RD_TRY(do_wgrad());
  } else {
    RD_TRY(do_wgrad());
    RD_TRY(do_dgrad());
  }
  return 0;
}

}  // namespace

// Encoder levels >= stage_split belong to backward stage 1, the shallower ones to stage 2.  Parameter counts shrink 4x
// per level towards the input while the backward time per level grows: with the split at level 3 the stage-1 slice
// (bottleneck + deep levels, ~24 MB at depth 5) has the whole shallow backward pass (> 1 ms at batch 64) to hide its
// all-reduce, and the last slice, which nothing can hide, is 1.5 MB (latency-sized).
static int stage_split(const rd_handle* h) { return h->depth - 1 < 3 ? h->depth - 1 : 3; }

// stage 0: last_layer + decoder (down to the first up-conv); 1: bottleneck + encoder levels >= stage_split; 2: the
// shallower encoder levels; -1: everything.  Stages must run in order 0, 1, 2 after one saving forward pass.
static int backward_stages(rd_handle* h, const float* x, const float* dy, int stage, void* stream) {
  if (!h || !dy || !x) return fail("rd_backward: null argument");
  if (!h->G) return fail("rd_backward: gradient arena not bound (rd_bind)");
  if (h->fwd_mode != RD_FWD_TRAIN && h->fwd_mode != RD_FWD_EVAL_SAVE)
    return fail("rd_backward: no saved forward pass (call rd_forward with RD_FWD_TRAIN or RD_FWD_EVAL_SAVE first)");
  RD_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = h->fwd_batch, T = h->fwd_tile, D = h->depth;
  const int C0 = h->cfg.start_kernel;
  const bool all = stage < 0;
  if (!all && stage != h->bw_next_stage)
    return fail("rd_backward_stage: stage %d out of order (expected %d)", stage, h->bw_next_stage);
  bool& gp_bf16 = h->bw_gp_bf16;                         // is the gradient at the current pooled tensor held in gp_b?
  bool& gh_bf16 = h->bw_gh_bf16;                         // is the gradient at the last up-conv's input held in gh_b?
  auto join = [&]() -> int {                             // the caller's stream continues only after every weight gradient
    if (h->overlap && h->overlap_allowed) {
      RD_CUDA(cudaEventRecord(h->ev_join, h->side));
      RD_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
      h->wg_pending[0] = h->wg_pending[1] = false;
    }
    return 0;
  };

  if (all || stage == 0) {
  gp_bf16 = false;
  h->xcol_early = false;
  if (h->overlap && h->overlap_allowed && h->enc[0].bb && h->enc[0].tc_wgrad.valid && h->xcol_b) {
    // The im2col expansion of the network input (first layer's weight gradient) depends on x only: run it on the side
    // stream now, under the decoder's backward pass, instead of at the very end of the step where nothing is left on the
    // main stream to overlap it (0.18 ms of exposed tail at batch 64).
    ConvBlock& b0 = h->enc[0];
    RD_CUDA(cudaEventRecord(h->ev_main, s));
    RD_CUDA(cudaStreamWaitEvent(h->side, h->ev_main, 0));
    const double px = (double)B * T * T;
    ProfScope ps(h, RD_PROF_FIRST_WGRAD, 0.0, px * (4.0 * b0.Cin + 2.0 * b0.Cin * 9), h->side);
    RD_TRY(launch_im2col_first_bf16(x, h->xcol_b, B, b0.Cin, T, T, h->xcol_b_pitch, h->side));
    if (h->tc_gram.valid) RD_TRY(first_layer_gram(h, h->side));
    h->xcol_early = true;
  }
  // last_layer (lib/UNet.py:184,227): du -> gradient at u_{D-1}, which is also the skip gradient of level 0
  {
    const double px = (double)B * T * T;
    ProfScope ps(h, RD_PROF_LAST_BWD, 2.0 * 2.0 * 9.0 * C0 * px, 4.0 * px * (2.0 * C0 + 1.0), s);
    // when the last up-conv takes bf16 operands, the gradient at u_{D-1} (= skip gradient of level 0) is kept in
    // bf16 only: its readers are that up-conv's GEMMs and the BatchNorm backward of encoder level 0
    RD_TRY(launch_conv_last_bwd(h->ups[D - 1].u, dy, h->P + h->last_w, h->ups[D - 1].bb ? nullptr : h->g_skip[0],
                                h->ups[D - 1].bb ? h->gs_b[0] : nullptr, h->G + h->last_w,
                                h->last_b >= 0 ? h->G + h->last_b : nullptr, h->G + h->ups[D - 1].bias, h->scratch,
                                h->scratch_floats, B, T, T, C0, s));
  }
  if (h->ob_gamma >= 0) {
    int nparts = 0;
    RD_TRY(launch_outer_bn_reduce(x, dy, h->ob_mean, h->partials, &nparts, B, h->cfg.n_input_channels, T * T, s));
    RD_TRY(launch_outer_bn_bwd_finalize(h->partials, nparts, h->ob_invstd, h->G + h->ob_gamma, h->G + h->ob_beta, s));
  }
  for (int j = D - 1; j >= 0; --j) {
    UpConv& u = h->ups[j];
    const int Hin = T >> (D - j);
    const float* Gu = h->g_skip[D - 1 - j];                  // gradient at u_j (and at skip a_{D-1-j})
    const float* X = j == 0 ? h->bott.a : h->dec[j - 1].a;   // input of the transposed conv
    const double px = (double)B * Hin * Hin, cc = (double)u.C * u.C;
    gh_bf16 = false;
    // bias gradient = per-channel sum of du_j: produced by the kernel that wrote du_j (last-conv backward for the
    // last level, the decoder conv's dgrad epilogue for the others)
    if (u.bilinear) {
      Gather g1 = gather_plain(Hin, Hin, u.C);
      int S = 0;
      RD_TRY(launch_bilinear_up_adjoint(Gu, h->gt, B, Hin, Hin, u.C, h->tf32() && u.tc, s));
      {
        ProfScope ps(h, RD_PROF_CONVT_WGRAD, 2.0 * cc * px, 4.0 * px * u.C * 2.0, s);
        if (u.tc_wgrad.valid) {
          RD_TRY(launch_gemm_reduce_tc(u.tc_wgrad, s));
          S = u.tc_wgrad.splits;
        } else {
          RD_TRY(launch_gemm_reduce_simt(X, g1, h->gt, B, u.C, h->part, h->part_floats, &S, s));
        }
      }
      RD_TRY(launch_unpack_conv_grad(h->part, S, h->G + u.w, u.C, u.C, 1, s));
      Epilogue e{};
      e.mode = EPI_PLAIN;
      e.out = h->gh;
      {
        ProfScope ps(h, RD_PROF_CONVT_DGRAD, 2.0 * cc * px, 4.0 * px * u.C * 2.0, s);
        if (u.tc) RD_TRY(launch_gemm_rows_tc(u.tc_dgrad, e, nullptr, s));
        else RD_TRY(launch_gemm_rows_simt(h->gt, g1, u.w_nk, B, u.C, e, nullptr, s));
      }
    } else {
    Gather g4 = gather_up2(Hin, Hin, u.C);
    int S = 0;
    const bool ov = h->overlap && h->overlap_allowed;
    cudaStream_t ws = ov ? h->side : s;                 // the gradient at u_j (bf16) is complete on the main stream here
    static const bool wg_first = getenv("RESDEPTH_WG_FIRST") != nullptr;
    auto up_wgrad = [&]() -> int {
      if (ov) {
        RD_CUDA(cudaEventRecord(h->ev_main, s));
        RD_CUDA(cudaStreamWaitEvent(ws, h->ev_main, 0));
      }
      {
        ProfScope ps(h, RD_PROF_CONVT_WGRAD, 2.0 * 4.0 * cc * px, 4.0 * px * u.C * 5.0, ws);
        if (u.tc_wgrad.valid) {
          RD_TRY(launch_gemm_reduce_tc(u.tc_wgrad, ws));
          S = u.tc_wgrad.splits;
        } else {
          RD_TRY(launch_gemm_reduce_simt(Gu, g4, X, B, u.C, h->part, h->part_floats, &S, ws));
        }
      }
      {
        ProfScope ps(h, RD_PROF_UNPACK, 0.0, 4.0 * 4.0 * cc * (S + 1.0), ws);
        RD_TRY(launch_unpack_convt_grad(h->part, S, h->G + u.w, u.C, u.C, ws));
      }
      return 0;
    };
    auto up_dgrad = [&]() -> int {
      Epilogue e{};
      e.mode = EPI_PLAIN;
      // bf16 backward on tcgen05: the gradient at the up-conv's input is only read by the BatchNorm backward of the
      // block that produced that input -- keep it in bf16 only, like the skip and pooled-tensor gradients
      gh_bf16 = u.bb && u.tc && h->gh_b != nullptr;
      e.out = gh_bf16 ? nullptr : h->gh;
      e.out_b = gh_bf16 ? h->gh_b : nullptr;
      ProfScope ps(h, RD_PROF_CONVT_DGRAD, 2.0 * 4.0 * cc * px, 4.0 * px * u.C * 5.0, s);
      if (u.tc) RD_TRY(launch_gemm_rows_tc(u.tc_dgrad, e, nullptr, s));
      else RD_TRY(launch_gemm_rows_simt(Gu, g4, u.w_nk, B, u.C, e, nullptr, s));
      return 0;
    };
    if (ov && !wg_first) {                              // data gradient first (see block_backward)
      RD_TRY(up_dgrad());
      RD_TRY(up_wgrad());
    } else {
      RD_TRY(up_wgrad());
      RD_TRY(up_dgrad());
    }
    }
    if (j == 0) {
      break;                                               // the bottleneck block belongs to stage 1
    } else {
      // du_{j-1} is also the A operand of the next transposed-conv dgrad / wgrad: store it TF32-rounded (fp32
      // backward) or only as bf16 (bf16 backward)
      const bool only_b = skip_grad_bf16(h, D - j);
      RD_TRY(block_backward(h, h->dec[j - 1], gh_bf16 ? GradRef(h->gh_b, 1) : GradRef(h->gh), GradRef(), B, Hin, h->ups[j - 1].u, false,
                            only_b ? nullptr : h->g_skip[D - j], h->tf32() && h->ups[j - 1].tc,
                            h->G + h->ups[j - 1].bias, only_b ? h->gs_b[D - j] : nullptr, s));
    }
  }
  if (!all) { RD_TRY(join()); h->bw_next_stage = 1; return 0; }
  }
  if (all || stage == 1) {
    gp_bf16 = h->gp_b && h->bott.tc;
    RD_TRY(block_backward(h, h->bott, h->bw_gh_bf16 ? GradRef(h->gh_b, 1) : GradRef(h->gh), GradRef(), B, T >> D, h->enc[D - 1].p, false, gp_bf16 ? nullptr : h->gp, 0,
                          nullptr, gp_bf16 ? h->gp_b : nullptr, s));
  }
  for (int i = D - 1; i >= 0; --i) {
    if (!all && (stage == 1) != (i >= stage_split(h))) continue;   // stage 1: deep encoder levels; stage 2: levels below the split
    const int H = T >> i;
    const GradRef gs = skip_grad_bf16(h, i) ? GradRef(h->gs_b[i], 1) : GradRef(h->g_skip[i]);
    const GradRef gp = gp_bf16 ? GradRef(h->gp_b, 1) : GradRef(h->gp);
    const bool out_b = i > 0 && h->gp_b && h->enc[i].tc;
    RD_TRY(block_backward(h, h->enc[i], gs, gp, B, H, i == 0 ? x : h->enc[i - 1].p, i == 0,
                          (i == 0 || out_b) ? nullptr : h->gp, 0, nullptr, out_b ? h->gp_b : nullptr, s));
    gp_bf16 = out_b;
  }
  RD_TRY(join());
  h->bw_next_stage = all ? 0 : (stage == 1 ? 2 : 0);
  return 0;
}

int rd_backward(rd_handle* h, const float* x, const float* dy, void* stream) {
  if (h) h->bw_next_stage = 0;
  return backward_stages(h, x, dy, -1, stream);
}

int rd_backward_stage(rd_handle* h, const float* x, const float* dy, int stage, void* stream) {
  if (stage < 0 || stage > 2) return fail("rd_backward_stage: stage must be 0, 1 or 2 (got %d)", stage);
  return backward_stages(h, x, dy, stage, stream);
}

int rd_grad_stage_range(const rd_handle* h, int stage, int64_t* offset, int64_t* numel) {
  if (!h || stage < 0 || stage > 2 || !offset || !numel) return fail("rd_grad_stage_range: bad argument");
  // parameters are laid out encoder.0 .. encoder.D-1, bottleneck, decoder.*, last_layer (+ outer-skip BatchNorm)
  const long long e_deep = h->enc[stage_split(h)].w, d_first = h->ups[0].w;
  const long long lo = stage == 2 ? 0 : (stage == 1 ? e_deep : d_first);
  const long long hi = stage == 2 ? e_deep : (stage == 1 ? d_first : h->param_floats);
  *offset = lo;
  *numel = hi - lo;
  return 0;
}

int rd_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail("rd_adam_step: null argument");
  if (step < 1) return fail("rd_adam_step: step must be >= 1");
  return launch_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                     reinterpret_cast<cudaStream_t>(stream));
}

int rd_sgd_step(float* params, const float* grads, int64_t n, float lr, float weight_decay, float grad_scale,
                void* stream) {
  if (!params || !grads) return fail("rd_sgd_step: null argument");
  return launch_sgd(params, grads, n, lr, weight_decay, grad_scale, reinterpret_cast<cudaStream_t>(stream));
}

int rd_blend_accumulate(const float* tiles, const float* mean, const float* std, const int32_t* geom, int n, int tile,
                        int stride, double* raster, int rows, int cols, void* stream) {
  if (!tiles || !mean || !std || !geom || !raster) return fail("rd_blend_accumulate: null argument");
  return launch_blend(tiles, mean, std, geom, n, tile, stride, raster, rows, cols, reinterpret_cast<cudaStream_t>(stream));
}

static int debug_gather(int kind, int H, int W, int C, Gather* g) {
  if (kind == 0) *g = gather_conv3x3(H, W, C);
  else if (kind == 1) *g = gather_plain(H, W, C);
  else if (kind == 2) *g = gather_up2(H, W, C);
  else return fail("rd_debug: unknown kind %d", kind);
  return 0;
}

__global__ void debug_sum_splits_kernel(const float* __restrict__ part, int S, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * n + i];
    out[i] = a;
  }
}

int rd_debug_rows(int engine, int kind, const float* src, int batch, int hh, int ww, int c, const float* w_kn,
                  const float* w_nk, int n, float* out, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  Gather g;
  RD_TRY(debug_gather(kind, hh, ww, c, &g));
  Epilogue e{};
  e.mode = EPI_PLAIN;
  e.out = out;
  if (engine == 0) return launch_gemm_rows_simt(src, g, w_kn, batch, n, e, nullptr, s);
  TcRowsPlan plan;
  RD_TRY(tc_make_rows_plan(&plan, src, g, batch, w_nk, n, engine == 2));   // engine 2: src / w_nk are bf16
  return launch_gemm_rows_tc(plan, e, nullptr, s);
}

int rd_debug_reduce(int engine, int kind, const float* src, int batch, int hh, int ww, int c, const float* gm, int n,
                    float* out, float* scratch, int64_t scratch_floats, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  Gather g;
  RD_TRY(debug_gather(kind, hh, ww, c, &g));
  int S = 0;
  if (engine == 0) {
    RD_TRY(launch_gemm_reduce_simt(src, g, gm, batch, n, scratch, (size_t)scratch_floats, &S, s));
  } else {
    TcReducePlan plan;
    RD_TRY(tc_make_reduce_plan(&plan, src, g, batch, gm, n, scratch, (size_t)scratch_floats, engine == 2));
    RD_TRY(launch_gemm_reduce_tc(plan, s));
    S = plan.splits;
  }
  const long long total = (long long)g.ntaps * c * n;
  debug_sum_splits_kernel<<<(int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), 256, 0, s>>>(scratch, S, total, out);
  RD_LAUNCHED();
  return 0;
}

int rd_residuals(const void* raster, int raster_f64, const void* gt, int gt_f64, const uint8_t* mask_gt, int64_t n,
                 double nodata, double* res, uint8_t* valid, void* stream) {
  if (!raster || !gt || !res || !valid || n < 0) return fail("rd_residuals: bad argument");
  return launch_residuals(raster, raster_f64, gt, gt_f64, mask_gt, n, nodata, res, valid,
                          reinterpret_cast<cudaStream_t>(stream));
}

int rd_residual_stats(const double* res, const uint8_t* valid, int64_t n, double threshold, double* out16, void* stream) {
  if (!res || !valid || !out16 || n < 0) return fail("rd_residual_stats: bad argument");
  return residual_statistics(res, valid, n, threshold, out16, reinterpret_cast<cudaStream_t>(stream));
}

int rd_tile_stds(const float* dsm, int rows, int cols, const int32_t* pos, int n, int tile, float nodata, double* stds,
                 void* stream) {
  if (!dsm || !pos || !stds || n < 0 || tile <= 0 || tile > rows || tile > cols) return fail("rd_tile_stds: bad argument");
  return launch_tile_stds(dsm, rows, cols, pos, n, tile, nodata, stds, reinterpret_cast<cudaStream_t>(stream));
}

int rd_freeze_params(rd_handle* h, int on) {
  if (!h) return fail("rd_freeze_params: null handle");
  if (on) ++h->freeze_gen;
  h->frozen = on != 0;
  return 0;
}

int rd_set_overlap(rd_handle* h, int on) {
  if (!h) return fail("rd_set_overlap: null handle");
  h->overlap_allowed = on != 0;
  return 0;
}

int rd_profile_enable(rd_handle* h, int on) {
  if (!h) return fail("rd_profile_enable: null handle");
  RD_CUDA(cudaSetDevice(h->device));
  if (!on) RD_TRY(rd_profile_collect(h));
  h->profiling = on != 0;
  for (int i = 0; i < RD_PROF_NUM; ++i) {
    h->prof_ms[i] = h->prof_flops[i] = h->prof_bytes[i] = 0.0;
    h->prof_launches[i] = h->prof_calls[i] = 0;
  }
  return 0;
}

int rd_profile_collect(rd_handle* h) {
  if (!h) return fail("rd_profile_collect: null handle");
  RD_CUDA(cudaSetDevice(h->device));
  // RESDEPTH_TIMELINE=1: print every bracket as (category, start, end) in ms relative to the first one -- brackets of the
  // main and the side stream share the clock, so this is the step's two-stream timeline (profiles/timeline.py)
  static const bool timeline = getenv("RESDEPTH_TIMELINE") != nullptr;
  if (timeline && !h->prof.empty()) {
    for (auto& r : h->prof) RD_CUDA(cudaEventSynchronize(r.e1));
    for (auto& r : h->prof) {
      float t0 = 0.f, t1 = 0.f;
      RD_CUDA(cudaEventElapsedTime(&t0, h->prof[0].e0, r.e0));
      RD_CUDA(cudaEventElapsedTime(&t1, h->prof[0].e0, r.e1));
      fprintf(stderr, "TL %s %.4f %.4f\n", kProfNames[r.cat], t0, t1);
    }
  }
  for (auto& r : h->prof) {
    RD_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    RD_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    h->prof_ms[r.cat] += ms;
    h->prof_flops[r.cat] += r.flops;
    h->prof_bytes[r.cat] += r.bytes;
    h->prof_launches[r.cat] += r.launches;
    h->prof_calls[r.cat] += 1;
    h->event_pool.push_back(r.e0);
    h->event_pool.push_back(r.e1);
  }
  h->prof.clear();
  return 0;
}

int rd_profile_read(const rd_handle* h, int category, char* name64, double* ms, double* flops, double* bytes,
                    int64_t* launches, int64_t* calls) {
  if (!h || category < 0 || category >= RD_PROF_NUM) return fail("rd_profile_read: bad argument");
  if (name64) { std::strncpy(name64, kProfNames[category], 63); name64[63] = 0; }
  if (ms) *ms = h->prof_ms[category];
  if (flops) *flops = h->prof_flops[category];
  if (bytes) *bytes = h->prof_bytes[category];
  if (launches) *launches = h->prof_launches[category];
  if (calls) *calls = h->prof_calls[category];
  return 0;
}

int rd_make_tiles(const float* dsm_in, const float* dsm_gt, const float* orthos, int rows, int cols,
                  int n_views_total, const int32_t* pos, const int32_t* views, const int32_t* aug, int n, int tile,
                  int n_ortho, int include_dsm, float nodata, float dsm_std, float ortho_std, float dsm_mean_in,
                  float ortho_mean_in, float* input, float* target, uint8_t* mask, float* dsm_mean_out,
                  float* scratch, void* stream) {
  if (!dsm_in || !dsm_gt || !pos || !aug || !input || !target || !mask || !dsm_mean_out || !scratch)
    return fail("rd_make_tiles: null argument");
  return launch_make_tiles(dsm_in, dsm_gt, orthos, rows, cols, n_views_total, pos, views, aug, n, tile, n_ortho,
                           include_dsm, nodata, dsm_std, ortho_std, dsm_mean_in, ortho_mean_in, input, target, mask,
                           dsm_mean_out, scratch, reinterpret_cast<cudaStream_t>(stream));
}

int64_t rd_launch_count(int reset) {
  const long long v = g_launch_count;
  if (reset) g_launch_count = 0;
  return v;
}

const char* rd_bwd_mode_name(const rd_handle* h) {
  if (!h) return "none";
  if (!h->tf32()) return "fp32";
  return h->bwd_bf16 ? "bf16" : "tf32";
}

const char* rd_math_mode_name(const rd_handle* h) {
  if (!h) return "none";
  return h->tf32() ? "tf32 (tcgen05 kind::tf32, fp32 accumulate)" : "fp32 (CUDA-core FMA)";
}

}  // extern "C"
