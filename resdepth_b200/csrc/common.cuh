// Shared declarations of the resdepth_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace rd {

extern thread_local std::string g_last_error;
extern long long g_launch_count;

int fail(const char* fmt, ...);

#define RD_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return ::rd::fail("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)

#define RD_TRY(expr)            \
  do {                          \
    int r__ = (expr);           \
    if (r__ != 0) return r__;   \
  } while (0)

// count + check a kernel launch (no sync)
#define RD_LAUNCHED()                                                                      \
  do {                                                                                     \
    ++::rd::g_launch_count;                                                                \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return ::rd::fail("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Tensor views (NHWC fp32 unless stated otherwise)
// ---------------------------------------------------------------------------------------------
struct Act {           // per-layer activation parameters
  int kind;            // RD_ACT_*
  const float* slope;  // device pointer to the negative slope (0 for relu, 0.01 lrelu, alpha prelu)
};

// ---- direct (CUDA-core) kernels for the two thin ends of the network -------------------------
// first conv: x NCHW [B,Cin,H,W] (Cin<=8) -> z NHWC [B,H,W,Cout]; per-block channel sums for BN
int launch_conv_first_fwd(const float* x, const float* w_oihw, float* z, float* partials, int* n_partials,
                          int B, int Cin, int H, int W, int Cout, cudaStream_t s);
// wgrad of the first conv: dW[co][ci][r][s] (OIHW, written) from x NCHW and dz NHWC
int launch_conv_first_wgrad(const float* x, const float* dz, float* dw_oihw, float* scratch, size_t scratch_floats,
                            int B, int Cin, int H, int W, int Cout, cudaStream_t s);
// last conv C->1 (+bias +residual x[:,0]):  u NHWC [B,H,W,C] -> y [B,H,W]
//   x_affine (optional, 2 floats: scale, shift): the residual is x0*scale+shift (outer_skip_BN)
//   tap_scratch (optional, 9*B*H*W floats): enables the thread-per-pixel two-pass form for C = 32 / 64
int launch_conv_last_fwd(const float* u, const float* w_oihw, const float* bias, const float* x_nchw, int x_cstride_b,
                         const float* x_affine, float* y, int B, int H, int W, int C, float* tap_scratch, cudaStream_t s);
// backward of the last conv: du NHWC (written), dW (OIHW [1,C,3,3], written), dbias (written if non-null)
//   du_channel_sum (optional, [C]): per-channel sums of du = bias gradient of the transposed conv that produced u
//   du_b (optional): bf16 copy of du
int launch_conv_last_bwd(const float* u, const float* dy, const float* w_oihw, float* du, void* du_b, float* dw,
                         float* dbias, float* du_channel_sum, float* scratch, size_t scratch_floats, int B, int H, int W,
                         int C, cudaStream_t s);

// ---- elementwise / reduction kernels -----------------------------------------------------------
struct BnLayer {
  int C;
  const float* gamma;   // params (null when !do_bn)
  const float* beta;
  const float* conv_bias;  // used when !do_bn
  float* running_mean;  // buffers (null when !do_bn)
  float* running_var;
  float* mean;          // saved batch stats [C]
  float* invstd;        // [C]
  float* scale;         // [C] fused affine:  a = act(z*scale + shift)
  float* shift;         // [C]
};
// partials: [nparts][C][2] (sum, sumsq) -> mean/invstd/scale/shift (+ running stats when training)
int launch_bn_finalize(const BnLayer& L, const float* partials, int nparts, long long count, int training,
                       int do_bn, cudaStream_t s);
// eval-mode finalize (running statistics -> scale / shift) of all layers in one launch
static constexpr int BN_EVAL_MAX_JOBS = 24;
struct BnEvalJob {
  const float *gamma, *beta, *conv_bias, *running_mean, *running_var;
  float *mean, *invstd, *scale, *shift;
  int C, block0;
};
struct BnEvalJobs {
  int n = 0, total_blocks = 0;
  BnEvalJob job[BN_EVAL_MAX_JOBS];
};
int bn_eval_jobs_add(BnEvalJobs& J, const BnLayer& L, int do_bn);
int launch_bn_eval_batched(const BnEvalJobs& J, cudaStream_t s);
// a = act(z*scale+shift)  (+ 2x2 max-pool into p when p != null).  round_*: store TF32-rounded values.
//   a_b / p_b (optional): bf16 copies of a / p (operands of the bf16 backward GEMMs)
int launch_bn_act_pool(const float* z, const float* scale, const float* shift, Act act, float* a, float* p,
                       int B, int H, int W, int C, int round_a, int round_p, void* a_b, void* p_b, cudaStream_t s);
// backward pass 1:  gA = unpool(g_pool, a) + g_full ; gY = gA*act'(y) ; per-block partial sums
//   partials [nblk][C][3] = (sum gY, sum gY*(z-mean), sum gA*min(y,0))
// g_full / g_pool: fp32 tensors, or bf16 tensors when the matching *_bf16 flag is set
//   gy_b (optional, pooled blocks): also store gY as bf16 (first encoder block: its apply pass is replaced by
//   launch_first_grad_correct)
int launch_bn_bwd_reduce(const void* g_full, const void* g_pool, int gf_bf16, int gp_bf16, const float* z,
                         const BnLayer& L, Act act, float* partials, int* n_partials, int B, int H, int W,
                         cudaStream_t s, void* gy_b = nullptr);
// dW of the first encoder conv from X^T gY (already in dw), the Gram matrix of the im2col expansion and the BatchNorm
// backward coefficients (see first_grad_correct_kernel)
int launch_first_grad_correct(float* dw, const float* w, const float* gram, const void* coef, int Co, int K, int Kc,
                              cudaStream_t s);
// finalize: dgamma, dbeta (or conv dbias) -> grads, PReLU slope gradient, coefficients for pass 2 (coef: 4*C floats)
int launch_bn_bwd_finalize(const BnLayer& L, const float* partials, int nparts, long long count, int do_bn,
                           int batch_stats, float* dgamma, float* dbeta, float* dslope, float* dslope_scratch, void* coef,
                           cudaStream_t s);
// pass 2:  dz = cs*(gY - c1 - (z-mean)*c2), gY recomputed from the same inputs as pass 1
//   dz (fp32) and/or dz_b (bf16) receive the result; either may be null
int launch_bn_bwd_apply(const void* g_full, const void* g_pool, int gf_bf16, int gp_bf16, const float* z,
                        const BnLayer& L, Act act, const void* coef, float* dz, void* dz_b, int B, int H, int W, int round_out, cudaStream_t s);

int launch_loss(const float* y_pred, const float* target, const uint8_t* mask, const float* mean, const float* std,
                float* loss_out, float* dy_out, float* scratch, int B, int HW, cudaStream_t s);
static constexpr size_t LOSS_SCRATCH_FLOATS = 2 * 2 * 1184 + 4;
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                float wd, long long step, float gscale, cudaStream_t s);
int launch_sgd(float* p, const float* g, long long n, float lr, float wd, float gscale, cudaStream_t s);
int launch_blend(const float* tiles, const float* mean, const float* std, const int32_t* geom, int n, int T,
                 int stride, double* raster, int rows, int cols, cudaStream_t s);
int launch_fill(float* p, float v, long long n, cudaStream_t s);
// training-tile producer (kernels_tiles.cu); scratch: 2*n floats
int launch_make_tiles(const float* dsm_in, const float* dsm_gt, const float* orthos, int rows, int cols, int nvt,
                      const int32_t* pos, const int32_t* views, const int32_t* aug, int n, int T, int n_ortho,
                      int include_dsm, float nodata, float dsm_std, float ortho_std, float dsm_mean_in,
                      float ortho_mean_in, float* input, float* target, uint8_t* mask, float* dsm_mean_out,
                      float* scratch, cudaStream_t s);
// out[i] = sum over p of part[p*row_stride + i*col_stride], i < n (double accumulation)
int launch_sum_partials(const float* part, int nparts, int n, int row_stride, int col_stride, float* out,
                        cudaStream_t s);

// weight packing (see kernels_elementwise.cu for the layouts); null outputs are skipped
//   dnk_b / kn_b: bf16 copies of the dgrad weight matrices for the bf16 backward GEMMs
int launch_pack_conv3x3(const float* w, float* kn, float* nk, float* dkn, float* dnk, void* dnk_b, int Co, int Ci,
                        int round_tf32, cudaStream_t s);
int launch_pack_convt(const float* w, float* kn, float* nk, void* kn_b, int Ci, int Co, int round_tf32, cudaStream_t s);
// all layers in one launch: collect jobs with pack_jobs_add (same output meanings as the per-layer launchers above;
// conv1x1: o0 = copy, o1 = transpose), then launch_pack_batched
enum { PACK_CONV3X3 = 0, PACK_CONV3X3_TILED = 1, PACK_CONVT = 2, PACK_CONV1X1 = 3 };
static constexpr int PACK_MAX_JOBS = 40;
struct PackJob {
  const float* w;
  float *o0, *o1, *o2, *o3;
  void* ob;
  int kind, Co, Ci, rnd, block0;
};
struct PackJobs {
  int n = 0, total_blocks = 0;
  PackJob job[PACK_MAX_JOBS];
};
int pack_jobs_add(PackJobs& J, int kind, const float* w, float* o0, float* o1, float* o2, float* o3, void* ob, int Co,
                  int Ci, int rnd);
int launch_pack_batched(const PackJobs& J, cudaStream_t s);
// reduce split partials and un-pack to the PyTorch layouts
//   conv: part [S][(t,ci)][co] -> dW OIHW ;  convT: part [S][(a,b,co)][ci] -> dW [ci][co][2][2]
int launch_unpack_conv_grad(const float* part, int S, float* dw, int Co, int Ci, int ntaps, cudaStream_t s);
int launch_unpack_convt_grad(const float* part, int S, float* dw, int Ci, int Co, cudaStream_t s);
// outer_skip_BN (BatchNorm2d(1) on input channel 0): statistics / backward reductions over x[:,0] (NCHW)
//   dy == null: partials (sum x0, sum x0^2) in the [nparts][1][2] layout of launch_bn_finalize
//   dy != null: partials (sum dy, sum dy*(x0-mean))
int launch_outer_bn_reduce(const float* x, const float* dy, const float* mean, float* partials, int* n_partials, int B,
                           int Cin, int HW, cudaStream_t s);
int launch_outer_bn_bwd_finalize(const float* partials, int nparts, const float* invstd, float* dgamma, float* dbeta,
                                 cudaStream_t s);
// up_mode='bilinear': x2 bilinear interpolation (+bias +skip) of the low-resolution 1x1-conv output, and its adjoint
int launch_bilinear_up_add(const float* t, const float* bias, const float* skip, float* u, int B, int Hin, int Win,
                           int C, int round_tf32, cudaStream_t s);
int launch_bilinear_up_adjoint(const float* du, float* dt, int B, int Hin, int Win, int C, int round_tf32,
                               cudaStream_t s);
int launch_pack_conv1x1(const float* w, float* w_copy, float* w_t, int Co, int Ci, int round_tf32, cudaStream_t s);
// first-layer wgrad on tensor cores: im2col expansion of the NCHW input and un-packing of the reduce-GEMM result
int launch_im2col_first(const float* x, float* xcol, int B, int Cin, int H, int W, int Kc, int round_tf32,
                        cudaStream_t s);
int launch_im2col_first_bf16(const float* x, void* xcol, int B, int Cin, int H, int W, int Kc, cudaStream_t s);
int launch_unpack_conv_grad_wide(const float* part, int S, int ldn, float* dw, int Co, int Ci, int by_ci, cudaStream_t s);
// ---- evaluation-side kernels (kernels_stats.cu) ---------------------------------------------------
int launch_residuals(const void* raster, int raster_f64, const void* gt, int gt_f64, const uint8_t* mask_gt, long long n,
                     double nodata, double* res, uint8_t* valid, cudaStream_t s);
int residual_statistics(const double* res, const uint8_t* valid, long long n, double threshold, double* out16,
                        cudaStream_t s);
int launch_tile_stds(const float* dsm, int rows, int cols, const int32_t* pos, int n, int tile, float nodata,
                     double* stds, cudaStream_t s);
int launch_im2col_first_bf16_clear(void* xcol, size_t bytes, cudaStream_t s);
int launch_unpack_first_grad(const float* part, int S, float* dw, int Co, int K, int Kc, cudaStream_t s, int pitch = 0);
// column sums: out[c] = sum over pixels of g[p][c]   (bias gradient of the transposed convs)
int launch_channel_sum(const float* g, long long npix, int C, float* out, float* scratch, size_t scratch_floats,
                       cudaStream_t s);

// ---- GEMM-shaped layers, CUDA-core fp32 path ("exact" mode) ------------------------------------
struct Gather {      // how GEMM rows (output pixels) map to source pixels per tap
  int ntaps;         // 9 (conv3x3), 1 (plain), 4 (2x2 stride-2 gather)
  int ups;           // 1 or 2: source pixel = ups*(h,w) + (dh,dw)
  int Hs, Ws;        // source spatial size
  int Ho, Wo;        // GEMM-row spatial size
  int C;             // channels per tap (source tensor channel count)
  signed char dh[9], dw[9];
};
enum { EPI_PLAIN = 0, EPI_STATS = 1, EPI_CONVT = 2, EPI_BNACT = 3 };
struct Epilogue {
  int mode;
  float* out;             // PLAIN/STATS/BNACT: [M][N] ; CONVT: u NHWC [B,2Ho,2Wo,N/4]
  float* partials;        // STATS: [m_tiles][N][2]
  const float* bias;      // CONVT: [N/4]
  const float* skip;      // CONVT: NHWC same shape as out (may be null)
  // CONVT on tcgen05: when skip_scale != null, `skip` holds the RAW conv output z of the encoder level and the
  // epilogue applies that level's BatchNorm + activation on the fly: skip value = act(z*scale[c] + shift[c])
  // (the training forward then never writes the full-resolution activated encoder tensor)
  const float* skip_scale;
  const float* skip_shift;
  const float* skip_slope;   // device scalar
  int round_tf32;         // round the stored values to TF32 (they feed a tcgen05 GEMM)
  // BNACT (eval-mode BatchNorm folded into the conv, tcgen05 kernel only): out = act(acc*scale + shift),
  // optionally also the 2x2 max-pooled tensor
  const float* scale;     // [N]
  const float* shift;     // [N]
  const float* slope;     // device scalar: negative slope of the activation
  float* pool_out;        // [B,Ho/2,Wo/2,N] or null
  int round_pool;
  void* out_b;            // optional bf16 copy of `out` (tcgen05 kernel only)
};
// C[M=B*Ho*Wo][N] = gather(src)[M][ntaps*C] * Bm[ntaps*C][N]
int launch_gemm_rows_simt(const float* src, const Gather& g, const float* Bm, int B, int N, const Epilogue& e,
                          int* m_tiles_out, cudaStream_t s);
// part[split][tap][Ca][N] = sum over pixels  gather(src)[p][tap][Ca] * G[p][N]
int launch_gemm_reduce_simt(const float* src, const Gather& g, const float* G, int B, int N, float* part,
                            size_t part_floats, int* splits_out, cudaStream_t s);

// ---- GEMM-shaped layers, tcgen05 path ------------------------------------------------------------
bool tc_available();

}  // namespace rd
