/*
 * resdepth_b200.h -- C ABI of the B200-native ResDepth hot path.
 *
 * The reference (prs-eth/ResDepth) has no FFI of its own: its hot path is reached through two
 * Python classes, lib.UNet.UNet (lib/UNet.py:104-246) and lib.Trainer.Trainer
 * (lib/Trainer.py:13-318), plus lib.evaluation.predict_linear_blend (lib/evaluation.py:460-513).
 * Each entry point below names the reference statement(s) it replaces.  The Python mirror
 * classes in resdepth_b200/lib/ bind these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; rd_last_error() gives the text
 *     of the last failure on the calling thread.  Nothing throws across the ABI.
 *   - all pointers except handles are DEVICE pointers borrowed from the caller (PyTorch owns
 *     parameters, gradients, optimizer state, inputs and outputs); the library owns only its
 *     workspace (activations, packed weights, partial sums, TMA descriptors).
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     a handle is bound to one device and is not re-entrant.
 *   - activations are fp32; the external tensor layout is the reference's (NCHW contiguous).
 */
#ifndef RESDEPTH_B200_H_
#define RESDEPTH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RD_ABI_VERSION 4

typedef struct rd_handle rd_handle;

enum { RD_ACT_RELU = 0, RD_ACT_LRELU = 1, RD_ACT_PRELU = 2 };      /* lib/UNet.py:27-33 */
enum { RD_MATH_FP32 = 0, RD_MATH_TF32 = 1 };  /* CUDA-core fp32 FMA | tcgen05 kind::tf32 */
enum { RD_UP_TRANSPOSE = 0, RD_UP_BILINEAR = 1 };                  /* lib/UNet.py:17-24 */
/* operand type of the backward GEMMs in RD_MATH_TF32 mode: AUTO = bf16 unless the environment says RESDEPTH_BWD=tf32 */
enum { RD_BWD_AUTO = 0, RD_BWD_TF32 = 1, RD_BWD_BF16 = 2 };
/* rd_forward modes: model.eval() under no_grad | model.train() | model.eval() with autograd recording */
enum { RD_FWD_EVAL = 0, RD_FWD_TRAIN = 1, RD_FWD_EVAL_SAVE = 2 };

/* Constructor arguments of lib.UNet.UNet (lib/UNet.py:105-107) + execution knobs. */
typedef struct rd_config {
  int32_t n_input_channels;   /* 1..8 */
  int32_t start_kernel;       /* multiple of 16 */
  int32_t max_filter_depth;
  int32_t depth;              /* 1..8 */
  int32_t act_encoder, act_decoder, act_bottleneck;   /* RD_ACT_* */
  int32_t do_bn;
  int32_t bias_conv_layer;    /* bias of last_layer, lib/UNet.py:184 */
  int32_t outer_skip;
  int32_t outer_skip_bn;      /* BatchNorm2d(1) on input channel 0 before the outer residual, lib/UNet.py:192-193 */
  int32_t math_mode;          /* RD_MATH_* for the GEMM-shaped layers */
  int32_t up_mode;            /* RD_UP_* */
  int32_t bwd_mode;           /* RD_BWD_*: the reference's autograd computes gradients in fp32/TF32 (lib/Trainer.py:179);
                                 bf16 operands (fp32 accumulation) are this library's faster default */
} rd_config;

int rd_abi_version(void);
const char* rd_last_error(void);

/* UNet.__init__ (lib/UNet.py:104-194): builds the layer plan; no device memory yet. */
int rd_create(const rd_config* cfg, int device, rd_handle** out);
int rd_destroy(rd_handle* h);

/* Flat parameter arena layout, in nn.Module.named_parameters() order.
 * rd_param_info: name (<=63 chars), element count and float offset of parameter `index`.
 * rd_buffer_info: same for the fp32 BatchNorm buffers (running_mean, running_var). */
int rd_num_params(const rd_handle* h);
int rd_param_info(const rd_handle* h, int index, char* name64, int64_t* numel, int64_t* offset);
int64_t rd_param_arena_size(const rd_handle* h);
int rd_num_buffers(const rd_handle* h);
int rd_buffer_info(const rd_handle* h, int index, char* name64, int64_t* numel, int64_t* offset);
int64_t rd_buffer_arena_size(const rd_handle* h);

/* Borrow the caller's arenas (replaces nn.Module parameter/buffer storage and param.grad). */
int rd_bind(rd_handle* h, float* params, float* grads, float* bn_buffers);

/* Select (building it on first use) the workspace layout for batches of `batch` tiles of `tile` x `tile` pixels.
 * A handle keeps up to four layouts alive (training batch, validation batch, a partial last batch ...), so that
 * alternating shapes costs neither a device synchronisation nor a rebuild of the TMA descriptors; the least
 * recently used one is freed beyond that.  rd_workspace_id: id (> 0) of the current layout; rd_workspace_alive:
 * whether the layout with that id still exists -- what a caller that replays a CUDA graph captured on it checks. */
int rd_reserve(rd_handle* h, int batch, int tile, int with_backward);
int64_t rd_workspace_bytes(const rd_handle* h);
int64_t rd_workspace_id(const rd_handle* h);
int rd_workspace_alive(const rd_handle* h, int64_t id);

/* UNet.forward (lib/UNet.py:196-246).  x: [B,C,T,T] fp32 NCHW, y: [B,1,T,T].
 * mode RD_FWD_TRAIN: BatchNorm uses batch statistics, updates running stats in the bound buffer
 * arena (num_batches_tracked is the caller's job) and keeps activations for rd_backward.
 * RD_FWD_EVAL_SAVE: running statistics, activations kept.  RD_FWD_EVAL: inference only. */
int rd_forward(rd_handle* h, const float* x, float* y, int batch, int tile, int mode, void* stream);

/* Trainer._compute_denormalized_loss (lib/Trainer.py:87-100) + the seed of loss.backward()
 * (lib/Trainer.py:179).  mask: uint8 [B,1,T,T]; mean/std: [B]; loss_out: device scalar;
 * dy_out (may be NULL): d loss / d y_pred, [B,1,T,T]. */
int rd_loss(rd_handle* h, const float* y_pred, const float* target, const uint8_t* mask,
            const float* mean, const float* std, float* loss_out, float* dy_out,
            int batch, int tile, void* stream);

/* loss.backward() through the network (lib/Trainer.py:179): dy [B,1,T,T] -> gradients of all
 * parameters, written (not accumulated) into the bound gradient arena.  x is the input of the
 * matching rd_forward call (needed for the first layer's weight gradient). */
int rd_backward(rd_handle* h, const float* x, const float* dy, void* stream);

/* The same backward pass in three consecutive stages, so that a data-parallel caller can start the gradient
 * all-reduce of one stage (lib/Trainer.py has none: the reference is single-device) while the next
 * stage computes.  Stage 0: last_layer + decoder; 1: bottleneck + encoder levels >= min(3, depth-1); 2: the
 * shallower encoder levels (few parameters: the one slice that cannot overlap anything is latency-sized).  Call 0, 1, 2 in order after one saving rd_forward; when a stage returns, `stream` is ordered after
 * every gradient of that stage.  rd_grad_stage_range: the contiguous slice [offset, offset + numel) of the
 * gradient arena (floats) that stage `stage` completes. */
int rd_backward_stage(rd_handle* h, const float* x, const float* dy, int stage, void* stream);
int rd_grad_stage_range(const rd_handle* h, int stage, int64_t* offset, int64_t* numel);

/* torch.optim.Adam.step / SGD.step as built by lib/utils.py:329-334 (coupled L2 decay), over
 * flat arenas of n floats.  step >= 1 is the Adam time step after the increment. */
int rd_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                 float grad_scale, void* stream);
int rd_sgd_step(float* params, const float* grads, int64_t n, float lr, float weight_decay,
                float grad_scale, void* stream);

/* Accumulation loop of predict_linear_blend (lib/evaluation.py:484-511) with
 * denormalize_numpy (lib/data_normalization.py:41-53) and _get_blend_weights
 * (lib/evaluation.py:516-567).  tiles: [n,1,T,T] fp32 predictions; mean/std: [n];
 * geom: int32 [n,6] = (y, x, uly, ulx, lry, lrx); raster: float64 [rows, cols], accumulated. */
int rd_blend_accumulate(const float* tiles, const float* mean, const float* std, const int32_t* geom,
                        int n, int tile, int stride, double* raster, int rows, int cols, void* stream);

/* DsmOrthoDataset.__getitem__ (lib/DsmOrthoDataset.py:161-291, training strategy) for a batch of n tiles, on
 * rasters resident in device memory: crop at pos (y, x), per-tile masked mean-centring / sigma scaling of the
 * DSMs, ortho-image gather + normalisation, loss mask, rot90 / flipud / fliplr augmentation
 * (lib/torch_transforms.py:15-157).  The random decisions are inputs.
 *   dsm_in, dsm_gt: [rows][cols] f32; orthos: planar [n_views_total][rows][cols] f32 (the reference's
 *   np.dstack raster, lib/DsmOrthoDataset.py:293-314, transposed once at upload; NULL when n_ortho == 0)
 *   pos int32 [n][2] = (y, x); views int32 [n][n_ortho] (already permuted); aug int32 [n][3] = (k, vflip, hflip)
 *   dsm_mean_in / ortho_mean_in: user-specified means, or NaN to centre every tile on its own mean
 *   include_dsm: channel 0 of the network input is the DSM ('geom*' configurations)
 *   outputs: input [n][C][T][T] f32, target [n][1][T][T] f32, mask uint8 [n][1][T][T], dsm_mean_out f32 [n];
 *   scratch: 2*n floats. */
int rd_make_tiles(const float* dsm_in, const float* dsm_gt, const float* orthos, int rows, int cols,
                  int n_views_total, const int32_t* pos, const int32_t* views, const int32_t* aug, int n, int tile,
                  int n_ortho, int include_dsm, float nodata, float dsm_std, float ortho_std, float dsm_mean_in,
                  float ortho_mean_in, float* input, float* target, uint8_t* mask, float* dsm_mean_out,
                  float* scratch, void* stream);

/* compute_residuals (lib/evaluation.py:11-37) on device arrays of n elements: a pixel is valid unless
 * gt == nodata, raster == nodata or (mask_gt != NULL and mask_gt == 0); res = raster - gt (float64; the float32
 * difference when both inputs are float32, as numpy would).  raster_f64 / gt_f64: element type of the inputs
 * (0 = float32, 1 = float64).  Outputs: res float64 [n] (0 where invalid), valid uint8 [n]. */
int rd_residuals(const void* raster, int raster_f64, const void* gt, int gt_f64, const uint8_t* mask_gt, int64_t n,
                 double nodata, double* res, uint8_t* valid, void* stream);

/* get_statistics (lib/evaluation.py:51-131) over the valid residuals: out16 (HOST memory) = count_total, diff_max,
 * diff_min, MAE, RMSE, absolute_median, median, NMAD, then count_total, MAE, RMSE, absolute_median, median, NMAD of
 * the residuals truncated to [-threshold, threshold] (lib/evaluation.py:40-48; NaN when threshold <= 0); medians
 * are exact order statistics (radix select), even counts average the two middle values like np.ma.median.
 * Synchronises `stream`. */
int rd_residual_stats(const double* res, const uint8_t* valid, int64_t n, double threshold, double* out16, void* stream);

/* Per-tile part of compute_local_dsm_std_per_centered_patch (lib/utils.py:111-158): for each of the n tiles at
 * pos int32 [n][2] = (y, x) of the device raster dsm [rows][cols], the standard deviation of the valid heights
 * around the tile's own mean, sqrt(sum (x - mean)^2 / (count - 1)), in float64 -> stds [n] (device). */
int rd_tile_stds(const float* dsm, int rows, int cols, const int32_t* pos, int n, int tile, float nodata, double* stds,
                 void* stream);

/* rd_backward runs the weight-gradient GEMMs on a handle-owned side stream (forked after each dz, joined before it
 * returns control of `stream`) when every GEMM of the backward pass takes the bf16 tcgen05 path.  on = 0 serialises
 * everything on the caller's stream (used by the per-kernel profile of bench.py); default 1. */
int rd_set_overlap(rd_handle* h, int on);

/* Inference with constant weights (test.py's tile loop, lib/evaluation.py:38-76; validation, lib/Trainer.py:269-300):
 * while on, the caller promises that the parameter and BatchNorm-buffer arenas do not change, and RD_FWD_EVAL forwards
 * re-use the packed GEMM copies of the weights and the BatchNorm scale / shift vectors of the first such forward after
 * the switch-on instead of rebuilding them every call (one launch over all parameters + one over all BatchNorm
 * layers).  Every switch-on starts a new generation (the next forward packs again), as do rd_bind, a training-mode
 * forward on the same workspace layout and a new workspace layout.  on = 0 (default): every forward re-packs. */
int rd_freeze_params(rd_handle* h, int on);

/* Per-category device timing (CUDA events on the launching stream around the library's own launches).
 * rd_profile_enable(h, 1) starts recording; rd_profile_collect synchronises the recorded events and folds
 * them into per-category totals; rd_profile_read returns one category: total milliseconds, algorithmic
 * FLOPs and algorithmic HBM bytes of the bracketed launches, number of kernel launches and of brackets.
 * rd_profile_enable(h, 0) stops recording and clears the totals. */
enum {
  RD_PROF_CONV_FWD = 0, RD_PROF_CONV_DGRAD, RD_PROF_CONV_WGRAD, RD_PROF_CONVT_FWD, RD_PROF_CONVT_DGRAD,
  RD_PROF_CONVT_WGRAD, RD_PROF_FIRST_FWD, RD_PROF_FIRST_WGRAD, RD_PROF_LAST_FWD, RD_PROF_LAST_BWD,
  RD_PROF_BN_FINALIZE, RD_PROF_BN_ACT_POOL, RD_PROF_BN_BWD_REDUCE, RD_PROF_BN_BWD_APPLY, RD_PROF_PACK,
  RD_PROF_UNPACK, RD_PROF_BIAS_GRAD, RD_PROF_LOSS, RD_PROF_NUM
};
int rd_profile_enable(rd_handle* h, int on);
int rd_profile_collect(rd_handle* h);
int rd_profile_read(const rd_handle* h, int category, char* name64, double* ms, double* flops, double* bytes,
                    int64_t* launches, int64_t* calls);

/* Test hooks: run ONE GEMM-shaped kernel in isolation (used by tests/ to compare the tcgen05 kernels with
 * the CUDA-core kernels and the oracle layer by layer).  engine: 0 = CUDA-core fp32, 1 = tcgen05 TF32,
 * 2 = tcgen05 bf16 (backward GEMMs; src, w_nk and g then point to bf16 tensors, out stays fp32).
 * kind: 0 = 3x3 taps (conv3x3 forward / dgrad / wgrad), 1 = single tap (transposed-conv forward),
 *       2 = 2x2 stride-2 gather (transposed-conv dgrad / wgrad; src is [B, 2H, 2W, C]).
 * rows:   out[B*H*W][N] = gather(src)[.][ntaps*C] x W, with w_kn = [ntaps*C][N] and w_nk = [N][ntaps*C].
 * reduce: out[ntaps*C][N] = sum over the B*H*W pixels of gather(src)[p][.]^T G[p][N] (splits already summed;
 *         scratch holds the split partials). */
int rd_debug_rows(int engine, int kind, const float* src, int batch, int h, int w, int c, const float* w_kn,
                  const float* w_nk, int n, float* out, void* stream);
int rd_debug_reduce(int engine, int kind, const float* src, int batch, int h, int w, int c, const float* g, int n,
                    float* out, float* scratch, int64_t scratch_floats, void* stream);

/* Number of kernels launched by this library since the last call with reset != 0. */
int64_t rd_launch_count(int reset);

/* "bf16", "tf32" or "fp32": operand type of the backward GEMMs of this handle (see rd_config.bwd_mode). */
const char* rd_bwd_mode_name(const rd_handle* h);

/* Debug/profiling: name of the math path actually compiled for the GEMM-shaped layers. */
const char* rd_math_mode_name(const rd_handle* h);

#ifdef __cplusplus
}
#endif
#endif  /* RESDEPTH_B200_H_ */
