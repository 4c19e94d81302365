// Memory-bound kernels of the ResDepth hot path (NHWC fp32, 128-bit accesses):
//   * BatchNorm statistics finalize + running-stat update       (nn.BatchNorm2d, lib/UNet.py:45,66,86)
//   * fused BN-apply + activation (+ 2x2 max-pool dual write)   (lib/UNet.py:27-33,161,167)
//   * block backward: un-pool + skip-grad add + act' + BN backward (two passes)
//   * masked denormalised L1 loss + gradient seed               (lib/Trainer.py:87-100,179)
//   * Adam / SGD over a flat arena                              (lib/utils.py:329-334, lib/Trainer.py:218)
//   * linear-blend accumulation                                 (lib/evaluation.py:484-567)
//   * weight packing / gradient un-packing for the GEMM-shaped layers
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "../../include/resdepth_b200.h"

namespace rd {

static constexpr int EW_THREADS = 256;
static constexpr float BN_EPS = 1e-5f;
static constexpr float BN_MOMENTUM = 0.1f;

__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_rn4(float4 v) {
  return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
}
__device__ __forceinline__ float act1(float y, float slope) { return y > 0.f ? y : y * slope; }
// 4 floats -> 4 bf16 (round to nearest even), one 8-byte store
__device__ __forceinline__ void st4_bf16(void* base, size_t elem, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem) = pk;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 4 consecutive elements of a gradient tensor stored as fp32 or (bf16 != 0) as bf16
__device__ __forceinline__ float4 ld4g(const void* base, size_t elem, int bf16) {
  if (!bf16) return ld4(reinterpret_cast<const float*>(base) + elem);
  const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem);
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}

static int ew_grid(long long work_items) {
  long long g = (work_items + EW_THREADS - 1) / EW_THREADS;
  const long long cap = 148LL * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ----------------------------------------------------------------------------------------------
// BN finalize: partials [nparts][C][2] (sum, sumsq) -> mean / invstd / scale / shift
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count, int training, int do_bn,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ conv_bias, float* __restrict__ running_mean,
                   float* __restrict__ running_var, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                   float* __restrict__ scale, float* __restrict__ shift) {
  __shared__ double red[32][33][2];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (do_bn && training && c < C) {
    // four independent loads in flight per thread: with one, the ~20 iterations of a 592-row partial table were 20
    // serialised L2 round trips (10 us for a kernel that moves 300 KB; timeline of round 2)
    int p = pl;
    for (; p + 96 < nparts; p += 128) {
      const float2 v0 = __ldg(reinterpret_cast<const float2*>(partials + ((size_t)p * C + c) * 2));
      const float2 v1 = __ldg(reinterpret_cast<const float2*>(partials + ((size_t)(p + 32) * C + c) * 2));
      const float2 v2 = __ldg(reinterpret_cast<const float2*>(partials + ((size_t)(p + 64) * C + c) * 2));
      const float2 v3 = __ldg(reinterpret_cast<const float2*>(partials + ((size_t)(p + 96) * C + c) * 2));
      s1 += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
      s2 += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
    }
    for (; p < nparts; p += 32) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partials + ((size_t)p * C + c) * 2));
      s1 += (double)v.x;
      s2 += (double)v.y;
    }
  }
  red[pl][cl][0] = s1;
  red[pl][cl][1] = s2;
  __syncthreads();
  if (pl != 0 || c >= C) return;
  if (!do_bn) {
    scale[c] = 1.f;
    shift[c] = conv_bias ? conv_bias[c] : 0.f;
    return;
  }
  float mean, invstd;
  if (training) {
#pragma unroll
    for (int i = 1; i < 32; ++i) { s1 += red[i][cl][0]; s2 += red[i][cl][1]; }
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
    const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
    running_mean[c] = (1.f - BN_MOMENTUM) * running_mean[c] + BN_MOMENTUM * mean;
    running_var[c] = (1.f - BN_MOMENTUM) * running_var[c] + BN_MOMENTUM * (float)unbiased;
  } else {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + BN_EPS);
  }
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
}

int launch_bn_finalize(const BnLayer& L, const float* partials, int nparts, long long count, int training,
                       int do_bn, cudaStream_t s) {
  bn_finalize_kernel<<<cdiv(L.C, 32), 1024, 0, s>>>(partials, nparts, L.C, (double)count, training, do_bn, L.gamma,
                                                   L.beta, L.conv_bias, L.running_mean, L.running_var, L.mean,
                                                   L.invstd, L.scale, L.shift);
  RD_LAUNCHED();
  return 0;
}

// Eval-mode BatchNorm of EVERY layer in one launch: scale / shift from the running statistics (no batch data involved),
// so an inference forward pass does not pay one tiny finalize launch per layer (10 x ~7 us of 2.2 ms at batch 32).
__global__ void __launch_bounds__(256) bn_eval_batched_kernel(const __grid_constant__ BnEvalJobs J) {
  int j = 0;
  while (j + 1 < J.n && (int)blockIdx.x >= J.job[j + 1].block0) ++j;
  const BnEvalJob& L = J.job[j];
  const int c = ((int)blockIdx.x - L.block0) * 256 + threadIdx.x;
  if (c >= L.C) return;
  if (!L.gamma) {                                       // no BatchNorm: identity scale, the conv bias as shift
    L.scale[c] = 1.f;
    L.shift[c] = L.conv_bias ? L.conv_bias[c] : 0.f;
    return;
  }
  const float mean = L.running_mean[c];
  const float invstd = 1.f / sqrtf(L.running_var[c] + BN_EPS);
  L.mean[c] = mean;
  L.invstd[c] = invstd;
  const float sc = L.gamma[c] * invstd;
  L.scale[c] = sc;
  L.shift[c] = L.beta[c] - mean * sc;
}
int bn_eval_jobs_add(BnEvalJobs& J, const BnLayer& L, int do_bn) {
  if (J.n >= BN_EVAL_MAX_JOBS) return fail("bn_eval: more than %d layers", BN_EVAL_MAX_JOBS);
  BnEvalJob& j = J.job[J.n++];
  j.C = L.C;
  j.gamma = do_bn ? L.gamma : nullptr; j.beta = L.beta; j.conv_bias = L.conv_bias;
  j.running_mean = L.running_mean; j.running_var = L.running_var;
  j.mean = L.mean; j.invstd = L.invstd; j.scale = L.scale; j.shift = L.shift;
  j.block0 = J.total_blocks;
  J.total_blocks += (L.C + 255) / 256;
  return 0;
}
int launch_bn_eval_batched(const BnEvalJobs& J, cudaStream_t s) {
  if (J.n == 0) return 0;
  bn_eval_batched_kernel<<<J.total_blocks, 256, 0, s>>>(J);
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// a = act(z*scale + shift), optional 2x2 max-pool (first-max-wins is irrelevant in the forward)
// ----------------------------------------------------------------------------------------------
template <bool POOL>
__global__ void __launch_bounds__(EW_THREADS)
bn_act_pool_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ slope_p, float* __restrict__ a, float* __restrict__ p, int B, int H,
                   int W, int C, int round_a, int round_p, void* __restrict__ a_b, void* __restrict__ p_b) {
  const int Q = C >> 2;
  const float slope = *slope_p;
  if (POOL) {
    const int Hp = H >> 1, Wp = W >> 1;
    const long long total = (long long)B * Hp * Wp * Q;
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
      const int q = (int)(i % Q);
      long long w_ = i / Q;
      const int wp = (int)(w_ % Wp); w_ /= Wp;
      const int hp = (int)(w_ % Hp);
      const int b = (int)(w_ / Hp);
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + q);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + q);
      const size_t base = (((size_t)b * H + 2 * hp) * W + 2 * wp) * C + q * 4;
      float4 m;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t o = base + ((size_t)(k >> 1) * W + (k & 1)) * C;
        const float4 v = ld4(z + o);
        float4 r;
        r.x = act1(fmaf(v.x, sc.x, sh.x), slope);
        r.y = act1(fmaf(v.y, sc.y, sh.y), slope);
        r.z = act1(fmaf(v.z, sc.z, sh.z), slope);
        r.w = act1(fmaf(v.w, sc.w, sh.w), slope);
        if (a) st4(a + o, round_a ? tf32_rn4(r) : r);      // a == null: consumers re-derive it from z (CONVT epilogue)
        if (a_b) st4_bf16(a_b, o, r);
        if (k == 0) m = r;
        else { m.x = fmaxf(m.x, r.x); m.y = fmaxf(m.y, r.y); m.z = fmaxf(m.z, r.z); m.w = fmaxf(m.w, r.w); }
      }
      const size_t po = (((size_t)b * Hp + hp) * Wp + wp) * C + q * 4;
      st4(p + po, round_p ? tf32_rn4(m) : m);
      if (p_b) st4_bf16(p_b, po, m);
    }
  } else {
    const long long total = (long long)B * H * W * Q;
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
      const int q = (int)(i % Q);
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + q);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + q);
      const float4 v = ld4(z + i * 4);
      float4 r;
      r.x = act1(fmaf(v.x, sc.x, sh.x), slope);
      r.y = act1(fmaf(v.y, sc.y, sh.y), slope);
      r.z = act1(fmaf(v.z, sc.z, sh.z), slope);
      r.w = act1(fmaf(v.w, sc.w, sh.w), slope);
      st4(a + i * 4, round_a ? tf32_rn4(r) : r);
      if (a_b) st4_bf16(a_b, (size_t)i * 4, r);
    }
  }
}

int launch_bn_act_pool(const float* z, const float* scale, const float* shift, Act act, float* a, float* p, int B,
                       int H, int W, int C, int round_a, int round_p, void* a_b, void* p_b, cudaStream_t s) {
  if (C % 4) return fail("bn_act_pool: C=%d not a multiple of 4", C);
  if (p) {
    if ((H | W) & 1) return fail("bn_act_pool: odd size %dx%d cannot be pooled", H, W);
    const long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
    bn_act_pool_kernel<true><<<ew_grid(total), EW_THREADS, 0, s>>>(z, scale, shift, act.slope, a, p, B, H, W, C,
                                                                    round_a, round_p, a_b, p_b);
  } else {
    const long long total = (long long)B * H * W * (C / 4);
    bn_act_pool_kernel<false><<<ew_grid(total), EW_THREADS, 0, s>>>(z, scale, shift, act.slope, a, p, B, H, W, C,
                                                                     round_a, round_p, a_b, p_b);
  }
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Block backward.  With y = z*scale+shift, a = act(y):
//   gA = unpool_first_max(g_pool, a) + g_full ;  gY = gA * act'(y)
//   pass 1 (reduce): per-channel  S1 = sum gY,  S2 = sum gY*(z-mean),  S3 = sum gA*min(y,0)   (PReLU slope grad)
//   pass 2 (apply) : dz = cs * (gY - c1 - (z-mean)*c2)
// Threads keep a fixed channel quad so the sums stay in registers; one thread handles one 2x2 window (pooled
// layers) or one pixel (others) per iteration.
// ----------------------------------------------------------------------------------------------
struct BwdCoef {            // per channel, built by bn_bwd_finalize
  float mean, cs, c1, c2;
};

// RELU: the activation is a plain ReLU (slope 0, no slope gradient) -- the common case, with a shorter inner loop
// STORE_GY (reduce pass only): also store gY (bf16).  Used for the FIRST encoder block, whose dz feeds nothing but the
// weight gradient: there dW = X^T dz is rebuilt from X^T gY and small correction terms (first_grad_correct_kernel), and
// the second full pass over z (the apply pass) is not run at all.
template <bool POOL, bool APPLY, bool RELU, bool STORE_GY = false>
__global__ void __launch_bounds__(EW_THREADS, APPLY ? 4 : 3)
bn_bwd_kernel(const void* __restrict__ g_full, const void* __restrict__ g_pool, int gf_bf16, int gp_bf16,
              const float* __restrict__ z,
              const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ slope_p,
              const float* __restrict__ mean, const BwdCoef* __restrict__ coef, float* __restrict__ dz,
              float* __restrict__ partials, int B, int H, int W, int C, int round_out, void* __restrict__ dz_b) {
  extern __shared__ float red[];                       // reduce: [PL][C][3]
  const int Q = C >> 2;
  const int PL = EW_THREADS / Q;                       // pixel lanes per block
  const int q = threadIdx.x % Q, pl = threadIdx.x / Q;
  const float slope = RELU ? 0.f : *slope_p;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, s3[4] = {0, 0, 0, 0};
  if (pl < PL) {
    const float4 sc4 = ld4(scale + q * 4), sh4 = ld4(shift + q * 4);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
    float mu[4], cs[4], c1[4], c2[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (APPLY) {
        const BwdCoef k = coef[q * 4 + c];
        mu[c] = k.mean; cs[c] = k.cs; c1[c] = k.c1; c2[c] = k.c2;
      } else {
        mu[c] = mean[q * 4 + c]; cs[c] = c1[c] = c2[c] = 0.f;
      }
    }
    const int Hw = POOL ? (H >> 1) : H, Ww = POOL ? (W >> 1) : W;
    const long long nwin = (long long)B * Hw * Ww;
    for (long long wi = (long long)blockIdx.x * PL + pl; wi < nwin; wi += (long long)gridDim.x * PL) {
      if (POOL) {
        const unsigned wiu = (unsigned)wi;             // launchers guarantee B*H*W < 2^32: 32-bit index math
        const unsigned t_ = wiu / (unsigned)Ww;
        const int wp = (int)(wiu - t_ * (unsigned)Ww);
        const int b = (int)(t_ / (unsigned)Hw);
        const int hp = (int)(t_ - (unsigned)b * (unsigned)Hw);
        const size_t base = (((size_t)b * H + 2 * hp) * W + 2 * wp) * C + q * 4;
        float zz[4][4], yy[4][4], gf[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const size_t o = base + ((size_t)(k >> 1) * W + (k & 1)) * C;
          const float4 v = ld4(z + o);
          zz[k][0] = v.x; zz[k][1] = v.y; zz[k][2] = v.z; zz[k][3] = v.w;
          if (g_full) {
            const float4 g = ld4g(g_full, o, gf_bf16);
            gf[k][0] = g.x; gf[k][1] = g.y; gf[k][2] = g.z; gf[k][3] = g.w;
          } else {
            gf[k][0] = gf[k][1] = gf[k][2] = gf[k][3] = 0.f;
          }
        }
        const float4 gp4 = ld4g(g_pool, (size_t)wi * C + q * 4, gp_bf16);
        const float gp[4] = {gp4.x, gp4.y, gp4.z, gp4.w};
        float out[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int arg = 0;
          float best = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            yy[k][c] = fmaf(zz[k][c], sc[c], sh[c]);
            // ReLU: the first maximum of max(y,0) is the first maximum of y wherever the window's gradient survives
            // the ReLU mask (an all-nonpositive window has gY = 0 at every site)
            const float av = RELU ? yy[k][c] : act1(yy[k][c], slope);
            if (k == 0 || av > best) { best = av; arg = k; }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float gA = gf[k][c] + (k == arg ? gp[c] : 0.f);
            const float y = yy[k][c];
            const float gY = RELU ? (y > 0.f ? gA : 0.f) : (y > 0.f ? gA : gA * slope);
            const float zc = zz[k][c] - mu[c];
            if (APPLY) {
              out[k][c] = cs[c] * (gY - c1[c] - zc * c2[c]);
            } else {
              if (STORE_GY) out[k][c] = gY;
              s1[c] += gY;
              s2[c] = fmaf(gY, zc, s2[c]);
              if (!RELU) s3[c] += y > 0.f ? 0.f : gA * y;
            }
          }
        }
        if (APPLY || STORE_GY) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const size_t o = base + ((size_t)(k >> 1) * W + (k & 1)) * C;
            float4 r = make_float4(out[k][0], out[k][1], out[k][2], out[k][3]);
            if (dz) st4(dz + o, round_out ? tf32_rn4(r) : r);
            if (dz_b) st4_bf16(dz_b, o, r);
          }
        }
      } else {
        const size_t o = (size_t)wi * C + q * 4;
        const float4 v = ld4(z + o), g = ld4g(g_full, o, gf_bf16);
        const float zz[4] = {v.x, v.y, v.z, v.w}, gA[4] = {g.x, g.y, g.z, g.w};
        float out[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float y = fmaf(zz[c], sc[c], sh[c]);
          const float gY = RELU ? (y > 0.f ? gA[c] : 0.f) : (y > 0.f ? gA[c] : gA[c] * slope);
          const float zc = zz[c] - mu[c];
          if (APPLY) {
            out[c] = cs[c] * (gY - c1[c] - zc * c2[c]);
          } else {
            if (STORE_GY) out[c] = gY;
            s1[c] += gY;
            s2[c] = fmaf(gY, zc, s2[c]);
            if (!RELU) s3[c] += y > 0.f ? 0.f : gA[c] * y;
          }
        }
        if (APPLY || STORE_GY) {
          float4 r = make_float4(out[0], out[1], out[2], out[3]);
          if (dz) st4(dz + o, round_out ? tf32_rn4(r) : r);
          if (dz_b) st4_bf16(dz_b, o, r);
        }
      }
    }
  }
  if (!APPLY) {
    if (pl < PL) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* r = red + ((size_t)pl * C + q * 4 + c) * 3;
        r[0] = s1[c]; r[1] = s2[c]; r[2] = s3[c];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 3; i += EW_THREADS) {
      float acc = 0.f;
      for (int p = 0; p < PL; ++p) acc += red[(size_t)p * C * 3 + i];
      partials[(size_t)blockIdx.x * C * 3 + i] = acc;
    }
  }
}

static int bwd_grid(long long nwin, int PL, int per_sm = 4) {
  long long g = (nwin + PL - 1) / PL;
  if (g > 148 * per_sm) g = 148 * per_sm;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_bn_bwd_reduce(const void* g_full, const void* g_pool, int gf_bf16, int gp_bf16, const float* z,
                         const BnLayer& L, Act act, float* partials, int* n_partials, int B, int H, int W,
                         cudaStream_t s, void* gy_b) {
  const int C = L.C, Q = C / 4;
  if (C % 4 || Q > EW_THREADS) return fail("bn_bwd: unsupported C=%d", C);
  if ((long long)B * H * W >= (1LL << 32)) return fail("bn_bwd: more than 2^32 pixels");
  const int PL = EW_THREADS / Q;
  const size_t smem = (size_t)PL * C * 3 * sizeof(float);
  int grid;
  const bool relu = act.kind == RD_ACT_RELU;
  if (g_pool) {
    grid = bwd_grid((long long)B * (H / 2) * (W / 2), PL, 3);     // 3 resident CTAs per SM: one wave
#define RD_BWD_R(RL, ST)                                                                                          \
  {                                                                                                               \
    RD_CUDA(cudaFuncSetAttribute(bn_bwd_kernel<true, false, RL, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 64 * 1024));                                                                     \
    bn_bwd_kernel<true, false, RL, ST><<<grid, EW_THREADS, smem, s>>>(g_full, g_pool, gf_bf16, gp_bf16, z, L.scale, \
                                                                  L.shift, act.slope, L.mean, nullptr, nullptr,   \
                                                                  partials, B, H, W, C, 0, gy_b);                 \
  }
    if (gy_b) { if (relu) RD_BWD_R(true, true) else RD_BWD_R(false, true) }
    else { if (relu) RD_BWD_R(true, false) else RD_BWD_R(false, false) }
#undef RD_BWD_R
  } else {
    if (!g_full) return fail("bn_bwd: no incoming gradient");
    if (gy_b) return fail("bn_bwd: the gY-storing reduce pass exists for pooled blocks only");
    grid = bwd_grid((long long)B * H * W, PL, 3);
#define RD_BWD_R(RL)                                                                                              \
  {                                                                                                               \
    RD_CUDA(cudaFuncSetAttribute(bn_bwd_kernel<false, false, RL>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                 64 * 1024));                                                                     \
    bn_bwd_kernel<false, false, RL><<<grid, EW_THREADS, smem, s>>>(g_full, nullptr, gf_bf16, 0, z, L.scale,       \
                                                                   L.shift, act.slope, L.mean, nullptr, nullptr,  \
                                                                   partials, B, H, W, C, 0, nullptr);             \
  }
    if (relu) RD_BWD_R(true) else RD_BWD_R(false)
#undef RD_BWD_R
  }
  RD_LAUNCHED();
  *n_partials = grid;
  return 0;
}

// partials [nparts][C][3] -> dgamma/dbeta (or conv dbias), PReLU slope grad partial, pass-2 coefficients
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count, int do_bn,
                       int batch_stats, const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ dslope_part, BwdCoef* __restrict__ coef) {
  __shared__ double red[32][33][3];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < C) {
    int p = pl;
    for (; p + 96 < nparts; p += 128) {                 // four partial rows in flight per thread (see bn_finalize_kernel)
      const float* v0 = partials + ((size_t)p * C + c) * 3;
      const float* v1 = partials + ((size_t)(p + 32) * C + c) * 3;
      const float* v2 = partials + ((size_t)(p + 64) * C + c) * 3;
      const float* v3 = partials + ((size_t)(p + 96) * C + c) * 3;
      const float a0 = __ldg(v0), a1 = __ldg(v0 + 1), a2 = __ldg(v0 + 2), b0 = __ldg(v1), b1 = __ldg(v1 + 1), b2 = __ldg(v1 + 2);
      const float c0 = __ldg(v2), c1 = __ldg(v2 + 1), c2 = __ldg(v2 + 2), d0 = __ldg(v3), d1 = __ldg(v3 + 1), d2 = __ldg(v3 + 2);
      s1 += ((double)a0 + (double)b0) + ((double)c0 + (double)d0);
      s2 += ((double)a1 + (double)b1) + ((double)c1 + (double)d1);
      s3 += ((double)a2 + (double)b2) + ((double)c2 + (double)d2);
    }
    for (; p < nparts; p += 32) {
      const float* v = partials + ((size_t)p * C + c) * 3;
      s1 += (double)__ldg(v); s2 += (double)__ldg(v + 1); s3 += (double)__ldg(v + 2);
    }
  }
  red[pl][cl][0] = s1; red[pl][cl][1] = s2; red[pl][cl][2] = s3;
  __syncthreads();
  if (pl != 0) return;
#pragma unroll
  for (int i = 1; i < 32; ++i) { s1 += red[i][cl][0]; s2 += red[i][cl][1]; s3 += red[i][cl][2]; }
  if (c < C) {
    BwdCoef k;
    if (do_bn) {
      const double is = (double)invstd[c];
      const double dg = s2 * is;                     // sum gY * xhat
      dgamma[c] = (float)dg;
      dbeta[c] = (float)s1;
      k.mean = mean[c];
      k.cs = gamma[c] * invstd[c];
      k.c1 = (float)(s1 / count);
      k.c2 = (float)(is * is * s2 / count);          // xhat*dgamma/N = (z-mean)*invstd^2*S2/N
      if (!batch_stats) k.c1 = k.c2 = 0.f;           // eval-mode BN: statistics are constants
    } else {
      dbeta[c] = (float)s1;                          // gradient of the conv bias
      k.mean = 0.f; k.cs = 1.f; k.c1 = 0.f; k.c2 = 0.f;
    }
    coef[c] = k;
  }
  // slope gradient: sum over channels of S3 (one value per block; reduced by the caller-side kernel below)
  if (dslope_part) {
    double v = c < C ? s3 : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (cl == 0) dslope_part[blockIdx.x] = (float)v;
  }
}

__global__ void sum_small_kernel(const float* __restrict__ in, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) a += (double)in[i];
    out[0] = (float)a;
  }
}

int launch_bn_bwd_finalize(const BnLayer& L, const float* partials, int nparts, long long count, int do_bn,
                           int batch_stats, float* dgamma, float* dbeta, float* dslope, float* dslope_scratch, void* coef,
                           cudaStream_t s) {
  const int nb = cdiv(L.C, 32);
  bn_bwd_finalize_kernel<<<nb, 1024, 0, s>>>(partials, nparts, L.C, (double)count, do_bn, batch_stats, L.gamma, L.mean,
                                            L.invstd, dgamma, dbeta, dslope ? dslope_scratch : nullptr,
                                            reinterpret_cast<BwdCoef*>(coef));
  RD_LAUNCHED();
  if (dslope) {
    sum_small_kernel<<<1, 32, 0, s>>>(dslope_scratch, nb, dslope);
    RD_LAUNCHED();
  }
  return 0;
}

int launch_bn_bwd_apply(const void* g_full, const void* g_pool, int gf_bf16, int gp_bf16, const float* z,
                        const BnLayer& L, Act act, const void* coef, float* dz, void* dz_b, int B, int H, int W, int round_out, cudaStream_t s) {
  const int C = L.C, Q = C / 4;
  const int PL = EW_THREADS / Q;
  const bool relu = act.kind == RD_ACT_RELU;
  const BwdCoef* cf = reinterpret_cast<const BwdCoef*>(coef);
  if (g_pool) {
    const int grid = bwd_grid((long long)B * (H / 2) * (W / 2), PL) * 2;
    if (relu)
      bn_bwd_kernel<true, true, true><<<grid, EW_THREADS, 0, s>>>(g_full, g_pool, gf_bf16, gp_bf16, z, L.scale, L.shift,
                                                                  act.slope, L.mean, cf, dz, nullptr, B, H, W, C,
                                                                  round_out, dz_b);
    else
      bn_bwd_kernel<true, true, false><<<grid, EW_THREADS, 0, s>>>(g_full, g_pool, gf_bf16, gp_bf16, z, L.scale, L.shift,
                                                                   act.slope, L.mean, cf, dz, nullptr, B, H, W, C,
                                                                   round_out, dz_b);
  } else {
    const int grid = bwd_grid((long long)B * H * W, PL) * 2;
    if (relu)
      bn_bwd_kernel<false, true, true><<<grid, EW_THREADS, 0, s>>>(g_full, nullptr, gf_bf16, 0, z, L.scale, L.shift,
                                                                   act.slope, L.mean, cf, dz, nullptr, B, H, W, C,
                                                                   round_out, dz_b);
    else
      bn_bwd_kernel<false, true, false><<<grid, EW_THREADS, 0, s>>>(g_full, nullptr, gf_bf16, 0, z, L.scale, L.shift,
                                                                    act.slope, L.mean, cf, dz, nullptr, B, H, W, C,
                                                                    round_out, dz_b);
  }
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Loss (lib/Trainer.py:87-100 with lib/data_normalization.py:29-38 and L1Loss(mean)):
//   yp = y_pred*std_i + mean_i ; yt = y*std_i + mean_i   (separate fp32 multiply and add, as torch does)
//   loss = mean(|yp - yt| over ALL pixels, masked ones zeroed) * numel / sum(mask)
//   d loss / d y_pred = sign(yp - yt) * mask * std_i / sum(mask)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loss_partial_kernel(const float* __restrict__ y_pred, const float* __restrict__ target,
                    const uint8_t* __restrict__ mask, const float* __restrict__ mean, const float* __restrict__ std,
                    double* __restrict__ part, int B, int HW) {
  __shared__ double r1[256], r2[256];
  double a = 0.0, m = 0.0;
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int b = (int)(i / HW);
    if (mask[i]) {
      const float sd = std[b], mu = mean[b];
      const float yp = __fadd_rn(__fmul_rn(y_pred[i], sd), mu);
      const float yt = __fadd_rn(__fmul_rn(target[i], sd), mu);
      a += (double)fabsf(__fsub_rn(yp, yt));
      m += 1.0;
    }
  }
  r1[threadIdx.x] = a; r2[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[blockIdx.x * 2] = r1[0]; part[blockIdx.x * 2 + 1] = r2[0]; }
}

__global__ void __launch_bounds__(256)
loss_finalize_kernel(const double* __restrict__ part, int nparts, float* __restrict__ loss_out,
                     float* __restrict__ inv_count) {
  __shared__ double ra[256], rm[256];
  double a = 0.0, m = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) { a += part[i * 2]; m += part[i * 2 + 1]; }
  ra[threadIdx.x] = a; rm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {                  // fixed-order tree: deterministic
    if (threadIdx.x < o) { ra[threadIdx.x] += ra[threadIdx.x + o]; rm[threadIdx.x] += rm[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss_out[0] = (float)(ra[0] / rm[0]);
    inv_count[0] = (float)(1.0 / rm[0]);
  }
}

__global__ void __launch_bounds__(256)
loss_grad_kernel(const float* __restrict__ y_pred, const float* __restrict__ target, const uint8_t* __restrict__ mask,
                 const float* __restrict__ mean, const float* __restrict__ std, const float* __restrict__ inv_count,
                 float* __restrict__ dy, int B, int HW) {
  const float ic = inv_count[0];
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int b = (int)(i / HW);
    float g = 0.f;
    if (mask[i]) {
      const float sd = std[b], mu = mean[b];
      const float d = __fsub_rn(__fadd_rn(__fmul_rn(y_pred[i], sd), mu), __fadd_rn(__fmul_rn(target[i], sd), mu));
      g = d > 0.f ? sd * ic : (d < 0.f ? -sd * ic : 0.f);
    }
    dy[i] = g;
  }
}

int launch_loss(const float* y_pred, const float* target, const uint8_t* mask, const float* mean, const float* std,
                float* loss_out, float* dy_out, float* scratch, int B, int HW, cudaStream_t s) {
  // scratch: >= 2*LOSS_BLOCKS doubles + 2 floats
  const int nb = 1184;
  double* part = reinterpret_cast<double*>(scratch);
  float* inv_count = scratch + 2 * 2 * nb;
  loss_partial_kernel<<<nb, 256, 0, s>>>(y_pred, target, mask, mean, std, part, B, HW);
  RD_LAUNCHED();
  loss_finalize_kernel<<<1, 256, 0, s>>>(part, nb, loss_out, inv_count);
  RD_LAUNCHED();
  if (dy_out) {
    loss_grad_kernel<<<ew_grid((long long)B * HW), 256, 0, s>>>(y_pred, target, mask, mean, std, inv_count, dy_out, B,
                                                                HW);
    RD_LAUNCHED();
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Optimizers over a flat arena (torch.optim.Adam / SGD with coupled L2 decay, lib/utils.py:329-334)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float b1, float b2, float eps, float wd, float step_size, float inv_bc2_sqrt, float gscale) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 pp = ld4(p + i * 4), gg = ld4(g + i * 4), mm = ld4(m + i * 4), vv = ld4(v + i * 4);
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = fmaf(wd, P[k], G[k] * gscale);
      M[k] = M[k] + (1.f - b1) * (gr - M[k]);
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[k]) * inv_bc2_sqrt + eps;
      P[k] = P[k] - step_size * (M[k] / denom);
    }
    st4(p + i * 4, pp); st4(m + i * 4, mm); st4(v + i * 4, vv);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float gr = fmaf(wd, p[i], g[i] * gscale);
    const float mk = m[i] + (1.f - b1) * (gr - m[i]);
    const float vk = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mk; v[i] = vk;
    p[i] = p[i] - step_size * (mk / (sqrtf(vk) * inv_bc2_sqrt + eps));
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                float wd, long long step, float gscale, cudaStream_t s) {
  if (n <= 0) return 0;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return fail("adam: arenas must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  adam_kernel<<<ew_grid(n / 4 + 1), 256, 0, s>>>(p, g, m, v, n, b1, b2, eps, wd, (float)((double)lr / bc1),
                                                 (float)(1.0 / sqrt(bc2)), gscale);
  RD_LAUNCHED();
  return 0;
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, float lr, float wd, float gscale) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    p[i] = p[i] - lr * fmaf(wd, p[i], g[i] * gscale);
}

int launch_sgd(float* p, const float* g, long long n, float lr, float wd, float gscale, cudaStream_t s) {
  if (n <= 0) return 0;
  sgd_kernel<<<ew_grid(n), 256, 0, s>>>(p, g, n, lr, wd, gscale);
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Linear blending (lib/evaluation.py:484-567, denormalize_numpy lib/data_normalization.py:41-53)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double ramp_weight(int c, int lo_edge, int hi_edge, int T, int overlap, double step) {
  // lo_edge = ulx (or uly), hi_edge = lrx (or lry); follows _get_blend_weights for one axis
  double w = 1.0;
  if (lo_edge > 0) {
    if (c < lo_edge - overlap) return 0.0;
    if (c < lo_edge) {
      const int k = c - (lo_edge - overlap);
      w *= (k == overlap - 1) ? 1.0 : (double)k * step;
    }
  }
  if (hi_edge < T - 1 && c > hi_edge) {
    const int k = overlap - 1 - (c - (hi_edge + 1));
    if (k >= 0 && k < overlap) w *= (k == overlap - 1) ? 1.0 : (double)k * step;
  }
  return w;
}

__global__ void __launch_bounds__(256)
blend_kernel(const float* __restrict__ tiles, const float* __restrict__ mean, const float* __restrict__ std,
             const int32_t* __restrict__ geom, int T, int stride, double* __restrict__ raster, int rows, int cols) {
  const int t = blockIdx.y;
  const int y0 = geom[t * 6 + 0], x0 = geom[t * 6 + 1];
  const int uly = geom[t * 6 + 2], ulx = geom[t * 6 + 3], lry = geom[t * 6 + 4], lrx = geom[t * 6 + 5];
  const int overlap = T - stride;
  const double step = overlap > 1 ? 1.0 / (double)(overlap - 1) : 0.0;
  const float sd = std[t], mu = mean[t];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < T * T; i += gridDim.x * 256) {
    const int r = i / T, c = i % T;
    const int gy = y0 + r, gx = x0 + c;
    if (gy < 0 || gy >= rows || gx < 0 || gx >= cols) continue;
    const double w = ramp_weight(c, ulx, lrx, T, overlap, step) * ramp_weight(r, uly, lry, T, overlap, step);
    const float den = __fadd_rn(__fmul_rn(tiles[(size_t)t * T * T + i], sd), mu);
    atomicAdd(raster + (size_t)gy * cols + gx, (double)den * w);
  }
}

int launch_blend(const float* tiles, const float* mean, const float* std, const int32_t* geom, int n, int T,
                 int stride, double* raster, int rows, int cols, cudaStream_t s) {
  if (n <= 0) return 0;
  if (stride <= 0 || stride > T) return fail("blend: bad stride %d for tile %d", stride, T);
  dim3 grid(cdiv((long long)T * T, 256 * 4), n);
  blend_kernel<<<grid, 256, 0, s>>>(tiles, mean, std, geom, T, stride, raster, rows, cols);
  RD_LAUNCHED();
  return 0;
}

__global__ void fill_kernel(float* p, float v, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) p[i] = v;
}
int launch_fill(float* p, float v, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  fill_kernel<<<ew_grid(n), 256, 0, s>>>(p, v, n);
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Weight packing (every forward in training: the optimizer rewrites the weights each step)
//   conv3x3 OIHW [Co,Ci,3,3]:  kn[(t,ci)][co]  = W[co,ci,t]         (forward, K rows x N)
//                              nk[co][(t,ci)]  = W[co,ci,t]         (forward, N rows x K: tcgen05 K-major B)
//                              dkn[(t,co)][ci] = W[co,ci,8-t]       (dgrad: taps rotated by 180 degrees)
//                              dnk[ci][(t,co)] = W[co,ci,8-t]
//   convT [Ci,Co,2,2]:         kn[ci][(ab,co)] = W[ci,co,ab]        (forward)     nk[(ab,co)][ci] (its transpose)
//     the transpose is also the K x N matrix of the dgrad GEMM (K = (ab,co), N = ci), and kn its N x K form.
// ----------------------------------------------------------------------------------------------
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, float* __restrict__ kn, float* __restrict__ nk,
                                    float* __restrict__ dkn, float* __restrict__ dnk, __nv_bfloat16* __restrict__ dnk_b,
                                    int Co, int Ci, int rnd) {
  const long long total = (long long)Co * Ci * 9;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int t = (int)(i % 9);
    const int ci = (int)((i / 9) % Ci);
    const int co = (int)(i / (9LL * Ci));
    float v = w[i];
    if (rnd) v = tf32_rn(v);
    if (kn) kn[((size_t)t * Ci + ci) * Co + co] = v;
    if (nk) nk[(size_t)co * 9 * Ci + (size_t)t * Ci + ci] = v;
    const int tr = 8 - t;
    if (dkn) dkn[((size_t)tr * Co + co) * Ci + ci] = v;
    if (dnk) dnk[(size_t)ci * 9 * Co + (size_t)tr * Co + co] = v;
    if (dnk_b) dnk_b[(size_t)ci * 9 * Co + (size_t)tr * Co + co] = __float2bfloat16_rn(w[i]);
  }
}
// Tiled variant (Co multiple of 32, Ci multiple of 8): one block stages a 32 co x 8 ci x 9 tap tile in shared memory
// (coalesced 288-byte pieces of the OIHW rows) and writes every requested layout with the layout's own fastest index
// across the lanes (full 32-byte sectors or better) -- the element-wise kernel above scatters 4-byte (2-byte) stores
// over 32 sectors per warp.  Small tiles on purpose: the layers are small and the grid must fill the chip.
constexpr int PK_CI = 8, PK_ROW = PK_CI * 9;            // 72 floats per co row of the tile
__global__ void __launch_bounds__(256)
pack_conv3x3_tiled_kernel(const float* __restrict__ w, float* __restrict__ kn, float* __restrict__ nk,
                          float* __restrict__ dkn, float* __restrict__ dnk, __nv_bfloat16* __restrict__ dnk_b, int Co,
                          int Ci, int rnd) {
  __shared__ float tl[32][PK_ROW + 1];
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * PK_CI;
  for (int idx = threadIdx.x; idx < 32 * PK_ROW; idx += 256) {
    const int r = idx / PK_ROW, c = idx - r * PK_ROW;
    tl[r][c] = w[((size_t)(co0 + r) * Ci + ci0) * 9 + c];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * PK_ROW; idx += 256) {
    {                                                                 // ci fastest: (ci, t, co)
      const int ci = idx % PK_CI, t = (idx / PK_CI) % 9, co = idx / PK_ROW;
      if (nk) {                                                       // nk[co][(t,ci)]
        const float v = tl[co][ci * 9 + t];
        nk[(size_t)(co0 + co) * 9 * Ci + (size_t)t * Ci + ci0 + ci] = rnd ? tf32_rn(v) : v;
      }
      if (dkn) {                                                      // dkn[(tr,co)][ci]
        const float v = tl[co][ci * 9 + (8 - t)];
        dkn[((size_t)t * Co + co0 + co) * Ci + ci0 + ci] = rnd ? tf32_rn(v) : v;
      }
    }
    {                                                                 // co fastest: (co, t, ci)
      const int co = idx & 31, t = (idx >> 5) % 9, ci = idx / 288;
      if (kn) {                                                       // kn[(t,ci)][co]
        const float v = tl[co][ci * 9 + t];
        kn[((size_t)t * Ci + ci0 + ci) * Co + co0 + co] = rnd ? tf32_rn(v) : v;
      }
      if (dnk || dnk_b) {                                             // dnk[ci][(tr,co)]
        const float v = tl[co][ci * 9 + (8 - t)];
        const size_t o = (size_t)(ci0 + ci) * 9 * Co + (size_t)t * Co + co0 + co;
        if (dnk) dnk[o] = rnd ? tf32_rn(v) : v;
        if (dnk_b) dnk_b[o] = __float2bfloat16_rn(v);
      }
    }
  }
}
int launch_pack_conv3x3(const float* w, float* kn, float* nk, float* dkn, float* dnk, void* dnk_b, int Co, int Ci,
                        int rnd, cudaStream_t s) {
  if (Co % 32 == 0 && Ci % PK_CI == 0) {
    pack_conv3x3_tiled_kernel<<<dim3(Ci / PK_CI, Co / 32), 256, 0, s>>>(w, kn, nk, dkn, dnk,
                                                                   reinterpret_cast<__nv_bfloat16*>(dnk_b), Co, Ci, rnd);
    RD_LAUNCHED();
    return 0;
  }
  pack_conv3x3_kernel<<<ew_grid((long long)Co * Ci * 9), 256, 0, s>>>(w, kn, nk, dkn, dnk,
                                                                      reinterpret_cast<__nv_bfloat16*>(dnk_b), Co, Ci, rnd);
  RD_LAUNCHED();
  return 0;
}

__global__ void pack_convt_kernel(const float* __restrict__ w, float* __restrict__ kn, float* __restrict__ nk,
                                  __nv_bfloat16* __restrict__ kn_b, int Ci, int Co, int rnd) {
  const long long total = (long long)Ci * Co * 4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ab = (int)(i % 4);
    const int co = (int)((i / 4) % Co);
    const int ci = (int)(i / (4LL * Co));
    float v = w[i];
    if (rnd) v = tf32_rn(v);
    if (kn) kn[(size_t)ci * 4 * Co + (size_t)ab * Co + co] = v;
    if (kn_b) kn_b[(size_t)ci * 4 * Co + (size_t)ab * Co + co] = __float2bfloat16_rn(w[i]);
    if (nk) nk[((size_t)ab * Co + co) * Ci + ci] = v;
  }
}
int launch_pack_convt(const float* w, float* kn, float* nk, void* kn_b, int Ci, int Co, int rnd, cudaStream_t s) {
  pack_convt_kernel<<<ew_grid((long long)Ci * Co * 4), 256, 0, s>>>(w, kn, nk, reinterpret_cast<__nv_bfloat16*>(kn_b), Ci,
                                                                    Co, rnd);
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Batched packing: every layer's weights in ONE launch (the per-layer launches above cost 0.18 ms per step in 14
// launches for ~150 MB of traffic).  A block looks its job up in the by-value job table and runs the per-layer
// routine on its share: a (32 co x 8 ci) tile of a conv3x3, or 2048 consecutive elements of the element-wise kinds.
// ----------------------------------------------------------------------------------------------
static constexpr int PK_CHUNK = 2048;

__global__ void __launch_bounds__(256) pack_batched_kernel(const __grid_constant__ PackJobs J) {
  __shared__ float tl[32][PK_ROW + 1];
  int j = 0;
  while (j + 1 < J.n && (int)blockIdx.x >= J.job[j + 1].block0) ++j;
  const PackJob& job = J.job[j];
  const int lb = (int)blockIdx.x - job.block0;
  const float* __restrict__ w = job.w;
  const int Co = job.Co, Ci = job.Ci, rnd = job.rnd;
  if (job.kind == PACK_CONV3X3_TILED) {
    float *kn = job.o0, *nk = job.o1, *dkn = job.o2, *dnk = job.o3;
    __nv_bfloat16* dnk_b = reinterpret_cast<__nv_bfloat16*>(job.ob);
    const int nbx = Ci / PK_CI;
    const int co0 = (lb / nbx) * 32, ci0 = (lb % nbx) * PK_CI;
    for (int idx = threadIdx.x; idx < 32 * PK_ROW; idx += 256) {
      const int r = idx / PK_ROW, c = idx - r * PK_ROW;
      tl[r][c] = w[((size_t)(co0 + r) * Ci + ci0) * 9 + c];
    }
    __syncthreads();
    // vector stores for the tensor-core layouts: one (co, t) pair = 8 consecutive ci of nk (2 x 16 bytes), one
    // (ci, t, 8 co) group = 16 bytes of dnk_b / 2 x 16 bytes of dnk -- 288 items each, instead of 2304 scalar stores
    for (int item = threadIdx.x; item < 288; item += 256) {
      if (nk) {
        const int co = item / 9, t = item - co * 9;
        float v[8];
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) { const float x = tl[co][ci * 9 + t]; v[ci] = rnd ? tf32_rn(x) : x; }
        float4* dst = reinterpret_cast<float4*>(nk + (size_t)(co0 + co) * 9 * Ci + (size_t)t * Ci + ci0);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (dnk || dnk_b) {
        const int g = item & 3, t = (item >> 2) % 9, ci = item / 36;          // 8 co per item, co fastest
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = tl[g * 8 + q][ci * 9 + (8 - t)];
        const size_t o = (size_t)(ci0 + ci) * 9 * Co + (size_t)t * Co + co0 + g * 8;
        if (dnk) {
          float4* dst = reinterpret_cast<float4*>(dnk + o);
          dst[0] = rnd ? tf32_rn4(make_float4(v[0], v[1], v[2], v[3])) : make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = rnd ? tf32_rn4(make_float4(v[4], v[5], v[6], v[7])) : make_float4(v[4], v[5], v[6], v[7]);
        }
        if (dnk_b) {
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 bb = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
            pk[q] = *reinterpret_cast<uint32_t*>(&bb);
          }
          *reinterpret_cast<uint4*>(dnk_b + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    nk = nullptr; dnk = nullptr; dnk_b = nullptr;            // done above; the loop below serves the CUDA-core layouts
    if (!kn && !dkn) return;
    for (int idx = threadIdx.x; idx < 32 * PK_ROW; idx += 256) {
      {                                                                 // ci fastest: (ci, t, co)
        const int ci = idx % PK_CI, t = (idx / PK_CI) % 9, co = idx / PK_ROW;
        if (nk) {
          const float v = tl[co][ci * 9 + t];
          nk[(size_t)(co0 + co) * 9 * Ci + (size_t)t * Ci + ci0 + ci] = rnd ? tf32_rn(v) : v;
        }
        if (dkn) {
          const float v = tl[co][ci * 9 + (8 - t)];
          dkn[((size_t)t * Co + co0 + co) * Ci + ci0 + ci] = rnd ? tf32_rn(v) : v;
        }
      }
      {                                                                 // co fastest: (co, t, ci)
        const int co = idx & 31, t = (idx >> 5) % 9, ci = idx / 288;
        if (kn) {
          const float v = tl[co][ci * 9 + t];
          kn[((size_t)t * Ci + ci0 + ci) * Co + co0 + co] = rnd ? tf32_rn(v) : v;
        }
        if (dnk || dnk_b) {
          const float v = tl[co][ci * 9 + (8 - t)];
          const size_t o = (size_t)(ci0 + ci) * 9 * Co + (size_t)t * Co + co0 + co;
          if (dnk) dnk[o] = rnd ? tf32_rn(v) : v;
          if (dnk_b) dnk_b[o] = __float2bfloat16_rn(v);
        }
      }
    }
    return;
  }
  const long long total = job.kind == PACK_CONV3X3 ? (long long)Co * Ci * 9
                          : job.kind == PACK_CONVT ? (long long)Ci * Co * 4 : (long long)Co * Ci;
  const long long lo = (long long)lb * PK_CHUNK, hi = lo + PK_CHUNK < total ? lo + PK_CHUNK : total;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) {
    const float raw = w[i];
    const float v = rnd ? tf32_rn(raw) : raw;
    if (job.kind == PACK_CONV3X3) {                 // o0 = kn, o1 = nk, o2 = dkn, o3 = dnk, ob = dnk_b
      const int t = (int)(i % 9), ci = (int)((i / 9) % Ci), co = (int)(i / (9LL * Ci)), tr = 8 - t;
      if (job.o0) job.o0[((size_t)t * Ci + ci) * Co + co] = v;
      if (job.o1) job.o1[(size_t)co * 9 * Ci + (size_t)t * Ci + ci] = v;
      if (job.o2) job.o2[((size_t)tr * Co + co) * Ci + ci] = v;
      if (job.o3) job.o3[(size_t)ci * 9 * Co + (size_t)tr * Co + co] = v;
      if (job.ob) reinterpret_cast<__nv_bfloat16*>(job.ob)[(size_t)ci * 9 * Co + (size_t)tr * Co + co] = __float2bfloat16_rn(raw);
    } else if (job.kind == PACK_CONVT) {            // w [Ci][Co][2][2]: o0 = kn, o1 = nk, ob = kn_b
      const int ab = (int)(i % 4), co = (int)((i / 4) % Co), ci = (int)(i / (4LL * Co));
      if (job.o0) job.o0[(size_t)ci * 4 * Co + (size_t)ab * Co + co] = v;
      if (job.ob) reinterpret_cast<__nv_bfloat16*>(job.ob)[(size_t)ci * 4 * Co + (size_t)ab * Co + co] = __float2bfloat16_rn(raw);
      if (job.o1) job.o1[((size_t)ab * Co + co) * Ci + ci] = v;
    } else {                                        // conv1x1 [Co][Ci]: o0 = copy, o1 = transpose
      const int ci = (int)(i % Ci), co = (int)(i / Ci);
      job.o0[i] = v;
      job.o1[(size_t)ci * Co + co] = v;
    }
  }
}

int pack_jobs_add(PackJobs& J, int kind, const float* w, float* o0, float* o1, float* o2, float* o3, void* ob, int Co,
                  int Ci, int rnd) {
  if (J.n >= PACK_MAX_JOBS) return fail("pack: more than %d layers in one batch", PACK_MAX_JOBS);
  if (kind == PACK_CONV3X3 && Co % 32 == 0 && Ci % PK_CI == 0) kind = PACK_CONV3X3_TILED;
  PackJob& j = J.job[J.n++];
  j.kind = kind; j.w = w; j.o0 = o0; j.o1 = o1; j.o2 = o2; j.o3 = o3; j.ob = ob; j.Co = Co; j.Ci = Ci; j.rnd = rnd;
  j.block0 = J.total_blocks;
  long long blocks;
  if (kind == PACK_CONV3X3_TILED) blocks = (long long)(Ci / PK_CI) * (Co / 32);
  else {
    const long long total = kind == PACK_CONV3X3 ? (long long)Co * Ci * 9 : kind == PACK_CONVT ? (long long)Ci * Co * 4
                                                                                               : (long long)Co * Ci;
    blocks = (total + PK_CHUNK - 1) / PK_CHUNK;
  }
  J.total_blocks += (int)blocks;
  return 0;
}

int launch_pack_batched(const PackJobs& J, cudaStream_t s) {
  if (J.n == 0) return 0;
  pack_batched_kernel<<<J.total_blocks, 256, 0, s>>>(J);
  RD_LAUNCHED();
  return 0;
}

// gradient un-packing: part [S][(t,ci)][co] summed over S -> dW OIHW
__global__ void unpack_conv3x3_grad_kernel(const float* __restrict__ part, int S, float* __restrict__ dw, int Co,
                                           int Ci, int ntaps) {
  const long long total = (long long)Co * Ci * ntaps;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    // i indexes the packed layout (coalesced reads); scatter to OIHW
    const int co = (int)(i % Co);
    const int ci = (int)((i / Co) % Ci);
    const int t = (int)(i / ((long long)Co * Ci));
    float a = 0.f;
    for (int sp = 0; sp < S; ++sp) a += part[(size_t)sp * total + i];
    dw[((size_t)co * Ci + ci) * ntaps + t] = a;
  }
}
// (A shared-memory tiled un-pack, the mirror image of pack_conv3x3_tiled_kernel, was measured slower -- 0.31 vs 0.26 ms
// per step: the split partials are read S times, so the element-wise kernels' full-chip parallelism on the reads matters
// more than their scattered 4-byte stores.)
// Tiled form for the large layers (few splits): a block sums the splits of a (9 taps x 8 ci) x 32 co tile with coalesced
// 128-byte row reads into shared memory and writes 288 contiguous bytes per output channel -- the element-wise kernel
// above scatters its 4-byte stores over 32 sectors per warp (0.7 TB/s on the 512 x 512 layers, ncu round 2).
__global__ void __launch_bounds__(256)
unpack_conv3x3_grad_tiled_kernel(const float* __restrict__ part, int S, float* __restrict__ dw, int Co, int Ci) {
  __shared__ float tl[32][PK_ROW + 1];                    // [co][ci_local * 9 + t]
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * PK_CI;
  const size_t total = (size_t)Co * Ci * 9;
  for (int idx = threadIdx.x; idx < 32 * PK_ROW; idx += 256) {
    const int co = idx & 31, r = idx >> 5;                // r = t * PK_CI + ci_local: consecutive rows of the packed layout
    const int t = r / PK_CI, ci = r - t * PK_CI;
    const float* src = part + ((size_t)t * Ci + ci0 + ci) * Co + co0 + co;
    float a = 0.f;
    for (int sp = 0; sp < S; ++sp) a += __ldg(src + (size_t)sp * total);
    tl[co][ci * 9 + t] = a;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * PK_ROW; idx += 256) {
    const int co = idx / PK_ROW, c = idx - co * PK_ROW;
    dw[((size_t)(co0 + co) * Ci + ci0) * 9 + c] = tl[co][c];
  }
}
int launch_unpack_conv_grad(const float* part, int S, float* dw, int Co, int Ci, int ntaps, cudaStream_t s) {
  static const bool no_tiled = getenv("RESDEPTH_UNPACK_SIMPLE") != nullptr;
  if (!no_tiled && ntaps == 9 && Co % 32 == 0 && Ci % PK_CI == 0 && S <= 8 && (long long)Co * Ci >= 128 * 128) {
    unpack_conv3x3_grad_tiled_kernel<<<dim3(Ci / PK_CI, Co / 32), 256, 0, s>>>(part, S, dw, Co, Ci);
    RD_LAUNCHED();
    return 0;
  }
  unpack_conv3x3_grad_kernel<<<ew_grid((long long)Co * Ci * ntaps), 256, 0, s>>>(part, S, dw, Co, Ci, ntaps);
  RD_LAUNCHED();
  return 0;
}

// wide weight-gradient layouts (tc_make_reduce_plan_wide), summed over S -> dW OIHW:
//   by_ci != 0: part [S][ci][(t, co)] (row stride ldn)     by_ci == 0: part [S][co][(t, ci)]
__global__ void unpack_conv3x3_grad_wide_kernel(const float* __restrict__ part, int S, int ldn, float* __restrict__ dw,
                                                int Co, int Ci, int by_ci) {
  const int Cu = by_ci ? Ci : Co, Cs = by_ci ? Co : Ci;           // row channels, column channels per tap
  const long long total = (long long)Cu * 9 * Cs;
  const size_t split = (size_t)Cu * ldn;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int cs = (int)(i % Cs);
    const int t = (int)((i / Cs) % 9);
    const int cu = (int)(i / (9LL * Cs));
    const size_t src = (size_t)cu * ldn + (size_t)t * Cs + cs;
    float a = 0.f;
    for (int sp = 0; sp < S; ++sp) a += part[(size_t)sp * split + src];
    const int co = by_ci ? cs : cu, ci = by_ci ? cu : cs;
    dw[((size_t)co * Ci + ci) * 9 + t] = a;
  }
}
int launch_unpack_conv_grad_wide(const float* part, int S, int ldn, float* dw, int Co, int Ci, int by_ci, cudaStream_t s) {
  unpack_conv3x3_grad_wide_kernel<<<ew_grid((long long)Co * Ci * 9), 256, 0, s>>>(part, S, ldn, dw, Co, Ci, by_ci);
  RD_LAUNCHED();
  return 0;
}

// part [S][(ab,co)][ci] summed over S -> dW [ci][co][2][2]
__global__ void unpack_convt_grad_kernel(const float* __restrict__ part, int S, float* __restrict__ dw, int Ci,
                                         int Co) {
  const long long total = (long long)Ci * Co * 4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int ci = (int)(i % Ci);
    const int co = (int)((i / Ci) % Co);
    const int ab = (int)(i / ((long long)Ci * Co));
    float a = 0.f;
    for (int sp = 0; sp < S; ++sp) a += part[(size_t)sp * total + i];
    dw[((size_t)ci * Co + co) * 4 + ab] = a;
  }
}
int launch_unpack_convt_grad(const float* part, int S, float* dw, int Ci, int Co, cudaStream_t s) {
  unpack_convt_grad_kernel<<<ew_grid((long long)Ci * Co * 4), 256, 0, s>>>(part, S, dw, Ci, Co);
  RD_LAUNCHED();
  return 0;
}

// First-layer weight gradient on tensor cores: the NCHW network input is expanded to an NHWC "im2col" tensor
// xcol[b,h,w,k] with k = ci*9 + r*3 + s  (value x[b,ci,h+r-1,w+s-1], zero outside the image and for k >= Cin*9,
// Kc = Cin*9 rounded up to 32) so that dW[co][k] = sum_p xcol[p][k] * dz[p][co] is a plain reduce GEMM.
// one thread per pixel: reads its 3x3 neighbourhood of every input channel (neighbours hit in L1) and writes the
// Kc-float row of xcol as full 128-byte lines
template <int KC>
__global__ void __launch_bounds__(256)
im2col_first_kernel(const float* __restrict__ x, float* __restrict__ xcol, int B, int Cin, int H, int W, int rnd) {
  const long long npix = (long long)B * H * W;
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < npix; p += gridDim.x * 256LL) {
    const int w = (int)(p % W);
    const int h = (int)((p / W) % H);
    const int b = (int)(p / ((long long)W * H));
    float v[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) v[k] = 0.f;
    const float* xb = x + (size_t)b * Cin * H * W;
#pragma unroll
    for (int ci = 0; ci < KC / 9; ++ci) {
      if (ci < Cin) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int hh = h + r - 1;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int ww = w + q - 1;
            float val = 0.f;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(xb + ((size_t)ci * H + hh) * W + ww);
            v[ci * 9 + r * 3 + q] = rnd ? tf32_rn(val) : val;
          }
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(xcol + (size_t)p * KC);
#pragma unroll
    for (int j = 0; j < KC / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}
int launch_im2col_first(const float* x, float* xcol, int B, int Cin, int H, int W, int Kc, int rnd, cudaStream_t s) {
  const int grid = ew_grid((long long)B * H * W);
  if (Kc == 32) im2col_first_kernel<32><<<grid, 256, 0, s>>>(x, xcol, B, Cin, H, W, rnd);
  else if (Kc == 64) im2col_first_kernel<64><<<grid, 256, 0, s>>>(x, xcol, B, Cin, H, W, rnd);
  else if (Kc == 96) im2col_first_kernel<96><<<grid, 256, 0, s>>>(x, xcol, B, Cin, H, W, rnd);
  else return fail("im2col_first: unsupported Kc=%d", Kc);
  RD_LAUNCHED();
  return 0;
}
// bf16 variant: xcol is a bf16 tensor [B*H*W][Kc] with Kc = 64 (Cin <= 7) or 128.  One thread per (pixel, 16-byte
// chunk of 8 k-values): the warp's stores cover whole 32-byte sectors of consecutive rows, and only the
// ceil(Cin*9/8) chunks that hold data are written -- the padding of every row is zeroed once by rd_reserve
// (launch_im2col_first_bf16_clear) and never touched again.
template <int KC>
__global__ void __launch_bounds__(256)
im2col_first_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xcol, int B, int Cin, int H, int W,
                         int chunks) {
  const int K = Cin * 9;
  const int per_row = W * chunks;
  for (int row = blockIdx.x; row < B * H; row += gridDim.x) {          // one image row per iteration: 32-bit math
    const int b = row / H, h = row - b * H;
    const float* xb = x + (size_t)b * Cin * H * W;
    __nv_bfloat16* orow = xcol + (size_t)row * W * KC;
    for (int t = threadIdx.x; t < per_row; t += 256) {
      const int w = t / chunks, c = t - w * chunks;
      uint32_t pk[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = c * 8 + jj * 2 + e;
          const int ci = k / 9, rs = k - ci * 9;
          const int r = rs / 3, q = rs - r * 3;
          const int hh = h + r - 1, ww = w + q - 1;
          const bool ok = k < K && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
          // column K is the constant 1: the Gram GEMM xcol^T xcol then also yields the column sums of xcol (row K) and
          // the pixel count; the weight-gradient GEMM ignores that row
          v[e] = ok ? __ldg(xb + (ci * H + hh) * W + ww) : (k == K ? 1.f : 0.f);
        }
        __nv_bfloat162 bb = __floats2bfloat162_rn(v[0], v[1]);
        pk[jj] = *reinterpret_cast<uint32_t*>(&bb);
      }
      *reinterpret_cast<uint4*>(orow + (size_t)w * KC + c * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}
// Compile-time form for Cin <= 3 (the tensor-core first layer): one thread per pixel gathers its 3 x 3 x Cin
// neighbourhood with unrolled, constant tap offsets (three row and three column validity flags instead of a divide /
// modulo chain per element -- the generic kernel above needs ~200 instructions per 16-byte chunk and was SM-bound at
// 19 % DRAM, ncu round 2), appends the constant-one column and writes the live 16-byte chunks of its row.
template <int CIN, int KC>
__global__ void __launch_bounds__(256)
im2col_first_bf16_px_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xcol, int B, int H, int W) {
  constexpr int K = CIN * 9, CHUNKS = (K + 1 + 7) / 8;
  const long long NP = (long long)B * H * W;
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < NP; p += gridDim.x * 256LL) {
    const int w = (int)(p % W);
    const long long r_ = p / W;
    const int h = (int)(r_ % H);
    const long long b = r_ / H;
    const bool rok[3] = {h > 0, true, h + 1 < H}, cok[3] = {w > 0, true, w + 1 < W};
    float v[CHUNKS * 8];
#pragma unroll
    for (int k = 0; k < CHUNKS * 8; ++k) v[k] = k == K ? 1.f : 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* xc = x + ((size_t)(b * CIN + ci) * H + h) * W + w;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (rok[r] && cok[q]) v[ci * 9 + r * 3 + q] = __ldg(xc + (r - 1) * W + (q - 1));
    }
    uint4* dst = reinterpret_cast<uint4*>(xcol + (size_t)p * KC);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 bb = __floats2bfloat162_rn(v[c * 8 + 2 * j], v[c * 8 + 2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&bb);
      }
      dst[c] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

int launch_im2col_first_bf16(const float* x, void* xcol, int B, int Cin, int H, int W, int Kc, cudaStream_t s) {
  static const bool generic = getenv("RESDEPTH_IM2COL_GENERIC") != nullptr;
  if ((!generic || Kc == 32) && (Kc == 64 || Kc == 32) && Cin >= 1 && Cin <= 3) {
    // Kc = 32: rows of 32 bf16 (64 bytes) -- the reduce GEMMs read them through tensor maps whose boxes are 64 wide
    const long long NP = (long long)B * H * W;
    const int grid = (int)((NP + 255) / 256 < 148 * 8 ? (NP + 255) / 256 : 148 * 8);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(xcol);
#define RD_IM2COL(CI)                                                                     \
  {                                                                                       \
    if (Kc == 64) im2col_first_bf16_px_kernel<CI, 64><<<grid, 256, 0, s>>>(x, o, B, H, W); \
    else im2col_first_bf16_px_kernel<CI, 32><<<grid, 256, 0, s>>>(x, o, B, H, W);          \
  }
    if (Cin == 3) RD_IM2COL(3) else if (Cin == 2) RD_IM2COL(2) else RD_IM2COL(1)
#undef RD_IM2COL
    RD_LAUNCHED();
    return 0;
  }
  const int chunks = (Cin * 9 + 7) / 8;
  if (chunks * 8 > Kc) return fail("im2col_first_bf16: Kc=%d too small for Cin=%d", Kc, Cin);
  const int grid = B * H < 148 * 8 ? B * H : 148 * 8;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(xcol);
  if (Kc == 64) im2col_first_bf16_kernel<64><<<grid, 256, 0, s>>>(x, out, B, Cin, H, W, chunks);
  else if (Kc == 128) im2col_first_bf16_kernel<128><<<grid, 256, 0, s>>>(x, out, B, Cin, H, W, chunks);
  else return fail("im2col_first_bf16: unsupported Kc=%d", Kc);
  RD_LAUNCHED();
  return 0;
}
// zero the whole expansion once (rows are only partially rewritten by the kernel above)
int launch_im2col_first_bf16_clear(void* xcol, size_t bytes, cudaStream_t s) {
  RD_CUDA(cudaMemsetAsync(xcol, 0, bytes, s));
  return 0;
}
// part [S][Kc][Co] summed over S -> dW [Co][K] (K = Cin*9 <= Kc; OIHW flattening of the first conv)
// 32 consecutive outputs per block (coalesced 128-byte reads of every split), the splits spread over the 8 warps and
// combined in a fixed order: the one-wave reduce GEMM of this layer produces ~148 splits of only Co*K values each
__global__ void __launch_bounds__(256)
unpack_first_grad_kernel(const float* __restrict__ part, int S, float* __restrict__ dw, int Co, int K, int Kc, int pitch) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane, total = Co * K;
  const int co = i % Co, k = i / Co;
  float a = 0.f;
  if (i < total)
    for (int sp = w; sp < S; sp += 8) a += part[((size_t)sp * Kc + k) * pitch + co];
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && i < total) {
#pragma unroll
    for (int j = 1; j < 8; ++j) a += red[j][lane];
    dw[(size_t)co * K + k] = a;
  }
}
int launch_unpack_first_grad(const float* part, int S, float* dw, int Co, int K, int Kc, cudaStream_t s, int pitch) {
  unpack_first_grad_kernel<<<cdiv(Co * K, 32), 256, 0, s>>>(part, S, dw, Co, K, Kc, pitch ? pitch : Co);
  RD_LAUNCHED();
  return 0;
}

// First encoder block, weight gradient without the apply pass.  With X = im2col(x) [pixels][K], A = X^T gY (what the
// reduce GEMM + unpack_first_grad have just written into dw as dw[co][k]) and dz = cs (gY - c1 - (z - mean) c2):
//   dW[co][k] = sum_p X[p][k] dz[p][co] = cs[co] ( A[k][co] - c1[co] cx[k] - c2[co] ( (G W[co])[k] - cx[k] mean[co] ) )
// because z = X W^T (bias-free conv under BatchNorm): sum_p X[p][k] z[p][co] = sum_k' G[k][k'] W[co][k'] with the Gram
// matrix G = X^T X.  gram [Kc][Kc] comes from the same reduce GEMM run on (xcol, xcol); its column K holds cx (the
// im2col kernel writes a constant-one column there).  One thread per (co, k); K <= 27.
__global__ void __launch_bounds__(256)
first_grad_correct_kernel(float* __restrict__ dw, const float* __restrict__ w, const float* __restrict__ gram,
                          const BwdCoef* __restrict__ coef, int Co, int K, int Kc) {   // Kc: row pitch of gram
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Co * K) return;
  const int co = i / K, k = i - co * K;
  const BwdCoef cf = coef[co];
  const float cx = gram[(size_t)k * Kc + K];
  double gw = 0.0;
  for (int k2 = 0; k2 < K; ++k2) gw += (double)gram[(size_t)k * Kc + k2] * (double)w[(size_t)co * K + k2];
  const double corr = (double)cf.c1 * cx + (double)cf.c2 * (gw - (double)cx * (double)cf.mean);
  dw[i] = (float)((double)cf.cs * ((double)dw[i] - corr));
}
int launch_first_grad_correct(float* dw, const float* w, const float* gram, const void* coef, int Co, int K, int Kc,
                              cudaStream_t s) {
  first_grad_correct_kernel<<<cdiv(Co * K, 256), 256, 0, s>>>(dw, w, gram, reinterpret_cast<const BwdCoef*>(coef), Co, K, Kc);
  RD_LAUNCHED();
  return 0;
}

// out[c] = sum over pixels of g[p][c]
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ g, long long npix, int C, float* __restrict__ part) {
  extern __shared__ float red[];                        // [PL][C]
  const int Q = C >> 2, PL = 256 / Q;
  const int q = threadIdx.x % Q, pl = threadIdx.x / Q;
  float4 a = make_float4(0, 0, 0, 0);
  if (pl < PL) {
    for (long long p = (long long)blockIdx.x * PL + pl; p < npix; p += (long long)gridDim.x * PL) {
      const float4 v = ld4(g + (size_t)p * C + q * 4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    st4(red + (size_t)pl * C + q * 4, a);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    float acc = 0.f;
    for (int p = 0; p < PL; ++p) acc += red[(size_t)p * C + i];
    part[(size_t)blockIdx.x * C + i] = acc;
  }
}
// out[i] = sum_p part[p][i] (double accumulation): 32 columns x 32 partial lanes per block
__global__ void __launch_bounds__(1024)
sum_partials_kernel(const float* __restrict__ part, int nparts, int n, int row_stride, int col_stride,
                    float* __restrict__ out) {
  __shared__ double red[32][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  double a = 0.0;
  if (i < n) {
    const float* col = part + (size_t)i * col_stride;
    int p = pl;
    for (; p + 96 < nparts; p += 128) {                  // four rows in flight per thread
      const float v0 = __ldg(col + (size_t)p * row_stride), v1 = __ldg(col + (size_t)(p + 32) * row_stride);
      const float v2 = __ldg(col + (size_t)(p + 64) * row_stride), v3 = __ldg(col + (size_t)(p + 96) * row_stride);
      a += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; p < nparts; p += 32) a += (double)__ldg(col + (size_t)p * row_stride);
  }
  red[pl][cl] = a;
  __syncthreads();
  if (pl == 0 && i < n) {
#pragma unroll
    for (int k = 1; k < 32; ++k) a += red[k][cl];
    out[i] = (float)a;
  }
}
int launch_sum_partials(const float* part, int nparts, int n, int row_stride, int col_stride, float* out,
                        cudaStream_t s) {
  sum_partials_kernel<<<cdiv(n, 32), 1024, 0, s>>>(part, nparts, n, row_stride, col_stride, out);
  RD_LAUNCHED();
  return 0;
}
int launch_channel_sum(const float* g, long long npix, int C, float* out, float* scratch, size_t scratch_floats,
                       cudaStream_t s) {
  const int Q = C / 4;
  if (C % 4 || Q > 256) return fail("channel_sum: unsupported C=%d", C);
  const int PL = 256 / Q;
  int grid = bwd_grid(npix, PL);
  while ((size_t)grid * C > scratch_floats && grid > 1) grid /= 2;
  if ((size_t)grid * C > scratch_floats) return fail("channel_sum: scratch too small");
  channel_sum_kernel<<<grid, 256, (size_t)PL * C * sizeof(float), s>>>(g, npix, C, scratch);
  RD_LAUNCHED();
  return launch_sum_partials(scratch, grid, C, C, 1, out, s);
}

// ----------------------------------------------------------------------------------------------
// outer_skip_BN: BatchNorm2d(1) on channel 0 of the NCHW input (lib/UNet.py:192-193,231-237)
//   stats : partials [nblk][1][2] = (sum x0, sum x0^2)          -> bn_finalize with C = 1
//   bwd   : partials [nblk][2]    = (sum dy, sum dy*(x0-mean))  -> dgamma = invstd * S2, dbeta = S1
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
outer_bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                       float* __restrict__ partials, int B, int Cin, int HW) {
  __shared__ float r1[256], r2[256];
  const float mu = dy ? mean[0] : 0.f;
  float a = 0.f, b = 0.f;
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long bi = i / HW, p = i - bi * HW;
    const float v = x[(size_t)bi * Cin * HW + p];
    if (dy) {
      const float g = dy[i];
      a += g;
      b = fmaf(g, v - mu, b);
    } else {
      a += v;
      b = fmaf(v, v, b);
    }
  }
  r1[threadIdx.x] = a; r2[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partials[blockIdx.x * 2] = r1[0]; partials[blockIdx.x * 2 + 1] = r2[0]; }
}
int launch_outer_bn_reduce(const float* x, const float* dy, const float* mean, float* partials, int* n_partials, int B,
                           int Cin, int HW, cudaStream_t s) {
  const int nb = 296;
  outer_bn_reduce_kernel<<<nb, 256, 0, s>>>(x, dy, mean, partials, B, Cin, HW);
  RD_LAUNCHED();
  *n_partials = nb;
  return 0;
}
__global__ void outer_bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts,
                                             const float* __restrict__ invstd, float* __restrict__ dgamma,
                                             float* __restrict__ dbeta) {
  if (threadIdx.x != 0) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < nparts; ++i) { s1 += (double)partials[i * 2]; s2 += (double)partials[i * 2 + 1]; }
  dbeta[0] = (float)s1;
  dgamma[0] = (float)(s2 * (double)invstd[0]);
}
int launch_outer_bn_bwd_finalize(const float* partials, int nparts, const float* invstd, float* dgamma, float* dbeta,
                                 cudaStream_t s) {
  outer_bn_bwd_finalize_kernel<<<1, 32, 0, s>>>(partials, nparts, invstd, dgamma, dbeta);
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// up_mode='bilinear': nn.Upsample(scale_factor=2, mode='bilinear') (align_corners=False) followed by a 1x1 conv
// (lib/UNet.py:20).  The 1x1 conv commutes with the interpolation (the weights of every output pixel sum to 1),
// so it runs at the LOW resolution as a 1-tap GEMM and these kernels interpolate its output:
//   forward : u[b,oy,ox,c] = bilinear(t)[oy,ox,c] + bias[c] + skip[b,oy,ox,c]
//   adjoint : dt[b,i,j,c]  = sum over the output pixels that read (i,j) of weight * du
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, int n_in, int& i0, int& i1, float& l1) {
  float src = 0.5f * ((float)o + 0.5f) - 0.5f;          // area_pixel_compute_source_index, align_corners=False
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
__global__ void __launch_bounds__(256)
bilinear_up_add_kernel(const float* __restrict__ t, const float* __restrict__ bias, const float* __restrict__ skip,
                       float* __restrict__ u, int B, int Hin, int Win, int C, int rnd) {
  const int Q = C >> 2, Ho = 2 * Hin, Wo = 2 * Win;
  const long long total = (long long)B * Ho * Wo * Q;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int q = (int)(i % Q);
    long long p = i / Q;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(oy, Hin, y0, y1, ly);
    bilinear_src(ox, Win, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* base = t + (size_t)b * Hin * Win * C + q * 4;
    const float4 v00 = ld4(base + ((size_t)y0 * Win + x0) * C), v01 = ld4(base + ((size_t)y0 * Win + x1) * C);
    const float4 v10 = ld4(base + ((size_t)y1 * Win + x0) * C), v11 = ld4(base + ((size_t)y1 * Win + x1) * C);
    const float4 bv = ld4(bias + q * 4);
    float4 r;
    r.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x) + bv.x;
    r.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y) + bv.y;
    r.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z) + bv.z;
    r.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w) + bv.w;
    if (skip) {
      const float4 sv = ld4(skip + i * 4);
      r.x += sv.x; r.y += sv.y; r.z += sv.z; r.w += sv.w;
    }
    st4(u + i * 4, rnd ? tf32_rn4(r) : r);
  }
}
int launch_bilinear_up_add(const float* t, const float* bias, const float* skip, float* u, int B, int Hin, int Win,
                           int C, int rnd, cudaStream_t s) {
  if (C % 4) return fail("bilinear_up: C=%d not a multiple of 4", C);
  bilinear_up_add_kernel<<<ew_grid((long long)B * 4 * Hin * Win * (C / 4)), 256, 0, s>>>(t, bias, skip, u, B, Hin, Win,
                                                                                        C, rnd);
  RD_LAUNCHED();
  return 0;
}
__global__ void __launch_bounds__(256)
bilinear_up_adjoint_kernel(const float* __restrict__ du, float* __restrict__ dt, int B, int Hin, int Win, int C,
                           int rnd) {
  const int Q = C >> 2, Ho = 2 * Hin, Wo = 2 * Win;
  const long long total = (long long)B * Hin * Win * Q;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int q = (int)(i % Q);
    long long p = i / Q;
    const int ix = (int)(p % Win); p /= Win;
    const int iy = (int)(p % Hin);
    const int b = (int)(p / Hin);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int oy = 2 * iy - 2; oy <= 2 * iy + 2; ++oy) {
      if (oy < 0 || oy >= Ho) continue;
      int y0, y1;
      float ly;
      bilinear_src(oy, Hin, y0, y1, ly);
      const float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int ox = 2 * ix - 2; ox <= 2 * ix + 2; ++ox) {
        if (ox < 0 || ox >= Wo) continue;
        int x0, x1;
        float lx;
        bilinear_src(ox, Win, x0, x1, lx);
        const float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
        if (wx == 0.f) continue;
        const float4 g = ld4(du + (((size_t)b * Ho + oy) * Wo + ox) * C + q * 4);
        const float w = wy * wx;
        acc.x = fmaf(w, g.x, acc.x); acc.y = fmaf(w, g.y, acc.y); acc.z = fmaf(w, g.z, acc.z); acc.w = fmaf(w, g.w, acc.w);
      }
    }
    st4(dt + i * 4, rnd ? tf32_rn4(acc) : acc);
  }
}
int launch_bilinear_up_adjoint(const float* du, float* dt, int B, int Hin, int Win, int C, int rnd, cudaStream_t s) {
  bilinear_up_adjoint_kernel<<<ew_grid((long long)B * Hin * Win * (C / 4)), 256, 0, s>>>(du, dt, B, Hin, Win, C, rnd);
  RD_LAUNCHED();
  return 0;
}
// 1x1 conv weight [Co][Ci] -> its transpose (and TF32 rounding of both copies when requested)
__global__ void pack_conv1x1_kernel(const float* __restrict__ w, float* __restrict__ w_copy, float* __restrict__ w_t,
                                    int Co, int Ci, int rnd) {
  const int total = Co * Ci;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int ci = i % Ci, co = i / Ci;
    float v = w[i];
    if (rnd) v = tf32_rn(v);
    w_copy[i] = v;
    w_t[(size_t)ci * Co + co] = v;
  }
}
int launch_pack_conv1x1(const float* w, float* w_copy, float* w_t, int Co, int Ci, int rnd, cudaStream_t s) {
  pack_conv1x1_kernel<<<ew_grid((long long)Co * Ci), 256, 0, s>>>(w, w_copy, w_t, Co, Ci, rnd);
  RD_LAUNCHED();
  return 0;
}

}  // namespace rd
