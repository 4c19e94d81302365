// Direct (CUDA-core) kernels for the two thin ends of the U-Net, where the contraction is too
// narrow for tensor cores and the layer is HBM-bound (enc0 AI~10, last AI~4):
//   * first conv  Cin(<=8) -> Cout, NCHW input -> NHWC output, + BatchNorm partial sums
//       replaces nn.Conv2d of encoder[0]            (reference lib/UNet.py:158-162, 4-5)
//   * last  conv  C -> 1 (+bias, +outer residual)     (reference lib/UNet.py:184,227,229-244)
//   * their weight/input gradients (autograd of the above, reference lib/Trainer.py:179)
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"

namespace rd {

// Packed fp32x2 FMA (FFMA2, sm_100): the two thin-end kernels of the network are issue-bound on CUDA cores, and one
// FFMA2 does the work of two FFMAs.  Same rounding as fmaf per component.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

static constexpr int TH = 8, TW = 32;          // pixel tile of one block
static constexpr int HALO_W = TW + 2, HALO_H = TH + 2;

// ----------------------------------------------------------------------------------------------
// first conv forward
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ z,
                      float* __restrict__ partials, int B, int Cin, int H, int W, int Cout,
                      int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ float smem[];
  float* xs = smem;                                   // [Cin][HALO_H][HALO_W]
  float* ws = xs + Cin * HALO_H * HALO_W;             // [Cin*9][Cout]
  float* red = ws + Cin * 9 * Cout;                   // [PG][Cout][2]
  const int tid = threadIdx.x;
  const int Q = Cout >> 2;
  const int PG = 256 / Q;
  for (int i = tid; i < Cin * 9 * Cout; i += 256) {   // w[co][ci][r][s] -> ws[(ci*9+r*3+s)][co]
    int co = i % Cout, k = i / Cout;
    ws[i] = w[(size_t)co * Cin * 9 + k];
  }
  const int q = tid % Q, pg = tid / Q;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int h0 = ty * TH, w0 = tx * TW;
    __syncthreads();
    for (int i = tid; i < Cin * HALO_H * HALO_W; i += 256) {
      int ww = i % HALO_W, hh = (i / HALO_W) % HALO_H, ci = i / (HALO_W * HALO_H);
      int gh = h0 + hh - 1, gw = w0 + ww - 1;
      float v = 0.f;
      if (gh >= 0 && gh < H && gw >= 0 && gw < W) v = x[(((size_t)b * Cin + ci) * H + gh) * W + gw];
      xs[i] = v;
    }
    __syncthreads();
    if (pg < PG) {
      for (int g = pg; g < (TH * TW) / 4; g += PG) {
        const int lh = g / (TW / 4), lw = (g % (TW / 4)) * 4;
        float acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;
        for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float* row = xs + (ci * HALO_H + lh + r) * HALO_W + lw;
            float in[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) in[i] = row[i];
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const float4 wv = *reinterpret_cast<const float4*>(ws + ((ci * 3 + r) * 3 + s) * Cout + q * 4);
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                acc[p][0] = fmaf(in[p + s], wv.x, acc[p][0]);
                acc[p][1] = fmaf(in[p + s], wv.y, acc[p][1]);
                acc[p][2] = fmaf(in[p + s], wv.z, acc[p][2]);
                acc[p][3] = fmaf(in[p + s], wv.w, acc[p][3]);
              }
            }
          }
        }
        const int gh = h0 + lh;
        if (gh < H) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int gw = w0 + lw + p;
            if (gw < W) {
              *reinterpret_cast<float4*>(z + (((size_t)b * H + gh) * W + gw) * Cout + q * 4) =
                  make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
#pragma unroll
              for (int c = 0; c < 4; ++c) { s1[c] += acc[p][c]; s2[c] = fmaf(acc[p][c], acc[p][c], s2[c]); }
            }
          }
        }
      }
    }
  }
  if (pg < PG) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      red[(pg * Cout + q * 4 + c) * 2 + 0] = s1[c];
      red[(pg * Cout + q * 4 + c) * 2 + 1] = s2[c];
    }
  }
  __syncthreads();
  if (partials != nullptr) {
    for (int i = tid; i < Cout * 2; i += 256) {
      float a = 0.f;
      for (int p = 0; p < PG; ++p) a += red[p * Cout * 2 + i];
      partials[(size_t)blockIdx.x * Cout * 2 + i] = a;
    }
  }
}

int launch_conv_first_fwd(const float* x, const float* w, float* z, float* partials, int* n_partials, int B,
                          int Cin, int H, int W, int Cout, cudaStream_t s) {
  if (Cout % 4 || Cout > 1024 || Cin > 8) return fail("conv_first: unsupported Cin=%d Cout=%d", Cin, Cout);
  const int tiles_x = cdiv(W, TW), tiles_y = cdiv(H, TH);
  const int ntiles = tiles_x * tiles_y * B;
  const int grid = ntiles < 148 * 4 ? ntiles : 148 * 4;
  const int Q = Cout / 4, PG = 256 / Q;
  size_t smem = sizeof(float) * ((size_t)Cin * HALO_H * HALO_W + (size_t)Cin * 9 * Cout + (size_t)PG * Cout * 2);
  if (smem > 48 * 1024)
    RD_CUDA(cudaFuncSetAttribute(conv_first_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_first_fwd_kernel<<<grid, 256, smem, s>>>(x, w, z, partials, B, Cin, H, W, Cout, tiles_x, tiles_y, ntiles);
  RD_LAUNCHED();
  if (n_partials) *n_partials = grid;
  return 0;
}

// ----------------------------------------------------------------------------------------------
// first conv wgrad:  dW[co][ci][r][s] = sum_p x[b,ci,h+r-1,w+s-1] * dz[p,co]
// ----------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256)
conv_first_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ part,
                        int B, int Cin, int H, int W, int Cout, int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ float smem[];
  float* xs = smem;                                   // [Cin][HALO_H][HALO_W]
  const int tid = threadIdx.x;
  const int Q = Cout >> 2;
  const int NS = 256 / (Q * 4);                       // pixel streams per block
  const int q = tid % Q, tg = (tid / Q) % 4, st = tid / (Q * 4);
  const int ntaps = Cin * 9;
  int off[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    int k = tg * NT + i;
    if (k >= ntaps) k = 0;
    const int ci = k / 9, r = (k % 9) / 3, sx = k % 3;
    off[i] = (ci * HALO_H + r) * HALO_W + sx;
  }
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int h0 = ty * TH, w0 = tx * TW;
    __syncthreads();
    for (int i = tid; i < Cin * HALO_H * HALO_W; i += 256) {
      int ww = i % HALO_W, hh = (i / HALO_W) % HALO_H, ci = i / (HALO_W * HALO_H);
      int gh = h0 + hh - 1, gw = w0 + ww - 1;
      float v = 0.f;
      if (gh >= 0 && gh < H && gw >= 0 && gw < W) v = x[(((size_t)b * Cin + ci) * H + gh) * W + gw];
      xs[i] = v;
    }
    __syncthreads();
    if (st < NS) {
      for (int p = st; p < TH * TW; p += NS) {
        const int lh = p / TW, lw = p % TW;
        const int gh = h0 + lh, gw = w0 + lw;
        if (gh >= H || gw >= W) continue;
        const float4 g = *reinterpret_cast<const float4*>(dz + (((size_t)b * H + gh) * W + gw) * Cout + q * 4);
        const float* base = xs + lh * HALO_W + lw;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const float xv = base[off[i]];
          acc[i][0] = fmaf(xv, g.x, acc[i][0]);
          acc[i][1] = fmaf(xv, g.y, acc[i][1]);
          acc[i][2] = fmaf(xv, g.z, acc[i][2]);
          acc[i][3] = fmaf(xv, g.w, acc[i][3]);
        }
      }
    }
  }
  // reduce the NS streams through shared memory, then write the block partial [Cout][ntaps]
  __syncthreads();
  float* red = smem;                                  // [NS][Cout*ntaps]  (re-uses xs; sized by host)
  if (st < NS) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int k = tg * NT + i;
      if (k < ntaps) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[(size_t)st * Cout * ntaps + (q * 4 + c) * ntaps + k] = acc[i][c];
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < Cout * ntaps; i += 256) {
    float a = 0.f;
    for (int s2 = 0; s2 < NS; ++s2) a += red[(size_t)s2 * Cout * ntaps + i];
    part[(size_t)blockIdx.x * Cout * ntaps + i] = a;
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int p = 0; p < nparts; ++p) a += (double)part[(size_t)p * n + i];
  out[i] = (float)a;
}

int launch_conv_first_wgrad(const float* x, const float* dz, float* dw, float* scratch, size_t scratch_floats,
                            int B, int Cin, int H, int W, int Cout, cudaStream_t s) {
  const int Q = Cout / 4;
  if (Cout % 4 || Q * 4 > 256 || Cin > 8) return fail("conv_first_wgrad: unsupported Cin=%d Cout=%d", Cin, Cout);
  const int tiles_x = cdiv(W, TW), tiles_y = cdiv(H, TH);
  const int ntiles = tiles_x * tiles_y * B;
  const int ntaps = Cin * 9;
  int grid = ntiles < 296 ? ntiles : 296;
  while ((size_t)grid * Cout * ntaps > scratch_floats && grid > 1) grid /= 2;
  if ((size_t)grid * Cout * ntaps > scratch_floats) return fail("conv_first_wgrad: scratch too small");
  const int NS = 256 / (Q * 4);
  size_t smem_x = (size_t)Cin * HALO_H * HALO_W, smem_r = (size_t)NS * Cout * ntaps;
  size_t smem = sizeof(float) * (smem_x > smem_r ? smem_x : smem_r);
  const int NT = (ntaps + 3) / 4;
#define RD_WG(NTV)                                                                                            \
  {                                                                                                           \
    RD_CUDA(cudaFuncSetAttribute(conv_first_wgrad_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem));                                                                 \
    conv_first_wgrad_kernel<NTV><<<grid, 256, smem, s>>>(x, dz, scratch, B, Cin, H, W, Cout, tiles_x, tiles_y, \
                                                         ntiles);                                             \
  }
  if (NT <= 3) RD_WG(3)
  else if (NT <= 5) RD_WG(5)
  else if (NT <= 7) RD_WG(7)
  else if (NT <= 9) RD_WG(9)
  else if (NT <= 14) RD_WG(14)
  else RD_WG(18)
#undef RD_WG
  RD_LAUNCHED();
  const int n = Cout * ntaps;
  return launch_sum_partials(scratch, grid, n, n, 1, dw, s);
}

// ----------------------------------------------------------------------------------------------
// last conv forward:  y[b,h,w] = sum_{c,r,s} u[b,h+r-1,w+s-1,c] W[c,r,s] + bias + x[b,0,h,w]
//   phase 1: per input pixel, 9 partial dot products over channels (each u element read once)
//   phase 2: per output pixel, gather the 9 partials of its 3x3 neighbourhood
// ----------------------------------------------------------------------------------------------
static constexpr int LT_H = 16, LT_W = 32;                       // output tile of one block
static constexpr int LH_H = LT_H + 2, LH_W = LT_W + 2;           // with halo
// phase 2 of both forward variants: one thread per output pixel gathers the 9 partials of its 3x3 neighbourhood
__device__ __forceinline__ void conv_last_gather(const float (*ts)[9], const float* __restrict__ bias,
                                                 const float* __restrict__ x0, long long x_bstride,
                                                 const float* __restrict__ x_affine, float* __restrict__ y, int b,
                                                 int h0, int w0, int H, int W) {
  const float bv = bias ? bias[0] : 0.f;
  for (int o = threadIdx.x; o < LT_H * LT_W; o += 256) {
    const int lh = o / LT_W, lw = o - lh * LT_W;
    const int gh = h0 + lh, gw = w0 + lw;
    if (gh < H && gw < W) {
      float a = bv;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) a += ts[(lh + r) * LH_W + lw + q][r * 3 + q];
      if (x0) {
        const float xv = x0[(size_t)b * x_bstride + (size_t)gh * W + gw];
        a += x_affine ? fmaf(xv, x_affine[0], x_affine[1]) : xv;
      }
      y[((size_t)b * H + gh) * W + gw] = a;
    }
  }
}

// Fast variant for C <= 64.  Phase 1: a half-warp owns four consecutive halo pixels per iteration; lane l holds
// channels 4l..4l+3 (one coalesced 256-byte row per pixel and half-warp, weights resident in registers), forms
// the 9 tap partials of each of the 4 pixels over its channels (36 accumulators) and the half-warp reduces them
// with a transposing butterfly: 35 shuffle+add pairs instead of 144.  The butterfly needs no selects because every
// lane keeps its accumulators in a lane-specific ORDER: pixel slot i holds pixel i ^ (lane bits 3,2) and tap slot
// s holds tap kTapPerm[lane bits 1,0][s], so that "keep the low half, send the high half" is the same instruction
// stream for both partners of every exchange.  Lane l ends with pixel (l>>2)&3 and taps kTapPerm[l&3][0..1]
// (lane bits 00 also with tap 4).
__constant__ signed char kTapPerm[4][9] = {{0, 1, 2, 3, 4, 5, 6, 7, 8},
                                           {2, 3, 0, 1, 4, 7, 8, 5, 6},
                                           {5, 6, 7, 8, 4, 0, 1, 2, 3},
                                           {7, 8, 5, 6, 4, 2, 3, 0, 1}};
__global__ void __launch_bounds__(256, 2)
conv_last_fwd64_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ x0, long long x_bstride, const float* __restrict__ x_affine,
                       float* __restrict__ y, int B, int H, int W, int C, int tiles_x, int tiles_y) {
  __shared__ float ts[LH_H * LH_W][9];
  __shared__ float wsm[64 * 9];
  const int tid = threadIdx.x;
  const int l16 = tid & 15, hw = tid >> 4;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int h0 = ty * LT_H, w0 = tx * LT_W;
  const int c0 = l16 * 4;
  for (int i = tid; i < C * 9; i += 256) wsm[i] = w[i];
  __syncthreads();
  const int tsel = l16 & 3;                 // tap order of this lane
  const int pperm = (l16 >> 2) & 3;         // pixel order of this lane (slot i holds pixel i ^ pperm)
  int tap[9];
  float4 wr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    tap[k] = kTapPerm[tsel][k];
    wr[k] = c0 < C ? make_float4(wsm[(c0 + 0) * 9 + tap[k]], wsm[(c0 + 1) * 9 + tap[k]], wsm[(c0 + 2) * 9 + tap[k]],
                                 wsm[(c0 + 3) * 9 + tap[k]])
                   : make_float4(0, 0, 0, 0);
  }
  const int tap0 = tap[0], tap1 = tap[1];
  constexpr int NPIX = LH_H * LH_W, NGROUPS = (NPIX + 3) / 4;
  auto load_group = [&](int g, float4 (&dst)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = g * 4 + (i ^ pperm);
      const int hh = p / LH_W, ww = p - hh * LH_W;
      const int gh = h0 + hh - 1, gw = w0 + ww - 1;
      const bool ok = p < NPIX && (unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W && c0 < C;
      dst[i] = ok ? __ldg(reinterpret_cast<const float4*>(u + (((size_t)b * H + gh) * W + gw) * C + c0))
                  : make_float4(0, 0, 0, 0);
    }
  };
  float4 uv[4], nx[4];
  load_group(hw, uv);
#pragma unroll 1
  for (int g0 = 0; g0 < NGROUPS; g0 += 16) {
    const int g = g0 + hw;
    load_group(g + 16, nx);                            // next iteration's rows are in flight during the butterfly
    float v[36];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 ux = make_float2(uv[i].x, uv[i].x), uy = make_float2(uv[i].y, uv[i].y);
      const float2 uz = make_float2(uv[i].z, uv[i].z), uw = make_float2(uv[i].w, uv[i].w);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {                   // two tap slots per FFMA2, same summation order as the scalar form
        float2 acc = make_float2(uv[i].w * wr[k].w, uv[i].w * wr[k + 1].w);
        acc = ffma2(uz, make_float2(wr[k].z, wr[k + 1].z), acc);
        acc = ffma2(uy, make_float2(wr[k].y, wr[k + 1].y), acc);
        acc = ffma2(ux, make_float2(wr[k].x, wr[k + 1].x), acc);
        v[i * 9 + k] = acc.x;
        v[i * 9 + k + 1] = acc.y;
      }
      (void)uw;
      v[i * 9 + 8] = fmaf(uv[i].x, wr[8].x, fmaf(uv[i].y, wr[8].y, fmaf(uv[i].z, wr[8].z, uv[i].w * wr[8].w)));
    }
#pragma unroll
    for (int j = 0; j < 18; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j + 18], 8);     // pixel pairs
#pragma unroll
    for (int j = 0; j < 9; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j + 9], 4);       // pixels
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j + 5], 2);       // tap groups {0-3|5-8}, 4
    v[4] += __shfl_xor_sync(0xffffffffu, v[4], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[2], 1);                                       // tap pairs
    v[1] += __shfl_xor_sync(0xffffffffu, v[3], 1);
    v[4] += __shfl_xor_sync(0xffffffffu, v[4], 1);
    const int p = g * 4 + pperm;
    if (p < NPIX) {
      ts[p][tap0] = v[0];
      ts[p][tap1] = v[1];
      if (tsel == 0) ts[p][4] = v[4];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) uv[i] = nx[i];
  }
  __syncthreads();
  conv_last_gather(ts, bias, x0, x_bstride, x_affine, y, b, h0, w0, H, W);
}

__device__ __forceinline__ void lf_cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const uint32_t n = pred ? 16u : 0u;                  // src-size 0: zero fill, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

__global__ void __launch_bounds__(256)
conv_last_fwd_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ x0, long long x_bstride, const float* __restrict__ x_affine,
                     float* __restrict__ y, int B, int H, int W, int C, int tiles_x, int tiles_y) {
  __shared__ float ts[LH_H * LH_W][9];                            // per input pixel: its 9 tap partial sums
  __shared__ __align__(16) float ws[128 * 9 + 12];                // W[c][t]
  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int h0 = ty * LT_H, w0 = tx * LT_W;
  for (int i = tid; i < C * 9; i += 256) ws[i] = w[i];
  __syncthreads();
  // phase 1: one thread per (halo) input pixel streams its C channels once and forms all 9 tap dot products
  for (int p = tid; p < LH_H * LH_W; p += 256) {
    const int hh = p / LH_W, ww = p - hh * LH_W;
    const int gh = h0 + hh - 1, gw = w0 + ww - 1;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    if (gh >= 0 && gh < H && gw >= 0 && gw < W) {
      const float4* up = reinterpret_cast<const float4*>(u + (((size_t)b * H + gh) * W + gw) * C);
#pragma unroll 4
      for (int cq = 0; cq < (C >> 2); ++cq) {
        const float4 v = __ldg(up + cq);
        const float* wq = ws + cq * 36;                            // 4 channels x 9 taps
        float wr[36];
#pragma unroll
        for (int j = 0; j < 9; ++j) *reinterpret_cast<float4*>(wr + 4 * j) = *reinterpret_cast<const float4*>(wq + 4 * j);
#pragma unroll
        for (int k = 0; k < 9; ++k)
          acc[k] += v.x * wr[k] + v.y * wr[9 + k] + v.z * wr[18 + k] + v.w * wr[27 + k];
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) ts[p][k] = acc[k];
  }
  __syncthreads();
  conv_last_gather(ts, bias, x0, x_bstride, x_affine, y, b, h0, w0, H, W);
}

// ----------------------------------------------------------------------------------------------
// last conv forward, thread-per-pixel form (C = 32 / 64).  The half-warp-per-pixel kernels above spend most of their
// issue slots on the cross-lane reduction and on index arithmetic (ncu, round 2: 58 % SM throughput at 37 % DRAM).
// Here a pixel's 9 tap partials t[p][k] = sum_c u[p][c] W[c][k] are formed by ONE thread, so there is nothing to
// reduce, and a warp instruction serves 32 pixels with one broadcast weight load:
//   pass 1 (conv_last_taps_kernel): persistent CTAs stream contiguous chunks of 256 pixels (256 x C floats, one
//     contiguous block of the NHWC tensor) through a 3-stage cp.async ring into shared memory rows padded to C*4 + 16
//     bytes (conflict-free LDS.128 per thread); per channel quad 1 + 9 LDS.128 and 18 FFMA2; the 9 partials go to
//     planar scratch t[k][pixel] (coalesced 128-byte stores).  No halo: every u element is read exactly once.
//   pass 2 (conv_last_gather_kernel): y[q] = bias + x0[q] + sum_k t[k][q + off(k)]   (36 B per pixel, L2-resident)
// ----------------------------------------------------------------------------------------------
static constexpr int LTP_PIX = 256, LTP_NST = 3;
// PPT pixels per thread (256 / PPT threads per CTA): the 9 weight loads of a channel quad are shared by the PPT pixels --
// with one pixel per thread the kernel was bound by the shared-memory pipe (ncu round 2: stall_mio + short scoreboard
// 48 %, 160 LDS.128 per pixel); two pixels per thread need 88 per pixel.
template <int C, int PPT>
__global__ void __launch_bounds__(256 / PPT, 1)
conv_last_taps_kernel(const float* __restrict__ u, const float* __restrict__ w, float* __restrict__ taps, long long NP) {
  constexpr int NT = 256 / PPT;
  constexpr int Q = C / 4, ROWB = C * 4 + 16, STAGE = LTP_PIX * ROWB;
  extern __shared__ __align__(16) uint8_t ltp_smem[];
  float* wq = reinterpret_cast<float*>(ltp_smem + LTP_NST * STAGE);       // [Q][36]: 4 x (k 0..7), then k = 8 of the 4 channels
  const int tid = threadIdx.x;
  for (int i = tid; i < C * 9; i += NT) {
    const int c = i / 9, k = i - c * 9;
    wq[(c >> 2) * 36 + (k < 8 ? (c & 3) * 8 + k : 32 + (c & 3))] = w[i];
  }
  const long long nchunks = (NP + LTP_PIX - 1) / LTP_PIX;
  auto issue = [&](long long n) {                      // n-th chunk of this CTA
    const long long chunk = (long long)blockIdx.x + n * gridDim.x;
    if (chunk < nchunks) {
      uint8_t* dst = ltp_smem + (int)(n % LTP_NST) * STAGE;
      const long long p0 = chunk * LTP_PIX;
#pragma unroll
      for (int j = 0; j < Q * PPT; ++j) {
        const int idx = j * NT + tid;
        const int px = idx / Q, ch = idx % Q;
        const bool ok = p0 + px < NP;
        lf_cp_async16(dst + px * ROWB + ch * 16, ok ? u + (size_t)(p0 + px) * C + ch * 4 : u, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int n = 0; n < LTP_NST - 1; ++n) issue(n);
  for (long long n = 0;; ++n) {
    const long long chunk = (long long)blockIdx.x + n * gridDim.x;
    if (chunk >= nchunks) break;
    asm volatile("cp.async.wait_group %0;" ::"n"(LTP_NST - 2) : "memory");
    __syncthreads();                                   // chunk n has landed for every thread; chunk n-1's stage is free
    issue(n + LTP_NST - 1);
    const uint8_t* stage = ltp_smem + (int)(n % LTP_NST) * STAGE;
    float2 a01[PPT], a23[PPT], a45[PPT], a67[PPT], a8[PPT];
#pragma unroll
    for (int e = 0; e < PPT; ++e) a01[e] = a23[e] = a45[e] = a67[e] = a8[e] = make_float2(0.f, 0.f);
#pragma unroll
    for (int cq = 0; cq < Q; ++cq) {
      const float4* wp = reinterpret_cast<const float4*>(wq + cq * 36);
      float4 uv[PPT];
#pragma unroll
      for (int e = 0; e < PPT; ++e) uv[e] = *reinterpret_cast<const float4*>(stage + (tid + e * NT) * ROWB + cq * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 wa = wp[2 * j], wb = wp[2 * j + 1];
#pragma unroll
        for (int e = 0; e < PPT; ++e) {
          const float us = j == 0 ? uv[e].x : (j == 1 ? uv[e].y : (j == 2 ? uv[e].z : uv[e].w));
          const float2 ub = make_float2(us, us);
          a01[e] = ffma2(ub, make_float2(wa.x, wa.y), a01[e]);
          a23[e] = ffma2(ub, make_float2(wa.z, wa.w), a23[e]);
          a45[e] = ffma2(ub, make_float2(wb.x, wb.y), a45[e]);
          a67[e] = ffma2(ub, make_float2(wb.z, wb.w), a67[e]);
        }
      }
      const float4 w8 = wp[8];
#pragma unroll
      for (int e = 0; e < PPT; ++e) {
        a8[e] = ffma2(make_float2(uv[e].x, uv[e].y), make_float2(w8.x, w8.y), a8[e]);
        a8[e] = ffma2(make_float2(uv[e].z, uv[e].w), make_float2(w8.z, w8.w), a8[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < PPT; ++e) {
      const long long p = chunk * LTP_PIX + tid + e * NT;
      if (p < NP) {
        taps[0 * NP + p] = a01[e].x; taps[1 * NP + p] = a01[e].y; taps[2 * NP + p] = a23[e].x; taps[3 * NP + p] = a23[e].y;
        taps[4 * NP + p] = a45[e].x; taps[5 * NP + p] = a45[e].y; taps[6 * NP + p] = a67[e].x; taps[7 * NP + p] = a67[e].y;
        taps[8 * NP + p] = a8[e].x + a8[e].y;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256)
conv_last_gather_kernel(const float* __restrict__ taps, const float* __restrict__ bias, const float* __restrict__ x0,
                        long long x_bstride, const float* __restrict__ x_affine, float* __restrict__ y, int B, int H, int W) {
  const long long NP = (long long)B * H * W;
  const float bv = bias ? bias[0] : 0.f;
  const float xs = x_affine ? x_affine[0] : 1.f, xo = x_affine ? x_affine[1] : 0.f;
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < NP; p += gridDim.x * 256LL) {
    const int gw = (int)(p % W);
    const long long r_ = p / W;
    const int gh = (int)(r_ % H);
    const long long b = r_ / H;
    float a = bv;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hh = gh + r - 1;
      if ((unsigned)hh >= (unsigned)H) continue;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int ww = gw + q - 1;
        if ((unsigned)ww < (unsigned)W) a += __ldg(taps + (size_t)(r * 3 + q) * NP + p + (long long)(r - 1) * W + (q - 1));
      }
    }
    if (x0) a += fmaf(__ldg(x0 + (size_t)b * x_bstride + (size_t)gh * W + gw), xs, xo);
    y[p] = a;
  }
}

int launch_conv_last_fwd(const float* u, const float* w, const float* bias, const float* x, int x_bstride,
                         const float* x_affine, float* y, int B, int H, int W, int C, float* tap_scratch, cudaStream_t s) {
  if (C % 4 || C > 128) return fail("conv_last: unsupported C=%d (needs C%%4==0, C<=128)", C);
  static const bool no_tp = getenv("RESDEPTH_LAST_FWD_OLD") != nullptr;
  if (tap_scratch && !no_tp && (C == 64 || C == 32)) {
    // thread-per-pixel tap partials (planar scratch, 9 floats per pixel), then the 3x3 gather
    const long long NP = (long long)B * H * W;
    const long long nchunks = (NP + LTP_PIX - 1) / LTP_PIX;
    const int grid = (int)(nchunks < 148 ? nchunks : 148);
    static const int ppt = getenv("RESDEPTH_TAPS_PPT") ? atoi(getenv("RESDEPTH_TAPS_PPT")) : 2;
#define RD_TAPS(CC, PP)                                                                                              \
  {                                                                                                                  \
    const int smem = LTP_NST * LTP_PIX * (CC * 4 + 16) + CC * 9 * 4;                                                 \
    RD_CUDA(cudaFuncSetAttribute(conv_last_taps_kernel<CC, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    conv_last_taps_kernel<CC, PP><<<grid, 256 / PP, smem, s>>>(u, w, tap_scratch, NP);                               \
  }
    if (C == 64) { if (ppt == 1) RD_TAPS(64, 1) else RD_TAPS(64, 2) }
    else { if (ppt == 1) RD_TAPS(32, 1) else RD_TAPS(32, 2) }
#undef RD_TAPS
    RD_LAUNCHED();
    const long long blocks = (NP + 255) / 256;
    conv_last_gather_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, s>>>(tap_scratch, bias, x, x_bstride, x_affine,
                                                                                    y, B, H, W);
    RD_LAUNCHED();
    return 0;
  }
  const int tiles_x = cdiv(W, LT_W), tiles_y = cdiv(H, LT_H);
  const int grid = tiles_x * tiles_y * B;
  if (C <= 64)
    conv_last_fwd64_kernel<<<grid, 256, 0, s>>>(u, w, bias, x, x_bstride, x_affine, y, B, H, W, C, tiles_x, tiles_y);
  else
    conv_last_fwd_kernel<<<grid, 256, 0, s>>>(u, w, bias, x, x_bstride, x_affine, y, B, H, W, C, tiles_x, tiles_y);
  RD_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// last conv backward: for input pixel q and channel c, with n[r][s] = dy[q - (r-1, s-1)]:
//   du[q,c] = sum_{r,s} n[r][s] W[c,r,s];   dW[c,r,s] += u[q,c] n[r][s];   db += dy[q]
// ----------------------------------------------------------------------------------------------
// MODE 0: everything in one pass (default); MODE 1: du (+ its channel sums, db) only; MODE 2: dW only.  The split
// pair (RESDEPTH_LAST_BWD_SPLIT=1) was tried against the fused pass's register pressure (126 registers, 2 CTAs per
// SM, ncu: 37 % DRAM) and measured slower on B200: 0.29 + 0.34 ms against 0.55 ms.
// pixel tile of one block iteration: large, so that the two block-wide barriers around the dy halo load amortise
static constexpr int LB_TH = 32, LB_TW = 32, LB_HH = LB_TH + 2, LB_HW = LB_TW + 2;
template <int QPL, int MODE>
__global__ void __launch_bounds__(256, MODE == 0 ? 2 : 3)
conv_last_bwd_kernel(const float* __restrict__ u, const float* __restrict__ dy, const float* __restrict__ w,
                     float* __restrict__ du, __nv_bfloat16* __restrict__ du_b, float* __restrict__ part, int B, int H,
                     int W, int C, int tiles_x, int tiles_y, int ntiles) {
  __shared__ float dys[LB_HH * LB_HW];
  extern __shared__ float red_dyn[];                  // [16][RS]
  constexpr int NDW = 9 * 4 * 16 * QPL;               // weight-gradient entries, then 64*QPL channel sums of du, then db
  constexpr int RS = NDW + 4 * 16 * QPL + 1;
  const int tid = threadIdx.x;
  const int lane16 = tid & 15, grp = tid >> 4;
  float4 wr[QPL][9], dwacc[QPL][9], dusum[QPL];
  float dbacc = 0.f;
#pragma unroll
  for (int j = 0; j < QPL; ++j) {
    const int c = (lane16 + 16 * j) * 4;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (MODE != 2 && c < C) wr[j][k] = make_float4(w[(c + 0) * 9 + k], w[(c + 1) * 9 + k], w[(c + 2) * 9 + k], w[(c + 3) * 9 + k]);
      else wr[j][k] = make_float4(0, 0, 0, 0);
      dwacc[j][k] = make_float4(0, 0, 0, 0);
    }
    dusum[j] = make_float4(0, 0, 0, 0);
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int h0 = ty * LB_TH, w0 = tx * LB_TW;
    __syncthreads();
    for (int i = tid; i < LB_HH * LB_HW; i += 256) {
      const int hh = i / LB_HW, ww = i % LB_HW;
      const int gh = h0 + hh - 1, gw = w0 + ww - 1;
      dys[i] = (gh >= 0 && gh < H && gw >= 0 && gw < W) ? dy[((size_t)b * H + gh) * W + gw] : 0.f;
    }
    __syncthreads();
    // four pixels per iteration: their u loads are issued together (memory-level parallelism), then consumed
    for (int p0 = grp; p0 < LB_TH * LB_TW; p0 += 64) {
      float4 uvs[4][QPL];
      bool okp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = p0 + 16 * i;
        const int lh = p / LB_TW, lw = p % LB_TW;
        const int gh = h0 + lh, gw = w0 + lw;
        okp[i] = p < LB_TH * LB_TW && gh < H && gw < W;
        const size_t o = (((size_t)b * H + gh) * W + gw) * C;
#pragma unroll
        for (int j = 0; j < QPL; ++j) {
          const int c = (lane16 + 16 * j) * 4;
          uvs[i][j] = (MODE != 1 && okp[i] && c < C) ? __ldg(reinterpret_cast<const float4*>(u + o + c)) : make_float4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!okp[i]) continue;
        const int p = p0 + 16 * i;
        const int lh = p / LB_TW, lw = p % LB_TW;
        const int gh = h0 + lh, gw = w0 + lw;
        float n[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) n[r * 3 + s2] = dys[(lh + 1 - (r - 1)) * LB_HW + (lw + 1 - (s2 - 1))];
        if (MODE != 2 && lane16 == 0) dbacc += n[4];
        const size_t o = (((size_t)b * H + gh) * W + gw) * C;
#pragma unroll
        for (int j = 0; j < QPL; ++j) {
          const int c = (lane16 + 16 * j) * 4;
          if (c < C) {
            const float4 uv = uvs[i][j];
            float4 d = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
              const float2 nk = make_float2(n[k], n[k]);
              if (MODE != 2) {
                const float2 lo = ffma2(nk, make_float2(wr[j][k].x, wr[j][k].y), make_float2(d.x, d.y));
                const float2 hi = ffma2(nk, make_float2(wr[j][k].z, wr[j][k].w), make_float2(d.z, d.w));
                d = make_float4(lo.x, lo.y, hi.x, hi.y);
              }
              if (MODE != 1) {
                const float2 lo = ffma2(make_float2(uv.x, uv.y), nk, make_float2(dwacc[j][k].x, dwacc[j][k].y));
                const float2 hi = ffma2(make_float2(uv.z, uv.w), nk, make_float2(dwacc[j][k].z, dwacc[j][k].w));
                dwacc[j][k] = make_float4(lo.x, lo.y, hi.x, hi.y);
              }
            }
            if (MODE == 2) continue;
            if (du) *reinterpret_cast<float4*>(du + o + c) = d;
            if (du_b) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(d.x, d.y), hi = __floats2bfloat162_rn(d.z, d.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(du_b + o + c) = pk;
            }
            dusum[j].x += d.x; dusum[j].y += d.y; dusum[j].z += d.z; dusum[j].w += d.w;
          }
        }
      }
    }
  }
  // block reduction over the 16 pixel groups -> part[blk][C*9 + 1]
#pragma unroll
  for (int j = 0; j < QPL; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int base = ((j * 16 + lane16) * 9 + k) * 4;
      float* rr = red_dyn + grp * RS + base;
      rr[0] = dwacc[j][k].x; rr[1] = dwacc[j][k].y; rr[2] = dwacc[j][k].z; rr[3] = dwacc[j][k].w;
    }
#pragma unroll
  for (int j = 0; j < QPL; ++j) {
    float* rr = red_dyn + grp * RS + NDW + (j * 16 + lane16) * 4;
    rr[0] = dusum[j].x; rr[1] = dusum[j].y; rr[2] = dusum[j].z; rr[3] = dusum[j].w;
  }
  if (lane16 == 0) red_dyn[grp * RS + RS - 1] = dbacc;
  __syncthreads();
  const size_t prow = (size_t)blockIdx.x * (C * 10 + 1);   // partial row: [C*9 dW][1 db][C channel sums of du]
  for (int i = tid; i < RS; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) a += red_dyn[g * RS + i];
    if (i == RS - 1) {
      if (MODE != 2) part[prow + C * 9] = a;
    } else if (i >= NDW) {
      const int c = i - NDW;
      if (MODE != 2 && c < C) part[prow + C * 9 + 1 + c] = a;
    } else if (MODE != 1) {
      const int cc = i & 3, k = (i >> 2) % 9, ql = (i >> 2) / 9;   // ql = j*16 + lane16 = channel quad
      const int c = ql * 4 + cc;
      if (c < C) part[prow + c * 9 + k] = a;
    }
  }
}

// Same pass for C <= 64, restructured after two ncu captures (round 2).  (1) The kernel above issued its u loads in the
// iteration that consumes them: 37 % of the warp samples sat on the first FFMA2 of a fresh row.  Here the rows are
// fetched by cp.async into a per-16-lane-group ring, LB_NST - 1 iterations (4 pixels = 1 KB per group) ahead and across
// tile boundaries; every lane reads back exactly the 16 bytes it copied, so the ring needs no cross-lane
// synchronisation, and its shared memory is reused by the block reduction at the end.  (2) With the latency hidden the
// kernel was issue-bound on index arithmetic (~1000 instructions per 4-pixel iteration for 144 FFMA2): an iteration now
// covers four CONSECUTIVE pixels of one row, so the nine dy taps of each pixel come from one 3 x 6 register window
// (six vector loads), output addresses are one base plus i*C, and tile coordinates are decoded once per tile.
static constexpr int LB_NST = 4;
static constexpr int LB_DW = 36;                        // padded row length of the dy halo tile (vector loads)
__device__ __forceinline__ void lb_cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const uint32_t n = pred ? 16u : 0u;                  // src-size 0: zero fill, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
struct LbCursor {                                       // position of a group in its CTA's tile sequence
  int t, j, b, h0, w0;
};
__global__ void __launch_bounds__(256, 2)
conv_last_bwd_ring_kernel(const float* __restrict__ u, const float* __restrict__ dy, const float* __restrict__ w,
                          float* __restrict__ du, __nv_bfloat16* __restrict__ du_b, float* __restrict__ part, int B, int H,
                          int W, int C, int tiles_x, int tiles_y, int ntiles) {
  __shared__ __align__(16) float dys[LB_HH * LB_DW];
  extern __shared__ __align__(16) float red_dyn[];    // ring [16 groups][LB_NST][4 pixels][16 lanes] float4, then [16][RS]
  constexpr int NDW = 9 * 4 * 16;
  constexpr int RS = NDW + 4 * 16 + 1;
  constexpr int ITERS = LB_TH * LB_TW / 64;           // iterations of one group per tile (16)
  static_assert(LB_TW == 32 && LB_TH == 32 && ITERS == 16, "unit decomposition below assumes 32 x 32 tiles");
  const int tid = threadIdx.x;
  const int lane16 = tid & 15, grp = tid >> 4;
  const int rpar = grp >> 3, cb4 = (grp & 7) * 4;     // this group's row parity and first column inside a tile
  const int c = lane16 * 4;
  const bool cok = c < C;
  float4 wr[9], dwacc[9], dusum = make_float4(0, 0, 0, 0);
  float dbacc = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wr[k] = cok ? make_float4(w[(c + 0) * 9 + k], w[(c + 1) * 9 + k], w[(c + 2) * 9 + k], w[(c + 3) * 9 + k])
                : make_float4(0, 0, 0, 0);
    dwacc[k] = make_float4(0, 0, 0, 0);
  }
  float4* ring = reinterpret_cast<float4*>(red_dyn) + (size_t)grp * LB_NST * 64;
  auto decode = [&](LbCursor& cu) {
    int t = cu.t;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    cu.b = t / tiles_y;
    cu.h0 = ty * LB_TH; cu.w0 = tx * LB_TW;
  };
  LbCursor pf{(int)blockIdx.x, 0, 0, 0, 0}, cs = pf;
  if (pf.t < ntiles) { decode(pf); cs = pf; }
  int pf_slot = 0, cs_slot = 0;
  auto issue = [&]() {
    if (pf.t < ntiles) {
      const int gh = pf.h0 + 2 * pf.j + rpar, gw0 = pf.w0 + cb4;
      const bool rowok = cok && gh < H;
      const float* src = u + (((size_t)pf.b * H + gh) * W + gw0) * C + c;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = rowok && gw0 + i < W;
        lb_cp_async16(&ring[(pf_slot * 4 + i) * 16 + lane16], ok ? src + (size_t)i * C : u, ok);
      }
      if (++pf.j == ITERS) {
        pf.j = 0;
        pf.t += gridDim.x;
        if (pf.t < ntiles) decode(pf);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // empty groups keep the group count uniform
    pf_slot = (pf_slot + 1) & (LB_NST - 1);
  };
#pragma unroll
  for (int k = 0; k < LB_NST - 1; ++k) issue();
#pragma unroll 1
  while (cs.t < ntiles) {
    if (cs.j == 0) {                                    // next tile: its dy halo (block-uniform branch)
      __syncthreads();
      for (int i = tid; i < LB_HH * LB_DW; i += 256) {
        const int hh = i / LB_DW, ww = i - hh * LB_DW;
        const int gh = cs.h0 + hh - 1, gw = cs.w0 + ww - 1;
        dys[i] = (ww < LB_HW && gh >= 0 && gh < H && gw >= 0 && gw < W) ? dy[((size_t)cs.b * H + gh) * W + gw] : 0.f;
      }
      __syncthreads();
    }
    issue();
    asm volatile("cp.async.wait_group %0;" ::"n"(LB_NST - 1) : "memory");
    const int lh = 2 * cs.j + rpar;
    const int gh = cs.h0 + lh, gw0 = cs.w0 + cb4;
    if (gh < H && gw0 < W) {
      // dy window: halo rows lh..lh+2, halo columns cb4..cb4+5; tap (r, s) of pixel i reads win[2 - r][i + 2 - s]
      float win[3][6];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float4 q4 = *reinterpret_cast<const float4*>(&dys[(lh + a) * LB_DW + cb4]);
        const float2 q2 = *reinterpret_cast<const float2*>(&dys[(lh + a) * LB_DW + cb4 + 4]);
        win[a][0] = q4.x; win[a][1] = q4.y; win[a][2] = q4.z; win[a][3] = q4.w; win[a][4] = q2.x; win[a][5] = q2.y;
      }
      const size_t o0 = (((size_t)cs.b * H + gh) * W + gw0) * C + c;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (gw0 + i >= W) break;
        if (lane16 == 0) dbacc += win[1][i + 1];
        if (!cok) continue;
        const float4 uv = ring[(cs_slot * 4 + i) * 16 + lane16];
        float4 d = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float nv = win[2 - k / 3][i + 2 - k % 3];
          const float2 nk = make_float2(nv, nv);
          const float2 lo = ffma2(nk, make_float2(wr[k].x, wr[k].y), make_float2(d.x, d.y));
          const float2 hi = ffma2(nk, make_float2(wr[k].z, wr[k].w), make_float2(d.z, d.w));
          d = make_float4(lo.x, lo.y, hi.x, hi.y);
          const float2 lo2 = ffma2(make_float2(uv.x, uv.y), nk, make_float2(dwacc[k].x, dwacc[k].y));
          const float2 hi2 = ffma2(make_float2(uv.z, uv.w), nk, make_float2(dwacc[k].z, dwacc[k].w));
          dwacc[k] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
        }
        const size_t o = o0 + (size_t)i * C;
        if (du) *reinterpret_cast<float4*>(du + o) = d;
        if (du_b) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(d.x, d.y), hi = __floats2bfloat162_rn(d.z, d.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(du_b + o) = pk;
        }
        dusum.x += d.x; dusum.y += d.y; dusum.z += d.z; dusum.w += d.w;
      }
    }
    cs_slot = (cs_slot + 1) & (LB_NST - 1);
    if (++cs.j == ITERS) {
      cs.j = 0;
      cs.t += gridDim.x;
      if (cs.t < ntiles) decode(cs);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();                                      // the ring is dead: its memory becomes the reduction buffer
  // block reduction over the 16 pixel groups -> part[blk][C*9 dW][1 db][C channel sums of du]
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float* rr = red_dyn + grp * RS + (lane16 * 9 + k) * 4;
    rr[0] = dwacc[k].x; rr[1] = dwacc[k].y; rr[2] = dwacc[k].z; rr[3] = dwacc[k].w;
  }
  {
    float* rr = red_dyn + grp * RS + NDW + lane16 * 4;
    rr[0] = dusum.x; rr[1] = dusum.y; rr[2] = dusum.z; rr[3] = dusum.w;
  }
  if (lane16 == 0) red_dyn[grp * RS + RS - 1] = dbacc;
  __syncthreads();
  const size_t prow = (size_t)blockIdx.x * (C * 10 + 1);
  for (int i = tid; i < RS; i += 256) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) a += red_dyn[g * RS + i];
    if (i == RS - 1) {
      part[prow + C * 9] = a;
    } else if (i >= NDW) {
      const int cc = i - NDW;
      if (cc < C) part[prow + C * 9 + 1 + cc] = a;
    } else {
      const int cc = i & 3, k = (i >> 2) % 9, ql = (i >> 2) / 9;
      const int ch = ql * 4 + cc;
      if (ch < C) part[prow + ch * 9 + k] = a;
    }
  }
}

int launch_conv_last_bwd(const float* u, const float* dy, const float* w, float* du, void* du_b, float* dw,
                         float* dbias, float* du_channel_sum, float* scratch, size_t scratch_floats, int B, int H, int W,
                         int C, cudaStream_t s) {
  if (C % 4 || C > 128) return fail("conv_last_bwd: unsupported C=%d", C);
  const int tiles_x = cdiv(W, LB_TW), tiles_y = cdiv(H, LB_TH);
  const int ntiles = tiles_x * tiles_y * B;
  int grid = ntiles < 148 * 4 ? ntiles : 148 * 4;
  const int PN = C * 10 + 1;
  if ((size_t)grid * PN > scratch_floats) return fail("conv_last_bwd: scratch too small");
  static const bool split = getenv("RESDEPTH_LAST_BWD_SPLIT") != nullptr;
  if (C <= 64 && split) {
    // two passes (same grid, disjoint columns of the same partial rows): du + channel sums + db, then dW
    const int smem = 16 * (10 * 4 * 16 * 1 + 1) * (int)sizeof(float);
    conv_last_bwd_kernel<1, 1><<<grid, 256, smem, s>>>(u, dy, w, du, reinterpret_cast<__nv_bfloat16*>(du_b), scratch,
                                                       B, H, W, C, tiles_x, tiles_y, ntiles);
    RD_LAUNCHED();
    conv_last_bwd_kernel<1, 2><<<grid, 256, smem, s>>>(u, dy, w, nullptr, nullptr, scratch, B, H, W, C, tiles_x, tiles_y,
                                                       ntiles);
  } else if (C <= 64) {
    static const bool no_ring = getenv("RESDEPTH_LAST_BWD_NORING") != nullptr;
    const int red = 16 * (10 * 4 * 16 * 1 + 1) * (int)sizeof(float);
    if (no_ring) {
      conv_last_bwd_kernel<1, 0><<<grid, 256, red, s>>>(u, dy, w, du, reinterpret_cast<__nv_bfloat16*>(du_b), scratch, B,
                                                        H, W, C, tiles_x, tiles_y, ntiles);
    } else {
      const int ring = 16 * LB_NST * 64 * (int)sizeof(float4);
      const int smem = ring > red ? ring : red;
      RD_CUDA(cudaFuncSetAttribute(conv_last_bwd_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      conv_last_bwd_ring_kernel<<<grid, 256, smem, s>>>(u, dy, w, du, reinterpret_cast<__nv_bfloat16*>(du_b), scratch, B,
                                                        H, W, C, tiles_x, tiles_y, ntiles);
    }
  } else {
    const int smem = 16 * (10 * 4 * 16 * 2 + 1) * (int)sizeof(float);
    RD_CUDA(cudaFuncSetAttribute(conv_last_bwd_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv_last_bwd_kernel<2, 0><<<grid, 256, smem, s>>>(u, dy, w, du, reinterpret_cast<__nv_bfloat16*>(du_b), scratch, B,
                                                       H, W, C, tiles_x, tiles_y, ntiles);
  }
  RD_LAUNCHED();
  RD_TRY(launch_sum_partials(scratch, grid, C * 9, PN, 1, dw, s));
  if (dbias) RD_TRY(launch_sum_partials(scratch + C * 9, grid, 1, PN, 1, dbias, s));
  if (du_channel_sum) RD_TRY(launch_sum_partials(scratch + C * 9 + 1, grid, C, PN, 1, du_channel_sum, s));
  return 0;
}

}  // namespace rd
