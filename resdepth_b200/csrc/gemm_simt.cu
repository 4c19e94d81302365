// GEMM-shaped layers on CUDA cores in full fp32 (RD_MATH_FP32, the "exact" mode; also the path for channel
// counts the tcgen05 kernels do not tile).  Two kernels cover all six contractions of the network:
//
//   rows   : C[M = B*Ho*Wo][N] = gather(src)[M][ntaps*C] * Bm[ntaps*C][N]
//            conv3x3 forward / dgrad (9 shifted taps), transposed-conv forward (1 tap, scatter epilogue with
//            bias + additive skip), transposed-conv dgrad (4 taps of the 2x2 stride-2 gather)
//            replaces nn.Conv2d / nn.ConvTranspose2d forward and their input gradients
//            (reference lib/UNet.py:4-5,21 ; autograd of lib/Trainer.py:179)
//   reduce : P[split][(tap,ca)][N] = sum over pixels gather(src)[p][(tap,ca)] * G[p][N]
//            conv3x3 wgrad and transposed-conv wgrad (split over pixels, reduced by the un-pack kernels)
#include "common.cuh"

namespace rd {

__device__ __forceinline__ float tf32_rn_(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

static constexpr int BM = 128, BK = 16;

template <int BN>
__global__ void __launch_bounds__(256)
gemm_rows_kernel(const float* __restrict__ src, const Gather g, const float* __restrict__ Bm, int Bsz, int N,
                 const Epilogue e) {
  constexpr int TN = BN / 16;                         // columns per thread (2, 4 or 8)
  constexpr int VW = TN >= 4 ? 4 : 2;                 // vector width of one column group
  constexpr int NG = TN / VW;                         // column groups per thread
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float red[16][BN][2];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN;
  const long long M = (long long)Bsz * g.Ho * g.Wo;
  const int m_tiles = (int)((M + BM - 1) / BM);
  const int K = g.ntaps * g.C;
  float cs1[TN], cs2[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) cs1[j] = cs2[j] = 0.f;

  for (int mt = blockIdx.y; mt < m_tiles; mt += gridDim.y) {
    const long long m0 = (long long)mt * BM;
    // the two A rows this thread loads: r = tid/4 and 64 + tid/4; k quad = tid%4
    int rb[2], rh[2], rw[2];
    bool rv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long m = m0 + (tid >> 2) + 64 * i;
      rv[i] = m < M;
      const long long mm = rv[i] ? m : 0;
      rw[i] = (int)(mm % g.Wo);
      rh[i] = (int)((mm / g.Wo) % g.Ho);
      rb[i] = (int)(mm / ((long long)g.Wo * g.Ho));
    }
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
      const int tap = k0 / g.C, c0 = k0 - tap * g.C;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0, 0, 0, 0);
        const int hs = g.ups * rh[i] + g.dh[tap], ws = g.ups * rw[i] + g.dw[tap];
        if (rv[i] && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws)
          v = *reinterpret_cast<const float4*>(src + (((size_t)rb[i] * g.Hs + hs) * g.Ws + ws) * g.C + c0 + (tid & 3) * 4);
        const int r = (tid >> 2) + 64 * i, kk = (tid & 3) * 4;
        As[kk + 0][r] = v.x; As[kk + 1][r] = v.y; As[kk + 2][r] = v.z; As[kk + 3][r] = v.w;
      }
      for (int idx = tid; idx < (BK * BN) / 4; idx += 256) {
        const int kk = idx / (BN / 4), nq = idx % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[kk][nq * 4]) =
            *reinterpret_cast<const float4*>(Bm + (size_t)(k0 + kk) * N + n0 + nq * 4);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[8], b[TN];
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
#pragma unroll
        for (int gq = 0; gq < NG; ++gq)
#pragma unroll
          for (int v = 0; v < VW; ++v) b[gq * VW + v] = Bs[kk][gq * (BN / 2) + tx * VW + v];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= M) continue;
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
        const int n = n0 + gq * (BN / 2) + tx * VW;
        float v[VW];
#pragma unroll
        for (int q = 0; q < VW; ++q) v[q] = acc[i][gq * VW + q];
        float* dst;
        if (e.mode == EPI_CONVT) {
          const int Co = N >> 2;
          const int ab = n / Co, co = n - ab * Co;
          const int wo = (int)(m % g.Wo);
          const int ho = (int)((m / g.Wo) % g.Ho);
          const int b = (int)(m / ((long long)g.Wo * g.Ho));
          const size_t o = ((((size_t)b * 2 * g.Ho + 2 * ho + (ab >> 1)) * 2 * g.Wo) + 2 * wo + (ab & 1)) * Co + co;
#pragma unroll
          for (int q = 0; q < VW; ++q) {
            v[q] += e.bias[co + q];
            if (e.skip) v[q] += e.skip[o + q];
          }
          dst = e.out + o;
        } else {
          dst = e.out + (size_t)m * N + n;
        }
        if (e.round_tf32) {
#pragma unroll
          for (int q = 0; q < VW; ++q) v[q] = tf32_rn_(v[q]);
        }
        if (VW == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        else *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
        if (e.mode == EPI_STATS) {
#pragma unroll
          for (int q = 0; q < VW; ++q) {
            cs1[gq * VW + q] += v[q];
            cs2[gq * VW + q] = fmaf(v[q], v[q], cs2[gq * VW + q]);
          }
        }
      }
    }
  }
  if (e.mode == EPI_STATS) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = (j / VW) * (BN / 2) + tx * VW + (j % VW);
      red[ty][col][0] = cs1[j];
      red[ty][col][1] = cs2[j];
    }
    __syncthreads();
    for (int i = tid; i < BN * 2; i += 256) {
      const int col = i >> 1, w = i & 1;
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) a += red[t][col][w];
      e.partials[((size_t)blockIdx.y * N + n0 + col) * 2 + w] = a;
    }
  }
}

int launch_gemm_rows_simt(const float* src, const Gather& g, const float* Bm, int B, int N, const Epilogue& e,
                          int* n_partials, cudaStream_t s) {
  if (g.C % BK) return fail("gemm_rows: channel count %d is not a multiple of %d", g.C, BK);
  if (N % 32) return fail("gemm_rows: N=%d is not a multiple of 32", N);
  if (e.mode == EPI_CONVT && (N / 4) % 4) return fail("gemm_rows: convT channel count %d not a multiple of 4", N / 4);
  const long long M = (long long)B * g.Ho * g.Wo;
  const int m_tiles = cdiv(M, BM);
  const int BN = (N % 128 == 0) ? 128 : (N % 64 == 0 ? 64 : 32);
  const int n_tiles = N / BN;
  int gy = cdiv(148 * 2, n_tiles);
  if (gy > m_tiles) gy = m_tiles;
  if (gy > 1024) gy = 1024;
  dim3 grid(n_tiles, gy);
  if (BN == 128) gemm_rows_kernel<128><<<grid, 256, 0, s>>>(src, g, Bm, B, N, e);
  else if (BN == 64) gemm_rows_kernel<64><<<grid, 256, 0, s>>>(src, g, Bm, B, N, e);
  else gemm_rows_kernel<32><<<grid, 256, 0, s>>>(src, g, Bm, B, N, e);
  RD_LAUNCHED();
  if (n_partials) *n_partials = gy;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// reduce GEMM (weight gradients)
// ---------------------------------------------------------------------------------------------
static constexpr int RK = 8;                           // pixels per k-step

template <int BN>
__global__ void __launch_bounds__(256)
gemm_reduce_kernel(const float* __restrict__ src, const Gather g, const float* __restrict__ G, int Bsz, int N,
                   float* __restrict__ part, long long pix_per_split) {
  constexpr int TN = BN / 16;
  constexpr int VW = TN >= 4 ? 4 : 2;
  constexpr int NG = TN / VW;
  __shared__ __align__(16) float As[2][RK][BM];
  __shared__ __align__(16) float Bs[2][RK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int Mrows = g.ntaps * g.C;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const long long npix = (long long)Bsz * g.Ho * g.Wo;
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > npix) p_end = npix;

  // A load assignment: pixel slot ak = tid/32, row quad aq = tid%32 -> rows m0 + aq*4 .. +3 (same tap: C%4==0)
  const int ak = tid >> 5, aq = tid & 31;
  const int arow = m0 + aq * 4;
  const bool arow_ok = arow < Mrows;
  const int atap = arow_ok ? arow / g.C : 0;
  const int ac = arow_ok ? arow - atap * g.C : 0;
  const int adh = g.dh[atap], adw = g.dw[atap];
  // B load assignment
  const int bk = tid / (BN / 4), bq = tid % (BN / 4);
  const bool b_active = tid < RK * (BN / 4);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto load_a = [&](long long p) -> float4 {
    float4 v = make_float4(0, 0, 0, 0);
    if (arow_ok && p < p_end) {
      const int w = (int)(p % g.Wo);
      const int h = (int)((p / g.Wo) % g.Ho);
      const int b = (int)(p / ((long long)g.Wo * g.Ho));
      const int hs = g.ups * h + adh, ws = g.ups * w + adw;
      if (hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws)
        v = *reinterpret_cast<const float4*>(src + (((size_t)b * g.Hs + hs) * g.Ws + ws) * g.C + ac);
    }
    return v;
  };
  auto load_b = [&](long long p) -> float4 {
    float4 v = make_float4(0, 0, 0, 0);
    if (b_active && p < p_end) v = *reinterpret_cast<const float4*>(G + (size_t)p * N + n0 + bq * 4);
    return v;
  };

  const long long nsteps = (p_end - p_begin + RK - 1) / RK;
  float4 ra = load_a(p_begin + ak), rb = load_b(p_begin + bk);
  for (long long st = 0; st < nsteps; ++st) {
    const int buf = (int)(st & 1);
    *reinterpret_cast<float4*>(&As[buf][ak][aq * 4]) = ra;
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][bk][bq * 4]) = rb;
    __syncthreads();
    if (st + 1 < nsteps) {
      const long long pn = p_begin + (st + 1) * RK;
      ra = load_a(pn + ak);
      rb = load_b(pn + bk);
    }
#pragma unroll
    for (int kk = 0; kk < RK; ++kk) {
      float a[8], b[TN];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
#pragma unroll
      for (int gq = 0; gq < NG; ++gq)
#pragma unroll
        for (int v = 0; v < VW; ++v) b[gq * VW + v] = Bs[buf][kk][gq * (BN / 2) + tx * VW + v];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // double buffering: the next iteration writes the other buffer; one barrier per step is enough because a
    // thread can only be one step ahead (it needs the barrier of step st+1 before touching buffer `buf` again)
  }
  float* out = part + (size_t)blockIdx.z * Mrows * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= Mrows) continue;
#pragma unroll
    for (int gq = 0; gq < NG; ++gq) {
      const int n = n0 + gq * (BN / 2) + tx * VW;
      float* dst = out + (size_t)m * N + n;
      if (VW == 4)
        *reinterpret_cast<float4*>(dst) =
            make_float4(acc[i][gq * 4 + 0], acc[i][gq * 4 + 1], acc[i][gq * 4 + 2], acc[i][gq * 4 + 3]);
      else
        *reinterpret_cast<float2*>(dst) = make_float2(acc[i][gq * 2 + 0], acc[i][gq * 2 + 1]);
    }
  }
}

int launch_gemm_reduce_simt(const float* src, const Gather& g, const float* G, int B, int N, float* part,
                            size_t part_floats, int* splits_out, cudaStream_t s) {
  if (g.C % 4) return fail("gemm_reduce: channel count %d is not a multiple of 4", g.C);
  if (N % 32) return fail("gemm_reduce: N=%d is not a multiple of 32", N);
  const int Mrows = g.ntaps * g.C;
  const long long npix = (long long)B * g.Ho * g.Wo;
  const int BN = (N % 128 == 0) ? 128 : (N % 64 == 0 ? 64 : 32);
  const int tiles = cdiv(Mrows, BM) * (N / BN);
  int S = cdiv(148 * 2, tiles);
  const long long max_s_pix = (npix + 255) / 256;       // at least 256 pixels per split
  if (S > max_s_pix) S = (int)max_s_pix;
  const size_t per = (size_t)Mrows * N;
  while (S > 1 && (size_t)S * per > part_floats) --S;
  if ((size_t)S * per > part_floats) return fail("gemm_reduce: scratch too small (%zu floats needed)", per);
  if (S < 1) S = 1;
  long long pps = (npix + S - 1) / S;
  pps = (pps + RK - 1) / RK * RK;
  S = (int)((npix + pps - 1) / pps);
  dim3 grid(N / BN, cdiv(Mrows, BM), S);
  if (BN == 128) gemm_reduce_kernel<128><<<grid, 256, 0, s>>>(src, g, G, B, N, part, pps);
  else if (BN == 64) gemm_reduce_kernel<64><<<grid, 256, 0, s>>>(src, g, G, B, N, part, pps);
  else gemm_reduce_kernel<32><<<grid, 256, 0, s>>>(src, g, G, B, N, part, pps);
  RD_LAUNCHED();
  *splits_out = S;
  return 0;
}

}  // namespace rd
