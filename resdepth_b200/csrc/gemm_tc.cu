// GEMM-shaped layers on the 5th-generation tensor cores (RD_MATH_TF32).
//
// rows kernel (conv3x3 forward / dgrad, transposed-conv forward / dgrad; reference lib/UNet.py:4-5,21 and their
// autograd): implicit GEMM  D[128 pixels][BN] = sum_taps sum_c A_tap[pixel][c] * W[n][(tap,c)]
//   * A tiles: TMA 4-D boxes (32 channels x tw x th x tb pixels) of the NHWC activation tensor, one box per
//     (tap, 32-channel chunk); negative / overflowing coordinates are zero-filled by the TMA unit, which IS the
//     convolution padding.  128-byte swizzle, K-major (32 fp32 of K per 128-byte row).
//   * B tiles: TMA 2-D boxes (32 k x BN) of the packed weights [N][K] (K-major).
//   * tcgen05.mma kind::tf32, M = 128, N = BN, K = 8 per instruction, fp32 accumulators in TMEM, two
//     accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * warp roles (384 threads): warp 0 = activation-TMA producer, warp 1 = TMEM allocator + MMA issuer, warp 2 =
//     weight-TMA producer of the halo variant, warps 4..11 = epilogue (tcgen05.ld 32 lanes x 32 columns -> registers ->
//     fused epilogue -> transposition through swizzled shared memory -> 128-byte row stores).  Register budget by
//     role (setmaxnreg): the first warpgroup releases down to 72 registers per thread, the two epilogue warpgroups
//     take 216 -- the epilogue is the instruction- and latency-bound part of every layer with short K.
//   * variants: HALO (3x3 layers with N <= 128: one 18 x 18 halo patch per channel chunk serves all nine taps and two
//     sub-tiles; weight tiles of three taps per slot), RING (HBM-bound up-convs: skip rows by per-warp cp.async rings),
//     BF16 (backward GEMMs, kind::f16).  The file also holds the reduce kernels (weight gradients, incl. the wide-N
//     formulation) and the first-conv kernel (software im2col producer, 3xTF32).
//   * persistent CTAs, static tile schedule with a fixed N tile per CTA so BatchNorm column sums accumulate in
//     registers across all of a CTA's tiles (one partial row per CTA-warp instead of one per tile); they are taken in
//     the transposed domain of the store path (a lane's four channels over its rows), with one cross-lane step per kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace rd {

using namespace tc;

static constexpr int TC_THREADS = 192;          // reduce kernel: TMA warp, MMA warp, 4 epilogue warps
static constexpr int ROWS_THREADS = 384;        // rows kernel: A-TMA warp, MMA warp, B-TMA warp (halo variant), spare, 8 epilogue warps

__device__ __forceinline__ float tf32_round(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// HALO variant (3x3 taps, image at least 16x16): one CTA tile is 16 x 16 pixels = TWO 128-row MMA sub-tiles
// (16 rows x 8 pixels each) with separate TMEM accumulators.
//   * The 18x18 halo patch of one 32-channel chunk is loaded ONCE (40.5 KB) and reused by all nine taps and both
//     sub-tiles: a (tap, sub-tile) is just a different start row of the UMMA descriptor inside the patch
//     (start address + (dh*18 + dw + 8*sub)*128 bytes, 8-row groups 18 rows = 2304 bytes apart).  The 128-byte
//     swizzle tolerates the unaligned start rows because it is a function of the shared-memory address.
//   * Every weight tile (32 k x BN) is used by both sub-tiles, halving the weight traffic per MMA -- with the
//     plain kernel the operand traffic is 96-187 B per MMA cycle per SM, here 40-50.
static constexpr int HALO_W = 18, HALO_H = 18, HALO_SUB = 2;
static constexpr int HALO_BYTES = HALO_W * HALO_H * 128;             // 41472
static constexpr int HALO_SLOT = 41 * 1024;                           // 1024-byte aligned slot

// RING: transposed-conv variant for HBM-bound up-convs (K = C <= 128): a 2-stage operand ring frees 96 KB of shared
// memory for per-warp cp.async rings that hold the additive-skip rows of the next chunks (see convt_ring_epilogue)
__host__ __device__ constexpr int ring_slots(int BN) { return BN == 256 ? 3 : 4; }   // per-warp 4 KB slots that fit
template <int BN, bool HALO, bool RING = false>
struct RowsCfg {
  static constexpr int A_BYTES = HALO ? HALO_SLOT : 128 * 128;       // plain: 128 rows x 32 fp32
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;              // plain variant: one ring of (A, B) stages
  static constexpr int STAGES = RING ? 2 : (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static_assert(!RING || (!HALO && (BN == 256 || BN == 128)), "skip-ring variant: plain 256/128-column tiles only");
  static_assert(!HALO || BN <= 128, "halo variant: two sub-tiles x two accumulator stages x BN columns <= 512");
  static constexpr int A_SLOTS = 2;                                  // halo variant: separate rings
  // halo variant: one weight slot = the tiles of 3 taps (one 3-D TMA, one barrier round trip per 24 MMAs: with N = 64 an
  // MMA lasts 32 cycles and the per-tap wait / commit of the issuing warp was what the tensor pipe waited for)
  static constexpr int B_TAPS = 3;
  static constexpr int B_SLOT_BYTES = B_TAPS * B_BYTES;
  static constexpr int B_SLOTS = (BN == 128) ? 2 : 4;
  static constexpr int SUB = HALO ? HALO_SUB : 1;                    // 128-row sub-tiles per CTA tile
  static constexpr int ACC_COLS = SUB * BN;                          // TMEM columns of one accumulator stage
  static constexpr int DATA_BYTES = HALO ? A_SLOTS * HALO_SLOT + B_SLOTS * B_SLOT_BYTES : STAGES * STAGE_BYTES;
  static constexpr int NBAR_A = HALO ? A_SLOTS : STAGES;
  static constexpr int NBAR_B = HALO ? B_SLOTS : 0;
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr int RING_OFF = DATA_BYTES + 512 + 8 * 4096;
  static constexpr int SMEM_BYTES = DATA_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/ + 8 * 4096 /*epilogue staging*/ +
                                    (RING ? 8 * ring_slots(BN) * 4096 : 0) /*skip rings*/;
};

// Epilogue store of one warp's 32 rows x 32 columns.  After tcgen05.ld a lane owns one ROW (32 consecutive floats of
// one pixel); storing that directly makes every warp-level STG.128 touch 32 different 128-byte lines (32 L1 wavefronts
// for 512 bytes).  The tile is therefore transposed through 4 KB of XOR-swizzled shared memory so that 8 lanes
// cover one row: each warp-level access then touches 4 rows x 128 contiguous bytes (4 wavefronts).  `off` is the
// element offset of the lane's row.  (The transposed-conv epilogues, which also add the skip tensor, have their own
// versions of this: convt_epilogue / convt_ring_epilogue.)
// Row offsets of a (sub-)tile in the transposed domain (lane -> rows i*4 + lane/8): shuffled ONCE per 128-row tile, not
// per 32-column chunk.  32-bit element offsets: the launcher rejects outputs of 2^32 elements or more.
__device__ __forceinline__ void rows_to_transposed(int lane, unsigned row_off, bool ok, unsigned (&R)[8], unsigned& okm) {
  okm = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    R[i] = __shfl_sync(0xffffffffu, row_off, r);
    okm |= (unsigned)__shfl_sync(0xffffffffu, (int)ok, r) << i;
  }
}
// `col` = element offset of the chunk inside a row + this lane's 4-channel group (n + 4 * (lane % 8)).
template <bool SUMS = false>
__device__ __forceinline__ void warp_store_rows(float* stg, int lane, const float (&v)[32], float* __restrict__ out,
                                                const unsigned (&R)[8], unsigned okm, unsigned col, int rnd,
                                                __nv_bfloat16* __restrict__ outb = nullptr, float4* s1 = nullptr,
                                                float4* s2 = nullptr) {
  const uint32_t stg_s = smem_u32(stg);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    sts128(stg_s + (lane * 32 + ((c4 ^ (lane & 7)) << 2)) * 4,
             make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]));
  __syncwarp();
  const int c4 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    float4 val = lds128(stg_s + (r * 32 + ((c4 ^ (r & 7)) << 2)) * 4);
    if (SUMS && ((okm >> i) & 1)) {
      // column sums (BatchNorm statistics / bias gradient) of the valid rows, per lane: this lane's four channels over its
      // eight rows -- the cross-lane part (four lane groups) is done ONCE per kernel, not with a 31-shuffle butterfly per chunk
      s1->x += val.x; s1->y += val.y; s1->z += val.z; s1->w += val.w;
      s2->x = fmaf(val.x, val.x, s2->x); s2->y = fmaf(val.y, val.y, s2->y);
      s2->z = fmaf(val.z, val.z, s2->z); s2->w = fmaf(val.w, val.w, s2->w);
    }
    if (rnd) { val.x = tf32_round(val.x); val.y = tf32_round(val.y); val.z = tf32_round(val.z); val.w = tf32_round(val.w); }
    if ((okm >> i) & 1) {
      const size_t o = (size_t)(R[i] + col);
      if (out) *reinterpret_cast<float4*>(out + o) = val;
      if (outb) {
        __nv_bfloat162 lo2 = __floats2bfloat162_rn(val.x, val.y), hi2 = __floats2bfloat162_rn(val.z, val.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo2);
        pk.y = *reinterpret_cast<uint32_t*>(&hi2);
        *reinterpret_cast<uint2*>(outb + o) = pk;
      }
    }
  }
  __syncwarp();
}

// Inference epilogue of one warp's 32 rows x 32 columns: BatchNorm(running statistics) + activation (+ 2x2 max-pool), all
// in the TRANSPOSED domain (lane -> rows i*4 + lane/8, columns 4*(lane%8)..+3): the per-channel constants are two float4
// per lane (requested before the accumulator load) instead of sixteen row-domain loads behind it, one staging round trip
// serves both outputs, and the pooling partners of a row are a register (row ^ tw: same lane for tw >= 4) and lane ^ 8
// (row ^ 1).  ncu (round 2): the row-domain version held the 64-channel inference convs at 41 % tensor pipe.
__device__ __forceinline__ void warp_bnact_store_rows(float* stg, int lane, const float (&v)[32], float4 sc, float4 sh,
                                                      float slope, float* __restrict__ out, long long off, bool ok, int rnd,
                                                      float* __restrict__ pool_out, long long poff, bool pok, int rnd_pool,
                                                      int tw) {
  const uint32_t stg_s = smem_u32(stg);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    sts128(stg_s + (lane * 32 + ((c4 ^ (lane & 7)) << 2)) * 4, make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]));
  __syncwarp();
  const int c4 = lane & 7;
  const unsigned off_lo = (unsigned)(off & 0xffffffffu), off_hi = (unsigned)((unsigned long long)off >> 32);
  const unsigned pof_lo = (unsigned)(poff & 0xffffffffu), pof_hi = (unsigned)((unsigned long long)poff >> 32);
  float4 val[8];
  unsigned okm = 0, pokm = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    float4 a = lds128(stg_s + (r * 32 + ((c4 ^ (r & 7)) << 2)) * 4);
    a.x = fmaf(a.x, sc.x, sh.x); a.y = fmaf(a.y, sc.y, sh.y); a.z = fmaf(a.z, sc.z, sh.z); a.w = fmaf(a.w, sc.w, sh.w);
    a.x = a.x > 0.f ? a.x : a.x * slope; a.y = a.y > 0.f ? a.y : a.y * slope;
    a.z = a.z > 0.f ? a.z : a.z * slope; a.w = a.w > 0.f ? a.w : a.w * slope;
    val[i] = a;
    const unsigned lo = __shfl_sync(0xffffffffu, off_lo, r), hi = __shfl_sync(0xffffffffu, off_hi, r);
    const bool okr = __shfl_sync(0xffffffffu, (int)ok, r) != 0;
    okm |= (unsigned)okr << i;
    if (okr) {
      float4 w4 = a;
      if (rnd) { w4.x = tf32_round(w4.x); w4.y = tf32_round(w4.y); w4.z = tf32_round(w4.z); w4.w = tf32_round(w4.w); }
      *reinterpret_cast<float4*>(out + (long long)(((unsigned long long)hi << 32) | lo) + c4 * 4) = w4;
    }
  }
  if (pool_out) {
    // vertical partner: row ^ tw
    float4 m[8];
    if (tw >= 4) {
      const int k = tw >> 2;                              // 1, 2 or 4: register index bit
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 o = k == 1 ? val[i ^ 1] : (k == 2 ? val[i ^ 2] : val[i ^ 4]);
        m[i] = make_float4(fmaxf(val[i].x, o.x), fmaxf(val[i].y, o.y), fmaxf(val[i].z, o.z), fmaxf(val[i].w, o.w));
      }
    } else {                                              // tw == 2: row bit 1 = lane bit 4
#pragma unroll
      for (int i = 0; i < 8; ++i)
        m[i] = make_float4(fmaxf(val[i].x, __shfl_xor_sync(0xffffffffu, val[i].x, 16)),
                           fmaxf(val[i].y, __shfl_xor_sync(0xffffffffu, val[i].y, 16)),
                           fmaxf(val[i].z, __shfl_xor_sync(0xffffffffu, val[i].z, 16)),
                           fmaxf(val[i].w, __shfl_xor_sync(0xffffffffu, val[i].w, 16)));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      float4 p4 = make_float4(fmaxf(m[i].x, __shfl_xor_sync(0xffffffffu, m[i].x, 8)),      // horizontal partner: row ^ 1
                              fmaxf(m[i].y, __shfl_xor_sync(0xffffffffu, m[i].y, 8)),
                              fmaxf(m[i].z, __shfl_xor_sync(0xffffffffu, m[i].z, 8)),
                              fmaxf(m[i].w, __shfl_xor_sync(0xffffffffu, m[i].w, 8)));
      const unsigned lo = __shfl_sync(0xffffffffu, pof_lo, r), hi = __shfl_sync(0xffffffffu, pof_hi, r);
      const bool pk = __shfl_sync(0xffffffffu, (int)pok, r) != 0;
      pokm |= (unsigned)pk << i;
      if (pk) {
        if (rnd_pool) { p4.x = tf32_round(p4.x); p4.y = tf32_round(p4.y); p4.z = tf32_round(p4.z); p4.w = tf32_round(p4.w); }
        *reinterpret_cast<float4*>(pool_out + (long long)(((unsigned long long)hi << 32) | lo) + c4 * 4) = p4;
      }
    }
  }
  (void)okm; (void)pokm;
  __syncwarp();
}

// Transposed-conv epilogue of one warp for one 128-row (sub-)tile: + bias, + additive skip (optionally BatchNorm +
// activation of the raw encoder output on the fly), scatter to the 2x2 output sites, optional TF32 rounding and
// bf16 copy.  The layer is HBM-bound and the skip read sits on the critical path of every store, so the skip rows
// of the NEXT 32-column chunk are requested before the current chunk is stored (two chunks = 8 KB per warp in
// flight).  Output offsets split into a row part (shuffled once per tile into the transposed domain: lane -> rows
// i*4 + lane/8, columns 4*(lane%8)..+3) and a warp-uniform chunk part.
template <int BN>
__device__ __forceinline__ void convt_epilogue(const TcRowsParams& P, float* stg, int lane, int half, uint32_t t_row,
                                               int nt, int b, int h, int w, bool valid) {
  constexpr int NCH = BN / 32, NCH2 = (NCH + 1) / 2;
  const uint32_t stg_s = smem_u32(stg);
  const int Co = P.N >> 2;
  const int c4 = lane & 7;
  // 32-bit element offsets: the launcher rejects outputs of 2^32 elements or more
  const unsigned rrow = (unsigned)((((size_t)b * 2 * P.Ho + 2 * h) * 2 * P.Wo + 2 * w) * Co);
  unsigned R[8];
  unsigned okm = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    R[i] = __shfl_sync(0xffffffffu, rrow, r) + c4 * 4;
    okm |= (unsigned)__shfl_sync(0xffffffffu, (int)valid, r) << i;
  }
  auto chunk_off = [&](int ch, int& co) -> unsigned {
    const int n = nt * BN + ch * 32;
    const int ab = n / Co;
    co = n - ab * Co;
    return (unsigned)(((ab >> 1) * 2 * P.Wo + (ab & 1)) * Co + co);
  };
  const bool has_skip = P.skip != nullptr;
  float4 a_cur[8], a_nxt[8];
  auto prefetch = [&](int ch, float4 (&a)[8]) {
    int co;
    const unsigned S = chunk_off(ch, co);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      a[i] = (okm >> i) & 1 ? __ldg(reinterpret_cast<const float4*>(P.skip + (size_t)(R[i] + S))) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
#pragma unroll
  for (int i = 0; i < 8; ++i) a_cur[i] = a_nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has_skip && half < NCH) prefetch(half, a_cur);
  __nv_bfloat16* outb = reinterpret_cast<__nv_bfloat16*>(P.out_b);
#pragma unroll
  for (int ci = 0; ci < NCH2; ++ci) {
    const int ch = 2 * ci + half;
    if (ch >= NCH) break;
    float v[32];
    tmem_ld32(t_row + ch * 32, v);
    int co;
    const unsigned S = chunk_off(ch, co);
    // bias / BatchNorm constants of this lane's four channels in the transposed domain: requested before the staging
    // stores, consumed after them (they used to be eight row-domain loads in front of the stores)
    const float4 bi = __ldg(reinterpret_cast<const float4*>(P.bias + co) + c4);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    float sl = 0.f;
    if (has_skip && P.skip_scale) {
      sc = __ldg(reinterpret_cast<const float4*>(P.skip_scale + co) + c4);
      sh = __ldg(reinterpret_cast<const float4*>(P.skip_shift + co) + c4);
      sl = __ldg(P.skip_slope);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(stg_s + (lane * 32 + ((j ^ (lane & 7)) << 2)) * 4, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    __syncwarp();
    if (has_skip && ch + 2 < NCH) prefetch(ch + 2, a_nxt);     // v[] is dead here: a_nxt takes its registers
    if (has_skip && P.skip_scale) {    // the skip tensor is a raw conv output: BatchNorm + activation of its layer
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y0 = fmaf(a_cur[i].x, sc.x, sh.x), y1 = fmaf(a_cur[i].y, sc.y, sh.y);
        const float y2 = fmaf(a_cur[i].z, sc.z, sh.z), y3 = fmaf(a_cur[i].w, sc.w, sh.w);
        a_cur[i].x = y0 > 0.f ? y0 : y0 * sl; a_cur[i].y = y1 > 0.f ? y1 : y1 * sl;
        a_cur[i].z = y2 > 0.f ? y2 : y2 * sl; a_cur[i].w = y3 > 0.f ? y3 : y3 * sl;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      float4 val = lds128(stg_s + (r * 32 + ((c4 ^ (r & 7)) << 2)) * 4);
      val.x = (val.x + bi.x) + a_cur[i].x; val.y = (val.y + bi.y) + a_cur[i].y;     // (acc + bias) + skip
      val.z = (val.z + bi.z) + a_cur[i].z; val.w = (val.w + bi.w) + a_cur[i].w;
      if (P.round_tf32) {
        val.x = tf32_round(val.x); val.y = tf32_round(val.y); val.z = tf32_round(val.z); val.w = tf32_round(val.w);
      }
      if ((okm >> i) & 1) {
        const size_t o = (size_t)(R[i] + S);
        *reinterpret_cast<float4*>(P.out + o) = val;
        if (outb) {
          __nv_bfloat162 lo2 = __floats2bfloat162_rn(val.x, val.y), hi2 = __floats2bfloat162_rn(val.z, val.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo2);
          pk.y = *reinterpret_cast<uint32_t*>(&hi2);
          *reinterpret_cast<uint2*>(outb + o) = pk;
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) a_cur[i] = a_nxt[i];
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const uint32_t n = pred ? 16u : 0u;                  // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Whole epilogue loop of one warp for the skip-ring transposed-conv variant.  The layer is HBM-bound and every
// output store waits for its skip row, so the skip rows are fetched with cp.async RING_SLOTS-1 chunks (8 KB per
// warp, 64 KB per SM) ahead of their use -- across tile boundaries, without holding registers.  The copy runs
// in the transposed domain (lane -> rows i*4 + lane/8, 16-byte column group lane%8): each lane later reads back
// exactly the 16 bytes it copied, so no cross-lane synchronisation is needed.
template <int BN>
__device__ __forceinline__ void convt_ring_epilogue(const TcRowsParams& P, uint8_t* ring, float* stg, int lane, int q,
                                                    int half, uint32_t tmem_base, uint64_t* tfull_bar,
                                                    uint64_t* tempty_bar, int num_tiles, int n_tiles) {
  constexpr int NCH2 = BN / 64;                         // chunks of this warp per tile (even / odd 32-column chunks)
  constexpr int RING_SLOTS = ring_slots(BN);
  const uint32_t stg_s = smem_u32(stg);
  const int Co = P.N >> 2;
  const int c4 = lane & 7;
  const int row = q * 32 + lane;
  const int iw = row % P.tw, ih = (row / P.tw) % P.th, ib = row / (P.tw * P.th);
  const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total = my_tiles * NCH2;
  auto tile_row = [&](int it, int& nt, bool& valid) -> unsigned {      // element offset of this lane's row, tile `it`
    const int tile = blockIdx.x + it * gridDim.x;
    nt = tile % n_tiles;
    int mt = tile / n_tiles;
    const int tw_i = mt % P.tiles_w; mt /= P.tiles_w;
    const int th_i = mt % P.tiles_h;
    const int tb_i = mt / P.tiles_h;
    const int w = tw_i * P.tw + iw, h = th_i * P.th + ih, b = tb_i * P.tb + ib;
    valid = (w < P.Wo) && (h < P.Ho) && (b < P.Bo);
    return (unsigned)((((size_t)b * 2 * P.Ho + 2 * h) * 2 * P.Wo + 2 * w) * Co);
  };
  auto chunk_off = [&](int nt, int ch, int& co) -> unsigned {
    const int n = nt * BN + ch * 32;
    const int ab = n / Co;
    co = n - ab * Co;
    return (unsigned)(((ab >> 1) * 2 * P.Wo + (ab & 1)) * Co + co);
  };
  auto issue = [&](int m) {                             // request the skip rows of this warp's m-th chunk
    if (m < total) {
      int nt, co;
      bool valid;
      const unsigned rrow = tile_row(m / NCH2, nt, valid);
      const unsigned S = chunk_off(nt, 2 * (m % NCH2) + half, co);
      uint8_t* slot = ring + (m % RING_SLOTS) * 4096;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        const unsigned o = __shfl_sync(0xffffffffu, rrow, r) + c4 * 4 + S;
        const bool ok = __shfl_sync(0xffffffffu, (int)valid, r) != 0;
        cp_async16(slot + (i * 32 + lane) * 16, P.skip + (ok ? (size_t)o : 0), ok);
      }
    }
    cp_async_commit();                                  // empty groups keep the group count uniform
  };
#pragma unroll
  for (int m = 0; m < RING_SLOTS - 1; ++m) issue(m);
  __nv_bfloat16* outb = reinterpret_cast<__nv_bfloat16*>(P.out_b);
  const float sl = P.skip_scale ? __ldg(P.skip_slope) : 0.f;
  // per-channel constants of a chunk (bias, BatchNorm scale / shift of the skip's layer) for THIS lane's four channels of
  // the transposed domain, fetched one chunk ahead: ncu (round 2) had a third of the epilogue's warp samples on the first
  // use of these L1-resident loads when they were issued inside the chunk that consumes them
  auto chunk_consts = [&](int m, float4& bi, float4& sc, float4& sh) {
    bi = make_float4(0.f, 0.f, 0.f, 0.f); sc = make_float4(1.f, 1.f, 1.f, 1.f); sh = bi;
    if (m < total) {
      const int tile = blockIdx.x + (m / NCH2) * gridDim.x;
      int co;
      (void)chunk_off(tile % n_tiles, 2 * (m % NCH2) + half, co);
      bi = __ldg(reinterpret_cast<const float4*>(P.bias + co) + c4);
      if (P.skip_scale) {
        sc = __ldg(reinterpret_cast<const float4*>(P.skip_scale + co) + c4);
        sh = __ldg(reinterpret_cast<const float4*>(P.skip_shift + co) + c4);
      }
    }
  };
  float4 bi, sc, sh, bi_n, sc_n, sh_n;
  chunk_consts(0, bi, sc, sh);
  unsigned R[8];
  unsigned okm = 0;
  int nt = 0, acc = 0;
  for (int m = 0; m < total; ++m) {
    const int it = m / NCH2, ci = m - it * NCH2;
    if (ci == 0) {
      acc = it & 1;
      bool valid;
      const unsigned rrow = tile_row(it, nt, valid);
      okm = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        R[i] = __shfl_sync(0xffffffffu, rrow, r) + c4 * 4;
        okm |= (unsigned)__shfl_sync(0xffffffffu, (int)valid, r) << i;
      }
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
    }
    issue(m + RING_SLOTS - 1);
    const int ch = 2 * ci + half;
    int co;
    const unsigned S = chunk_off(nt, ch, co);
    chunk_consts(m + 1, bi_n, sc_n, sh_n);
    {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + ch * 32, v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(stg_s + (lane * 32 + ((j ^ (lane & 7)) << 2)) * 4, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    }
    __syncwarp();
    cp_async_wait<RING_SLOTS - 1>();                    // this chunk's skip rows have landed
    const uint32_t slot_s = smem_u32(ring) + (m % RING_SLOTS) * 4096;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      float4 a = lds128(slot_s + (i * 32 + lane) * 16);
      if (P.skip_scale) {               // the skip tensor is a raw conv output: BatchNorm + activation of its layer
        const float y0 = fmaf(a.x, sc.x, sh.x), y1 = fmaf(a.y, sc.y, sh.y);
        const float y2 = fmaf(a.z, sc.z, sh.z), y3 = fmaf(a.w, sc.w, sh.w);
        a.x = y0 > 0.f ? y0 : y0 * sl; a.y = y1 > 0.f ? y1 : y1 * sl;
        a.z = y2 > 0.f ? y2 : y2 * sl; a.w = y3 > 0.f ? y3 : y3 * sl;
      }
      float4 val = lds128(stg_s + (r * 32 + ((c4 ^ (r & 7)) << 2)) * 4);
      val.x = (val.x + bi.x) + a.x; val.y = (val.y + bi.y) + a.y;           // (acc + bias) + skip
      val.z = (val.z + bi.z) + a.z; val.w = (val.w + bi.w) + a.w;
      if (P.round_tf32) {
        val.x = tf32_round(val.x); val.y = tf32_round(val.y); val.z = tf32_round(val.z); val.w = tf32_round(val.w);
      }
      if ((okm >> i) & 1) {
        const size_t o = (size_t)(R[i] + S);
        *reinterpret_cast<float4*>(P.out + o) = val;
        if (outb) {
          __nv_bfloat162 lo2 = __floats2bfloat162_rn(val.x, val.y), hi2 = __floats2bfloat162_rn(val.z, val.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo2);
          pk.y = *reinterpret_cast<uint32_t*>(&hi2);
          *reinterpret_cast<uint2*>(outb + o) = pk;
        }
      }
    }
    __syncwarp();
    if (ci == NCH2 - 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    bi = bi_n; sc = sc_n; sh = sh_n;
  }
  cp_async_wait<0>();
}

#define RD_REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 72;")
#define RD_REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 216;")
template <int BN, bool HALO, bool BF16, bool RING = false>
__global__ void __launch_bounds__(ROWS_THREADS, 1)
gemm_rows_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const TcRowsParams P) {
  using Cfg = RowsCfg<BN, HALO, RING>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::DATA_BYTES);     // plain: stage ring; halo: A ring
  uint64_t* empty_bar = full_bar + Cfg::NBAR_A;
  uint64_t* bfull_bar = empty_bar + Cfg::NBAR_A;                                // halo: B ring
  uint64_t* bempty_bar = bfull_bar + Cfg::NBAR_B;
  uint64_t* tfull_bar = bempty_bar + Cfg::NBAR_B;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int i = 0; i < Cfg::NBAR_A; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < Cfg::NBAR_B; ++i) { mbar_init(&bfull_bar[i], 1); mbar_init(&bempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Warp-specialised register budget: the producer / MMA-issuer warpgroup (warps 0-3) gives registers back, the two
  // epilogue warpgroups (warps 4-11) take them: 128 x 72 + 256 x 216 = 64 512 <= 65 536.  The epilogue is the
  // instruction- and latency-bound part of every HBM-bound layer (ncu round 2) and was held at 168 registers with spills.
  // (the instructions sit at the top of each role's branch: ptxas bounds the registers of the code a setmaxnreg dominates)

  const int n_tiles = P.N / BN;
  const int m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  const int num_tiles = m_tiles * n_tiles;
  const int kblocks = P.ntaps * P.cchunks;

  if (HALO && (warp == 0 || warp == 2)) {
    // ===== halo variant: warp 0 streams the halo patches (one per 32-channel chunk), warp 2 the weight tiles
    // (one per chunk and tap); the two rings advance independently =====
    RD_REG_DEC();
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      uint8_t* bring = smem + Cfg::A_SLOTS * HALO_SLOT;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        int mt = tile / n_tiles;
        const int tw_i = mt % P.tiles_w; mt /= P.tiles_w;
        const int th_i = mt % P.tiles_h;
        const int tb_i = mt / P.tiles_h;
        for (int cc = 0; cc < P.cchunks; ++cc) {
          if (warp == 0) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[slot], HALO_BYTES);
            tma_load_4d(smem + slot * HALO_SLOT, &mapA, &full_bar[slot], cc * (BF16 ? 64 : 32), tw_i * P.tw * HALO_SUB - 1,
                        th_i * P.th - 1, tb_i);
            if (++slot == Cfg::A_SLOTS) { slot = 0; phase ^= 1; }
          } else {
            for (int tg = 0; tg < 9 / Cfg::B_TAPS; ++tg) {      // mapB is 3-D here: (k inside the tap, n, tap)
              mbar_wait(&bempty_bar[slot], phase ^ 1);
              mbar_arrive_expect_tx(&bfull_bar[slot], Cfg::B_SLOT_BYTES);
              tma_load_3d(bring + slot * Cfg::B_SLOT_BYTES, &mapB, &bfull_bar[slot], cc * (BF16 ? 64 : 32), nt * BN,
                          tg * Cfg::B_TAPS);
              if (++slot == Cfg::B_SLOTS) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (HALO && warp == 1) {
    // ===== halo variant MMA issuer: the whole warp walks the loop, one elected lane issues =====
    RD_REG_DEC();
    {
      constexpr uint32_t idesc = BF16 ? idesc_bf16(128, BN, 0, 0) : idesc_tf32(128, BN, 0, 0);
      int aslot = 0, bslot = 0;
      uint32_t aphase = 0, bphase = 0;
      const uint32_t a_base = smem_u32(smem), b_base = a_base + Cfg::A_SLOTS * HALO_SLOT;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int cc = 0; cc < P.cchunks; ++cc) {
          mbar_wait(&full_bar[aslot], aphase);
          const uint32_t sa = a_base + aslot * HALO_SLOT;
          for (int tg = 0; tg < 9 / Cfg::B_TAPS; ++tg) {
            mbar_wait(&bfull_bar[bslot], bphase);
            tc_fence_after();
            // descriptors are built once per tap; (sub-tile, k) only add a constant to the 16-byte address field (no
            // carry out of its 14 bits: shared memory ends below 256 KB)
            uint64_t da_tap[Cfg::B_TAPS], db_tap[Cfg::B_TAPS];
#pragma unroll
            for (int t = 0; t < Cfg::B_TAPS; ++t) {
              const int tap = tg * Cfg::B_TAPS + t;
              da_tap[t] = smem_desc_sw128(sa + ((tap / 3) * HALO_W + (tap % 3)) * 128, 16, HALO_W * 128);
              db_tap[t] = smem_desc_sw128(b_base + bslot * Cfg::B_SLOT_BYTES + t * Cfg::B_BYTES, 16, 1024);
            }
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < Cfg::B_TAPS; ++t)
#pragma unroll
                for (int sub = 0; sub < HALO_SUB; ++sub)
#pragma unroll
                  for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K: 8 fp32 (kind::tf32) or 16 bf16 (kind::f16)
                    const uint64_t da = da_tap[t] + (uint64_t)((sub * 8 * 128 + k * 32) >> 4);
                    const uint64_t db = db_tap[t] + (uint64_t)((k * 32) >> 4);
                    if (BF16) mma_bf16(d_tmem + sub * BN, da, db, idesc, (cc | tg | t | k) != 0);
                    else mma_tf32(d_tmem + sub * BN, da, db, idesc, (cc | tg | t | k) != 0);
                  }
              tc_commit(&bempty_bar[bslot]);
              if (tg == 9 / Cfg::B_TAPS - 1) tc_commit(&empty_bar[aslot]);
              if (tg == 9 / Cfg::B_TAPS - 1 && cc == P.cchunks - 1) tc_commit(&tfull_bar[acc]);
            }
            __syncwarp();
            if (++bslot == Cfg::B_SLOTS) { bslot = 0; bphase ^= 1; }
          }
          if (++aslot == Cfg::A_SLOTS) { aslot = 0; aphase ^= 1; }
        }
      }
    }
  } else if (!HALO && warp == 0) {
    // ===== TMA producer: lane 0 waits for the slot and arms the barrier, then lane 0 issues the A box and
    // lane 1 the B box (coordinates in registers, no indexed arrays) =====
    RD_REG_DEC();
    const bool up2 = P.coord_b < 0;                   // (c, w, a, b*h) view of the 2x-upsampled tensor; else (c, w, h, b)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % n_tiles;
      int mt = tile / n_tiles;
      const int tw_i = mt % P.tiles_w; mt /= P.tiles_w;
      const int th_i = mt % P.tiles_h;
      const int tb_i = mt / P.tiles_h;
      const int w0 = tw_i * P.tw, h0 = th_i * P.th, b0 = tb_i * P.tb;
      for (int tap = 0; tap < P.ntaps; ++tap) {
        const int off0 = P.tap_off[tap][0], off1 = P.tap_off[tap][1], off2 = P.tap_off[tap][2];
        for (int cc = 0; cc < P.cchunks; ++cc) {
          if (lane == 0) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          }
          __syncwarp();
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if (lane == 0)
            tma_load_4d(sa, &mapA, &full_bar[stage], off0 + cc * (BF16 ? 64 : 32), w0 + off1, up2 ? off2 : h0 + off2, up2 ? h0 : b0);
          else if (lane == 1)
            tma_load_2d(sa + Cfg::A_BYTES, &mapB, &full_bar[stage], (tap * P.cchunks + cc) * (BF16 ? 64 : 32), nt * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (!HALO && warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop, one elected lane issues =====
    RD_REG_DEC();
    {
      constexpr uint32_t idesc = BF16 ? idesc_bf16(128, BN, 0, 0) : idesc_tf32(128, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint64_t da0 = smem_desc_sw128(smem_u32(smem), 16, 1024);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // stage descriptors differ only in the 16-byte address field (no carry: shared memory ends below 256 KB)
          const uint64_t da = da0 + (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
          const uint64_t db = da + (uint64_t)(Cfg::A_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { // 4 x 32 bytes of K (8 fp32 / 16 bf16) inside the 128-byte swizzle atom
              if (BF16) mma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              else mma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            tc_commit(&empty_bar[stage]);           // frees the smem stage when these MMAs have read it
            if (kb == kblocks - 1) tc_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== 8 epilogue warps: TMEM lane quarter = warp % 4; the two warps of a quarter take the even / odd
    // 32-column chunks, doubling the loads and stores in flight for the HBM-bound layers =====
    RD_REG_INC();
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    float* stg = reinterpret_cast<float*>(smem + Cfg::DATA_BYTES + 512) + (warp - 4) * 1024;   // 4 KB per warp
    if constexpr (RING) {
      convt_ring_epilogue<BN>(P, smem + Cfg::RING_OFF + (warp - 4) * ring_slots(BN) * 4096, stg, lane, q, half, tmem_base,
                              tfull_bar, tempty_bar, num_tiles, n_tiles);
    } else {
    constexpr int NCH = BN / 32, NCH2 = (NCH + 1) / 2;
    float4 cs1[NCH2], cs2[NCH2];                        // transposed-domain column sums: channels 4*(lane%8)..+3 of chunk ci
#pragma unroll
    for (int i = 0; i < NCH2; ++i) cs1[i] = cs2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int iw = row % P.tw, ih = (row / P.tw) % P.th, ib = row / (P.tw * P.th);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int nt = tile % n_tiles;
      int mt = tile / n_tiles;
      const int tw_i = mt % P.tiles_w; mt /= P.tiles_w;
      const int th_i = mt % P.tiles_h;
      const int tb_i = mt / P.tiles_h;
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < Cfg::SUB; ++sub) {       // halo variant: two 16x8-pixel sub-tiles side by side
      const int w = (tw_i * Cfg::SUB + sub) * P.tw + iw, h = th_i * P.th + ih, b = tb_i * P.tb + ib;
      const bool valid = (w < P.Wo) && (h < P.Ho) && (b < P.Bo);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS + sub * BN;
      const size_t pix = ((size_t)b * P.Ho + h) * P.Wo + w;
      if (P.epi_mode == EPI_CONVT) {
        convt_epilogue<BN>(P, stg, lane, half, t_row, nt, b, h, w, valid);
        continue;
      }
      unsigned Rt[8], okt = 0;                           // PLAIN / STATS: transposed-domain row offsets of this sub-tile
      if (P.epi_mode != EPI_BNACT) rows_to_transposed(lane, (unsigned)(pix * P.N), valid, Rt, okt);
#pragma unroll
      for (int ci = 0; ci < NCH2; ++ci) {
        const int ch = 2 * ci + half;
        if (ch >= NCH) break;
        const int n = nt * BN + ch * 32;
        float4 bn_sc = make_float4(1.f, 1.f, 1.f, 1.f), bn_sh = make_float4(0.f, 0.f, 0.f, 0.f);
        float bn_slope = 0.f;
        if (P.epi_mode == EPI_BNACT) {                     // this lane's four channels of the transposed domain
          bn_sc = __ldg(reinterpret_cast<const float4*>(P.scale + n) + (lane & 7));
          bn_sh = __ldg(reinterpret_cast<const float4*>(P.shift + n) + (lane & 7));
          bn_slope = __ldg(P.slope);
        }
        float v[32];
        tmem_ld32(t_row + ch * 32, v);
        if (P.epi_mode == EPI_BNACT) {
          // eval-mode BatchNorm folded into the conv: a = act(acc*scale + shift); optional fused 2x2 max-pool
          const size_t pp = ((size_t)b * (P.Ho >> 1) + (h >> 1)) * (P.Wo >> 1) + (w >> 1);
          warp_bnact_store_rows(stg, lane, v, bn_sc, bn_sh, bn_slope, P.out, (long long)(pix * P.N + n), valid, P.round_tf32,
                                P.pool_out, (long long)(pp * P.N + n), valid && !(w & 1) && !(h & 1), P.round_pool, P.tw);
        } else {
          if (P.epi_mode == EPI_STATS)
            warp_store_rows<true>(stg, lane, v, P.out, Rt, okt, (unsigned)(n + 4 * (lane & 7)), P.round_tf32,
                                  reinterpret_cast<__nv_bfloat16*>(P.out_b), &cs1[ci], &cs2[ci]);
          else
            warp_store_rows(stg, lane, v, P.out, Rt, okt, (unsigned)(n + 4 * (lane & 7)), P.round_tf32,
                            reinterpret_cast<__nv_bfloat16*>(P.out_b));
        }
      }
      }   // sub-tiles
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (P.epi_mode == EPI_STATS) {
      // lane holds the sums of its four channels over the rows it handled (all tiles of this CTA); the four lane groups
      // (lane / 8) cover disjoint rows: two xor steps complete the warp's column sums in lanes 0..7.
      // partial row = (CTA group, warp); the n_tiles CTAs of a group (fixed N tile each: gridDim.x is a multiple
      // of n_tiles) fill disjoint column ranges of the same partial rows
      const int nt = blockIdx.x % n_tiles;
      float* dst = P.partials + ((size_t)((blockIdx.x / n_tiles) * 4 + q) * P.N + nt * BN) * 2;
      const bool had_tiles = blockIdx.x < num_tiles;
#pragma unroll
      for (int ci = 0; ci < NCH2; ++ci) {
        const int ch = 2 * ci + half;
        if (ch >= NCH) break;
        float a[8] = {cs1[ci].x, cs1[ci].y, cs1[ci].z, cs1[ci].w, cs2[ci].x, cs2[ci].y, cs2[ci].z, cs2[ci].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          a[e] += __shfl_xor_sync(0xffffffffu, a[e], 8);
          a[e] += __shfl_xor_sync(0xffffffffu, a[e], 16);
        }
        if (lane < 8) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = ch * 32 + lane * 4 + e;
            dst[col * 2 + 0] = had_tiles ? a[e] : 0.f;
            dst[col * 2 + 1] = had_tiles ? a[4 + e] : 0.f;
          }
        }
      }
    }
    }   // !RING
  } else {
    RD_REG_DEC();                                       // idle warps of the first warpgroup: the whole group must release
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}
#undef RD_REG_DEC
#undef RD_REG_INC

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int tc_encode_map(CUtensorMap* map, const void* base, int rank, const long long* dims, const long long* strides_bytes,
                  const int* box, int swizzle_atom32, int elem_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = (cuuint64_t)strides_bytes[i - 1];
  }
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                      : (swizzle_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled failed with code %d (rank %d dims %lld %lld %lld %lld box %d %d %d %d)", (int)r,
                rank, dims[0], rank > 1 ? dims[1] : 0, rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0, box[0],
                rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return 0;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

bool tc_rows_eligible(const Gather& g, int N, int bf16) {
  if (g.C % (bf16 ? 64 : 32) || N % 32) return false;
  if (g.ntaps == 4 && g.ups != 2) return false;
  return true;
}

int tc_pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 32;
}

int tc_make_rows_plan(TcRowsPlan* plan, const void* src, const Gather& g, int B, const void* w_nk, int N, int bf16) {
  if (!tc_rows_eligible(g, N, bf16)) return fail("tc rows plan: shape not eligible (C=%d N=%d bf16=%d)", g.C, N, bf16);
  TcRowsParams& P = plan->p;
  std::memset(&P, 0, sizeof(P));
  const int EB = bf16 ? 2 : 4;                 // element bytes; an operand row is always 128 bytes
  P.N = N;
  P.ntaps = g.ntaps;
  P.bf16 = bf16;
  P.kchunk = 128 / EB;
  P.cchunks = g.C / P.kchunk;
  plan->BN = tc_pick_bn(N);
  {
    // small layers (bottleneck, deepest levels at small batch): with 256-column tiles fewer than half of the SMs get a
    // tile (ncu round 1: 64 CTAs, 33 % tensor pipe on the bottleneck dgrad); 128-column tiles double the tile count
    const long long m_tiles = ((long long)B * g.Ho * g.Wo + 127) / 128;
    if (plan->BN == 256 && N % 128 == 0 && m_tiles * (N / 256) <= 74) plan->BN = 128;
  }
  {
    // HBM-bound up-conv forward (1 tap, N = 4C, K = C <= 128): 128-column tiles leave room for deeper skip rings
    static const bool ring128 = getenv("RESDEPTH_RING_BN128") != nullptr;
    if (ring128 && !bf16 && g.ntaps == 1 && N == 4 * g.C && g.C <= 128 && N % 128 == 0) plan->BN = 128;
  }
  plan->halo = false;
  long long dims[4], strides[3];
  int box[4];
  if (g.ups == 1) {
    // source NHWC [B, H, W, C] seen as (C, W, H, B)
    P.Wo = g.Wo; P.Ho = g.Ho; P.Bo = B;
    P.tw = pow2_floor(g.Wo < 16 ? g.Wo : 16);
    int th_max = 128 / P.tw;
    P.th = pow2_floor(g.Ho < th_max ? g.Ho : th_max);
    P.tb = 128 / (P.tw * P.th);
    P.coord_w = 1; P.coord_h = 2; P.coord_b = 3;
    for (int t = 0; t < g.ntaps; ++t) { P.tap_off[t][0] = 0; P.tap_off[t][1] = g.dw[t]; P.tap_off[t][2] = g.dh[t]; P.tap_off[t][3] = 0; }
    dims[0] = g.C; dims[1] = g.Ws; dims[2] = g.Hs; dims[3] = B;
    strides[0] = (long long)g.C * EB; strides[1] = (long long)g.Ws * g.C * EB; strides[2] = (long long)g.Hs * g.Ws * g.C * EB;
    box[0] = P.kchunk; box[1] = P.tw; box[2] = P.th; box[3] = P.tb;
    static const bool no_halo = getenv("RESDEPTH_NO_HALO") != nullptr;
    if (g.ntaps == 9 && g.Ho >= 16 && g.Wo >= 16 && plan->BN <= 128 && !no_halo) {
      // halo-reuse variant: 16 x 16 pixel tiles (two 16 x 8 MMA sub-tiles), one 18 x 18 halo patch per 32-channel
      // chunk serves all nine taps and both sub-tiles; N tile capped at 128 (TMEM: 2 stages x 2 sub-tiles x BN)
      plan->halo = true;
      P.tw = 8; P.th = 16; P.tb = 1;
      box[1] = HALO_W; box[2] = HALO_H; box[3] = 1;
    }
  } else {
    // 2x2 stride-2 gather of dU NHWC [B, 2Hin, 2Win, C] seen as (2C, Win, 2, B*Hin): tap (a, b) = (coord2, coord0 / C)
    P.Wo = g.Wo; P.Ho = B * g.Ho; P.Bo = 1;
    P.tw = pow2_floor(g.Wo < 16 ? g.Wo : 16);
    int th_max = 128 / P.tw;
    P.th = pow2_floor(P.Ho < th_max ? P.Ho : th_max);
    P.tb = 128 / (P.tw * P.th);
    if (P.tb != 1) { P.th = 128 / P.tw; P.tb = 1; }     // rows beyond B*Hin are zero-filled and masked
    P.coord_w = 1; P.coord_h = 3; P.coord_b = -1;
    for (int t = 0; t < 4; ++t) { P.tap_off[t][0] = g.dw[t] * g.C; P.tap_off[t][1] = 0; P.tap_off[t][2] = g.dh[t]; P.tap_off[t][3] = 0; }
    dims[0] = 2LL * g.C; dims[1] = g.Wo; dims[2] = 2; dims[3] = (long long)B * g.Ho;
    strides[0] = 2LL * g.C * EB; strides[1] = (long long)g.Wo * 2 * g.C * EB; strides[2] = 2LL * g.Wo * 2 * g.C * EB;
    box[0] = P.kchunk; box[1] = P.tw; box[2] = 1; box[3] = P.th;
  }
  P.tiles_w = cdiv(P.Wo, P.tw * (plan->halo ? HALO_SUB : 1));
  P.tiles_h = cdiv(P.Ho, P.th);
  P.tiles_b = cdiv(P.Bo, P.tb);
  RD_TRY(tc_encode_map(&plan->mapA, src, 4, dims, strides, box, 0, EB));
  const long long K = (long long)g.ntaps * g.C;
  long long wd[2] = {K, N}, ws[1] = {K * EB};
  int wb[2] = {P.kchunk, plan->BN};
  if (plan->halo) {
    // weights [N][(tap, c)] as a 3-D tensor (c, n, tap): one box = the tiles of 3 consecutive taps, back to back
    long long wd3[3] = {g.C, N, 9}, ws3[2] = {K * EB, (long long)g.C * EB};
    int wb3[3] = {P.kchunk, plan->BN, 3};
    RD_TRY(tc_encode_map(&plan->mapB, w_nk, 3, wd3, ws3, wb3, 0, EB));
  } else {
    RD_TRY(tc_encode_map(&plan->mapB, w_nk, 2, wd, ws, wb, 0, EB));
  }
  plan->valid = true;
  return 0;
}

template <int BN, bool HALO, bool BF16, bool RING = false>
static int launch_rows_t(const TcRowsPlan& plan, const TcRowsParams& P, int grid, cudaStream_t s) {
  using Cfg = RowsCfg<BN, HALO, RING>;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(cudaFuncSetAttribute(gemm_rows_tc_kernel<BN, HALO, BF16, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES));
    attr_set = true;
  }
  gemm_rows_tc_kernel<BN, HALO, BF16, RING><<<grid, ROWS_THREADS, Cfg::SMEM_BYTES, s>>>(plan.mapA, plan.mapB, P);
  RD_LAUNCHED();
  return 0;
}
template <int BN, bool HALO>
static int launch_rows(const TcRowsPlan& plan, const TcRowsParams& P, int grid, cudaStream_t s) {
  return P.bf16 ? launch_rows_t<BN, HALO, true>(plan, P, grid, s) : launch_rows_t<BN, HALO, false>(plan, P, grid, s);
}

int launch_gemm_rows_tc(const TcRowsPlan& plan, const Epilogue& e, int* n_partials, cudaStream_t s) {
  if (!plan.valid) return fail("tc rows: plan not built");
  TcRowsParams P = plan.p;
  P.epi_mode = e.mode;
  P.out = e.out;
  P.partials = e.partials;
  P.bias = e.bias;
  P.skip = e.skip;
  P.skip_scale = e.skip_scale;
  P.skip_shift = e.skip_shift;
  P.skip_slope = e.skip_slope;
  P.round_tf32 = e.round_tf32;
  P.scale = e.scale;
  P.shift = e.shift;
  P.slope = e.slope;
  P.pool_out = e.pool_out;
  P.round_pool = e.round_pool;
  P.out_b = e.out_b;
  if (e.mode != EPI_BNACT && (double)P.Bo * P.Ho * P.Wo * P.N >= 4294967296.0)
    return fail("tc rows: output of 2^32 elements or more (32-bit element offsets in the epilogue: reduce the batch)");
  if (e.mode == EPI_BNACT && e.pool_out && (P.tw < 2 || P.th < 2 || (P.Wo & 1) || (P.Ho & 1)))
    return fail("tc rows: fused pooling needs even image sizes and a tile of at least 2x2 pixels");
  const int n_tiles = P.N / plan.BN;
  const int num_tiles = P.tiles_w * P.tiles_h * P.tiles_b * n_tiles;
  int grid = (148 / n_tiles) * n_tiles;
  if (grid < n_tiles) grid = n_tiles;
  if (grid > num_tiles) grid = ((num_tiles + n_tiles - 1) / n_tiles) * n_tiles;
  if (n_partials) *n_partials = (grid / n_tiles) * 4;
  if (plan.halo) {
    switch (plan.BN) {
      case 128: return launch_rows<128, true>(plan, P, grid, s);
      case 64: return launch_rows<64, true>(plan, P, grid, s);
      case 32: return launch_rows<32, true>(plan, P, grid, s);
    }
  } else {
    switch (plan.BN) {
      case 256:
        // HBM-bound up-convs (short K): skip rows prefetched through per-warp cp.async rings
        if (e.mode == EPI_CONVT && e.skip && !P.bf16 && P.ntaps * P.cchunks <= 4)
          return launch_rows_t<256, false, false, true>(plan, P, grid, s);
        return launch_rows<256, false>(plan, P, grid, s);
      case 128:
        if (e.mode == EPI_CONVT && e.skip && !P.bf16 && P.ntaps * P.cchunks <= 4)
          return launch_rows_t<128, false, false, true>(plan, P, grid, s);
        return launch_rows<128, false>(plan, P, grid, s);
      case 64: return launch_rows<64, false>(plan, P, grid, s);
      case 32: return launch_rows<32, false>(plan, P, grid, s);
    }
  }
  return fail("tc rows: unsupported BN=%d", plan.BN);
}

// ---------------------------------------------------------------------------------------------
// reduce kernel (weight gradients): D[(tap,ca)][n] = sum over pixels A[p + off(tap)][ca] * G[p][n]
//   * the contraction index is the PIXEL, so both operands are MN-major: a TMA box of 32 pixels x 32 channels
//     lands as 32 rows (k) x 128 bytes (m or n).  MN-major TF32 operands must use the "128-byte swizzle with
//     32-byte atomicity" layout (TMA SWIZZLE_128B_ATOM_32B / descriptor layout SWIZZLE_128B_BASE32B): 4-row x
//     128-byte atoms, 512 bytes apart along K.  A 128-row A tile is four such boxes (one per 32-channel chunk,
//     each with its own tap shift), the G tile is BN/32 boxes.
//   * one CTA = one (128-row tile, BN-column tile, pixel split); accumulates in TMEM over its pixel boxes and
//     writes its partial [128][BN] block; the un-pack kernels sum the splits.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct ReduceCfg {
  // pixels per pipeline stage: the TMA unit costs ~130 cycles per instruction, so narrow N tiles (few MMA cycles
  // per pixel) take 64 pixels per stage to amortise it
  static constexpr int KP = (BN == 256) ? 32 : 64;
  static constexpr int BOX_BYTES = KP * 128;
  static constexpr int NB = BN / 32;
  static constexpr int STAGE_BYTES = (4 + NB) * BOX_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};
static int reduce_kp(int BN) { return BN == 256 ? 32 : 64; }

template <int BN, bool EXPERIMENT_KMAJOR = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_reduce_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapG,
                      const TcReduceParams P) {
  using Cfg = ReduceCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* done_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapG);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * 128;
  const int total_boxes = P.tiles_w * P.tiles_h * P.tiles_b;
  const int box_begin = blockIdx.z * P.boxes_per_split;
  int box_end = box_begin + P.boxes_per_split;
  if (box_end > total_boxes) box_end = total_boxes;
  const int nsteps = box_end > box_begin ? box_end - box_begin : 0;
  int nchunks = (P.Mrows - m0 + 31) / 32;             // live 32-row chunks of this M tile
  if (nchunks > 4) nchunks = 4;

  if (warp == 0 && P.a_nch > 0) {
    // ===== TMA producer, grouped loads: the TMA unit needs ~120-150 cycles per instruction whatever the box
    // size, so a stage is fetched with as few instructions as possible -- lane g < groups: a_nch consecutive
    // 32-channel chunks of A (one tap) through a 5-D map, lane 4: all BN/32 chunks of G =====
    const int groups = 4 / P.a_nch;
    const bool is_a = lane < 4;
    int chunk0 = 0, off1 = 0, off2 = 0;
    bool active;
    if (is_a) {
      const int r = m0 + lane * P.a_nch * 32;
      active = lane < groups && r < P.Mrows;
      const int tap = active ? r / P.Ca : 0;
      chunk0 = P.a_chunk_off[tap] + (active ? (r % P.Ca) / 32 : 0);
      off1 = P.tap_off[tap][1];
      off2 = P.tap_off[tap][2];
    } else {
      active = lane == 4;
      chunk0 = n0 / 32;
    }
    int live_chunks = 0;                                 // A chunks actually fetched (zero-filled ones included)
    for (int gI = 0; gI < groups; ++gI)
      if (m0 + gI * P.a_nch * 32 < P.Mrows) live_chunks += P.a_nch;
    const CUtensorMap* map = is_a ? &mapA : &mapG;
    const bool up2 = P.coord_b < 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int bx = box_begin; bx < box_end; ++bx) {
      int t = bx;
      const int tw_i = t % P.tiles_w; t /= P.tiles_w;
      const int th_i = t % P.tiles_h;
      const int tb_i = t / P.tiles_h;
      const int w0 = tw_i * P.tw, h0 = th_i * P.th, b0 = tb_i * P.tb;
      if (lane == 0) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], (live_chunks + Cfg::NB) * Cfg::BOX_BYTES);
      }
      __syncwarp();
      if (active) {
        uint8_t* dst = smem + stage * Cfg::STAGE_BYTES + (is_a ? lane * P.a_nch : 4) * Cfg::BOX_BYTES;
        tma_load_5d(dst, map, &full_bar[stage], 0, w0 + off1, up2 ? off2 : h0 + off2, up2 ? h0 : b0, chunk0);
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 0) {
    // ===== TMA producer: lane i owns box i of every stage (A chunks 0..3, then the G chunks), so the
    // per-instruction issue cost of the 4 KB boxes is spread over up to 12 lanes; all coordinates live in
    // registers (no indexed arrays) =====
    const bool is_a = lane < 4;
    const bool active = is_a ? (lane < nchunks) : (lane < 4 + Cfg::NB);
    int off0 = 0, off1 = 0, off2 = 0;
    if (is_a) {
      const int r = m0 + lane * 32;
      const int tap = r < P.Mrows ? r / P.Ca : 0;
      off0 = P.tap_off[tap][0] + (r < P.Mrows ? r % P.Ca : 0);
      off1 = P.tap_off[tap][1];
      off2 = P.tap_off[tap][2];
    } else {
      off0 = n0 + (lane - 4) * 32;
    }
    const CUtensorMap* map = is_a ? &mapA : &mapG;
    const bool up2 = P.coord_b < 0;                   // (c, w, a, b*h) view; else (c, w, h, b)
    int stage = 0;
    uint32_t phase = 0;
    for (int bx = box_begin; bx < box_end; ++bx) {
      int t = bx;
      const int tw_i = t % P.tiles_w; t /= P.tiles_w;
      const int th_i = t % P.tiles_h;
      const int tb_i = t / P.tiles_h;
      const int w0 = tw_i * P.tw, h0 = th_i * P.th, b0 = tb_i * P.tb;
      if (lane == 0) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], (nchunks + Cfg::NB) * Cfg::BOX_BYTES);
      }
      __syncwarp();
      if (active) {
        uint8_t* dst = smem + stage * Cfg::STAGE_BYTES + lane * Cfg::BOX_BYTES;
        tma_load_4d(dst, map, &full_bar[stage], off0, w0 + off1, up2 ? off2 : h0 + off2, up2 ? h0 : b0);
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = EXPERIMENT_KMAJOR ? idesc_tf32(128, BN, 0, 0) : idesc_tf32(128, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nsteps; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sg = sa + 4 * Cfg::BOX_BYTES;
        if (elect_one()) {
#pragma unroll
        for (int k = 0; k < Cfg::KP / 8; ++k) {                         // 8 pixels = two 4-row swizzle atoms per MMA
          const uint64_t da = EXPERIMENT_KMAJOR ? smem_desc_sw128(sa + k * 32, 16, 1024)
                                                : smem_desc_mn_sw128_32b(sa + k * 1024, Cfg::BOX_BYTES, 512);
          const uint64_t dg = EXPERIMENT_KMAJOR ? smem_desc_sw128(sg + k * 32, 16, 1024)
                                                : smem_desc_mn_sw128_32b(sg + k * 1024, Cfg::BOX_BYTES, 512);
          mma_tf32(tmem_base, da, dg, idesc, (i | k) != 0);
        }
        tc_commit(&empty_bar[stage]);
        if (i == nsteps - 1) tc_commit(done_bar);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    float* out = P.part + ((size_t)blockIdx.z * P.Mrows + m) * P.N + n0;
    if (nsteps > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
      float v[32];
      if (nsteps > 0) {
        tmem_ld32(t_row + ch * 32, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (m < P.Mrows) {
        float4* op = reinterpret_cast<float4*>(out + ch * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// bf16 variant of the reduce kernel (backward only; SURVEY.md 0.5: the gradients tolerate bf16 operands).
// 64-channel chunks (128 bytes of bf16), standard 128-byte swizzle, MN-major atoms of 8 pixel rows, K = 16 pixels
// per tcgen05.mma kind::f16 -- half the operand bytes and twice the MMA rate of the TF32 kernel.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct ReduceBCfg {
  static constexpr int KP = 64;                      // pixels per stage = 4 MMAs of K = 16
  static constexpr int BOX_BYTES = KP * 128;
  static constexpr int NB = BN / 64;
  static constexpr int STAGE_BYTES = (2 + NB) * BOX_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = BN < 32 ? 32 : (BN == 192 ? 256 : BN);     // power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_reduce_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapG,
                        const TcReduceParams P) {
  using Cfg = ReduceBCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* done_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapG);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * 128;
  const int total_boxes = P.tiles_w * P.tiles_h * P.tiles_b;
  const int box_begin = blockIdx.z * P.boxes_per_split;
  int box_end = box_begin + P.boxes_per_split;
  if (box_end > total_boxes) box_end = total_boxes;
  const int nsteps = box_end > box_begin ? box_end - box_begin : 0;

  if (warp == 0) {
    const int groups = 2 / P.a_nch;                    // A: two 64-channel chunks per 128-row tile
    const bool is_a = lane < 4;
    int chunk0 = 0, off1 = 0, off2 = 0, g_slot = 2;
    bool active;
    if (is_a) {
      const int r = m0 + lane * P.a_nch * 64;
      active = lane < groups && r < P.Mrows;
      const int tap = active ? r / P.Ca : 0;
      chunk0 = P.a_chunk_off[tap] + (active ? (r % P.Ca) / 64 : 0);
      off1 = P.tap_off[tap][1];
      off2 = P.tap_off[tap][2];
    } else if (P.g_taps > 0) {
      // wide mode: lane 4+t fetches tap (N tile * g_taps + t) of the N operand at its own pixel shift; a tap
      // beyond the ninth is fetched far outside the tensor (zero-filled box, columns never read back)
      const int tl = lane - 4;
      active = tl < P.g_taps;
      const int tap = blockIdx.x * P.g_taps + (active ? tl : 0);
      off1 = tap < 9 ? P.g_off[tap][0] : (1 << 20);
      off2 = tap < 9 ? P.g_off[tap][1] : 0;
      g_slot = 2 + tl * P.g_nch;
    } else {
      active = lane == 4;
      chunk0 = n0 / 64;
      if (P.g_self) { chunk0 = 0; g_slot = 1; }         // [A chunk | G chunk]: G lands right behind the (single) A box
    }
    int live_chunks = 0;
    for (int gI = 0; gI < groups; ++gI)
      if (m0 + gI * P.a_nch * 64 < P.Mrows) live_chunks += P.a_nch;
    const int g_boxes = P.g_self ? 1 : Cfg::NB;
    const CUtensorMap* map = is_a ? &mapA : &mapG;
    const bool up2 = P.coord_b < 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int bx = box_begin; bx < box_end; ++bx) {
      int t = bx;
      const int tw_i = t % P.tiles_w; t /= P.tiles_w;
      const int th_i = t % P.tiles_h;
      const int tb_i = t / P.tiles_h;
      const int w0 = tw_i * P.tw, h0 = th_i * P.th, b0 = tb_i * P.tb;
      if (lane == 0) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], (live_chunks + g_boxes) * Cfg::BOX_BYTES);
      }
      __syncwarp();
      if (active) {
        uint8_t* dst = smem + stage * Cfg::STAGE_BYTES + (is_a ? lane * P.a_nch : g_slot) * Cfg::BOX_BYTES;
        tma_load_5d(dst, map, &full_bar[stage], 0, w0 + off1, up2 ? off2 : h0 + off2, up2 ? h0 : b0, chunk0);
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = idesc_bf16(128, BN, 1, 1);       // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nsteps; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sg = P.g_self ? sa : sa + 2 * Cfg::BOX_BYTES;     // self mode: N operand = [A box | G box]
        const uint64_t da0 = smem_desc_sw128(sa, Cfg::BOX_BYTES, 1024), dg0 = smem_desc_sw128(sg, Cfg::BOX_BYTES, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < Cfg::KP / 16; ++k)                    // 16 pixels = two 8-row swizzle atoms per MMA
            mma_bf16(tmem_base, da0 + (uint64_t)(k * (2048 >> 4)), dg0 + (uint64_t)(k * (2048 >> 4)), idesc, (i | k) != 0);
          tc_commit(&empty_bar[stage]);
          if (i == nsteps - 1) tc_commit(done_bar);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    float* out = P.part + ((size_t)blockIdx.z * P.Mrows + m) * P.N + n0;
    if (nsteps > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
      float v[32];
      if (nsteps > 0) {
        tmem_ld32(t_row + ch * 32, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (m < P.Mrows) {
        float4* op = reinterpret_cast<float4*>(out + ch * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

bool tc_reduce_eligible(const Gather& g, int N, int bf16) {
  if (!(g.ups == 1 || (g.ups == 2 && g.ntaps == 4))) return false;
  if (bf16) return g.C % 64 == 0 && N % 64 == 0 && (g.C % 128 == 0 || g.C == 64);
  return g.C % 32 == 0 && N % 32 == 0;
}

// a_pitch / g_pitch (elements, 0 = dense): the A / G tensor stores only its first `pitch` columns per pixel (row pitch =
// pitch elements); the tensor map declares that inner extent and the TMA unit zero-fills the rest of the 128-byte box,
// so a 32-column bf16 expansion feeds the same 64-column operand tiles at half the HBM traffic.
int tc_make_reduce_plan(TcReducePlan* plan, const void* src, const Gather& g, int B, const void* G, int N,
                        float* part, size_t part_floats, int bf16, int a_pitch, int g_pitch) {
  if (!tc_reduce_eligible(g, N, bf16)) return fail("tc reduce plan: shape not eligible (C=%d N=%d bf16=%d)", g.C, N, bf16);
  if ((a_pitch || g_pitch) && !(bf16 && g.ups == 1 && g.ntaps == 1 && g.C == 64 && (!g_pitch || N == 64)))
    return fail("tc reduce plan: pitched operands are supported for the bf16 1-tap 64-column case only");
  TcReduceParams& P = plan->p;
  std::memset(&P, 0, sizeof(P));
  P.bf16 = bf16;
  const int EB = bf16 ? 2 : 4, CH = 128 / EB;          // element bytes, channels per 128-byte chunk
  P.Mrows = g.ntaps * g.C;
  P.Ca = g.C;
  P.N = N;
  P.ntaps = g.ntaps;
  P.part = part;
  plan->BN = tc_pick_bn(N);
  if (bf16 && plan->BN < 64) return fail("tc reduce plan: bf16 needs N %% 64 == 0");
  long long dims[4], strides[3], gdims[4], gstrides[3];
  int box[4];
  int Wg, Hg, Bg;                                      // pixel grid of the contraction
  if (g.ups == 1) {
    Wg = g.Wo; Hg = g.Ho; Bg = B;
    P.coord_w = 1; P.coord_h = 2; P.coord_b = 3;
    for (int t = 0; t < g.ntaps; ++t) { P.tap_off[t][0] = 0; P.tap_off[t][1] = g.dw[t]; P.tap_off[t][2] = g.dh[t]; P.tap_off[t][3] = 0; }
    const long long ap = a_pitch ? a_pitch : g.C, gp = g_pitch ? g_pitch : N;
    dims[0] = ap; dims[1] = g.Ws; dims[2] = g.Hs; dims[3] = B;
    strides[0] = ap * EB; strides[1] = (long long)g.Ws * ap * EB; strides[2] = (long long)g.Hs * g.Ws * ap * EB;
    gdims[0] = gp; gdims[1] = Wg; gdims[2] = Hg; gdims[3] = B;
    gstrides[0] = gp * EB; gstrides[1] = (long long)Wg * gp * EB; gstrides[2] = (long long)Hg * Wg * gp * EB;
  } else {
    Wg = g.Wo; Hg = B * g.Ho; Bg = 1;
    P.coord_w = 1; P.coord_h = 3; P.coord_b = -1;
    for (int t = 0; t < 4; ++t) { P.tap_off[t][0] = g.dw[t] * g.C; P.tap_off[t][1] = 0; P.tap_off[t][2] = g.dh[t]; P.tap_off[t][3] = 0; }
    dims[0] = 2LL * g.C; dims[1] = g.Wo; dims[2] = 2; dims[3] = (long long)B * g.Ho;
    strides[0] = 2LL * g.C * EB; strides[1] = (long long)g.Wo * 2 * g.C * EB; strides[2] = 2LL * g.Wo * 2 * g.C * EB;
    gdims[0] = N; gdims[1] = Wg; gdims[2] = 1; gdims[3] = Hg;
    gstrides[0] = (long long)N * EB; gstrides[1] = (long long)Wg * N * EB; gstrides[2] = (long long)Wg * N * EB;
  }
  const int KP = bf16 ? 64 : reduce_kp(plan->BN);
  P.tw = pow2_floor(Wg < KP ? Wg : KP);
  int th_max = KP / P.tw;
  P.th = pow2_floor(Hg < th_max ? Hg : th_max);
  P.tb = KP / (P.tw * P.th);
  if (g.ups == 2 && P.tb != 1) { P.th = KP / P.tw; P.tb = 1; }
  P.tiles_w = cdiv(Wg, P.tw);
  P.tiles_h = cdiv(Hg, P.th);
  P.tiles_b = cdiv(Bg, P.tb);
  if (g.ups == 1) { box[0] = CH; box[1] = P.tw; box[2] = P.th; box[3] = P.tb; }
  else { box[0] = CH; box[1] = P.tw; box[2] = 1; box[3] = P.th; }
  // grouped 5-D maps (chunk index slowest) when the 128-row tile splits into whole same-tap groups
  static const bool no_group = getenv("RESDEPTH_NO_GROUPED_TMA") != nullptr;
  P.a_nch = 0;
  const int per_tile = 128 / CH;                         // chunks per 128-row M tile: 4 (fp32) or 2 (bf16)
  const int a_nch = g.C % 128 == 0 ? per_tile : (g.C == 64 ? 64 / CH : (g.C == 32 && !bf16 ? 1 : 0));
  bool grouped = false;
  if (a_nch > 0 && (!no_group || bf16)) {
    long long d5[5], s5[4], gd5[5], gs5[4];
    int b5[5], gb5[5];
    const long long csrc = g.ups == 1 ? g.C : 2LL * g.C;          // channels of one pixel row of the A tensor
    d5[0] = a_pitch ? a_pitch : CH; d5[1] = dims[1]; d5[2] = dims[2]; d5[3] = dims[3]; d5[4] = a_pitch ? 1 : csrc / CH;
    s5[0] = strides[0]; s5[1] = strides[1]; s5[2] = strides[2]; s5[3] = 128;
    b5[0] = CH; b5[1] = box[1]; b5[2] = box[2]; b5[3] = box[3]; b5[4] = a_nch;
    gd5[0] = g_pitch ? g_pitch : CH; gd5[1] = gdims[1]; gd5[2] = gdims[2]; gd5[3] = gdims[3]; gd5[4] = g_pitch ? 1 : N / CH;
    gs5[0] = gstrides[0]; gs5[1] = gstrides[1]; gs5[2] = gstrides[2]; gs5[3] = 128;
    gb5[0] = CH; gb5[1] = box[1]; gb5[2] = box[2]; gb5[3] = box[3]; gb5[4] = plan->BN / CH;
    // fp32 MN-major operands need the 32-byte-atom swizzle; bf16 the standard 128-byte swizzle
    if (tc_encode_map(&plan->mapA, src, 5, d5, s5, b5, bf16 ? 0 : 1, EB) == 0 &&
        tc_encode_map(&plan->mapG, G, 5, gd5, gs5, gb5, bf16 ? 0 : 1, EB) == 0) {
      grouped = true;
      P.a_nch = a_nch;
      for (int t = 0; t < g.ntaps; ++t) P.a_chunk_off[t] = P.tap_off[t][0] / CH;
    }
  }
  if (bf16 && !grouped) return fail("tc reduce plan: bf16 path needs grouped 5-D tensor maps (C=%d)", g.C);
  if (!grouped) {
    RD_TRY(tc_encode_map(&plan->mapA, src, 4, dims, strides, box, 1));
    RD_TRY(tc_encode_map(&plan->mapG, G, 4, gdims, gstrides, box, 1));
  }
  // split the pixel boxes so that the CTAs fill ONE wave of the 148 SMs (1 CTA per SM: the kernel is not
  // persistent, a second partial wave would idle most of the chip), bounded by the partial buffer
  const int total_boxes = P.tiles_w * P.tiles_h * P.tiles_b;
  const int tiles = cdiv(P.Mrows, 128) * (N / plan->BN);
  int S = 148 / tiles;
  if (S > total_boxes / 8) S = total_boxes / 8;         // at least 8 K steps per CTA
  if (S < 1) S = 1;
  const size_t per = (size_t)P.Mrows * N;
  while (S > 1 && (size_t)S * per > part_floats) --S;
  if ((size_t)S * per > part_floats) return fail("tc reduce plan: partial buffer too small (%zu floats needed)", per);
  P.boxes_per_split = cdiv(total_boxes, S);
  plan->splits = cdiv(total_boxes, P.boxes_per_split);
  plan->valid = true;
  return 0;
}

int tc_reduce_plan_add_gram(TcReducePlan* plan, size_t part_floats) {
  TcReduceParams& P = plan->p;
  if (!plan->valid || !P.bf16 || P.ntaps != 1 || P.Ca != 64 || P.Mrows != 64 || P.N != 64 || plan->BN != 64 || P.g_taps)
    return fail("tc reduce plan: the Gram-fused form needs a bf16 1-tap plan with 64-channel operands");
  const size_t per = (size_t)P.Mrows * 128;
  if ((size_t)plan->splits * per > part_floats) return fail("tc reduce plan: partial buffer too small for the Gram-fused form");
  P.g_self = 1;
  P.N = 128;
  plan->BN = 128;
  return 0;
}

bool tc_reduce_wide_eligible(int Cu, int Cs, int H, int W) {
  return tc_available() && Cu % 128 == 0 && (Cs == 64 || Cs == 128) && H >= 16 && W >= 16;
}

int tc_make_reduce_plan_wide(TcReducePlan* plan, const void* U, int Cu, const void* S, int Cs, int sign, int B, int H,
                             int W, float* part, size_t part_floats) {
  plan->valid = false;
  if (!tc_reduce_wide_eligible(Cu, Cs, H, W)) return fail("tc wide reduce plan: shape not eligible (Cu=%d Cs=%d)", Cu, Cs);
  TcReduceParams& P = plan->p;
  std::memset(&P, 0, sizeof(P));
  P.bf16 = 1;
  const int EB = 2, CH = 64, KP = 64;
  P.g_nch = Cs / 64;
  P.g_taps = Cs == 64 ? 3 : 2;                          // N tile: 3 x 64 = 192 or 2 x 128 = 256 columns
  plan->BN = P.g_taps * Cs;
  const int n_tiles = cdiv(9, P.g_taps);
  P.Mrows = Cu;
  P.Ca = Cu;
  P.N = n_tiles * plan->BN;                             // padded: taps beyond the ninth are zero columns
  P.ntaps = 1;
  P.part = part;
  P.coord_w = 1; P.coord_h = 2; P.coord_b = 3;
  for (int t = 0; t < 9; ++t) {
    P.g_off[t][0] = sign * (t % 3 - 1);
    P.g_off[t][1] = sign * (t / 3 - 1);
  }
  P.tw = pow2_floor(W < KP ? W : KP);
  const int th_max = KP / P.tw;
  P.th = pow2_floor(H < th_max ? H : th_max);
  P.tb = KP / (P.tw * P.th);
  P.tiles_w = cdiv(W, P.tw);
  P.tiles_h = cdiv(H, P.th);
  P.tiles_b = cdiv(B, P.tb);
  P.a_nch = 2;
  P.a_chunk_off[0] = 0;
  long long d5[5], s5[4];
  int b5[5];
  d5[0] = CH; d5[1] = W; d5[2] = H; d5[3] = B; d5[4] = Cu / CH;
  s5[0] = (long long)Cu * EB; s5[1] = (long long)W * Cu * EB; s5[2] = (long long)H * W * Cu * EB; s5[3] = 128;
  b5[0] = CH; b5[1] = P.tw; b5[2] = P.th; b5[3] = P.tb; b5[4] = 2;
  RD_TRY(tc_encode_map(&plan->mapA, U, 5, d5, s5, b5, 0, EB));
  d5[4] = Cs / CH;
  s5[0] = (long long)Cs * EB; s5[1] = (long long)W * Cs * EB; s5[2] = (long long)H * W * Cs * EB;
  b5[4] = P.g_nch;
  RD_TRY(tc_encode_map(&plan->mapG, S, 5, d5, s5, b5, 0, EB));
  const int total_boxes = P.tiles_w * P.tiles_h * P.tiles_b;
  const int tiles = (Cu / 128) * n_tiles;
  int Sp = 148 / tiles;
  if (Sp > total_boxes / 8) Sp = total_boxes / 8;
  if (Sp < 1) Sp = 1;
  const size_t per = (size_t)P.Mrows * P.N;
  while (Sp > 1 && (size_t)Sp * per > part_floats) --Sp;
  if ((size_t)Sp * per > part_floats) return fail("tc wide reduce plan: partial buffer too small (%zu floats needed)", per);
  P.boxes_per_split = cdiv(total_boxes, Sp);
  plan->splits = cdiv(total_boxes, P.boxes_per_split);
  plan->valid = true;
  return 0;
}

template <int BN>
static int launch_reduce(const TcReducePlan& plan, cudaStream_t s) {
  using Cfg = ReduceCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(cudaFuncSetAttribute(gemm_reduce_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(plan.p.N / BN, cdiv(plan.p.Mrows, 128), plan.splits);
  static const bool kmajor_experiment = getenv("RD_EXPERIMENT_REDUCE_KMAJOR") != nullptr;
  if (kmajor_experiment) {
    static bool attr2 = false;
    if (!attr2) {
      RD_CUDA(cudaFuncSetAttribute(gemm_reduce_tc_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      attr2 = true;
    }
    gemm_reduce_tc_kernel<BN, true><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(plan.mapA, plan.mapG, plan.p);
  } else
  gemm_reduce_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(plan.mapA, plan.mapG, plan.p);
  RD_LAUNCHED();
  return 0;
}

template <int BN>
static int launch_reduce_bf16(const TcReducePlan& plan, cudaStream_t s) {
  using Cfg = ReduceBCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(cudaFuncSetAttribute(gemm_reduce_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(plan.p.N / BN, cdiv(plan.p.Mrows, 128), plan.splits);
  gemm_reduce_bf16_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(plan.mapA, plan.mapG, plan.p);
  RD_LAUNCHED();
  return 0;
}

int launch_gemm_reduce_tc(const TcReducePlan& plan, cudaStream_t s) {
  if (!plan.valid) return fail("tc reduce: plan not built");
  if (plan.p.bf16) {
    switch (plan.BN) {
      case 256: return launch_reduce_bf16<256>(plan, s);
      case 192: return launch_reduce_bf16<192>(plan, s);
      case 128: return launch_reduce_bf16<128>(plan, s);
      case 64: return launch_reduce_bf16<64>(plan, s);
    }
    return fail("tc reduce bf16: unsupported BN=%d", plan.BN);
  }
  switch (plan.BN) {
    case 256: return launch_reduce<256>(plan, s);
    case 128: return launch_reduce<128>(plan, s);
    case 64: return launch_reduce<64>(plan, s);
    case 32: return launch_reduce<32>(plan, s);
  }
  return fail("tc reduce: unsupported BN=%d", plan.BN);
}

bool tc_available() { return encode_fn() != nullptr; }

// ---------------------------------------------------------------------------------------------
// First encoder conv (Cin <= 3: K = 9*Cin <= 27, padded to 32) on tcgen05 with a SOFTWARE im2col producer.
// The layer is HBM-bound (writes 64 channels per pixel, reads 3) but on CUDA cores its 1728 FMAs per pixel made it
// issue-bound at 2.5x the HBM time.  fp32 accuracy is kept with the 3xTF32 split: x = xh + xl, w = wh + wl (each
// exactly representable in TF32), D = xh*wh + xl*wh + xh*wl (the dropped xl*wl term is ~2^-22 relative).
//   warp 13  : TMA producer -- the (tw+2) x (th+2) x Cin halo patch of every 128-pixel tile (tw x th pixels of one
//              image) straight from the NCHW input, FIRST_XS patches ahead; out-of-image coordinates are zero-filled
//              by the TMA unit = the convolution padding (a first version gathered the 27 values with global loads
//              in the builder threads and was bound by one DRAM round trip per tile: ncu long-scoreboard stalls)
//   warps 0-3: builders -- thread r reads the 27 input values of pixel r from the patch in shared memory, splits
//              them and writes row r of the K-major SWIZZLE_128B operand tiles A_hi / A_lo (+ proxy fence)
//   warp 4   : TMEM allocator, MMA issuer (12 tcgen05.mma of K = 8 per 128-pixel tile)
//   warps 5-12: epilogue (TMEM lane quarter = warp % 4, even / odd 32-column chunk = (warp - 5) / 4): a lane owns one
//              output row (pixel); it writes its 128 bytes into a SWIZZLE_128B staging box and one lane issues a TMA
//              store of the box (32 channels x 32 pixels) -- no transposition, no STG; BatchNorm column sums are kept
//              per lane in registers across the CTA's tiles.  Fused inference form: BatchNorm + activation before the
//              store, 2x2 max-pool read back from the staged rows.
// ---------------------------------------------------------------------------------------------
static constexpr int FIRST_THREADS = 448;         // 4 builder warps, MMA warp, 8 epilogue warps, input-TMA warp
static constexpr int FIRST_XS = 4;                // input halo patches in flight
static constexpr int FIRST_X_BYTES = 8192;        // slot of one halo patch (Cin x (th+2) x (tw+8) fp32 <= 4896 B)
static constexpr int FIRST_STAGES = 3;
template <int N>
struct FirstCfg {
  static constexpr int A_BYTES = 128 * 128;                 // one operand tile: 128 rows x 32 fp32
  static constexpr int STAGE_BYTES = 2 * A_BYTES;           // hi + lo
  static constexpr int B_BYTES = N * 128;
  static constexpr int B_OFF = FIRST_STAGES * STAGE_BYTES;  // B_hi, then B_lo
  static constexpr int STG_OFF = B_OFF + 2 * B_BYTES;       // 8 warps x 2 x 4 KB epilogue staging (TMA-store sources)
  static constexpr int X_OFF = STG_OFF + 8 * 8192;          // FIRST_XS input halo patches
  static constexpr int BAR_OFF = X_OFF + FIRST_XS * FIRST_X_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
  static constexpr int TMEM_COLS = 2 * N <= 32 ? 32 : (2 * N <= 64 ? 64 : (2 * N <= 128 ? 128 : (2 * N <= 256 ? 256 : 512)));
};

// element (row, k) of a K-major SWIZZLE_128B tile: 8-row groups of 1024 bytes, 16-byte chunks XOR-ed with row % 8
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// FUSE (inference, RD_FWD_EVAL): BatchNorm(running statistics) + activation + 2x2 max-pool in the epilogue -- the raw
// conv output z is never written (at batch 32 it is 0.54 GB written here and read again by bn_act_pool).  The tile is
// then 64 columns x 2 image rows with row r = pixel (r >> 1, r & 1), so that the four pixels of a pooling window are
// the lanes l, l^1, l^2, l^3 of one warp; out_a = activated tensor (the additive skip), out_p = pooled tensor.
struct FirstFuse {
  const float* scale;
  const float* shift;
  const float* slope;
  float* out_a;
  float* out_p;
  int round_p;
};
template <int N, bool FUSE = false>
__global__ void __launch_bounds__(FIRST_THREADS, 1)
conv_first_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ,
                     const float* __restrict__ w, float* __restrict__ partials, int B, int Cin, int H, int W, int tw,
                     int th, const FirstFuse F) {
  using Cfg = FirstCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* a_empty = a_full + FIRST_STAGES;
  uint64_t* t_full = a_empty + FIRST_STAGES;
  uint64_t* t_empty = t_full + 2;
  uint64_t* x_full = t_empty + 2;
  uint64_t* x_empty = x_full + FIRST_XS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + FIRST_XS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = Cin * 9;
  const int tiles_w = (W + tw - 1) / tw, tiles_h = (H + th - 1) / th;
  const int num_tiles = tiles_w * tiles_h * B;
  // halo patch [Cin][BH][BW] floats starting at pixel (w0 - 4, h0 - 1): the innermost TMA coordinate must keep the
  // global address 16-byte aligned (an unaligned start is an illegal instruction), so the left halo is 4 wide
  const int BW = tw + 8, BH = th + 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FIRST_STAGES; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
    for (int i = 0; i < FIRST_XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 128); }
    fence_mbar_init();
    prefetch_tmap(&mapX);
    prefetch_tmap(&mapZ);
  }
  if (warp == 4) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  // weights [N][K] (OIHW flattening) -> B_hi / B_lo operand tiles, zero-padded to K = 32
  {
    const uint32_t bh = smem_u32(smem + Cfg::B_OFF), bl = bh + Cfg::B_BYTES;
    for (int i = threadIdx.x; i < N * 8; i += FIRST_THREADS) {
      const int n = i >> 3, c = i & 7;
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = c * 4 + e;
        const float v = k < K ? w[(size_t)n * K + k] : 0.f;
        hi[e] = tf32_round(v);
        lo[e] = tf32_round(v - hi[e]);
      }
      sts128(bh + sw128_off(n, c), make_float4(hi[0], hi[1], hi[2], hi[3]));
      sts128(bl + sw128_off(n, c), make_float4(lo[0], lo[1], lo[2], lo[3]));
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===== builders =====
    const int r = threadIdx.x;                              // row of the tile = pixel (r % tw, r / tw)
    const int iw = FUSE ? (r >> 1) : r % tw, ih = FUSE ? (r & 1) : r / tw;
    int stage = 0, xs = 0;
    uint32_t phase = 0, xphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&x_full[xs], xphase);
      const float* patch = reinterpret_cast<const float*>(smem + Cfg::X_OFF + xs * FIRST_X_BYTES);
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int ci = k / 9, rs = k - ci * 9;
        const int dr = rs / 3, ds = rs - dr * 3;            // patch origin is pixel (-4, -1) of the tile
        v[k] = k < K ? patch[(ci * BH + ih + dr) * BW + iw + ds + 3] : 0.f;
      }
      mbar_arrive(&x_empty[xs]);                            // values are in registers: the slot may be refilled
      if (++xs == FIRST_XS) { xs = 0; xphase ^= 1; }
      mbar_wait(&a_empty[stage], phase ^ 1);
      const uint32_t ah = smem_u32(smem + stage * Cfg::STAGE_BYTES), al = ah + Cfg::A_BYTES;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[e] = tf32_round(v[c * 4 + e]);
          lo[e] = tf32_round(v[c * 4 + e] - hi[e]);
        }
        sts128(ah + sw128_off(r, c), make_float4(hi[0], hi[1], hi[2], hi[3]));
        sts128(al + sw128_off(r, c), make_float4(lo[0], lo[1], lo[2], lo[3]));
      }
      fence_proxy_async();                                  // generic-proxy stores -> visible to tcgen05.mma
      mbar_arrive(&a_full[stage]);
      if (++stage == FIRST_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 13) {
    // ===== input TMA producer =====
    if (lane == 0) {
      int xs = 0;
      uint32_t xphase = 0;
      const uint32_t bytes = (uint32_t)(BW * BH * Cin * 4);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t = tile;
        const int tw_i = t % tiles_w; t /= tiles_w;
        const int th_i = t % tiles_h;
        const int b = t / tiles_h;
        mbar_wait(&x_empty[xs], xphase ^ 1);
        mbar_arrive_expect_tx(&x_full[xs], bytes);
        tma_load_4d(smem + Cfg::X_OFF + xs * FIRST_X_BYTES, &mapX, &x_full[xs], tw_i * tw - 4, th_i * th - 1, 0, b);
        if (++xs == FIRST_XS) { xs = 0; xphase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = idesc_tf32(128, N, 0, 0);
    const uint32_t bh = smem_u32(smem + Cfg::B_OFF), bl = bh + Cfg::B_BYTES;
    const uint64_t dbh = smem_desc_sw128(bh, 16, 1024), dbl = smem_desc_sw128(bl, 16, 1024);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&t_empty[acc], ((it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      const uint32_t ah = smem_u32(smem + stage * Cfg::STAGE_BYTES);
      const uint64_t dah = smem_desc_sw128(ah, 16, 1024), dal = smem_desc_sw128(ah + Cfg::A_BYTES, 16, 1024);
      const uint32_t d_tmem = tmem_base + acc * N;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(d_tmem, dah + (uint64_t)(k * 2), dbh + (uint64_t)(k * 2), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(d_tmem, dal + (uint64_t)(k * 2), dbh + (uint64_t)(k * 2), idesc, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32(d_tmem, dah + (uint64_t)(k * 2), dbl + (uint64_t)(k * 2), idesc, 1);
        tc_commit(&a_empty[stage]);
        tc_commit(&t_full[acc]);
      }
      __syncwarp();
      if (++stage == FIRST_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===== epilogue: warps 5..12; a warp may only read the TMEM lane quarter warp % 4, the two warps of a quarter
    // take the even / odd 32-column chunk.  BatchNorm sums: every lane keeps per-column running sums of ITS rows
    // over all tiles of the CTA (64 registers); the cross-lane reduction happens once, after the last tile =====
    const int q = warp & 3;
    const int half = (warp - 5) >> 2;
    const int row = q * 32 + lane;
    // Full-resolution output rows leave through TMA stores: a lane writes ITS row (32 channels = 128 bytes) into a
    // SWIZZLE_128B box [32 pixels][128 B] (conflict-free STS.128), one lane issues the bulk store.  The transposing
    // round trip this replaces (STS + LDS + STG per 16 bytes) was two thirds of the epilogue's L1 wavefronts in a
    // kernel ncu showed at 81 % l1tex throughput and 42 % DRAM.  Box = (32 ch, bw, 32 / bw, 1) pixels of one image;
    // pixels outside the image are clipped by the TMA unit.
    uint8_t* stg_raw = smem + Cfg::STG_OFF + (warp - 5) * 8192;
    constexpr int NCH = N / 32;
    static_assert(NCH <= 2, "one 32-column chunk per epilogue warp");
    const bool has_chunk = half < NCH;
    float s1[32], s2[32];                                   // FUSE: scale / shift of this warp's 32 columns
#pragma unroll
    for (int j = 0; j < 32; ++j) s1[j] = s2[j] = 0.f;
    float slope = 0.f;
    if (FUSE && has_chunk) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { s1[j] = __ldg(F.scale + half * 32 + j); s2[j] = __ldg(F.shift + half * 32 + j); }
      slope = __ldg(F.slope);
    }
    const int iw = FUSE ? (row >> 1) : row % tw, ih = FUSE ? (row & 1) : row / tw;
    // first pixel of the warp's box inside the tile and this lane's row inside the box (w fastest, then h)
    const int bw0 = FUSE ? q * 16 : (q * 32) % tw, bh0 = FUSE ? 0 : (q * 32) / tw;
    const int srow = FUSE ? (lane & 1) * 16 + (lane >> 1) : lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      int t = tile;
      const int tw_i = t % tiles_w; t /= tiles_w;
      const int th_i = t % tiles_h;
      const int b = t / tiles_h;
      const int wq = tw_i * tw + iw, hq = th_i * th + ih;
      const bool valid = wq < W && hq < H;
      mbar_wait(&t_full[acc], (it >> 1) & 1);
      tc_fence_after();
      if (has_chunk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * N + half * 32, v);
        // the bulk store that last read this staging buffer must have finished reading it
        uint8_t* sbuf = stg_raw + (it & 1) * 4096;
        if (lane == 0) bulk_wait_group_read<1>();
        __syncwarp();
        if (FUSE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y = fmaf(v[j], s1[j], s2[j]);
            v[j] = y > 0.f ? y : y * slope;
          }
        } else if (partials && valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
        }
        const uint32_t sb = smem_u32(sbuf);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(sb + sw128_off(srow, c), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
        fence_proxy_async();                                // generic-proxy stores -> visible to the TMA unit
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&mapZ, sbuf, half * 32, tw_i * tw + bw0, th_i * th + bh0, b);
          bulk_commit_group();
        }
        if (FUSE) {
          // 2x2 max-pool straight from the staged rows (the TMA store only reads them): the box holds image rows 0 / 1 of
          // 16 columns, pooled pixel pc = rows {2pc, 2pc+1, 16+2pc, 17+2pc}.  Lane -> (pc = lane / 4, two 16-byte channel
          // groups); the group order alternates with pc so that the eight lanes of a shared-memory phase hit eight
          // different swizzled chunk positions, and four lanes write 64 contiguous bytes of a pooled pixel.
          // (Round 2: the 64-shuffle butterfly + transposing store this replaces made the epilogue warps the
          // bottleneck of the kernel -- ncu: MMA warp 69 % of its samples waiting for a free accumulator.)
          const int pc = lane >> 2, jq = lane & 3;
          const int pw = ((tw_i * tw + bw0) >> 1) + pc, ph = th_i;               // th == 2: one pooled row per tile
          const bool pvalid = (tw_i * tw + bw0 + 2 * pc) < W && th_i * 2 < H;
          float* prow = F.out_p + (((long long)b * (H >> 1) + ph) * (W >> 1) + pw) * N + half * 32;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int j = jq + 4 * ((pc & 1) ^ g);
            float4 m = lds128(sb + sw128_off(2 * pc, j));
            const float4 m1 = lds128(sb + sw128_off(2 * pc + 1, j));
            const float4 m2 = lds128(sb + sw128_off(16 + 2 * pc, j));
            const float4 m3 = lds128(sb + sw128_off(17 + 2 * pc, j));
            m.x = fmaxf(fmaxf(m.x, m1.x), fmaxf(m2.x, m3.x)); m.y = fmaxf(fmaxf(m.y, m1.y), fmaxf(m2.y, m3.y));
            m.z = fmaxf(fmaxf(m.z, m1.z), fmaxf(m2.z, m3.z)); m.w = fmaxf(fmaxf(m.w, m1.w), fmaxf(m2.w, m3.w));
            if (F.round_p) { m.x = tf32_round(m.x); m.y = tf32_round(m.y); m.z = tf32_round(m.z); m.w = tf32_round(m.w); }
            if (pvalid) *reinterpret_cast<float4*>(prow + 4 * j) = m;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    if (has_chunk && lane == 0) bulk_wait_group_read<0>();   // shared memory must outlive the reads of the last stores
    if (!FUSE && partials && has_chunk) {
      const float c1 = warp_colsum32(s1, lane), c2 = warp_colsum32(s2, lane);
      float* dst = partials + (size_t)(blockIdx.x * 4 + q) * N * 2;
      dst[(half * 32 + lane) * 2 + 0] = c1;
      dst[(half * 32 + lane) * 2 + 1] = c2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

bool conv_first_tc_eligible(int Cin, int Cout) { return tc_available() && Cin * 9 <= 32 && (Cout == 32 || Cout == 64); }

// the image must tile into tw x th = 128-pixel boxes of one image (tw a power of two >= 16) and rows must be
// 16-byte multiples (TMA global strides)
bool conv_first_tc_shape_ok(int Cin, int Cout, int H, int W) {
  if (!conv_first_tc_eligible(Cin, Cout) || W < 16 || W % 4 != 0) return false;
  const int tw = pow2_floor(W < 128 ? W : 128), th = 128 / tw;
  return H >= th && (long long)H * W < (1LL << 31);
}

template <int N, bool FUSE>
static int launch_first_t(const CUtensorMap& mapX, const CUtensorMap& mapZ, const float* w, float* partials, int B, int Cin,
                          int H, int W, int tw, int th, int grid, const FirstFuse& F, cudaStream_t s) {
  using Cfg = FirstCfg<N>;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(cudaFuncSetAttribute(conv_first_tc_kernel<N, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  conv_first_tc_kernel<N, FUSE><<<grid, FIRST_THREADS, Cfg::SMEM_BYTES, s>>>(mapX, mapZ, w, partials, B, Cin, H, W, tw, th, F);
  RD_LAUNCHED();
  return 0;
}

// the fused inference epilogue pools inside a 64 x 2 pixel tile
bool conv_first_tc_fuse_ok(int Cin, int Cout, int H, int W) {
  return conv_first_tc_shape_ok(Cin, Cout, H, W) && W >= 64 && H % 2 == 0 && W % 2 == 0;
}

int launch_conv_first_tc(const float* x, const float* w, float* z, float* partials, int* n_partials, int B, int Cin,
                         int H, int W, int Cout, cudaStream_t s, const float* scale, const float* shift, const float* slope,
                         float* out_a, float* out_p, int round_p) {
  if (!conv_first_tc_shape_ok(Cin, Cout, H, W)) return fail("conv_first_tc: unsupported Cin=%d Cout=%d %dx%d", Cin, Cout, H, W);
  const bool fuse = scale != nullptr;
  if (fuse && !conv_first_tc_fuse_ok(Cin, Cout, H, W)) return fail("conv_first_tc: fused epilogue needs W >= 64, even H and W");
  const FirstFuse F{scale, shift, slope, out_a, out_p, round_p};
  const int tw = fuse ? 64 : pow2_floor(W < 128 ? W : 128), th = 128 / tw;
  // one tensor map per call: x is the caller's tensor (a different address every batch)
  CUtensorMap mapX;
  const long long dims[4] = {W, H, Cin, B};
  const long long strides[3] = {(long long)W * 4, (long long)H * W * 4, (long long)Cin * H * W * 4};
  const int box[4] = {tw + 8, th + 2, Cin, 1};
  RD_TRY(tc_encode_map(&mapX, x, 4, dims, strides, box, 2));
  // output rows (z, or the activated tensor of the fused inference epilogue) as NHWC boxes of 32 channels x 32 pixels
  CUtensorMap mapZ;
  const int bw = fuse ? 16 : (tw < 32 ? tw : 32);
  const long long zdims[4] = {Cout, W, H, B};
  const long long zstrides[3] = {(long long)Cout * 4, (long long)W * Cout * 4, (long long)H * W * Cout * 4};
  const int zbox[4] = {32, bw, 32 / bw, 1};
  RD_TRY(tc_encode_map(&mapZ, fuse ? out_a : z, 4, zdims, zstrides, zbox, 0));
  const long long tiles = (long long)cdiv(W, tw) * cdiv(H, th) * B;
  const int grid = tiles < 148 ? (int)tiles : 148;
  if (n_partials) *n_partials = grid * 4;
  switch (Cout) {
    case 32: return fuse ? launch_first_t<32, true>(mapX, mapZ, w, partials, B, Cin, H, W, tw, th, grid, F, s)
                         : launch_first_t<32, false>(mapX, mapZ, w, partials, B, Cin, H, W, tw, th, grid, F, s);
    case 64: return fuse ? launch_first_t<64, true>(mapX, mapZ, w, partials, B, Cin, H, W, tw, th, grid, F, s)
                         : launch_first_t<64, false>(mapX, mapZ, w, partials, B, Cin, H, W, tw, th, grid, F, s);
  }
  return fail("conv_first_tc: unsupported Cout=%d", Cout);
}


}  // namespace rd
