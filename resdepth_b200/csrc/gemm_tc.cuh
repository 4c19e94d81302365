// Host-visible types of the tcgen05 GEMM kernels (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cstring>

#include "common.cuh"

namespace rd {

struct TcRowsParams {
  int N;                 // GEMM N (output channels of the layer / 4*C for the transposed conv)
  int ntaps, cchunks;    // K = ntaps * cchunks * kchunk
  int bf16;              // operands are bf16 (tcgen05 kind::f16, backward GEMMs) instead of fp32/TF32
  int kchunk;            // channels per 128-byte operand row: 32 (fp32) or 64 (bf16)
  int tw, th, tb;        // pixel box of one 128-row tile (tw*th*tb == 128)
  int tiles_w, tiles_h, tiles_b;
  int Wo, Ho, Bo;        // output pixel grid the tile coordinates index
  int coord_w, coord_h, coord_b;   // which tensor-map coordinate receives w0 / h0 / b0 (-1: none)
  int tap_off[9][4];     // per-tap coordinate offsets (dimension 0 = channel)
  int epi_mode;          // EPI_*
  int round_tf32;
  float* out;
  float* partials;
  const float* bias;
  const float* skip;
  const float* skip_scale;   // see Epilogue
  const float* skip_shift;
  const float* skip_slope;
  const float* scale;    // EPI_BNACT
  const float* shift;
  const float* slope;
  float* pool_out;
  int round_pool;
  void* out_b;           // optional bf16 copy of `out` (same indexing): operand of a later bf16 GEMM
};

struct TcRowsPlan {
  CUtensorMap mapA, mapB;
  TcRowsParams p;
  int BN = 0;
  bool halo = false;     // 3x3 halo-reuse variant (16x8 tiles)
  bool valid = false;
};

struct TcReduceParams {
  int Mrows, Ca, N;      // output rows (tap, ca), channels per tap, GEMM N
  int ntaps;
  int tw, th, tb;        // pixel box of one K step (tw*th*tb == 32)
  int tiles_w, tiles_h, tiles_b;
  int coord_w, coord_h, coord_b;
  int tap_off[9][4];
  int boxes_per_split;
  float* part;           // [splits][Mrows][N]
  // grouped loads (5-D tensor maps with the 32-channel chunk index as the slowest dimension): one TMA fetches
  // a_nch consecutive chunks of A (same tap) / all BN/32 chunks of G; 0 = one 4-D TMA per chunk
  int a_nch;
  int a_chunk_off[9];    // per tap: chunk offset of the tap inside the A tensor map (2x2 gather: (b*C)/32)
  int bf16;              // bf16 operands (64-channel chunks, kind::f16); always uses grouped loads
  // "wide" 3x3 weight-gradient mode (bf16 only): the M operand is an UNSHIFTED tensor (128-channel tiles), the N
  // tile is g_taps taps x Cs channels of the other tensor, each tap fetched at its own pixel shift -- a 128 x 192/256
  // output tile per CTA instead of 128 x 64/128 (the narrow tiles are L2-bandwidth-bound)
  int g_taps;            // taps per N tile (0: normal mode)
  int g_nch;             // 64-channel chunks per tap (Cs / 64)
  int g_off[9][2];       // per tap: (w, h) shift of the N operand
  // "self" mode (bf16, 64-row M operand, 64-column G): the N tile is [A chunk | G chunk] = 128 columns, the A box in shared
  // memory serving as both operands -- D = A^T [A | G]: the Gram matrix of A and A^T G from ONE pass over both tensors
  int g_self;
};

struct TcReducePlan {
  CUtensorMap mapA, mapG;
  TcReduceParams p;
  int BN = 0, splits = 0;
  bool valid = false;
};

bool tc_reduce_eligible(const Gather& g, int N, int bf16 = 0);
// src: tensor the gather reads (rows of the result); G: [pixels][N] matrix; part_floats: capacity of the split buffer
// bf16 != 0: src and G point to bf16 tensors
int tc_make_reduce_plan(TcReducePlan* plan, const void* src, const Gather& g, int B, const void* G, int N,
                        float* part, size_t part_floats, int bf16 = 0, int a_pitch = 0, int g_pitch = 0);
int launch_gemm_reduce_tc(const TcReducePlan& plan, cudaStream_t s);
// wide 3x3 weight gradient (bf16): D[cu][(tap, cs)] = sum_p U[p][cu] * S[p + sign*off(tap)][cs]; U, S: bf16 NHWC
// [B,H,W,Cu] / [B,H,W,Cs], Cu % 128 == 0, Cs in {64, 128}.  part: [splits][Cu][ntiles*BN] (see plan->p.N)
bool tc_reduce_wide_eligible(int Cu, int Cs, int H, int W);
// turns a bf16 1-tap plan with a 64-channel A operand and N = 64 into the "self" form (TcReduceParams::g_self): output
// [64][128] per split, columns 0..63 = A^T A, columns 64..127 = A^T G
int tc_reduce_plan_add_gram(TcReducePlan* plan, size_t part_floats);
int tc_make_reduce_plan_wide(TcReducePlan* plan, const void* U, int Cu, const void* S, int Cs, int sign, int B, int H,
                             int W, float* part, size_t part_floats);

// first encoder conv (Cin <= 3) on tcgen05: software im2col producer, 3xTF32 split (fp32-equivalent accuracy);
// x NCHW, w OIHW, z NHWC; partials: [n_partials][Cout][2] BatchNorm column sums (may be null)
bool conv_first_tc_eligible(int Cin, int Cout);
bool conv_first_tc_shape_ok(int Cin, int Cout, int H, int W);
// scale != null: fused inference epilogue -- out_a = act(z*scale+shift), out_p = its 2x2 max-pool, z is not written
bool conv_first_tc_fuse_ok(int Cin, int Cout, int H, int W);
int launch_conv_first_tc(const float* x, const float* w, float* z, float* partials, int* n_partials, int B, int Cin,
                         int H, int W, int Cout, cudaStream_t s, const float* scale = nullptr, const float* shift = nullptr,
                         const float* slope = nullptr, float* out_a = nullptr, float* out_p = nullptr, int round_p = 0);

bool tc_rows_eligible(const Gather& g, int N, int bf16 = 0);
int tc_pick_bn(int N);
// strides in BYTES; elem_bytes 4 (fp32) or 2 (bf16)
int tc_encode_map(CUtensorMap* map, const void* base, int rank, const long long* dims, const long long* strides_bytes,
                  const int* box, int swizzle_atom32, int elem_bytes = 4);
// src: activation tensor the gather reads; w_nk: packed weights [N][ntaps*C] (K contiguous)
// bf16 != 0: src and w_nk point to bf16 tensors (same logical shapes)
int tc_make_rows_plan(TcRowsPlan* plan, const void* src, const Gather& g, int B, const void* w_nk, int N, int bf16 = 0);
int launch_gemm_rows_tc(const TcRowsPlan& plan, const Epilogue& e, int* n_partials, cudaStream_t s);

}  // namespace rd
