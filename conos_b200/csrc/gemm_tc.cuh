// Dense contractions of the reduction stage on the sm_100a tensor cores.
//
//   C[m][n] = sum_k A[k][m] * B[k][n]          (both operands "K-major": k is the fast index)
//
// with FP32-grade accuracy from BF16 tensor cores: every operand is stored as hi + lo BF16 panels
// (x = hi + lo up to 2^-18 |x|) and the product is accumulated as hi*hi + hi*lo + lo*hi in FP32 in TMEM
// (3 tcgen05.mma per k-step).  Used for the per-sample gene Gram matrix X'X (src/spcov.cpp:15, K = cells) and
// for the block products M (D Q) of the eigen / common-PCA iterations (K = genes).
//
// Operand panels live in HBM pre-tiled so that a (rows x 64 k) tile is ONE contiguous block that a single 1-D
// bulk async copy (TMA engine) moves, already in the K-major no-swizzle core-matrix order tcgen05 reads:
//   panel[slab = k/64][rg = r/8][kc = (k%64)/8][ri = r%8][ki = k%8]     (bf16)
// rows padded to a multiple of 256, k padded to a multiple of 64, zero filled.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace cg {

constexpr int GEMM_ROW_PAD = 256;
constexpr int GEMM_K_PAD = 64;

struct TiledOperand {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int rows = 0, k = 0;          // valid extents
  int rows_pad = 0, k_pad = 0;  // padded extents
  static size_t elems(int rows, int k) { return round_up(rows, GEMM_ROW_PAD) * round_up(k, GEMM_K_PAD); }
};

__host__ __device__ inline size_t tiled_operand_offset(int r, int k, int rows_pad) {
  return (((size_t)(k >> 6) * (size_t)(rows_pad >> 3) + (size_t)(r >> 3)) * 8 + (size_t)((k & 63) >> 3)) * 64 +
         (size_t)(r & 7) * 8 + (size_t)(k & 7);
}

struct GemmProblem {
  const __nv_bfloat16* a_hi;  // rows = M
  const __nv_bfloat16* a_lo;
  const __nv_bfloat16* b_hi;  // rows = N
  const __nv_bfloat16* b_lo;
  int m_pad, n_pad, k_pad;
  int M, N;        // valid output extents
  void* C;         // column-major M x N, leading dimension ldc: element (m, n) at n*ldc + m
  long long ldc;
  int out_fp64;    // 1: double output, 0: float
  int symmetric = 0;  // 1: A == B (X'X): only the tiles that touch the upper triangle (m <= n) are computed
  int chunk_slabs = 0;  // > 0: K is accumulated in FP32 over chunks of this many 64-wide slabs, the chunks are summed
                        //      in the output type by the epilogue (long contractions: Gram matrices, out_fp64 = 1)
  int fp16 = 0;         // 1: the operand panels hold fp16 hi + lo pairs (11 + 11 bits) instead of bf16 (8 + 8)
};

// batched: every problem is cut into 128 x 256 output tiles, one CTA per tile
void gemm_tn_bf16x3(Context& ctx, const GemmProblem* h_probs, int n_probs);

// The same batch launched many times (the filter steps of the eigen iteration): the problem and tile tables are
// uploaded once, gemm_run only launches -- no host round trip between the steps of a sweep.
struct GemmTileRef { int prob, mt, nt; };
struct GemmPlan {
  DevBuf<GemmProblem> probs;
  DevBuf<GemmTileRef> tiles;
  unsigned n_tiles = 0;
};
void gemm_plan(Context& ctx, const GemmProblem* h_probs, int n_probs, GemmPlan& plan);
void gemm_run(Context& ctx, const GemmPlan& plan);

// ---- operand builders (zero padded, hi/lo split) -------------------------------------------------------------
// sparse cells x genes CSC -> operand with rows = genes, k = cells (op.hi/lo pre-allocated, elems(rows, k))
void tiled_from_csc(Context& ctx, const int* csc_p, const int* csc_i, const double* csc_x, int n_cells, int n_genes,
                    TiledOperand& op);
// same operand from the CSR-by-cell form (columns ascending inside a row): tile-wise, coalesced, no memset
// fp16 = true stores fp16 hi + lo pairs (values must fit fp16's range; GemmProblem::fp16 must be set for both operands)
void tiled_from_csr(Context& ctx, const int* csr_p, const int* csr_i, const double* csr_x, int n_cells, int n_genes,
                    TiledOperand& op, bool fp16 = false);
// dense column-major (k x rows, element (k, r) at r*ld + k) FP64 or FP32 -> operand
void tiled_from_colmajor_f64(Context& ctx, const double* src, int k, int rows, long long ld, TiledOperand& op);
void tiled_from_colmajor_f32(Context& ctx, const float* src, int k, int rows, long long ld, TiledOperand& op);

}  // namespace cg
