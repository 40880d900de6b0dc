// Exact k-nearest-neighbour search between projected cell matrices
// (replaces N2R::crossKnn / N2R::Knn, reference call sites R/conos.R:595,851-855).
//
// Canonical float32 arithmetic (identical in oracle/conos_oracle.c, bit for bit):
//   angular: ss = fma-chain_i(x_i*x_i), i ascending; inv = 1/sqrtf(ss) (0 if ss==0); xh_i = x_i*inv;
//            four interleaved fma chains a_j = sum_{i = j mod 4} qh_i*rh_i (i ascending, from 0) -- the
//            4-lane SIMD accumulation a CPU dot product uses -- dot = (a_0 + a_1) + (a_2 + a_3);
//            dist = 1.0f - dot
//   L2     : same four chains over (q_i - r_i)^2; dist = (a_0 + a_1) + (a_2 + a_3)  (squared Euclidean)
//   order  : (dist ascending, reference index ascending); the first k are returned.
//
// Operand panels live in HBM in "core-matrix tiled" fp32 layout
//   tiled[(row/8)][chunk = col/4][row%8][col%4],   KC = d_pad/4 chunks, d_pad in {32, 64}
// i.e. exactly the shared-memory image a K-major, no-swizzle tcgen05 (kind::tf32) operand needs, so a
// tile of R rows (R % 8 == 0) is one contiguous byte range that a single 1-D bulk copy moves.
// The panel is stored three times: the exact fp32 values in the tiled layout (knn_tc.cu, knn_simt.cu); n_pad * d_pad
// floats later the same values rounded to nearest fp16 in the swizzled row-major layout of knn_tiled16_offset (the
// kind::f16 tensor-core operand of knn_tc2.cu); 1.5 * n_pad * d_pad floats later the exact fp32 values row-major
// [row][d_pad] (the re-score of knn_tc2.cu reads whole rows: 128 / 256 contiguous bytes per candidate).
// Rows are zero-padded up to a multiple of 256.  When d < d_pad the LAST padded column of every row holds -1:
// the tensor-core kernel uses it to subtract a per-query threshold inside the MMA; every exact computation
// (re-rank, SIMT kernel) only touches the d real columns.
#pragma once
#include <vector>

#include "common.cuh"

namespace cg {

constexpr int KNN_ROW_PAD = 256;  // panel rows are padded to a multiple of this
constexpr int KNN_MAX_D = 128;    // widest panel: the tensor-core kernels serve d <= 64, the exact SIMT kernel d <= 128
// padded width of a d-column panel
__host__ __device__ inline int knn_d_pad(int d) { return d <= 32 ? 32 : (d <= 64 ? 64 : 128); }
constexpr int KNN_TC_MAXK = 32;   // largest k served by the tensor-core kernel

enum Metric : int { METRIC_ANGULAR = 0, METRIC_L2 = 1 };

struct KnnPanel {
  int n = 0;      // valid rows
  int d = 0;      // valid columns
  int d_pad = 0;  // 32, 64 or 128
  int n_pad = 0;  // multiple of KNN_ROW_PAD
  // 2.5 * n_pad * d_pad floats (not owned): the exact fp32 panel (tiled), its fp16 copy, the exact panel row-major
  float* tiled = nullptr;
  static size_t elems(int n, int d) { return round_up(n, KNN_ROW_PAD) * (size_t)knn_d_pad(d) * 5 / 2; }
};

__host__ __device__ inline size_t knn_tiled_offset(int row, int col, int d_pad) {
  return (size_t)(row >> 3) * (size_t)(8 * d_pad) + (size_t)(col >> 2) * 32 + (size_t)(row & 7) * 4 + (col & 3);
}

// element (row, col) of the fp16 copy: row-major rows of d_pad halves (64 or 128 bytes) whose 16-byte chunks are
// XOR-swizzled inside every 8-row atom -- the SWIZZLE_64B (d_pad 32) / SWIZZLE_128B (d_pad 64) K-major shared-memory
// image of tcgen05.mma, conflict-free for the tensor core's operand fetch
__host__ __device__ inline size_t knn_tiled16_offset(int row, int col, int d_pad) {
  const int chunk = col >> 3;
  const int sw = d_pad == 32 ? (chunk ^ ((row >> 1) & 3)) : (chunk ^ (row & 7));
  return (size_t)row * (size_t)d_pad + (size_t)sw * 8 + (col & 7);
}

struct KnnProblem {
  const float* q_tiled;
  const float* r_tiled;
  int nq;
  int nr;
  int* out_idx;     // [nq][k], -1 where fewer than k references exist
  float* out_dist;  // [nq][k], +inf where idx == -1
};

struct KnnItem {
  int prob;
  int qblock;  // block of 256 query rows
};

// Two-phase tensor-core kernel (knn_tc2.cu): per-problem plan and scratch
constexpr int KNN_TC2_CAND_STRIDE = 64;  // candidate slots per (query row, 64-column half)
struct KnnTc2Aux {
  int sg = 0;   // columns per group inside a tile half: 64 (whole half), 32 or 16
  int tg = 0;   // tiles per group (sg == 64 only)
  int ngt = 0;  // groups per (row, half)
  int pad = 0;
  long long row0 = 0;   // first row of this problem in the chunk's scratch arrays
  int* cand = nullptr;  // [rows][2][KNN_TC2_CAND_STRIDE] reference indices that passed the filter
  int* cnt = nullptr;   // [rows][2] survivors seen (may exceed the capacity: overflow)
  float* thr = nullptr; // [rows] threshold the filter applied to the approximate dot product
};

// Fill a tiled panel from a column-major double matrix (R layout) -- angular panels are L2-normalised.
void knn_prep_from_f64_colmajor(Context& ctx, const double* d_src, int n, int d, int ld, Metric metric,
                                KnnPanel& out);
// Same from a row-major fp32 device matrix [n][ld] (output of the projection kernels).
void knn_prep_from_f32_rowmajor(Context& ctx, const float* d_src, int n, int d, int ld, Metric metric,
                                KnnPanel& out);

// Exact SIMT kernel: any k, both metrics. One warp per query row.
void knn_simt_launch(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, int n_probs, int k,
                     int d_pad, int d, Metric metric);

// Tensor-core path for the angular metric, k <= KNN_TC_MAXK: problems that qualify go to the two-phase kernel
// (knn_tc2.cu); small problems and the 256-row blocks it hands back run through the streaming kernel of knn_tc.cu
// (TF32 tcgen05 tiles as a filter, canonical fp32 re-rank of the survivors, sorted top-k lists in shared memory).
void knn_tc_launch(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, int n_probs, int k,
                   int d_pad, int d);

// Two-phase kernel: decides whether a problem with nr references qualifies and fills the plan.
bool knn_tc2_plan(int nr, int k, KnnTc2Aux& ax);
// Runs the problems in `plist`; 256-row blocks that need the exact streaming kernel are appended to `redo`.
void knn_tc2_run(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, const std::vector<int>& plist,
                 std::vector<KnnTc2Aux>& aux, int n_probs, int k, int d_pad, int d, std::vector<KnnItem>& redo);

// Large-k tensor-core kernel (knn_tcl.cu): 32 < k <= KNN_TCL_MAXK, angular metric.  Thresholds by counting passes.
constexpr int KNN_TCL_MAXK = 3072;
struct KnnTclAux {
  long long row0 = 0;   // first row of this problem in the chunk's scratch arrays
  int* cand = nullptr;  // [rows][2][cap / 2]
  int* cnt = nullptr;   // [rows][2]
  float* thr = nullptr; // [rows]
};
bool knn_tcl_applies(int nr, int k);
// Runs the problems in `plist`; problems that need the exact SIMT kernel (candidate overflow) are appended to `redo`.
void knn_tcl_run(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, const std::vector<int>& plist,
                 int n_probs, int k, int d_pad, int d, std::vector<int>& redo);

// Dispatch by (k, metric): the tensor-core kernel whenever it applies.
void knn_launch(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, int n_probs, int k, int d_pad,
                int d, Metric metric);

}  // namespace cg
