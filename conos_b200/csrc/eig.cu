#include <cooperative_groups.h>

#include <functional>

#include "eig.cuh"
#include "gemm_tc.cuh"

namespace cg {
namespace {

constexpr int EIG_THREADS = 256;

struct EigWork {
  const double* C;
  int G;
  double* Q;       // [G][B] row-major
  double* Z;       // [G][B] row-major
  double* theta;   // [B] current Ritz values
  double* V;       // out
  double* theta_out;
  int* converged;  // 1 when the leading r residuals are below tol * theta_1
  double* resid;   // optional: worst relative residual of the leading r pairs
  double* colres = nullptr;  // optional [B]: relative residual of every leading Ritz pair (the filter degree uses it)
  double* X = nullptr;       // FP64 driver: previous Chebyshev iterate [G][B]
  int* nlock = nullptr;      // FP64 driver: leading Ritz pairs that met the tolerance (locked: no longer filtered)
};

__device__ __forceinline__ double hash_unit(unsigned a, unsigned b) {
  unsigned x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return (double)x * (2.0 / 4294967296.0) - 1.0;
}

template <int B>
__global__ void eig_init_kernel(const EigWork* __restrict__ works) {
  const EigWork w = works[blockIdx.y];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (size_t)w.G * B) w.Z[t] = hash_unit((unsigned)(t / B), (unsigned)(t % B));
}

// Z = C Q.  One warp per row of C (symmetric: row r == column r, contiguous), lanes over the B columns of Q.
template <int B>
__global__ void __launch_bounds__(256) eig_matmul_kernel(const EigWork* __restrict__ works,
                                                         const int* __restrict__ deg, int step) {
  const EigWork w = works[blockIdx.y];
  if (w.converged && *w.converged != 0) return;  // locked: a converged problem is never touched again
  // inside a filter sweep: a problem whose own (shorter) sweep is over already holds its filtered block in Z --
  // it must not be overwritten because a batch mate runs a higher degree
  if (step > 0 && step > deg[blockIdx.y] + 0) return;
  __shared__ double crow[8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  const int G = w.G;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  const double* col = w.C + (size_t)(r < G ? r : 0) * G;
  for (int c0 = 0; c0 < G; c0 += 128) {
    __syncwarp();
    for (int t = lane; t < 128; t += 32) crow[warp][t] = (c0 + t < G) ? col[c0 + t] : 0.0;
    __syncwarp();
    const int lim = G - c0 < 128 ? G - c0 : 128;
    for (int t = 0; t < lim; ++t) {
      const double cv = crow[warp][t];
      const double* q = w.Q + (size_t)(c0 + t) * B;
#pragma unroll
      for (int u = 0; u < NC; ++u) acc[u] = fma(cv, q[lane + 32 * u], acc[u]);
    }
  }
  if (r < G) {
#pragma unroll
    for (int u = 0; u < NC; ++u) w.Z[(size_t)r * B + lane + 32 * u] = acc[u];
  }
}

// S[i][j] = sum_r X[r][i] * Y[r][j]  (B x B, into shared memory), all 256 threads
// Rows [r0, r1) only: a problem is shared by the CTAs of a thread-block cluster, each owning a slice of the rows
// (RowSlice below); the partial Gram matrices are summed over the cluster by cluster_sum.
template <int B>
__device__ void block_gram_rows(const double* __restrict__ X, const double* __restrict__ Y, int r0, int G,
                                double (*S)[B + 1]) {
  if constexpr (B == 32) {
    // rows split over the 8 warps (fixed assignment r = warp, warp + 8, ...), lane l owns column l of S: one coalesced
    // 256-byte load of each operand per row, x broadcast by shuffles, 32 independent FMA chains per lane; the eight
    // partial sums are then added in warp order (deterministic).
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = EIG_THREADS / 32;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    int r = r0 + warp;
    for (; r + NW < G; r += 2 * NW) {  // two rows in flight
      const double x0 = X[(size_t)r * 32 + lane], y0 = Y[(size_t)r * 32 + lane];
      const double x1 = X[(size_t)(r + NW) * 32 + lane], y1 = Y[(size_t)(r + NW) * 32 + lane];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(__shfl_sync(0xffffffffu, x0, i), y0, acc[i]);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(__shfl_sync(0xffffffffu, x1, i), y1, acc[i]);
    }
    if (r < G) {
      const double x0 = X[(size_t)r * 32 + lane], y0 = Y[(size_t)r * 32 + lane];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(__shfl_sync(0xffffffffu, x0, i), y0, acc[i]);
    }
    for (int w = 0; w < NW; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < 32; ++i) S[i][lane] = (w == 0 ? 0.0 : S[i][lane]) + acc[i];
      }
      __syncthreads();
    }
    return;
  }
  if constexpr (B == 64) {
    // four 32 x 32 blocks of S, two warps per block (even / odd rows); lane l owns column l of its block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = warp & 3, par = warp >> 2;
    const int bi = (blk >> 1) * 32, bj = (blk & 1) * 32;
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    int r = r0 + par;
    for (; r + 2 < G; r += 4) {  // two rows in flight
      const double x0 = X[(size_t)r * 64 + bi + lane], y0 = Y[(size_t)r * 64 + bj + lane];
      const double x1 = X[(size_t)(r + 2) * 64 + bi + lane], y1 = Y[(size_t)(r + 2) * 64 + bj + lane];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(__shfl_sync(0xffffffffu, x0, i), y0, acc[i]);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(__shfl_sync(0xffffffffu, x1, i), y1, acc[i]);
    }
    if (r < G) {
      const double x0 = X[(size_t)r * 64 + bi + lane], y0 = Y[(size_t)r * 64 + bj + lane];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(__shfl_sync(0xffffffffu, x0, i), y0, acc[i]);
    }
    for (int w = 0; w < 2; ++w) {
      if (par == w) {
#pragma unroll
        for (int i = 0; i < 32; ++i) S[bi + i][bj + lane] = (w == 0 ? 0.0 : S[bi + i][bj + lane]) + acc[i];
      }
      __syncthreads();
    }
    return;
  }
  constexpr int TB = B / 16;
  const int i0 = TB * (threadIdx.x / 16), j0 = TB * (threadIdx.x % 16);
  double acc[TB][TB];
#pragma unroll
  for (int a = 0; a < TB; ++a)
#pragma unroll
    for (int b = 0; b < TB; ++b) acc[a][b] = 0.0;
  for (int r = r0; r < G; ++r) {
    double xv[TB], yv[TB];
#pragma unroll
    for (int a = 0; a < TB; ++a) { xv[a] = X[(size_t)r * B + i0 + a]; yv[a] = Y[(size_t)r * B + j0 + a]; }
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
      for (int b = 0; b < TB; ++b) acc[a][b] = fma(xv[a], yv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < TB; ++a)
#pragma unroll
    for (int b = 0; b < TB; ++b) S[i0 + a][j0 + b] = acc[a][b];
  __syncthreads();
}

// dst[r][:] = src[r][:] * W  (W B x B in shared memory), one warp per row; src may equal dst
template <int B>
__device__ void rows_times_matrix(const double* __restrict__ src, double* __restrict__ dst, int r0, int G,
                                  double (*W)[B + 1], double (*rowbuf)[B]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NC = B / 32;
  for (int r = r0 + warp; r < G; r += EIG_THREADS / 32) {
#pragma unroll
    for (int u = 0; u < NC; ++u) rowbuf[warp][lane + 32 * u] = src[(size_t)r * B + lane + 32 * u];
    __syncwarp();
    double acc[NC];
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] = 0.0;
    for (int i = 0; i < B; ++i) {
      const double x = rowbuf[warp][i];
#pragma unroll
      for (int u = 0; u < NC; ++u) acc[u] = fma(x, W[i][lane + 32 * u], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < NC; ++u) dst[(size_t)r * B + lane + 32 * u] = acc[u];
    __syncwarp();
  }
  __syncthreads();
}

// Cyclic Jacobi (round-robin parallel ordering) on the symmetric A (destroyed: diagonal = eigenvalues),
// W = eigenvectors in columns.  All 256 threads.
template <int B>
__device__ void block_jacobi(double (*A)[B + 1], double (*W)[B + 1], double* cs, double* red) {
  const int tid = threadIdx.x;
  for (int t = tid; t < B * B; t += EIG_THREADS) W[t / B][t % B] = (t / B == t % B) ? 1.0 : 0.0;
  __syncthreads();
  constexpr int HP = B / 2;
  for (int sweep = 0; sweep < 30; ++sweep) {
    // convergence: largest off-diagonal against largest diagonal
    double off = 0.0, dg = 0.0;
    for (int t = tid; t < B * B; t += EIG_THREADS) {
      const double v = fabs(A[t / B][t % B]);
      if (t / B == t % B) dg = fmax(dg, v); else off = fmax(off, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
      off = fmax(off, __shfl_xor_sync(0xffffffffu, off, d));
      dg = fmax(dg, __shfl_xor_sync(0xffffffffu, dg, d));
    }
    if ((tid & 31) == 0) { red[tid >> 5] = off; red[8 + (tid >> 5)] = dg; }
    __syncthreads();
    off = 0.0; dg = 0.0;
    for (int wv = 0; wv < 8; ++wv) { off = fmax(off, red[wv]); dg = fmax(dg, red[8 + wv]); }
    __syncthreads();
    if (off <= 1e-15 * dg || dg == 0.0) break;
    for (int step = 0; step < B - 1; ++step) {
      if (tid < HP) {
        int p, q;
        if (tid == 0) { p = B - 1; q = step; }
        else { p = (step + tid) % (B - 1); q = (step - tid + (B - 1)) % (B - 1); }
        const double apq = A[p][q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300) {
          const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
          const double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + tt * tt);
          s = tt * c;
        }
        cs[3 * tid] = c; cs[3 * tid + 1] = s; cs[3 * tid + 2] = (double)(p * B + q);
      }
      __syncthreads();
      // columns: A <- A J, W <- W J
      for (int t = tid; t < HP * B; t += EIG_THREADS) {
        const int pr = t / B, i = t % B;
        const double c = cs[3 * pr], s = cs[3 * pr + 1];
        const int pq = (int)cs[3 * pr + 2];
        const int p = pq / B, q = pq % B;
        const double aip = A[i][p], aiq = A[i][q];
        A[i][p] = c * aip - s * aiq;
        A[i][q] = s * aip + c * aiq;
        const double wip = W[i][p], wiq = W[i][q];
        W[i][p] = c * wip - s * wiq;
        W[i][q] = s * wip + c * wiq;
      }
      __syncthreads();
      // rows: A <- J' A
      for (int t = tid; t < HP * B; t += EIG_THREADS) {
        const int pr = t / B, j = t % B;
        const double c = cs[3 * pr], s = cs[3 * pr + 1];
        const int pq = (int)cs[3 * pr + 2];
        const int p = pq / B, q = pq % B;
        const double apj = A[p][j], aqj = A[q][j];
        A[p][j] = c * apj - s * aqj;
        A[q][j] = s * apj + c * aqj;
      }
      __syncthreads();
    }
  }
}

// V = leading r columns of Q (already the sorted Ritz vectors), theta_out = leading r Ritz values
template <int B>
__global__ void eig_export_kernel(const EigWork* __restrict__ works, int r) {
  const EigWork w = works[blockIdx.y];
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < (size_t)w.G * r; t += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(t % w.G), c = (int)(t / w.G);
    w.V[t] = w.Q[(size_t)row * B + c];
  }
  if (blockIdx.x == 0 && threadIdx.x < r && w.theta_out) w.theta_out[threadIdx.x] = w.theta[threadIdx.x];
}

template <int B>
struct EigSmem {
  double S[B][B + 1];
  double W[B][B + 1];
  double rowbuf[EIG_THREADS / 32][B];
  double cs[3 * (B / 2)];
  double red[16];
  double ev[B];
  int perm[B];
  int flag;
  int wellcond;  // last Cholesky-QR pass: smallest pivot > 1e-4 * largest diagonal (orthogonality error ~ 1e-12)
};

// A problem may be shared by the CTAs of one thread-block cluster (EIG_CLUSTER of them): the O(G) parts -- block Gram
// products, row updates, residuals -- are split by rows, the B x B factorisations are repeated by every CTA on the
// same (cluster-summed) matrix.  One CTA per problem leaves most of the 148 SMs idle and makes the orthonormalisation
// and Rayleigh-Ritz steps the latency floor of the eigen stage whatever the number of problems.
//
// Every sum over rows is taken over the SAME EIG_CLUSTER base slices in the same order whether one CTA or a cluster
// works on the problem (a lone CTA walks the base slices one after the other): the result of a problem does not
// depend on the cluster width, i.e. on how many other problems share its launch -- which is what keeps the graph
// bit-identical across batch cuts and GPU counts.
constexpr int EIG_CLUSTER = 4;
struct RowSlice {
  int r0, r1;       // rows of this CTA
  int per;          // rows of a base slice (a multiple of 8: the warp-strided loops of the helpers start at r0 + warp)
  int G;
  unsigned rank, size;
};
__device__ __forceinline__ RowSlice row_slice(int G, int csize) {
  RowSlice sl;
  sl.size = (unsigned)csize;
  sl.rank = csize > 1 ? cooperative_groups::this_cluster().block_rank() : 0u;
  sl.per = ((G + EIG_CLUSTER - 1) / EIG_CLUSTER + 7) & ~7;
  sl.G = G;
  if (csize > 1) {
    sl.r0 = min(G, (int)sl.rank * sl.per);
    sl.r1 = min(G, sl.r0 + sl.per);
  } else {
    sl.r0 = 0;
    sl.r1 = G;
  }
  return sl;
}
// rows of base slice b
__device__ __forceinline__ int base_r0(const RowSlice& sl, int b) { return min(sl.G, b * sl.per); }
__device__ __forceinline__ int base_r1(const RowSlice& sl, int b) { return min(sl.G, (b + 1) * sl.per); }
// problem of this CTA: clusters are consecutive CTAs
__device__ __forceinline__ int cluster_problem(int csize) { return (int)blockIdx.x / csize; }

// S = X'Y over all rows of the problem: every CTA of the cluster forms the partial product of its slice in `part`
// (shared memory the caller does not need at this point), then adds the partials of all CTAs in rank order through
// distributed shared memory -- the same order in every CTA, so all of them hold bit-identical S.
template <int B>
__device__ void block_gram(const double* __restrict__ X, const double* __restrict__ Y, const RowSlice& sl,
                           double (*S)[B + 1], double (*part)[B + 1]) {
  if (sl.size == 1) {
    // the base slices one after the other, added like the cluster adds its partials: 0 + p0 + p1 + p2 + p3
    for (int b = 0; b < EIG_CLUSTER; ++b) {
      block_gram_rows<B>(X, Y, base_r0(sl, b), base_r1(sl, b), part);
      for (int t = threadIdx.x; t < B * B; t += EIG_THREADS)
        S[t / B][t % B] = (b == 0 ? 0.0 : S[t / B][t % B]) + part[t / B][t % B];
      __syncthreads();
    }
    return;
  }
  namespace cgr = cooperative_groups;
  cgr::cluster_group cluster = cgr::this_cluster();
  block_gram_rows<B>(X, Y, sl.r0, sl.r1, part);
  cluster.sync();
  for (int t = threadIdx.x; t < B * B; t += EIG_THREADS) {
    double acc = 0.0;
    for (unsigned r = 0; r < sl.size; ++r) {
      const double* remote = cluster.map_shared_rank(&part[0][0], r);
      acc += remote[(t / B) * (B + 1) + (t % B)];
    }
    S[t / B][t % B] = acc;
  }
  cluster.sync();  // nobody overwrites its partial while a neighbour still reads it
}

// sum of one shared-memory vector over the cluster, same order everywhere (v is overwritten by the total)
__device__ void cluster_sum_vector(double* v, int n, double* scratch, const RowSlice& sl) {
  if (sl.size == 1) return;
  namespace cgr = cooperative_groups;
  cgr::cluster_group cluster = cgr::this_cluster();
  cluster.sync();
  if ((int)threadIdx.x < n) {
    double acc = 0.0;
    for (unsigned r = 0; r < sl.size; ++r) acc += cluster.map_shared_rank(v, r)[threadIdx.x];
    scratch[threadIdx.x] = acc;
  }
  cluster.sync();
  if ((int)threadIdx.x < n) v[threadIdx.x] = scratch[threadIdx.x];
  __syncthreads();
}

// Orthonormalise the columns of Z into Q: Cholesky-QR (Q = Z R^-1 with Z'Z = R'R); when a pivot collapses
// (rank-deficient block) the symmetric eigen-decomposition Z'Z = U L U' is used instead, Q = Z U L^-1/2,
// and the numerically null directions are zeroed.
template <int B>
__device__ void orth_pass(const double* __restrict__ src, double* __restrict__ dst, const RowSlice& sl,
                          EigSmem<B>& sm) {
  block_gram<B>(src, src, sl, sm.S, sm.W);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) sm.flag = 0;
  // equilibrate: the columns of a filtered block differ by many orders of magnitude in norm, which says nothing
  // about their independence.  S <- D^-1 S D^-1 with D = diag(column norms) (a zero column keeps D = 1 and shows up
  // as a null direction below); the multiplier is rescaled by D^-1 at the end.
  if (tid < B) {
    const double d = sm.S[tid][tid];
    sm.ev[tid] = d > 0.0 ? 1.0 / sqrt(d) : 1.0;   // D^-1
  }
  __syncthreads();
  for (int t = tid; t < B * B; t += EIG_THREADS) sm.S[t / B][t % B] *= sm.ev[t / B] * sm.ev[t % B];
  __syncthreads();
  if (tid < B) sm.rowbuf[0][tid] = sm.ev[tid];  // D^-1 survives the fallback (which reuses ev) here
  __syncthreads();
  if (tid < 32) {
    double maxd = 0.0;
    for (int i = lane; i < B; i += 32) maxd = fmax(maxd, sm.S[i][i]);
    for (int d = 16; d > 0; d >>= 1) maxd = fmax(maxd, __shfl_xor_sync(0xffffffffu, maxd, d));
    // copy S to W (Cholesky works on W so that S survives for the fallback)
    for (int t = lane; t < B * B; t += 32) sm.W[t / B][t % B] = sm.S[t / B][t % B];
    __syncwarp();
    bool ok = maxd > 0.0;
    double minpiv = maxd;
    for (int k = 0; k < B && ok; ++k) {
      const double piv = sm.W[k][k];
      if (!(piv > 1e-13 * maxd)) { ok = false; break; }
      minpiv = fmin(minpiv, piv);
      const double dk = sqrt(piv);
      __syncwarp();
      for (int i = k + lane; i < B; i += 32) sm.W[i][k] = (i == k) ? dk : sm.W[i][k] / dk;   // L column k
      __syncwarp();
      for (int j = k + 1; j < B; ++j) {
        const double ljk = sm.W[j][k];
        for (int i = j + lane; i < B; i += 32) sm.W[i][j] -= sm.W[i][k] * ljk;
      }
      __syncwarp();
    }
    if (lane == 0) sm.wellcond = (ok && minpiv > 1e-4 * maxd) ? 1 : 0;
    if (!ok) {
      if (lane == 0) sm.flag = 1;
    } else {
      // invert L (lower) in place column by column: X = L^-1, then the multiplier is R^-1 = X'
      // solve L X = I by forward substitution, lane per column
      for (int c = lane; c < B; c += 32) {
        for (int i = 0; i < B; ++i) {
          double v = (i == c) ? 1.0 : 0.0;
          for (int t = c; t < i; ++t) v -= sm.W[i][t] * sm.S[t][c];   // S reused as X (only rows >= c touched)
          sm.S[i][c] = (i < c) ? 0.0 : v / sm.W[i][i];
        }
      }
      __syncwarp();
      for (int t = lane; t < B * B; t += 32) sm.W[t / B][t % B] = 0.0;
      __syncwarp();
      for (int t = lane; t < B * B; t += 32) {
        const int i = t / B, j = t % B;
        sm.W[j][i] = sm.S[i][j];  // W = X' = R^-1
      }
    }
  }
  __syncthreads();
  if (sm.flag) {
    block_jacobi<B>(sm.S, sm.W, sm.cs, sm.red);
    double lmax = 0.0;
    for (int i = 0; i < B; ++i) lmax = fmax(lmax, sm.S[i][i]);
    for (int t = tid; t < B * B; t += EIG_THREADS) {
      const int i = t / B, j = t % B;
      const double l = sm.S[j][j];
      sm.W[i][j] = (l > 1e-14 * lmax && l > 0.0) ? sm.W[i][j] / sqrt(l) : 0.0;
    }
    if (tid < B) sm.ev[tid] = (sm.S[tid][tid] > 1e-14 * lmax && sm.S[tid][tid] > 0.0) ? 0.0 : 1.0;
    __syncthreads();
  }
  // undo the equilibration: Q = Z D^-1 W
  for (int t = tid; t < B * B; t += EIG_THREADS) sm.W[t / B][t % B] *= sm.rowbuf[0][t / B];
  __syncthreads();
  rows_times_matrix<B>(src, dst, sl.r0, sl.r1, sm.W, sm.rowbuf);
  if (sm.flag) {
    // a numerically null direction is replaced by a fresh pseudo-random column (the next pass orthonormalises it):
    // a zero column would stay zero for the rest of the iteration
    for (size_t t = (size_t)sl.r0 * B + tid; t < (size_t)sl.r1 * B; t += EIG_THREADS) {
      const int j = (int)(t % B);
      if (sm.ev[j] != 0.0) dst[t] = hash_unit((unsigned)(t / B) + 0x9e37u, (unsigned)j + 4099u);
    }
    __syncthreads();
    if (tid == 0) sm.wellcond = 0;
  }
  // the second pass (and the caller's next kernel) reads rows other CTAs of the cluster have just written
  if (sl.size > 1) {
    __threadfence();
    cooperative_groups::this_cluster().sync();
  }
}

template <int B>
__global__ void __launch_bounds__(EIG_THREADS) eig_orth_kernel(const EigWork* __restrict__ works, int csize) {
  extern __shared__ __align__(16) unsigned char raw[];
  EigSmem<B>& sm = *reinterpret_cast<EigSmem<B>*>(raw);
  const EigWork w = works[cluster_problem(csize)];
  // a problem is locked the moment it converges: its result cannot depend on how long its batch mates take
  // (every CTA of the cluster takes the same branch: the flag is only written by the Rayleigh-Ritz kernel)
  if (w.converged && *w.converged != 0) return;
  const RowSlice sl = row_slice(w.G, csize);
  orth_pass<B>(w.Z, w.Q, sl, sm);
  // CholeskyQR2: the second pass only when the first saw an ill-conditioned block
  if (!sm.wellcond) orth_pass<B>(w.Q, w.Q, sl, sm);
}

// Rayleigh-Ritz: T = Q'Z (Z = C Q), eigen-decompose, rotate Q and Z to the Ritz basis (descending values),
// test stagnation of the leading r Ritz values.
template <int B>
__global__ void __launch_bounds__(EIG_THREADS) eig_rr_kernel(const EigWork* __restrict__ works, int r, double tol,
                                                             int finalize, int csize) {
  extern __shared__ __align__(16) unsigned char raw[];
  EigSmem<B>& sm = *reinterpret_cast<EigSmem<B>*>(raw);
  const EigWork w = works[cluster_problem(csize)];
  if (*w.converged != 0) return;  // locked (see eig_orth_kernel): Q already holds the sorted Ritz vectors
  const int tid = threadIdx.x;
  const RowSlice sl = row_slice(w.G, csize);
  block_gram<B>(w.Q, w.Z, sl, sm.S, sm.W);
  for (int t = tid; t < B * B; t += EIG_THREADS) {
    const int i = t / B, j = t % B;
    if (i < j) {
      const double v = 0.5 * (sm.S[i][j] + sm.S[j][i]);
      sm.S[i][j] = v;
      sm.S[j][i] = v;
    }
  }
  __syncthreads();
  block_jacobi<B>(sm.S, sm.W, sm.cs, sm.red);
  // sort descending (rank by counting; ties by index)
  if (tid < B) {
    const double v = sm.S[tid][tid];
    int rank = 0;
    for (int j = 0; j < B; ++j) {
      const double u = sm.S[j][j];
      rank += (u > v || (u == v && j < tid)) ? 1 : 0;
    }
    sm.perm[rank] = tid;
    sm.ev[rank] = v;
  }
  __syncthreads();
  // permute the columns of W: S reused as the permuted copy
  for (int t = tid; t < B * B; t += EIG_THREADS) {
    const int i = t / B, j = t % B;
    sm.S[i][j] = sm.W[i][sm.perm[j]];
  }
  __syncthreads();
  for (int t = tid; t < B * B; t += EIG_THREADS) sm.W[t / B][t % B] = sm.S[t / B][t % B];
  __syncthreads();
  if (tid < B && sl.rank == 0) w.theta[tid] = sm.ev[tid];
  __syncthreads();
  rows_times_matrix<B>(w.Q, w.Q, sl.r0, sl.r1, sm.W, sm.rowbuf);
  rows_times_matrix<B>(w.Z, w.Z, sl.r0, sl.r1, sm.W, sm.rowbuf);
  // residuals ||C q_j - theta_j q_j||^2 of the leading r Ritz pairs: EIG_THREADS / B partial sums per column over the
  // rows of this CTA (S is free by now), added in a fixed order, then summed over the cluster in rank order
  {
    constexpr int NP = EIG_THREADS / B;
    const int j = tid % B, part = tid / B;
    // one pass per base slice: a cluster has one slice per CTA, a lone CTA walks all of them (same sums, same order)
    const int nb = sl.size == 1 ? EIG_CLUSTER : 1;
    for (int b = 0; b < nb; ++b) {
      const int q0 = sl.size == 1 ? base_r0(sl, b) : sl.r0, q1 = sl.size == 1 ? base_r1(sl, b) : sl.r1;
      double acc = 0.0;
      if (j < r) {
        const double th = sm.ev[j];
        for (int row = q0 + part; row < q1; row += NP) {
          const double d = w.Z[(size_t)row * B + j] - th * w.Q[(size_t)row * B + j];
          acc = fma(d, d, acc);
        }
      }
      sm.S[part][j] = acc;
      __syncthreads();
      if (tid < B) {
        double tot = 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) tot += sm.S[q][tid];
        // cs holds 3 * B / 2 >= B doubles; the cluster form adds 0 + v0 + v1 + v2 + v3 in cluster_sum_vector
        sm.cs[tid] = sl.size == 1 ? ((b == 0 ? 0.0 : sm.cs[tid]) + tot) : tot;
      }
      __syncthreads();
    }
    cluster_sum_vector(sm.cs, B, &sm.rowbuf[0][0], sl);
  }
  if (tid == 0 && sl.rank == 0) {
    double worst = 0.0;
    for (int j = 0; j < r; ++j) worst = fmax(worst, sqrt(sm.cs[j]));
    const double scale = fmax(fabs(sm.ev[0]), 1e-300);
    // a block that lost rank shows up as zero / non-finite Ritz values with zero residuals: never "converged"
    const bool sane = isfinite(worst) && isfinite(sm.ev[0]) && (sm.ev[r - 1] > 0.0 || sm.ev[0] <= 0.0);
    *w.converged = finalize ? 2 : ((sane && worst <= tol * scale) ? 1 : 0);
    if (w.resid && !finalize) *w.resid = worst / scale;
    if (w.colres && !finalize)
      for (int j = 0; j < r; ++j) w.colres[j] = sqrt(sm.cs[j]) / scale;
    if (w.nlock && !finalize) {
      int nl = 0;
      while (nl < r && sane && sqrt(sm.cs[nl]) <= tol * scale) ++nl;
      *w.nlock = nl;
    }
  }
  if (finalize) {
    for (size_t t = tid; t < (size_t)w.G * r; t += EIG_THREADS) {
      const int row = (int)(t % w.G), c = (int)(t / w.G);
      if (row >= sl.r0 && row < sl.r1) w.V[t] = w.Q[(size_t)row * B + c];
    }
    if (tid < r && w.theta_out && sl.rank == 0) w.theta_out[tid] = sm.ev[tid];
  }
}

// Chebyshev coefficients of one sweep of the FP64 driver (same recurrence and degree rule as tc_coeff_kernel; FP64
// panels keep 16 digits, so a wanted column may carry 1e3 times its own size of a better-conditioned direction).
template <int B>
__global__ void eig_coeff_kernel(const EigWork* __restrict__ W, int n_probs, int degree, int r, double budget,
                                 double* __restrict__ al, double* __restrict__ be, double* __restrict__ cc,
                                 int* __restrict__ deg) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_probs) return;
  const double* theta = W[p].theta;
  const int nl = W[p].nlock ? *W[p].nlock : 0;
  const double a0 = theta[nl < B ? nl : B - 1];  // the filter is scaled at the first pair that is still iterated
  double b = theta[B - 1];
  if (!(b > 1e-12 * theta[0])) b = 1e-12 * theta[0];
  if (!(a0 > b) || nl >= r) {  // nothing left to separate (or a zero matrix): one plain power step
    for (int i = 1; i <= degree; ++i) { al[(size_t)i * n_probs + p] = 0.0; be[(size_t)i * n_probs + p] = 0.0; }
    al[(size_t)1 * n_probs + p] = a0 > 0.0 ? 1.0 / a0 : 0.0;
    cc[p] = 0.0;
    deg[p] = 1;
    return;
  }
  const double e = b / 2.0, c = b / 2.0;
  double sigma1 = e / (a0 - c), sigma = sigma1;
  cc[p] = c;
  al[(size_t)1 * n_probs + p] = sigma1 / e;
  be[(size_t)1 * n_probs + p] = 0.0;
  for (int i = 2; i <= degree; ++i) {
    const double s2 = 1.0 / (2.0 / sigma1 - sigma);
    al[(size_t)i * n_probs + p] = 2.0 * s2 / e;
    be[(size_t)i * n_probs + p] = sigma * s2;
    sigma = s2;
  }
  // degree: eps_i * T_m(x_i) / T_m(x_j) <= budget over the pairs that are still iterated
  double m_safe = (double)degree;
  for (int i = nl; i + 1 < r; ++i) {
    double xi = (theta[i] - c) / e;
    if (!(xi > 1.0)) xi = 1.0;
    const double li = log(xi + sqrt(xi * xi - 1.0));
    double eps = W[p].colres ? W[p].colres[i] : 1.0;
    if (!(eps > 1e-14)) eps = 1e-14;
    if (!(eps < 1.0)) eps = 1.0;
    const double room = log(budget / eps);
    for (int j = i + 1; j < r; ++j) {
      double xj = (theta[j] - c) / e;
      if (!(xj > 1.0)) xj = 1.0;
      const double dl = li - log(xj + sqrt(xj * xj - 1.0));
      if (dl > 1e-12) m_safe = fmin(m_safe, room / dl);
    }
  }
  const int m = (int)floor(m_safe);
  deg[p] = m < 1 ? 1 : (m > degree ? degree : m);
}

// One step of the scaled three-term recurrence on the whole block of a problem (elementwise):
//   first:   X = Q;  Y = al (Z - c Q)                         (Z = Op Q from the Rayleigh-Ritz step)
//   later:   Y = al (Z - c Q) - be X;  X = Q                  (Z = Op Q, Q the current iterate)
// Y goes to Q (next operand) or, on the problem's last step, to Z (what eig_orth_kernel orthonormalises).
template <int B>
__global__ void eig_cheb_kernel(const EigWork* __restrict__ W, const double* __restrict__ al,
                                const double* __restrict__ be, const double* __restrict__ cc,
                                const int* __restrict__ deg, int n_probs, int step) {
  const int p = blockIdx.y;
  const EigWork w = W[p];
  if (*w.converged != 0) return;
  const int m = deg[p];
  if (step > m) return;
  const double a = al[(size_t)step * n_probs + p], b = be[(size_t)step * n_probs + p], c = cc[p];
  const bool last = step == m;
  const int nl = w.nlock ? *w.nlock : 0;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < (size_t)w.G * B;
       t += (size_t)gridDim.x * blockDim.x) {
    if ((int)(t % B) < nl) {  // locked Ritz vector: kept as it is
      if (last) w.Z[t] = w.Q[t];
      continue;
    }
    const double q = w.Q[t], z = w.Z[t];
    const double y = a * (z - c * q) - (step == 1 ? 0.0 : b * w.X[t]);
    w.X[t] = q;
    if (last) w.Z[t] = y; else w.Q[t] = y;
  }
}

// The filter polynomial is huge at the locked eigenvalues, so whatever the iterated columns still hold of the locked
// directions is removed after every step:  Y <- Y - Q_L (Q_L' Y)  (Q_L = the first nlock columns of Q).
template <int B>
__global__ void __launch_bounds__(EIG_THREADS) eig_project_kernel(const EigWork* __restrict__ W,
                                                                  const int* __restrict__ deg, int step, int csize) {
  extern __shared__ __align__(16) unsigned char raw[];
  EigSmem<B>& sm = *reinterpret_cast<EigSmem<B>*>(raw);
  const int prob = cluster_problem(csize);
  const EigWork w = W[prob];
  if (*w.converged != 0) return;
  const int m = deg[prob];
  if (step > m) return;
  const int nl = w.nlock ? *w.nlock : 0;
  if (nl == 0) return;
  double* Y = step == m ? w.Z : w.Q;
  const RowSlice sl = row_slice(w.G, csize);
  block_gram<B>(w.Q, Y, sl, sm.S, sm.W);   // S[l][j] = q_l' y_j
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NC = B / 32;
  for (int g = sl.r0 + warp; g < sl.r1; g += EIG_THREADS / 32) {
#pragma unroll
    for (int u = 0; u < NC; ++u) sm.rowbuf[warp][lane + 32 * u] = w.Q[(size_t)g * B + lane + 32 * u];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      const int j = lane + 32 * u;
      if (j >= nl) {
        double acc = Y[(size_t)g * B + j];
        for (int l = 0; l < nl; ++l) acc = fma(-sm.rowbuf[warp][l], sm.S[l][j], acc);
        Y[(size_t)g * B + j] = acc;
      }
    }
    __syncwarp();
  }
}

// <<<n_probs clusters of csize CTAs>>>
template <class... KArgs, class... Args>
void launch_clustered(void (*kernel)(KArgs...), int n_probs, int csize, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_probs * csize));
  cfg.blockDim = dim3(EIG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = csize > 1 ? 1 : 0;
  CG_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
// cluster width for problems of G rows: slices below ~250 rows are not worth the cluster barriers
// and with hundreds of problems the single-CTA form already fills the machine (the repeated B x B factorisations of a
// cluster would only add work: 496 CPCA problems ran 1.5x slower clustered)
inline int eig_cluster_size(const Context& ctx, int maxG, int n_probs) {
  if (ctx.eig_cluster == 1 || ctx.eig_cluster == EIG_CLUSTER) return ctx.eig_cluster;  // tests force either form
  if (maxG < 1024) return 1;
  return n_probs <= 96 ? EIG_CLUSTER : 1;   // 1 or EIG_CLUSTER only: the base slices are EIG_CLUSTER wide
}

// Generic FP64 driver: `matmul(works)` must fill Z = Op(Q) for every problem (Q, Z row-major [G][B] on the device).
// Chebyshev-filtered subspace iteration: Rayleigh-Ritz, residual test (||C q - theta q|| <= tol * theta_1 for the
// leading r pairs -- irlba's criterion), then a polynomial filter of per-problem degree on the unwanted interval
// [0, theta_B], CholeskyQR2, repeat.  `max_iter` bounds the number of operator applications.
template <int B, class MatMul>
void run_eig_generic(Context& ctx, const int* rows, double* const* Vout, double* const* theta_out, const double* const* Cmat,
                     int n_probs, int r, double tol, int max_iter, MatMul&& matmul, int* h_converged = nullptr) {
  cudaStream_t s = ctx.stream;
  constexpr int DEGREE = 12;
  std::vector<EigWork> works(n_probs);
  size_t tot = 0;
  int maxG = 0;
  for (int p = 0; p < n_probs; ++p) {
    tot += (size_t)rows[p] * B;
    maxG = rows[p] > maxG ? rows[p] : maxG;
    CG_CHECK(rows[p] >= B, "eigen block wider than the matrix");
  }
  DevBuf<double> Q(tot), Z(tot), X(tot), theta((size_t)n_probs * B), colres((size_t)n_probs * B);
  DevBuf<double> al((size_t)(DEGREE + 1) * n_probs), be((size_t)(DEGREE + 1) * n_probs), cc((size_t)n_probs);
  DevBuf<int> conv((size_t)n_probs), deg((size_t)n_probs), nlock((size_t)n_probs);
  theta.zero(s);
  conv.zero(s);
  nlock.zero(s);
  {
    std::vector<double> ones((size_t)n_probs * B, 1.0);
    colres.upload(ones.data(), ones.size(), s);
  }
  // how far a wanted column may be swamped by a better-conditioned one before the FP64 panel loses the accuracy
  // the tolerance asks for; the Gram matrix of CholeskyQR squares the condition number, so 1e4 at most
  const double budget = fmin(1e4, fmax(1e3, tol * 1e13));
  size_t o = 0;
  for (int p = 0; p < n_probs; ++p) {
    works[p] = EigWork{Cmat ? Cmat[p] : nullptr, rows[p], Q.p + o, Z.p + o, theta.p + (size_t)p * B, Vout[p],
                       theta_out ? theta_out[p] : nullptr, conv.p + p, nullptr, colres.p + (size_t)p * B, X.p + o,
                       nlock.p + p};
    o += (size_t)rows[p] * B;
  }
  DevBuf<EigWork> dw;
  dw.upload(works.data(), works.size(), s);
  const size_t smem = sizeof(EigSmem<B>);
  CG_CUDA(cudaFuncSetAttribute(eig_orth_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CG_CUDA(cudaFuncSetAttribute(eig_rr_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CG_CUDA(cudaFuncSetAttribute(eig_project_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 gi((unsigned)(((size_t)maxG * B + 255) / 256), n_probs);
  const int csize = eig_cluster_size(ctx, maxG, n_probs);
  eig_init_kernel<B><<<gi, 256, 0, s>>>(dw.p);
  launch_clustered(eig_orth_kernel<B>, n_probs, csize, smem, s, (const EigWork*)dw.p, csize);
  ctx.stats.kernels += 2;
  std::vector<int> hc(n_probs, 0), hdeg(n_probs, 1);
  int applied = 0, outer = 0;
  const unsigned cheb_blocks = (unsigned)std::min<size_t>(((size_t)maxG * B + 255) / 256, 64);
  while (true) {
    ++outer;
    {
      StageScope sc(ctx, "reduction.eig.matmul");
      matmul(works, dw.p, maxG, nullptr, 0);
      ++applied;
    }
    {
      StageScope sc(ctx, "reduction.eig.rr");
      launch_clustered(eig_rr_kernel<B>, n_probs, csize, smem, s, (const EigWork*)dw.p, r, tol, 0, csize);
      ctx.stats.kernels++;
      conv.download(hc.data(), (size_t)n_probs, s);
      CG_CUDA(cudaStreamSynchronize(s));
    }
    bool all = true;
    for (int p = 0; p < n_probs; ++p) all = all && hc[p] != 0;
    if (all || applied >= max_iter) break;
    eig_coeff_kernel<B><<<(n_probs + 127) / 128, 128, 0, s>>>(dw.p, n_probs, DEGREE, r, budget, al.p, be.p, cc.p,
                                                             deg.p);
    ctx.stats.kernels++;
    deg.download(hdeg.data(), (size_t)n_probs, s);
    CG_CUDA(cudaStreamSynchronize(s));
    int sweep = 1;
    for (int p = 0; p < n_probs; ++p)
      if (!hc[p] && hdeg[p] > sweep) sweep = hdeg[p];
    for (int i = 1; i <= sweep; ++i) {
      if (i > 1) {
        StageScope sc(ctx, "reduction.eig.matmul");
        matmul(works, dw.p, maxG, deg.p, i);
        ++applied;
      }
      eig_cheb_kernel<B><<<dim3(cheb_blocks, n_probs), 256, 0, s>>>(dw.p, al.p, be.p, cc.p, deg.p, n_probs, i);
      launch_clustered(eig_project_kernel<B>, n_probs, csize, smem, s, (const EigWork*)dw.p, (const int*)deg.p, i, csize);
      ctx.stats.kernels += 2;
    }
    {
      StageScope sc(ctx, "reduction.eig.orth");
      // converged problems are locked: every kernel of the loop skips them
      launch_clustered(eig_orth_kernel<B>, n_probs, csize, smem, s, (const EigWork*)dw.p, csize);
      ctx.stats.kernels++;
    }
  }
  if (ctx.profile) {
    ctx.stage_ms["reduction.eig.iterations"] += applied;
    ctx.stage_ms["reduction.eig.outer_iterations"] += outer;
    int nconv = 0;
    for (int p = 0; p < n_probs; ++p) nconv += hc[p] != 0;
    ctx.stage_ms["reduction.eig.converged_problems"] += nconv;
    ctx.stage_ms["reduction.eig.problems"] += n_probs;
  }
  if (h_converged)
    for (int p = 0; p < n_probs; ++p) h_converged[p] = hc[p] != 0 ? 1 : 0;
  // the loop ends right after a Rayleigh-Ritz step: Q holds the sorted Ritz vectors
  eig_export_kernel<B><<<dim3(8, n_probs), 256, 0, s>>>(dw.p, r);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
}

template <int B>
void run_eig(Context& ctx, const EigProblem* h_probs, int n_probs, int r, double tol, int max_iter) {
  std::vector<int> rows(n_probs);
  std::vector<double*> V(n_probs), th(n_probs);
  std::vector<const double*> Cm(n_probs);
  for (int p = 0; p < n_probs; ++p) {
    rows[p] = h_probs[p].G;
    V[p] = h_probs[p].V;
    th[p] = h_probs[p].theta;
    Cm[p] = h_probs[p].C;
  }
  cudaStream_t s = ctx.stream;
  run_eig_generic<B>(ctx, rows.data(), V.data(), th.data(), Cm.data(), n_probs, r, tol, max_iter,
                     [&](std::vector<EigWork>&, const EigWork* dw, int maxG, const int* deg, int step) {
                       dim3 gm((maxG + 7) / 8, n_probs);
                       eig_matmul_kernel<B><<<gm, 256, 0, s>>>(dw, deg, step);
                       ctx.stats.kernels++;
                     });
}

// =================================================================== tensor-core path (shared Gram operator)
constexpr int TB = 32;  // block width of the tensor-core path

struct TcSampleDev {
  const double* colsum;
  int G, n_cells;
  int n_cols;          // 32 * problems of this sample
  float* X;            // column-major G x n_cols panels
  float* Y;
  float* V;
  __nv_bfloat16* Uh;   // B operand (rows = columns of the panel, k = genes), pre-zeroed padding
  __nv_bfloat16* Ul;
  int u_rows_pad;
  double* t;           // [n_cols]  m' D y per column
  int first_prob;
};
struct TcProbDev {
  int sample, col0;
  unsigned seed;  // stable identity of the problem (sample pair + side): the start block never depends on the batch
  const double* f;
  double* Q;  // [G][32] row-major
  double* Z;
  double* theta;
  int* converged;
};

// Mc = M - n m m'  (FP64; the centring is done before the bf16 split so that no cancellation happens in low precision)
__global__ void tc_center_gram_kernel(const double* __restrict__ M, const double* __restrict__ colsum, int G, double n,
                                      double* __restrict__ Mc) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * G) return;
  const int i = (int)(t % G), j = (int)(t / G);
  Mc[t] = M[t] - (colsum[i] / n) * (colsum[j] / n) * n;
}

__global__ void tc_init_kernel(const TcSampleDev* __restrict__ S, const TcProbDev* __restrict__ P) {
  const TcProbDev pd = P[blockIdx.y];
  const TcSampleDev sd = S[pd.sample];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)sd.G * TB) return;
  const unsigned g = (unsigned)(t % sd.G), j = (unsigned)(t / sd.G);
  sd.Y[(size_t)(pd.col0 + j) * sd.G + g] = (float)hash_unit(g, j + 64u * pd.seed);
}

__global__ void tc_to64_kernel(const TcSampleDev* __restrict__ S, const TcProbDev* __restrict__ P) {
  const TcProbDev pd = P[blockIdx.y];
  if (*pd.converged != 0) return;  // locked
  const TcSampleDev sd = S[pd.sample];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)sd.G * TB) return;
  const int g = (int)(t / TB), j = (int)(t % TB);
  pd.Z[t] = (double)sd.Y[(size_t)(pd.col0 + j) * sd.G + g];
}

__device__ __forceinline__ void store_u(const TcSampleDev& sd, int col, int g, float u) {
  const size_t o = tiled_operand_offset(col, g, sd.u_rows_pad);
  const __nv_bfloat16 h = __float2bfloat16_rn(u);
  sd.Uh[o] = h;
  sd.Ul[o] = __float2bfloat16_rn(u - __bfloat162float(h));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// mode 0: U = D Q (operand of the Rayleigh-Ritz product), t = m' U
// mode 1: first Chebyshev step from the Ritz pairs  X = Q, Y = a1 (Z - c Q), then U, t from Y
// mode 2: generic Chebyshev step  z = D (V - n m t) / (n-1); Ynew = al (z - c Y) - be X; X = Y; Y = Ynew; U, t
// mode 3: Z (FP64 row-major) = D (V - n m t) / (n-1)                    (the Rayleigh-Ritz product itself)
__global__ void __launch_bounds__(256) tc_column_kernel(const TcSampleDev* __restrict__ S, const TcProbDev* __restrict__ P,
                                                        const int* __restrict__ col_sample,
                                                        const int* __restrict__ col_local, int n_cols_total, int mode,
                                                        const float* __restrict__ al, const float* __restrict__ be,
                                                        const float* __restrict__ cc, const int* __restrict__ deg,
                                                        int step) {
  // one CTA per column: 256 threads share the G rows (the column sum t is a block reduction)
  __shared__ double t_part[8];
  const int lane = threadIdx.x & 31;
  const int gc = blockIdx.x;
  if (gc >= n_cols_total) return;
  const TcSampleDev sd = S[col_sample[gc]];
  const int col = col_local[gc];
  const int pidx = sd.first_prob + col / TB, j = col % TB;
  const TcProbDev pd = P[pidx];
  if (*pd.converged != 0) return;  // locked: its columns of the GEMM are idle
  if (mode == 2 && step > deg[pidx]) return;  // this problem's sweep is shorter: keep its X, Y
  const int G = sd.G;
  const double n = (double)sd.n_cells;
  const double t_old = sd.t[col];
  const float a = al ? al[pidx] : 0.f, b = be ? be[pidx] : 0.f, c = cc ? cc[pidx] : 0.f;
  double tacc = 0.0;
  for (int g = threadIdx.x; g < G; g += 256) {
    const double f = pd.f[g];
    const double m = sd.colsum ? sd.colsum[g] / n : 0.0;
    const size_t pc = (size_t)col * G + g;
    if (mode == 0) {
      const float u = (float)(f * pd.Q[(size_t)g * TB + j]);
      store_u(sd, col, g, u);
      tacc += m * (double)u;
    } else if (mode == 1) {
      const float q = (float)pd.Q[(size_t)g * TB + j];
      const float z = (float)pd.Z[(size_t)g * TB + j];
      const float y = a * (z - c * q);
      sd.X[pc] = q;
      sd.Y[pc] = y;
      const float u = (float)(f * (double)y);
      store_u(sd, col, g, u);
      tacc += m * (double)u;
    } else if (mode == 2) {
      const float z = (float)(f * ((double)sd.V[pc] - n * m * t_old) / (n - 1.0));
      const float x = sd.X[pc], y = sd.Y[pc];
      const float yn = a * (z - c * y) - b * x;
      sd.X[pc] = y;
      sd.Y[pc] = yn;
      const float u = (float)(f * (double)yn);
      store_u(sd, col, g, u);
      tacc += m * (double)u;
    } else {
      pd.Z[(size_t)g * TB + j] = f * ((double)sd.V[pc] - n * m * t_old) / (n - 1.0);
    }
  }
  if (mode != 3) {
    tacc = warp_sum_d(tacc);
    if (lane == 0) t_part[threadIdx.x >> 5] = tacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += t_part[w];
      sd.t[col] = t;
    }
  }
}

// Chebyshev coefficients of every step from the current Ritz values: unwanted interval [0, theta_{B-1}],
// scaling at theta_0 (Zhou & Saad's scaled three-term recurrence).
// The filter multiplies the component of a column along eigenvector i by T_m(x_i); the panels are fp32 and the
// GEMM operands bf16 hi+lo (16 bits), so a wanted column j must never be swamped by what is left in it of a
// better-conditioned direction i < j.  With eps_i = relative residual of Ritz pair i (how much of v_i still leaks
// into the other columns) the degree of this sweep is the largest m with eps_i * T_m(x_i) / T_m(x_j) <= 8 for all
// i < j < r: narrow wanted spectra run the full degree from the start, wide ones ramp up as the leading pairs
// converge.
__global__ void tc_coeff_kernel(const TcProbDev* __restrict__ P, int n_probs, int degree, int r,
                                const double* __restrict__ colres, float* __restrict__ al, float* __restrict__ be,
                                float* __restrict__ cc, int* __restrict__ deg) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_probs) return;
  if (*P[p].converged != 0) {  // locked: no filter steps
    deg[p] = 0;
    return;
  }
  const double a0 = P[p].theta[0];
  double b = P[p].theta[TB - 1];
  if (!(b > 1e-7 * a0)) b = 1e-7 * a0;
  if (!(a0 > 0.0)) {  // degenerate (zero matrix): identity filter
    for (int i = 1; i <= degree; ++i) { al[(size_t)i * n_probs + p] = 0.f; be[(size_t)i * n_probs + p] = 0.f; }
    cc[p] = 0.f;
    deg[p] = 1;
    return;
  }
  const double e = b / 2.0, c = b / 2.0;
  const double denom = a0 - c;
  double sigma1 = e / denom, sigma = sigma1;
  cc[p] = (float)c;
  al[(size_t)1 * n_probs + p] = (float)(sigma1 / e);
  be[(size_t)1 * n_probs + p] = 0.f;
  for (int i = 2; i <= degree; ++i) {
    const double s2 = 1.0 / (2.0 / sigma1 - sigma);
    al[(size_t)i * n_probs + p] = (float)(2.0 * s2 / e);
    be[(size_t)i * n_probs + p] = (float)(sigma * s2);
    sigma = s2;
  }
  // log growth per degree of direction j: T_m(x) ~ (x + sqrt(x^2 - 1))^m / 2 for x >= 1
  double lg[TB];
  for (int j = 0; j < r; ++j) {
    double x = (P[p].theta[j] - c) / e;
    if (!(x > 1.0)) x = 1.0;
    lg[j] = log(x + sqrt(x * x - 1.0));
  }
  double m_safe = (double)degree;
  for (int i = 0; i + 1 < r; ++i) {
    double eps = colres[(size_t)p * TB + i];
    if (!(eps > 1e-7)) eps = 1e-7;
    if (!(eps < 1.0)) eps = 1.0;
    const double room = log(8.0 / eps);
    for (int j = i + 1; j < r; ++j) {
      const double dl = lg[i] - lg[j];
      if (dl > 1e-12) m_safe = fmin(m_safe, room / dl);
    }
  }
  int m = (int)floor(m_safe);
  deg[p] = m < 1 ? 1 : (m > degree ? degree : m);
}

}  // namespace

void top_eigenvectors_batched(Context& ctx, const EigProblem* h_probs, int n_probs, int r, double tol, int max_iter) {
  if (n_probs == 0) return;
  CG_CHECK(r >= 1 && r <= 56, "1..56 eigenvectors per problem");
  if (r <= 24) run_eig<32>(ctx, h_probs, n_probs, r, tol, max_iter);
  else run_eig<64>(ctx, h_probs, n_probs, r, tol, max_iter);
}


void top_eigenvectors_operator(Context& ctx, int n_rows, int r, double tol, int max_iter,
                               const std::function<void(const double* Q, double* Z, int B)>& apply, double* V,
                               double* theta) {
  CG_CHECK(r >= 1 && r <= 56, "1..56 eigenvectors per problem");
  double* Vp[1] = {V};
  double* tp[1] = {theta};
  auto mm = [&](std::vector<EigWork>& works, const EigWork*, int, const int*, int) {
    apply(works[0].Q, works[0].Z, r <= 24 ? 32 : 64);
  };
  if (r <= 24) run_eig_generic<32>(ctx, &n_rows, Vp, tp, nullptr, 1, r, tol, max_iter, mm);
  else run_eig_generic<64>(ctx, &n_rows, Vp, tp, nullptr, 1, r, tol, max_iter, mm);
}

// =================================================================== pair operators  S = wa Da Mc_a Da + wb Db Mc_b Db
// (the cell-weighted mean covariance whose leading eigenvectors start the common-PCA iteration, R/conos.R:180-185).
// The FP64 driver above iterates; the operator is applied for ALL pairs at once through one bf16x3 tensor-core GEMM
// per sample against its centred Gram matrix: a pair contributes a block of B columns to each of its two samples.
namespace {

struct PairOpDev {
  const int* cols[2];
  const double* f[2];
  double w[2];
  __nv_bfloat16* Uh[2];  // this pair's B columns inside the samples' operands (tiled, rows = columns of U)
  __nv_bfloat16* Ul[2];
  int col0[2];           // first column of the pair's block inside the sample's operand / product
  int u_rows_pad[2];
  const float* V[2];     // the samples' products, column-major G_s x ncols_s
  int Gs[2];
  int G;                 // od genes of the pair (<= rows of its Q block)
};

// U_s[:, col0 + j] = D_s Q[:, j] for both samples of every pair (zero outside the pair's gene set: never written)
template <int B>
__global__ void pairop_scatter_kernel(const EigWork* __restrict__ W, const PairOpDev* __restrict__ P,
                                      const int* __restrict__ deg, int step) {
  const EigWork w = W[blockIdx.y];
  if (*w.converged != 0) return;
  if (step > 0 && step > deg[blockIdx.y]) return;
  const PairOpDev pd = P[blockIdx.y];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)pd.G * B) return;
  const int g = (int)(t / B), j = (int)(t % B);
  const double q = w.Q[t];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const float u = (float)(pd.f[side][g] * q);
    const size_t o = tiled_operand_offset(pd.col0[side] + j, pd.cols[side][g], pd.u_rows_pad[side]);
    const __nv_bfloat16 h = __float2bfloat16_rn(u);
    pd.Uh[side][o] = h;
    pd.Ul[side][o] = __float2bfloat16_rn(u - __bfloat162float(h));
  }
}

// Z[g][j] = wa fa[g] V_a[cols_a[g], col0_a + j] + wb fb[g] V_b[cols_b[g], col0_b + j]; padding rows of Q stay zero
template <int B>
__global__ void pairop_gather_kernel(const EigWork* __restrict__ W, const PairOpDev* __restrict__ P,
                                     const int* __restrict__ deg, int step) {
  const EigWork w = W[blockIdx.y];
  if (*w.converged != 0) return;
  if (step > 0 && step > deg[blockIdx.y]) return;
  const PairOpDev pd = P[blockIdx.y];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)w.G * B) return;
  const int g = (int)(t / B), j = (int)(t % B);
  double z = 0.0;
  if (g < pd.G) {
#pragma unroll
    for (int side = 0; side < 2; ++side)
      z += pd.w[side] * pd.f[side][g] *
           (double)pd.V[side][(size_t)(pd.col0[side] + j) * pd.Gs[side] + pd.cols[side][g]];
  }
  w.Z[t] = z;
}

template <int B>
void run_pair_sum_tc(Context& ctx, const PairOpSample* samples, int n_samples, const PairOpProblem* probs, int n_probs,
                     int r, double tol, int max_iter, int* h_converged) {
  cudaStream_t s = ctx.stream;
  // columns of every sample
  std::vector<int> ncols(n_samples, 0);
  std::vector<PairOpDev> hp(n_probs);
  for (int p = 0; p < n_probs; ++p) {
    const int sidx[2] = {probs[p].a, probs[p].b};
    for (int side = 0; side < 2; ++side) {
      CG_CHECK(sidx[side] >= 0 && sidx[side] < n_samples, "bad sample index");
      hp[p].col0[side] = ncols[sidx[side]];
      ncols[sidx[side]] += B;
    }
  }
  std::vector<DevBuf<__nv_bfloat16>> Mh(n_samples), Ml(n_samples), Uh(n_samples), Ul(n_samples);
  std::vector<DevBuf<float>> V(n_samples);
  std::vector<TiledOperand> Mop(n_samples);
  std::vector<int> u_rows_pad(n_samples, 0);
  std::vector<GemmProblem> gprobs;
  for (int si = 0; si < n_samples; ++si) {
    if (ncols[si] == 0) continue;
    const int G = samples[si].G;
    Mh[si].alloc(TiledOperand::elems(G, G));
    Ml[si].alloc(TiledOperand::elems(G, G));
    Mop[si].hi = Mh[si].p;
    Mop[si].lo = Ml[si].p;
    tiled_from_colmajor_f64(ctx, samples[si].Mc, G, G, G, Mop[si]);
    u_rows_pad[si] = (int)round_up(ncols[si], GEMM_ROW_PAD);
    const size_t ue = (size_t)u_rows_pad[si] * round_up(G, GEMM_K_PAD);
    Uh[si].alloc(ue);
    Ul[si].alloc(ue);
    Uh[si].zero(s);
    Ul[si].zero(s);
    V[si].alloc((size_t)G * ncols[si]);
    gprobs.push_back(GemmProblem{Mop[si].hi, Mop[si].lo, Uh[si].p, Ul[si].p, Mop[si].rows_pad, u_rows_pad[si],
                                 Mop[si].k_pad, G, ncols[si], V[si].p, (long long)G, 0});
  }
  GemmPlan gplan;
  gemm_plan(ctx, gprobs.data(), (int)gprobs.size(), gplan);
  std::vector<int> rows(n_probs);
  std::vector<double*> Vout(n_probs), th(n_probs);
  for (int p = 0; p < n_probs; ++p) {
    const PairOpProblem& pb = probs[p];
    const int sidx[2] = {pb.a, pb.b};
    for (int side = 0; side < 2; ++side) {
      const int si = sidx[side];
      hp[p].cols[side] = pb.cols[side];
      hp[p].f[side] = pb.f[side];
      hp[p].w[side] = pb.w[side];
      hp[p].Uh[side] = Uh[si].p;
      hp[p].Ul[side] = Ul[si].p;
      hp[p].u_rows_pad[side] = u_rows_pad[si];
      hp[p].V[side] = V[si].p;
      hp[p].Gs[side] = samples[si].G;
    }
    hp[p].G = pb.G;
    rows[p] = pb.rows;
    Vout[p] = pb.V;
    th[p] = pb.theta;
  }
  DevBuf<PairOpDev> dP;
  dP.upload(hp.data(), hp.size(), s);
  run_eig_generic<B>(
      ctx, rows.data(), Vout.data(), th.data(), nullptr, n_probs, r, tol, max_iter,
      [&](std::vector<EigWork>&, const EigWork* dw, int maxG, const int* deg, int step) {
        dim3 grid((unsigned)(((size_t)maxG * B + 255) / 256), n_probs);
        pairop_scatter_kernel<B><<<grid, 256, 0, s>>>(dw, dP.p, deg, step);
        gemm_run(ctx, gplan);
        pairop_gather_kernel<B><<<grid, 256, 0, s>>>(dw, dP.p, deg, step);
        ctx.stats.kernels += 2;
      },
      h_converged);
}

}  // namespace

void top_eigenvectors_pair_sum_tc(Context& ctx, const PairOpSample* samples, int n_samples, const PairOpProblem* probs,
                                  int n_probs, int r, double tol, int max_iter, int* h_converged) {
  if (n_probs == 0) return;
  CG_CHECK(r >= 1 && r <= 56, "1..56 eigenvectors per problem");
  if (r <= 24) run_pair_sum_tc<32>(ctx, samples, n_samples, probs, n_probs, r, tol, max_iter, h_converged);
  else run_pair_sum_tc<64>(ctx, samples, n_samples, probs, n_probs, r, tol, max_iter, h_converged);
}

void top_eigenvectors_shared_gram(Context& ctx, const TcEigSample* samples, int n_samples,
                                  const TcEigProblem* probs, int n_probs, int r, double tol, int max_outer,
                                  int degree, int* h_converged) {
  if (n_probs == 0) return;
  CG_CHECK(r >= 1 && r <= 24, "tensor-core eigen path: 1..24 eigenvectors per problem");
  cudaStream_t s = ctx.stream;
  constexpr int B = TB;
  // problems grouped by sample (stable)
  std::vector<std::vector<int>> by_sample(n_samples);
  for (int p = 0; p < n_probs; ++p) {
    CG_CHECK(probs[p].sample >= 0 && probs[p].sample < n_samples, "bad sample index");
    by_sample[probs[p].sample].push_back(p);
  }
  std::vector<int> order;  // internal problem order
  std::vector<TcSampleDev> hs(n_samples);
  std::vector<TcProbDev> hp;
  std::vector<int> col_sample, col_local;
  size_t panel_elems = 0, u_elems = 0, qz_elems = 0, n_cols_total = 0;
  for (int si = 0; si < n_samples; ++si) {
    const int nc = B * (int)by_sample[si].size();
    const int G = nc > 0 ? samples[si].G : 0;  // samples without a problem (not resident on this rank) take no space
    CG_CHECK(nc == 0 || G >= B, "tensor-core eigen path needs at least 32 genes");
    hs[si].colsum = nullptr;  // the operand is the centred Gram matrix: no rank-one correction left
    hs[si].G = G;
    hs[si].n_cells = samples[si].n_cells;
    hs[si].n_cols = nc;
    hs[si].u_rows_pad = (int)round_up(nc > 0 ? nc : 1, GEMM_ROW_PAD);
    hs[si].first_prob = (int)order.size();
    panel_elems += (size_t)G * nc;
    u_elems += (size_t)hs[si].u_rows_pad * round_up(G, GEMM_K_PAD);
    for (int t = 0; t < (int)by_sample[si].size(); ++t) {
      order.push_back(by_sample[si][t]);
      qz_elems += (size_t)G * B;
    }
    for (int c = 0; c < nc; ++c) { col_sample.push_back(si); col_local.push_back(c); }
    n_cols_total += nc;
  }
  DevBuf<float> X(panel_elems), Y(panel_elems), V(panel_elems);
  DevBuf<__nv_bfloat16> Uh(u_elems), Ul(u_elems);
  DevBuf<double> tvec(n_cols_total), Q(qz_elems), Z(qz_elems), theta((size_t)n_probs * B);
  DevBuf<int> conv((size_t)n_probs);
  DevBuf<double> resid((size_t)n_probs), colres((size_t)n_probs * B);
  resid.zero(s);
  {
    std::vector<double> ones((size_t)n_probs * B, 1.0);  // nothing is known about the random start
    colres.upload(ones.data(), ones.size(), s);
  }
  Uh.zero(s);
  Ul.zero(s);
  tvec.zero(s);
  theta.zero(s);
  conv.zero(s);
  // per-sample Gram operands (bf16 hi / lo, tiled)
  std::vector<DevBuf<__nv_bfloat16>> Mh(n_samples), Ml(n_samples);
  std::vector<TiledOperand> Mop(n_samples);
  {
    size_t po = 0, uo = 0, qo = 0, co = 0;
    int pi = 0;
    for (int si = 0; si < n_samples; ++si) {
      const int G = hs[si].G, nc = hs[si].n_cols;
      hs[si].X = X.p + po; hs[si].Y = Y.p + po; hs[si].V = V.p + po;
      hs[si].Uh = Uh.p + uo; hs[si].Ul = Ul.p + uo;
      hs[si].t = tvec.p + co;
      po += (size_t)G * nc;
      uo += (size_t)hs[si].u_rows_pad * round_up(G, GEMM_K_PAD);
      co += nc;
      for (int t = 0; t < (int)by_sample[si].size(); ++t, ++pi) {
        const TcEigProblem& ep = probs[order[pi]];
        hp.push_back(TcProbDev{si, t * B, ep.seed, ep.f, Q.p + qo, Z.p + qo, theta.p + (size_t)pi * B, conv.p + pi});
        qo += (size_t)G * B;
      }
      if (nc == 0) continue;
      Mh[si].alloc(TiledOperand::elems(G, G));
      Ml[si].alloc(TiledOperand::elems(G, G));
      Mop[si].hi = Mh[si].p;
      Mop[si].lo = Ml[si].p;
      {
        DevBuf<double> Mc((size_t)G * G);
        tc_center_gram_kernel<<<(unsigned)(((size_t)G * G + 255) / 256), 256, 0, s>>>(
            samples[si].gram, samples[si].colsum, G, (double)samples[si].n_cells, Mc.p);
        tiled_from_colmajor_f64(ctx, Mc.p, G, G, G, Mop[si]);
        ctx.stats.kernels++;
      }
    }
  }
  DevBuf<TcSampleDev> dS;
  DevBuf<TcProbDev> dP;
  DevBuf<int> dcs, dcl;
  dS.upload(hs.data(), hs.size(), s);
  dP.upload(hp.data(), hp.size(), s);
  dcs.upload(col_sample.data(), col_sample.size(), s);
  dcl.upload(col_local.data(), col_local.size(), s);
  DevBuf<float> al((size_t)(degree + 1) * n_probs), be((size_t)(degree + 1) * n_probs), cc((size_t)n_probs);
  // EigWork views for the FP64 orthonormalisation / Rayleigh-Ritz kernels
  std::vector<EigWork> works(n_probs);
  for (int pi = 0; pi < n_probs; ++pi)
    works[pi] = EigWork{nullptr, hs[hp[pi].sample].G, hp[pi].Q, hp[pi].Z, hp[pi].theta, probs[order[pi]].V,
                        probs[order[pi]].theta, hp[pi].converged, resid.p + pi, colres.p + (size_t)pi * B};
  DevBuf<EigWork> dw;
  dw.upload(works.data(), works.size(), s);
  std::vector<GemmProblem> gprobs;
  int maxG = 0;
  for (int si = 0; si < n_samples; ++si) {
    if (hs[si].n_cols == 0) continue;
    const int G = hs[si].G;
    maxG = G > maxG ? G : maxG;
    gprobs.push_back(GemmProblem{Mop[si].hi, Mop[si].lo, hs[si].Uh, hs[si].Ul, Mop[si].rows_pad, hs[si].u_rows_pad,
                                 Mop[si].k_pad, G, hs[si].n_cols, hs[si].V, (long long)G, 0});
  }
  GemmPlan gplan;  // the same batch at every step of every sweep
  gemm_plan(ctx, gprobs.data(), (int)gprobs.size(), gplan);
  const size_t smem = sizeof(EigSmem<B>);
  CG_CUDA(cudaFuncSetAttribute(eig_orth_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CG_CUDA(cudaFuncSetAttribute(eig_rr_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned colblocks = (unsigned)n_cols_total;  // one CTA per column
  const int csize = eig_cluster_size(ctx, maxG, n_probs);
  DevBuf<int> deg((size_t)n_probs);
  auto column = [&](int mode, const float* a, const float* b, const float* c, int step) {
    tc_column_kernel<<<colblocks, 256, 0, s>>>(dS.p, dP.p, dcs.p, dcl.p, (int)n_cols_total, mode, a, b, c, deg.p,
                                               step);
    ctx.stats.kernels++;
  };
  {
    dim3 gi((unsigned)(((size_t)maxG * B + 255) / 256), n_probs);
    tc_init_kernel<<<gi, 256, 0, s>>>(dS.p, dP.p);
    ctx.stats.kernels++;
  }
  std::vector<int> hc(n_probs), hdeg(n_probs);
  int outer_done = 0, gemm_steps = 0;
  for (int outer = 0; outer <= max_outer; ++outer) {
    outer_done = outer;
    // orthonormal basis of the filtered block
    dim3 g64((unsigned)(((size_t)maxG * B + 255) / 256), n_probs);
    {
      StageScope sc(ctx, "reduction.eig.orth");
      tc_to64_kernel<<<g64, 256, 0, s>>>(dS.p, dP.p);
      launch_clustered(eig_orth_kernel<B>, n_probs, csize, smem, s, (const EigWork*)dw.p, csize);
      ctx.stats.kernels += 2;
    }
    // Rayleigh-Ritz: Z = C Q through the shared Gram GEMM
    {
      StageScope sc(ctx, "reduction.eig.rr_product");
      column(0, nullptr, nullptr, nullptr, 0);
      gemm_run(ctx, gplan);
      column(3, nullptr, nullptr, nullptr, 0);
    }
    {
      StageScope sc(ctx, "reduction.eig.rr");
      launch_clustered(eig_rr_kernel<B>, n_probs, csize, smem, s, (const EigWork*)dw.p, r, tol, 0, csize);
      ctx.stats.kernels++;
    }
    // Chebyshev filter of the next sweep: per-problem degree (<= degree) from the current residuals, see
    // tc_coeff_kernel.  Launched before the host looks at the convergence flags so that one round trip per sweep
    // brings both the flags and the degrees.
    tc_coeff_kernel<<<(n_probs + 127) / 128, 128, 0, s>>>(dP.p, n_probs, degree, r, colres.p, al.p, be.p, cc.p, deg.p);
    ctx.stats.kernels++;
    conv.download(hc.data(), (size_t)n_probs, s);
    deg.download(hdeg.data(), (size_t)n_probs, s);
    CG_CUDA(cudaStreamSynchronize(s));
    bool all = true;
    for (int p = 0; p < n_probs; ++p) all = all && hc[p] != 0;
    if (all || outer == max_outer) break;
    int sweep = 1;
    for (int p = 0; p < n_probs; ++p)
      if (!hc[p] && hdeg[p] > sweep) sweep = hdeg[p];
    gemm_steps += sweep;
    StageScope sc(ctx, "reduction.eig.filter");
    column(1, al.p + (size_t)1 * n_probs, nullptr, cc.p, 1);
    for (int i = 2; i <= sweep; ++i) {
      gemm_run(ctx, gplan);
      column(2, al.p + (size_t)i * n_probs, be.p + (size_t)i * n_probs, cc.p, i);
    }
  }
  if (ctx.profile) {
    ctx.stage_ms["reduction.eig.outer_iterations"] += outer_done;
    ctx.stage_ms["reduction.eig.filter_gemms"] += gemm_steps;
    std::vector<double> hr(n_probs);
    resid.download(hr.data(), (size_t)n_probs, s);
    CG_CUDA(cudaStreamSynchronize(s));
    double wr = 0.0;
    for (double v : hr) wr = v > wr ? v : wr;
    ctx.stage_ms["reduction.eig.worst_rel_residual_x1e9"] = wr * 1e9;
    int nconv = 0;
    for (int p = 0; p < n_probs; ++p) nconv += hc[p] != 0;
    ctx.stage_ms["reduction.eig.converged_problems"] += nconv;
    ctx.stage_ms["reduction.eig.problems"] += n_probs;
  }
  if (h_converged)
    for (int pi = 0; pi < n_probs; ++pi) h_converged[order[pi]] = hc[pi] != 0 ? 1 : 0;
  // the loop always ends right after a Rayleigh-Ritz step: Q holds the sorted Ritz vectors
  eig_export_kernel<B><<<dim3(8, n_probs), 256, 0, s>>>(dw.p, r);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
}

}  // namespace cg
