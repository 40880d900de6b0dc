// General neighbour matching (any k1, mNN or NN) and the k1 > k hub-edge reduction.
//
// Matching (R/conos.R:859-874): both neighbour lists of a pair become triples (column j in B, row i in A, source),
// bucketed by column, sorted by (row, source) inside the column, and adjacent triples of the same (i, j) are
// merged: mNN keeps pairs present in both lists with sqrt(s_ij * s_ji), NN keeps the union with the mean
// (a missing direction counts 0).  Entries below min.similarity disappear.
//
// Hub paring (R/conos.R:890-933, src/edgeFilter.cpp:23-60): per pair one CTA walks the columns (then the rows)
// in order of decreasing live degree (stable), and for a line with more than k live entries removes the entries
// whose opposite vertex has the highest current degree (stable descending sort; arma::sort_index is unstable in
// the reference, the tie rule is fixed to order of appearance like the oracle) while that degree exceeds klow,
// updating the degrees as it goes.  The walk is inherently sequential inside a pair; pairs run concurrently.
#include <cmath>

#include "hub.cuh"
#include "prims.cuh"

namespace cg {
namespace {

typedef unsigned long long u64;
inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ double sim(float d, int metric, double cor_base, double l2_sigma) {
  if (metric == METRIC_ANGULAR) return fmax(0.0, cor_base - (double)d);
  return exp(-(double)d / l2_sigma);
}

// ------------------------------------------------------------------ matching by sort-merge (one problem)
__global__ void triple_count_kernel(const int* __restrict__ idx_ab, int nA, const int* __restrict__ idx_ba, int nB,
                                    int k1, int* __restrict__ cnt) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nba = (long long)nB * k1, nab = (long long)nA * k1;
  if (t < nba) {
    if (idx_ba[t] >= 0) atomicAdd(&cnt[t / k1], 1);
  } else if (t < nba + nab) {
    const int j = idx_ab[t - nba];
    if (j >= 0) atomicAdd(&cnt[j], 1);
  }
}

__global__ void triple_scatter_kernel(const int* __restrict__ idx_ab, int nA, const int* __restrict__ idx_ba, int nB,
                                      int k1, const long long* __restrict__ col_off, int* __restrict__ cursor,
                                      u64* __restrict__ keys) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nba = (long long)nB * k1, nab = (long long)nA * k1;
  int i, j;
  if (t < nba) {
    i = idx_ba[t];
    j = (int)(t / k1);
  } else if (t < nba + nab) {
    j = idx_ab[t - nba];
    i = (int)((t - nba) / k1);
  } else {
    return;
  }
  if (i < 0 || j < 0) return;
  const int slot = atomicAdd(&cursor[j], 1);
  keys[col_off[j] + slot] = ((u64)(unsigned)i << 32) | (u64)(unsigned)t;  // BA triples (small t) sort first
}

// head of every run of equal (column, row); weight of the merged entry; keep flag
__global__ void triple_merge_kernel(const u64* __restrict__ keys, const long long* __restrict__ col_off, int nB,
                                    int nA, int k1, const float* __restrict__ dist_ab,
                                    const float* __restrict__ dist_ba, SimParams sp, int* __restrict__ keep,
                                    double* __restrict__ wgt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nB) return;
  const long long nba = (long long)nB * k1;
  const long long b = col_off[j], e = col_off[j + 1];
  long long t = b;
  while (t < e) {
    const unsigned row = (unsigned)(keys[t] >> 32);
    double s_ba = 0.0, s_ab = 0.0;
    bool has_ba = false, has_ab = false;
    long long u = t;
    while (u < e && (unsigned)(keys[u] >> 32) == row) {
      const long long id = (long long)(keys[u] & 0xffffffffull);
      if (id < nba) { s_ba = sim(dist_ba[id], sp.metric, sp.cor_base, sp.l2_sigma); has_ba = true; }
      else { s_ab = sim(dist_ab[id - nba], sp.metric, sp.cor_base, sp.l2_sigma); has_ab = true; }
      keep[u] = 0;
      ++u;
    }
    double w;
    if (sp.matching == 0) w = (has_ba && has_ab) ? sqrt(s_ba * s_ab) : 0.0;  // drop0(n21 * t(n12)); sqrt
    else w = (s_ba + s_ab) / 2.0;                                              // drop0(n21 + t(n12)) / 2
    keep[t] = (w > 0.0 && !(w < sp.min_similarity)) ? 1 : 0;
    wgt[t] = w;
    t = u;
  }
}

__global__ void triple_emit_kernel(const u64* __restrict__ keys, const long long* __restrict__ col_off, int nB,
                                   const int* __restrict__ keep, const long long* __restrict__ pos,
                                   const double* __restrict__ wgt, int* __restrict__ o_row, int* __restrict__ o_col,
                                   double* __restrict__ o_w, int* __restrict__ col_cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nB) return;
  int c = 0;
  for (long long t = col_off[j]; t < col_off[j + 1]; ++t) {
    if (!keep[t]) continue;
    const long long o = pos[t];
    o_row[o] = (int)(keys[t] >> 32);
    o_col[o] = j;
    o_w[o] = wgt[t];
    ++c;
  }
  col_cnt[j] = c;
}

// ------------------------------------------------------------------ hub paring
struct PareProblem {
  // matrix entries in CSC order (by column, rows ascending): row / col / x / alive
  const int* row;
  const int* col;
  int* alive;
  long long nnz;
  int nA, nB;
  // line structures: columns (CSC) and rows (CSR permutation)
  const long long* col_ptr;  // [nB + 1]
  const long long* row_ptr;  // [nA + 1]
  const int* row_perm;       // entry ids ordered by (row, col)
  int* rowN;                 // [nA] live degree of every row
  int* colN;                 // [nB] live degree of every column
  const u64* order;          // line visiting order for the current side (low 32 bits = line id)
  u64* scratch;              // per-problem sort scratch (max line length, power of two padded)
  int k, klow;               // parameters of this pass (k < 0: skip)
};

__device__ void block_bitonic(u64* s, int np, int tid, int nthreads) {
  for (int size = 2; size <= np; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (np >> 1); t += nthreads) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        u64 a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
}

constexpr int PARE_THREADS = 512;
constexpr int PARE_SMEM_KEYS = 4096;
constexpr int PARE_BINS = 4096;      // counting select: degrees below this many are binned in shared memory
constexpr int PARE_LINE_MAX = 4096;  // ... for lines of at most this many stored entries

// The entries a line loses are the m live ones with the highest opposite degree (ties: lowest position first).  That is
// a selection, not a sort: degrees are small integers, so a histogram gives the cut degree d* (everything above it goes,
// of the entries AT d* the first m - #above in position order go), and one ordered pass over the line removes them.
// ~10 CTA barriers per line instead of the ~80 of a 4096-key bitonic sort; same result, tie rule included.
// Returns false when the line does not qualify (too long / degrees too large): the caller sorts instead.
__device__ bool pare_line_by_counting(const PareProblem& pb, int side, long long b, int len, int* __restrict__ N,
                                      int* s_deg, int* s_hist, int* s_misc) {
  if (len > PARE_LINE_MAX) return false;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = PARE_THREADS / 32;
  // ---- degrees of the live entries (dead: -1), their number and maximum
  __syncthreads();               // the previous line's readers of s_misc are done
  if (tid < 4) s_misc[tid] = 0;  // [0] live, [1] max degree, [2] above klow, [3] scratch
  __syncthreads();
  int live = 0, dmax = 0, above = 0;
  for (int t = tid; t < len; t += PARE_THREADS) {
    const int ent = side == 0 ? (int)(b + t) : pb.row_perm[b + t];
    int d = -1;
    if (pb.alive[ent]) {
      d = N[side == 0 ? pb.row[ent] : pb.col[ent]];
      ++live;
      dmax = max(dmax, d);
      above += d > pb.klow ? 1 : 0;
    }
    s_deg[t] = d;
  }
  for (int o = 16; o > 0; o >>= 1) {
    live += __shfl_xor_sync(0xffffffffu, live, o);
    above += __shfl_xor_sync(0xffffffffu, above, o);
    dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if (lane == 0) {
    atomicAdd(&s_misc[0], live);
    atomicAdd(&s_misc[2], above);
    atomicMax(&s_misc[1], dmax);
  }
  __syncthreads();
  live = s_misc[0];
  dmax = s_misc[1];
  above = s_misc[2];
  if (dmax >= PARE_BINS) return false;  // uniform: every thread read the same shared values
  int m = live - pb.k;
  if (above < m) m = above;
  if (m <= 0) return true;  // nothing to remove (also uniform)
  // ---- histogram of the degrees, cut degree d* = largest d with #(deg >= d) >= m
  const int nb = dmax + 1;
  for (int t = tid; t < nb; t += PARE_THREADS) s_hist[t] = 0;
  __syncthreads();
  for (int t = tid; t < len; t += PARE_THREADS)
    if (s_deg[t] >= 0) atomicAdd(&s_hist[s_deg[t]], 1);
  __syncthreads();
  // suffix sums: thread t owns the bins [t * per, (t + 1) * per) counted from the TOP (bin index dmax - i)
  const int per = (nb + PARE_THREADS - 1) / PARE_THREADS;
  int mine = 0;
  for (int i = tid * per; i < min(nb, (tid + 1) * per); ++i) mine += s_hist[dmax - i];
  // exclusive prefix of `mine` over the threads (warp scan + warp totals)
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __shared__ int s_wtot[NW];
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w2 = 0; w2 < warp; ++w2) base += s_wtot[w2];
  int run = base + incl - mine;  // entries with degree above this thread's first bin
  for (int i = tid * per; i < min(nb, (tid + 1) * per); ++i) {
    const int d = dmax - i, c = s_hist[d];
    if (run < m && run + c >= m) {  // exactly one (thread, bin) satisfies this
      s_misc[1] = d;                // d*
      s_misc[3] = m - run;          // how many entries at d* go
    }
    run += c;
  }
  __syncthreads();
  const int dstar = s_misc[1], take_eq = s_misc[3];
  // ---- ordered removal: every warp owns a contiguous piece of the line; entries at d* are ranked by position
  const int piece = (len + NW - 1) / NW;
  const int p0 = warp * piece, p1 = min(len, p0 + piece);
  int eq = 0;
  for (int t = p0 + lane; t < p1; t += 32) eq += s_deg[t] == dstar ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) eq += __shfl_xor_sync(0xffffffffu, eq, o);
  if (lane == 0) s_wtot[warp] = eq;
  __syncthreads();
  int before = 0;
  for (int w2 = 0; w2 < warp; ++w2) before += s_wtot[w2];
  for (int t0 = p0; t0 < p1; t0 += 32) {
    const int t = t0 + lane;
    const int d = t < p1 ? s_deg[t] : -1;
    const unsigned eqm = __ballot_sync(0xffffffffu, d == dstar);
    const int rank = before + __popc(eqm & ((1u << lane) - 1u));
    if (d > dstar || (d == dstar && rank < take_eq)) {
      const int ent = side == 0 ? (int)(b + t) : pb.row_perm[b + t];
      pb.alive[ent] = 0;
      N[side == 0 ? pb.row[ent] : pb.col[ent]] -= 1;  // opposite vertices are distinct inside a line
    }
    before += __popc(eqm);
  }
  __syncthreads();
  return true;
}

// side 0: lines = columns, opposite = rows (N = rowN); side 1: lines = rows, opposite = columns (N = colN)
__global__ void __launch_bounds__(PARE_THREADS) pare_kernel(const PareProblem* __restrict__ probs, int side) {
  // sort keys; the counting select uses the same 32 KB as histogram (first PARE_BINS ints) + degree cache (the rest)
  __shared__ u64 skeys[PARE_SMEM_KEYS];
  static_assert(PARE_SMEM_KEYS * 2 >= PARE_BINS + PARE_LINE_MAX, "shared memory of the counting select");
  int* s_deg = reinterpret_cast<int*>(skeys) + PARE_BINS;
  __shared__ int s_misc[4];
  __shared__ int s_cnt, s_above;
  const PareProblem pb = probs[blockIdx.x];
  if (pb.k < 0) return;
  const int tid = threadIdx.x;
  const int n_lines = side == 0 ? pb.nB : pb.nA;
  const long long* lptr = side == 0 ? pb.col_ptr : pb.row_ptr;
  int* N = side == 0 ? pb.rowN : pb.colN;
  for (int li = 0; li < n_lines; ++li) {
    const int line = (int)(pb.order[li] & 0xffffffffull);
    const long long b = lptr[line], e = lptr[line + 1];
    const int len = (int)(e - b);
    if (len <= pb.k) continue;  // cannot exceed k live entries
    // the live degree of the line itself is current (live_counts ran before this side; the entries of a line die
    // only while that line is visited): nothing to pare
    const int* own = side == 0 ? pb.colN : pb.rowN;  // absent in the standalone pareDownHubEdges entry point
    if (own && own[line] <= pb.k) continue;
    // selection by counting (the common case); the sort below serves very long lines and degrees >= PARE_BINS
    if (pare_line_by_counting(pb, side, b, len, N, s_deg, reinterpret_cast<int*>(skeys), s_misc)) continue;
    // live entries of the line, compacted to the front (the sort below restores the order: the key carries the
    // position among the stored entries); only the live ones are sorted -- after the first pass of the ladder most
    // of a hub line is dead
    if (tid == 0) { s_cnt = 0; s_above = 0; }
    __syncthreads();
    u64* keys = len <= PARE_SMEM_KEYS ? skeys : pb.scratch;
    for (int t = tid; t < len; t += PARE_THREADS) {
      const int ent = side == 0 ? (int)(b + t) : pb.row_perm[b + t];
      if (pb.alive[ent]) {
        const int opp = side == 0 ? pb.row[ent] : pb.col[ent];
        const unsigned deg = (unsigned)N[opp];
        // degree descending, position ascending
        keys[atomicAdd(&s_cnt, 1)] = ((u64)(0x7fffffffu - deg) << 32) | (u64)(unsigned)t;
      }
    }
    __syncthreads();
    const int nedges = s_cnt;
    int np = 2;
    while (np < nedges) np <<= 1;
    for (int t = nedges + tid; t < np; t += PARE_THREADS) keys[t] = ~0ull;
    __syncthreads();
    if (nedges > pb.k) {
      block_bitonic(keys, np, tid, PARE_THREADS);
      // removable prefix: entries whose opposite degree exceeds klow (sorted descending => a prefix)
      int above = 0;
      for (int t = tid; t < nedges; t += PARE_THREADS) {
        const unsigned deg = 0x7fffffffu - (unsigned)(keys[t] >> 32);
        above += (int)deg > pb.klow ? 1 : 0;
      }
      atomicAdd(&s_above, above);
      __syncthreads();
      int m = nedges - pb.k;
      if (s_above < m) m = s_above;
      for (int t = tid; t < m; t += PARE_THREADS) {
        const int posn = (int)(keys[t] & 0xffffffffull);
        const int ent = side == 0 ? (int)(b + posn) : pb.row_perm[b + posn];
        pb.alive[ent] = 0;
        const int opp = side == 0 ? pb.row[ent] : pb.col[ent];
        N[opp] -= 1;  // opposite vertices are distinct inside a line
      }
    }
    __syncthreads();
  }
}

__global__ void live_counts_kernel(const int* __restrict__ row, const int* __restrict__ col,
                                   const int* __restrict__ alive, long long nnz, int* __restrict__ rowN,
                                   int* __restrict__ colN) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz || !alive[t]) return;
  atomicAdd(&rowN[row[t]], 1);
  atomicAdd(&colN[col[t]], 1);
}

__global__ void order_keys_kernel(const int* __restrict__ cnt, int n, u64* __restrict__ keys) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) keys[t] = ((u64)(0x7fffffffu - (unsigned)cnt[t]) << 32) | (u64)(unsigned)t;  // order(-cnt), stable
}

__global__ void max_int_kernel(const int* __restrict__ a, int n, int* __restrict__ out) {
  int m = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) m = max(m, a[t]);
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void fill_ones_kernel(int* a, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) a[t] = 1;
}

__global__ void row_scatter_kernel(const int* __restrict__ row, const int* __restrict__ col, long long nnz,
                                   const long long* __restrict__ row_ptr, int* __restrict__ cursor,
                                   u64* __restrict__ keys) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  const int r = row[t];
  keys[row_ptr[r] + atomicAdd(&cursor[r], 1)] = ((u64)(unsigned)col[t] << 32) | (u64)(unsigned)t;
}
__global__ void low32_kernel(const u64* __restrict__ keys, long long n, int* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (int)(keys[t] & 0xffffffffull);
}

__global__ void emit_alive_kernel(const int* __restrict__ row, const int* __restrict__ col, const double* __restrict__ w,
                                  const int* __restrict__ alive, const long long* __restrict__ pos, long long nnz,
                                  const int* __restrict__ rows_a, const int* __restrict__ rows_b, int offA, int offB,
                                  int* __restrict__ o_from, int* __restrict__ o_to, double* __restrict__ o_w,
                                  int* __restrict__ o_type) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz || !alive[t]) return;
  const long long o = pos[t];
  o_from[o] = offA + (rows_a ? rows_a[row[t]] : row[t]);
  o_to[o] = offB + (rows_b ? rows_b[col[t]] : col[t]);
  o_w[o] = w[t];
  o_type[o] = 1;
}

int r_round(double x) { return (int)std::nearbyint(x); }

// One pair: matched COO in CSC order (device)
struct Matched {
  DevBuf<int> row, col, alive, col_cnt;
  DevBuf<double> w;
  long long nnz = 0;
};

void match_one(Context& ctx, const MnnProblem& pb, int k1, const SimParams& sp, Matched& out) {
  cudaStream_t s = ctx.stream;
  const long long ntrip = (long long)(pb.nA + pb.nB) * k1;
  CG_CHECK(ntrip < (1ll << 32), "too many neighbour triples for one pair");
  DevBuf<int> cnt((size_t)pb.nB), cursor((size_t)pb.nB);
  cnt.zero(s);
  cursor.zero(s);
  triple_count_kernel<<<nblk((size_t)ntrip), 256, 0, s>>>(pb.idx_ab, pb.nA, pb.idx_ba, pb.nB, k1, cnt.p);
  DevBuf<long long> col_off((size_t)pb.nB + 1);
  exclusive_scan_i32_to_i64(ctx, cnt.p, col_off.p, (size_t)pb.nB, col_off.p + pb.nB);
  long long nt = 0;
  CG_CUDA(cudaMemcpyAsync(&nt, col_off.p + pb.nB, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  out.nnz = 0;
  out.col_cnt.alloc((size_t)pb.nB);
  out.col_cnt.zero(s);
  if (nt == 0) return;
  DevBuf<u64> keys((size_t)nt);
  triple_scatter_kernel<<<nblk((size_t)ntrip), 256, 0, s>>>(pb.idx_ab, pb.nA, pb.idx_ba, pb.nB, k1, col_off.p, cursor.p,
                                                            keys.p);
  segmented_sort_u64(ctx, keys.p, col_off.p, pb.nB);
  DevBuf<int> keep((size_t)nt);
  DevBuf<double> wgt((size_t)nt);
  DevBuf<long long> pos((size_t)nt + 1);
  triple_merge_kernel<<<nblk((size_t)pb.nB), 256, 0, s>>>(keys.p, col_off.p, pb.nB, pb.nA, k1, pb.dist_ab, pb.dist_ba,
                                                          sp, keep.p, wgt.p);
  exclusive_scan_i32_to_i64(ctx, keep.p, pos.p, (size_t)nt, pos.p + nt);
  long long nnz = 0;
  CG_CUDA(cudaMemcpyAsync(&nnz, pos.p + nt, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  out.nnz = nnz;
  out.row.alloc((size_t)nnz);
  out.col.alloc((size_t)nnz);
  out.w.alloc((size_t)nnz);
  out.alive.alloc((size_t)nnz);
  if (nnz) {
    triple_emit_kernel<<<nblk((size_t)pb.nB), 256, 0, s>>>(keys.p, col_off.p, pb.nB, keep.p, pos.p, wgt.p, out.row.p,
                                                           out.col.p, out.w.p, out.col_cnt.p);
    fill_ones_kernel<<<nblk((size_t)nnz), 256, 0, s>>>(out.alive.p, nnz);
  }
  ctx.stats.kernels += 5;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
}

// reduceEdgesInGraphIteratively on one matched pair (alive flags updated in place)
// reduceEdgesInGraphIteratively (R/conos.R:906-933) for a batch of pairs.  The greedy paring of one pair is sequential
// by definition (every line sees the degrees the previous lines left behind), so the parallelism is across pairs: every
// pass is ONE pare_kernel launch with a block per pair, the host decisions of the ladder (max degree, number of steps,
// cleanup pass) are taken for all pairs from one download.
struct PareState {
  Matched* m = nullptr;
  int nA = 0, nB = 0;
  DevBuf<int> rowN, colN, row_perm;
  DevBuf<long long> col_ptr, row_ptr, one_seg;
  DevBuf<u64> scratch, order;
  std::vector<int> ks;
  int klow_final = 0;
  bool active = false;
};

void reduce_edges_batch(Context& ctx, const MnnProblem* probs, Matched* ms, int n, int k) {
  cudaStream_t s = ctx.stream;
  if (n == 0) return;
  std::vector<PareState> st((size_t)n);
  DevBuf<int> dmax((size_t)n);
  std::vector<int> hmax((size_t)n, 0);
  auto live_counts = [&](PareState& q) {
    q.rowN.zero(s);
    q.colN.zero(s);
    live_counts_kernel<<<nblk((size_t)q.m->nnz), 256, 0, s>>>(q.m->row.p, q.m->col.p, q.m->alive.p, q.m->nnz, q.rowN.p,
                                                             q.colN.p);
    ctx.stats.kernels++;
  };
  auto max_degrees = [&]() {  // live degrees must be current; one download for the whole batch
    dmax.zero(s);
    for (int p = 0; p < n; ++p) {
      if (!st[p].active) continue;
      max_int_kernel<<<64, 256, 0, s>>>(st[p].rowN.p, st[p].nA, dmax.p + p);
      max_int_kernel<<<64, 256, 0, s>>>(st[p].colN.p, st[p].nB, dmax.p + p);
      ctx.stats.kernels += 2;
    }
    dmax.download(hmax.data(), (size_t)n, s);
    CG_CUDA(cudaStreamSynchronize(s));
  };
  for (int p = 0; p < n; ++p) {
    PareState& q = st[p];
    q.m = &ms[p];
    q.nA = probs[p].nA;
    q.nB = probs[p].nB;
    q.active = ms[p].nnz > 0;
    if (!q.active) continue;
    q.rowN.alloc((size_t)q.nA);
    q.colN.alloc((size_t)q.nB);
    q.col_ptr.alloc((size_t)q.nB + 1);
    exclusive_scan_i32_to_i64(ctx, q.m->col_cnt.p, q.col_ptr.p, (size_t)q.nB, q.col_ptr.p + q.nB);
    live_counts(q);
  }
  max_degrees();
  const int max_kdiff = 5;
  size_t max_steps = 0;
  for (int p = 0; p < n; ++p) {
    PareState& q = st[p];
    if (!q.active) continue;
    const int maxd = hmax[p];
    if (maxd <= k) {  // nothing to be done - already below k
      q.active = false;
      continue;
    }
    // CSR permutation (stored entries by row, columns ascending)
    q.row_ptr.alloc((size_t)q.nA + 1);
    q.row_perm.alloc((size_t)q.m->nnz);
    exclusive_scan_i32_to_i64(ctx, q.rowN.p, q.row_ptr.p, (size_t)q.nA, q.row_ptr.p + q.nA);
    {
      DevBuf<u64> rk((size_t)q.m->nnz);
      DevBuf<int> cursor((size_t)q.nA);
      cursor.zero(s);
      row_scatter_kernel<<<nblk((size_t)q.m->nnz), 256, 0, s>>>(q.m->row.p, q.m->col.p, q.m->nnz, q.row_ptr.p, cursor.p,
                                                               rk.p);
      segmented_sort_u64(ctx, rk.p, q.row_ptr.p, q.nA);
      low32_kernel<<<nblk((size_t)q.m->nnz), 256, 0, s>>>(rk.p, q.m->nnz, q.row_perm.p);
      ctx.stats.kernels += 2;
    }
    q.klow_final = std::max(std::min(k, 3), k - max_kdiff);
    const int n_steps = std::min(3, r_round((double)maxd / k));
    if (n_steps > 1) {
      // round(exp(seq(log(maxd), log(k), length.out = n.steps + 1))[-1])
      for (int t = 1; t <= n_steps; ++t) {
        const double v = std::exp(std::log((double)maxd) + (std::log((double)k) - std::log((double)maxd)) * t / n_steps);
        q.ks.push_back(r_round(v));
      }
    } else {
      q.ks.push_back(k);
    }
    max_steps = std::max(max_steps, q.ks.size());
    int max_line = 2;  // scratch sized for the longest stored line
    while (max_line < maxd) max_line <<= 1;
    q.scratch.alloc((size_t)max_line);
    q.order.alloc((size_t)std::max(q.nA, q.nB));
    q.one_seg.alloc(2);
  }
  DevBuf<PareProblem> dprob((size_t)n);
  std::vector<PareProblem> hprob((size_t)n);
  auto one_pass = [&](const std::vector<int>& kk, const std::vector<int>& klow) {
    // reduceEdgesInGraph: column side, then row side (R/conos.R:890-901)
    bool any = false;
    for (int p = 0; p < n; ++p) any = any || (st[p].active && kk[p] >= 0);
    if (!any) return;
    for (int side = 0; side < 2; ++side) {
      for (int p = 0; p < n; ++p) {
        PareState& q = st[p];
        hprob[p] = PareProblem{};
        hprob[p].k = -1;
        if (!q.active || kk[p] < 0) continue;
        live_counts(q);
        const int n_lines = side == 0 ? q.nB : q.nA;
        order_keys_kernel<<<nblk((size_t)n_lines), 256, 0, s>>>(side == 0 ? q.colN.p : q.rowN.p, n_lines, q.order.p);
        const long long seg[2] = {0, n_lines};
        q.one_seg.upload(seg, 2, s);
        segmented_sort_u64(ctx, q.order.p, q.one_seg.p, 1);
        hprob[p] = PareProblem{q.m->row.p, q.m->col.p, q.m->alive.p, q.m->nnz, q.nA, q.nB, q.col_ptr.p, q.row_ptr.p,
                               q.row_perm.p, q.rowN.p, q.colN.p, q.order.p, q.scratch.p, kk[p], klow[p]};
        ctx.stats.kernels++;
      }
      dprob.upload(hprob.data(), (size_t)n, s);
      pare_kernel<<<n, PARE_THREADS, 0, s>>>(dprob.p, side);
      ctx.stats.kernels++;
      CG_CUDA(cudaGetLastError());
      CG_CUDA(cudaStreamSynchronize(s));
    }
  };
  std::vector<int> kk((size_t)n), kl((size_t)n);
  for (size_t t = 0; t < max_steps; ++t) {
    for (int p = 0; p < n; ++p) {
      kk[p] = (st[p].active && t < st[p].ks.size()) ? st[p].ks[t] : -1;
      kl[p] = kk[p];
    }
    one_pass(kk, kl);
  }
  bool any_active = false;
  for (int p = 0; p < n; ++p)
    if (st[p].active) {
      live_counts(st[p]);
      any_active = true;
    }
  if (!any_active) return;
  max_degrees();
  for (int p = 0; p < n; ++p) {  // cleanup step
    const bool need = st[p].active && hmax[p] - k > max_kdiff;
    kk[p] = need ? k : -1;
    kl[p] = st[p].klow_final;
  }
  one_pass(kk, kl);
}

}  // namespace

void match_append_edges(Context& ctx, const MnnProblem* h_probs, int n_probs, int k, int k1, const SimParams& sp,
                        EdgeTable& out, long long* pair_counts) {
  if (sp.matching == 0 && k1 == k && k1 <= 32) {
    mnn_append_edges(ctx, h_probs, n_probs, k1, sp, out, pair_counts);
    return;
  }
  cudaStream_t s = ctx.stream;
  // sub-batches bounded by the matched structures that are alive at the same time (~24 B per neighbour triple)
  // what stays alive per pair is the matched matrix (row, col, w, alive + the row permutation: 24 B per entry, at most
  // (nA + nB) k1 / 2 entries); the triples of match_one are temporaries of one pair.  The greedy paring of a pair is
  // sequential, so the throughput of the stage is the number of pairs a launch works on side by side.
  const size_t budget = (size_t)40 << 30;
  int p0 = 0;
  while (p0 < n_probs) {
    int p1 = p0;
    size_t used = 0;
    while (p1 < n_probs && p1 - p0 < 1024) {
      const size_t need = (size_t)(h_probs[p1].nA + h_probs[p1].nB) * (size_t)k1 * 12;
      if (p1 > p0 && used + need > budget) break;
      used += need;
      ++p1;
    }
    const int nb = p1 - p0;
    std::vector<Matched> ms((size_t)nb);
    for (int q = 0; q < nb; ++q) match_one(ctx, h_probs[p0 + q], k1, sp, ms[q]);
    if (k1 > k) reduce_edges_batch(ctx, h_probs + p0, ms.data(), nb, k);
    for (int q = 0; q < nb; ++q) {
      const MnnProblem& pb = h_probs[p0 + q];
      Matched& m = ms[q];
      long long n_alive = 0;
      if (m.nnz) {
        DevBuf<long long> pos((size_t)m.nnz + 1);
        exclusive_scan_i32_to_i64(ctx, m.alive.p, pos.p, (size_t)m.nnz, pos.p + m.nnz);
        CG_CUDA(cudaMemcpyAsync(&n_alive, pos.p + m.nnz, sizeof(long long), cudaMemcpyDeviceToHost, s));
        CG_CUDA(cudaStreamSynchronize(s));
        out.reserve(out.n + (size_t)n_alive, s);
        if (n_alive) {
          emit_alive_kernel<<<nblk((size_t)m.nnz), 256, 0, s>>>(m.row.p, m.col.p, m.w.p, m.alive.p, pos.p, m.nnz,
                                                                pb.rows_a, pb.rows_b, pb.offA, pb.offB,
                                                                out.from.p + out.n, out.to.p + out.n, out.w.p + out.n,
                                                                out.type.p + out.n);
          ctx.stats.kernels++;
          CG_CUDA(cudaGetLastError());
          CG_CUDA(cudaStreamSynchronize(s));
        }
        out.n += (size_t)n_alive;
      }
      if (pair_counts) pair_counts[p0 + q] = n_alive;
    }
    p0 = p1;
  }
}

// pareDownHubEdges on a CSC matrix (the reference's own entry point): columns are visited in storage order,
// entries with x != 0 are live.
namespace {
__global__ void alive_from_x_kernel(const double* __restrict__ x, long long n, int* __restrict__ alive) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) alive[t] = 1;  // the reference counts every stored entry of the column (p1 - p0), zero or not
}
__global__ void col_of_entry_kernel(const int* __restrict__ p, int ncols, int* __restrict__ col) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ncols) return;
  for (int t = p[g]; t < p[g + 1]; ++t) col[t] = g;
}
__global__ void identity_order_kernel(int n, u64* __restrict__ o) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) o[t] = (u64)(unsigned)t;
}
__global__ void widen_kernel(const int* __restrict__ p, int n, long long* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = p[t];
}
__global__ void zero_dead_kernel(double* __restrict__ x, const int* __restrict__ alive, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && !alive[t]) x[t] = 0.0;
}
}  // namespace

void pare_down_hub_edges_device(Context& ctx, int ncols, const int* d_p, const int* d_i, double* d_x, int* d_rowN,
                                int k, int klow) {
  cudaStream_t s = ctx.stream;
  int h_last = 0;
  CG_CUDA(cudaMemcpyAsync(&h_last, d_p + ncols, sizeof(int), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  const long long nnz = h_last;
  if (nnz == 0 || ncols == 0) return;
  DevBuf<int> alive((size_t)nnz), col((size_t)nnz);
  DevBuf<long long> col_ptr((size_t)ncols + 1);
  DevBuf<u64> order((size_t)ncols);
  alive_from_x_kernel<<<nblk((size_t)nnz), 256, 0, s>>>(d_x, nnz, alive.p);
  col_of_entry_kernel<<<nblk((size_t)ncols), 256, 0, s>>>(d_p, ncols, col.p);
  widen_kernel<<<nblk((size_t)ncols + 1), 256, 0, s>>>(d_p, ncols + 1, col_ptr.p);
  identity_order_kernel<<<nblk((size_t)ncols), 256, 0, s>>>(ncols, order.p);
  // longest column
  std::vector<int> hp((size_t)ncols + 1);
  CG_CUDA(cudaMemcpyAsync(hp.data(), d_p, ((size_t)ncols + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  int max_line = 2, longest = 0;
  for (int g = 0; g < ncols; ++g) longest = std::max(longest, hp[g + 1] - hp[g]);
  while (max_line < longest) max_line <<= 1;
  DevBuf<u64> scratch((size_t)max_line);
  PareProblem pp{d_i, col.p, alive.p, nnz, 0, ncols, col_ptr.p, nullptr, nullptr, d_rowN, nullptr, order.p, scratch.p,
                 k, klow};
  DevBuf<PareProblem> dprob(1);
  dprob.upload(&pp, 1, s);
  pare_kernel<<<1, PARE_THREADS, 0, s>>>(dprob.p, 0);
  zero_dead_kernel<<<nblk((size_t)nnz), 256, 0, s>>>(d_x, alive.p, nnz);
  ctx.stats.kernels += 6;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
}

}  // namespace cg
