#include <climits>

#include "edges.cuh"
#include "prims.cuh"

namespace cg {

void EdgeTable::reserve(size_t cap, cudaStream_t s) {
  if (cap <= from.n) return;
  size_t newcap = cap + cap / 4 + 1024;
  DevBuf<int> f(newcap), t(newcap), ty(newcap);
  DevBuf<double> ww(newcap);
  if (n) {
    CG_CUDA(cudaMemcpyAsync(f.p, from.p, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.p, to.p, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaMemcpyAsync(ty.p, type.p, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaMemcpyAsync(ww.p, w.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    // the old blocks are released in stream order, after these copies
  }
  from = std::move(f);
  to = std::move(t);
  type = std::move(ty);
  w = std::move(ww);
}

namespace {

__device__ __forceinline__ double similarity(float d, int metric, double cor_base, double l2_sigma) {
  // R/conos.R:767-773, double arithmetic on the float32-valued distance
  if (metric == METRIC_ANGULAR) return fmax(0.0, cor_base - (double)d);
  return exp(-(double)d / l2_sigma);
}

// ------------------------------------------------------------------ mNN (k1 <= 32), one warp per column j of B
__global__ void __launch_bounds__(256) mnn_columns_kernel(const MnnProblem* __restrict__ probs,
                                                          const long long* __restrict__ col_base, int n_probs,
                                                          long long n_cols, int k1, SimParams sp,
                                                          int* __restrict__ tmp_i, double* __restrict__ tmp_w,
                                                          int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const long long col = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (col >= n_cols) return;
  int lo = 0, hi = n_probs - 1;  // last problem with col_base <= col
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (col_base[mid] <= col) lo = mid; else hi = mid - 1;
  }
  const MnnProblem pb = probs[lo];
  const int j = (int)(col - col_base[lo]);
  int i = -1;
  double w = 0.0;
  if (lane < k1) {
    i = pb.idx_ba[(size_t)j * k1 + lane];
    if (i >= 0) {
      const float dji = pb.dist_ba[(size_t)j * k1 + lane];
      const int* row = pb.idx_ab + (size_t)i * k1;
      int s = -1;
      for (int r = 0; r < k1; ++r)
        if (row[r] == j) { s = r; break; }
      if (s >= 0) {
        const float dij = pb.dist_ab[(size_t)i * k1 + s];
        const double sji = similarity(dji, sp.metric, sp.cor_base, sp.l2_sigma);
        const double sij = similarity(dij, sp.metric, sp.cor_base, sp.l2_sigma);
        w = sqrt(sji * sij);  // adj = n21 * t(n12); sqrt(adj@x)
      }
    }
  }
  const bool keep = (w > 0.0) && !(w < sp.min_similarity);
  // rank among kept entries by ascending row index i (CSC order inside the column)
  int rank = 0;
  for (int t = 0; t < k1; ++t) {
    const int it = __shfl_sync(0xffffffffu, i, t);
    const bool kt = __shfl_sync(0xffffffffu, (int)keep, t) != 0;
    rank += (kt && it < i) ? 1 : 0;
  }
  const unsigned km = __ballot_sync(0xffffffffu, keep);
  if (keep) {
    tmp_i[col * k1 + rank] = pb.offA + (pb.rows_a ? pb.rows_a[i] : i);
    tmp_w[col * k1 + rank] = w;
  }
  if (lane == 0) cnt[col] = __popc(km);
}

__global__ void mnn_compact_kernel(const MnnProblem* __restrict__ probs, const long long* __restrict__ col_base,
                                   int n_probs, long long n_cols, int k1, const int* __restrict__ tmp_i,
                                   const double* __restrict__ tmp_w, const int* __restrict__ cnt,
                                   const long long* __restrict__ off, int* __restrict__ from, int* __restrict__ to,
                                   double* __restrict__ w, int* __restrict__ type) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long col = t / k1;
  const int r = (int)(t - col * k1);
  if (col >= n_cols || r >= cnt[col]) return;
  int lo = 0, hi = n_probs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (col_base[mid] <= col) lo = mid; else hi = mid - 1;
  }
  const MnnProblem pb = probs[lo];
  const int j = (int)(col - col_base[lo]);
  const long long o = off[col] + r;
  from[o] = tmp_i[col * k1 + r];
  to[o] = pb.offB + (pb.rows_b ? pb.rows_b[j] : j);
  w[o] = tmp_w[col * k1 + r];
  type[o] = 1;
}

// ------------------------------------------------------------------ local edges, one warp per query column
__global__ void __launch_bounds__(256) local_columns_kernel(const int* __restrict__ idx, const float* __restrict__ dist,
                                                            int n, int kp1, int off, int metric, double l2_sigma,
                                                            double ksw, int* __restrict__ tmp_i,
                                                            double* __restrict__ tmp_w, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= n) return;
  // kp1 may exceed 32: every lane owns entries lane, lane+32, ...
  int total = 0;
  for (int r0 = 0; r0 < kp1; r0 += 32) {
    const int r = r0 + lane;
    int nb = -1;
    float d = 0.f;
    if (r < kp1) {
      nb = idx[(size_t)q * kp1 + r];
      d = dist[(size_t)q * kp1 + r];
    }
    // diag(xk) <- 0 ; drop0: the self entry and exact-zero distances disappear (R/conos.R:596-601)
    const bool keep = nb >= 0 && nb != q && d != 0.0f;
    if (keep) {
      int rank = 0;
      for (int u = 0; u < kp1; ++u) {
        const int nu = idx[(size_t)q * kp1 + u];
        const float du = dist[(size_t)q * kp1 + u];
        rank += (nu >= 0 && nu != q && du != 0.0f && nu < nb) ? 1 : 0;
      }
      tmp_i[(size_t)q * kp1 + rank] = off + nb;
      tmp_w[(size_t)q * kp1 + rank] = similarity(d, metric, 1.0, l2_sigma) * ksw;  // cor.base = 1 for local edges
    }
    total += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if (lane == 0) cnt[q] = total;
}

__global__ void local_compact_kernel(int n, int kp1, int off, const int* __restrict__ tmp_i,
                                     const double* __restrict__ tmp_w, const int* __restrict__ cnt,
                                     const long long* __restrict__ offs, int* __restrict__ from, int* __restrict__ to,
                                     double* __restrict__ w, int* __restrict__ type) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long q = t / kp1;
  const int r = (int)(t - q * kp1);
  if (q >= n || r >= cnt[q]) return;
  const long long o = offs[q] + r;
  from[o] = tmp_i[q * kp1 + r];
  to[o] = off + (int)q;
  w[o] = tmp_w[q * kp1 + r];
  type[o] = 0;
}

}  // namespace

void mnn_append_edges(Context& ctx, const MnnProblem* h_probs, int n_probs, int k1, const SimParams& sp,
                      EdgeTable& out, long long* pair_counts) {
  if (n_probs == 0) return;
  CG_CHECK(k1 >= 1 && k1 <= 32, "mutual-neighbour fast path serves k1 <= 32");
  CG_CHECK(sp.matching == 0, "fast path is mNN only");
  std::vector<long long> base(n_probs + 1, 0);
  for (int p = 0; p < n_probs; ++p) base[p + 1] = base[p] + h_probs[p].nB;
  const long long n_cols = base[n_probs];
  if (n_cols == 0) {
    if (pair_counts) for (int p = 0; p < n_probs; ++p) pair_counts[p] = 0;
    return;
  }
  DevBuf<MnnProblem> d_probs;
  d_probs.upload(h_probs, n_probs, ctx.stream);
  DevBuf<long long> d_base;
  d_base.upload(base.data(), base.size(), ctx.stream);
  DevBuf<int> tmp_i((size_t)n_cols * k1), cnt((size_t)n_cols);
  DevBuf<double> tmp_w((size_t)n_cols * k1);
  DevBuf<long long> off((size_t)n_cols + 1);
  mnn_columns_kernel<<<(unsigned)((n_cols + 7) / 8), 256, 0, ctx.stream>>>(d_probs.p, d_base.p, n_probs, n_cols, k1, sp,
                                                                            tmp_i.p, tmp_w.p, cnt.p);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
  exclusive_scan_i32_to_i64(ctx, cnt.p, off.p, (size_t)n_cols, off.p + n_cols);
  std::vector<long long> h_off;
  long long total = 0;
  CG_CUDA(cudaMemcpyAsync(&total, off.p + n_cols, sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
  if (pair_counts) {
    // offsets at the problem boundaries
    h_off.resize(n_probs + 1);
    for (int p = 0; p <= n_probs; ++p)
      CG_CUDA(cudaMemcpyAsync(&h_off[p], off.p + base[p], sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
  }
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
  if (pair_counts) for (int p = 0; p < n_probs; ++p) pair_counts[p] = h_off[p + 1] - h_off[p];
  out.reserve(out.n + (size_t)total, ctx.stream);
  if (total > 0) {
    const long long nt = n_cols * k1;
    mnn_compact_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx.stream>>>(
        d_probs.p, d_base.p, n_probs, n_cols, k1, tmp_i.p, tmp_w.p, cnt.p, off.p, out.from.p + out.n, out.to.p + out.n,
        out.w.p + out.n, out.type.p + out.n);
    ctx.stats.kernels++;
    CG_CUDA(cudaGetLastError());
  }
  out.n += (size_t)total;   // temporaries are released in stream order: no host wait
}

void local_append_edges(Context& ctx, const int* d_idx, const float* d_dist, int n, int kp1, int off, int metric,
                        double l2_sigma, double k_self_weight, EdgeTable& out) {
  if (n == 0 || kp1 <= 0) return;
  DevBuf<int> tmp_i((size_t)n * kp1), cnt((size_t)n);
  DevBuf<double> tmp_w((size_t)n * kp1);
  DevBuf<long long> offs((size_t)n + 1);
  local_columns_kernel<<<(n + 7) / 8, 256, 0, ctx.stream>>>(d_idx, d_dist, n, kp1, off, metric, l2_sigma, k_self_weight,
                                                            tmp_i.p, tmp_w.p, cnt.p);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
  exclusive_scan_i32_to_i64(ctx, cnt.p, offs.p, (size_t)n, offs.p + n);
  long long total = 0;
  CG_CUDA(cudaMemcpyAsync(&total, offs.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
  out.reserve(out.n + (size_t)total, ctx.stream);
  if (total > 0) {
    const long long nt = (long long)n * kp1;
    local_compact_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx.stream>>>(
        n, kp1, off, tmp_i.p, tmp_w.p, cnt.p, offs.p, out.from.p + out.n, out.to.p + out.n, out.w.p + out.n,
        out.type.p + out.n);
    ctx.stats.kernels++;
    CG_CUDA(cudaGetLastError());
  }
  out.n += (size_t)total;
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
}

// The same for several samples with ONE host round trip: the columns of all samples are counted into one array, one
// scan gives every row its place in the table, the total and the per-sample counts come back together.  Rows are
// appended sample by sample, row by row -- the order of consecutive local_append_edges calls.
void local_append_edges_batch(Context& ctx, const LocalProblem* probs, int n_probs, int kp1, int metric,
                              double l2_sigma, double k_self_weight, EdgeTable& out, long long* counts) {
  std::vector<long long> base(n_probs + 1, 0);
  for (int p = 0; p < n_probs; ++p) base[p + 1] = base[p] + probs[p].n;
  const long long N = base[n_probs];
  if (counts)
    for (int p = 0; p < n_probs; ++p) counts[p] = 0;
  if (N == 0 || kp1 <= 0) return;
  DevBuf<int> tmp_i((size_t)N * kp1), cnt((size_t)N);
  DevBuf<double> tmp_w((size_t)N * kp1);
  DevBuf<long long> offs((size_t)N + 1);
  for (int p = 0; p < n_probs; ++p) {
    if (probs[p].n == 0) continue;
    local_columns_kernel<<<(probs[p].n + 7) / 8, 256, 0, ctx.stream>>>(
        probs[p].idx, probs[p].dist, probs[p].n, kp1, probs[p].off, metric, l2_sigma, k_self_weight,
        tmp_i.p + base[p] * kp1, tmp_w.p + base[p] * kp1, cnt.p + base[p]);
    ctx.stats.kernels++;
  }
  CG_CUDA(cudaGetLastError());
  exclusive_scan_i32_to_i64(ctx, cnt.p, offs.p, (size_t)N, offs.p + N);
  std::vector<long long> h_off(n_probs + 1, 0);
  for (int p = 0; p <= n_probs; ++p)
    CG_CUDA(cudaMemcpyAsync(&h_off[p], offs.p + base[p], sizeof(long long), cudaMemcpyDeviceToHost, ctx.stream));
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
  const long long total = h_off[n_probs];
  if (counts)
    for (int p = 0; p < n_probs; ++p) counts[p] = h_off[p + 1] - h_off[p];
  out.reserve(out.n + (size_t)total, ctx.stream);
  if (total > 0)
    for (int p = 0; p < n_probs; ++p) {
      if (probs[p].n == 0) continue;
      const long long nt = (long long)probs[p].n * kp1;
      local_compact_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx.stream>>>(
          probs[p].n, kp1, probs[p].off, tmp_i.p + base[p] * kp1, tmp_w.p + base[p] * kp1, cnt.p + base[p],
          offs.p + base[p], out.from.p + out.n, out.to.p + out.n, out.w.p + out.n, out.type.p + out.n);
      ctx.stats.kernels++;
    }
  CG_CUDA(cudaGetLastError());
  out.n += (size_t)total;
}

// ==================================================================== graph assembly
namespace {

__global__ void flag_positive_kernel(const double* __restrict__ w, size_t n, int* __restrict__ flag) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) flag[e] = w[e] > 0.0 ? 1 : 0;
}

__global__ void first_pos_kernel(const int* __restrict__ from, const int* __restrict__ to, const int* __restrict__ flag,
                                 const long long* __restrict__ kept_off, size_t n, int* __restrict__ firstpos) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !flag[e]) return;
  const long long ke = kept_off[e];  // row of the filtered table
  atomicMin(&firstpos[from[e]], (int)(2 * ke));
  atomicMin(&firstpos[to[e]], (int)(2 * ke + 1));
}

__global__ void mark_first_kernel(const int* __restrict__ firstpos, int n_cells, int* __restrict__ posflag) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n_cells && firstpos[v] != INT_MAX) posflag[firstpos[v]] = 1;
}

__global__ void assign_vid_kernel(const int* __restrict__ firstpos, const int* __restrict__ posrank, int n_cells,
                                  int* __restrict__ vid, int* __restrict__ vertices) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_cells) return;
  if (firstpos[v] == INT_MAX) { vid[v] = -1; return; }
  int id = posrank[firstpos[v]];
  vid[v] = id;
  vertices[id] = v;
}

// degree of the low endpoint (upper-triangular bucket sizes); loops are dropped
__global__ void lo_degree_kernel(const int* __restrict__ from, const int* __restrict__ to, const int* __restrict__ flag,
                                 size_t n, const int* __restrict__ vid, int* __restrict__ deg) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !flag[e]) return;
  int a = vid[from[e]], b = vid[to[e]];
  if (a == b) return;
  atomicAdd(&deg[a < b ? a : b], 1);
}

__global__ void lo_scatter_kernel(const int* __restrict__ from, const int* __restrict__ to, const int* __restrict__ flag,
                                  size_t n, const int* __restrict__ vid, const long long* __restrict__ seg_off,
                                  int* __restrict__ cursor, unsigned long long* __restrict__ keys) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !flag[e]) return;
  int a = vid[from[e]], b = vid[to[e]];
  if (a == b) return;
  int lo = a < b ? a : b, hi = a < b ? b : a;
  int slot = atomicAdd(&cursor[lo], 1);
  keys[seg_off[lo] + slot] = ((unsigned long long)(unsigned)hi << 32) | (unsigned long long)(unsigned)e;
}

// after the segmented sort: head of every run of equal `hi` inside a segment
__global__ void run_heads_kernel(const unsigned long long* __restrict__ keys, const long long* __restrict__ seg_off,
                                 int n_vertices, int* __restrict__ head) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vertices) return;
  long long b = seg_off[v], e = seg_off[v + 1];
  unsigned prev = 0xffffffffu;
  for (long long t = b; t < e; ++t) {
    unsigned hi = (unsigned)(keys[t] >> 32);
    head[t] = (t == b || hi != prev) ? 1 : 0;
    prev = hi;
  }
}

// one thread per run head: weight = sum in table order (keys are sorted by original row), type = first
__global__ void merge_runs_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ head,
                                  const long long* __restrict__ out_pos, long long n_keys,
                                  const long long* __restrict__ seg_off, int n_vertices,
                                  const double* __restrict__ w, const int* __restrict__ type,
                                  int* __restrict__ o_to, double* __restrict__ o_w, int* __restrict__ o_type) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_keys || !head[t]) return;
  const unsigned hi = (unsigned)(keys[t] >> 32);
  double s = 0.0;
  long long u = t;
  do {
    s += w[(unsigned)(keys[u] & 0xffffffffull)];
    ++u;
  } while (u < n_keys && !head[u]);
  const long long o = out_pos[t];
  o_to[o] = (int)hi;
  o_w[o] = s;
  o_type[o] = type[(unsigned)(keys[t] & 0xffffffffull)];
}

// `from` of the merged edges: per vertex, number of run heads in its segment
__global__ void fill_from_kernel(const long long* __restrict__ seg_off, const long long* __restrict__ out_pos,
                                 const int* __restrict__ head, long long n_keys, int n_vertices,
                                 int* __restrict__ o_from) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vertices) return;
  for (long long t = seg_off[v]; t < seg_off[v + 1]; ++t)
    if (head[t]) o_from[out_pos[t]] = v;
}

__global__ void adj_degree_kernel(const int* __restrict__ from, const int* __restrict__ to, long long n,
                                  int* __restrict__ deg) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  atomicAdd(&deg[from[e]], 1);
  atomicAdd(&deg[to[e]], 1);
}

__global__ void adj_scatter_kernel(const int* __restrict__ from, const int* __restrict__ to, long long n,
                                   const long long* __restrict__ rowp, int* __restrict__ cursor,
                                   unsigned long long* __restrict__ keys) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  int a = from[e], b = to[e];
  keys[rowp[a] + atomicAdd(&cursor[a], 1)] = ((unsigned long long)(unsigned)b << 32) | (unsigned long long)e;
  keys[rowp[b] + atomicAdd(&cursor[b], 1)] = ((unsigned long long)(unsigned)a << 32) | (unsigned long long)e;
}

__global__ void adj_fill_kernel(const unsigned long long* __restrict__ keys, long long n, const double* __restrict__ w,
                                int* __restrict__ ci, double* __restrict__ x) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  ci[t] = (int)(keys[t] >> 32);
  x[t] = w[keys[t] & 0xffffffffull];
}

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void assemble_graph(Context& ctx, const EdgeTable& el, int n_cells_total, Graph& g) {
  cudaStream_t s = ctx.stream;
  const size_t n = el.n;
  CG_CHECK(n < (size_t)1 << 30, "edge table too large for 32-bit positions");
  g.n_vertices = 0;
  g.n_edges = 0;
  if (n == 0 || n_cells_total == 0) return;
  // 1. el <- el[w > 0, ]
  DevBuf<int> flag(n);
  DevBuf<long long> kept_off(n + 1);
  flag_positive_kernel<<<nblk(n), 256, 0, s>>>(el.w.p, n, flag.p);
  exclusive_scan_i32_to_i64(ctx, flag.p, kept_off.p, n, kept_off.p + n);
  long long n_kept = 0;
  CG_CUDA(cudaMemcpyAsync(&n_kept, kept_off.p + n, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  ctx.stats.kernels += 1;
  if (n_kept == 0) return;
  // 2. vertex ids by first appearance scanning (from_1, to_1, from_2, to_2, ...)
  DevBuf<int> firstpos(n_cells_total), posflag(2 * (size_t)n_kept), posrank(2 * (size_t)n_kept), vid(n_cells_total);
  CG_CUDA(cudaMemsetAsync(firstpos.p, 0x7f, (size_t)n_cells_total * sizeof(int), s));  // 0x7f7f7f7f, fixed below
  {
    // INT_MAX fill (memset gives 0x7f7f7f7f; use a kernel-free trick: cudaMemset is byte-wise, so fill by copy)
    std::vector<int> h(n_cells_total, INT_MAX);
    CG_CUDA(cudaMemcpyAsync(firstpos.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaStreamSynchronize(s));
  }
  posflag.zero(s);
  first_pos_kernel<<<nblk(n), 256, 0, s>>>(el.from.p, el.to.p, flag.p, kept_off.p, n, firstpos.p);
  mark_first_kernel<<<nblk(n_cells_total), 256, 0, s>>>(firstpos.p, n_cells_total, posflag.p);
  DevBuf<int> nv_dev(1);
  exclusive_scan_i32(ctx, posflag.p, posrank.p, 2 * (size_t)n_kept, nv_dev.p);
  int nv = 0;
  nv_dev.download(&nv, 1, s);
  CG_CUDA(cudaStreamSynchronize(s));
  g.n_vertices = nv;
  g.vertices.alloc(nv);
  assign_vid_kernel<<<nblk(n_cells_total), 256, 0, s>>>(firstpos.p, posrank.p, n_cells_total, vid.p, g.vertices.p);
  ctx.stats.kernels += 3;
  // 3. bucket by low endpoint, sort each bucket by (high endpoint, table row)
  DevBuf<int> deg((size_t)nv), cursor((size_t)nv);
  deg.zero(s);
  cursor.zero(s);
  lo_degree_kernel<<<nblk(n), 256, 0, s>>>(el.from.p, el.to.p, flag.p, n, vid.p, deg.p);
  DevBuf<long long> seg_off((size_t)nv + 1);
  exclusive_scan_i32_to_i64(ctx, deg.p, seg_off.p, (size_t)nv, seg_off.p + nv);
  long long n_keys = 0;
  CG_CUDA(cudaMemcpyAsync(&n_keys, seg_off.p + nv, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  ctx.stats.kernels += 1;
  if (n_keys == 0) return;
  DevBuf<unsigned long long> keys((size_t)n_keys);
  lo_scatter_kernel<<<nblk(n), 256, 0, s>>>(el.from.p, el.to.p, flag.p, n, vid.p, seg_off.p, cursor.p, keys.p);
  ctx.stats.kernels += 1;
  segmented_sort_u64(ctx, keys.p, seg_off.p, nv);
  // 4. simplify: merge runs of equal (lo, hi)
  DevBuf<int> head((size_t)n_keys);
  DevBuf<long long> out_pos((size_t)n_keys + 1);
  run_heads_kernel<<<nblk(nv), 256, 0, s>>>(keys.p, seg_off.p, nv, head.p);
  exclusive_scan_i32_to_i64(ctx, head.p, out_pos.p, (size_t)n_keys, out_pos.p + n_keys);
  long long n_edges = 0;
  CG_CUDA(cudaMemcpyAsync(&n_edges, out_pos.p + n_keys, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  g.n_edges = n_edges;
  g.from.alloc((size_t)n_edges);
  g.to.alloc((size_t)n_edges);
  g.type.alloc((size_t)n_edges);
  g.weight.alloc((size_t)n_edges);
  merge_runs_kernel<<<nblk((size_t)n_keys), 256, 0, s>>>(keys.p, head.p, out_pos.p, n_keys, seg_off.p, nv, el.w.p,
                                                         el.type.p, g.to.p, g.weight.p, g.type.p);
  fill_from_kernel<<<nblk(nv), 256, 0, s>>>(seg_off.p, out_pos.p, head.p, n_keys, nv, g.from.p);
  ctx.stats.kernels += 3;
  // 5. symmetric adjacency in vertex order (as_adj(graph, attr = 'weight'))
  deg.zero(s);
  cursor.zero(s);
  adj_degree_kernel<<<nblk((size_t)n_edges), 256, 0, s>>>(g.from.p, g.to.p, n_edges, deg.p);
  g.adj_p.alloc((size_t)nv + 1);
  exclusive_scan_i32_to_i64(ctx, deg.p, g.adj_p.p, (size_t)nv, g.adj_p.p + nv);
  DevBuf<unsigned long long> akeys((size_t)(2 * n_edges));
  adj_scatter_kernel<<<nblk((size_t)n_edges), 256, 0, s>>>(g.from.p, g.to.p, n_edges, g.adj_p.p, cursor.p, akeys.p);
  segmented_sort_u64(ctx, akeys.p, g.adj_p.p, nv);
  g.adj_i.alloc((size_t)(2 * n_edges));
  g.adj_x.alloc((size_t)(2 * n_edges));
  adj_fill_kernel<<<nblk((size_t)(2 * n_edges)), 256, 0, s>>>(akeys.p, 2 * n_edges, g.weight.p, g.adj_i.p, g.adj_x.p);
  ctx.stats.kernels += 4;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
}

}  // namespace cg
