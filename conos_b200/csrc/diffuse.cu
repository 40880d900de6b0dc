// Diffusion over the joint graph: smooth_count_matrix / propagate_labels,
// src/propagate_labels.cpp:65-222, driven by Conos$propagateLabels(method = 'diffusion') (R/conclass.R:910-925,
// R/conos.R:1071-1083) and Conos$correctGenes (R/conclass.R:862-893).
//
// The reference walks the edge list once per iteration and adds  exp(-a * (1 - w' + b)) * old row of one end  to the
// new row of the other end.  Here every vertex owns its incidence list IN EDGE ORDER (stable: key = 2 * edge + side),
// so one warp sums a row in exactly the reference's order with separate multiply and add (no FMA): the iteration is
// a deterministic, HBM-bound SpMM over a row-major copy of the matrix, and the results differ from the reference's
// only through exp() (1 ulp between libm and the device).
//
// Bytes per iteration (algorithmic): every incidence entry reads one matrix row (8 * n_cols B) plus 12 B of
// (neighbour, weight); every vertex writes one row.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "api_core.cuh"
#include "prims.cuh"

namespace cg {
namespace {

inline unsigned nblocks(size_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

// ---- min / max of the edge weights: per-block partials, then one block (fixed order: deterministic)
__global__ void minmax_partial_kernel(const double* __restrict__ w, long long n, double* __restrict__ part) {
  __shared__ double smn[256], smx[256];
  double mn = INFINITY, mx = -INFINITY;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const double v = w[t];
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  smn[threadIdx.x] = mn;
  smx[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      smn[threadIdx.x] = smn[threadIdx.x + s] < smn[threadIdx.x] ? smn[threadIdx.x + s] : smn[threadIdx.x];
      smx[threadIdx.x] = smx[threadIdx.x + s] > smx[threadIdx.x] ? smx[threadIdx.x + s] : smx[threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = smn[0];
    part[2 * blockIdx.x + 1] = smx[0];
  }
}
__global__ void minmax_final_kernel(const double* __restrict__ part, int n_part, double* __restrict__ out) {
  double mn = INFINITY, mx = -INFINITY;
  for (int t = 0; t < n_part; ++t) {
    mn = part[2 * t] < mn ? part[2 * t] : mn;
    mx = part[2 * t + 1] > mx ? part[2 * t + 1] : mx;
  }
  out[0] = mn;
  out[1] = mx;
}

// parse_edges (:40-62) + the weight of smooth_count_matrix_c (:137); flags[0] |= 1 on a negative length,
// flags[0] |= 2 on a vertex id outside [0, n_vertices)
__global__ void edge_weight_kernel(const int* __restrict__ from, const int* __restrict__ to,
                                   const double* __restrict__ w, long long n, int n_vertices,
                                   const double* __restrict__ mm, double fading, double fading_const,
                                   double* __restrict__ ew, int* __restrict__ deg, int* __restrict__ flags) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double min_w = mm[0] > 0.0 ? mm[0] : 0.0;
  const double max_w = mm[1] - min_w;
  const double len = 1.0 - (w[e] - min_w) / max_w;
  if (len < 0.0) atomicOr(flags, 1);
  ew[e] = exp(-fading * (len + fading_const));
  const int a = from[e], b = to[e];
  if (a < 0 || a >= n_vertices || b < 0 || b >= n_vertices) {
    atomicOr(flags, 2);
    return;
  }
  atomicAdd(deg + a, 1);
  atomicAdd(deg + b, 1);
}

// incidence entries: key = (2 * edge + side) << 31 | other end; the segmented sort restores edge order per vertex
__global__ void incidence_fill_kernel(const int* __restrict__ from, const int* __restrict__ to, long long n,
                                      const long long* __restrict__ row_p, int* __restrict__ cursor,
                                      unsigned long long* __restrict__ keys) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int a = from[e], b = to[e];
  const long long pa = row_p[a] + atomicAdd(cursor + a, 1);
  keys[pa] = ((unsigned long long)(2 * e) << 31) | (unsigned)b;
  const long long pb = row_p[b] + atomicAdd(cursor + b, 1);
  keys[pb] = ((unsigned long long)(2 * e + 1) << 31) | (unsigned)a;
}
__global__ void incidence_unpack_kernel(const unsigned long long* __restrict__ keys, long long n,
                                        const double* __restrict__ ew, int* __restrict__ other,
                                        double* __restrict__ wv) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const unsigned long long k = keys[t];
  other[t] = (int)(k & 0x7fffffffull);
  wv[t] = ew[k >> 32];
}

// R matrix (column-major, n_vertices x n_cols) <-> row-major working copy; 32 x 32 tiles through shared memory
__global__ void to_row_major_kernel(const double* __restrict__ cm, int n_vertices, int n_cols, double* __restrict__ rm) {
  __shared__ double tile[32][33];
  const int v0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int v = v0 + threadIdx.x, c = c0 + j;
    tile[j][threadIdx.x] = (v < n_vertices && c < n_cols) ? cm[(size_t)c * n_vertices + v] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int v = v0 + j, c = c0 + threadIdx.x;
    if (v < n_vertices && c < n_cols) rm[(size_t)v * n_cols + c] = tile[threadIdx.x][j];
  }
}
__global__ void to_col_major_kernel(const double* __restrict__ rm, int n_vertices, int n_cols, double* __restrict__ cm) {
  __shared__ double tile[32][33];
  const int v0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int v = v0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (v < n_vertices && c < n_cols) ? rm[(size_t)v * n_cols + c] : 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int v = v0 + threadIdx.x, c = c0 + j;
    if (v < n_vertices && c < n_cols) cm[(size_t)c * n_vertices + v] = tile[threadIdx.x][j];
  }
}

// one iteration: a warp per vertex, a lane per column (chunks of 32 columns); the row's incidence list is walked in
// edge order.  norm_bits receives max |new - old| (bit pattern of a non-negative double orders like an integer).
__global__ void __launch_bounds__(256)
diffuse_step_kernel(const long long* __restrict__ row_p, const int* __restrict__ other, const double* __restrict__ wv,
                    const unsigned char* __restrict__ fixed, const double* __restrict__ old_m,
                    double* __restrict__ new_m, int n_vertices, int n_cols, int normalize,
                    unsigned long long* __restrict__ norm_bits) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  double worst = 0.0;
  for (int v = blockIdx.x * warps_per_block + (threadIdx.x >> 5); v < n_vertices; v += gridDim.x * warps_per_block) {
    const bool fx = fixed != nullptr && fixed[v] != 0;
    const long long p0 = row_p[v], p1 = fx ? p0 : row_p[v + 1];
    const double* orow = old_m + (size_t)v * n_cols;
    double* nrow = new_m + (size_t)v * n_cols;
    for (int c = lane; c < n_cols; c += 32) {
      const double o = orow[c];
      double acc = o, sw = 1.0;
      long long t = p0;
      // eight neighbour rows in flight per lane (the gathers are L2 latency bound); the sums keep the edge order
      for (; t + 8 <= p1; t += 8) {
        double w[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          w[u] = wv[t + u];
          x[u] = old_m[(size_t)other[t + u] * n_cols + c];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc = __dadd_rn(acc, __dmul_rn(x[u], w[u]));
          sw = __dadd_rn(sw, w[u]);
        }
      }
      for (; t < p1; ++t) {
        const double w = wv[t];
        acc = __dadd_rn(acc, __dmul_rn(old_m[(size_t)other[t] * n_cols + c], w));
        sw = __dadd_rn(sw, w);
      }
      if (normalize) acc = __ddiv_rn(acc, sw);
      nrow[c] = acc;
      const double d = fabs(acc - o);
      worst = d > worst ? d : worst;   // NaN never wins, like Eigen's lpNorm<Infinity> on a max reduction
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, worst, s);
    worst = o > worst ? o : worst;
  }
  if (lane == 0 && worst > 0.0) atomicMax(norm_bits, (unsigned long long)__double_as_longlong(worst));
}

// label_probs (:83-93): one-hot rows of the labelled vertices, row-major
__global__ void one_hot_kernel(const int* __restrict__ label, int n_vertices, int n_labels, int fix,
                               double* __restrict__ m, unsigned char* __restrict__ fixed) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vertices) return;
  const int l = label[v];
  if (l >= 0 && l < n_labels) m[(size_t)v * n_labels + l] = 1.0;
  fixed[v] = (fix && l >= 0 && l < n_labels) ? 1 : 0;
}
// res(v, .) / sum(res(v, .))  (:100-106), left to right
__global__ void row_normalize_kernel(double* __restrict__ m, int n_vertices, int n_cols) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vertices) return;
  double* r = m + (size_t)v * n_cols;
  double s = 0.0;
  for (int c = 0; c < n_cols; ++c) s = __dadd_rn(s, r[c]);
  for (int c = 0; c < n_cols; ++c) r[c] = __ddiv_rn(r[c], s);
}

struct Incidence {
  DevBuf<long long> row_p;
  DevBuf<int> other;
  DevBuf<double> wv;
  long long n_entries = 0;
};

// device edge list -> per-vertex incidence lists in edge order
void build_incidence(Context& c, const int* d_from, const int* d_to, const double* d_w, long long n_edges,
                     int n_vertices, double fading, double fading_const, Incidence& inc) {
  cudaStream_t st = c.stream;
  DevBuf<double> mm(2), ew((size_t)std::max<long long>(n_edges, 1));
  DevBuf<int> deg((size_t)n_vertices + 1), flags(1);
  deg.zero(st);
  flags.zero(st);
  inc.row_p.alloc((size_t)n_vertices + 1);
  if (n_edges > 0) {
    const int n_part = (int)std::min<long long>(1024, (n_edges + 255) / 256);
    DevBuf<double> part((size_t)2 * n_part);
    minmax_partial_kernel<<<n_part, 256, 0, st>>>(d_w, n_edges, part.p);
    minmax_final_kernel<<<1, 1, 0, st>>>(part.p, n_part, mm.p);
    edge_weight_kernel<<<nblocks((size_t)n_edges), 256, 0, st>>>(d_from, d_to, d_w, n_edges, n_vertices, mm.p, fading,
                                                                fading_const, ew.p, deg.p, flags.p);
    c.stats.kernels += 3;
  }
  int h_flags = 0;
  flags.download(&h_flags, 1, st);
  exclusive_scan_i32_to_i64(c, deg.p, inc.row_p.p, (size_t)n_vertices + 1, nullptr);
  CG_CUDA(cudaStreamSynchronize(st));
  CG_CHECK(!(h_flags & 2), "edge refers to a vertex outside the matrix");
  CG_CHECK(!(h_flags & 1), "Negative edge length");  // Edge::Edge, src/propagate_labels.cpp:24-26
  inc.n_entries = 2 * n_edges;
  inc.other.alloc((size_t)std::max<long long>(inc.n_entries, 1));
  inc.wv.alloc((size_t)std::max<long long>(inc.n_entries, 1));
  if (n_edges > 0) {
    CG_CHECK(n_edges < (1ll << 31), "too many edges");
    DevBuf<unsigned long long> keys((size_t)inc.n_entries);
    deg.zero(st);  // reused as the fill cursors
    incidence_fill_kernel<<<nblocks((size_t)n_edges), 256, 0, st>>>(d_from, d_to, n_edges, inc.row_p.p, deg.p, keys.p);
    segmented_sort_u64(c, keys.p, inc.row_p.p, n_vertices);
    incidence_unpack_kernel<<<nblocks((size_t)inc.n_entries), 256, 0, st>>>(keys.p, inc.n_entries, ew.p, inc.other.p,
                                                                           inc.wv.p);
    c.stats.kernels += 2;
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaStreamSynchronize(st));  // keys / ew are released after their last use
  }
}

// smooth_count_matrix_c (:120-175) on a row-major device matrix; returns the iteration index the loop stopped at
int diffuse(Context& c, const Incidence& inc, const unsigned char* d_fixed, DevBuf<double>& cur, int n_vertices,
            int n_cols, int max_n_iters, double tol, int normalize, double* inf_norm_out) {
  cudaStream_t st = c.stream;
  DevBuf<double> nxt((size_t)n_vertices * n_cols);
  DevBuf<unsigned long long> norm(1);
  // one resident wave (8 CTAs of 8 warps per SM) with a grid-stride loop; 4x .. 32x and one warp per vertex without
  // a loop were measured equal or slower (0.37-0.50 ms per iteration on the configs[1] graph)
  const int grid = (int)std::min(nblocks((size_t)n_vertices, 8), (unsigned)(c.sm_count * 8));
  double inf_norm = 1e10;
  int iter = 0;
  for (iter = 0; iter < max_n_iters; ++iter) {
    norm.zero(st);
    diffuse_step_kernel<<<grid, 256, 0, st>>>(inc.row_p.p, inc.other.p, inc.wv.p, d_fixed, cur.p, nxt.p, n_vertices,
                                             n_cols, normalize, norm.p);
    c.stats.kernels++;
    unsigned long long bits = 0;
    norm.download(&bits, 1, st);
    CG_CUDA(cudaStreamSynchronize(st));
    memcpy(&inf_norm, &bits, sizeof(double));
    if (inf_norm < tol) break;  // the reference leaves the matrix of the previous iteration in place (:166-170)
    std::swap(cur.p, nxt.p);
  }
  CG_CUDA(cudaGetLastError());
  if (inf_norm_out) *inf_norm_out = inf_norm;
  return iter;
}

struct EdgeSource {
  DevBuf<int> from, to;
  DevBuf<double> w;
  const int* d_from = nullptr;
  const int* d_to = nullptr;
  const double* d_w = nullptr;
  long long n = 0;
};

// the graph handle's simplified edge list (already on the device) or a host edge list in R's 0-based vertex ids
void edge_source(Context& c, const cg_graph* g, const int32_t* from, const int32_t* to, const double* weight,
                 int64_t n_edges, int32_t n_vertices, EdgeSource& es) {
  if (g) {
    CG_CHECK(n_vertices == g->g.n_vertices, "matrix rows must follow the vertices of the graph");
    es.d_from = g->g.from.p;
    es.d_to = g->g.to.p;
    es.d_w = g->g.weight.p;
    es.n = g->g.n_edges;
    return;
  }
  CG_CHECK(n_edges == 0 || (from && to && weight), "NULL edge list");
  es.n = n_edges;
  if (n_edges > 0) {
    es.from.upload(from, (size_t)n_edges, c.stream);
    es.to.upload(to, (size_t)n_edges, c.stream);
    es.w.upload(weight, (size_t)n_edges, c.stream);
  }
  es.d_from = es.from.p;
  es.d_to = es.to.p;
  es.d_w = es.w.p;
}

}  // namespace
}  // namespace cg

using namespace cg;

extern "C" {

int cg_smooth_matrix(cg_context* ctx, const cg_graph* graph, const int32_t* from, const int32_t* to,
                     const double* weight, int64_t n_edges, int32_t n_vertices, double* matrix, int32_t n_cols,
                     const uint8_t* is_fixed, int32_t max_n_iters, double fading, double fading_const, double tol,
                     int32_t normalize, int32_t* n_iters, double* inf_norm) {
  CG_API_BEGIN
  CG_CHECK(ctx && n_vertices >= 0 && n_cols >= 0 && n_edges >= 0 && max_n_iters >= 0, "bad argument");
  if (n_iters) *n_iters = 0;
  if (inf_norm) *inf_norm = 1e10;
  if (n_vertices == 0 || n_cols == 0) return 0;  // "Empty matrix passed" (:122-125, :181-184)
  CG_CHECK(matrix != nullptr, "NULL matrix");
  CG_BIND(ctx);
  Context& c = ctx->c;
  StageScope sc(c, "diffusion");
  EdgeSource es;
  edge_source(c, graph, from, to, weight, n_edges, n_vertices, es);
  if (es.n == 0 && !graph) return 0;             // "Empty graph passed" (:189-192): the matrix comes back unchanged
  Incidence inc;
  build_incidence(c, es.d_from, es.d_to, es.d_w, es.n, n_vertices, fading, fading_const, inc);
  const size_t N = (size_t)n_vertices * n_cols;
  DevBuf<double> col(N), cur(N);
  DevBuf<unsigned char> fx;
  col.upload(matrix, N, c.stream);
  if (is_fixed) fx.upload(is_fixed, (size_t)n_vertices, c.stream);
  const dim3 tg((n_vertices + 31) / 32, (n_cols + 31) / 32), tb(32, 8);
  to_row_major_kernel<<<tg, tb, 0, c.stream>>>(col.p, n_vertices, n_cols, cur.p);
  double nrm = 1e10;
  const int it = diffuse(c, inc, is_fixed ? fx.p : nullptr, cur, n_vertices, n_cols, max_n_iters, tol, normalize, &nrm);
  to_col_major_kernel<<<tg, tb, 0, c.stream>>>(cur.p, n_vertices, n_cols, col.p);
  c.stats.kernels += 2;
  col.download(matrix, N, c.stream);
  CG_CUDA(cudaStreamSynchronize(c.stream));
  if (n_iters) *n_iters = it;
  if (inf_norm) *inf_norm = nrm;
  CG_API_END(ctx)
}

int cg_propagate_labels(cg_context* ctx, const cg_graph* graph, const int32_t* from, const int32_t* to,
                        const double* weight, int64_t n_edges, int32_t n_vertices, const int32_t* label_of_vertex,
                        int32_t n_labels, int32_t max_n_iters, double fading, double fading_const, double tol,
                        int32_t fixed_initial_labels, double* label_dist, int32_t* n_iters, double* inf_norm) {
  CG_API_BEGIN
  CG_CHECK(ctx && n_vertices >= 0 && n_labels >= 0 && n_edges >= 0 && max_n_iters >= 0, "bad argument");
  if (n_iters) *n_iters = 0;
  if (inf_norm) *inf_norm = 1e10;
  if (n_vertices == 0 || n_labels == 0) return 0;
  CG_CHECK(label_of_vertex && label_dist, "NULL argument");
  CG_BIND(ctx);
  Context& c = ctx->c;
  StageScope sc(c, "diffusion");
  EdgeSource es;
  edge_source(c, graph, from, to, weight, n_edges, n_vertices, es);
  Incidence inc;
  build_incidence(c, es.d_from, es.d_to, es.d_w, es.n, n_vertices, fading, fading_const, inc);
  const size_t N = (size_t)n_vertices * n_labels;
  DevBuf<double> cur(N), col(N);
  DevBuf<int> lab;
  DevBuf<unsigned char> fx((size_t)n_vertices);
  lab.upload(label_of_vertex, (size_t)n_vertices, c.stream);
  cur.zero(c.stream);
  one_hot_kernel<<<nblocks((size_t)n_vertices), 256, 0, c.stream>>>(lab.p, n_vertices, n_labels, fixed_initial_labels,
                                                                   cur.p, fx.p);
  double nrm = 1e10;
  const int it = diffuse(c, inc, fx.p, cur, n_vertices, n_labels, max_n_iters, tol, 1, &nrm);
  row_normalize_kernel<<<nblocks((size_t)n_vertices), 256, 0, c.stream>>>(cur.p, n_vertices, n_labels);
  const dim3 tg((n_vertices + 31) / 32, (n_labels + 31) / 32), tb(32, 8);
  to_col_major_kernel<<<tg, tb, 0, c.stream>>>(cur.p, n_vertices, n_labels, col.p);
  c.stats.kernels += 3;
  col.download(label_dist, N, c.stream);
  CG_CUDA(cudaStreamSynchronize(c.stream));
  if (n_iters) *n_iters = it;
  if (inf_norm) *inf_norm = nrm;
  CG_API_END(ctx)
}

}  // extern "C"
