#include <cmath>
#include "prims.cuh"
#include <cuda_fp16.h>

#include "sparse.cuh"
#include "gemm_tc.cuh"
#include <map>
#include <algorithm>

namespace cg {
namespace {

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

__global__ void row_count_kernel(const int* __restrict__ ci, long long nnz, int* __restrict__ cnt) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) atomicAdd(&cnt[ci[t]], 1);
}

// CSC -> CSR keys: (column << 32 | source position), bucketed by row; the segmented sort orders a row by column
__global__ void csr_scatter_kernel(const int* __restrict__ csc_p, const int* __restrict__ csc_i, int n_genes,
                                   const long long* __restrict__ rowp, int* __restrict__ cursor,
                                   unsigned long long* __restrict__ keys) {
  // one warp per gene column
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= n_genes) return;
  const int lane = threadIdx.x & 31;
  for (int t = csc_p[g] + lane; t < csc_p[g + 1]; t += 32) {
    int r = csc_i[t];
    int slot = atomicAdd(&cursor[r], 1);
    keys[rowp[r] + slot] = ((unsigned long long)(unsigned)g << 32) | (unsigned long long)(unsigned)t;
  }
}

__global__ void csr_fill_kernel(const unsigned long long* __restrict__ keys, long long nnz,
                                const double* __restrict__ csc_x, const long long* __restrict__ rowp64, int n_rows,
                                int* __restrict__ csr_i, double* __restrict__ csr_x) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  csr_i[t] = (int)(keys[t] >> 32);
  csr_x[t] = csc_x[keys[t] & 0xffffffffull];
}

__global__ void rowp_narrow_kernel(const long long* __restrict__ in, int n, int* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (int)in[t];
}

// deterministic column sums: one warp per column, fixed lane stride and shuffle tree
// out[g] = column sum, out[n_genes + g] = largest |value| of the column (decides whether the Gram operand fits fp16)
__global__ void colsum_kernel(const int* __restrict__ csc_p, const double* __restrict__ csc_x, int n_genes,
                              double* __restrict__ out) {
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= n_genes) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0, mx = 0.0;
  for (int t = csc_p[g] + lane; t < csc_p[g + 1]; t += 32) {
    const double v = csc_x[t];
    s += v;
    mx = fmax(mx, fabs(v));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, d);
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, d));
  }
  if (lane == 0) {
    out[g] = s;
    out[n_genes + g] = mx;
  }
}

// ------------------------------------------------------------------ projection operands
__global__ void wbuild_kernel(const WBuildProblem* __restrict__ probs, int d, int d_pad) {
  const WBuildProblem pb = probs[blockIdx.y];
  // zero W, then scatter the od-gene rows (two phases separated by a grid-wide kernel boundary would be
  // cleaner; here every CTA zeroes and fills disjoint column ranges of the SAME od genes, so no race:
  // CTA b owns od genes [b*chunk, (b+1)*chunk) and the W rows they map to are distinct per gene)
  const int chunk = (pb.G + gridDim.x - 1) / gridDim.x;
  const int g0 = blockIdx.x * chunk, g1 = min(pb.G, g0 + chunk);
  for (int g = g0 + (threadIdx.x / d_pad); g < g1; g += blockDim.x / d_pad) {
    const int c = threadIdx.x % d_pad;
    const double v = c < d ? pb.f[g] * pb.rot[(size_t)c * pb.G + g] : 0.0;
    pb.W[(size_t)pb.cols[g] * d_pad + c] = v;
  }
  if (blockIdx.x == 0) {
    // shift[c] = sum_g cvec[g] * rot[g, c]: one warp per component, fixed order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = warp; c < d_pad; c += nw) {
      double s = 0.0;
      if (c < d)
        for (int g = lane; g < pb.G; g += 32) s += pb.cvec[g] * pb.rot[(size_t)c * pb.G + g];
#pragma unroll
      for (int dl = 16; dl > 0; dl >>= 1) s += __shfl_down_sync(0xffffffffu, s, dl);
      if (lane == 0) pb.shift[c] = s;
    }
  }
}

// one warp per cell row: fp64 accumulate over the row's non-zeros, subtract the shift, cast to fp32,
// canonical normalisation (sequential fma chain over the d components), write the tiled kNN panel
template <int D_PAD>
__global__ void __launch_bounds__(256) project_kernel(const ProjProblem* __restrict__ probs, int d, int metric) {
  const ProjProblem pb = probs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n_pad = (int)((pb.n_rows + KNN_ROW_PAD - 1) / KNN_ROW_PAD * KNN_ROW_PAD);
  if (row >= n_pad) return;
  constexpr int NC = D_PAD / 32;  // components per lane
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  if (row < pb.n_rows) {
    const int b = pb.rp[row], e = pb.rp[row + 1];
    for (int t = b; t < e; ++t) {
      const double x = pb.xv[t];
      const double* w = pb.W + (size_t)pb.ci[t] * D_PAD;
#pragma unroll
      for (int u = 0; u < NC; ++u) acc[u] = fma(x, w[lane + 32 * u], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] -= pb.shift[lane + 32 * u];
  }
  float v[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) {
    const int c = lane + 32 * u;
    v[u] = (row < pb.n_rows && c < d) ? (float)acc[u] : 0.0f;
    if (pb.p64 && row < pb.n_rows && c < d) pb.p64[(size_t)c * pb.n_rows + row] = acc[u];
  }
  float inv = 1.0f;
  if (metric == METRIC_ANGULAR) {
    float ss = 0.0f;
    for (int i = 0; i < d; ++i) {
      const float xi = __shfl_sync(0xffffffffu, v[i >> 5], i & 31);
      ss = fmaf(xi, xi, ss);
    }
    inv = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  // exact fp32 panel, then its fp16 copy (the tensor-core operand of the two-phase kNN kernel)
  float* dst = pb.tiled + (size_t)(row >> 3) * (8 * D_PAD) + (size_t)(row & 7) * 4;
  __half* dst16 = reinterpret_cast<__half*>(pb.tiled + (size_t)n_pad * D_PAD);
  float* dstrm = pb.tiled + (size_t)n_pad * D_PAD * 3 / 2 + (size_t)row * D_PAD;  // row-major exact copy
#pragma unroll
  for (int u = 0; u < NC; ++u) {
    const int c = lane + 32 * u;
    const float x = (d < D_PAD && c == D_PAD - 1) ? -1.0f : v[u] * inv;
    dst[(c >> 2) * 32 + (c & 3)] = x;
    dst16[knn_tiled16_offset(row, c, D_PAD)] = __float2half_rn(x);
    dstrm[c] = x;
  }
}

}  // namespace

void csc_to_csr(Context& ctx, int n_rows, int n_cols, const int* csc_p, const int* csc_i, const double* csc_x,
                long long nnz, int* csr_p, int* csr_i, double* csr_x) {
  cudaStream_t st = ctx.stream;
  if (nnz == 0) {
    CG_CUDA(cudaMemsetAsync(csr_p, 0, ((size_t)n_rows + 1) * sizeof(int), st));
    return;
  }
  DevBuf<int> cnt((size_t)n_rows), cursor((size_t)n_rows);
  cnt.zero(st);
  cursor.zero(st);
  DevBuf<long long> rowp64((size_t)n_rows + 1);
  row_count_kernel<<<nblk((size_t)nnz), 256, 0, st>>>(csc_i, nnz, cnt.p);
  exclusive_scan_i32_to_i64(ctx, cnt.p, rowp64.p, (size_t)n_rows, rowp64.p + n_rows);
  DevBuf<unsigned long long> keys((size_t)nnz);
  csr_scatter_kernel<<<(n_cols + 7) / 8, 256, 0, st>>>(csc_p, csc_i, n_cols, rowp64.p, cursor.p, keys.p);
  segmented_sort_u64(ctx, keys.p, rowp64.p, n_rows);
  csr_fill_kernel<<<nblk((size_t)nnz), 256, 0, st>>>(keys.p, nnz, csc_x, rowp64.p, n_rows, csr_i, csr_x);
  rowp_narrow_kernel<<<nblk((size_t)n_rows + 1), 256, 0, st>>>(rowp64.p, n_rows + 1, csr_p);
  ctx.stats.kernels += 4;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(st));
}

// Two halves so that a panel's uploads pipeline: sample_upload enqueues the host->device copies of one sample on
// `copy_st` (its own stream when the caller has one; the buffers are allocated on the context's stream first and
// handed over through an event) and records `ready`; sample_finish makes the context's stream wait for `ready` and
// builds the derived arrays (CSR by cell, column sums).  With pinned host memory the copies of sample s+1 run
// while sample s is converted.
void sample_upload(Context& ctx, Sample& s, int n_cells, int n_genes, const int* p, const int* i, const double* x,
                   const double* var_v, const double* var_qv, const double* pca, int n_pcs, int pca_ld,
                   cudaStream_t copy_st, cudaEvent_t ready) {
  cudaStream_t st = ctx.stream;
  if (!copy_st) copy_st = st;
  s.has_counts = p != nullptr;
  CG_CHECK(n_cells > 0 && (n_genes > 0 || !s.has_counts), "sample must have cells and genes");
  s.n_cells = n_cells;
  s.n_genes = s.has_counts ? n_genes : 0;
  s.nnz = s.has_counts ? p[n_genes] : 0;
  if (s.has_counts) {
    CG_CHECK(i != nullptr && x != nullptr, "sample counts matrix incomplete");
    for (int g = 0; g < n_genes; ++g) CG_CHECK(p[g + 1] >= p[g], "dgCMatrix @p must be non-decreasing");
    s.csc_p.alloc((size_t)n_genes + 1);
    s.csc_i.alloc((size_t)s.nnz);
    s.csc_x.alloc((size_t)s.nnz);
    s.csr_p.alloc((size_t)n_cells + 1);
    s.csr_i.alloc((size_t)s.nnz);
    s.csr_x.alloc((size_t)s.nnz);
    s.colsum.alloc((size_t)2 * n_genes);  // sums, then column maxima
  }
  s.n_pcs = 0;
  s.pca_metric = -1;
  if (pca && n_pcs > 0) {
    s.n_pcs = n_pcs;
    s.pca.alloc((size_t)n_cells * n_pcs);
  }
  if (copy_st != st) {  // stream-ordered allocations: usable on another stream only after this point of `st`
    cudaEvent_t ea;
    CG_CUDA(cudaEventCreateWithFlags(&ea, cudaEventDisableTiming));
    CG_CUDA(cudaEventRecord(ea, st));
    CG_CUDA(cudaStreamWaitEvent(copy_st, ea, 0));
    CG_CUDA(cudaEventDestroy(ea));
  }
  if (s.has_counts) {
    s.csc_p.upload(p, (size_t)n_genes + 1, copy_st);
    s.csc_i.upload(i, (size_t)s.nnz, copy_st);
    s.csc_x.upload(x, (size_t)s.nnz, copy_st);
    if (var_v) s.h_v.assign(var_v, var_v + n_genes);
    if (var_qv) s.h_qv.assign(var_qv, var_qv + n_genes);
  }
  s.h_logqv.resize(s.h_qv.size());
  s.h_expv.resize(s.h_v.size());
  for (size_t g = 0; g < s.h_qv.size(); ++g) s.h_logqv[g] = std::log(s.h_qv[g]);
  for (size_t g = 0; g < s.h_v.size(); ++g) s.h_expv[g] = std::exp(s.h_v[g]);
  if (s.n_pcs > 0) {
    if (pca_ld == n_cells) {
      s.pca.upload(pca, (size_t)n_cells * n_pcs, copy_st);
    } else {
      CG_CUDA(cudaMemcpy2DAsync(s.pca.p, sizeof(double) * (size_t)n_cells, pca, sizeof(double) * (size_t)pca_ld,
                                sizeof(double) * (size_t)n_cells, (size_t)n_pcs, cudaMemcpyHostToDevice, copy_st));
    }
  }
  if (ready) CG_CUDA(cudaEventRecord(ready, copy_st));
}

void sample_finish(Context& ctx, Sample& s, cudaEvent_t ready) {
  cudaStream_t st = ctx.stream;
  if (ready) CG_CUDA(cudaStreamWaitEvent(st, ready, 0));
  if (!s.has_counts) return;
  csc_to_csr(ctx, s.n_cells, s.n_genes, s.csc_p.p, s.csc_i.p, s.csc_x.p, s.nnz, s.csr_p.p, s.csr_i.p, s.csr_x.p);
  colsum_kernel<<<(s.n_genes + 7) / 8, 256, 0, st>>>(s.csc_p.p, s.csc_x.p, s.n_genes, s.colsum.p);
  ctx.stats.kernels++;
  s.h_colsum.resize((size_t)2 * s.n_genes);
  s.colsum.download(s.h_colsum.data(), (size_t)2 * s.n_genes, st);
  CG_CUDA(cudaStreamSynchronize(st));
  s.max_abs = 0.0;
  for (int g = 0; g < s.n_genes; ++g) s.max_abs = std::max(s.max_abs, s.h_colsum[(size_t)s.n_genes + g]);
  s.h_colsum.resize(s.n_genes);
}

void build_projection_operands(Context& ctx, const WBuildProblem* h_probs, int n_probs, int d, int d_pad,
                               bool zeroed) {
  if (n_probs == 0) return;
  DevBuf<WBuildProblem> dp;
  dp.upload(h_probs, n_probs, ctx.stream);
  if (!zeroed)   // the caller may have cleared all operands with one memset (they are slices of one allocation)
    for (int p = 0; p < n_probs; ++p)
      CG_CUDA(cudaMemsetAsync(h_probs[p].W, 0, (size_t)h_probs[p].n_genes_sample * d_pad * sizeof(double), ctx.stream));
  dim3 grid(32, n_probs);
  wbuild_kernel<<<grid, 256, 0, ctx.stream>>>(dp.p, d, d_pad);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

void project_and_prep(Context& ctx, const ProjProblem* h_probs, int n_probs, int d, int d_pad, Metric metric) {
  if (n_probs == 0) return;
  DevBuf<ProjProblem> dp;
  dp.upload(h_probs, n_probs, ctx.stream);
  int max_rows = 0;
  for (int p = 0; p < n_probs; ++p) max_rows = h_probs[p].n_rows > max_rows ? h_probs[p].n_rows : max_rows;
  int max_pad = (int)round_up(max_rows, KNN_ROW_PAD);
  for (int p0 = 0; p0 < n_probs; p0 += 65535) {
    int np = n_probs - p0 < 65535 ? n_probs - p0 : 65535;
    dim3 grid((max_pad + 7) / 8, np);
    if (d_pad == 32) project_kernel<32><<<grid, 256, 0, ctx.stream>>>(dp.p + p0, d, (int)metric);
    else if (d_pad == 64) project_kernel<64><<<grid, 256, 0, ctx.stream>>>(dp.p + p0, d, (int)metric);
    else project_kernel<128><<<grid, 256, 0, ctx.stream>>>(dp.p + p0, d, (int)metric);
    ctx.stats.kernels++;
  }
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
}


// ---------------------------------------------------------------------------------------- tensor-core projection
namespace {

// dense bf16 hi/lo operand with rows = cells, k = genes, built tile-wise from the CSR form: a CTA owns 256 cells x
// 64 genes (32 KB contiguous in the operand layout); every thread drops the non-zeros of one cell row that fall into
// the gene range into the zeroed shared-memory image, which leaves with coalesced 16-byte stores.
__global__ void __launch_bounds__(256) cells_operand_kernel(const int* __restrict__ csr_p, const int* __restrict__ csr_i,
                                                            const double* __restrict__ csr_x, int n_cells, int rows_pad,
                                                            __nv_bfloat16* __restrict__ hi,
                                                            __nv_bfloat16* __restrict__ lo) {
  extern __shared__ __align__(16) unsigned char co_smem[];
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(co_smem);
  __nv_bfloat16* sl = sh + 256 * 64;
  const int kb = blockIdx.y, r0 = blockIdx.x * 256, k0 = kb * 64;
  const int tid = threadIdx.x;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int t = tid; t < 256 * 64 / 8; t += 256) {
    reinterpret_cast<uint4*>(sh)[t] = z;
    reinterpret_cast<uint4*>(sl)[t] = z;
  }
  __syncthreads();
  const int c = r0 + tid;
  if (c < n_cells) {
    const int b = csr_p[c], e = csr_p[c + 1];
    int lo_i = b, hi_i = e;  // first entry with column >= k0
    while (lo_i < hi_i) {
      const int mid = (lo_i + hi_i) >> 1;
      if (csr_i[mid] < k0) lo_i = mid + 1; else hi_i = mid;
    }
    for (int t = lo_i; t < e; ++t) {
      const int col = csr_i[t];
      if (col >= k0 + 64) break;
      const float x = (float)csr_x[t];
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const int kk = col - k0;
      const int o = ((tid >> 3) * 8 + (kk >> 3)) * 64 + (tid & 7) * 8 + (kk & 7);
      sh[o] = h;
      sl[o] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
  }
  __syncthreads();
  const size_t base = ((size_t)kb * (size_t)(rows_pad >> 3) + (size_t)(r0 >> 3)) * 512;
  for (int t = tid; t < 256 * 64 / 8; t += 256) {
    reinterpret_cast<uint4*>(hi + base)[t] = reinterpret_cast<const uint4*>(sh)[t];
    reinterpret_cast<uint4*>(lo + base)[t] = reinterpret_cast<const uint4*>(sl)[t];
  }
}

struct WSide {
  const double* W;  // [n_genes][32]
};

// B operand of the projection GEMM: rows = 32 * side + component (256 rows = up to 8 sides), k = gene
__global__ void w_operand_kernel(const WSide* __restrict__ sides, int n_sides, int n_genes, int k_pad,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)256 * k_pad) return;
  const int r = (int)(t & 255), kk = (int)(t >> 8);
  const int side = r >> 5, comp = r & 31;
  float x = 0.f;
  if (side < n_sides && kk < n_genes) x = (float)sides[side].W[(size_t)kk * 32 + comp];
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const size_t o = tiled_operand_offset(r, kk, 256);
  hi[o] = h;
  lo[o] = __float2bfloat16_rn(x - __bfloat162float(h));
}

struct ProjFinish {
  const float* C;       // GEMM output, column-major n_rows x 256: this side's 32 columns start at C
  const double* shift;  // [32]
  float* tiled;         // the kNN panel (three copies)
  int n_rows;
};

// one thread per cell row: shift, cast, canonical normalisation, the three panel copies (same tail as project_kernel)
__global__ void __launch_bounds__(128) project_finish_kernel(const ProjFinish* __restrict__ probs, int d, int metric) {
  constexpr int D_PAD = 32;
  const ProjFinish pb = probs[blockIdx.y];
  const int n_pad = (int)((pb.n_rows + KNN_ROW_PAD - 1) / KNN_ROW_PAD * KNN_ROW_PAD);
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_pad) return;
  float v[D_PAD];
#pragma unroll
  for (int c = 0; c < D_PAD; ++c)
    v[c] = (row < pb.n_rows && c < d) ? (float)((double)pb.C[(size_t)c * pb.n_rows + row] - pb.shift[c]) : 0.0f;
  float inv = 1.0f;
  if (metric == METRIC_ANGULAR) {
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < D_PAD; ++c)
      if (c < d) ss = fmaf(v[c], v[c], ss);
    inv = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  float* dst = pb.tiled + (size_t)(row >> 3) * (8 * D_PAD) + (size_t)(row & 7) * 4;
  __half* dst16 = reinterpret_cast<__half*>(pb.tiled + (size_t)n_pad * D_PAD);
  float* dstrm = pb.tiled + (size_t)n_pad * D_PAD * 3 / 2 + (size_t)row * D_PAD;
#pragma unroll
  for (int c = 0; c < D_PAD; ++c) {
    const float x = (d < D_PAD && c == D_PAD - 1) ? -1.0f : v[c] * inv;
    dst[(c >> 2) * 32 + (c & 3)] = x;
    dst16[knn_tiled16_offset(row, c, D_PAD)] = __float2half_rn(x);
    dstrm[c] = x;
  }
}

}  // namespace

void project_and_prep_tc(Context& ctx, const ProjProblem* h_probs, const int* sample_of, const int* n_genes_of,
                         int n_probs, int d, Metric metric) {
  if (n_probs == 0) return;
  cudaStream_t st = ctx.stream;
  // sides grouped by sample (stable)
  std::map<int, std::vector<int>> by_sample;
  for (int p = 0; p < n_probs; ++p) by_sample[sample_of[p]].push_back(p);
  constexpr int co_smem = 2 * 256 * 64 * (int)sizeof(__nv_bfloat16);
  CG_CUDA(cudaFuncSetAttribute(cells_operand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, co_smem));
  // one set of operand / output buffers sized for the largest sample, reused in stream order
  size_t max_x = 0, max_w = 0, max_c = 0, n_chunks = 0;
  for (auto& kv : by_sample) {
    const size_t rp = round_up(h_probs[kv.second[0]].n_rows, GEMM_ROW_PAD);
    const size_t kp = round_up(n_genes_of[kv.second[0]], GEMM_K_PAD);
    max_x = std::max(max_x, rp * kp);
    max_w = std::max(max_w, (size_t)256 * kp);
    max_c = std::max(max_c, (size_t)h_probs[kv.second[0]].n_rows * 256);
    n_chunks += (kv.second.size() + 7) / 8;
  }
  DevBuf<__nv_bfloat16> xh(max_x), xl(max_x), wh(max_w * n_chunks), wl(max_w * n_chunks);
  DevBuf<float> Cbuf(max_c);
  // the side tables of every chunk travel in one upload each
  std::vector<WSide> hw;
  std::vector<ProjFinish> hf;
  for (auto& kv : by_sample)
    for (size_t s0 = 0; s0 < kv.second.size(); s0 += 8) {
      const int ns = (int)std::min<size_t>(8, kv.second.size() - s0);
      for (int j = 0; j < 8; ++j) {
        if (j < ns) {
          const ProjProblem& pp = h_probs[kv.second[s0 + j]];
          hw.push_back(WSide{pp.W});
          hf.push_back(ProjFinish{Cbuf.p + (size_t)j * 32 * pp.n_rows, pp.shift, pp.tiled, pp.n_rows});
        } else {
          hw.push_back(WSide{nullptr});
          hf.push_back(ProjFinish{nullptr, nullptr, nullptr, 0});
        }
      }
    }
  DevBuf<WSide> dws;
  DevBuf<ProjFinish> dfs;
  dws.upload(hw.data(), hw.size(), st);
  dfs.upload(hf.data(), hf.size(), st);
  size_t chunk = 0;
  for (auto& kv : by_sample) {
    const std::vector<int>& sides = kv.second;
    const ProjProblem& first = h_probs[sides[0]];
    const int n_cells = first.n_rows, n_genes = n_genes_of[sides[0]];
    TiledOperand opx;
    opx.rows = n_cells;
    opx.k = n_genes;
    opx.rows_pad = (int)round_up(n_cells, GEMM_ROW_PAD);
    opx.k_pad = (int)round_up(n_genes, GEMM_K_PAD);
    opx.hi = xh.p;
    opx.lo = xl.p;
    cells_operand_kernel<<<dim3((unsigned)(opx.rows_pad / 256), (unsigned)(opx.k_pad / 64)), 256, co_smem, st>>>(
        first.rp, first.ci, first.xv, n_cells, opx.rows_pad, opx.hi, opx.lo);
    ctx.stats.kernels++;
    // the W operands of all chunks of this sample first, so that the GEMMs and finishes run back to back
    for (size_t s0 = 0; s0 < sides.size(); s0 += 8, ++chunk) {
      const int ns = (int)std::min<size_t>(8, sides.size() - s0);
      __nv_bfloat16* whc = wh.p + chunk * max_w;
      __nv_bfloat16* wlc = wl.p + chunk * max_w;
      w_operand_kernel<<<(unsigned)(((size_t)256 * opx.k_pad + 255) / 256), 256, 0, st>>>(dws.p + chunk * 8, ns, n_genes,
                                                                                         opx.k_pad, whc, wlc);
      ctx.stats.kernels++;
      GemmProblem gp{opx.hi, opx.lo, whc, wlc, opx.rows_pad, 256, opx.k_pad, n_cells, 32 * ns, Cbuf.p,
                     (long long)n_cells, 0};
      GemmPlan plan;  // stream-ordered: the tables are freed behind the kernel, no host synchronisation per sample
      gemm_plan(ctx, &gp, 1, plan);
      gemm_run(ctx, plan);
      const int n_pad = (int)round_up(n_cells, KNN_ROW_PAD);
      project_finish_kernel<<<dim3((unsigned)((n_pad + 127) / 128), (unsigned)ns), 128, 0, st>>>(dfs.p + chunk * 8, d,
                                                                                                 (int)metric);
      ctx.stats.kernels++;
    }
  }
  CG_CUDA(cudaGetLastError());
}

}  // namespace cg
