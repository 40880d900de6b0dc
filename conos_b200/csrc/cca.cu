// quickCCA on the device (R/conos.R:327-369, PMA = FALSE), without ever forming the dense n1 x n2 cross product
// sm1 %*% t(sm2) of the reference (:344):  with Ya, Yb the column-centred scaled matrices of the kept cells,
//   A = Ya Yb',   A A' = Ya Kb Ya',   Kb = Yb'Yb  (G x G),
// so the left singular vectors of A are the leading eigenvectors of the implicit n1 x n1 operator Ya Kb Ya', applied
// as  U -> Ya' U (sparse, by gene) -> Kb (.) (dense G x G) -> Ya (.) (sparse, by cell), with the centring folded in
// as rank-one terms.  v = A'u / sigma, then both are scaled by sqrt(sigma) / max sqrt(sigma) (:363-366).
#include <algorithm>
#include <cmath>

#include "eig.cuh"
#include "prims.cuh"
#include "reduce.cuh"

namespace cg {
namespace {

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// fcol[col] = f[g] for the pair's od genes, 0 elsewhere; gidx[col] = g or -1
__global__ void cca_colmap_kernel(const int* __restrict__ cols, const double* __restrict__ f, int G,
                                  double* __restrict__ fcol, int* __restrict__ gidx) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  fcol[cols[g]] = f[g];
  gidx[cols[g]] = g;
}

// keep[i] = rowSums(scaled m) > 0    (R/conos.R:336)
__global__ void cca_keep_kernel(const int* __restrict__ rp, const int* __restrict__ ci, const double* __restrict__ xv,
                                const double* __restrict__ fcol, int n, int* __restrict__ keep) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  double s = 0.0;
  for (int t = rp[row] + lane; t < rp[row + 1]; t += 32) s += xv[t] * fcol[ci[t]];
  s = warp_sum(s);
  if (lane == 0) keep[row] = s > 0.0 ? 1 : 0;
}

__global__ void cca_rows_kernel(const int* __restrict__ keep, const int* __restrict__ pos, int n, int* __restrict__ rows,
                                int* __restrict__ pos_or_neg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keep[i]) { rows[pos[i]] = i; pos_or_neg[i] = pos[i]; } else pos_or_neg[i] = -1;
}

// K = D (M[cols, cols] - colsum colsum' / n_kept) D   (= Y'Y of the centred scaled kept cells), zero padded
__global__ void cca_gram_kernel(const double* __restrict__ M, int Gs, const double* __restrict__ colsum,
                                const int* __restrict__ cols, const double* __restrict__ f, int G, double n_kept,
                                double* __restrict__ K) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * G) return;
  const int i = (int)(t % G), j = (int)(t / G);
  const int ci = cols[i], cj = cols[j];
  K[t] = f[i] * f[j] * (M[(size_t)cj * Gs + ci] - colsum[ci] * colsum[cj] / n_kept);
}

// column sums of a row-major [n][B] block: out[b] = sum_i U[i][b]   (single CTA, deterministic)
template <int B>
__global__ void cca_colsum_kernel(const double* __restrict__ U, int n, double* __restrict__ out) {
  __shared__ double red[8][B];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  for (int i = warp; i < n; i += 8)
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] += U[(size_t)i * B + lane + 32 * u];
#pragma unroll
  for (int u = 0; u < NC; ++u) red[warp][lane + 32 * u] = acc[u];
  __syncthreads();
  if (warp == 0)
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w][lane + 32 * u];
      out[lane + 32 * u] = s;
    }
}

// W[g][:] = f[g] * sum_{kept i in column} x_ig U[pos(i)][:]  -  mean[g] * colsumU[:]      (Ya' U, one warp per od gene)
template <int B>
__global__ void __launch_bounds__(256) cca_yt_u_kernel(const int* __restrict__ cp, const int* __restrict__ ci,
                                                       const double* __restrict__ cx, const int* __restrict__ cols,
                                                       const double* __restrict__ f, const double* __restrict__ mean,
                                                       const int* __restrict__ pos, const double* __restrict__ U,
                                                       const double* __restrict__ colsumU, int G,
                                                       double* __restrict__ W) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= G) return;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  const int col = cols[g];
  for (int t = cp[col]; t < cp[col + 1]; ++t) {
    const int p = pos[ci[t]];
    if (p < 0) continue;
    const double x = cx[t];
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] = fma(x, U[(size_t)p * B + lane + 32 * u], acc[u]);
  }
#pragma unroll
  for (int u = 0; u < NC; ++u)
    W[(size_t)g * B + lane + 32 * u] = f[g] * acc[u] - mean[g] * colsumU[lane + 32 * u];
}

// Z = K W  (K symmetric column-major G x G, W row-major [G][B]): one warp per row
template <int B>
__global__ void __launch_bounds__(256) cca_kw_kernel(const double* __restrict__ K, const double* __restrict__ W, int G,
                                                     double* __restrict__ Z) {
  __shared__ double crow[8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  const double* col = K + (size_t)(r < G ? r : 0) * G;
  for (int c0 = 0; c0 < G; c0 += 128) {
    __syncwarp();
    for (int t = lane; t < 128; t += 32) crow[warp][t] = (c0 + t < G) ? col[c0 + t] : 0.0;
    __syncwarp();
    const int lim = G - c0 < 128 ? G - c0 : 128;
    for (int t = 0; t < lim; ++t) {
      const double cv = crow[warp][t];
#pragma unroll
      for (int u = 0; u < NC; ++u) acc[u] = fma(cv, W[(size_t)(c0 + t) * B + lane + 32 * u], acc[u]);
    }
  }
  if (r < G)
#pragma unroll
    for (int u = 0; u < NC; ++u) Z[(size_t)r * B + lane + 32 * u] = acc[u];
}

// shift[b] = sum_g mean[g] W[g][b]  (single CTA)
template <int B>
__global__ void cca_shift_kernel(const double* __restrict__ mean, const double* __restrict__ W, int G,
                                 double* __restrict__ shift) {
  __shared__ double red[8][B];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  for (int g = warp; g < G; g += 8)
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] = fma(mean[g], W[(size_t)g * B + lane + 32 * u], acc[u]);
#pragma unroll
  for (int u = 0; u < NC; ++u) red[warp][lane + 32 * u] = acc[u];
  __syncthreads();
  if (warp == 0)
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w][lane + 32 * u];
      shift[lane + 32 * u] = s;
    }
}

// out[r][:] = sum_nnz x * f[g] * W[g][:] - shift[:]   over the kept rows (Y W, one warp per kept row)
template <int B>
__global__ void __launch_bounds__(256) cca_y_w_kernel(const int* __restrict__ rp, const int* __restrict__ ci,
                                                      const double* __restrict__ xv, const int* __restrict__ gidx,
                                                      const double* __restrict__ fcol, const int* __restrict__ rows,
                                                      int n_kept, const double* __restrict__ W,
                                                      const double* __restrict__ shift, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n_kept) return;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  const int row = rows[r];
  for (int t = rp[row]; t < rp[row + 1]; ++t) {
    const int g = gidx[ci[t]];
    if (g < 0) continue;
    const double x = xv[t] * fcol[ci[t]];
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] = fma(x, W[(size_t)g * B + lane + 32 * u], acc[u]);
  }
#pragma unroll
  for (int u = 0; u < NC; ++u) out[(size_t)r * B + lane + 32 * u] = acc[u] - shift[lane + 32 * u];
}

// column-major n x d  <->  row-major [n][B] helpers
__global__ void cca_cm_to_rm_kernel(const double* __restrict__ cm, int n, int d, int B, double* __restrict__ rm) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * B) return;
  const int i = (int)(t / B), j = (int)(t % B);
  rm[t] = j < d ? cm[(size_t)j * n + i] : 0.0;
}
// out (column-major n x d) = rm[:, :d] * scale[j]
__global__ void cca_rm_to_cm_scaled_kernel(const double* __restrict__ rm, int n, int d, int B,
                                           const double* __restrict__ scale, double* __restrict__ cm) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * d) return;
  const int i = (int)(t % n), j = (int)(t / n);
  cm[t] = rm[(size_t)i * B + j] * scale[j];
}

struct SideData {
  const Sample* s;
  DevBuf<double> f, fcol, mean, K;
  DevBuf<int> cols, gidx, keep, pos, rows, posmap;
  int n_kept = 0;
};

template <int B>
void cca_pair(Context& ctx, cg_panel& panel, const cg_pair& pr, const cg_params& P, PairReduction& R) {
  cudaStream_t st = ctx.stream;
  const int G = pr.n_genes;
  const int d = std::min(P.ncomps, G - 1);  // R/conos.R:334
  SideData S[2];
  S[0].s = &panel.samples[pr.a];
  S[1].s = &panel.samples[pr.b];
  const int* gcols[2] = {pr.genes_a, pr.genes_b};
  std::vector<double> f[2];
  f[0].assign(G, 1.0);
  f[1].assign(G, 1.0);
  if (P.var_scale) {
    CG_CHECK(!S[0].s->h_qv.empty() && !S[1].s->h_qv.empty() && !S[0].s->h_v.empty() && !S[1].s->h_v.empty(),
             "var.scale = TRUE needs misc$varinfo$v and $qv for every sample");
    for (int g = 0; g < G; ++g) {
      const int ca = pr.genes_a[g], cb = pr.genes_b[g];
      const double cqv = std::exp((S[0].s->h_logqv[ca] + S[1].s->h_logqv[cb]) / 2.0);
      const double fa = std::sqrt(cqv / S[0].s->h_expv[ca]);
      const double fb = std::sqrt(cqv / S[1].s->h_expv[cb]);
      f[0][g] = std::isfinite(fa) ? fa : 0.0;
      f[1][g] = std::isfinite(fb) ? fb : 0.0;
    }
  }
  for (int side = 0; side < 2; ++side) {
    SideData& D = S[side];
    const Sample& sm = *D.s;
    D.f.upload(f[side].data(), (size_t)G, st);
    D.cols.upload(gcols[side], (size_t)G, st);
    D.fcol.alloc((size_t)sm.n_genes);
    D.gidx.alloc((size_t)sm.n_genes);
    D.fcol.zero(st);
    CG_CUDA(cudaMemsetAsync(D.gidx.p, 0xff, (size_t)sm.n_genes * sizeof(int), st));
    cca_colmap_kernel<<<nblk((size_t)G), 256, 0, st>>>(D.cols.p, D.f.p, G, D.fcol.p, D.gidx.p);
    D.keep.alloc((size_t)sm.n_cells);
    D.pos.alloc((size_t)sm.n_cells + 1);
    D.posmap.alloc((size_t)sm.n_cells);
    cca_keep_kernel<<<(sm.n_cells + 7) / 8, 256, 0, st>>>(sm.csr_p.p, sm.csr_i.p, sm.csr_x.p, D.fcol.p, sm.n_cells,
                                                          D.keep.p);
    exclusive_scan_i32(ctx, D.keep.p, D.pos.p, (size_t)sm.n_cells, D.pos.p + sm.n_cells);
    CG_CUDA(cudaMemcpyAsync(&D.n_kept, D.pos.p + sm.n_cells, sizeof(int), cudaMemcpyDeviceToHost, st));
    CG_CUDA(cudaStreamSynchronize(st));
    CG_CHECK(D.n_kept >= B, "too few expressing cells for the device CCA");
    D.rows.alloc((size_t)D.n_kept);
    cca_rows_kernel<<<nblk((size_t)sm.n_cells), 256, 0, st>>>(D.keep.p, D.pos.p, sm.n_cells, D.rows.p, D.posmap.p);
    // column means over the kept cells (dropped rows are all zero on the od genes)
    std::vector<double> mean((size_t)G);
    for (int g = 0; g < G; ++g) mean[g] = f[side][g] * sm.h_colsum[gcols[side][g]] / (double)D.n_kept;
    D.mean.upload(mean.data(), (size_t)G, st);
    D.K.alloc((size_t)G * G);
    cca_gram_kernel<<<nblk((size_t)G * G), 256, 0, st>>>(sm.gram.p, sm.n_genes, sm.colsum.p, D.cols.p, D.f.p, G,
                                                         (double)D.n_kept, D.K.p);
    ctx.stats.kernels += 4;
  }
  CG_CUDA(cudaGetLastError());
  const int n1 = S[0].n_kept, n2 = S[1].n_kept;
  DevBuf<double> csU(B), W1((size_t)G * B), W2((size_t)G * B), shift(B);
  // Z = Ya Kb Ya' Q
  auto apply = [&](const double* Q, double* Z, int) {
    const Sample& sa = *S[0].s;
    cca_colsum_kernel<B><<<1, 256, 0, st>>>(Q, n1, csU.p);
    cca_yt_u_kernel<B><<<(G + 7) / 8, 256, 0, st>>>(sa.csc_p.p, sa.csc_i.p, sa.csc_x.p, S[0].cols.p, S[0].f.p,
                                                    S[0].mean.p, S[0].posmap.p, Q, csU.p, G, W1.p);
    cca_kw_kernel<B><<<(G + 7) / 8, 256, 0, st>>>(S[1].K.p, W1.p, G, W2.p);
    cca_shift_kernel<B><<<1, 256, 0, st>>>(S[0].mean.p, W2.p, G, shift.p);
    cca_y_w_kernel<B><<<(n1 + 7) / 8, 256, 0, st>>>(sa.csr_p.p, sa.csr_i.p, sa.csr_x.p, S[0].gidx.p, S[0].fcol.p,
                                                    S[0].rows.p, n1, W2.p, shift.p, Z);
    ctx.stats.kernels += 5;
  };
  DevBuf<double> Ucm((size_t)n1 * d), theta((size_t)d);
  {
    StageScope sc(ctx, "reduction.eig");
    top_eigenvectors_operator(ctx, n1, d, ctx.exact_reduction ? 1e-10 : 1e-5, 600, apply, Ucm.p, theta.p);
  }
  // sigma = sqrt(eigenvalue of A A'); v = Yb (Ya' u) / sigma; both scaled by sqrt(sigma) / max sqrt(sigma)
  std::vector<double> th((size_t)d), su((size_t)d), sv((size_t)d);
  theta.download(th.data(), (size_t)d, st);
  CG_CUDA(cudaStreamSynchronize(st));
  double cwmax = 0.0;
  std::vector<double> sigma((size_t)d), cw((size_t)d);
  for (int j = 0; j < d; ++j) {
    sigma[j] = std::sqrt(std::max(th[j], 0.0));
    cw[j] = std::sqrt(sigma[j]);
    cwmax = std::max(cwmax, cw[j]);
  }
  for (int j = 0; j < d; ++j) {
    cw[j] /= cwmax;
    su[j] = cw[j];
    sv[j] = sigma[j] > 0.0 ? cw[j] / sigma[j] : 0.0;
  }
  DevBuf<double> Urm((size_t)n1 * B), Vrm((size_t)n2 * B), dsu, dsv;
  dsu.upload(su.data(), (size_t)d, st);
  dsv.upload(sv.data(), (size_t)d, st);
  cca_cm_to_rm_kernel<<<nblk((size_t)n1 * B), 256, 0, st>>>(Ucm.p, n1, d, B, Urm.p);
  {
    const Sample& sa = *S[0].s;
    const Sample& sb = *S[1].s;
    cca_colsum_kernel<B><<<1, 256, 0, st>>>(Urm.p, n1, csU.p);
    cca_yt_u_kernel<B><<<(G + 7) / 8, 256, 0, st>>>(sa.csc_p.p, sa.csc_i.p, sa.csc_x.p, S[0].cols.p, S[0].f.p,
                                                    S[0].mean.p, S[0].posmap.p, Urm.p, csU.p, G, W1.p);
    cca_shift_kernel<B><<<1, 256, 0, st>>>(S[1].mean.p, W1.p, G, shift.p);
    cca_y_w_kernel<B><<<(n2 + 7) / 8, 256, 0, st>>>(sb.csr_p.p, sb.csr_i.p, sb.csr_x.p, S[1].gidx.p, S[1].fcol.p,
                                                    S[1].rows.p, n2, W1.p, shift.p, Vrm.p);
    ctx.stats.kernels += 5;
  }
  // gene loadings and variance fractions (R/conos.R:349-357): ul = Ya'u, vl = Yb'v with unit columns (W1 already is
  // Ya'u; the 1 / sigma of v drops out of the normalisation); nv_s[j] = var(Y_s l_j) / sum of column variances
  // = l_j' K_s l_j / trace(K_s) with K_s = Y_s'Y_s of the centred kept cells
  {
    const Sample& sb = *S[1].s;
    DevBuf<double> T1((size_t)G * B), T2((size_t)G * B);
    cca_colsum_kernel<B><<<1, 256, 0, st>>>(Vrm.p, n2, csU.p);
    cca_yt_u_kernel<B><<<(G + 7) / 8, 256, 0, st>>>(sb.csc_p.p, sb.csc_i.p, sb.csc_x.p, S[1].cols.p, S[1].f.p,
                                                    S[1].mean.p, S[1].posmap.p, Vrm.p, csU.p, G, W2.p);
    cca_kw_kernel<B><<<(G + 7) / 8, 256, 0, st>>>(S[0].K.p, W1.p, G, T1.p);
    cca_kw_kernel<B><<<(G + 7) / 8, 256, 0, st>>>(S[1].K.p, W2.p, G, T2.p);
    ctx.stats.kernels += 4;
    std::vector<double> hw[2], ht[2], hd[2];
    const DevBuf<double>* Wd[2] = {&W1, &W2};
    const DevBuf<double>* Td[2] = {&T1, &T2};
    for (int side = 0; side < 2; ++side) {
      hw[side].resize((size_t)G * B);
      ht[side].resize((size_t)G * B);
      hd[side].resize((size_t)G);
      Wd[side]->download(hw[side].data(), hw[side].size(), st);
      Td[side]->download(ht[side].data(), ht[side].size(), st);
      CG_CUDA(cudaMemcpy2DAsync(hd[side].data(), sizeof(double), S[side].K.p, ((size_t)G + 1) * sizeof(double),
                                sizeof(double), (size_t)G, cudaMemcpyDeviceToHost, st));
    }
    CG_CUDA(cudaStreamSynchronize(st));
    R.h_nv.assign((size_t)2 * d, 0.0);
    for (int side = 0; side < 2; ++side) {
      std::vector<double>& L = side == 0 ? R.h_ul : R.h_vl;
      L.assign((size_t)G * d, 0.0);
      double tr = 0.0;
      for (int g = 0; g < G; ++g) tr += hd[side][g];
      for (int j = 0; j < d; ++j) {
        double nn = 0.0, q = 0.0;
        for (int g = 0; g < G; ++g) {
          nn += hw[side][(size_t)g * B + j] * hw[side][(size_t)g * B + j];
          q += hw[side][(size_t)g * B + j] * ht[side][(size_t)g * B + j];
        }
        const double inv = nn > 0.0 ? 1.0 / std::sqrt(nn) : 0.0;
        for (int g = 0; g < G; ++g) L[(size_t)j * G + g] = hw[side][(size_t)g * B + j] * inv;
        R.h_nv[(size_t)side * d + j] = (nn > 0.0 && tr > 0.0) ? q / nn / tr : 0.0;
      }
    }
  }
  R.n_u = n1;
  R.n_v = n2;
  R.d = d;
  R.n_rows = G;
  R.n_cols = d;
  R.u.alloc((size_t)n1 * d);
  R.v.alloc((size_t)n2 * d);
  cca_rm_to_cm_scaled_kernel<<<nblk((size_t)n1 * d), 256, 0, st>>>(Urm.p, n1, d, B, dsu.p, R.u.p);
  cca_rm_to_cm_scaled_kernel<<<nblk((size_t)n2 * d), 256, 0, st>>>(Vrm.p, n2, d, B, dsv.p, R.v.p);
  ctx.stats.kernels += 2;
  R.h_rows_u.resize((size_t)n1);
  R.h_rows_v.resize((size_t)n2);
  S[0].rows.download(R.h_rows_u.data(), (size_t)n1, st);
  S[1].rows.download(R.h_rows_v.data(), (size_t)n2, st);
  R.h_sigma = sigma;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(st));
}

}  // namespace

void ensure_gram(Context& ctx, Sample& s);

void compute_cca_reduction(Context& ctx, cg_panel& panel, const cg_pair& pr, const cg_params& P, PairReduction& R) {
  CG_CHECK(pr.n_genes >= 5, "fewer than 5 common od genes: the reference drops such pairs (R/conos.R:330-332)");
  ensure_gram(ctx, panel.samples[pr.a]);
  ensure_gram(ctx, panel.samples[pr.b]);
  const int d = std::min(P.ncomps, pr.n_genes - 1);
  CG_CHECK(d >= 1 && d <= 56, "device CCA supports up to 56 components");
  if (d <= 24) cca_pair<32>(ctx, panel, pr, P, R);
  else cca_pair<64>(ctx, panel, pr, P, R);
}

}  // namespace cg
