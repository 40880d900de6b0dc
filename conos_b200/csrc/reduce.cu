#include <algorithm>
#include <cmath>
#include <map>

#include "eig.cuh"
#include "gemm_tc.cuh"
#include <thread>

#include "reduce.cuh"
#include "stage.cuh"

namespace cg {
namespace {

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// scaled covariance of one sample on the pair's od genes (spcov of the column-scaled matrix, without ever
// forming it):  C[i][j] = f_i f_j (M[c_i][c_j] - n m_i m_j) / (n - 1),  m = colSums / n of the unscaled counts,
// M = X'X of the whole sample (cached per sample).  Zero padded to Gp x Gp.
__global__ void scaled_cov_kernel(const double* __restrict__ M, int Gs, const double* __restrict__ colsum,
                                  const int* __restrict__ cols, const double* __restrict__ f, int G, int Gp, double n,
                                  double* __restrict__ C) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)Gp * Gp) return;
  const int i = (int)(t % Gp), j = (int)(t / Gp);
  double v = 0.0;
  if (i < G && j < G) {
    const int ci = cols[i], cj = cols[j];
    const double mi = colsum[ci] / n, mj = colsum[cj] / n;
    v = f[i] * f[j] * (M[(size_t)cj * Gs + ci] - n * mi * mj) / (n - 1.0);
  }
  C[t] = v;
}

__global__ void mirror_upper_kernel(double* __restrict__ M, int G) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * G) return;
  const int r = (int)(t % G), c = (int)(t / G);
  if (r > c) M[t] = M[(size_t)r * G + c];
}

// rot[:, 2j] = Va[:, j], rot[:, 2j+1] = Vb[:, j]   (interleave <- order(c((1:n1)-0.5, 1:n2)), R/conos.R:300-304)
__global__ void interleave_kernel(const double* __restrict__ Va, const double* __restrict__ Vb, int G, int Gp, int r,
                                  double* __restrict__ rot) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * 2 * r) return;
  const int g = (int)(t % G), c = (int)(t / G);
  const double* src = (c & 1) ? Vb : Va;
  rot[t] = src[(size_t)(c >> 1) * Gp + g];
}

// same with a row gather: pair gene g lives in column cols_a[g] / cols_b[g] of its sample's gene universe
__global__ void interleave_gather_kernel(const double* __restrict__ Va, const double* __restrict__ Vb,
                                         const int* __restrict__ cols_a, const int* __restrict__ cols_b, int G,
                                         int Gsa, int Gsb, int r, double* __restrict__ rot) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * 2 * r) return;
  const int g = (int)(t % G), c = (int)(t / G);
  rot[t] = (c & 1) ? Vb[(size_t)(c >> 1) * Gsb + cols_b[g]] : Va[(size_t)(c >> 1) * Gsa + cols_a[g]];
}

__global__ void trace_kernel(const double* __restrict__ C, int Gp, double* __restrict__ out) {
  // single block
  double s = 0.0;
  for (int i = threadIdx.x; i < Gp; i += blockDim.x) s += C[(size_t)i * Gp + i];
  __shared__ double red[32];
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    *out = tot;
  }
}

__global__ void weighted_sum_kernel(const double* __restrict__ A, const double* __restrict__ B, double wa, double wb,
                                    size_t n, double* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = wa * A[t] + wb * B[t];
}

int r_round_half_even(double x) { return (int)std::nearbyint(x); }  // default rounding mode = to nearest even

// S = wa Da Mc_a Da + wb Db Mc_b Db on the pair's od genes (wa = n_a / n / (n_a - 1)): the cell-weighted mean of the two
// covariance slices, R/conos.R:180-183.  Zero padded to Gp x Gp.
__global__ void mean_cov_kernel(const double* __restrict__ Ma, int Gsa, const double* __restrict__ Mb, int Gsb,
                                const int* __restrict__ cols_a, const int* __restrict__ cols_b,
                                const double* __restrict__ fa, const double* __restrict__ fb, int G, int Gp, double wa,
                                double wb, double* __restrict__ S) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)Gp * Gp) return;
  const int i = (int)(t % Gp), j = (int)(t / Gp);
  double v = 0.0;
  if (i < G && j < G)
    v = wa * fa[i] * fa[j] * Ma[(size_t)cols_a[j] * Gsa + cols_a[i]] +
        wb * fb[i] * fb[j] * Mb[(size_t)cols_b[j] * Gsb + cols_b[i]];
  S[t] = v;
}

// scaledMatricesP2 (R/conos.R:18-46): cqv = exp(rowMeans(log qv)), cgsf = sqrt(cqv / exp(v)), non-finite -> 0
void pair_factors(const Sample& sa, const Sample& sb, const cg_pair& pr, bool var_scale, std::vector<double>& fa,
                  std::vector<double>& fb) {
  const int G = pr.n_genes;
  fa.assign(G, 1.0);
  fb.assign(G, 1.0);
  if (!var_scale) return;
  CG_CHECK(!sa.h_qv.empty() && !sb.h_qv.empty() && !sa.h_v.empty() && !sb.h_v.empty(),
           "var.scale = TRUE needs misc$varinfo$v and $qv for every sample");
  for (int g = 0; g < G; ++g) {
    const int ca = pr.genes_a[g], cb = pr.genes_b[g];
    const double cqv = std::exp((sa.h_logqv[ca] + sb.h_logqv[cb]) / 2.0);
    const double a = std::sqrt(cqv / sa.h_expv[ca]);
    const double b = std::sqrt(cqv / sb.h_expv[cb]);
    fa[g] = std::isfinite(a) ? a : 0.0;
    fb[g] = std::isfinite(b) ? b : 0.0;
  }
}

}  // namespace

// X'X of the whole sample in FP64 (upper triangle accumulated, then mirrored); cached on the sample.
void ensure_gram(Context& ctx, Sample& s);
void ensure_grams(Context& ctx, Sample* const* samples, int n);

// host copy of a device-computed rotation: asynchronous, into the context's page-locked arena
static void download_rotation(Context& ctx, PairReduction& R) {
  const size_t n = (size_t)R.n_rows * R.n_cols;
  double* h = static_cast<double*>(ctx.rot_arena.take(n * sizeof(double)));
  CG_CUDA(cudaMemcpyAsync(h, R.rot.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  R.h_rot_pinned = h;
}

void compute_pair_reductions(Context& ctx, cg_panel& panel, const cg_pair* pairs, int n_pairs, const cg_params& P,
                             std::vector<PairReduction>& out) {
  cudaStream_t st = ctx.stream;
  // ---- supplied reductions (the cache-hit path of updatePairs, R/conclass.R:1062-1071)
  std::vector<int> todo;
  for (int q = 0; q < n_pairs; ++q) {
    const cg_pair& pr = pairs[q];
    PairReduction& R = out[q];
    if (P.space == CG_SPACE_CCA) {
      if (pr.u && pr.v) {
        CG_CHECK(pr.n_u > 0 && pr.n_v > 0 && pr.rows_u && pr.rows_v, "CCA needs the row maps of u and v");
        CG_CHECK(pr.uv_cols >= 1, "cached CCA variates need their column count (uv_cols)");
        R.n_u = pr.n_u;
        R.n_v = pr.n_v;
        R.d = pr.uv_cols;  // every cached column, whatever ncomps says now (R/conclass.R:268-270)
        R.u.upload(pr.u, (size_t)pr.n_u * pr.uv_cols, st);
        R.v.upload(pr.v, (size_t)pr.n_v * pr.uv_cols, st);
        R.h_rows_u.assign(pr.rows_u, pr.rows_u + pr.n_u);
        R.h_rows_v.assign(pr.rows_v, pr.rows_v + pr.n_v);
      } else {
        todo.push_back(q);
      }
    } else {
      CG_CHECK(pr.n_genes >= 5, "fewer than 5 common od genes: the reference drops such pairs (R/conos.R:206,278)");
      if (pr.rot && pr.rot_cols > 0) {
        CG_CHECK(pr.rot_rows == pr.n_genes,
                 "cached rotation: rot_rows must equal n_genes (the pair's genes are the rownames of the cached CPC)");
        R.n_rows = pr.n_genes;
        R.n_cols = pr.rot_cols;
        R.d = P.ncomps < pr.rot_cols ? P.ncomps : pr.rot_cols;
        R.rot.upload(pr.rot, (size_t)pr.n_genes * pr.rot_cols, st);
        R.h_rot.assign(pr.rot, pr.rot + (size_t)pr.n_genes * pr.rot_cols);
      } else {
        todo.push_back(q);
      }
    }
  }
  CG_CUDA(cudaStreamSynchronize(st));
  if (todo.empty()) return;
  if (P.space == CG_SPACE_CCA) {
    StageScope sc(ctx, "reduction.cca");
    for (int q : todo) compute_cca_reduction(ctx, panel, pairs[q], P, out[q]);
    return;
  }

  // ---- per-sample Gram matrices
  {
    StageScope sc(ctx, "reduction.gram");
    std::vector<Sample*> need;
    for (int q : todo) {
      need.push_back(&panel.samples[pairs[q].a]);
      need.push_back(&panel.samples[pairs[q].b]);
    }
    ensure_grams(ctx, need.data(), (int)need.size());
  }

  if (P.space == CG_SPACE_PCA) {
    // quickPlainPCA (R/conos.R:273-315): r = round(ncomps / 2) leading right singular vectors of each
    // centred, scaled sample == leading eigenvectors of its scaled covariance; columns interleaved.
    const int r = r_round_half_even(P.ncomps / 2.0);
    CG_CHECK(r >= 1 && r <= 56, "ncomps out of range for the device PCA");
    // tensor-core path: the operator of every pair acts on the per-sample centred Gram matrix over the sample's
    // whole gene universe; the pair's gene set (any subset, any order) enters through the diagonal scaling (zero
    // outside the set) and a row gather of the eigenvectors.  The Gram matrix is dense: bounded universe only.
    // The choice is made pair by pair (never for a batch as a whole: the result of a pair must not depend on which
    // other pairs share its call); pairs that do not qualify take the FP64 driver below.
    std::vector<int> fp64_pairs;
    if (!ctx.exact_reduction && r <= 24) {
      std::vector<int> tc_pairs;
      for (int q : todo) {
        const cg_pair& pr = pairs[q];
        const Sample& sa = panel.samples[pr.a];
        const Sample& sb = panel.samples[pr.b];
        // 8192 genes: 537 MB of FP64 Gram matrix per sample (the union of 100 real top-2000 od lists is 5-8 thousand)
        const bool ok = pr.n_genes >= 32 && sa.n_genes >= 32 && sb.n_genes >= 32 && sa.n_genes <= 8192 &&
                        sb.n_genes <= 8192;
        (ok ? tc_pairs : fp64_pairs).push_back(q);
      }
      todo = tc_pairs;
    } else {
      fp64_pairs = todo;
      todo.clear();
    }
    if (!todo.empty()) {
      const int S = (int)panel.samples.size();
      std::vector<TcEigSample> ts(S);
      for (int si = 0; si < S; ++si) {
        Sample& sm = panel.samples[si];
        ts[si] = TcEigSample{sm.gram.p, sm.colsum.p, sm.n_cells, sm.n_genes};
      }
      struct PB {
        const double* f[2];   // into the staging buffer: scale factor per gene of the sample's universe
        const int* cols[2];   // into the staging buffer: column of every pair gene
        double* V[2];         // into the batch buffers below
        double* theta[2];
        std::vector<double> hf[2];  // per pair gene
      };
      std::vector<PB> pbs(todo.size());
      std::vector<TcEigProblem> tp;
      std::vector<int> tc_conv(2 * todo.size(), 1), redo;
      // one staging upload and one allocation per kind for the whole batch
      size_t n_f = 0, n_c = 0, n_v = 0;
      for (size_t t = 0; t < todo.size(); ++t) {
        const cg_pair& pr = pairs[todo[t]];
        n_f += (size_t)panel.samples[pr.a].n_genes + panel.samples[pr.b].n_genes;
        n_c += 2 * (size_t)pr.n_genes;
        n_v += ((size_t)panel.samples[pr.a].n_genes + panel.samples[pr.b].n_genes) * r;
      }
      // staged in the context's page-locked arena (valid until the next cg_pairs_edges call): the loop hands out the
      // slices, several host threads fill them (the atlas has 9900 of them per call)
      double* h_f = static_cast<double*>(ctx.rot_arena.take(std::max<size_t>(n_f, 1) * sizeof(double)));
      int* h_c = static_cast<int*>(ctx.rot_arena.take(std::max<size_t>(n_c, 1) * sizeof(int)));
      DevBuf<double> d_f(n_f), d_V(n_v), d_theta(2 * (size_t)r * todo.size());
      DevBuf<int> d_c(n_c);
      size_t o_v = 0, o_f = 0, o_c = 0;
      std::vector<size_t> off_f(2 * todo.size()), off_c(2 * todo.size());
      for (size_t t = 0; t < todo.size(); ++t) {
        const cg_pair& pr = pairs[todo[t]];
        const int G = pr.n_genes;
        Sample* sm[2] = {&panel.samples[pr.a], &panel.samples[pr.b]};
        PairReduction& R = out[todo[t]];
        for (int g = 0; g < G; ++g)
          CG_CHECK(pr.genes_a[g] >= 0 && pr.genes_a[g] < sm[0]->n_genes && pr.genes_b[g] >= 0 &&
                       pr.genes_b[g] < sm[1]->n_genes,
                   "pair gene column out of range");
        if ((int)R.h_fa.size() != G || (int)R.h_fb.size() != G)   // not computed up front by the caller
          pair_factors(*sm[0], *sm[1], pr, P.var_scale != 0, R.h_fa, R.h_fb);
        for (int side = 0; side < 2; ++side) {
          const int Gs = sm[side]->n_genes;
          if (P.score_component_variance) pbs[t].hf[side] = side == 0 ? R.h_fa : R.h_fb;
          off_f[2 * t + side] = o_f;
          off_c[2 * t + side] = o_c;
          pbs[t].f[side] = d_f.p + o_f;
          pbs[t].cols[side] = d_c.p + o_c;
          o_f += (size_t)Gs;
          o_c += (size_t)G;
          pbs[t].V[side] = d_V.p + o_v;
          o_v += (size_t)Gs * r;
          pbs[t].theta[side] = d_theta.p + (2 * t + side) * (size_t)r;
          // identity of the problem: (sample a, sample b, side) -- the same on every rank and in every batch
          const unsigned seed = ((unsigned)pr.a * 40503u + (unsigned)pr.b) * 2u + (unsigned)side;
          tp.push_back(TcEigProblem{side == 0 ? pr.a : pr.b, seed, pbs[t].f[side], pbs[t].V[side], pbs[t].theta[side]});
        }
      }
      {
        // scale factor per gene of the sample's universe (zero outside the pair's gene set) and the gene columns
        auto fill = [&](size_t t0, size_t t1) {
          for (size_t t = t0; t < t1; ++t) {
            const cg_pair& pr = pairs[todo[t]];
            const int G = pr.n_genes;
            const PairReduction& R = out[todo[t]];
            for (int side = 0; side < 2; ++side) {
              const int Gs = panel.samples[side == 0 ? pr.a : pr.b].n_genes;
              const int* cols = side == 0 ? pr.genes_a : pr.genes_b;
              const std::vector<double>& f = side == 0 ? R.h_fa : R.h_fb;
              double* fu = h_f + off_f[2 * t + side];
              std::fill(fu, fu + Gs, 0.0);
              for (int g = 0; g < G; ++g) fu[cols[g]] = f[g];
              memcpy(h_c + off_c[2 * t + side], cols, (size_t)G * sizeof(int));
            }
          }
        };
        const int nt = host_fill_threads(ctx, todo.size(), 64);
        if (nt == 1) {
          fill(0, todo.size());
        } else {
          std::vector<std::thread> th;
          for (int k = 0; k < nt; ++k) th.emplace_back(fill, todo.size() * k / nt, todo.size() * (k + 1) / nt);
          for (auto& x : th) x.join();
        }
      }
      if (o_f) CG_CUDA(cudaMemcpyAsync(d_f.p, h_f, o_f * sizeof(double), cudaMemcpyHostToDevice, st));
      if (o_c) CG_CUDA(cudaMemcpyAsync(d_c.p, h_c, o_c * sizeof(int), cudaMemcpyHostToDevice, st));
      {
        StageScope sc(ctx, "reduction.eig");
        top_eigenvectors_shared_gram(ctx, ts.data(), S, tp.data(), (int)tp.size(), r, ctx.tc_eig_tol, ctx.tc_eig_max_sweeps, ctx.tc_eig_degree,
                                     tc_conv.data());
      }
      for (size_t t = 0; t < todo.size(); ++t) {
        if (!tc_conv[2 * t] || !tc_conv[2 * t + 1]) {  // not at irlba's tolerance: this pair goes to the FP64 driver
          redo.push_back(todo[t]);
          continue;
        }
        const cg_pair& pr = pairs[todo[t]];
        const int G = pr.n_genes;
        PairReduction& R = out[todo[t]];
        R.n_rows = G;
        R.n_cols = 2 * r;
        R.d = P.ncomps < R.n_cols ? P.ncomps : R.n_cols;
        R.rot.alloc((size_t)G * R.n_cols);
        interleave_gather_kernel<<<nblk((size_t)G * R.n_cols), 256, 0, st>>>(
            pbs[t].V[0], pbs[t].V[1], pbs[t].cols[0], pbs[t].cols[1], G, panel.samples[pr.a].n_genes,
            panel.samples[pr.b].n_genes, r, R.rot.p);
        ctx.stats.kernels++;
        download_rotation(ctx, R);
        if (P.score_component_variance) {
          // trace(C) = sum_g f_g^2 (M_cc - n m_c^2) / (n - 1), c = the gene's column: from the host copies
          std::vector<double> th(2 * (size_t)r);
          CG_CUDA(cudaMemcpyAsync(th.data(), pbs[t].theta[0], 2 * (size_t)r * sizeof(double), cudaMemcpyDeviceToHost, st));
          CG_CUDA(cudaStreamSynchronize(st));
          R.h_nv.assign(2 * (size_t)R.n_cols, 0.0);
          for (int side = 0; side < 2; ++side) {
            Sample& sm = panel.samples[side == 0 ? pr.a : pr.b];
            if (sm.h_gram_diag.empty()) {
              sm.h_gram_diag.resize(sm.n_genes);
              CG_CUDA(cudaMemcpy2DAsync(sm.h_gram_diag.data(), sizeof(double), sm.gram.p,
                                        ((size_t)sm.n_genes + 1) * sizeof(double), sizeof(double), sm.n_genes,
                                        cudaMemcpyDeviceToHost, st));
              CG_CUDA(cudaStreamSynchronize(st));
            }
            const std::vector<double>& fh = pbs[t].hf[side];
            const int* cols = side == 0 ? pr.genes_a : pr.genes_b;
            const double n = sm.n_cells;
            double tr = 0.0;
            for (int g = 0; g < G; ++g) {
              const int c = cols[g];
              const double m = sm.h_colsum[c] / n;
              tr += fh[g] * fh[g] * (sm.h_gram_diag[c] - n * m * m) / (n - 1.0);
            }
            for (int j = 0; j < r; ++j) R.h_nv[(size_t)side * R.n_cols + j] = th[(size_t)side * r + j] / tr;
          }
        }
      }
      CG_CUDA(cudaStreamSynchronize(st));
      ctx.eig_redo_pairs += redo.size();
      if (ctx.profile) ctx.stage_ms["reduction.eig.fp64_redo_pairs"] += (double)redo.size();
      fp64_pairs.insert(fp64_pairs.end(), redo.begin(), redo.end());
      std::sort(fp64_pairs.begin(), fp64_pairs.end());
    }
    todo = fp64_pairs;
    if (todo.empty()) return;
    const int B = r <= 24 ? 32 : 64;
    // chunks of pairs bounded by memory (2 covariance matrices per pair)
    size_t i0 = 0;
    while (i0 < todo.size()) {
      size_t need_fp64 = 0;
      for (size_t t = i0; t < todo.size(); ++t) {
        const size_t Gp = (size_t)std::max(pairs[todo[t]].n_genes, B);
        need_fp64 += 2 * (Gp * Gp * 8 + Gp * (2 * B + r) * 8);
      }
      const size_t freeb = ctx.free_bytes_for(need_fp64);
      size_t per = 0;
      size_t i1 = i0;
      size_t used = 0;
      while (i1 < todo.size()) {
        const int G = pairs[todo[i1]].n_genes;
        const int Gp = G > B ? G : B;
        per = 2 * ((size_t)Gp * Gp * 8 + (size_t)Gp * (2 * B + r) * 8);
        if (i1 > i0 && used + per > freeb / 3) break;
        used += per;
        ++i1;
      }
      const int nq = (int)(i1 - i0);
      struct Buf {
        DevBuf<double> C[2], V[2], theta[2], f[2], tr;
        DevBuf<int> cols[2];
        int G, Gp;
      };
      std::vector<Buf> bufs(nq);
      std::vector<EigProblem> eprobs;
      for (int t = 0; t < nq; ++t) {
        const cg_pair& pr = pairs[todo[i0 + t]];
        Buf& b = bufs[t];
        b.G = pr.n_genes;
        b.Gp = b.G > B ? b.G : B;
        Sample* sm[2] = {&panel.samples[pr.a], &panel.samples[pr.b]};
        const int* gcols[2] = {pr.genes_a, pr.genes_b};
        // scale factors (same arithmetic as api_pairs.cu: scaledMatricesP2)
        std::vector<double> f[2];
        pair_factors(*sm[0], *sm[1], pr, P.var_scale != 0, f[0], f[1]);
        b.tr.alloc(2);
        for (int side = 0; side < 2; ++side) {
          b.C[side].alloc((size_t)b.Gp * b.Gp);
          b.V[side].alloc((size_t)b.Gp * r);
          b.theta[side].alloc((size_t)r);
          b.f[side].upload(f[side].data(), (size_t)b.G, st);
          b.cols[side].upload(gcols[side], (size_t)b.G, st);
          scaled_cov_kernel<<<nblk((size_t)b.Gp * b.Gp), 256, 0, st>>>(sm[side]->gram.p, sm[side]->n_genes,
                                                                        sm[side]->colsum.p, b.cols[side].p,
                                                                        b.f[side].p, b.G, b.Gp,
                                                                        (double)sm[side]->n_cells, b.C[side].p);
          ctx.stats.kernels++;
          if (P.score_component_variance) {
            trace_kernel<<<1, 256, 0, st>>>(b.C[side].p, b.Gp, b.tr.p + side);
            ctx.stats.kernels++;
          }
          eprobs.push_back(EigProblem{b.C[side].p, b.Gp, b.V[side].p, b.theta[side].p});
        }
      }
      CG_CUDA(cudaGetLastError());
      {
        StageScope sc(ctx, "reduction.eig");
        top_eigenvectors_batched(ctx, eprobs.data(), (int)eprobs.size(), r, ctx.exact_reduction ? 1e-10 : 1e-5, 400);
      }
      for (int t = 0; t < nq; ++t) {
        Buf& b = bufs[t];
        PairReduction& R = out[todo[i0 + t]];
        R.n_rows = b.G;
        R.n_cols = 2 * r;
        R.d = P.ncomps < R.n_cols ? P.ncomps : R.n_cols;
        R.rot.alloc((size_t)b.G * R.n_cols);
        interleave_kernel<<<nblk((size_t)b.G * R.n_cols), 256, 0, st>>>(b.V[0].p, b.V[1].p, b.G, b.Gp, r, R.rot.p);
        ctx.stats.kernels++;
        download_rotation(ctx, R);
        if (P.score_component_variance) {
          // nv_j = var(projection on v_j) / sum of column variances = theta_j / trace(C)   (R/conos.R:289-296)
          std::vector<double> th(2 * (size_t)r), tr(2);
          b.theta[0].download(th.data(), (size_t)r, st);
          b.theta[1].download(th.data() + r, (size_t)r, st);
          b.tr.download(tr.data(), 2, st);
          CG_CUDA(cudaStreamSynchronize(st));
          R.h_nv.assign(2 * (size_t)R.n_cols, 0.0);
          for (int side = 0; side < 2; ++side)
            for (int j = 0; j < r; ++j) R.h_nv[(size_t)side * R.n_cols + j] = th[(size_t)side * r + j] / tr[side];
        }
      }
      CG_CUDA(cudaStreamSynchronize(st));
      i0 = i1;
    }
    return;
  }
  // ---- CPCA: quickCPCA (R/conos.R:204-259) -> cpcaFast (:175-191) -> cpcaF (src/cpca.cpp)
  CG_CHECK(P.space == CG_SPACE_CPCA, "unknown space");
  // The covariance slices of a pair are never built: C_s^(pair) = D Mc_s D / (n_s - 1) with the sample's centred Gram
  // matrix Mc_s shared by all of its pairs.  Initial vectors: leading eigenvectors of the cell-weighted mean covariance
  // S (explicit G x G, FP64 block iteration, chunks of pairs); then every pair of the call iterates in lock step
  // (cpca_batch.cu).
  const int S = (int)panel.samples.size();
  std::vector<DevBuf<double>> Mc(S);
  std::vector<const double*> Mc_ptr(S, nullptr);
  for (int q : todo)
    for (int si : {pairs[q].a, pairs[q].b})
      if (!Mc[si].n) {
        Mc[si].alloc((size_t)panel.samples[si].n_genes * panel.samples[si].n_genes);
        center_gram(ctx, panel.samples[si], Mc[si].p);
        Mc_ptr[si] = Mc[si].p;
      }
  struct CpcaBuf {
    DevBuf<double> f[2], V0, CPC, D, niter, tr;
    DevBuf<int> cols[2];
    int G = 0, Gp = 0, ncomp = 0;
  };
  std::vector<CpcaBuf> bufs(todo.size());
  std::vector<CpcaBatchPair> bp(todo.size());
  for (size_t t = 0; t < todo.size(); ++t) {
    const cg_pair& pr = pairs[todo[t]];
    CpcaBuf& b = bufs[t];
    const int G = pr.n_genes;
    const int ncomp = std::min(P.ncomps, G - 1);  // R/conos.R:210
    CG_CHECK(ncomp >= 1 && ncomp <= 56, "device CPCA supports up to 56 common components");
    const int B = ncomp <= 24 ? 32 : 64;
    b.G = G;
    b.Gp = G > B ? G : B;
    b.ncomp = ncomp;
    Sample* sm[2] = {&panel.samples[pr.a], &panel.samples[pr.b]};
    std::vector<double> f[2];
    pair_factors(*sm[0], *sm[1], pr, P.var_scale != 0, f[0], f[1]);
    const int* gcols[2] = {pr.genes_a, pr.genes_b};
    for (int side = 0; side < 2; ++side) {
      b.f[side].upload(f[side].data(), (size_t)G, st);
      b.cols[side].upload(gcols[side], (size_t)G, st);
    }
    b.V0.alloc((size_t)b.Gp * ncomp);
    b.CPC.alloc((size_t)G * ncomp);
    b.D.alloc((size_t)2 * ncomp);
    b.niter.alloc((size_t)ncomp);
    b.tr.alloc(2);
    bp[t] = CpcaBatchPair{pr.a, pr.b, G, ncomp, b.Gp, {b.cols[0].p, b.cols[1].p}, {b.f[0].p, b.f[1].p}, b.V0.p,
                          b.CPC.p, b.D.p, b.niter.p, b.tr.p};
  }
  CG_CUDA(cudaStreamSynchronize(st));  // the host vectors of the loop above are gone
  // initial vectors.  Default mode: S is applied as an operator for all pairs at once (one tensor-core GEMM per sample
  // and application, eig.cu); pairs that miss irlba's tolerance there -- and every pair in exact mode -- get the
  // explicit G x G matrix and the FP64 block iteration, chunk by chunk (a batched solve shares the block width).
  std::vector<size_t> expl;
  if (!ctx.exact_reduction) {
    StageScope sc(ctx, "reduction.cpca_init_tc");
    std::vector<PairOpSample> ps(S);
    for (int si = 0; si < S; ++si) ps[si] = PairOpSample{Mc_ptr[si], panel.samples[si].n_genes};
    std::map<int, std::vector<size_t>> by_ncomp;
    for (size_t t = 0; t < todo.size(); ++t) by_ncomp[bufs[t].ncomp].push_back(t);
    for (auto& kv : by_ncomp) {
      std::vector<PairOpProblem> pp;
      std::vector<DevBuf<double>> th(kv.second.size());
      for (size_t u = 0; u < kv.second.size(); ++u) {
        const size_t t = kv.second[u];
        const cg_pair& pr = pairs[todo[t]];
        CpcaBuf& b = bufs[t];
        const double na = panel.samples[pr.a].n_cells, nb = panel.samples[pr.b].n_cells;
        th[u].alloc((size_t)b.ncomp);
        PairOpProblem q{};
        q.a = pr.a;
        q.b = pr.b;
        q.G = b.G;
        q.rows = b.Gp;
        q.cols[0] = b.cols[0].p;
        q.cols[1] = b.cols[1].p;
        q.f[0] = b.f[0].p;
        q.f[1] = b.f[1].p;
        q.w[0] = na / (na + nb) / (na - 1.0);
        q.w[1] = nb / (na + nb) / (nb - 1.0);
        q.V = b.V0.p;
        q.theta = th[u].p;
        pp.push_back(q);
      }
      std::vector<int> conv(pp.size(), 1);
      top_eigenvectors_pair_sum_tc(ctx, ps.data(), S, pp.data(), (int)pp.size(), kv.first, 1e-5, 400, conv.data());
      for (size_t u = 0; u < kv.second.size(); ++u)
        if (!conv[u]) expl.push_back(kv.second[u]);
    }
    std::sort(expl.begin(), expl.end());
    ctx.eig_redo_pairs += expl.size();
    if (ctx.profile) ctx.stage_ms["reduction.eig.fp64_redo_pairs"] += (double)expl.size();
  } else {
    for (size_t t = 0; t < todo.size(); ++t) expl.push_back(t);
  }
  if (!expl.empty()) {
    StageScope sc(ctx, "reduction.cpca_init");
    size_t c0 = 0;
    while (c0 < expl.size()) {
      size_t need_init = 0;
      for (size_t t = c0; t < expl.size(); ++t)
        need_init += (size_t)bufs[expl[t]].Gp * bufs[expl[t]].Gp * 8 + (size_t)bufs[expl[t]].Gp * 64 * 8 * 4;
      const size_t freeb = ctx.free_bytes_for(need_init);
      size_t c1 = c0, used = 0;
      while (c1 < expl.size()) {
        if (bufs[expl[c1]].ncomp != bufs[expl[c0]].ncomp) break;
        const size_t need = (size_t)bufs[expl[c1]].Gp * bufs[expl[c1]].Gp * 8 + (size_t)bufs[expl[c1]].Gp * 64 * 8 * 4;
        if (c1 > c0 && used + need > std::min(freeb / 3, ctx.max_batch_bytes)) break;
        used += need;
        ++c1;
      }
      std::vector<DevBuf<double>> Sm(c1 - c0), th(c1 - c0);
      std::vector<EigProblem> eprobs;
      for (size_t tt = c0; tt < c1; ++tt) {
        const size_t t = expl[tt];
        const cg_pair& pr = pairs[todo[t]];
        CpcaBuf& b = bufs[t];
        const Sample& sa = panel.samples[pr.a];
        const Sample& sb = panel.samples[pr.b];
        Sm[tt - c0].alloc((size_t)b.Gp * b.Gp);
        th[tt - c0].alloc((size_t)b.ncomp);
        const double na = sa.n_cells, nb = sb.n_cells;
        mean_cov_kernel<<<nblk((size_t)b.Gp * b.Gp), 256, 0, st>>>(
            Mc_ptr[pr.a], sa.n_genes, Mc_ptr[pr.b], sb.n_genes, b.cols[0].p, b.cols[1].p, b.f[0].p, b.f[1].p, b.G, b.Gp,
            na / (na + nb) / (na - 1.0), nb / (na + nb) / (nb - 1.0), Sm[tt - c0].p);
        ctx.stats.kernels++;
        eprobs.push_back(EigProblem{Sm[tt - c0].p, b.Gp, b.V0.p, th[tt - c0].p});
      }
      CG_CUDA(cudaGetLastError());
      {
        StageScope sc2(ctx, "reduction.eig");
        top_eigenvectors_batched(ctx, eprobs.data(), (int)eprobs.size(), bufs[expl[c0]].ncomp,
                                 ctx.exact_reduction ? 1e-10 : 1e-5, 600);
      }
      c0 = c1;
    }
  }
  {
    StageScope sc(ctx, "reduction.cpca");
    cpca_batched(ctx, panel, bp.data(), (int)bp.size(), Mc_ptr.data(), P.cpca_maxit, P.cpca_tol);
  }
  for (size_t t = 0; t < todo.size(); ++t) {
    CpcaBuf& b = bufs[t];
    PairReduction& R = out[todo[t]];
    R.n_rows = b.G;
    R.n_cols = b.ncomp;
    R.d = P.ncomps < b.ncomp ? P.ncomps : b.ncomp;
    R.rot = std::move(b.CPC);
    download_rotation(ctx, R);
    R.h_D.resize((size_t)2 * b.ncomp);
    R.h_niter.resize((size_t)b.ncomp);
    b.D.download(R.h_D.data(), R.h_D.size(), st);
    b.niter.download(R.h_niter.data(), R.h_niter.size(), st);
    if (P.score_component_variance) {
      std::vector<double> htr(2);
      b.tr.download(htr.data(), 2, st);
      CG_CUDA(cudaStreamSynchronize(st));
      R.h_nv.assign((size_t)2 * b.ncomp, 0.0);
      for (int side = 0; side < 2; ++side)
        for (int j = 0; j < b.ncomp; ++j)
          R.h_nv[(size_t)side * b.ncomp + j] = R.h_D[(size_t)side * b.ncomp + j] / htr[side];
    }
  }
  CG_CUDA(cudaStreamSynchronize(st));
}

void ensure_gram(Context& ctx, Sample& s) {
  Sample* one[1] = {&s};
  ensure_grams(ctx, one, 1);
}

// Gram matrices X'X of several samples: the dense bf16 hi/lo operands of a chunk are built first, then ONE batched
// GEMM computes the tiles that touch the upper triangles (a single 2000 x 2000 product is 128 tiles on 148 SMs --
// one under-filled wave however few of them are needed), and the lower triangles are mirrored (which also makes the
// matrix exactly symmetric: the two cross terms of the bf16x3 product are accumulated in a different order at (m, n)
// and (n, m)).
void ensure_grams(Context& ctx, Sample* const* samples, int n) {
  std::vector<Sample*> todo;
  for (int t = 0; t < n; ++t) {
    Sample* s = samples[t];
    if (s->gram.n) continue;
    bool dup = false;
    for (Sample* o : todo) dup = dup || o == s;
    if (!dup) todo.push_back(s);
  }
  if (todo.empty()) return;
  {
    StageScope sc_a(ctx, "gram.alloc");
    for (Sample* s : todo) {
      CG_CHECK(s->n_genes <= 16384, "subset the counts matrix to the od-gene universe before upload (Gram matrix too big)");
      s->gram.alloc((size_t)s->n_genes * s->n_genes);
    }
  }
  if (ctx.exact_reduction) {
    for (Sample* s : todo) {
      const size_t G = (size_t)s->n_genes;
      s->gram.zero(ctx.stream);
      gram_upper_from_csr(ctx, s->csr_p.p, s->csr_i.p, s->csr_x.p, s->n_cells, s->n_genes, s->gram.p);
      mirror_upper_kernel<<<nblk(G * G), 256, 0, ctx.stream>>>(s->gram.p, s->n_genes);
      ctx.stats.kernels++;
    }
    CG_CUDA(cudaGetLastError());
    return;
  }
  // X'X on the tensor cores: the sparse counts become dense bf16 hi/lo operands (rows = genes, k = cells)
  const size_t budget = (size_t)8 << 30;
  size_t i0 = 0;
  while (i0 < todo.size()) {
    size_t i1 = i0, used = 0;
    while (i1 < todo.size()) {
      const size_t need = 2 * TiledOperand::elems(todo[i1]->n_genes, todo[i1]->n_cells) * sizeof(__nv_bfloat16);
      if (i1 > i0 && used + need > budget) break;
      used += need;
      ++i1;
    }
    std::vector<DevBuf<__nv_bfloat16>> hi(i1 - i0), lo(i1 - i0);
    std::vector<GemmProblem> gps;
    StageScope* sc_o = new StageScope(ctx, "gram.operands");
    for (size_t t = i0; t < i1; ++t) {
      Sample* s = todo[t];
      hi[t - i0].alloc(TiledOperand::elems(s->n_genes, s->n_cells));
      lo[t - i0].alloc(TiledOperand::elems(s->n_genes, s->n_cells));
      TiledOperand op;
      op.hi = hi[t - i0].p;
      op.lo = lo[t - i0].p;
      // fp16 hi + lo (22 significant bits) whenever the values fit: expression values are log-scale, 0 .. ~6; the
      // contraction over the cells runs in 1024-cell FP32 chunks summed in FP64 (see gemm_tc.cu)
      const bool fp16 = s->max_abs < 3.0e4;
      tiled_from_csr(ctx, s->csr_p.p, s->csr_i.p, s->csr_x.p, s->n_cells, s->n_genes, op, fp16);
      GemmProblem gp{op.hi, op.lo, op.hi, op.lo, op.rows_pad, op.rows_pad, op.k_pad, s->n_genes, s->n_genes, s->gram.p,
                     (long long)s->n_genes, 1, 1};
      gp.chunk_slabs = ctx.gram_chunk_slabs;
      gp.fp16 = fp16 ? 1 : 0;
      gps.push_back(gp);
    }
    delete sc_o;
    {
      StageScope sc_g(ctx, "gram.gemm");
      gemm_tn_bf16x3(ctx, gps.data(), (int)gps.size());
    }
    StageScope sc_m(ctx, "gram.mirror_and_release");
    for (size_t t = i0; t < i1; ++t) {
      const size_t G = (size_t)todo[t]->n_genes;
      mirror_upper_kernel<<<nblk(G * G), 256, 0, ctx.stream>>>(todo[t]->gram.p, todo[t]->n_genes);
      ctx.stats.kernels++;
    }
    CG_CUDA(cudaGetLastError());   // the operand panels are released in stream order
    i0 = i1;
  }
}

}  // namespace cg
