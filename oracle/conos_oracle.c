/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the loop-heavy pieces of the conos
 * buildGraph hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library; the product (conos_b200 / libconosgpu.so) never does.
 *
 * PARITY UNPINNED: the reference ships no golden vectors for this path and cannot be run
 * here (no R; N2R/irlba/Matrix/igraph are not vendored), so these functions restate the reference sources
 * cited at each function, plus the published behaviour of the un-vendored N2R/n2 (exact search assumed).
 *
 * Reference files restated (paths relative to the conos source tree):
 *   N2R::crossKnn / N2R::Knn call sites     R/conos.R:595, 851-855   (n2 is NOT vendored: float32 cast,
 *                                           angular = 1 - dot of L2-normalised rows, L2 = squared Euclidean)
 *   pareDownHubEdges                        src/edgeFilter.cpp:23-60
 *   cpcaF                                   src/cpca.cpp:16-101
 *   getSumWeightMatrix / adjustWeights...C  src/edge_rebalancing.cpp:14-48
 *   ReferenceEdges / referenceWij           src/edgeweights.cpp:25-184
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads the OpenMP loops of this library will use (bench.py prints it next to the CPU baseline) */
int oracle_max_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* ------------------------------------------------------------------ kNN (canonical fp32 arithmetic)
 * normalise: ss = fma chain of x_i*x_i (i ascending) ; inv = 1/sqrtf(ss) (0 when ss == 0); xh = x*inv.
 * dot      : four interleaved fma chains a_j = sum_{i = j mod 4} q_i*r_i (the 4-lane SIMD accumulation of
 *            a CPU dot product [3P: n2 uses SSE/AVX kernels]), dot = (a_0 + a_1) + (a_2 + a_3).
 * in : X column-major double n x d (leading dimension ld) -- R's matrix layout
 * out: row-major float [n][d]                                                                     */
void oracle_prepare_f32(const double* X, int n, int d, int ld, int metric, float* out) {
  for (int r = 0; r < n; ++r) {
    float inv = 1.0f;
    if (metric == 0) {
      float ss = 0.0f;
      for (int c = 0; c < d; ++c) {
        float x = (float)X[(size_t)c * ld + r];
        ss = fmaf(x, x, ss);
      }
      inv = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
    }
    for (int c = 0; c < d; ++c) out[(size_t)r * d + c] = (float)X[(size_t)c * ld + r] * inv;
  }
}

static inline void list_insert(float* ld, int* li, int k, float dd, int j) {
  int p = k - 1;
  while (p > 0 && ld[p - 1] > dd) {
    ld[p] = ld[p - 1];
    li[p] = li[p - 1];
    --p;
  }
  ld[p] = dd;
  li[p] = j;
}

/* Exact search: for every row of Q the k rows of R with smallest (dist, index).
 * metric 0 (angular): dist = 1 - fma-chain dot of prepared (normalised) rows; metric 1: squared L2 chain.
 * Q, R row-major float, outputs [nq][k] (idx -1 / dist +inf when nr < k). nthreads <= 0: all cores.   */
#define RB 16
void oracle_knn_f32(const float* Q, int nq, const float* R, int nr, int d, int k, int metric, int* idx,
                    float* dist, int nthreads) {
  /* transposed, padded copy of R so that 16 references advance in lock step (each lane is its own
   * sequential chain over i, so the result is bit-identical to the scalar loop) */
  size_t nrp = ((size_t)nr + RB - 1) / RB * RB;
  float* RT = (float*)calloc((size_t)d * nrp + 1, sizeof(float));
  for (int j = 0; j < nr; ++j)
    for (int c = 0; c < d; ++c) RT[(size_t)c * nrp + j] = R[(size_t)j * d + c];
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int i = 0; i < nq; ++i) {
    float* ld = dist + (size_t)i * k;
    int* li = idx + (size_t)i * k;
    for (int p = 0; p < k; ++p) {
      ld[p] = INFINITY;
      li[p] = -1;
    }
    const float* q = Q + (size_t)i * d;
    for (size_t jb = 0; jb < nrp; jb += RB) {
      float acc[RB];
      float a4[4][RB];
      for (int j4 = 0; j4 < 4; ++j4)
        for (int l = 0; l < RB; ++l) a4[j4][l] = 0.0f;
      if (metric == 0) {
        for (int c = 0; c < d; ++c) {
          const float qc = q[c];
          const float* rt = RT + (size_t)c * nrp + jb;
          float* a = a4[c & 3];
          for (int l = 0; l < RB; ++l) a[l] = __builtin_fmaf(qc, rt[l], a[l]);
        }
        for (int l = 0; l < RB; ++l) acc[l] = 1.0f - ((a4[0][l] + a4[1][l]) + (a4[2][l] + a4[3][l]));
      } else {
        for (int c = 0; c < d; ++c) {
          const float qc = q[c];
          const float* rt = RT + (size_t)c * nrp + jb;
          float* a = a4[c & 3];
          for (int l = 0; l < RB; ++l) {
            float df = qc - rt[l];
            a[l] = __builtin_fmaf(df, df, a[l]);
          }
        }
        for (int l = 0; l < RB; ++l) acc[l] = (a4[0][l] + a4[1][l]) + (a4[2][l] + a4[3][l]);
      }
      const float kth = ld[k - 1];
      for (int l = 0; l < RB; ++l) {
        size_t j = jb + l;
        if (j < (size_t)nr && acc[l] < kth && acc[l] < ld[k - 1]) list_insert(ld, li, k, acc[l], (int)j);
      }
    }
  }
  free(RT);
}

/* FP64 brute force (recall ground truth): dist = 1 - <q,r>/(|q||r|) or squared L2, all in double. */
void oracle_knn_f64(const double* Q, int nq, const double* R, int nr, int d, int k, int metric, int* idx,
                    double* dist) {
  double* rn = (double*)malloc(sizeof(double) * (size_t)(nr > 0 ? nr : 1));
  for (int j = 0; j < nr; ++j) {
    double s = 0;
    for (int c = 0; c < d; ++c) s += R[(size_t)j * d + c] * R[(size_t)j * d + c];
    rn[j] = sqrt(s);
  }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int i = 0; i < nq; ++i) {
    double* ld = dist + (size_t)i * k;
    int* li = idx + (size_t)i * k;
    for (int p = 0; p < k; ++p) {
      ld[p] = INFINITY;
      li[p] = -1;
    }
    double qn = 0;
    for (int c = 0; c < d; ++c) qn += Q[(size_t)i * d + c] * Q[(size_t)i * d + c];
    qn = sqrt(qn);
    for (int j = 0; j < nr; ++j) {
      double s = 0;
      if (metric == 0) {
        for (int c = 0; c < d; ++c) s += Q[(size_t)i * d + c] * R[(size_t)j * d + c];
        double den = qn * rn[j];
        s = den > 0 ? 1.0 - s / den : 1.0;
      } else {
        for (int c = 0; c < d; ++c) {
          double df = Q[(size_t)i * d + c] - R[(size_t)j * d + c];
          s += df * df;
        }
      }
      if (s < ld[k - 1]) {
        int p = k - 1;
        while (p > 0 && ld[p - 1] > s) {
          ld[p] = ld[p - 1];
          li[p] = li[p - 1];
          --p;
        }
        ld[p] = s;
        li[p] = j;
      }
    }
  }
  free(rn);
}

/* ------------------------------------------------------------------ pareDownHubEdges
 * src/edgeFilter.cpp:23-60.  CSC matrix (p, i, x), rowN = current row degrees (mutated, as the reference
 * does at :52).  arma::sort_index(.,"descend") at :47 is an unstable std::sort; the restatement fixes the
 * tie rule to STABLE descending (equal row degrees keep their order of appearance in the column).     */
typedef struct {
  int v;
  int pos;
} deg_pos;
static int cmp_deg_desc(const void* a, const void* b) {
  const deg_pos* x = (const deg_pos*)a;
  const deg_pos* y = (const deg_pos*)b;
  if (x->v != y->v) return y->v - x->v;
  return x->pos - y->pos;
}
void oracle_pare_down_hub_edges(int ncols, const int* p, const int* i, double* x, int* rowN, int k, int klow) {
  if (klow < 0) klow = k;
  int cap = 0;
  for (int g = 0; g < ncols; ++g)
    if (p[g + 1] - p[g] > cap) cap = p[g + 1] - p[g];
  deg_pos* cid = (deg_pos*)malloc(sizeof(deg_pos) * (size_t)(cap > 0 ? cap : 1));
  for (int g = 0; g < ncols; ++g) {
    int p0 = p[g], p1 = p[g + 1];
    int nedges = p1 - p0;
    if (nedges <= k) continue;
    for (int j = p0, jl = 0; j < p1; ++j, ++jl) {
      cid[jl].v = rowN[i[j]];
      cid[jl].pos = jl;
    }
    qsort(cid, (size_t)(p1 - p0), sizeof(deg_pos), cmp_deg_desc);
    int j = 0;
    while (nedges > k && j < p1 - p0 && cid[j].v > klow) {
      x[cid[j].pos + p0] = 0;
      rowN[i[cid[j].pos + p0]]--;
      nedges--;
      j++;
    }
  }
  free(cid);
}

/* ------------------------------------------------------------------ cpcaF
 * src/cpca.cpp:16-101 with eigenvR given (the only branch conos reaches, R/conos.R:185-187).
 * cov: p x p x k column-major slices; ng[k]; eigv: p x ncomp column-major initial vectors.
 * out: D (ncomp x k column-major == R's t(D)), CPC (p x ncomp column-major), niter[ncomp].             */
void oracle_cpca(const double* cov, int p, int k, const double* ng, int ncomp, int maxit, double tol,
                 const double* eigv, double* D, double* CPC, double* niter) {
  size_t pp = (size_t)p * p;
  double* Qw = (double*)calloc(pp, sizeof(double));
  double* S = (double*)malloc(pp * sizeof(double));
  double* q = (double*)malloc(sizeof(double) * p);
  double* w = (double*)malloc(sizeof(double) * p);
  double* w2 = (double*)malloc(sizeof(double) * p);
  double* t = (double*)malloc(sizeof(double) * p);
  double* dv = (double*)malloc(sizeof(double) * k);
  for (int a = 0; a < p; ++a) Qw[(size_t)a * p + a] = 1.0;
  for (int comp = 0; comp < ncomp; ++comp) {
    for (int a = 0; a < p; ++a) q[a] = eigv[(size_t)comp * p + a];
    for (int s = 0; s < k; ++s) {
      const double* C = cov + pp * s;
      for (int a = 0; a < p; ++a) t[a] = 0;
      for (int b = 0; b < p; ++b) {
        double qb = q[b];
        for (int a = 0; a < p; ++a) t[a] += C[(size_t)b * p + a] * qb;
      }
      double v = 0;
      for (int a = 0; a < p; ++a) v += q[a] * t[a];
      dv[s] = v;
    }
    double cost0 = 0;
    int it = 0;
    for (it = 0; it < maxit; ++it) {
      memset(S, 0, pp * sizeof(double));
      for (int s = 0; s < k; ++s) {
        const double* C = cov + pp * s;
        double f = ng[s] / dv[s];
        for (size_t e = 0; e < pp; ++e) S[e] += C[e] * f;
      }
      for (int a = 0; a < p; ++a) w[a] = 0;
      for (int b = 0; b < p; ++b) {
        double qb = q[b];
        for (int a = 0; a < p; ++a) w[a] += S[(size_t)b * p + a] * qb;
      }
      if (comp != 0) {
        for (int a = 0; a < p; ++a) w2[a] = 0;
        for (int b = 0; b < p; ++b) {
          double wb = w[b];
          for (int a = 0; a < p; ++a) w2[a] += Qw[(size_t)b * p + a] * wb;
        }
        memcpy(w, w2, sizeof(double) * p);
      }
      double nn = 0;
      for (int a = 0; a < p; ++a) nn += w[a] * w[a];
      nn = sqrt(nn);
      for (int a = 0; a < p; ++a) q[a] = w[a] / nn;
      for (int s = 0; s < k; ++s) {
        const double* C = cov + pp * s;
        for (int a = 0; a < p; ++a) t[a] = 0;
        for (int b = 0; b < p; ++b) {
          double qb = q[b];
          for (int a = 0; a < p; ++a) t[a] += C[(size_t)b * p + a] * qb;
        }
        double v = 0;
        for (int a = 0; a < p; ++a) v += q[a] * t[a];
        dv[s] = v;
      }
      double cost = 0;
      for (int s = 0; s < k; ++s) cost += log(dv[s]) * ng[s];
      double delta = fabs((cost - cost0) / cost);
      if (delta < tol) break;
      cost0 = cost;
    }
    niter[comp] = it;
    for (int s = 0; s < k; ++s) D[(size_t)s * ncomp + comp] = dv[s];
    for (int a = 0; a < p; ++a) CPC[(size_t)comp * p + a] = q[a];
    for (int b = 0; b < p; ++b)
      for (int a = 0; a < p; ++a) Qw[(size_t)b * p + a] -= q[a] * q[b];
  }
  free(Qw); free(S); free(q); free(w); free(w2); free(t); free(dv);
}

/* ------------------------------------------------------------------ edge re-balancing
 * src/edge_rebalancing.cpp:14-36 and :39-48.  m is n_cells x n_levels column-major (R matrix).       */
void oracle_sum_weight_matrix(const double* weights, const int* row_inds, const int* col_inds, int n_edges,
                              const int* factor_levels, int n_cells, int n_levels, int normalize, double* m) {
  memset(m, 0, sizeof(double) * (size_t)n_cells * n_levels);
  for (int e = 0; e < n_edges; ++e) {
    int r = row_inds[e], c = col_inds[e];
    int cf = factor_levels[c] - 1, rf = factor_levels[r] - 1;
    m[(size_t)cf * n_cells + r] += weights[e];
    m[(size_t)rf * n_cells + c] += weights[e];
  }
  if (!normalize) return;
  for (int r = 0; r < n_cells; ++r) {
    double rs = 0;
    for (int l = 0; l < n_levels; ++l) rs += m[(size_t)l * n_cells + r];
    double dn = rs > 1e-10 ? rs : 1e-10;
    for (int l = 0; l < n_levels; ++l) m[(size_t)l * n_cells + r] /= dn;
  }
}
void oracle_adjust_weights(double* weights, const int* row_inds, const int* col_inds, int n_edges,
                           const int* factor_levels, int n_cells, const double* dividers) {
  for (int e = 0; e < n_edges; ++e) {
    int r = row_inds[e], c = col_inds[e];
    int cf = factor_levels[c] - 1, rf = factor_levels[r] - 1;
    weights[e] /= sqrt(dividers[(size_t)cf * n_cells + r] * dividers[(size_t)rf * n_cells + c]);
  }
}

/* ------------------------------------------------------------------ referenceWij
 * src/edgeweights.cpp:25-184.  Edges (from sorted ascending, to, d); the constructor squares d (:59).
 * Output: dense-free triplets of the final sparse matrix BEFORE duplicate summation, exactly as getWIJ
 * emits them (:147-157): out_from/out_to/out_w sized 2*n_edges; returns the number of triplets.
 * searchReverse (:98-106) is restated as written (it scans the vertex's own list for a self edge).    */
int oracle_reference_wij(const int* from, const int* to, const double* d, int n_edges, double perplexity,
                         int* out_from, int* out_to, double* out_w) {
  if (n_edges == 0) return 0;
  int n_vertices = from[n_edges - 1] + 1;
  int cap = 2 * n_edges;
  int* head = (int*)malloc(sizeof(int) * n_vertices);
  int* next = (int*)malloc(sizeof(int) * cap);
  int* reverse = (int*)malloc(sizeof(int) * cap);
  for (int v = 0; v < n_vertices; ++v) head[v] = -1;
  int ne = 0;
  for (int x = 0; x < n_vertices; ++x) {
    while (ne < n_edges && from[ne] == x) {
      out_from[ne] = x;
      out_to[ne] = to[ne];
      out_w[ne] = d[ne] * d[ne];
      next[ne] = head[x];
      reverse[ne] = -1;
      head[x] = ne++;
    }
  }
  int n_listed = ne; /* edges whose `from` was not ascending are silently ignored, as in the reference */
  double logp = log(perplexity);
  for (int id = 0; id < n_vertices; ++id) {
    double beta = 1, lo = -1, hi = -1, sw = 0, tmp;
    for (int iter = 0; iter < 200; ++iter) {
      double H = 0;
      sw = 0;
      for (int p = head[id]; p >= 0; p = next[p]) {
        sw += tmp = exp(-beta * out_w[p]);
        H += beta * (out_w[p] * tmp);
      }
      H = (H / sw) + log(sw);
      if (fabs(H - logp) < 1e-5) break;
      if (H > logp) {
        lo = beta;
        if (hi < 0) beta *= 2; else beta = (beta + hi) / 2;
      } else {
        hi = beta;
        if (lo < 0) beta /= 2; else beta = (lo + beta) / 2;
      }
    }
    sw = 0;
    for (int p = head[id]; p >= 0; p = next[p]) sw += out_w[p] = exp(-beta * out_w[p]);
    for (int p = head[id]; p >= 0; p = next[p]) out_w[p] /= sw;
  }
  for (int id = 0; id < n_vertices; ++id) {
    for (int p = head[id]; p >= 0; p = next[p]) {
      int q;
      for (q = head[id]; q >= 0; q = next[q])
        if (out_to[q] == id) break;
      reverse[p] = q;
    }
  }
  ne = n_listed;
  for (int id = 0; id != n_vertices; ++id) {
    for (int p = head[id]; p >= 0; p = next[p]) {
      int y = out_to[p];
      int q = reverse[p];
      if (q == -1) {
        out_from[ne] = y;
        out_to[ne] = id;
        out_w[ne] = 0;
        next[ne] = head[y];
        reverse[ne] = p;
        q = reverse[p] = head[y] = ne++;
      }
      if (id > y) {
        double a = (out_w[p] + out_w[q]) / 2;
        out_w[p] = out_w[q] = a;
      }
    }
  }
  free(head); free(next); free(reverse);
  return ne;
}

/* ------------------------------------------------------------------ graph diffusion
 * src/propagate_labels.cpp:120-175 (smooth_count_matrix_c) with the edge weights of parse_edges (:40-62):
 * w' = (w - min_w) / max_w, min_w = max(min w, 0), max_w = max w - min_w; length = 1 - w' (negative: error -1);
 * per iteration every edge adds exp(-fading * (length + fading_const)) times the OLD row of one end to the new row
 * of the other (unless that vertex is fixed), rows are divided by 1 + sum of the weights when normalize, and the
 * loop stops BEFORE taking the new matrix when max |new - old| < tol (:166-170).  cm is n_vertices x n_cols
 * column-major (R matrix), updated in place; is_fixed NULL or n_vertices bytes.  Returns the iteration index at
 * which the loop stopped (max_n_iters if it ran out), inf_norm = the last norm.                                   */
int oracle_smooth_matrix(const int* from, const int* to, const double* weights, long long n_edges, int n_vertices,
                         double* cm, int n_cols, const unsigned char* is_fixed, int max_n_iters, double fading,
                         double fading_const, double tol, int normalize, double* inf_norm_out) {
  if (n_vertices == 0 || n_cols == 0) return 0;
  double mn = 1e300, mx = -1e300;
  for (long long e = 0; e < n_edges; ++e) { if (weights[e] < mn) mn = weights[e]; if (weights[e] > mx) mx = weights[e]; }
  double min_w = mn > 0.0 ? mn : 0.0, max_w = mx - min_w;
  double* ew = (double*)malloc(sizeof(double) * (size_t)(n_edges > 0 ? n_edges : 1));
  for (long long e = 0; e < n_edges; ++e) {
    double len = 1.0 - (weights[e] - min_w) / max_w;
    if (len < 0) { free(ew); return -1; }
    ew[e] = exp(-fading * (len + fading_const));
  }
  size_t N = (size_t)n_vertices * n_cols;
  double* nw = (double*)malloc(sizeof(double) * N);
  double* sw = (double*)malloc(sizeof(double) * (size_t)n_vertices);
  for (int v = 0; v < n_vertices; ++v) sw[v] = 1.0;
  double inf_norm = 1e10;
  int iter = 0;
  for (iter = 0; iter < max_n_iters; ++iter) {
    memcpy(nw, cm, sizeof(double) * N);
    for (long long e = 0; e < n_edges; ++e) {
      int a = from[e], b = to[e];
      double w = ew[e];
      if (!is_fixed || !is_fixed[a]) {
        for (int c = 0; c < n_cols; ++c) nw[(size_t)c * n_vertices + a] += cm[(size_t)c * n_vertices + b] * w;
        if (normalize) sw[a] += w;
      }
      if (!is_fixed || !is_fixed[b]) {
        for (int c = 0; c < n_cols; ++c) nw[(size_t)c * n_vertices + b] += cm[(size_t)c * n_vertices + a] * w;
        if (normalize) sw[b] += w;
      }
    }
    if (normalize)
      for (int v = 0; v < n_vertices; ++v) {
        for (int c = 0; c < n_cols; ++c) nw[(size_t)c * n_vertices + v] /= sw[v];
        sw[v] = 1.0;
      }
    inf_norm = 0.0;
    for (size_t t = 0; t < N; ++t) { double d = fabs(nw[t] - cm[t]); if (d > inf_norm) inf_norm = d; }
    if (inf_norm < tol) break;
    memcpy(cm, nw, sizeof(double) * N);
  }
  if (inf_norm_out) *inf_norm_out = inf_norm;
  free(ew); free(nw); free(sw);
  return iter;
}
