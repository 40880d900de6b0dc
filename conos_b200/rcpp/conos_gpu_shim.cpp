// Rcpp shim a conos maintainer adds to src/ (NOT compiled in this repository: R and Rcpp are not installed here; the
// same C entry points are exercised by tests/cuda/job_test.cpp -- a C++ program without Python -- and by every GPU
// test through ctypes).  It keeps the reference's registered .Call entry points (src/RcppExports.cpp:349-370), routes
// the hot-path ones to libconosgpu.so through its C ABI (include/conos_gpu.h) and adds the entries that replace the
// N2R calls and the per-pair R loops.  Marshalling only: dense matrices are R's column-major doubles, dgCMatrix
// travels by its p / i / x slots, errors become R errors (Rcpp::stop, like BEGIN_RCPP / END_RCPP in the reference).
//
// Build: PKG_LIBS += -L$(CONOS_GPU_LIB) -lconosgpu ; PKG_CPPFLAGS += -I$(CONOS_GPU_INCLUDE)
#include <Rcpp.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "conos_gpu.h"

using namespace Rcpp;

// One job per R process, created lazily in the PARENT process (a CUDA context does not survive fork(): the GPU path
// replaces the mclapply / plapply loops over pairs, R/conclass.R:224 and :1074, by one batched native call).
static cg_job* g_job = nullptr;
static int g_job_devices = 0;
static cg_job* job_get(int n_devices) {
  if (g_job && g_job_devices != n_devices) {
    cg_job_free(g_job);
    g_job = nullptr;
  }
  if (!g_job) {
    if (cg_job_create(n_devices, &g_job) != 0) stop(cg_last_error(nullptr));
    g_job_devices = n_devices;
  }
  return g_job;
}
static cg_context* ctx_get() { return cg_job_context(job_get(g_job ? g_job_devices : 1), 0); }
static void check(int rc, cg_context* ctx = nullptr) {
  if (rc != 0) stop(cg_last_error(ctx ? ctx : ctx_get()));
}
static int metric_id(const std::string& t) {
  if (t == "angular") return CG_METRIC_ANGULAR;
  if (t == "L2") return CG_METRIC_L2;
  stop("only the following distance metrics are currently supported: ['L2' 'angular']");
  return 0;
}

// dgCMatrix from kNN lists: column q = query q, rows = its neighbours (ascending row index inside a column, as
// Matrix stores it), x = distance -- the layout N2R::crossKnn / N2R::Knn return (nrow(index) x nrow(queries))
static S4 knn_to_dgc(const std::vector<int32_t>& idx, const std::vector<float>& dist, int n_query, int n_index, int k) {
  std::vector<int> p(n_query + 1, 0);
  for (int q = 0; q < n_query; ++q) {
    int c = 0;
    for (int t = 0; t < k; ++t) c += idx[(size_t)q * k + t] >= 0;
    p[q + 1] = p[q] + c;
  }
  IntegerVector ri(p[n_query]);
  NumericVector rx(p[n_query]);
  for (int q = 0; q < n_query; ++q) {
    std::vector<std::pair<int, double>> col;
    for (int t = 0; t < k; ++t)
      if (idx[(size_t)q * k + t] >= 0) col.emplace_back(idx[(size_t)q * k + t], (double)dist[(size_t)q * k + t]);
    std::sort(col.begin(), col.end());
    for (size_t t = 0; t < col.size(); ++t) {
      ri[p[q] + t] = col[t].first;
      rx[p[q] + t] = col[t].second;
    }
  }
  S4 m("dgCMatrix");
  m.slot("i") = ri;
  m.slot("p") = IntegerVector(p.begin(), p.end());
  m.slot("x") = rx;
  m.slot("Dim") = IntegerVector::create(n_index, n_query);
  return m;
}

// [[Rcpp::export]]   replaces N2R::crossKnn(mA, mB, k, 1, FALSE, indexType) at R/conos.R:851-855 (exact search)
S4 crossKnnExact(NumericMatrix mA, NumericMatrix mB, int k, std::string indexType = "angular") {
  std::vector<int32_t> idx((size_t)k * mA.nrow());
  std::vector<float> dist((size_t)k * mA.nrow());
  check(cg_cross_knn(ctx_get(), mA.begin(), mA.nrow(), mA.nrow(), mB.begin(), mB.nrow(), mB.nrow(), mA.ncol(), k,
                     metric_id(indexType), idx.data(), dist.data(), nullptr, nullptr));
  return knn_to_dgc(idx, dist, mA.nrow(), mB.nrow(), k);
}

// [[Rcpp::export]]   replaces N2R::Knn(m, k, 1, verbose = FALSE, indexType) at R/conos.R:595 (self included)
S4 knnExact(NumericMatrix m, int k, std::string indexType = "angular") {
  std::vector<int32_t> idx((size_t)k * m.nrow());
  std::vector<float> dist((size_t)k * m.nrow());
  check(cg_self_knn(ctx_get(), m.begin(), m.nrow(), m.nrow(), m.ncol(), k, metric_id(indexType), idx.data(),
                    dist.data()));
  return knn_to_dgc(idx, dist, m.nrow(), m.nrow(), k);
}

// [[Rcpp::export]]   getNeighborMatrix(p1, p2, k, k1, matching, metric, l2.sigma, cor.base)  R/conos.R:848-886
S4 getNeighborMatrixGPU(NumericMatrix p1, NumericMatrix p2, int k, int k1, std::string matching, std::string metric,
                        double l2_sigma, double cor_base, double min_similarity = 1e-5) {
  if (matching != "mNN" && matching != "NN") stop("Unrecognized type of NN matching:" + matching);
  int64_t nnz = 0;
  check(cg_neighbor_matrix(ctx_get(), p1.begin(), p1.nrow(), p1.nrow(), p2.begin(), p2.nrow(), p2.nrow(), p1.ncol(),
                           k, k1, matching == "mNN" ? CG_MATCH_MNN : CG_MATCH_NN, metric_id(metric), l2_sigma,
                           cor_base, min_similarity, &nnz));
  IntegerVector i(nnz), j(nnz);
  NumericVector x(nnz);
  check(cg_fetch_coo(ctx_get(), i.begin(), j.begin(), x.begin()));
  S4 m("dgTMatrix");  // triplets in CSC order, like as(adj, 'TsparseMatrix') at R/conos.R:880-884
  m.slot("i") = i;
  m.slot("j") = j;
  m.slot("x") = x;
  m.slot("Dim") = IntegerVector::create(p1.nrow(), p2.nrow());
  m.slot("Dimnames") = List::create(rownames(p1), rownames(p2));
  return m;
}

// [[Rcpp::export]]   same name and arguments as src/spcov.cpp:14
NumericMatrix spcov(S4 m, NumericVector cm) {
  IntegerVector dims = m.slot("Dim"), p = m.slot("p"), i = m.slot("i");
  NumericVector x = m.slot("x");
  NumericMatrix out(dims[1], dims[1]);
  check(cg_spcov(ctx_get(), dims[0], dims[1], p.begin(), i.begin(), x.begin(), cm.begin(), out.begin()));
  return out;
}

// [[Rcpp::export]]   same name and arguments as src/cpca.cpp:16
List cpcaF(NumericVector cov /* p x p x k array */, NumericVector ng, int ncomp = 10, int maxit = 1000,
           double tol = 1e-6, Nullable<NumericMatrix> eigenvR = R_NilValue, bool verbose = true) {
  IntegerVector d = cov.attr("dim");
  if (eigenvR.isNull()) stop("the device cpcaF needs eigenvR (conos always supplies it, R/conos.R:185-187)");
  NumericMatrix ev(eigenvR.get());
  NumericMatrix D(ncomp, d[2]), CPC(d[0], ncomp);
  NumericVector niter(ncomp);
  check(cg_cpca(ctx_get(), cov.begin(), d[0], d[2], ng.begin(), ncomp, maxit, tol, ev.begin(), D.begin(), CPC.begin(),
                niter.begin()));
  return List::create(_["D"] = D, _["CPC"] = CPC, _["niter"] = niter);
}

// [[Rcpp::export]]   same name and arguments as src/edgeFilter.cpp:23 (x is returned, rowN mutated in place)
NumericVector pareDownHubEdges(S4 sY, IntegerVector rowN, int k, int klow = -1) {
  IntegerVector dims = sY.slot("Dim"), p = sY.slot("p"), i = sY.slot("i");
  NumericVector x = sY.slot("x");
  check(cg_pare_down_hub_edges(ctx_get(), dims[1], p.begin(), i.begin(), x.begin(), rowN.begin(), k, klow));
  return x;
}

// [[Rcpp::export]]   src/edge_rebalancing.cpp:14
NumericMatrix getSumWeightMatrix(NumericVector weights, IntegerVector row_inds, IntegerVector col_inds,
                                 IntegerVector factor_levels, bool normalize = true) {
  const int n_levels = max(factor_levels);
  NumericMatrix m(factor_levels.size(), n_levels);
  check(cg_sum_weight_matrix(ctx_get(), weights.begin(), row_inds.begin(), col_inds.begin(), weights.size(),
                             factor_levels.begin(), factor_levels.size(), n_levels, normalize, m.begin()));
  return m;
}

// [[Rcpp::export]]   src/edge_rebalancing.cpp:39
NumericVector adjustWeightsByCellBalancingC(NumericVector weights, IntegerVector row_inds, IntegerVector col_inds,
                                            IntegerVector factor_levels, NumericMatrix dividers) {
  NumericVector w = clone(weights);
  check(cg_adjust_weights(ctx_get(), w.begin(), row_inds.begin(), col_inds.begin(), w.size(), factor_levels.begin(),
                          factor_levels.size(), dividers.ncol(), dividers.begin()));
  return w;
}

// [[Rcpp::export]]   same name and arguments as src/edgeweights.cpp:160 (threads is ignored: the device is the pool)
S4 referenceWij(IntegerVector i, IntegerVector j, NumericVector d, int threads, double perplexity) {
  const int64_t n = i.size();
  std::vector<int32_t> of(2 * n), ot(2 * n);
  std::vector<double> ow(2 * n);
  int64_t n_out = 0;
  check(cg_reference_wij(ctx_get(), i.begin(), j.begin(), d.begin(), n, perplexity, of.data(), ot.data(), ow.data(),
                         &n_out));
  // setFromTriplets sums duplicates (src/edgeweights.cpp:178-181): Matrix::sparseMatrix does the same
  int nv = 0;
  for (int64_t t = 0; t < n; ++t) nv = std::max(nv, std::max(i[t], j[t]) + 1);
  Environment Matrix = Environment::namespace_env("Matrix");
  Function sparseMatrix = Matrix["sparseMatrix"];
  IntegerVector ri(of.begin(), of.begin() + n_out), rj(ot.begin(), ot.begin() + n_out);
  NumericVector rx(ow.begin(), ow.begin() + n_out);
  return sparseMatrix(_["i"] = ri + 1, _["j"] = rj + 1, _["x"] = rx, _["dims"] = IntegerVector::create(nv, nv));
}

// ---- diffusion over the graph.  The reference's entry points take vertex NAMES
// (StringMatrix edge_verts) and number them in C++ (parse_edges / order_strings, src/propagate_labels.cpp:33-62); the
// numbering is done here the same way, the arithmetic runs on the device.
static std::vector<int32_t> vertex_ids(const StringMatrix& edge_verts, std::vector<std::string> names,
                                       std::unordered_map<std::string, int32_t>& ids) {
  if (names.empty())
    for (int e = 0; e < edge_verts.nrow(); ++e) {
      names.push_back(as<std::string>(edge_verts(e, 0)));
      names.push_back(as<std::string>(edge_verts(e, 1)));
    }
  for (const auto& s : names) ids.emplace(s, (int32_t)ids.size());
  std::vector<int32_t> ft(2 * (size_t)edge_verts.nrow());
  for (int e = 0; e < edge_verts.nrow(); ++e) {
    ft[e] = ids.at(as<std::string>(edge_verts(e, 0)));
    ft[(size_t)edge_verts.nrow() + e] = ids.at(as<std::string>(edge_verts(e, 1)));
  }
  return ft;
}

// [[Rcpp::export]]   same name and arguments as src/propagate_labels.cpp:177
SEXP smooth_count_matrix(StringMatrix edge_verts, std::vector<double> edge_weights, NumericMatrix count_matrix,
                         std::vector<bool> is_label_fixed, int max_n_iters = 10, double diffusion_fading = 1.0,
                         double diffusion_fading_const = 0.1, double tol = 1e-3, bool verbose = true,
                         bool normalize = false) {
  if (count_matrix.nrow() == 0 || count_matrix.ncol() == 0) { warning("Empty matrix passed"); return count_matrix; }
  if ((size_t)edge_verts.nrow() != edge_weights.size()) stop("Size of edge_verts must match size of edge_weights");
  if (edge_verts.nrow() == 0) { warning("Empty graph passed"); return count_matrix; }
  if (edge_verts.ncol() != 2) stop("Matrix edge_verts must have exactly 2 columns");
  std::unordered_map<std::string, int32_t> ids;
  const std::vector<int32_t> ft = vertex_ids(edge_verts, as<std::vector<std::string>>(rownames(count_matrix)), ids);
  NumericMatrix res = clone(count_matrix);
  std::vector<uint8_t> fx(is_label_fixed.begin(), is_label_fixed.end());
  int32_t it = 0;
  double nrm = 0;
  const int64_t E = edge_verts.nrow();
  check(cg_smooth_matrix(ctx_get(), nullptr, ft.data(), ft.data() + E, edge_weights.data(), E, res.nrow(), res.begin(),
                         res.ncol(), fx.empty() ? nullptr : fx.data(), max_n_iters, diffusion_fading,
                         diffusion_fading_const, tol, normalize, &it, &nrm));
  if (verbose) Rcout << "Stop after " << it << " iterations. Norm: " << nrm << std::endl;
  return res;
}

// [[Rcpp::export]]   same name and arguments as src/propagate_labels.cpp:67
NumericMatrix propagate_labels(StringMatrix edge_verts, std::vector<double> edge_weights, StringVector vert_labels,
                               int max_n_iters = 10, bool verbose = true, double diffusion_fading = 10,
                               double diffusion_fading_const = 0.5, double tol = 5e-3,
                               bool fixed_initial_labels = false) {
  if ((size_t)edge_verts.nrow() != edge_weights.size() || edge_verts.ncol() != 2)
    stop("Incorrect dimension of input vectors");
  std::unordered_map<std::string, int32_t> ids, label_ids;
  const std::vector<int32_t> ft = vertex_ids(edge_verts, {}, ids);
  for (int t = 0; t < vert_labels.size(); ++t) label_ids.emplace(as<std::string>(vert_labels[t]), (int32_t)label_ids.size());
  std::vector<int32_t> lab(ids.size(), -1);
  StringVector vn = vert_labels.names();
  for (int t = 0; t < vert_labels.size(); ++t)
    lab[ids.at(as<std::string>(vn[t]))] = label_ids.at(as<std::string>(vert_labels[t]));
  NumericMatrix res((int)ids.size(), (int)label_ids.size());
  int32_t it = 0;
  double nrm = 0;
  const int64_t E = edge_verts.nrow();
  check(cg_propagate_labels(ctx_get(), nullptr, ft.data(), ft.data() + E, edge_weights.data(), E, (int32_t)ids.size(),
                            lab.data(), (int32_t)label_ids.size(), max_n_iters, diffusion_fading, diffusion_fading_const,
                            tol, fixed_initial_labels, res.begin(), &it, &nrm));
  if (verbose) Rcout << "Stop after " << it << " iterations. Norm: " << nrm << std::endl;
  CharacterVector rn(ids.size()), cn(label_ids.size());
  for (const auto& kv : ids) rn[kv.second] = kv.first;
  for (const auto& kv : label_ids) cn[kv.second] = kv.first;
  rownames(res) = rn;
  colnames(res) = cn;
  return res;
}

// [[Rcpp::export]]   the one native call Conos$buildGraph makes instead of updatePairs + its two papply loops + rbind +
// graph_from_edgelist + simplify + adjustWeightsByCellBalancing (R/conclass.R:199-368).
// samples: list of list(p, i, x, Dim, v, qv, pca); pairs: list of list(a, b, genes_a, genes_b [, k] [, rot] [, u, v,
// rows_u, rows_v]) in combn order; params: list; n_devices: GPUs of this node to use.
// Returns the graph and, per pair, the reduction with every field of self$pairs[[space]] (R/conclass.R:1090-1092).
List buildGraphGPU(List samples, List pairs, List params, int n_devices = 1) {
  cg_job* job = job_get(n_devices);
  cg_context* ctx0 = cg_job_context(job, 0);
  const int S = samples.size();
  std::vector<cg_sample> cs(S);
  for (int s = 0; s < S; ++s) {
    List sm = samples[s];
    IntegerVector dim = sm["Dim"], p = sm["p"], i = sm["i"];
    NumericVector x = sm["x"], v = sm["v"], qv = sm["qv"];
    cs[s] = cg_sample{dim[0], dim[1], p.begin(), i.begin(), x.begin(), v.begin(), qv.begin(), nullptr, 0, dim[0]};
    if (sm.containsElementNamed("pca") && !Rf_isNull(sm["pca"])) {
      NumericMatrix pca = sm["pca"];
      cs[s].pca = pca.begin();
      cs[s].n_pcs = pca.ncol();
      cs[s].pca_ld = pca.nrow();
    }
  }
  const int P = pairs.size();
  std::vector<cg_pair> cp(P);
  for (int q = 0; q < P; ++q) {
    List pr = pairs[q];
    IntegerVector ga = pr["genes_a"], gb = pr["genes_b"];
    cp[q] = cg_pair{};
    cp[q].a = as<int>(pr["a"]);
    cp[q].b = as<int>(pr["b"]);
    cp[q].n_genes = ga.size();
    cp[q].genes_a = ga.begin();
    cp[q].genes_b = gb.begin();
    if (pr.containsElementNamed("k")) cp[q].k = as<int>(pr["k"]);  // k.cur, R/conclass.R:230-233
    if (pr.containsElementNamed("rot")) {  // cached self$pairs[[space]][[name]]$CPC: its rownames ARE the pair's genes
      NumericMatrix rot = pr["rot"];
      cp[q].rot = rot.begin();
      cp[q].rot_rows = rot.nrow();
      cp[q].rot_cols = rot.ncol();
    }
    if (pr.containsElementNamed("u")) {    // cached self$pairs$CCA[[name]]$u / $v (R/conclass.R:268-270)
      NumericMatrix u = pr["u"], v = pr["v"];
      IntegerVector ru = pr["rows_u"], rv = pr["rows_v"];
      cp[q].u = u.begin();
      cp[q].v = v.begin();
      cp[q].rows_u = ru.begin();
      cp[q].rows_v = rv.begin();
      cp[q].n_u = u.nrow();
      cp[q].n_v = v.nrow();
      cp[q].uv_cols = u.ncol();
    }
  }
  cg_params par;
  cg_params_default(&par);
  par.k = as<int>(params["k"]);
  par.k1 = as<int>(params["k1"]);
  par.k_self = as<int>(params["k.self"]);
  par.ncomps = as<int>(params["ncomps"]);
  par.space = as<int>(params["space"]);
  par.matching = as<int>(params["matching"]);
  par.metric = as<int>(params["metric"]);
  par.var_scale = as<bool>(params["var.scale"]);
  par.common_centering = as<bool>(params["common.centering"]);
  par.score_component_variance = as<bool>(params["score.component.variance"]);
  par.k_self_weight = as<double>(params["k.self.weight"]);
  par.l2_sigma = as<double>(params["l2.sigma"]);
  par.cor_base = as<double>(params["cor.base"]);
  cg_graph* g = nullptr;
  if (cg_job_build_graph(job, cs.data(), S, cp.data(), P, &par, &g) != 0) stop(cg_job_last_error(job));
  int32_t nv = 0;
  int64_t ne = 0;
  cg_graph_sizes(g, &nv, &ne);
  // balancing (R/conclass.R:346-368): factor.per.vertex = 1-based level of every graph vertex, computed by the R side
  // from balancing.factor.per.cell / per.sample / getDatasetPerCell once the vertex order is known
  if (params.containsElementNamed("balance") && as<bool>(params["balance"])) {
    IntegerVector vertices0(nv);
    check(cg_graph_fetch(ctx0, g, vertices0.begin(), nullptr, nullptr, nullptr, nullptr), ctx0);
    IntegerVector level_of_cell = params["factor.level.per.cell"];  // 1-based levels in global cell order
    IntegerVector fac(nv);
    for (int v = 0; v < nv; ++v) fac[v] = level_of_cell[vertices0[v]];
    check(cg_graph_balance(ctx0, g, fac.begin(), max(level_of_cell), as<bool>(params["balance.edge.weights"]),
                           as<double>(params["same.factor.downweight"]), 50), ctx0);
  }
  IntegerVector vertices(nv), from(ne), to(ne), type(ne);
  NumericVector weight(ne);
  check(cg_graph_fetch(ctx0, g, vertices.begin(), from.begin(), to.begin(), weight.begin(), type.begin()), ctx0);
  cg_graph_free(g);
  // ---- self$pairs[[space]] write-back: every field the reference keeps
  List reductions(P);
  for (int q = 0; q < P; ++q) {
    int32_t dev = 0, slot = 0;
    cg_job_pair_location(job, q, &dev, &slot);
    cg_context* cx = cg_job_context(job, dev);
    if (par.space == CG_SPACE_CCA) {
      int32_t nu = 0, nvv = 0, d = 0, ng = 0, dl = 0;
      check(cg_pair_cca(cx, slot, &nu, &nvv, &d, nullptr, nullptr, nullptr, nullptr, nullptr), cx);
      if (nu == 0) continue;  // supplied by the caller: nothing new
      NumericMatrix u(nu, d), v(nvv, d);
      IntegerVector ru(nu), rv(nvv);
      NumericVector sigma(d);
      check(cg_pair_cca(cx, slot, nullptr, nullptr, nullptr, u.begin(), v.begin(), ru.begin(), rv.begin(),
                        sigma.begin()), cx);
      check(cg_pair_cca_loadings(cx, slot, &ng, &dl, nullptr, nullptr, nullptr), cx);
      NumericMatrix ul(ng, dl), vl(ng, dl), nvm(dl, 2);
      check(cg_pair_cca_loadings(cx, slot, nullptr, nullptr, ul.begin(), vl.begin(), nvm.begin()), cx);
      reductions[q] = List::create(_["d"] = sigma, _["u"] = u, _["v"] = v, _["ul"] = ul, _["vl"] = vl,
                                   _["nv"] = List::create(nvm(_, 0), nvm(_, 1)), _["rows_u"] = ru, _["rows_v"] = rv);
      continue;
    }
    int32_t nr = 0, nc = 0;
    check(cg_pair_reduction(cx, slot, &nr, &nc, nullptr, nullptr), cx);
    NumericMatrix rot(nr, nc), nvm(nc, 2);
    check(cg_pair_reduction(cx, slot, nullptr, nullptr, rot.begin(), par.score_component_variance ? nvm.begin() : nullptr),
          cx);
    List red = List::create(_["CPC"] = rot);   // R side: rownames(CPC) <- od.genes of the pair
    if (par.score_component_variance) red["nv"] = List::create(nvm(_, 0), nvm(_, 1));
    if (par.space == CG_SPACE_CPCA) {
      NumericMatrix D(nc, 2);   // t(D), src/cpca.cpp:97
      NumericVector niter(nc);
      check(cg_pair_cpca(cx, slot, nullptr, D.begin(), niter.begin()), cx);
      red["D"] = D;
      red["niter"] = niter;
    }
    reductions[q] = red;
  }
  // R side: g <- igraph::make_graph(rbind(from, to) + 1L, n = length(vertices), directed = FALSE);
  //         V(g)$name <- cells[vertices + 1]; E(g)$weight <- weight; E(g)$type <- type   (already simplified)
  return List::create(_["vertices"] = vertices, _["from"] = from, _["to"] = to, _["weight"] = weight,
                      _["type"] = type, _["pairs"] = reductions);
}

// ---- registration (what Rcpp::compileAttributes() generates into src/RcppExports.cpp for the entries above; the 16
// untouched entries of the reference's table -- src/RcppExports.cpp:349-370 -- stay as they are)
//
//   static const R_CallMethodDef CallEntries[] = {
//       {"_conos_spcov", (DL_FUNC) &_conos_spcov, 2},
//       {"_conos_cpcaF", (DL_FUNC) &_conos_cpcaF, 7},
//       {"_conos_pareDownHubEdges", (DL_FUNC) &_conos_pareDownHubEdges, 4},
//       {"_conos_getSumWeightMatrix", (DL_FUNC) &_conos_getSumWeightMatrix, 5},
//       {"_conos_adjustWeightsByCellBalancingC", (DL_FUNC) &_conos_adjustWeightsByCellBalancingC, 5},
//       {"_conos_referenceWij", (DL_FUNC) &_conos_referenceWij, 5},
//       {"_conos_propagate_labels", (DL_FUNC) &_conos_propagate_labels, 9},
//       {"_conos_smooth_count_matrix", (DL_FUNC) &_conos_smooth_count_matrix, 10},
//       {"_conos_crossKnnExact", (DL_FUNC) &_conos_crossKnnExact, 4},          // new
//       {"_conos_knnExact", (DL_FUNC) &_conos_knnExact, 3},                    // new
//       {"_conos_getNeighborMatrixGPU", (DL_FUNC) &_conos_getNeighborMatrixGPU, 9},  // new
//       {"_conos_buildGraphGPU", (DL_FUNC) &_conos_buildGraphGPU, 4},          // new
//       ... the reference's other 16 entries ...
//       {NULL, NULL, 0}};
//   RcppExport void R_init_conos(DllInfo* dll) {
//     R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
//     R_useDynamicSymbols(dll, FALSE);
//   }
//   // R/zzz.R: .onUnload <- function(libpath) { .Call('_conos_gpuShutdown'); library.dynam.unload('conos', libpath) }

// [[Rcpp::export]]   called from .onUnload: contexts, streams and the NCCL communicator go before the DLL does
void gpuShutdown() {
  if (g_job) cg_job_free(g_job);
  g_job = nullptr;
  g_job_devices = 0;
}
