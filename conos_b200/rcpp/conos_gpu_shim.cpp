// Rcpp shim a conos maintainer would add to src/ (NOT compiled in this repository: R and Rcpp are not
// installed here).  It keeps the reference's registered .Call entry points (src/RcppExports.cpp:349-370) and
// routes the hot-path ones to libconosgpu.so through its C ABI (include/conos_gpu.h).  Marshalling only: dense
// matrices are passed as R's column-major doubles, dgCMatrix by its p/i/x slots, errors become R errors.
//
// Build: PKG_LIBS += -L$(CONOS_GPU_LIB) -lconosgpu ; PKG_CPPFLAGS += -I$(CONOS_GPU_INCLUDE)
#include <Rcpp.h>
#include "conos_gpu.h"

using namespace Rcpp;

static cg_context* ctx_get() {
  // one context per R process, created lazily in the PARENT process: CUDA contexts do not survive fork(), so the
  // GPU path replaces the mclapply / plapply loops over pairs (R/conclass.R:224, 1074) by one batched call
  static cg_context* ctx = nullptr;
  if (!ctx && cg_init(0, &ctx) != 0) stop(cg_last_error(nullptr));
  return ctx;
}
static void check(int rc) {
  if (rc != 0) stop(cg_last_error(ctx_get()));
}

// [[Rcpp::export]]   replaces N2R::crossKnn(mA, mB, k, 1, FALSE, indexType) at R/conos.R:851-855
List crossKnnExact(NumericMatrix mA, NumericMatrix mB, int k, std::string indexType) {
  const int metric = indexType == "angular" ? CG_METRIC_ANGULAR : CG_METRIC_L2;
  IntegerMatrix idx(k, mA.nrow());   // column q = neighbours of query row q  (t() of the C layout [nA][k])
  std::vector<float> dist((size_t)k * mA.nrow());
  check(cg_cross_knn(ctx_get(), mA.begin(), mA.nrow(), mA.nrow(), mB.begin(), mB.nrow(), mB.nrow(), mA.ncol(), k,
                     metric, idx.begin(), dist.data(), nullptr, nullptr));
  // triplets (neighbour, query, dist) -> dgCMatrix nrow(mB) x nrow(mA), the layout N2R returns
  return List::create(_["i"] = idx, _["x"] = NumericVector(dist.begin(), dist.end()));
}

// [[Rcpp::export]]   getNeighborMatrix(p1, p2, k, k1, matching, metric, l2.sigma, cor.base)  R/conos.R:848-886
List getNeighborMatrixGPU(NumericMatrix p1, NumericMatrix p2, int k, int k1, std::string matching,
                          std::string metric, double l2_sigma, double cor_base, double min_similarity = 1e-5) {
  int64_t nnz = 0;
  check(cg_neighbor_matrix(ctx_get(), p1.begin(), p1.nrow(), p1.nrow(), p2.begin(), p2.nrow(), p2.nrow(), p1.ncol(),
                           k, k1, matching == "mNN" ? CG_MATCH_MNN : CG_MATCH_NN,
                           metric == "angular" ? CG_METRIC_ANGULAR : CG_METRIC_L2, l2_sigma, cor_base,
                           min_similarity, &nnz));
  IntegerVector i(nnz), j(nnz);
  NumericVector x(nnz);
  check(cg_fetch_coo(ctx_get(), i.begin(), j.begin(), x.begin()));
  return List::create(_["i"] = i, _["j"] = j, _["x"] = x, _["Dim"] = IntegerVector::create(p1.nrow(), p2.nrow()));
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

// [[Rcpp::export]]   the one native call Conos$buildGraph makes instead of its two papply loops
// samples: list of list(p, i, x, Dim, v, qv, pca); pairs: list of list(a, b, genes_a, genes_b [, rot]); params: list
List buildGraphGPU(List samples, List pairs, List params) {
  const int S = samples.size();
  std::vector<cg_sample> cs(S);
  for (int s = 0; s < S; ++s) {
    List sm = samples[s];
    IntegerVector dim = sm["Dim"], p = sm["p"], i = sm["i"];
    NumericVector x = sm["x"], v = sm["v"], qv = sm["qv"];
    NumericMatrix pca = sm["pca"];
    cs[s] = cg_sample{dim[0], dim[1], p.begin(), i.begin(), x.begin(), v.begin(), qv.begin(), pca.begin(),
                      (int32_t)pca.ncol(), (int32_t)pca.nrow()};
  }
  cg_panel* panel = nullptr;
  check(cg_panel_create(ctx_get(), cs.data(), S, &panel));
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
    if (pr.containsElementNamed("rot")) {  // cached self$pairs[[space]][[name]]$CPC
      NumericMatrix rot = pr["rot"];
      cp[q].rot = rot.begin();
      cp[q].rot_cols = rot.ncol();
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
  cg_edges* edges = nullptr;
  std::vector<int64_t> counts(P);
  check(cg_pairs_edges(ctx_get(), panel, cp.data(), P, &par, &edges, counts.data()));
  std::vector<int32_t> sidx(S);
  for (int s = 0; s < S; ++s) sidx[s] = s;
  check(cg_local_edges(ctx_get(), panel, sidx.data(), S, &par, &edges));
  std::vector<int32_t> offs(S + 1);
  cg_panel_offsets(panel, offs.data());
  cg_graph* g = nullptr;
  check(cg_assemble_graph(ctx_get(), edges, offs[S], &g));
  int32_t nv = 0;
  int64_t ne = 0;
  cg_graph_sizes(g, &nv, &ne);
  IntegerVector vertices(nv), from(ne), to(ne), type(ne);
  NumericVector weight(ne);
  check(cg_graph_fetch(ctx_get(), g, vertices.begin(), from.begin(), to.begin(), weight.begin(), type.begin()));
  cg_graph_free(g);
  cg_edges_free(edges);
  cg_panel_free(panel);
  // R side: g <- igraph::graph_from_edgelist(cbind(from, to) + 1L, directed = FALSE); V(g)$name <- cells[vertices+1]
  //         E(g)$weight <- weight; E(g)$type <- type      (already simplified: no call to simplify())
  return List::create(_["vertices"] = vertices, _["from"] = from, _["to"] = to, _["weight"] = weight,
                      _["type"] = type);
}
