// C++ host test of the drop-in boundary, no Python: links libconosgpu.so through include/conos_gpu.h only and runs
// the call sequence the Rcpp shim's buildGraphGPU makes (conos_b200/rcpp/conos_gpu_shim.cpp; reference:
// R/conclass.R:224-343, src/RcppExports.cpp:348-377) -- once call by call on one context, once as ONE cg_job call on
// one device and, when the box has them, on 2 devices -- and demands identical graphs (vertex order, edges, weights).
//   g++ -std=c++17 -O2 -Iinclude tests/cuda/job_test.cpp -Lconos_b200 -lconosgpu -Wl,-rpath,$PWD/conos_b200 -o build/job_test
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conos_gpu.h"

struct HostSample {
  int n_cells, n_genes;
  std::vector<int32_t> p, i;
  std::vector<double> x, v, qv, pca;
  int n_pcs;
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) / 9007199254740992.0;
}
static double nrand() { return std::sqrt(-2.0 * std::log(urand() + 1e-300)) * std::cos(6.283185307179586 * urand()); }

// cells of `n_types` latent types: type-specific marker genes on top of a sparse background (so that neighbours exist)
static HostSample make_sample(int n_cells, int n_genes, int n_types, int n_pcs, double shift) {
  HostSample s;
  s.n_cells = n_cells;
  s.n_genes = n_genes;
  s.n_pcs = n_pcs;
  std::vector<int> type(n_cells);
  for (int c = 0; c < n_cells; ++c) type[c] = (int)(urand() * n_types) % n_types;
  s.p.assign(1, 0);
  for (int g = 0; g < n_genes; ++g) {
    const int marker_of = g % (2 * n_types);  // genes >= n_types in this residue are background
    for (int c = 0; c < n_cells; ++c) {
      const bool marker = marker_of == type[c];
      const double pr = marker ? 0.7 : 0.06;
      if (urand() < pr) {
        s.i.push_back(c);
        s.x.push_back(std::log10(1.0 + (marker ? 40.0 : 8.0) * urand() + shift));
      }
    }
    s.p.push_back((int32_t)s.i.size());
  }
  s.v.resize(n_genes);
  s.qv.resize(n_genes);
  for (int g = 0; g < n_genes; ++g) {
    s.v[g] = std::log(0.05 + 0.2 * urand());
    s.qv[g] = std::exp(s.v[g]) * std::exp(0.2 * nrand());
  }
  s.pca.resize((size_t)n_cells * n_pcs);
  for (int c = 0; c < n_cells; ++c)
    for (int j = 0; j < n_pcs; ++j) s.pca[(size_t)j * n_cells + c] = (j == type[c] % n_pcs ? 3.0 : 0.0) + nrand();
  return s;
}

struct HostGraph {
  std::vector<int32_t> vertices, from, to, type;
  std::vector<double> weight;
};

#define CHECK(rc, ctx)                                                                       \
  do {                                                                                       \
    if ((rc) != 0) {                                                                         \
      std::fprintf(stderr, "FAILED at %s:%d: %s\n", __FILE__, __LINE__, cg_last_error(ctx)); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

static int fetch(cg_context* ctx, cg_graph* g, HostGraph& out) {
  int32_t nv = 0;
  int64_t ne = 0;
  cg_graph_sizes(g, &nv, &ne);
  out.vertices.resize(nv);
  out.from.resize(ne);
  out.to.resize(ne);
  out.type.resize(ne);
  out.weight.resize(ne);
  return cg_graph_fetch(ctx, g, out.vertices.data(), out.from.data(), out.to.data(), out.weight.data(), out.type.data());
}

static bool same(const HostGraph& a, const HostGraph& b) {
  return a.vertices == b.vertices && a.from == b.from && a.to == b.to && a.type == b.type &&
         a.weight.size() == b.weight.size() &&
         std::memcmp(a.weight.data(), b.weight.data(), a.weight.size() * sizeof(double)) == 0;
}

int main() {
  if (cg_device_count() < 1) {
    std::fprintf(stderr, "no CUDA device: libconosgpu has no CPU fallback\n");
    return 2;
  }
  const int S = 4, G = 320, n_pcs = 12;
  const int cells[S] = {900, 700, 1100, 650};
  std::vector<HostSample> hs;
  for (int s = 0; s < S; ++s) hs.push_back(make_sample(cells[s], G, 6, n_pcs, 0.02 * s));
  std::vector<cg_sample> cs(S);
  for (int s = 0; s < S; ++s)
    cs[s] = cg_sample{hs[s].n_cells, hs[s].n_genes, hs[s].p.data(), hs[s].i.data(), hs[s].x.data(), hs[s].v.data(),
                      hs[s].qv.data(), hs[s].pca.data(), n_pcs, hs[s].n_cells};
  std::vector<int32_t> ident(G);
  for (int g = 0; g < G; ++g) ident[g] = g;
  std::vector<cg_pair> pairs;
  for (int a = 0; a < S; ++a)
    for (int b = a + 1; b < S; ++b) {  // combn order (R/conclass.R:1058)
      cg_pair p;
      std::memset(&p, 0, sizeof(p));
      p.a = a;
      p.b = b;
      p.n_genes = G;
      p.genes_a = p.genes_b = ident.data();
      pairs.push_back(p);
    }
  cg_params P;
  cg_params_default(&P);
  P.k = P.k1 = 10;
  P.k_self = 5;
  P.ncomps = 16;

  // ---- (1) call by call on one context: what buildGraphGPU does on a single device
  HostGraph g_seq;
  {
    cg_context* ctx = nullptr;
    CHECK(cg_init(0, &ctx), nullptr);
    cg_panel* panel = nullptr;
    CHECK(cg_panel_create(ctx, cs.data(), S, &panel), ctx);
    cg_edges* edges = nullptr;
    std::vector<int64_t> counts(pairs.size());
    CHECK(cg_pairs_edges(ctx, panel, pairs.data(), (int)pairs.size(), &P, &edges, counts.data()), ctx);
    int64_t n_inter = 0;
    for (int64_t c : counts) n_inter += c;
    if (n_inter <= 0 || n_inter != cg_edges_count(edges)) {
      std::fprintf(stderr, "no inter-sample edges (%lld)\n", (long long)n_inter);
      return 1;
    }
    int32_t n_rows = 0, n_cols = 0;
    CHECK(cg_pair_reduction(ctx, 0, &n_rows, &n_cols, nullptr, nullptr), ctx);
    if (n_rows != G || n_cols != 16) {
      std::fprintf(stderr, "rotation of pair 0 is %d x %d\n", n_rows, n_cols);
      return 1;
    }
    std::vector<int32_t> sidx = {0, 1, 2, 3};
    CHECK(cg_local_edges(ctx, panel, sidx.data(), S, &P, &edges, nullptr), ctx);
    int32_t offs[S + 1];
    cg_panel_offsets(panel, offs);
    cg_graph* g = nullptr;
    CHECK(cg_assemble_graph(ctx, edges, offs[S], &g), ctx);
    CHECK(fetch(ctx, g, g_seq), ctx);
    std::printf("sequence: %zu vertices, %zu edges, %lld inter-sample rows\n", g_seq.vertices.size(), g_seq.from.size(),
                (long long)n_inter);
    {
      // the consumer next to the path (propagate_labels, src/propagate_labels.cpp:67): every 7th vertex labelled with
      // one of three labels, once on the device graph handle and once on the fetched edge list -- identical bits
      const int32_t nv = (int32_t)g_seq.vertices.size(), n_lab = 3;
      std::vector<int32_t> lab(nv, -1);
      for (int32_t v = 0; v < nv; v += 7) lab[v] = (v / 7) % n_lab;
      std::vector<double> da((size_t)nv * n_lab), db((size_t)nv * n_lab);
      int32_t ita = 0, itb = 0;
      double na = 0, nb = 0;
      CHECK(cg_propagate_labels(ctx, g, nullptr, nullptr, nullptr, 0, nv, lab.data(), n_lab, 20, 10.0, 0.1, 1e-3, 1,
                                da.data(), &ita, &na), ctx);
      CHECK(cg_propagate_labels(ctx, nullptr, g_seq.from.data(), g_seq.to.data(), g_seq.weight.data(),
                                (int64_t)g_seq.from.size(), nv, lab.data(), n_lab, 20, 10.0, 0.1, 1e-3, 1, db.data(), &itb,
                                &nb), ctx);
      if (ita != itb || std::memcmp(da.data(), db.data(), da.size() * sizeof(double)) != 0) {
        std::fprintf(stderr, "propagate_labels: graph handle and edge list disagree (%d vs %d iterations)\n", ita, itb);
        return 1;
      }
      for (int32_t v = 0; v < nv; v += 7)
        if (da[(size_t)lab[v] * nv + v] != 1.0) {
          std::fprintf(stderr, "a fixed label moved\n");
          return 1;
        }
      std::printf("propagate_labels: %d iterations, norm %.3g\n", ita, na);
    }
    cg_graph_free(g);
    cg_edges_free(edges);
    cg_panel_free(panel);
    // errors come back as status codes with a message, never as exceptions or aborts
    cg_pair bad = pairs[0];
    bad.a = 99;
    cg_edges* e2 = nullptr;
    if (cg_pairs_edges(ctx, nullptr, &bad, 1, &P, &e2, nullptr) == 0 || std::strlen(cg_last_error(ctx)) == 0) {
      std::fprintf(stderr, "bad arguments were accepted\n");
      return 1;
    }
    cg_shutdown(ctx);
  }
  // ---- (2) the same as ONE cg_job call, on 1 and (if present) 2 devices
  const int max_dev = cg_device_count() >= 2 ? 2 : 1;
  for (int nd = 1; nd <= max_dev; ++nd) {
    cg_job* job = nullptr;
    CHECK(cg_job_create(nd, &job), nullptr);
    cg_graph* g = nullptr;
    if (cg_job_build_graph(job, cs.data(), S, pairs.data(), (int)pairs.size(), &P, &g) != 0) {
      std::fprintf(stderr, "cg_job_build_graph(%d devices): %s\n", nd, cg_job_last_error(job));
      return 1;
    }
    HostGraph gj;
    CHECK(fetch(cg_job_context(job, 0), g, gj), cg_job_context(job, 0));
    cg_graph_free(g);
    for (int q = 0; q < (int)pairs.size(); ++q) {  // every reduction can be fetched from the device that computed it
      int32_t dev = -1, slot = -1, nr = 0, nc = 0;
      if (cg_job_pair_location(job, q, &dev, &slot) != 0 || dev < 0 || dev >= nd) return 1;
      CHECK(cg_pair_reduction(cg_job_context(job, dev), slot, &nr, &nc, nullptr, nullptr), cg_job_context(job, dev));
      if (nr != G || nc != 16) return 1;
    }
    cg_job_free(job);
    std::printf("job on %d device(s): %zu vertices, %zu edges -> %s\n", nd, gj.vertices.size(), gj.from.size(),
                same(gj, g_seq) ? "identical" : "DIFFERENT");
    if (!same(gj, g_seq)) return 1;
  }
  std::printf("JOB_TEST_OK\n");
  return 0;
}
