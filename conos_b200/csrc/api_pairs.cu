// C ABI, part 2: device-resident panel, per-pair reduction -> projection -> kNN -> matching, local edges.
#include <cmath>
#include <map>

#include <thread>

#include "api_core.cuh"
#include "hub.cuh"
#include "reduce.cuh"

using namespace cg;

namespace {
// a slice of the batch slab of cg_pairs_edges (same `.p` as a DevBuf, no ownership)
template <class T>
struct SlicePtr {
  T* p = nullptr;
};
}  // namespace

namespace {

// scaledMatricesP2 (R/conos.R:18-46): cqv = exp(rowMeans(log qv)), cgsf = sqrt(cqv / exp(v)), non-finite -> 0
void pair_scale_factors(const Sample& sa, const Sample& sb, const cg_pair& pr, bool var_scale, std::vector<double>& fa,
                        std::vector<double>& fb) {
  const int G = pr.n_genes;
  fa.assign(G, 1.0);
  fb.assign(G, 1.0);
  if (!var_scale) return;
  CG_CHECK(!sa.h_v.empty() && !sa.h_qv.empty() && !sb.h_v.empty() && !sb.h_qv.empty(),
           "var.scale = TRUE needs misc$varinfo$v and $qv for every sample");
  for (int g = 0; g < G; ++g) {
    const int ca = pr.genes_a[g], cb = pr.genes_b[g];
    const double cqv = std::exp((sa.h_logqv[ca] + sb.h_logqv[cb]) / 2.0);
    double a = std::sqrt(cqv / sa.h_expv[ca]);
    double b = std::sqrt(cqv / sb.h_expv[cb]);
    fa[g] = std::isfinite(a) ? a : 0.0;
    fb[g] = std::isfinite(b) ? b : 0.0;
  }
}


}  // namespace

extern "C" {

int cg_panel_create(cg_context* ctx, const cg_sample* samples, int n_samples, cg_panel** out) {
  cg_panel* pn = nullptr;
  try {
    CG_CHECK(ctx && samples && out && n_samples > 0, "bad argument");
    CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
    ::cg::PoolBinding pool_bind(ctx->c.panel_pool);
    pn = new cg_panel();
    pn->ctx = ctx;
    pn->stream = ctx->c.stream;
    pn->samples.resize(n_samples);
    pn->offsets.assign(n_samples + 1, 0);
    StageScope sc(ctx->c, "panel_upload");
    // all host->device copies go out first on their own stream; the conversions follow sample by sample
    struct CopyLane {
      cudaStream_t st = nullptr;
      std::vector<cudaEvent_t> ev;
      ~CopyLane() {
        for (auto e : ev) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
      }
    } lane;
    CG_CUDA(cudaStreamCreateWithFlags(&lane.st, cudaStreamNonBlocking));
    lane.ev.resize(n_samples, nullptr);
    for (int s = 0; s < n_samples; ++s) {
      const cg_sample& h = samples[s];
      // p == NULL: placeholder sample (its counts live on another rank); n_cells still numbers the cells
      CG_CHECK(h.p == nullptr || (h.i && h.x), "sample counts matrix incomplete");
      CG_CUDA(cudaEventCreateWithFlags(&lane.ev[s], cudaEventDisableTiming));
      sample_upload(ctx->c, pn->samples[s], h.n_cells, h.n_genes, h.p, h.i, h.x, h.var_v, h.var_qv, h.pca, h.n_pcs,
                    h.pca_ld > 0 ? h.pca_ld : h.n_cells, lane.st, lane.ev[s]);
      pn->samples[s].offset = pn->offsets[s];
      pn->offsets[s + 1] = pn->offsets[s] + h.n_cells;
    }
    if (ctx->c.profile) {  // profile mode serialises the two halves to time them separately
      const auto t0 = std::chrono::steady_clock::now();
      CG_CUDA(cudaStreamSynchronize(lane.st));
      ctx->c.stage_ms["panel_upload.copies"] +=
          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    {
      StageScope sc2(ctx->c, "panel_upload.convert");
      for (int s = 0; s < n_samples; ++s) sample_finish(ctx->c, pn->samples[s], lane.ev[s]);
    }
    CG_CUDA(cudaStreamSynchronize(lane.st));
    for (const auto& s : pn->samples) {
      if (s.has_counts) pn->device_bytes += (size_t)s.nnz * 24 + ((size_t)s.n_cells + s.n_genes) * 24;
      pn->device_bytes += (size_t)s.n_cells * s.n_pcs * 12;
    }
    ctx->c.free_bytes_seen -= std::min(ctx->c.free_bytes_seen, pn->device_bytes);
    *out = pn;
    return 0;
  } catch (const std::exception& e) {
    delete pn;
    return fail(ctx, e);
  }
}

void cg_panel_free(cg_panel* panel) {
  if (!panel) return;
  // the blocks go back to the panel pool, which keeps them: they are free again as far as this context is concerned
  if (::cg::stream_alive(panel->stream) && panel->ctx) panel->ctx->c.free_bytes_seen += panel->device_bytes;
  delete panel;
}

int cg_panel_offsets(const cg_panel* panel, int32_t* offsets) {
  if (!panel || !offsets) return 1;
  for (size_t s = 0; s < panel->offsets.size(); ++s) offsets[s] = panel->offsets[s];
  return 0;
}

int cg_panel_drop_cache(cg_panel* panel) {
  if (!panel) return 1;
  ::cg::StreamBinding bind(::cg::stream_alive(panel->stream) ? panel->stream : nullptr);
  for (auto& s : panel->samples) {
    s.gram.release();
    s.h_gram_diag.clear();
    s.pca_tiled.release();
    s.pca_metric = -1;
  }
  return 0;
}

int cg_keep_projections(cg_context* ctx, int enable) {
  if (!ctx) return 1;
  ctx->keep_projections = enable != 0;
  return 0;
}

int cg_pairs_edges(cg_context* ctx, cg_panel* panel, const cg_pair* pairs, int n_pairs, const cg_params* params,
                   cg_edges** edges, int64_t* pair_edge_counts) {
  CG_API_BEGIN
  CG_CHECK(ctx && panel && params && edges && (pairs || n_pairs == 0), "NULL argument");
  const cg_params& P = *params;
  CG_CHECK(P.space == CG_SPACE_PCA || P.space == CG_SPACE_CPCA || P.space == CG_SPACE_CCA,
           "only the following spaces are currently supported: [CPCA PCA CCA]");
  CG_CHECK(P.matching == CG_MATCH_MNN || P.matching == CG_MATCH_NN,
           "only the following matching methods are currently supported: ['mNN' 'NN']");
  CG_CHECK(P.metric == CG_METRIC_ANGULAR || P.metric == CG_METRIC_L2,
           "only the following distance metrics are currently supported: ['L2' 'angular']");
  CG_CHECK(P.k1 >= P.k && P.k >= 1, "k1 must be >= k");
  // the common space: up to 64 components on the tensor cores, 65..112 through the exact SIMT kNN kernel (the device
  // eigen solvers deliver at most 56 vectors per sample: PCA r = round(ncomps / 2) <= 56, CPCA / CCA ncomps <= 56)
  CG_CHECK(P.ncomps >= 1 && P.ncomps <= 112, "ncomps must be in 1..112");
  Context& c = ctx->c;
  CG_CUDA(cudaSetDevice(c.device));
  CG_BIND(ctx);
  if (!*edges) *edges = new cg_edges();
  EdgeTable& out = (*edges)->t;
  ctx->kept.clear();
  ctx->c.rot_arena.reset();
  ctx->kept.resize(n_pairs);
  const int S = (int)panel->samples.size();
  const Metric metric = (Metric)P.metric;
  SimParams sp{P.metric, P.matching, P.cor_base, P.l2_sigma, P.min_similarity};

  // ---- batches of pairs bounded by device memory
  // per-pair footprint (bytes); pairs with a bad sample index are rejected below
  auto need_of = [&](const cg_pair& pr) -> size_t {
    if (pr.a < 0 || pr.a >= S || pr.b < 0 || pr.b >= S) return 0;
    const Sample& sa = panel->samples[pr.a];
    const Sample& sb = panel->samples[pr.b];
    const size_t na = sa.n_cells, nb = sb.n_cells;
    size_t need = (round_up(na, KNN_ROW_PAD) + round_up(nb, KNN_ROW_PAD)) * 64 * 4  // panels
                  + (na + nb) * (size_t)P.k1 * 8 * 2                                   // kNN out + matching temps
                  + ((size_t)sa.n_genes + sb.n_genes) * 64 * 8                         // W
                  + (size_t)(pr.n_genes > 0 ? pr.n_genes : 0) * 64 * 8 * 2;
    if (ctx->keep_projections) need += (na + nb) * 64 * 8;
    return need;
  };
  size_t need_all = 0;
  for (int q = 0; q < n_pairs; ++q) need_all += need_of(pairs[q]);
  int p0 = 0;
  while (p0 < n_pairs) {
    // free memory: the context's estimate when the whole call fits comfortably, a fresh look otherwise (the query is
    // the one runtime call of this path that stalls for 1-100 ms on some hosts)
    const size_t budget = std::min<size_t>(c.free_bytes_for(need_all) / 3, c.max_batch_bytes);
    size_t used = 0;
    int p1 = p0;
    // a batch shares k.cur (R/conclass.R:230-233) and the width of the common space: it ends where either changes
    auto width_of = [&](const cg_pair& pr) {
      if (P.space == CG_SPACE_CCA) return pr.u && pr.v ? pr.uv_cols : std::min(P.ncomps, pr.n_genes - 1);
      if (pr.rot && pr.rot_cols > 0) return std::min(P.ncomps, pr.rot_cols);
      if (P.space == CG_SPACE_CPCA) return std::min(P.ncomps, pr.n_genes - 1);
      return std::min(P.ncomps, 2 * (int)std::nearbyint(P.ncomps / 2.0));
    };
    auto k_of = [&](const cg_pair& pr) { return pr.k > 0 ? std::min(pr.k, P.k1) : P.k; };
    const int batch_k = k_of(pairs[p0]), batch_d = width_of(pairs[p0]);
    while (p1 < n_pairs) {
      const cg_pair& pr = pairs[p1];
      if (p1 > p0 && (k_of(pr) != batch_k || width_of(pr) != batch_d)) break;
      CG_CHECK(pr.a >= 0 && pr.a < S && pr.b >= 0 && pr.b < S && pr.a != pr.b, "pair refers to a missing sample");
      const Sample& sa = panel->samples[pr.a];
      const Sample& sb = panel->samples[pr.b];
      CG_CHECK(sa.has_counts && sb.has_counts,
               "pair refers to a sample whose counts are not resident on this device (placeholder sample)");
      if (P.space != CG_SPACE_CCA || !(pr.u && pr.v)) {
        CG_CHECK(pr.n_genes >= 0 && (pr.n_genes == 0 || (pr.genes_a && pr.genes_b)), "pair gene columns missing");
        for (int g = 0; g < pr.n_genes; ++g)
          CG_CHECK(pr.genes_a[g] >= 0 && pr.genes_a[g] < sa.n_genes && pr.genes_b[g] >= 0 &&
                       pr.genes_b[g] < sb.n_genes,
                   "pair gene column out of range");
      }
      const size_t need = need_of(pr);
      if (p1 > p0 && used + need > budget) break;
      used += need;
      ++p1;
    }
    const int nb_pairs = p1 - p0;

    // ---- reductions (rotation per pair, column-major G x d_eff doubles on the device)
    std::vector<PairReduction> red(nb_pairs);
    if (P.space != CG_SPACE_CCA) {
      // scaledMatricesP2 factors of every pair of the batch, once, on several host threads (the reduction and the
      // projection operands both need them; at the atlas one thread spends 0.3 s per use on exp / sqrt)
      if (P.var_scale)
        for (int q = 0; q < nb_pairs; ++q) {
          const Sample& sa = panel->samples[pairs[p0 + q].a];
          const Sample& sb = panel->samples[pairs[p0 + q].b];
          CG_CHECK(!sa.h_v.empty() && !sa.h_qv.empty() && !sb.h_v.empty() && !sb.h_qv.empty(),
                   "var.scale = TRUE needs misc$varinfo$v and $qv for every sample");
        }
      auto factors = [&](int q0, int q1) {
        for (int q = q0; q < q1; ++q) {
          const cg_pair& pr = pairs[p0 + q];
          pair_scale_factors(panel->samples[pr.a], panel->samples[pr.b], pr, P.var_scale != 0, red[q].h_fa, red[q].h_fb);
        }
      };
      const int nt = host_fill_threads(c, (size_t)nb_pairs, 64);
      if (nt == 1) {
        factors(0, nb_pairs);
      } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(factors, nb_pairs * t / nt, nb_pairs * (t + 1) / nt);
        for (auto& t : th) t.join();
      }
    }
    {
      StageScope sc(c, "reduction");
      compute_pair_reductions(c, *panel, pairs + p0, nb_pairs, P, red);
    }
    StageScope* sc_setup = new StageScope(c, "pair_setup");

    // ---- projection operands
    // The per-pair device arrays of a batch (projected panels, projection operands, kNN output) are carved out of
    // ONE allocation: at the atlas nine stream-ordered allocations and nine frees per pair (4950 pairs) cost more host
    // time than everything else in this stage.  256-byte aligned slices; W slices are contiguous (one memset).
    struct PairBuf {
      SlicePtr<double> Wa, Wb, shift;
      DevBuf<double> p64a, p64b;        // only with cg_keep_projections
      DevBuf<int> rows_u, rows_v;       // only CCA
      SlicePtr<float> ta, tb;
      SlicePtr<int> i_ab, i_ba;
      SlicePtr<float> d_ab, d_ba;
      int na = 0, nb = 0, d = 0, d_pad = 32;
    };
    std::vector<PairBuf> pb(nb_pairs);
    auto al = [](size_t bytes) { return (bytes + 255) & ~(size_t)255; };
    size_t slab_w = 0, slab_rest = 0;
    for (int q = 0; q < nb_pairs; ++q) {
      const cg_pair& pr = pairs[p0 + q];
      const PairReduction& R = red[q];
      const size_t na = P.space == CG_SPACE_CCA ? (size_t)R.n_u : (size_t)panel->samples[pr.a].n_cells;
      const size_t nb = P.space == CG_SPACE_CCA ? (size_t)R.n_v : (size_t)panel->samples[pr.b].n_cells;
      const int dp = knn_d_pad(R.d);
      slab_rest += al(KnnPanel::elems((int)na, R.d) * sizeof(float)) + al(KnnPanel::elems((int)nb, R.d) * sizeof(float));
      slab_rest += 2 * al(na * P.k1 * sizeof(int)) + 2 * al(nb * P.k1 * sizeof(int));   // idx + dist (4 bytes each)
      if (P.space != CG_SPACE_CCA) {
        slab_w += al((size_t)panel->samples[pr.a].n_genes * dp * sizeof(double)) +
                  al((size_t)panel->samples[pr.b].n_genes * dp * sizeof(double));
        slab_rest += al((size_t)2 * dp * sizeof(double));
      }
    }
    DevBuf<unsigned char> slab(slab_w + slab_rest);
    size_t off_w = 0, off_rest = slab_w;
    auto take_w = [&](size_t bytes) {
      unsigned char* r = slab.p + off_w;
      off_w += al(bytes);
      return r;
    };
    auto take = [&](size_t bytes) {
      unsigned char* r = slab.p + off_rest;
      off_rest += al(bytes);
      return r;
    };
    // the small per-pair vectors (scale factors, centering, gene columns) of the whole batch travel in two uploads
    size_t stage_d = 0, stage_i = 0;
    if (P.space != CG_SPACE_CCA)
      for (int q = 0; q < nb_pairs; ++q) {
        stage_d += 4 * (size_t)pairs[p0 + q].n_genes;
        stage_i += 2 * (size_t)pairs[p0 + q].n_genes;
      }
    // staged in page-locked memory (the context's arena: valid until the next cg_pairs_edges call) and filled by
    // several host threads after the loop below has handed out the slices
    double* h_stage_d = stage_d ? static_cast<double*>(c.rot_arena.take(stage_d * sizeof(double))) : nullptr;
    int* h_stage_i = stage_i ? static_cast<int*>(c.rot_arena.take(stage_i * sizeof(int))) : nullptr;
    size_t used_d = 0, used_i = 0;
    std::vector<size_t> off_d(nb_pairs, 0), off_i(nb_pairs, 0);
    DevBuf<double> d_stage_d(stage_d);
    DevBuf<int> d_stage_i(stage_i);
    std::vector<WBuildProblem> wprobs;
    std::vector<ProjProblem> pprobs;
    std::vector<int> psample, pgenes;  // sample index and gene count of every projection problem
    std::vector<int> pproj_rot_supplied;  // per pair with projection problems: 1 when the caller supplied the rotation
    std::vector<KnnProblem> kprobs;
    std::vector<MnnProblem> mprobs;
    int d_pad_all = 0, d_all = -1;
    for (int q = 0; q < nb_pairs; ++q) {
      const cg_pair& pr = pairs[p0 + q];
      const Sample& sa = panel->samples[pr.a];
      const Sample& sb = panel->samples[pr.b];
      PairBuf& B = pb[q];
      PairReduction& R = red[q];
      cg_context::PairKeep& keep = ctx->kept[p0 + q];
      if (P.space == CG_SPACE_CCA) {
        // u, v are the projections themselves (R/conclass.R:268-270)
        B.na = R.n_u;
        B.nb = R.n_v;
        B.d = R.d;
      } else {
        B.na = sa.n_cells;
        B.nb = sb.n_cells;
        B.d = R.d;  // min(ncomps, ncol(rot)) -- R/conclass.R:254-258
      }
      B.d_pad = knn_d_pad(B.d);
      CG_CHECK(d_all < 0 || d_all == B.d, "all pairs of one call must share the number of components");
      d_all = B.d;
      d_pad_all = B.d_pad;
      B.ta.p = reinterpret_cast<float*>(take(KnnPanel::elems(B.na, B.d) * sizeof(float)));
      B.tb.p = reinterpret_cast<float*>(take(KnnPanel::elems(B.nb, B.d) * sizeof(float)));
      keep.n_rows = R.n_rows;
      keep.n_cols = R.n_cols;
      keep.rot = std::move(R.h_rot);
      keep.rot_host = R.h_rot_pinned ? R.h_rot_pinned : keep.rot.data();
      keep.nv = R.h_nv;
      keep.D = R.h_D;
      keep.niter = R.h_niter;
      keep.ul = std::move(R.h_ul);
      keep.vl = std::move(R.h_vl);
      if (P.space == CG_SPACE_CCA) {
        KnnPanel pa, pbn;
        pa.tiled = B.ta.p;
        pbn.tiled = B.tb.p;
        knn_prep_from_f64_colmajor(c, R.u.p, B.na, B.d, B.na, metric, pa);
        knn_prep_from_f64_colmajor(c, R.v.p, B.nb, B.d, B.nb, metric, pbn);
        B.rows_u.upload(R.h_rows_u.data(), R.h_rows_u.size(), c.stream);
        B.rows_v.upload(R.h_rows_v.data(), R.h_rows_v.size(), c.stream);
        if (!pr.u) {  // computed on the device: keep host copies for the self$pairs$CCA cache
          keep.cca_d = B.d;
          keep.u.resize((size_t)B.na * B.d);
          keep.v.resize((size_t)B.nb * B.d);
          R.u.download(keep.u.data(), keep.u.size(), c.stream);
          R.v.download(keep.v.data(), keep.v.size(), c.stream);
          keep.rows_u = R.h_rows_u;
          keep.rows_v = R.h_rows_v;
          keep.sigma = R.h_sigma;
        }
        if (ctx->keep_projections) {
          keep.pn[0] = B.na;
          keep.pn[1] = B.nb;
          keep.proj[0].resize((size_t)B.na * B.d);
          keep.proj[1].resize((size_t)B.nb * B.d);
          R.u.download(keep.proj[0].data(), keep.proj[0].size(), c.stream);
          R.v.download(keep.proj[1].data(), keep.proj[1].size(), c.stream);
        }
      } else {
        const int G = pr.n_genes;
        if (P.common_centering) CG_CHECK(G >= 2, "common centering needs at least two od genes");
        // slices of the staging arrays: [fa | fb | ca | cb] and [genes_a | genes_b]; filled after the loop
        off_d[q] = used_d;
        off_i[q] = used_i;
        const double* dfa = d_stage_d.p + used_d;
        const double* dfb = dfa + G;
        const double* dca = dfb + G;
        const double* dcb = dca + G;
        const int* dga = d_stage_i.p + used_i;
        const int* dgb = dga + G;
        used_d += 4 * (size_t)G;
        used_i += 2 * (size_t)G;
        B.Wa.p = reinterpret_cast<double*>(take_w((size_t)sa.n_genes * B.d_pad * sizeof(double)));
        B.Wb.p = reinterpret_cast<double*>(take_w((size_t)sb.n_genes * B.d_pad * sizeof(double)));
        B.shift.p = reinterpret_cast<double*>(take((size_t)2 * B.d_pad * sizeof(double)));
        wprobs.push_back(WBuildProblem{R.rot.p, dfa, dca, dga, B.Wa.p, B.shift.p, G, sa.n_genes});
        wprobs.push_back(WBuildProblem{R.rot.p, dfb, dcb, dgb, B.Wb.p, B.shift.p + B.d_pad, G, sb.n_genes});
        if (ctx->keep_projections) {
          B.p64a.alloc((size_t)B.na * B.d);
          B.p64b.alloc((size_t)B.nb * B.d);
        }
        pprobs.push_back(ProjProblem{sa.csr_p.p, sa.csr_i.p, sa.csr_x.p, B.Wa.p, B.shift.p, B.ta.p, B.p64a.p, B.na});
        pprobs.push_back(
            ProjProblem{sb.csr_p.p, sb.csr_i.p, sb.csr_x.p, B.Wb.p, B.shift.p + B.d_pad, B.tb.p, B.p64b.p, B.nb});
        pproj_rot_supplied.push_back(pr.rot != nullptr ? 1 : 0);
        psample.push_back(pr.a);
        psample.push_back(pr.b);
        pgenes.push_back(sa.n_genes);
        pgenes.push_back(sb.n_genes);
      }
      B.i_ab.p = reinterpret_cast<int*>(take((size_t)B.na * P.k1 * sizeof(int)));
      B.d_ab.p = reinterpret_cast<float*>(take((size_t)B.na * P.k1 * sizeof(float)));
      B.i_ba.p = reinterpret_cast<int*>(take((size_t)B.nb * P.k1 * sizeof(int)));
      B.d_ba.p = reinterpret_cast<float*>(take((size_t)B.nb * P.k1 * sizeof(float)));
      kprobs.push_back(KnnProblem{B.ta.p, B.tb.p, B.na, B.nb, B.i_ab.p, B.d_ab.p});
      kprobs.push_back(KnnProblem{B.tb.p, B.ta.p, B.nb, B.na, B.i_ba.p, B.d_ba.p});
      mprobs.push_back(MnnProblem{B.i_ab.p, B.d_ab.p, B.i_ba.p, B.d_ba.p,
                                  P.space == CG_SPACE_CCA ? B.rows_u.p : nullptr,
                                  P.space == CG_SPACE_CCA ? B.rows_v.p : nullptr, B.na, B.nb, sa.offset, sb.offset});
    }
    if (used_d) {
      StageScope sc_up(c, "pair_setup.upload");
      // scale factors, centering and gene columns of every pair into its slices
      auto fill = [&](int q0, int q1) {
        for (int q = q0; q < q1; ++q) {
          const cg_pair& pr = pairs[p0 + q];
          const Sample& sa = panel->samples[pr.a];
          const Sample& sb = panel->samples[pr.b];
          const int G = pr.n_genes;
          const std::vector<double>& fa = red[q].h_fa;   // computed for the whole batch above
          const std::vector<double>& fb = red[q].h_fb;
          double* o = h_stage_d + off_d[q];
          double *ofa = o, *ofb = o + G, *oca = o + 2 * (size_t)G, *ocb = o + 3 * (size_t)G;
          // centering (R/conos.R:784-789): per-gene scaled means, and the quirk of the default path:
          // `centering` is an atomic vector, so centering[[n]] is ONE number per sample (m[1], m[2])
          for (int g = 0; g < G; ++g) {
            ofa[g] = fa[g];
            ofb[g] = fb[g];
            oca[g] = fa[g] * sa.h_colsum[pr.genes_a[g]] / (double)sa.n_cells;
            ocb[g] = fb[g] * sb.h_colsum[pr.genes_b[g]] / (double)sb.n_cells;
          }
          if (P.common_centering) {
            const double na = sa.n_cells, nb = sb.n_cells;
            const double m0 = (oca[0] * na + ocb[0] * nb) / (na + nb);
            const double m1 = (oca[1] * na + ocb[1] * nb) / (na + nb);
            std::fill(oca, oca + G, m0);
            std::fill(ocb, ocb + G, m1);
          }
          memcpy(h_stage_i + off_i[q], pr.genes_a, (size_t)G * sizeof(int));
          memcpy(h_stage_i + off_i[q] + G, pr.genes_b, (size_t)G * sizeof(int));
        }
      };
      const int nt = host_fill_threads(c, (size_t)nb_pairs, 64);
      if (nt == 1) {
        fill(0, nb_pairs);
      } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(fill, nb_pairs * t / nt, nb_pairs * (t + 1) / nt);
        for (auto& t : th) t.join();
      }
      CG_CUDA(cudaMemcpyAsync(d_stage_d.p, h_stage_d, used_d * sizeof(double), cudaMemcpyHostToDevice, c.stream));
      CG_CUDA(cudaMemcpyAsync(d_stage_i.p, h_stage_i, used_i * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    }
    delete sc_setup;
    if (!wprobs.empty()) {
      StageScope sc(c, "projection");
      CG_CHECK(off_w <= slab_w && off_rest <= slab_w + slab_rest, "internal: batch slab overrun");
      if (off_w) CG_CUDA(cudaMemsetAsync(slab.p, 0, off_w, c.stream));   // all W operands at once
      build_projection_operands(c, wprobs.data(), (int)wprobs.size(), d_all, d_pad_all, /*zeroed=*/true);
      // default mode: one tensor-core GEMM per sample for all of its pair sides (fp32 accumulation, the accuracy class
      // of the tensor-core rotations it multiplies).  The FP64 SpMM serves the exact mode, rotations the caller
      // supplied (exact inputs: the projection stays exact too), d > 32, gene universes beyond the dense operand and
      // callers that want the raw FP64 projections back
      // The engine is chosen pair by pair, never for the batch as a whole (a pair's edges must not depend on which
      // other pairs share its call).
      const bool tc_mode = c.tc_projection && !c.exact_reduction && !ctx->keep_projections && d_pad_all == 32;
      std::vector<ProjProblem> tc_probs, fp_probs;
      std::vector<int> tc_sample, tc_genes;
      for (size_t t = 0; t + 1 < pprobs.size(); t += 2) {
        const bool tc = tc_mode && pgenes[t] <= 4096 && pgenes[t + 1] <= 4096 && pproj_rot_supplied[t / 2] == 0;
        for (int side = 0; side < 2; ++side) {
          if (tc) {
            tc_probs.push_back(pprobs[t + side]);
            tc_sample.push_back(psample[t + side]);
            tc_genes.push_back(pgenes[t + side]);
          } else {
            fp_probs.push_back(pprobs[t + side]);
          }
        }
      }
      if (!tc_probs.empty())
        project_and_prep_tc(c, tc_probs.data(), tc_sample.data(), tc_genes.data(), (int)tc_probs.size(), d_all, metric);
      if (!fp_probs.empty()) project_and_prep(c, fp_probs.data(), (int)fp_probs.size(), d_all, d_pad_all, metric);
      if (ctx->keep_projections) {
        for (int q = 0; q < nb_pairs; ++q) {
          cg_context::PairKeep& keep = ctx->kept[p0 + q];
          PairBuf& B = pb[q];
          keep.pn[0] = B.na;
          keep.pn[1] = B.nb;
          keep.proj[0].resize((size_t)B.na * B.d);
          keep.proj[1].resize((size_t)B.nb * B.d);
          B.p64a.download(keep.proj[0].data(), keep.proj[0].size(), c.stream);
          B.p64b.download(keep.proj[1].data(), keep.proj[1].size(), c.stream);
        }
        CG_CUDA(cudaStreamSynchronize(c.stream));
      }
    }
    CG_CUDA(cudaStreamSynchronize(c.stream));
    // ---- exact kNN, both directions of every pair in one launch
    {
      StageScope sc(c, "cross_knn");
      DevBuf<KnnProblem> dk;
      dk.upload(kprobs.data(), kprobs.size(), c.stream);
      knn_launch(c, dk.p, kprobs.data(), (int)kprobs.size(), P.k1, d_pad_all, d_all, metric);
      CG_CUDA(cudaStreamSynchronize(c.stream));
    }
    // ---- matching -> edges
    std::vector<long long> counts(nb_pairs, 0);
    {
      StageScope sc(c, "matching");
      match_append_edges(c, mprobs.data(), nb_pairs, batch_k, P.k1, sp, out, counts.data());
    }
    StageScope sc_free(c, "pair_free");
    if (pair_edge_counts)
      for (int q = 0; q < nb_pairs; ++q) pair_edge_counts[p0 + q] = counts[q];
    p0 = p1;
  }
  CG_API_END(ctx)
}

int cg_pair_reduction(cg_context* ctx, int which, int32_t* n_rows, int32_t* n_cols, double* rot, double* nv) {
  CG_API_BEGIN
  CG_CHECK(ctx && which >= 0 && which < (int)ctx->kept.size(), "no such pair in the last cg_pairs_edges call");
  const auto& k = ctx->kept[which];
  if (n_rows) *n_rows = k.n_rows;
  if (n_cols) *n_cols = k.n_cols;
  if (rot && k.rot_host) memcpy(rot, k.rot_host, (size_t)k.n_rows * k.n_cols * sizeof(double));
  if (nv && !k.nv.empty()) memcpy(nv, k.nv.data(), k.nv.size() * sizeof(double));
  CG_API_END(ctx)
}

int cg_pair_cca(cg_context* ctx, int which, int32_t* n_u, int32_t* n_v, int32_t* d, double* u, double* v,
                int32_t* rows_u, int32_t* rows_v, double* sigma) {
  CG_API_BEGIN
  CG_CHECK(ctx && which >= 0 && which < (int)ctx->kept.size(), "no such pair in the last cg_pairs_edges call");
  const auto& k = ctx->kept[which];
  CG_CUDA(cudaStreamSynchronize(ctx->c.stream));
  if (n_u) *n_u = (int32_t)k.rows_u.size();
  if (n_v) *n_v = (int32_t)k.rows_v.size();
  if (d) *d = k.cca_d;
  if (u && !k.u.empty()) memcpy(u, k.u.data(), k.u.size() * sizeof(double));
  if (v && !k.v.empty()) memcpy(v, k.v.data(), k.v.size() * sizeof(double));
  if (rows_u && !k.rows_u.empty()) memcpy(rows_u, k.rows_u.data(), k.rows_u.size() * sizeof(int));
  if (rows_v && !k.rows_v.empty()) memcpy(rows_v, k.rows_v.data(), k.rows_v.size() * sizeof(int));
  if (sigma && !k.sigma.empty()) memcpy(sigma, k.sigma.data(), k.sigma.size() * sizeof(double));
  CG_API_END(ctx)
}

// All rotations of the last call in one go (an atlas run caches thousands of pairs: one call instead of two per pair)
int cg_pair_reductions_all(cg_context* ctx, int32_t* n_rows, int32_t* n_cols, int64_t* offsets, double* rot,
                           int64_t rot_capacity) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  CG_CUDA(cudaStreamSynchronize(ctx->c.stream));  // the asynchronous copies into the pinned arena have landed
  int64_t off = 0;
  std::vector<int64_t> offs(ctx->kept.size() + 1, 0);
  for (size_t q = 0; q < ctx->kept.size(); ++q) {
    const auto& k = ctx->kept[q];
    const int64_t n = (int64_t)k.n_rows * k.n_cols;
    if (n_rows) n_rows[q] = k.n_rows;
    if (n_cols) n_cols[q] = k.n_cols;
    if (offsets) offsets[q] = off;
    offs[q] = off;
    if (rot && k.rot_host && n > 0) CG_CHECK(off + n <= rot_capacity, "rotation buffer too small");
    off += n;
  }
  offs[ctx->kept.size()] = off;
  if (offsets) offsets[ctx->kept.size()] = off;
  if (rot && off > 0) {
    // the atlas fetches 4950 rotations = 2.4 GB into freshly allocated (not yet touched) host memory: the copy is
    // bound by page faults of one thread (0.5 s); eight threads take contiguous ranges of the pair list
    const size_t nq = ctx->kept.size();
    const int nt = host_fill_threads(ctx->c, (size_t)off * sizeof(double), (size_t)64 << 20);
    auto copy_range = [&](size_t q0, size_t q1) {
      for (size_t q = q0; q < q1; ++q) {
        const auto& k = ctx->kept[q];
        const int64_t n = (int64_t)k.n_rows * k.n_cols;
        if (k.rot_host && n > 0) memcpy(rot + offs[q], k.rot_host, (size_t)n * sizeof(double));
      }
    };
    if (nt == 1) {
      copy_range(0, nq);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back(copy_range, nq * t / nt, nq * (t + 1) / nt);
      for (auto& t : th) t.join();
    }
  }
  CG_API_END(ctx)
}

int cg_pair_cpca(cg_context* ctx, int which, int32_t* n_cols, double* D, double* niter) {
  CG_API_BEGIN
  CG_CHECK(ctx && which >= 0 && which < (int)ctx->kept.size(), "no such pair in the last cg_pairs_edges call");
  const auto& k = ctx->kept[which];
  if (n_cols) *n_cols = (int32_t)k.niter.size();
  if (D && !k.D.empty()) memcpy(D, k.D.data(), k.D.size() * sizeof(double));
  if (niter && !k.niter.empty()) memcpy(niter, k.niter.data(), k.niter.size() * sizeof(double));
  CG_API_END(ctx)
}

int cg_pair_cca_loadings(cg_context* ctx, int which, int32_t* n_genes, int32_t* d, double* ul, double* vl, double* nv) {
  CG_API_BEGIN
  CG_CHECK(ctx && which >= 0 && which < (int)ctx->kept.size(), "no such pair in the last cg_pairs_edges call");
  const auto& k = ctx->kept[which];
  if (n_genes) *n_genes = k.ul.empty() ? 0 : k.n_rows;
  if (d) *d = k.cca_d;
  if (ul && !k.ul.empty()) memcpy(ul, k.ul.data(), k.ul.size() * sizeof(double));
  if (vl && !k.vl.empty()) memcpy(vl, k.vl.data(), k.vl.size() * sizeof(double));
  if (nv && !k.nv.empty()) memcpy(nv, k.nv.data(), k.nv.size() * sizeof(double));
  CG_API_END(ctx)
}

int cg_pair_projection(cg_context* ctx, int which, int side, int32_t* n_rows, int32_t* n_cols, double* p) {
  CG_API_BEGIN
  CG_CHECK(ctx && which >= 0 && which < (int)ctx->kept.size() && (side == 0 || side == 1), "bad argument");
  const auto& k = ctx->kept[which];
  CG_CHECK(!k.proj[side].empty(), "projections were not kept (cg_keep_projections)");
  if (n_rows) *n_rows = k.pn[side];
  if (n_cols) *n_cols = k.pn[side] ? (int)(k.proj[side].size() / (size_t)k.pn[side]) : 0;
  if (p) memcpy(p, k.proj[side].data(), k.proj[side].size() * sizeof(double));
  CG_API_END(ctx)
}

int cg_local_edges(cg_context* ctx, cg_panel* panel, const int32_t* samples, int n, const cg_params* params,
                   cg_edges** edges, int64_t* sample_edge_counts) {
  CG_API_BEGIN
  CG_CHECK(ctx && panel && params && edges && (samples || n == 0), "NULL argument");
  const cg_params& P = *params;
  CG_CHECK(P.metric == CG_METRIC_ANGULAR || P.metric == CG_METRIC_L2,
           "only the following distance metrics are currently supported: ['L2' 'angular']");
  Context& c = ctx->c;
  CG_CUDA(cudaSetDevice(c.device));
  CG_BIND(ctx);
  if (!*edges) *edges = new cg_edges();
  if (sample_edge_counts)
    for (int t = 0; t < n; ++t) sample_edge_counts[t] = 0;
  if (P.k_self <= 0) return 0;
  const int kp1 = P.k_self + 1;  // +1 accounts for the self edge removed afterwards (R/conos.R:595)
  // self-kNN of every listed sample: one launch per distinct number of PCA columns (each sample brings its own
  // reduction, R/conos.R:595 -- the bundled panel has 24 and 33 columns)
  std::vector<DevBuf<int>> idx(n);
  std::vector<DevBuf<float>> dist(n);
  std::map<int, std::vector<KnnProblem>> by_d;
  for (int t = 0; t < n; ++t) {
    CG_CHECK(samples[t] >= 0 && samples[t] < (int)panel->samples.size(), "no such sample");
    Sample& s = panel->samples[samples[t]];
    CG_CHECK(s.n_pcs > 0, "PCA must be estimated for all samples");
    // pagoda2's default nPcs is 100 (R/access_wrappers.R:9 hands over whatever the sample carries): up to 64 columns run
    // on the tensor cores, 65..128 through the exact SIMT kernel
    CG_CHECK(s.n_pcs <= KNN_MAX_D, "at most 128 PCA columns per sample are supported for the local neighbours");
    if (s.pca_metric != P.metric) {
      s.pca_tiled.alloc(KnnPanel::elems(s.n_cells, s.n_pcs));
      KnnPanel pn;
      pn.tiled = s.pca_tiled.p;
      knn_prep_from_f64_colmajor(c, s.pca.p, s.n_cells, s.n_pcs, s.n_cells, (Metric)P.metric, pn);
      s.pca_metric = P.metric;
    }
    idx[t].alloc((size_t)s.n_cells * kp1);
    dist[t].alloc((size_t)s.n_cells * kp1);
    by_d[s.n_pcs].push_back(KnnProblem{s.pca_tiled.p, s.pca_tiled.p, s.n_cells, s.n_cells, idx[t].p, dist[t].p});
  }
  {
    StageScope sc(c, "self_knn");
    for (auto& kv : by_d) {
      DevBuf<KnnProblem> dk;
      dk.upload(kv.second.data(), kv.second.size(), c.stream);
      knn_launch(c, dk.p, kv.second.data(), (int)kv.second.size(), kp1, knn_d_pad(kv.first), kv.first,
                 (Metric)P.metric);
      CG_CUDA(cudaStreamSynchronize(c.stream));
    }
  }
  StageScope sc_le(c, "local_edges");
  std::vector<LocalProblem> lp(n);
  for (int t = 0; t < n; ++t) {
    const Sample& s = panel->samples[samples[t]];
    lp[t] = LocalProblem{idx[t].p, dist[t].p, s.n_cells, s.offset};
  }
  std::vector<long long> lcounts(n, 0);
  local_append_edges_batch(c, lp.data(), n, kp1, P.metric, P.l2_sigma, P.k_self_weight, (*edges)->t, lcounts.data());
  if (sample_edge_counts)
    for (int t = 0; t < n; ++t) sample_edge_counts[t] = (int64_t)lcounts[t];
  // idx / dist go out of scope here and are released in stream order
  CG_API_END(ctx)
}

}  // extern "C"
