// cg_job: one buildGraph over several GPUs of a node driven from ONE host process -- the form an R session needs
// (R cannot fork with a CUDA context, and it has no torchrun): one host thread and one context per device, an NCCL
// communicator between them, and the same sequence every rank of a one-process-per-GPU run executes
// (partition -> resident samples -> shared Gram matrices -> pairs -> allgatherv -> local edges -> allgatherv ->
// assembly on device 0).  Replaces the fork pools of the reference (R/conclass.R:224-317 papply over the pairs,
// :1074-1088 inside updatePairs, R/conos.R:589 over the samples) and the rbind / graph_from_edgelist / simplify that
// follow (R/conclass.R:316-343).  With n_devices == 1 nothing but the library itself is needed (no NCCL).
#include <algorithm>
#include <mutex>
#include <thread>

#include "api_core.cuh"

struct cg_job {
  std::vector<cg_context*> ctx;
  std::vector<int> pair_device, pair_slot;  // where the reduction of every pair of the last call lives
  std::string last_error;
};

namespace cg {
void partition_pairs(const int* a, const int* b, const double* cost, int n_pairs, const double* sample_cost,
                     int n_samples, int world, int* owner);
void partition_units(const double* cost, int n, int world, const double* initial_load, int* owner);
}

using namespace cg;

namespace {

struct RankPlan {
  std::vector<cg_sample> samples;  // placeholders where the rank does not need the counts / the PCA
  std::vector<cg_pair> pairs;      // the rank's pairs, ascending
  std::vector<int> pair_ids;
  std::vector<int32_t> local_samples;
};

}  // namespace

extern "C" {

int cg_job_create(int n_devices, cg_job** out) {
  cg_job* job = nullptr;
  try {
    CG_CHECK(out != nullptr, "NULL argument");
    CG_CHECK(n_devices >= 1 && n_devices <= cg_device_count(), "n_devices exceeds the visible GPUs");
    job = new cg_job();
    job->ctx.assign((size_t)n_devices, nullptr);
    unsigned char id[128] = {0};
    if (n_devices > 1 && cg_comm_unique_id(id) != 0) throw Error(cg_last_error(nullptr));
    std::vector<std::string> errs((size_t)n_devices);
    std::vector<std::thread> th;
    for (int r = 0; r < n_devices; ++r)
      th.emplace_back([&, r]() {
        if (cg_init(r, &job->ctx[r]) != 0) { errs[r] = cg_last_error(nullptr); return; }
        if (cg_comm_init(job->ctx[r], r, n_devices, id) != 0) errs[r] = cg_last_error(job->ctx[r]);
      });
    for (auto& t : th) t.join();
    for (auto& e : errs)
      if (!e.empty()) throw Error(e);
    *out = job;
    return 0;
  } catch (const std::exception& e) {
    if (job) {
      for (cg_context* c : job->ctx) cg_shutdown(c);
      delete job;
    }
    return fail(nullptr, e);
  }
}

void cg_job_free(cg_job* job) {
  if (!job) return;
  // NCCL communicators are torn down by all ranks at the same time
  std::vector<std::thread> th;
  for (cg_context* c : job->ctx) th.emplace_back([c]() { cg_shutdown(c); });
  for (auto& t : th) t.join();
  delete job;
}

int cg_job_devices(const cg_job* job) { return job ? (int)job->ctx.size() : 0; }
cg_context* cg_job_context(cg_job* job, int device) {
  return (job && device >= 0 && device < (int)job->ctx.size()) ? job->ctx[device] : nullptr;
}
const char* cg_job_last_error(const cg_job* job) { return job ? job->last_error.c_str() : ""; }

int cg_job_pair_location(const cg_job* job, int pair, int32_t* device, int32_t* slot) {
  if (!job || pair < 0 || pair >= (int)job->pair_device.size()) return 1;
  if (device) *device = job->pair_device[pair];
  if (slot) *slot = job->pair_slot[pair];
  return 0;
}

int cg_job_build_graph(cg_job* job, const cg_sample* samples, int n_samples, const cg_pair* pairs, int n_pairs,
                       const cg_params* params, cg_graph** graph) {
  try {
    CG_CHECK(job && samples && params && graph && (pairs || n_pairs == 0) && n_samples >= 1, "NULL argument");
    const int world = (int)job->ctx.size();
    const cg_params& P = *params;
    for (int q = 0; q < n_pairs; ++q)
      CG_CHECK(pairs[q].a >= 0 && pairs[q].a < n_samples && pairs[q].b >= 0 && pairs[q].b < n_samples &&
                   pairs[q].a != pairs[q].b,
               "pair refers to a missing sample");
    // ---- who does what
    std::vector<int> pa(n_pairs), pb(n_pairs), owner(n_pairs, 0), lowner(n_samples, 0);
    std::vector<double> cost(n_pairs), scost(n_samples), lcost(n_samples);
    for (int s = 0; s < n_samples; ++s) {
      const double g = std::min<double>(samples[s].n_genes, 8000.0);
      scost[s] = 0.005 * samples[s].n_cells * g * g;  // Gram matrix + upload, in cell pairs (measured on configs[1])
      lcost[s] = 0.4 * (double)samples[s].n_cells * samples[s].n_cells;
    }
    for (int q = 0; q < n_pairs; ++q) {
      pa[q] = pairs[q].a;
      pb[q] = pairs[q].b;
      cost[q] = (double)samples[pa[q]].n_cells * samples[pb[q]].n_cells;
    }
    partition_pairs(pa.data(), pb.data(), cost.data(), n_pairs, scost.data(), n_samples, world, owner.data());
    if (P.k_self > 0) {  // local-neighbour searches go to the ranks the pairs left light
      std::vector<double> pload((size_t)world, 0.0);
      for (int q = 0; q < n_pairs; ++q) pload[owner[q]] += cost[q];
      partition_units(lcost.data(), n_samples, world, pload.data(), lowner.data());
    }
    std::vector<RankPlan> plan((size_t)world);
    std::vector<std::vector<char>> need((size_t)world, std::vector<char>((size_t)n_samples, 0));
    job->pair_device.assign((size_t)n_pairs, 0);
    job->pair_slot.assign((size_t)n_pairs, 0);
    bool any_uncached = false;
    for (int q = 0; q < n_pairs; ++q) {
      RankPlan& rp = plan[owner[q]];
      job->pair_device[q] = owner[q];
      job->pair_slot[q] = (int)rp.pairs.size();
      rp.pairs.push_back(pairs[q]);
      rp.pair_ids.push_back(q);
      need[owner[q]][pa[q]] = need[owner[q]][pb[q]] = 1;
      any_uncached = any_uncached || !(pairs[q].rot || (pairs[q].u && pairs[q].v));
    }
    std::vector<int32_t> gram_owner((size_t)n_samples, -1), n_genes((size_t)n_samples, 0);
    {
      std::vector<double> load((size_t)world, 0.0);
      for (int s = 0; s < n_samples; ++s) {
        n_genes[s] = samples[s].n_genes;
        int best = -1;
        for (int r = 0; r < world; ++r)
          if (need[r][s] && (best < 0 || load[r] < load[best])) best = r;
        gram_owner[s] = best;
        if (best >= 0) load[best] += samples[s].n_cells;
      }
    }
    for (int r = 0; r < world; ++r) {
      RankPlan& rp = plan[r];
      rp.samples.assign(samples, samples + n_samples);
      for (int s = 0; s < n_samples; ++s) {
        if (!need[r][s]) rp.samples[s].p = nullptr;  // placeholder: numbers its cells only
        const bool local = P.k_self > 0 && lowner[s] == r;
        if (!local) {
          rp.samples[s].pca = nullptr;
          rp.samples[s].n_pcs = 0;
        } else {
          rp.local_samples.push_back(s);
        }
      }
    }
    // ---- one host thread per device
    std::vector<std::string> errs((size_t)world);
    std::vector<cg_graph*> graphs((size_t)world, nullptr);
    auto rank_main = [&](int r) {
      cg_context* ctx = job->ctx[r];
      cg_panel* panel = nullptr;
      cg_edges* edges = nullptr;
      auto chk = [&](int rc) {
        if (rc != 0) throw Error(cg_last_error(ctx));
      };
      try {
        RankPlan& rp = plan[r];
        chk(cg_panel_create(ctx, rp.samples.data(), n_samples, &panel));
        if (world > 1 && any_uncached) chk(cg_panel_share_grams(ctx, panel, gram_owner.data(), n_genes.data()));
        // one exchange for everything: segments = the pairs (combn order), then the samples' local edges
        std::vector<int64_t> counts(rp.pairs.size() + 1, 0), seg((size_t)n_pairs + n_samples, 0);
        std::vector<int32_t> seg_owner(owner.begin(), owner.end());
        seg_owner.insert(seg_owner.end(), lowner.begin(), lowner.end());
        chk(cg_pairs_edges(ctx, panel, rp.pairs.data(), (int)rp.pairs.size(), &P, &edges, counts.data()));
        for (size_t t = 0; t < rp.pairs.size(); ++t) seg[rp.pair_ids[t]] = counts[t];
        if (P.k_self > 0) {
          std::vector<int64_t> lc(rp.local_samples.size() + 1, 0);
          chk(cg_local_edges(ctx, panel, rp.local_samples.data(), (int)rp.local_samples.size(), &P, &edges, lc.data()));
          for (size_t t = 0; t < rp.local_samples.size(); ++t) seg[(size_t)n_pairs + rp.local_samples[t]] = lc[t];
        }
        chk(cg_edges_allgatherv(ctx, &edges, seg_owner.data(), seg.data(), n_pairs + n_samples));
        if (r == 0) {
          int32_t total = 0;
          for (int s = 0; s < n_samples; ++s) total += samples[s].n_cells;
          chk(cg_assemble_graph(ctx, edges, total, &graphs[0]));
        }
      } catch (const std::exception& e) {
        errs[r] = e.what();
      }
      cg_edges_free(edges);
      cg_panel_free(panel);
    };
    if (world == 1) {
      rank_main(0);
    } else {
      std::vector<std::thread> th;
      for (int r = 0; r < world; ++r) th.emplace_back(rank_main, r);
      for (auto& t : th) t.join();
    }
    for (int r = 0; r < world; ++r)
      if (!errs[r].empty()) {
        if (graphs[0]) cg_graph_free(graphs[0]);
        throw Error("device " + std::to_string(r) + ": " + errs[r]);
      }
    *graph = graphs[0];
    return 0;
  } catch (const std::exception& e) {
    if (job) job->last_error = e.what();
    return fail(nullptr, e);
  }
}

}  // extern "C"
