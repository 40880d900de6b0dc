// Multi-GPU plumbing behind the C ABI: one context per GPU (one process per GPU under torchrun / mpirun, or one
// host thread per GPU inside a single R process), an NCCL communicator owned by the context, the deterministic
// partition of the sample-pair list and the allgatherv of the per-rank edge tables.
//
// Replaces the reference's fork pools over pairs and samples (R/conclass.R:224-317 `papply` over the pair list,
// :1074-1088 inside updatePairs, R/conos.R:396-416 / :589 over samples): pairs are independent units, the only
// exchange step is the concatenation of the per-pair edge data frames (`rbind`, R/conclass.R:316-321), which here is
// ONE ncclAllGather of every rank's packed table, unpacked straight into the reference's row order (pairs in combn
// order, then the samples' local edges) whatever the partition was.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy a host process such as torch has already mapped is
// reused): a single-GPU caller never needs the library.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "api_core.cuh"

namespace cg {

struct Comm {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

namespace {

#define CG_NCCL(cm, expr)                                                                              \
  do {                                                                                                 \
    ncclResult_t _r = (expr);                                                                          \
    if (_r != ncclSuccess) {                                                                           \
      char _buf[512];                                                                                  \
      snprintf(_buf, sizeof(_buf), "NCCL error at %s:%d: %s", __FILE__, __LINE__,                      \
               (cm).GetErrorString ? (cm).GetErrorString(_r) : "?");                                   \
      throw ::cg::Error(_buf);                                                                         \
    }                                                                                                  \
  } while (0)

void bind_nccl(Comm& cm) {
  if (cm.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    cm.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (cm.lib) break;
  }
  CG_CHECK(cm.lib != nullptr, "libnccl.so.2 not found: multi-GPU runs need NCCL");
  auto sym = [&](const char* name) {
    void* p = dlsym(cm.lib, name);
    CG_CHECK(p != nullptr, std::string("NCCL symbol missing: ") + name);
    return p;
  };
  cm.GetUniqueId = reinterpret_cast<decltype(cm.GetUniqueId)>(sym("ncclGetUniqueId"));
  cm.CommInitRank = reinterpret_cast<decltype(cm.CommInitRank)>(sym("ncclCommInitRank"));
  cm.CommDestroy = reinterpret_cast<decltype(cm.CommDestroy)>(sym("ncclCommDestroy"));
  cm.CommAbort = reinterpret_cast<decltype(cm.CommAbort)>(sym("ncclCommAbort"));
  cm.AllReduce = reinterpret_cast<decltype(cm.AllReduce)>(sym("ncclAllReduce"));
  cm.Broadcast = reinterpret_cast<decltype(cm.Broadcast)>(sym("ncclBroadcast"));
  cm.AllGather = reinterpret_cast<decltype(cm.AllGather)>(sym("ncclAllGather"));
  cm.GroupStart = reinterpret_cast<decltype(cm.GroupStart)>(sym("ncclGroupStart"));
  cm.GroupEnd = reinterpret_cast<decltype(cm.GroupEnd)>(sym("ncclGroupEnd"));
  cm.GetErrorString = reinterpret_cast<decltype(cm.GetErrorString)>(sym("ncclGetErrorString"));
}

// Every rank's table travels as ONE packed block [from | to | type | w] of `stride` bytes (n_pad rows per column, the
// largest rank's count rounded up to even) in a single ncclAllGather; this kernel unpacks the gathered blocks straight
// into the output table in segment order: dst row i belongs to the segment s with dst_off[s] <= i < dst_off[s + 1];
// it is row src_row[s] + (i - dst_off[s]) of the table of rank seg_owner[s].
__global__ void __launch_bounds__(256) unpack_segments_kernel(const long long* __restrict__ dst_off,
                                                              const long long* __restrict__ src_row,
                                                              const int* __restrict__ seg_owner, int n_seg,
                                                              long long n_rows, const unsigned char* __restrict__ packed,
                                                              long long stride, long long n_pad, int* __restrict__ df,
                                                              int* __restrict__ dt, int* __restrict__ dty,
                                                              double* __restrict__ dw) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows;
       i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_seg;  // last segment with dst_off <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (dst_off[mid] <= i) lo = mid; else hi = mid;
    }
    const long long r = src_row[lo] + (i - dst_off[lo]);
    const unsigned char* base = packed + (long long)seg_owner[lo] * stride;
    df[i] = reinterpret_cast<const int*>(base)[r];
    dt[i] = reinterpret_cast<const int*>(base + 4 * n_pad)[r];
    dty[i] = reinterpret_cast<const int*>(base + 8 * n_pad)[r];
    dw[i] = reinterpret_cast<const double*>(base + 12 * n_pad)[r];
  }
}

}  // namespace

void comm_destroy(Context& c) {
  if (!c.comm) return;
  if (c.comm->comm && c.comm->CommDestroy) c.comm->CommDestroy(c.comm->comm);
  delete c.comm;  // the library handle stays mapped: other contexts (or the host process) may use it
  c.comm = nullptr;
  c.comm_world = 1;
}

// Partition of the pair list with sample locality.  The samples are cut into g contiguous groups of about equal
// cell totals, g the smallest number with g (g + 1) / 2 >= world; the pair matrix then falls into g (g + 1) / 2
// tiles (group x group), a tile touches two groups only.  Tiles are assigned longest-processing-time first; after
// that single pairs move from the most loaded rank to the least loaded one (preferring pairs whose samples the
// receiver already holds) until the loads agree within 2 %.  A rank then touches about 2 S / g samples instead of all
// of them: fewer Gram matrices to build, fewer samples to upload.  Deterministic: every rank computes the same table.
void partition_pairs(const int* a, const int* b, const double* cost, int n_pairs, const double* sample_cost,
                     int n_samples, int world, int* owner) {
  CG_CHECK(world >= 1 && n_pairs >= 0 && n_samples >= 0, "bad partition arguments");
  for (int q = 0; q < n_pairs; ++q)
    CG_CHECK(a[q] >= 0 && a[q] < n_samples && b[q] >= 0 && b[q] < n_samples, "pair refers to a missing sample");
  if (world == 1 || n_pairs == 0) {
    for (int q = 0; q < n_pairs; ++q) owner[q] = 0;
    return;
  }
  int g = 1;
  while (g * (g + 1) / 2 < world) ++g;
  if (g > n_samples) g = n_samples > 0 ? n_samples : 1;
  // groups of samples: contiguous, balanced by the samples' weight (sample_cost when given, else their pair costs)
  std::vector<double> wgt(n_samples, 0.0);
  for (int q = 0; q < n_pairs; ++q) {
    wgt[a[q]] += cost[q];
    wgt[b[q]] += cost[q];
  }
  if (sample_cost)
    for (int s = 0; s < n_samples; ++s) wgt[s] = sample_cost[s] > 0 ? sample_cost[s] : wgt[s];
  double tot = 0.0;
  for (double v : wgt) tot += v;
  std::vector<int> group(n_samples, 0);
  {
    double run = 0.0;
    for (int s = 0; s < n_samples; ++s) {
      int gi = tot > 0 ? (int)((run + 0.5 * wgt[s]) / tot * g) : s * g / (n_samples > 0 ? n_samples : 1);
      group[s] = gi < 0 ? 0 : (gi >= g ? g - 1 : gi);
      run += wgt[s];
    }
  }
  // tiles
  const int n_tiles = g * g;
  std::vector<double> tile_cost(n_tiles, 0.0);
  std::vector<int> tile_of(n_pairs);
  for (int q = 0; q < n_pairs; ++q) {
    int gi = group[a[q]], gj = group[b[q]];
    if (gi > gj) std::swap(gi, gj);
    tile_of[q] = gi * g + gj;
    tile_cost[tile_of[q]] += cost[q];
  }
  std::vector<int> torder;
  for (int t = 0; t < n_tiles; ++t)
    if (tile_cost[t] > 0) torder.push_back(t);
  std::stable_sort(torder.begin(), torder.end(), [&](int x, int y) { return tile_cost[x] > tile_cost[y]; });
  std::vector<double> load(world, 0.0);
  std::vector<int> tile_owner(n_tiles, 0);
  for (int t : torder) {
    int best = 0;
    for (int r = 1; r < world; ++r)
      if (load[r] < load[best]) best = r;
    tile_owner[t] = best;
    load[best] += tile_cost[t];
  }
  std::vector<std::vector<int>> has(world, std::vector<int>(n_samples, 0));  // pairs of rank r that use sample s
  for (int q = 0; q < n_pairs; ++q) {
    owner[q] = tile_owner[tile_of[q]];
    has[owner[q]][a[q]]++;
    has[owner[q]][b[q]]++;
  }
  // single-pair moves from the fullest to the emptiest rank
  double mean = 0.0;
  for (double v : load) mean += v;
  mean /= world;
  for (int iter = 0; iter < 4 * n_pairs; ++iter) {
    int hi = 0, lo = 0;
    for (int r = 1; r < world; ++r) {
      if (load[r] > load[hi]) hi = r;
      if (load[r] < load[lo]) lo = r;
    }
    const double gap = load[hi] - load[lo];
    if (gap <= 0.02 * mean) break;
    int pick = -1, pick_loc = -1;
    double pick_cost = 0.0;
    for (int q = 0; q < n_pairs; ++q) {
      if (owner[q] != hi || !(cost[q] < gap)) continue;  // the move must lower the maximum of the two
      const int loc = (has[lo][a[q]] > 0) + (has[lo][b[q]] > 0);
      // prefer locality, then the cost closest to half the gap
      const double fit = -std::fabs(cost[q] - 0.5 * gap);
      if (pick < 0 || loc > pick_loc || (loc == pick_loc && fit > pick_cost)) {
        pick = q;
        pick_loc = loc;
        pick_cost = fit;
      }
    }
    if (pick < 0) break;
    owner[pick] = lo;
    load[hi] -= cost[pick];
    load[lo] += cost[pick];
    has[hi][a[pick]]--;
    has[hi][b[pick]]--;
    has[lo][a[pick]]++;
    has[lo][b[pick]]++;
  }
}

// Longest-processing-time assignment of independent units (the samples' local-neighbour searches) onto ranks that
// already carry `initial_load` (their pairs): the ranks the pair quantisation left light take the local work.
void partition_units(const double* cost, int n, int world, const double* initial_load, int* owner) {
  CG_CHECK(world >= 1 && n >= 0, "bad partition arguments");
  std::vector<double> load(world, 0.0);
  if (initial_load)
    for (int r = 0; r < world; ++r) load[r] = initial_load[r];
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
  for (int u : order) {
    int best = 0;
    for (int r = 1; r < world; ++r)
      if (load[r] < load[best]) best = r;
    owner[u] = best;
    load[best] += cost[u];
  }
}

void edges_allgatherv(Context& c, EdgeTable& t, const int* seg_owner, const long long* seg_counts_local, int n_seg) {
  const int world = c.comm ? c.comm->world : 1;
  const int me = c.comm ? c.comm->rank : 0;
  cudaStream_t st = c.stream;
  // ---- counts of every segment on every rank
  std::vector<long long> counts(n_seg, 0);
  long long mine = 0;
  for (int s = 0; s < n_seg; ++s) {
    CG_CHECK(seg_owner[s] >= 0 && seg_owner[s] < world, "segment owner out of range");
    if (seg_owner[s] == me) {
      CG_CHECK(seg_counts_local[s] >= 0, "negative segment count");
      counts[s] = seg_counts_local[s];
      mine += counts[s];
    }
  }
  CG_CHECK((size_t)mine == t.n, "segment counts do not add up to the local edge table");
  if (world == 1) return;  // one rank: the local table already is the gathered one (segments ascending)
  Comm& cm = *c.comm;
  {
    DevBuf<long long> d_counts((size_t)n_seg);
    d_counts.upload(counts.data(), (size_t)n_seg, st);
    CG_NCCL(cm, cm.AllReduce(d_counts.p, d_counts.p, (size_t)n_seg, ncclInt64, ncclSum, cm.comm, st));
    d_counts.download(counts.data(), (size_t)n_seg, st);
    CG_CUDA(cudaStreamSynchronize(st));  // the one host round trip of the exchange: the gathered size
  }
  // ---- layout: every rank's table as one packed block, final = segments ascending
  std::vector<long long> rank_total(world, 0);
  for (int s = 0; s < n_seg; ++s) rank_total[seg_owner[s]] += counts[s];
  long long total = 0, n_pad = 0;
  for (int r = 0; r < world; ++r) {
    total += rank_total[r];
    n_pad = std::max(n_pad, rank_total[r]);
  }
  n_pad = (n_pad + 1) & ~1ll;  // the double column of a block stays 8-byte aligned
  const long long stride = 20 * n_pad;
  std::vector<long long> dst_off(n_seg + 1, 0), src_row(n_seg, 0), cursor(world, 0);
  for (int s = 0; s < n_seg; ++s) {
    dst_off[s + 1] = dst_off[s] + counts[s];
    src_row[s] = cursor[seg_owner[s]];
    cursor[seg_owner[s]] += counts[s];
  }
  EdgeTable out;
  out.reserve((size_t)total, st);
  out.n = (size_t)total;
  if (total == 0) {
    t = std::move(out);
    return;
  }
  DevBuf<unsigned char> packed((size_t)stride * world);
  unsigned char* my = packed.p + (size_t)me * stride;
  if (mine) {
    CG_CUDA(cudaMemcpyAsync(my, t.from.p, (size_t)mine * 4, cudaMemcpyDeviceToDevice, st));
    CG_CUDA(cudaMemcpyAsync(my + 4 * n_pad, t.to.p, (size_t)mine * 4, cudaMemcpyDeviceToDevice, st));
    CG_CUDA(cudaMemcpyAsync(my + 8 * n_pad, t.type.p, (size_t)mine * 4, cudaMemcpyDeviceToDevice, st));
    CG_CUDA(cudaMemcpyAsync(my + 12 * n_pad, t.w.p, (size_t)mine * 8, cudaMemcpyDeviceToDevice, st));
  }
  // one collective over NVLink / NVSwitch, in place
  CG_NCCL(cm, cm.AllGather(my, packed.p, (size_t)stride, ncclChar, cm.comm, st));
  {
    DevBuf<long long> d_dst((size_t)n_seg + 1), d_src((size_t)n_seg);
    DevBuf<int> d_own((size_t)n_seg);
    d_dst.upload(dst_off.data(), dst_off.size(), st);
    d_src.upload(src_row.data(), src_row.size(), st);
    d_own.upload(seg_owner, (size_t)n_seg, st);
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)c.sm_count * 16);
    unpack_segments_kernel<<<blocks, 256, 0, st>>>(d_dst.p, d_src.p, d_own.p, n_seg, total, packed.p, stride, n_pad,
                                                   out.from.p, out.to.p, out.type.p, out.w.p);
    c.stats.kernels++;
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaStreamSynchronize(st));  // the host tables and `packed` go out of scope
  }
  t = std::move(out);
}

// X'X of every sample is computed ONCE in the whole job -- by its owner rank -- and broadcast over NVLink to the ranks
// whose pairs use the sample (everybody takes part in the broadcast; ranks that do not need a matrix receive it into
// a scratch buffer).  Without this every rank builds the Gram matrix of every sample it touches: at 8 GPUs and 8
// samples that is ~5 of them per rank, work that does not shrink with the GPU count.
void ensure_grams(Context& ctx, Sample* const* samples, int n);

void share_grams(Context& c, cg_panel& panel, const int* owner, const int* n_genes) {
  const int S = (int)panel.samples.size();
  const int world = c.comm ? c.comm->world : 1;
  const int me = c.comm ? c.comm->rank : 0;
  cudaStream_t st = c.stream;
  std::vector<Sample*> mine;
  for (int s = 0; s < S; ++s) {
    CG_CHECK(owner[s] >= -1 && owner[s] < world, "gram owner out of range");
    if (owner[s] == me) {
      CG_CHECK(panel.samples[s].has_counts, "the owner of a Gram matrix must hold the sample's counts");
      mine.push_back(&panel.samples[s]);
    }
  }
  {
    StageScope sc(c, "share_grams.compute");
    ensure_grams(c, mine.data(), (int)mine.size());
  }
  if (world == 1) return;
  Comm& cm = *c.comm;
  size_t scratch_elems = 0;
  for (int s = 0; s < S; ++s) {
    if (owner[s] < 0) continue;
    Sample& sm = panel.samples[s];
    const size_t n = (size_t)n_genes[s] * n_genes[s];
    if (sm.has_counts) {
      CG_CHECK(sm.n_genes == n_genes[s], "gene count of a shared Gram matrix does not match the resident sample");
      if (!sm.gram.n) sm.gram.alloc(n);
    } else {
      scratch_elems = std::max(scratch_elems, n);
    }
  }
  DevBuf<double> scratch(scratch_elems);
  StageScope* sc_q = new StageScope(c, "share_grams.enqueue");
  CG_NCCL(cm, cm.GroupStart());
  for (int s = 0; s < S; ++s) {
    if (owner[s] < 0) continue;
    Sample& sm = panel.samples[s];
    double* buf = sm.has_counts ? sm.gram.p : scratch.p;
    CG_NCCL(cm, cm.Broadcast(buf, buf, (size_t)n_genes[s] * n_genes[s], ncclFloat64, owner[s], cm.comm, st));
  }
  CG_NCCL(cm, cm.GroupEnd());
  delete sc_q;
  StageScope sc_w(c, "share_grams.wait");
  CG_CUDA(cudaStreamSynchronize(st));  // `scratch` is released on return
}

}  // namespace cg

using namespace cg;

extern "C" {

int cg_panel_share_grams(cg_context* ctx, cg_panel* panel, const int32_t* owner, const int32_t* n_genes) {
  CG_API_BEGIN
  CG_CHECK(ctx && panel && owner && n_genes, "NULL argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  StageScope sc(ctx->c, "share_grams");
  share_grams(ctx->c, *panel, owner, n_genes);
  CG_API_END(ctx)
}

int cg_comm_unique_id(void* id_out) {
  CG_API_BEGIN
  CG_CHECK(id_out != nullptr, "NULL argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  Comm cm;
  bind_nccl(cm);
  ncclUniqueId id;
  CG_NCCL(cm, cm.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  CG_API_END(nullptr)
}

int cg_comm_init(cg_context* ctx, int rank, int world, const void* id) {
  CG_API_BEGIN
  CG_CHECK(ctx && id, "NULL argument");
  CG_CHECK(world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  comm_destroy(ctx->c);
  Comm* cm = new Comm();
  ctx->c.comm = cm;
  ctx->c.comm_world = world;
  cm->rank = rank;
  cm->world = world;
  if (world > 1) {
    bind_nccl(*cm);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    CG_NCCL(*cm, cm->CommInitRank(&cm->comm, world, uid, rank));
  }
  CG_API_END(ctx)
}

int cg_comm_rank(const cg_context* ctx, int* rank, int* world) {
  if (!ctx) return 1;
  if (rank) *rank = ctx->c.comm ? ctx->c.comm->rank : 0;
  if (world) *world = ctx->c.comm ? ctx->c.comm->world : 1;
  return 0;
}

int cg_partition_pairs(const int32_t* a, const int32_t* b, const double* cost, int n_pairs, const double* sample_cost,
                       int n_samples, int world, int32_t* owner) {
  CG_API_BEGIN
  CG_CHECK((a && b && cost && owner) || n_pairs == 0, "NULL argument");
  partition_pairs(a, b, cost, n_pairs, sample_cost, n_samples, world, owner);
  CG_API_END(nullptr)
}

int cg_partition_units(const double* cost, int n, int world, const double* initial_load, int32_t* owner) {
  CG_API_BEGIN
  CG_CHECK((cost && owner) || n == 0, "NULL argument");
  partition_units(cost, n, world, initial_load, owner);
  CG_API_END(nullptr)
}

int cg_edges_allgatherv(cg_context* ctx, cg_edges** e, const int32_t* seg_owner, const int64_t* seg_counts_local,
                        int n_seg) {
  CG_API_BEGIN
  CG_CHECK(ctx && e && (n_seg == 0 || (seg_owner && seg_counts_local)), "NULL argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  if (!*e) *e = new cg_edges();
  StageScope sc(ctx->c, "edges_allgatherv");
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  edges_allgatherv(ctx->c, (*e)->t, seg_owner, reinterpret_cast<const long long*>(seg_counts_local), n_seg);
  CG_API_END(ctx)
}

}  // extern "C"
