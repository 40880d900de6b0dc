/*
 * conos_gpu.h -- C ABI of libconosgpu.so: the B200 (sm_100a) implementation of the joint-graph
 * construction hot path of conos (Conos$buildGraph).  This is the drop-in boundary: plain pointers and
 * sizes in R's native layouts (dense = column-major double, sparse = dgCMatrix slots p/i/x), int status
 * codes, no R, Rcpp or torch types.  The Rcpp shim (conos_b200/rcpp/conos_gpu_shim.cpp)
 * and the Python host mirror (conos_b200/conos.py, ctypes) both bind exactly these symbols.
 *
 * Every function returns 0 on success and a non-zero code on failure; the message is available through
 * cg_last_error().  Nothing throws or aborts across this boundary (the reference converts C++ exceptions
 * to R errors with BEGIN_RCPP/END_RCPP, src/RcppExports.cpp:74-86; the shim does Rcpp::stop(cg_last_error())).
 *
 * Reference interfaces replaced (paths relative to the conos source tree):
 *   N2R::crossKnn(mA, mB, k, nThreads, verbose, indexType)        R/conos.R:851-855   -> cg_cross_knn
 *   N2R::Knn(m, k, nThreads, verbose, indexType)                  R/conos.R:595       -> cg_self_knn
 *   getNeighborMatrix(p1, p2, k, k1, matching, metric, ...)       R/conos.R:848-886   -> cg_neighbor_matrix
 *   getLocalNeighbors / getLocalEdges                             R/conos.R:587-616   -> cg_local_edges
 *   getPcaBasedNeighborMatrix (scale, centre, project, match)     R/conos.R:776-833   -> cg_pairs_edges
 *   quickPlainPCA / quickCPCA / quickCCA                          R/conos.R:204-369   -> cg_pair_reduction
 *   _conos_spcov                 src/spcov.cpp:14-19,  src/RcppExports.cpp:349-370    -> cg_spcov
 *   _conos_cpcaF                 src/cpca.cpp:16-101                                  -> cg_cpca
 *   _conos_pareDownHubEdges      src/edgeFilter.cpp:23-60                             -> cg_pare_down_hub_edges
 *   adjustWeightsByCellBalancing (50 rounds + downweight)         R/conos.R:936-967    -> cg_graph_balance
 *   _conos_getSumWeightMatrix    src/edge_rebalancing.cpp:14-36                       -> cg_sum_weight_matrix
 *   _conos_adjustWeightsByCellBalancingC  src/edge_rebalancing.cpp:39-48              -> cg_adjust_weights
 *   _conos_referenceWij          src/edgeweights.cpp:160-184                          -> cg_reference_wij
 *   _conos_smooth_count_matrix   src/propagate_labels.cpp:177-222 -> cg_smooth_matrix
 *   _conos_propagate_labels      src/propagate_labels.cpp:67-118                       -> cg_propagate_labels
 *   edge table -> graph_from_edgelist -> simplify                 R/conclass.R:321-343 -> cg_assemble_graph
 *   papply / plapply fork pools over pairs and samples            R/conclass.R:224,1074; R/conos.R:399,589
 *                                                        -> cg_comm_init, cg_partition_pairs, cg_edges_allgatherv
 */
#ifndef CONOS_GPU_H
#define CONOS_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cg_context cg_context;  /* one per process and device; owns the stream and scratch memory */
typedef struct cg_panel cg_panel;      /* device-resident samples (counts, varinfo, per-sample PCA) */
typedef struct cg_edges cg_edges;      /* device-resident edge table in the reference's row order */
typedef struct cg_graph cg_graph;      /* device-resident simplified graph + symmetric adjacency */

enum { CG_METRIC_ANGULAR = 0, CG_METRIC_L2 = 1 };
enum { CG_MATCH_MNN = 0, CG_MATCH_NN = 1 };
enum { CG_SPACE_PCA = 0, CG_SPACE_CPCA = 1, CG_SPACE_CCA = 2 };

/* ---- context ------------------------------------------------------------------------------------- */
int cg_init(int device, cg_context** out);
void cg_shutdown(cg_context* ctx);
const char* cg_last_error(const cg_context* ctx); /* ctx may be NULL: last error of this thread */
int cg_device_count(void);
/* kernels launched by this context so far (bench.py reports the delta as gpu_launches) */
uint64_t cg_kernel_launches(const cg_context* ctx);
/* CUDA events on the context's stream: start, then stop returns the elapsed device time in ms */
int cg_timer_start(cg_context* ctx);
int cg_timer_stop(cg_context* ctx, double* ms);
/* options: "exact_reduction" = 1 runs the per-pair reductions in FP64 on the CUDA cores (Gram by FP64 atomics,
 * explicit covariance, subspace iteration to 1e-10) instead of the default tensor-core path (bf16x3 GEMMs,
 * Chebyshev-filtered iteration to irlba's own tolerance 1e-5) -- used by the parity tests */
/* "knn_engine": 0 = automatic (two-phase tensor-core kernel; the streaming tensor-core kernel for small problems
 * and for blocks whose candidate lists overflow), 1 = streaming tensor-core kernel only, 2 = exact SIMT kernel
 * only.  "knn_cand_cap": candidate slots per (query row, half tile) of the two-phase kernel, 1..64 (tests
 * shrink it to force the overflow path).  Results are identical for every setting.
 * "tc_eig_max_sweeps" (default 40): sweeps of the tensor-core eigen solver before a pair that has not reached
 * irlba's tolerance is handed to the FP64 driver (tests set 1 to force that path).
 * "gram_chunk_slabs" (default 16): the tensor-core Gram matrix X'X accumulates the cells in FP32 chunks of 64 x this
 * many cells and sums the chunks in FP64 (0: one FP32 accumulation over all cells, round 1's behaviour).
 * "tc_eig_tol" (default 1e-5 = irlba's): residual bound of that solver relative to the largest Ritz value.
 * "tc_eig_degree" (default 10, 2..64): largest Chebyshev degree of one sweep of that solver (8 / 10 / 12 / 14 were
 * measured on configs[1]: 10.4 / 8.8 / 9.1 / 9.2 ms for the 56 eigenproblems).
 * "tc_projection" (default 1): in the default mode the projections of device-computed rotations run as one bf16x3
 * tensor-core GEMM per sample (fp32 accumulation); 0 keeps the FP64 sparse product, which the exact mode, cached
 * rotations and cg_keep_projections always use.
 * "max_batch_bytes" (default 48 GB): scratch budget of one batch of pairs inside cg_pairs_edges; the result never
 * depends on how the pair list is cut into batches (tests set 1 to get one pair per batch).
 * "refresh_free_memory" (any value): look up the free device memory again.  The batch sizes of cg_pairs_edges come
 * from an estimate taken at cg_init and adjusted by the panels this context holds (the driver query stalls for up to
 * 100 ms on some hosts, so it stays out of the hot calls unless a job does not obviously fit); a caller that has since
 * given most of the device to something else should say so here.
 * "eig_cluster" (default 0 = automatic): 1 / 4 force the single-CTA / four-CTA-cluster form of the eigen
 * factorisation kernels; results are bit-identical (every row sum runs over the same four slices in the same order).
 * "knn_stagger" (cycles, default 0) and "knn_ld32" (default 0): scheduling experiments of the two-phase kernel's
 * epilogue (measured without effect, results identical); kept for re-measurement only. */
int cg_set_option(cg_context* ctx, const char* name, double value);
/* counters: "knn_redo_blocks" (256-row blocks recomputed by the streaming kernel), "eig_redo_pairs" (pairs the
 * tensor-core eigen solver handed to the FP64 driver), "kernel_launches" */
int cg_get_stat(cg_context* ctx, const char* name, double* value);
/* optional per-stage wall-clock profile: enable (resets), then dump "name=ms;name=ms;..." into buf.  enable = 1
 * synchronises the stream at both ends of every stage (device time per stage, slows the step); enable = 2 only reads
 * the host clock (where the host thread spends a step; free) */
int cg_profile(cg_context* ctx, int enable);
int cg_profile_dump(cg_context* ctx, char* buf, int buf_len);
/* optional CUDA-event timing of the dominant kernel (cross-kNN): enable, then read and reset.  `candidates` follows
 * n_a * n_b per sample pair, ONE distance matrix for both directions (a self search counts n^2);
 * cg_get_stat("knn_scores") gives the scores the kernels actually scanned (per direction). */
int cg_knn_timing(cg_context* ctx, int enable);
int cg_knn_timing_read(cg_context* ctx, double* ms, uint64_t* launches, double* candidates, double* bytes);

/* ---- exact kNN (float32 arithmetic of N2, exact search) -------------------------------------------- */
/* A: nA x d, B: nB x d column-major double (lda/ldb leading dimensions).
 * idx_ab/dist_ab: [nA][k] row-major -- for every row of A its k nearest rows of B (ascending distance, ties by
 * lower index; -1 / +inf when nB < k); idx_ba/dist_ba: [nB][k] likewise.  Either output pair may be NULL.
 * As sparse matrices these are N2R::crossKnn(A,B,k) (nB x nA, entry (neighbour, query)) and crossKnn(B,A,k). */
int cg_cross_knn(cg_context* ctx, const double* A, int nA, int lda, const double* B, int nB, int ldb, int d, int k,
                 int metric, int32_t* idx_ab, float* dist_ab, int32_t* idx_ba, float* dist_ba);
/* N2R::Knn(Z, k): the sample against itself, the point itself included. */
int cg_self_knn(cg_context* ctx, const double* Z, int n, int ldz, int d, int k, int metric, int32_t* idx,
                float* dist);

/* ---- getNeighborMatrix ---------------------------------------------------------------------------- */
/* p1: n1 x d, p2: n2 x d column-major double.  Result = TsparseMatrix n1 x n2 triplets in CSC order; fetch with
 * cg_fetch_coo after reading *nnz.  k1 > k applies reduceEdgesInGraphIteratively (R/conos.R:906-933). */
int cg_neighbor_matrix(cg_context* ctx, const double* p1, int n1, int ld1, const double* p2, int n2, int ld2, int d,
                       int k, int k1, int matching, int metric, double l2_sigma, double cor_base,
                       double min_similarity, int64_t* nnz);
int cg_fetch_coo(cg_context* ctx, int32_t* i, int32_t* j, double* x);

/* ---- samples ------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_cells, n_genes;
  const int32_t* p; /* counts: dgCMatrix cells x genes, @p [n_genes+1]; NULL = placeholder: the sample's counts are not
                       uploaded to this device (multi-GPU: only the ranks whose pairs touch a sample hold it); n_cells
                       still numbers its cells and pca may still be given for the local neighbours */
  const int32_t* i; /*                                   @i [nnz] (0-based cell)  */
  const double* x;  /*                                   @x [nnz]  */
  const double* var_v;  /* misc$varinfo$v  per gene (NULL when var.scale = FALSE) */
  const double* var_qv; /* misc$varinfo$qv per gene */
  const double* pca;    /* reductions$PCA, column-major n_cells x n_pcs (NULL when k.self = 0) */
  int32_t n_pcs, pca_ld;
} cg_sample;

int cg_panel_create(cg_context* ctx, const cg_sample* samples, int n_samples, cg_panel** out);
void cg_panel_free(cg_panel* panel);
int cg_panel_offsets(const cg_panel* panel, int32_t* offsets /* [n_samples + 1] global cell ids */);
/* drop derived per-sample data (Gram matrices, kNN panels of the PCA) so that the next call recomputes them */
int cg_panel_drop_cache(cg_panel* panel);

/* ---- per-pair reduction + matching --------------------------------------------------------------- */
typedef struct {
  int32_t a, b;           /* sample indices (combn order is the caller's: R/conclass.R:1058) */
  int32_t n_genes;        /* common od genes of the pair, in commonOverdispersedGenes order (R/conos.R:107-117) */
  const int32_t* genes_a; /* column of sample a holding od gene g */
  const int32_t* genes_b;
  const double* rot;      /* optional cached rotation, column-major rot_rows x rot_cols (self$pairs[[space]]$CPC);
                             rot_rows must equal n_genes: the pair's genes ARE rownames(CPC) (R/conclass.R:240-243) */
  int32_t rot_rows, rot_cols;
  const double* u;        /* CCA: optional cached u (n_u x uv_cols) and v (n_v x uv_cols), column-major; every cached */
  const double* v;        /* column is used, as the reference does (R/conclass.R:268-270)                            */
  const int32_t* rows_u;  /* CCA: cell index of every row of u / v (cells with empty rows are dropped) */
  const int32_t* rows_v;
  int32_t n_u, n_v, uv_cols;
  int32_t k;              /* 0: params->k; > 0: k.cur of this pair = min(k.same.factor, k1) for two samples of the same
                             balancing.factor.per.sample level (R/conclass.R:230-233) */
} cg_pair;

typedef struct {
  int32_t k, k1, k_self;
  int32_t ncomps;
  int32_t space, matching, metric;
  int32_t var_scale, common_centering, score_component_variance;
  double k_self_weight, l2_sigma, cor_base, min_similarity;
  int32_t cpca_maxit; /* 500  (R/conos.R:242) */
  double cpca_tol;    /* 1e-5 */
} cg_params;

void cg_params_default(cg_params* p); /* buildGraph defaults, R/conclass.R:161-165 */

/* Computes (or takes from pair.rot / pair.u,v) the reduction of every pair, projects, runs the two-directional
 * exact kNN, the mutual-neighbour matching and (k1 > k) hub paring, and appends the inter-sample edges
 * (type 1, from = cell of a, to = cell of b, global cell ids) pair after pair.  *edges may be NULL (created). */
int cg_pairs_edges(cg_context* ctx, cg_panel* panel, const cg_pair* pairs, int n_pairs, const cg_params* params,
                   cg_edges** edges, int64_t* pair_edge_counts /* [n_pairs] or NULL */);
/* The reduction of the most recent cg_pairs_edges call for pair `which`: PCA/CPCA: rot n_genes x ncols;
 * nv: per-sample variance fractions [2][ncols] when score_component_variance.  Query sizes with NULL outputs. */
int cg_pair_reduction(cg_context* ctx, int which, int32_t* n_rows, int32_t* n_cols, double* rot, double* nv);
/* CCA reduction of pair `which` computed by the most recent call: u (n_u x d), v (n_v x d) column-major, already
 * scaled by sqrt(d)/max (R/conos.R:363-366); rows_u / rows_v = cell index of every row (cells with empty rows are
 * dropped, :336); sigma = res$d.  Query the sizes with NULL outputs. */
int cg_pair_cca(cg_context* ctx, int which, int32_t* n_u, int32_t* n_v, int32_t* d, double* u, double* v,
                int32_t* rows_u, int32_t* rows_v, double* sigma);
/* The rotations of ALL pairs of the most recent cg_pairs_edges call at once: n_rows / n_cols [n_pairs], offsets
 * [n_pairs + 1] (doubles) into `rot`; call with rot == NULL to size the buffer (offsets[n_pairs]), then again. */
int cg_pair_reductions_all(cg_context* ctx, int32_t* n_rows, int32_t* n_cols, int64_t* offsets, double* rot,
                           int64_t rot_capacity);
/* CPCA: the other fields of cpcaF's result for pair `which` (src/cpca.cpp:93-100; self$pairs$CPCA keeps them,
 * R/conos.R:242): D = q'C_i q at convergence, column-major n_cols x 2 like the reference's t(D) (column 0 = sample a);
 * niter [n_cols]. */
int cg_pair_cpca(cg_context* ctx, int which, int32_t* n_cols, double* D, double* niter);
/* CCA: gene loadings ul, vl (n_genes x d column-major, unit columns) and variance fractions nv [2][d] of pair
 * `which` when it was computed by the most recent call (R/conos.R:349-357).  Query sizes with NULL outputs. */
int cg_pair_cca_loadings(cg_context* ctx, int which, int32_t* n_genes, int32_t* d, double* ul, double* vl, double* nv);
/* Projections of pair `which` of the most recent call (column-major n x ncomps doubles), for parity tests;
 * they are only kept after cg_keep_projections(ctx, 1). */
int cg_keep_projections(cg_context* ctx, int enable);
int cg_pair_projection(cg_context* ctx, int which, int side, int32_t* n_rows, int32_t* n_cols, double* p);

/* Within-sample edges of the listed samples (type 0), appended sample after sample; sample_edge_counts ([n] or NULL)
 * receives the rows appended per listed sample. */
int cg_local_edges(cg_context* ctx, cg_panel* panel, const int32_t* samples, int n, const cg_params* params,
                   cg_edges** edges, int64_t* sample_edge_counts);

/* ---- edge tables ---------------------------------------------------------------------------------- */
int64_t cg_edges_count(const cg_edges* e);
/* rows whose type equals `type` (1 = inter-sample: the "mNN edges" of the metric; 0 = within-sample); -1 on error */
int64_t cg_edges_count_type(cg_context* ctx, const cg_edges* e, int32_t type);
int cg_edges_fetch(cg_context* ctx, const cg_edges* e, int32_t* from, int32_t* to, double* w, int32_t* type);
/* append host rows (used to merge shards gathered from other ranks) */
int cg_edges_append_host(cg_context* ctx, cg_edges** e, const int32_t* from, const int32_t* to, const double* w,
                         const int32_t* type, int64_t n);
/* append rows that already live on this device (the shards gathered over NCCL) */
int cg_edges_append_device(cg_context* ctx, cg_edges** e, const void* from, const void* to, const void* w,
                           const void* type, int64_t n);
/* device pointers of the columns (for NCCL gathers driven by the host language) */
int cg_edges_device_ptrs(const cg_edges* e, void** from, void** to, void** w, void** type);
int cg_edges_reserve(cg_context* ctx, cg_edges** e, int64_t n_total);
int cg_edges_set_count(cg_edges* e, int64_t n);
void cg_edges_free(cg_edges* e);

/* ---- multi-GPU: one context per GPU ------------------------------------------------------------------ */
/* Replaces the reference's fork pools over the pair list and the samples (R/conclass.R:224-317, :1074-1088,
 * R/conos.R:396-416, :589).  One context per GPU -- one process per GPU (torchrun, mpirun) or one host thread per
 * GPU inside a single R process; rank 0 creates the id, every rank passes the same 128 bytes to cg_comm_init
 * (NCCL communicator over NVLink; world == 1 needs no NCCL at all). */
int cg_comm_unique_id(void* id_out /* 128 bytes */);
int cg_comm_init(cg_context* ctx, int rank, int world, const void* id /* 128 bytes */);
int cg_comm_rank(const cg_context* ctx, int* rank, int* world);
/* Deterministic partition of the pair list: longest-processing-time greedy over cost[q] (cells_a x cells_b) with
 * sample locality -- a rank is charged sample_cost[s] the first time it receives a pair of sample s (its Gram
 * matrix and its upload), so a rank touches about S / sqrt(world) samples.  owner[q] = rank.  sample_cost may be NULL. */
int cg_partition_pairs(const int32_t* a, const int32_t* b, const double* cost, int n_pairs, const double* sample_cost,
                       int n_samples, int world, int32_t* owner);
/* Independent units (the samples' local-neighbour searches, cost ~ 0.4 cells^2 in the unit of the pair costs) onto
 * ranks that already carry initial_load[r] (the sum of their pair costs; may be NULL): longest processing time first,
 * so the ranks the pair quantisation left light take the local work. */
int cg_partition_units(const double* cost, int n, int world, const double* initial_load, int32_t* owner);
/* Gram matrices X'X of the samples, computed once per job: owner[s] = rank that computes the matrix of sample s (it
 * must hold the sample's counts; -1: nobody needs it), n_genes[s] = its gene count (known to every rank).  Every
 * rank calls this with the same tables before cg_pairs_edges; the matrices travel by ncclBroadcast over NVLink.
 * Replaces the per-pair spcov calls of quickCPCA / the per-sample irlba of quickPlainPCA inside the reference's
 * fork pool (R/conclass.R:1074-1088): there every worker recomputes what it needs. */
int cg_panel_share_grams(cg_context* ctx, cg_panel* panel, const int32_t* owner, const int32_t* n_genes);
/* allgatherv of the ranks' edge tables.  The table is a sequence of segments (pairs for the inter-sample edges,
 * samples for the local ones): seg_owner[s] = rank that produced segment s, seg_counts_local[s] = its row count on
 * the owner (ignored elsewhere); the local table holds this rank's segments in ascending s.  On return every rank
 * holds all segments in ascending s == the reference's row order (pairs in combn order, then the samples' local
 * edges: R/conclass.R:316-321, 328-332).  One ncclAllReduce of the counts, ONE ncclAllGather of the packed tables
 * (from | to | type | w per rank), one kernel that unpacks into segment order; a single-rank context returns the
 * table as it is. */
int cg_edges_allgatherv(cg_context* ctx, cg_edges** e, const int32_t* seg_owner, const int64_t* seg_counts_local,
                        int n_seg);

/* ---- one buildGraph over several GPUs from ONE host process ------------------------------------------ */
/* What an R session needs (no fork with a CUDA context, no torchrun): cg_job_create starts one context per device
 * 0 .. n_devices-1 and an NCCL communicator between them (n_devices == 1: no NCCL); cg_job_build_graph runs, with
 * one host thread per device, partition -> resident samples -> shared Gram matrices -> cg_pairs_edges on the
 * device's pairs -> cg_edges_allgatherv -> cg_local_edges on the device's samples -> cg_edges_allgatherv -> graph
 * assembly on device 0.  `samples` / `pairs` are the full tables (every pair is computed once, on the device the
 * partition gives it).  The graph belongs to cg_job_context(job, 0): fetch it with cg_graph_fetch on that context.
 * The reduction of pair q stays on its device: cg_job_pair_location tells which context and which slot to pass to
 * cg_pair_reduction / cg_pair_cpca / cg_pair_cca / cg_pair_cca_loadings.
 * Replaces R/conclass.R:224-343 (both papply loops, rbind, graph_from_edgelist, simplify) and :1074-1088. */
typedef struct cg_job cg_job;
int cg_job_create(int n_devices, cg_job** out);
void cg_job_free(cg_job* job);
int cg_job_devices(const cg_job* job);
cg_context* cg_job_context(cg_job* job, int device);
const char* cg_job_last_error(const cg_job* job);
int cg_job_build_graph(cg_job* job, const cg_sample* samples, int n_samples, const cg_pair* pairs, int n_pairs,
                       const cg_params* params, cg_graph** graph);
int cg_job_pair_location(const cg_job* job, int pair, int32_t* device, int32_t* slot);

/* ---- graph ---------------------------------------------------------------------------------------- */
/* el[w > 0]; vertices numbered by first appearance (graph_from_edgelist on a character matrix);
 * simplify(edge.attr.comb = list(weight = "sum", type = "first")).  R/conclass.R:335-343. */
int cg_assemble_graph(cg_context* ctx, const cg_edges* e, int32_t n_cells_total, cg_graph** out);
int cg_graph_sizes(const cg_graph* g, int32_t* n_vertices, int64_t* n_edges);
int cg_graph_fetch(cg_context* ctx, const cg_graph* g, int32_t* vertices, int32_t* from, int32_t* to, double* weight,
                   int32_t* type);
/* symmetric adjacency (as_adj(graph, attr = "weight")): p [n_vertices+1], i/x [2*n_edges] */
int cg_graph_fetch_adjacency(cg_context* ctx, const cg_graph* g, int64_t* p, int32_t* i, double* x);
/* adjustWeightsByCellBalancing on the assembled graph, in place (R/conos.R:936-967 driven from R/conclass.R:346-368):
 * factor_per_vertex[v] = 1-based level of `as.factor(factor.per.cell[colnames(adj)])` for graph vertex v;
 * balance_weights != 0 runs the n_iters (50) rounds of getSumWeightMatrix / adjustWeightsByCellBalancingC
 * (src/edge_rebalancing.cpp:14-48) on the device without a host round trip; same_factor_downweight multiplies the
 * edges inside one level (applied when |f - 1| > 1e-5).  The edge weights are then re-read from the adjacency matrix
 * and `type` is set to -1: the reference rebuilds the graph from the adjacency matrix and loses the attribute. */
int cg_graph_balance(cg_context* ctx, cg_graph* g, const int32_t* factor_per_vertex, int n_levels, int balance_weights,
                     double same_factor_downweight, int n_iters);
void cg_graph_free(cg_graph* g);

/* ---- the reference's own .Call entry points, same argument meaning ----------------------------------- */
/* spcov(m, cm): m dgCMatrix n x G, cm [G]; out column-major G x G.   (V = m'm - n cm cm'; V /= n-1) */
int cg_spcov(cg_context* ctx, int n, int G, const int32_t* p, const int32_t* i, const double* x, const double* cm,
             double* out);
/* cpcaF(cov p x p x k, ng, ncomp, maxit, tol, eigenvR): D (k x ncomp), CPC (p x ncomp), niter [ncomp] */
int cg_cpca(cg_context* ctx, const double* cov, int p, int k, const double* ng, int ncomp, int maxit, double tol,
            const double* eigv, double* D, double* CPC, double* niter);
/* pareDownHubEdges(sY, rowN, k, klow): x is updated in place, rowN is mutated like the reference does */
int cg_pare_down_hub_edges(cg_context* ctx, int ncols, const int32_t* p, const int32_t* i, double* x, int32_t* rowN,
                           int k, int klow);
int cg_sum_weight_matrix(cg_context* ctx, const double* weights, const int32_t* row_inds, const int32_t* col_inds,
                         int64_t n_edges, const int32_t* factor_levels, int n_cells, int n_levels, int normalize,
                         double* m /* n_cells x n_levels column-major */);
int cg_adjust_weights(cg_context* ctx, double* weights, const int32_t* row_inds, const int32_t* col_inds,
                      int64_t n_edges, const int32_t* factor_levels, int n_cells, int n_levels,
                      const double* dividers);
/* referenceWij(i, j, d, threads, perplexity): triplets of the result before duplicate summation; outputs sized
 * 2 * n_edges; *n_out receives the triplet count */
int cg_reference_wij(cg_context* ctx, const int32_t* from, const int32_t* to, const double* d, int64_t n_edges,
                     double perplexity, int32_t* out_from, int32_t* out_to, double* out_w, int64_t* n_out);

/* ---- diffusion over the graph (the consumers of the joint graph next to the path: SURVEY.md 8(f) rank 3) -------- */
/* smooth_count_matrix(edge_verts, edge_weights, count_matrix, is_label_fixed, max_n_iters, diffusion_fading,
 * diffusion_fading_const, tol, verbose, normalize)  src/propagate_labels.cpp:177-222 -> :120-175; called by
 * Conos$correctGenes (R/conclass.R:862-893).  Vertices are 0-based integers = rows of `matrix` (the reference maps the
 * vertex names to the row names, :199-202); `matrix` is an R matrix (column-major n_vertices x n_cols), updated in
 * place.  Edges: either a graph handle (its simplified edge list is already on the device; from / to / weight are
 * ignored) or host arrays.  is_fixed: NULL or one byte per vertex.  *n_iters receives the iteration index at which the
 * loop stopped (the reference prints it), *inf_norm the last max |new - old|.  As in the reference the matrix of the
 * iteration BEFORE the one that met `tol` is returned.  Error "Negative edge length" like Edge::Edge (:24-26). */
int cg_smooth_matrix(cg_context* ctx, const cg_graph* graph, const int32_t* from, const int32_t* to,
                     const double* weight, int64_t n_edges, int32_t n_vertices, double* matrix, int32_t n_cols,
                     const uint8_t* is_fixed, int32_t max_n_iters, double fading, double fading_const, double tol,
                     int32_t normalize, int32_t* n_iters, double* inf_norm);
/* propagate_labels(edge_verts, edge_weights, vert_labels, max_n_iters, verbose, diffusion_fading,
 * diffusion_fading_const, tol, fixed_initial_labels)  src/propagate_labels.cpp:67-118; called by
 * propagateLabelsDiffusion (R/conos.R:1071-1083).  label_of_vertex: label id per vertex, -1 = unlabelled (the caller
 * numbers labels by first appearance like order_strings, :33-38); label_dist: n_vertices x n_labels column-major,
 * rows divided by their sums (:100-106; a vertex no label reaches gets 0 / 0 = NaN as in R). */
int cg_propagate_labels(cg_context* ctx, const cg_graph* graph, const int32_t* from, const int32_t* to,
                        const double* weight, int64_t n_edges, int32_t n_vertices, const int32_t* label_of_vertex,
                        int32_t n_labels, int32_t max_n_iters, double fading, double fading_const, double tol,
                        int32_t fixed_initial_labels, double* label_dist, int32_t* n_iters, double* inf_norm);

#ifdef __cplusplus
}
#endif
#endif /* CONOS_GPU_H */
