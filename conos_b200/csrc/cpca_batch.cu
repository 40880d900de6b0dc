// Common principal components of MANY sample pairs at once (quickCPCA R/conos.R:204-259 -> cpcaFast :175-191 ->
// cpcaF src/cpca.cpp:16-101, called once per pair by updatePairs R/conclass.R:1074-1088).
//
// One pair's iteration is a chain of ~10 x ncomp dependent p x p matrix-vector products: latency, not work (round 1:
// a cooperative kernel per pair, 35 ms each, 17 s for the 496 pairs of configs[2]).  The pairs are independent, so
// all of them advance in lock step; and the covariance slice of sample s inside pair p never has to exist:
//
//     C_s^(p) = D_p Mc_s D_p / (n_s - 1),   Mc_s = X_s'X_s - n_s m_s m_s'   (shared by every pair of sample s),
//
// D_p = diag(scale factors of the pair, zero outside its gene set).  One step therefore is, for every sample, ONE
// skinny FP64 product  Z_s = Mc_s [D_p q_p : pairs p of s]  (2000 x 2000 x 31 at configs[2]: every pair contributes
// one column to each of its two samples), followed by a per-pair update kernel that runs the reference's state
// machine (d_i = q'C_i q, cost, convergence test, w = sum_i (n_i / d_i) C_i q, deflation, normalisation, next
// component).  Each pair is at its own (component, iteration); a step serves them all.  FP64 throughout: the
// iteration count per component (`niter`) and D are part of the reference's result (src/cpca.cpp:93-100).
//
// Work per step: 8 B * G_s^2 read per sample (HBM-bound once the Gram matrices exceed L2: 1 GB for 32 samples) and
// 2 * G_s^2 * pairs_s FP64 flop per sample.  Deflation uses the stored components, w - Q (Q'w), instead of the
// explicit p x p projector the reference updates (Qw -= q q', src/cpca.cpp:90): the same operator.
#include <algorithm>

#include "api_core.cuh"
#include "reduce.cuh"

namespace cg {
namespace {

constexpr int CP_THREADS = 256;

__global__ void center_gram_kernel(const double* __restrict__ M, const double* __restrict__ colsum, int G, double n,
                                   double* __restrict__ Mc) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * G) return;
  const int i = (int)(t % G), j = (int)(t / G);
  Mc[t] = M[t] - (colsum[i] / n) * (colsum[j] / n) * n;
}

struct SkinnySample {
  const double* Mc;  // G x G column-major, symmetric
  const double* U;   // G x ncols column-major
  double* Z;         // G x ncols column-major
  int G, ncols;
};

// Z = Mc U for every sample: CTA = 128 rows x 32 columns of Z, thread = 8 rows x 2 columns, k in slabs of 32.
// Mc is symmetric, so the (rows r0.., k) slab is read as (k, rows r0..): contiguous in the row index, which is also the
// order the threads want it in shared memory (16-byte vector loads, 5 shared loads per 16 FP64 FMAs).
constexpr int SK_ROWS = 128, SK_COLS = 32, SK_K = 32;
__global__ void __launch_bounds__(CP_THREADS) skinny_gemm_kernel(const SkinnySample* __restrict__ S,
                                                                 const int* __restrict__ tile_sample,
                                                                 const int* __restrict__ tile_row0,
                                                                 const int* __restrict__ tile_col0) {
  __shared__ __align__(16) double As[SK_K][SK_ROWS];
  __shared__ __align__(16) double Bs[SK_K][SK_COLS + 2];
  const SkinnySample sd = S[tile_sample[blockIdx.x]];
  const int row0 = tile_row0[blockIdx.x], col0 = tile_col0[blockIdx.x];
  const int G = sd.G;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // columns 2 tx, 2 tx + 1; rows 8 ty .. 8 ty + 7
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
  for (int k0 = 0; k0 < G; k0 += SK_K) {
#pragma unroll
    for (int i = 0; i < SK_K * SK_ROWS / CP_THREADS; ++i) {
      const int e = threadIdx.x + i * CP_THREADS;
      const int r = e % SK_ROWS, k = e / SK_ROWS;
      As[k][r] = (row0 + r < G && k0 + k < G) ? sd.Mc[(size_t)(k0 + k) * G + row0 + r] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < SK_K * SK_COLS / CP_THREADS; ++i) {
      const int e = threadIdx.x + i * CP_THREADS;
      const int k = e % SK_K, c = e / SK_K;
      Bs[k][c] = (col0 + c < sd.ncols && k0 + k < G) ? sd.U[(size_t)(col0 + c) * G + k0 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < SK_K; ++k) {
      const double2 b = *reinterpret_cast<const double2*>(&Bs[k][2 * tx]);
      double a[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 v = *reinterpret_cast<const double2*>(&As[k][8 * ty + 2 * i]);
        a[2 * i] = v.x;
        a[2 * i + 1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fma(a[i], b.x, acc[i][0]);
        acc[i][1] = fma(a[i], b.y, acc[i][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = col0 + 2 * tx + j;
    if (c >= sd.ncols) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + 8 * ty + i;
      if (r < G) sd.Z[(size_t)c * G + r] = acc[i][j];
    }
  }
}

struct CpcaPair {
  const int* cols[2];     // [G] column of the pair's gene g in sample a / b
  const double* f[2];     // [G] scale factors
  double* U[2];           // this pair's column of the samples' operands (G_s entries, zero outside the gene set)
  const double* Z[2];     // this pair's column of the samples' products
  const double* eigv;     // G x ncomp column-major, leading dimension ldv: initial vectors (irlba(S, ncomp)$v)
  double* q;              // [G]
  double* w;              // [G] scratch
  double* CPC;            // G x ncomp column-major (leading dimension G)
  double* D;              // [2][ncomp]
  double* niter;          // [ncomp]
  double* trace;          // [2] trace of the two covariance slices (score.component.variance)
  const double* diag[2];  // diagonals of Mc_a / Mc_b (G_s entries)
  int* state;             // [4]: comp, it, phase (0 = first product of a component, 1 = iterating, 2 = done), -
  double* cost0;          // [1]
  double ng[2];
  int G, ncomp, ldv, maxit;
  double tol;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < CP_THREADS / 32; ++w) t += red[w];
  return t;
}

// start a component: q = initial vector, operand columns written; the next product gives d_i = q'C_i q (cpca.cpp:48-53)
__device__ void load_component(const CpcaPair& p, int comp) {
  for (int g = threadIdx.x; g < p.G; g += CP_THREADS) {
    const double qv = p.eigv[(size_t)comp * p.ldv + g];
    p.q[g] = qv;
    p.U[0][p.cols[0][g]] = p.f[0][g] * qv;
    p.U[1][p.cols[1][g]] = p.f[1][g] * qv;
  }
}

__global__ void __launch_bounds__(CP_THREADS) cpca_init_kernel(const CpcaPair* __restrict__ P) {
  const CpcaPair p = P[blockIdx.x];
  __shared__ double red[CP_THREADS / 32];
  for (int side = 0; side < 2; ++side) {
    double tr = 0.0;
    for (int g = threadIdx.x; g < p.G; g += CP_THREADS)
      tr += p.f[side][g] * p.f[side][g] * p.diag[side][p.cols[side][g]] / (p.ng[side] - 1.0);
    tr = block_sum(tr, red);
    if (threadIdx.x == 0) p.trace[side] = tr;
  }
  if (threadIdx.x == 0) {
    p.state[0] = 0;
    p.state[1] = 0;
    p.state[2] = 0;
    p.cost0[0] = 0.0;
  }
  load_component(p, 0);
}

// one step of every pair after the products Z = Mc U of the step are in place
__global__ void __launch_bounds__(CP_THREADS) cpca_step_kernel(const CpcaPair* __restrict__ P, int* __restrict__ n_done) {
  const CpcaPair p = P[blockIdx.x];
  if (p.state[2] == 2) return;
  __shared__ double red[CP_THREADS / 32];
  __shared__ double alpha[64];
  const int comp = p.state[0];
  int it = p.state[1];
  const int phase = p.state[2];
  // t_i = C_i q (kept in registers: thread-strided), d_i = q't_i                                (cpca.cpp:50-53, 72-76)
  double da = 0.0, db = 0.0;
  for (int g = threadIdx.x; g < p.G; g += CP_THREADS) {
    const double qv = p.q[g];
    da += qv * (p.f[0][g] * p.Z[0][p.cols[0][g]] / (p.ng[0] - 1.0));
    db += qv * (p.f[1][g] * p.Z[1][p.cols[1][g]] / (p.ng[1] - 1.0));
  }
  da = block_sum(da, red);
  db = block_sum(db, red);
  bool next_component = false;
  if (phase == 1) {
    const double cost = log(da) * p.ng[0] + log(db) * p.ng[1];                                   // cpca.cpp:78
    const double delta = fabs((cost - p.cost0[0]) / cost);
    __syncthreads();
    if (delta < p.tol) {
      next_component = true;                                                                      // break: niter = it
    } else {
      if (threadIdx.x == 0) p.cost0[0] = cost;
      ++it;
      if (it >= p.maxit) next_component = true;                                                   // loop ran out: niter = maxit
    }
  }
  if (next_component) {
    if (threadIdx.x == 0) {
      p.niter[comp] = (double)it;
      p.D[comp] = da;
      p.D[p.ncomp + comp] = db;
    }
    for (int g = threadIdx.x; g < p.G; g += CP_THREADS) p.CPC[(size_t)comp * p.G + g] = p.q[g];
    __syncthreads();
    if (comp + 1 >= p.ncomp) {
      if (threadIdx.x == 0) {
        p.state[2] = 2;
        atomicAdd(n_done, 1);
      }
      return;
    }
    load_component(p, comp + 1);
    if (threadIdx.x == 0) {
      p.state[0] = comp + 1;
      p.state[1] = 0;
      p.state[2] = 0;
      p.cost0[0] = 0.0;
    }
    return;
  }
  // one iteration: w = sum_i (ng_i / d_i) C_i q; deflate; normalise                               (cpca.cpp:58-70)
  const double ca = p.ng[0] / da / (p.ng[0] - 1.0), cb = p.ng[1] / db / (p.ng[1] - 1.0);
  for (int g = threadIdx.x; g < p.G; g += CP_THREADS)
    p.w[g] = ca * p.f[0][g] * p.Z[0][p.cols[0][g]] + cb * p.f[1][g] * p.Z[1][p.cols[1][g]];
  __syncthreads();
  if (comp > 0) {
    // w <- (I - sum_c q_c q_c') w : all coefficients from the same w, like the explicit projector
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < comp; c += CP_THREADS / 32) {
      double s = 0.0;
      for (int g = lane; g < p.G; g += 32) s += p.CPC[(size_t)c * p.G + g] * p.w[g];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      if (lane == 0) alpha[c] = s;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < p.G; g += CP_THREADS) {
      double acc = p.w[g];
      for (int c = 0; c < comp; ++c) acc -= alpha[c] * p.CPC[(size_t)c * p.G + g];
      p.w[g] = acc;
    }
    __syncthreads();
  }
  double nn = 0.0;
  for (int g = threadIdx.x; g < p.G; g += CP_THREADS) nn += p.w[g] * p.w[g];
  nn = sqrt(block_sum(nn, red));
  for (int g = threadIdx.x; g < p.G; g += CP_THREADS) {
    const double qv = p.w[g] / nn;
    p.q[g] = qv;
    p.U[0][p.cols[0][g]] = p.f[0][g] * qv;
    p.U[1][p.cols[1][g]] = p.f[1][g] * qv;
  }
  if (threadIdx.x == 0) {
    p.state[1] = it;
    p.state[2] = 1;
  }
}

__global__ void extract_diag_kernel(const double* __restrict__ M, int G, double* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) d[i] = M[(size_t)i * G + i];
}

}  // namespace

// Centred Gram matrices (Mc = X'X - n m m') of the listed samples, cached for the duration of one reduction call.
void center_gram(Context& ctx, const Sample& s, double* Mc) {
  const size_t n = (size_t)s.n_genes * s.n_genes;
  center_gram_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(s.gram.p, s.colsum.p, s.n_genes,
                                                                         (double)s.n_cells, Mc);
  ctx.stats.kernels++;
}

void cpca_batched(Context& ctx, cg_panel& panel, const CpcaBatchPair* pairs, int n_pairs, const double* const* Mc,
                  int maxit, double tol) {
  if (n_pairs == 0) return;
  cudaStream_t st = ctx.stream;
  const int S = (int)panel.samples.size();
  // ---- columns of every sample: one per (pair, side)
  std::vector<int> ncols(S, 0);
  std::vector<int> slot(2 * (size_t)n_pairs);
  for (int q = 0; q < n_pairs; ++q)
    for (int side = 0; side < 2; ++side) slot[2 * q + side] = ncols[side == 0 ? pairs[q].a : pairs[q].b]++;
  std::vector<size_t> uoff(S + 1, 0);
  for (int s = 0; s < S; ++s) uoff[s + 1] = uoff[s] + (size_t)panel.samples[s].n_genes * ncols[s];
  DevBuf<double> U(uoff[S]), Z(uoff[S]);
  U.zero(st);
  std::vector<DevBuf<double>> diag(S);
  std::vector<SkinnySample> hs(S);
  std::vector<int> t_sample, t_row0, t_col0;
  for (int s = 0; s < S; ++s) {
    const int G = panel.samples[s].n_genes;
    hs[s] = SkinnySample{Mc[s], U.p + uoff[s], Z.p + uoff[s], G, ncols[s]};
    if (ncols[s] == 0) continue;
    diag[s].alloc((size_t)G);
    extract_diag_kernel<<<(G + 255) / 256, 256, 0, st>>>(Mc[s], G, diag[s].p);
    ctx.stats.kernels++;
    for (int c0 = 0; c0 < ncols[s]; c0 += SK_COLS)
      for (int r0 = 0; r0 < G; r0 += SK_ROWS) {
        t_sample.push_back(s);
        t_row0.push_back(r0);
        t_col0.push_back(c0);
      }
  }
  DevBuf<SkinnySample> dS;
  DevBuf<int> dts, dtr, dtc;
  dS.upload(hs.data(), hs.size(), st);
  dts.upload(t_sample.data(), t_sample.size(), st);
  dtr.upload(t_row0.data(), t_row0.size(), st);
  dtc.upload(t_col0.data(), t_col0.size(), st);
  // ---- per-pair state
  size_t vec_elems = 0;
  for (int q = 0; q < n_pairs; ++q) vec_elems += 2 * (size_t)pairs[q].G;
  DevBuf<double> vecs(vec_elems), cost0((size_t)n_pairs);
  DevBuf<int> state((size_t)4 * n_pairs), done(1);
  done.zero(st);
  std::vector<CpcaPair> hp(n_pairs);
  size_t vo = 0;
  int max_steps = 0;
  for (int q = 0; q < n_pairs; ++q) {
    const CpcaBatchPair& b = pairs[q];
    CpcaPair& p = hp[q];
    const int sidx[2] = {b.a, b.b};
    for (int side = 0; side < 2; ++side) {
      const int s = sidx[side];
      const size_t col = uoff[s] + (size_t)slot[2 * q + side] * panel.samples[s].n_genes;
      p.cols[side] = b.cols[side];
      p.f[side] = b.f[side];
      p.U[side] = U.p + col;
      p.Z[side] = Z.p + col;
      p.diag[side] = diag[s].p;
      p.ng[side] = (double)panel.samples[s].n_cells;
    }
    p.eigv = b.eigv;
    p.ldv = b.ldv;
    p.q = vecs.p + vo;
    p.w = vecs.p + vo + b.G;
    vo += 2 * (size_t)b.G;
    p.CPC = b.CPC;
    p.D = b.D;
    p.niter = b.niter;
    p.trace = b.trace;
    p.state = state.p + 4 * (size_t)q;
    p.cost0 = cost0.p + q;
    p.G = b.G;
    p.ncomp = b.ncomp;
    p.maxit = maxit;
    p.tol = tol;
    CG_CHECK(b.ncomp >= 1 && b.ncomp <= 64, "batched CPCA supports up to 64 common components");
    max_steps = std::max(max_steps, b.ncomp * (maxit + 2));
  }
  DevBuf<CpcaPair> dP;
  dP.upload(hp.data(), hp.size(), st);
  cpca_init_kernel<<<n_pairs, CP_THREADS, 0, st>>>(dP.p);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
  // ---- lock step: product, update; the host looks at the number of finished pairs every few steps
  const unsigned n_tiles = (unsigned)t_sample.size();
  int h_done = 0, steps = 0;
  const int check_every = 16;
  while (steps < max_steps) {
    for (int t = 0; t < check_every; ++t, ++steps) {
      skinny_gemm_kernel<<<n_tiles, CP_THREADS, 0, st>>>(dS.p, dts.p, dtr.p, dtc.p);
      cpca_step_kernel<<<n_pairs, CP_THREADS, 0, st>>>(dP.p, done.p);
      ctx.stats.kernels += 2;
    }
    CG_CUDA(cudaGetLastError());
    done.download(&h_done, 1, st);
    CG_CUDA(cudaStreamSynchronize(st));
    if (h_done >= n_pairs) break;
  }
  if (ctx.profile) ctx.stage_ms["reduction.cpca.steps"] += steps;
  CG_CHECK(h_done >= n_pairs, "batched CPCA did not finish (internal error)");
}

}  // namespace cg
