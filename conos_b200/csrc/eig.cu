#include "eig.cuh"

namespace cg {
namespace {

constexpr int EIG_THREADS = 256;

struct EigWork {
  const double* C;
  int G;
  double* Q;       // [G][B] row-major
  double* Z;       // [G][B] row-major
  double* theta;   // [B] current Ritz values
  double* V;       // out
  double* theta_out;
  int* converged;  // 1 when the top-r Ritz values stagnated
};

__device__ __forceinline__ double hash_unit(unsigned a, unsigned b) {
  unsigned x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return (double)x * (2.0 / 4294967296.0) - 1.0;
}

template <int B>
__global__ void eig_init_kernel(const EigWork* __restrict__ works) {
  const EigWork w = works[blockIdx.y];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (size_t)w.G * B) w.Z[t] = hash_unit((unsigned)(t / B), (unsigned)(t % B));
}

// Z = C Q.  One warp per row of C (symmetric: row r == column r, contiguous), lanes over the B columns of Q.
template <int B>
__global__ void __launch_bounds__(256) eig_matmul_kernel(const EigWork* __restrict__ works) {
  const EigWork w = works[blockIdx.y];
  if (w.converged && *w.converged == 2) return;
  __shared__ double crow[8][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  const int G = w.G;
  constexpr int NC = B / 32;
  double acc[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) acc[u] = 0.0;
  const double* col = w.C + (size_t)(r < G ? r : 0) * G;
  for (int c0 = 0; c0 < G; c0 += 128) {
    __syncwarp();
    for (int t = lane; t < 128; t += 32) crow[warp][t] = (c0 + t < G) ? col[c0 + t] : 0.0;
    __syncwarp();
    const int lim = G - c0 < 128 ? G - c0 : 128;
    for (int t = 0; t < lim; ++t) {
      const double cv = crow[warp][t];
      const double* q = w.Q + (size_t)(c0 + t) * B;
#pragma unroll
      for (int u = 0; u < NC; ++u) acc[u] = fma(cv, q[lane + 32 * u], acc[u]);
    }
  }
  if (r < G) {
#pragma unroll
    for (int u = 0; u < NC; ++u) w.Z[(size_t)r * B + lane + 32 * u] = acc[u];
  }
}

// S[i][j] = sum_r X[r][i] * Y[r][j]  (B x B, into shared memory), all 256 threads
template <int B>
__device__ void block_gram(const double* __restrict__ X, const double* __restrict__ Y, int G, double (*S)[B + 1]) {
  constexpr int TB = B / 16;
  const int i0 = TB * (threadIdx.x / 16), j0 = TB * (threadIdx.x % 16);
  double acc[TB][TB];
#pragma unroll
  for (int a = 0; a < TB; ++a)
#pragma unroll
    for (int b = 0; b < TB; ++b) acc[a][b] = 0.0;
  for (int r = 0; r < G; ++r) {
    double xv[TB], yv[TB];
#pragma unroll
    for (int a = 0; a < TB; ++a) { xv[a] = X[(size_t)r * B + i0 + a]; yv[a] = Y[(size_t)r * B + j0 + a]; }
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
      for (int b = 0; b < TB; ++b) acc[a][b] = fma(xv[a], yv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < TB; ++a)
#pragma unroll
    for (int b = 0; b < TB; ++b) S[i0 + a][j0 + b] = acc[a][b];
  __syncthreads();
}

// dst[r][:] = src[r][:] * W  (W B x B in shared memory), one warp per row; src may equal dst
template <int B>
__device__ void rows_times_matrix(const double* __restrict__ src, double* __restrict__ dst, int G,
                                  double (*W)[B + 1], double (*rowbuf)[B]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NC = B / 32;
  for (int r = warp; r < G; r += EIG_THREADS / 32) {
#pragma unroll
    for (int u = 0; u < NC; ++u) rowbuf[warp][lane + 32 * u] = src[(size_t)r * B + lane + 32 * u];
    __syncwarp();
    double acc[NC];
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] = 0.0;
    for (int i = 0; i < B; ++i) {
      const double x = rowbuf[warp][i];
#pragma unroll
      for (int u = 0; u < NC; ++u) acc[u] = fma(x, W[i][lane + 32 * u], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < NC; ++u) dst[(size_t)r * B + lane + 32 * u] = acc[u];
    __syncwarp();
  }
  __syncthreads();
}

// Cyclic Jacobi (round-robin parallel ordering) on the symmetric A (destroyed: diagonal = eigenvalues),
// W = eigenvectors in columns.  All 256 threads.
template <int B>
__device__ void block_jacobi(double (*A)[B + 1], double (*W)[B + 1], double* cs, double* red) {
  const int tid = threadIdx.x;
  for (int t = tid; t < B * B; t += EIG_THREADS) W[t / B][t % B] = (t / B == t % B) ? 1.0 : 0.0;
  __syncthreads();
  constexpr int HP = B / 2;
  for (int sweep = 0; sweep < 30; ++sweep) {
    // convergence: largest off-diagonal against largest diagonal
    double off = 0.0, dg = 0.0;
    for (int t = tid; t < B * B; t += EIG_THREADS) {
      const double v = fabs(A[t / B][t % B]);
      if (t / B == t % B) dg = fmax(dg, v); else off = fmax(off, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
      off = fmax(off, __shfl_xor_sync(0xffffffffu, off, d));
      dg = fmax(dg, __shfl_xor_sync(0xffffffffu, dg, d));
    }
    if ((tid & 31) == 0) { red[tid >> 5] = off; red[8 + (tid >> 5)] = dg; }
    __syncthreads();
    off = 0.0; dg = 0.0;
    for (int wv = 0; wv < 8; ++wv) { off = fmax(off, red[wv]); dg = fmax(dg, red[8 + wv]); }
    __syncthreads();
    if (off <= 1e-17 * dg || dg == 0.0) break;
    for (int step = 0; step < B - 1; ++step) {
      if (tid < HP) {
        int p, q;
        if (tid == 0) { p = B - 1; q = step; }
        else { p = (step + tid) % (B - 1); q = (step - tid + (B - 1)) % (B - 1); }
        const double apq = A[p][q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300) {
          const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
          const double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + tt * tt);
          s = tt * c;
        }
        cs[3 * tid] = c; cs[3 * tid + 1] = s; cs[3 * tid + 2] = (double)(p * B + q);
      }
      __syncthreads();
      // columns: A <- A J, W <- W J
      for (int t = tid; t < HP * B; t += EIG_THREADS) {
        const int pr = t / B, i = t % B;
        const double c = cs[3 * pr], s = cs[3 * pr + 1];
        const int pq = (int)cs[3 * pr + 2];
        const int p = pq / B, q = pq % B;
        const double aip = A[i][p], aiq = A[i][q];
        A[i][p] = c * aip - s * aiq;
        A[i][q] = s * aip + c * aiq;
        const double wip = W[i][p], wiq = W[i][q];
        W[i][p] = c * wip - s * wiq;
        W[i][q] = s * wip + c * wiq;
      }
      __syncthreads();
      // rows: A <- J' A
      for (int t = tid; t < HP * B; t += EIG_THREADS) {
        const int pr = t / B, j = t % B;
        const double c = cs[3 * pr], s = cs[3 * pr + 1];
        const int pq = (int)cs[3 * pr + 2];
        const int p = pq / B, q = pq % B;
        const double apj = A[p][j], aqj = A[q][j];
        A[p][j] = c * apj - s * aqj;
        A[q][j] = s * apj + c * aqj;
      }
      __syncthreads();
    }
  }
}

template <int B>
struct EigSmem {
  double S[B][B + 1];
  double W[B][B + 1];
  double rowbuf[EIG_THREADS / 32][B];
  double cs[3 * (B / 2)];
  double red[16];
  double ev[B];
  int perm[B];
  int flag;
};

// Orthonormalise the columns of Z into Q: Cholesky-QR (Q = Z R^-1 with Z'Z = R'R); when a pivot collapses
// (rank-deficient block) the symmetric eigen-decomposition Z'Z = U L U' is used instead, Q = Z U L^-1/2,
// and the numerically null directions are zeroed.
template <int B>
__device__ void orth_pass(const double* __restrict__ src, double* __restrict__ dst, int G, EigSmem<B>& sm) {
  block_gram<B>(src, src, G, sm.S);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) sm.flag = 0;
  __syncthreads();
  if (tid < 32) {
    double maxd = 0.0;
    for (int i = lane; i < B; i += 32) maxd = fmax(maxd, sm.S[i][i]);
    for (int d = 16; d > 0; d >>= 1) maxd = fmax(maxd, __shfl_xor_sync(0xffffffffu, maxd, d));
    // copy S to W (Cholesky works on W so that S survives for the fallback)
    for (int t = lane; t < B * B; t += 32) sm.W[t / B][t % B] = sm.S[t / B][t % B];
    __syncwarp();
    bool ok = maxd > 0.0;
    for (int k = 0; k < B && ok; ++k) {
      const double piv = sm.W[k][k];
      if (!(piv > 1e-13 * maxd)) { ok = false; break; }
      const double dk = sqrt(piv);
      __syncwarp();
      for (int i = k + lane; i < B; i += 32) sm.W[i][k] = (i == k) ? dk : sm.W[i][k] / dk;   // L column k
      __syncwarp();
      for (int j = k + 1; j < B; ++j) {
        const double ljk = sm.W[j][k];
        for (int i = j + lane; i < B; i += 32) sm.W[i][j] -= sm.W[i][k] * ljk;
      }
      __syncwarp();
    }
    if (!ok) {
      if (lane == 0) sm.flag = 1;
    } else {
      // invert L (lower) in place column by column: X = L^-1, then the multiplier is R^-1 = X'
      // solve L X = I by forward substitution, lane per column
      for (int c = lane; c < B; c += 32) {
        for (int i = 0; i < B; ++i) {
          double v = (i == c) ? 1.0 : 0.0;
          for (int t = c; t < i; ++t) v -= sm.W[i][t] * sm.S[t][c];   // S reused as X (only rows >= c touched)
          sm.S[i][c] = (i < c) ? 0.0 : v / sm.W[i][i];
        }
      }
      __syncwarp();
      for (int t = lane; t < B * B; t += 32) sm.W[t / B][t % B] = 0.0;
      __syncwarp();
      for (int t = lane; t < B * B; t += 32) {
        const int i = t / B, j = t % B;
        sm.W[j][i] = sm.S[i][j];  // W = X' = R^-1
      }
    }
  }
  __syncthreads();
  if (sm.flag) {
    block_jacobi<B>(sm.S, sm.W, sm.cs, sm.red);
    double lmax = 0.0;
    for (int i = 0; i < B; ++i) lmax = fmax(lmax, sm.S[i][i]);
    for (int t = tid; t < B * B; t += EIG_THREADS) {
      const int i = t / B, j = t % B;
      const double l = sm.S[j][j];
      sm.W[i][j] = (l > 1e-20 * lmax && l > 0.0) ? sm.W[i][j] / sqrt(l) : 0.0;
    }
    __syncthreads();
  }
  rows_times_matrix<B>(src, dst, G, sm.W, sm.rowbuf);
}

template <int B>
__global__ void __launch_bounds__(EIG_THREADS) eig_orth_kernel(const EigWork* __restrict__ works) {
  extern __shared__ __align__(16) unsigned char raw[];
  EigSmem<B>& sm = *reinterpret_cast<EigSmem<B>*>(raw);
  const EigWork w = works[blockIdx.x];
  if (w.converged && *w.converged == 2) return;
  orth_pass<B>(w.Z, w.Q, w.G, sm);
  orth_pass<B>(w.Q, w.Q, w.G, sm);
}

// Rayleigh-Ritz: T = Q'Z (Z = C Q), eigen-decompose, rotate Q and Z to the Ritz basis (descending values),
// test stagnation of the leading r Ritz values.
template <int B>
__global__ void __launch_bounds__(EIG_THREADS) eig_rr_kernel(const EigWork* __restrict__ works, int r, double tol,
                                                             int finalize) {
  extern __shared__ __align__(16) unsigned char raw[];
  EigSmem<B>& sm = *reinterpret_cast<EigSmem<B>*>(raw);
  const EigWork w = works[blockIdx.x];
  if (*w.converged == 2) return;
  const int tid = threadIdx.x;
  block_gram<B>(w.Q, w.Z, w.G, sm.S);
  for (int t = tid; t < B * B; t += EIG_THREADS) {
    const int i = t / B, j = t % B;
    if (i < j) {
      const double v = 0.5 * (sm.S[i][j] + sm.S[j][i]);
      sm.S[i][j] = v;
      sm.S[j][i] = v;
    }
  }
  __syncthreads();
  block_jacobi<B>(sm.S, sm.W, sm.cs, sm.red);
  // sort descending (rank by counting; ties by index)
  if (tid < B) {
    const double v = sm.S[tid][tid];
    int rank = 0;
    for (int j = 0; j < B; ++j) {
      const double u = sm.S[j][j];
      rank += (u > v || (u == v && j < tid)) ? 1 : 0;
    }
    sm.perm[rank] = tid;
    sm.ev[rank] = v;
  }
  __syncthreads();
  // permute the columns of W: S reused as the permuted copy
  for (int t = tid; t < B * B; t += EIG_THREADS) {
    const int i = t / B, j = t % B;
    sm.S[i][j] = sm.W[i][sm.perm[j]];
  }
  __syncthreads();
  for (int t = tid; t < B * B; t += EIG_THREADS) sm.W[t / B][t % B] = sm.S[t / B][t % B];
  __syncthreads();
  if (tid < B) {
    w.theta[tid] = sm.ev[tid];
    sm.cs[tid] = 0.0;  // residual accumulators (cs holds 3*B/2 >= B doubles)
  }
  __syncthreads();
  rows_times_matrix<B>(w.Q, w.Q, w.G, sm.W, sm.rowbuf);
  rows_times_matrix<B>(w.Z, w.Z, w.G, sm.W, sm.rowbuf);
  // residuals ||C q_j - theta_j q_j||^2 of the leading r Ritz pairs
  {
    const int j = tid % B;
    double acc = 0.0;
    if (j < r) {
      const double th = sm.ev[j];
      for (int row = tid / B; row < w.G; row += EIG_THREADS / B) {
        const double d = w.Z[(size_t)row * B + j] - th * w.Q[(size_t)row * B + j];
        acc = fma(d, d, acc);
      }
      atomicAdd(&sm.cs[j], acc);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double worst = 0.0;
    for (int j = 0; j < r; ++j) worst = fmax(worst, sqrt(sm.cs[j]));
    const double scale = fmax(fabs(sm.ev[0]), 1e-300);
    *w.converged = finalize ? 2 : ((worst <= tol * scale) ? 1 : 0);
  }
  if (finalize) {
    for (size_t t = tid; t < (size_t)w.G * r; t += EIG_THREADS) {
      const int row = (int)(t % w.G), c = (int)(t / w.G);
      w.V[t] = w.Q[(size_t)row * B + c];
    }
    if (tid < r && w.theta_out) w.theta_out[tid] = sm.ev[tid];
  }
}

template <int B>
void run_eig(Context& ctx, const EigProblem* h_probs, int n_probs, int r, double tol, int max_iter) {
  cudaStream_t s = ctx.stream;
  std::vector<EigWork> works(n_probs);
  size_t tot = 0;
  int maxG = 0;
  for (int p = 0; p < n_probs; ++p) {
    tot += (size_t)h_probs[p].G * B;
    maxG = h_probs[p].G > maxG ? h_probs[p].G : maxG;
    CG_CHECK(h_probs[p].G >= B, "eigen block wider than the matrix");
  }
  DevBuf<double> Q(tot), Z(tot), theta((size_t)n_probs * B);
  DevBuf<int> conv((size_t)n_probs);
  theta.zero(s);
  conv.zero(s);
  size_t o = 0;
  for (int p = 0; p < n_probs; ++p) {
    works[p] = EigWork{h_probs[p].C, h_probs[p].G, Q.p + o, Z.p + o, theta.p + (size_t)p * B, h_probs[p].V,
                       h_probs[p].theta, conv.p + p};
    o += (size_t)h_probs[p].G * B;
  }
  DevBuf<EigWork> dw;
  dw.upload(works.data(), works.size(), s);
  const size_t smem = sizeof(EigSmem<B>);
  CG_CUDA(cudaFuncSetAttribute(eig_orth_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CG_CUDA(cudaFuncSetAttribute(eig_rr_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 gi((unsigned)(((size_t)maxG * B + 255) / 256), n_probs);
  dim3 gm((maxG + 7) / 8, n_probs);
  eig_init_kernel<B><<<gi, 256, 0, s>>>(dw.p);
  eig_orth_kernel<B><<<n_probs, EIG_THREADS, smem, s>>>(dw.p);
  ctx.stats.kernels += 2;
  std::vector<int> hc(n_probs);
  const int rr_every = 4;
  int it_done = 0;
  for (int it = 1; it <= max_iter; ++it) {
    it_done = it;
    {
      StageScope sc(ctx, "reduction.eig.matmul");
      eig_matmul_kernel<B><<<gm, 256, 0, s>>>(dw.p);
    }
    ctx.stats.kernels++;
    if (it % rr_every == 0) {
      {
        StageScope sc(ctx, "reduction.eig.rr");
        eig_rr_kernel<B><<<n_probs, EIG_THREADS, smem, s>>>(dw.p, r, tol, 0);
      }
      ctx.stats.kernels++;
      conv.download(hc.data(), (size_t)n_probs, s);
      CG_CUDA(cudaStreamSynchronize(s));
      bool all = true;
      for (int p = 0; p < n_probs; ++p) all = all && hc[p] != 0;
      if (all) break;
    }
    {
      StageScope sc(ctx, "reduction.eig.orth");
      eig_orth_kernel<B><<<n_probs, EIG_THREADS, smem, s>>>(dw.p);
    }
    ctx.stats.kernels++;
  }
  if (ctx.profile) {
    ctx.stage_ms["reduction.eig.iterations"] += it_done;
    int nconv = 0;
    for (int p = 0; p < n_probs; ++p) nconv += hc[p] != 0;
    ctx.stage_ms["reduction.eig.converged_problems"] += nconv;
    ctx.stage_ms["reduction.eig.problems"] += n_probs;
  }
  // final Rayleigh-Ritz from an orthonormal basis
  eig_orth_kernel<B><<<n_probs, EIG_THREADS, smem, s>>>(dw.p);
  eig_matmul_kernel<B><<<gm, 256, 0, s>>>(dw.p);
  eig_rr_kernel<B><<<n_probs, EIG_THREADS, smem, s>>>(dw.p, r, tol, 1);
  ctx.stats.kernels += 3;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

void top_eigenvectors_batched(Context& ctx, const EigProblem* h_probs, int n_probs, int r, double tol, int max_iter) {
  if (n_probs == 0) return;
  CG_CHECK(r >= 1 && r <= 56, "1..56 eigenvectors per problem");
  if (r <= 24) run_eig<32>(ctx, h_probs, n_probs, r, tol, max_iter);
  else run_eig<64>(ctx, h_probs, n_probs, r, tol, max_iter);
}

}  // namespace cg
