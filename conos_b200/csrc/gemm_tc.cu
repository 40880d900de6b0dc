#include <cuda_fp16.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace cg {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int G_THREADS = 192;  // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int G_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB per hi / lo tile
constexpr int B_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int G_OFF_BAR = G_STAGES * STAGE_BYTES;
constexpr int G_SMEM = G_OFF_BAR + 128;

struct GemmTile {
  int prob, mt, nt;
};

// K is consumed in chunks of pb.chunk_slabs 64-wide slabs (0: one chunk = everything).  Every chunk is accumulated in
// FP32 in its own TMEM accumulator (two accumulators, double buffered) and the chunks are summed in FP64 by the epilogue
// (C = chunk 0, C += chunk i; a tile belongs to one CTA, so the order is fixed and no atomics are needed).  The tensor
// core's FP32 accumulator rounds after every K=16 step: over the 50 000-cell contraction of a Gram matrix (9375 steps)
// that alone costs ~1e-5 relative -- enough to rotate the 15th eigenvector of a covariance with clustered
// eigenvalues by ~1e-2 and change 2 % of the mutual-neighbour edges; with 1024-cell chunks the FP32 part of every
// sum is 192 steps long.
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tn_bf16x3_kernel(const GemmProblem* __restrict__ probs, const GemmTile* __restrict__ tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_OFF_BAR);
  uint64_t* full = bars;                 // [2]
  uint64_t* empty = bars + G_STAGES;     // [2]
  uint64_t* acc_full = bars + 2 * G_STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * G_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmTile tl = tiles[blockIdx.x];
  const GemmProblem pb = probs[tl.prob];
  const int n_slabs = pb.k_pad / BK;
  const int chunk = (pb.chunk_slabs > 0 && pb.chunk_slabs < n_slabs) ? pb.chunk_slabs : n_slabs;
  const int n_chunks = (n_slabs + chunk - 1) / chunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      ptx::mbar_init(full + s, 1);
      ptx::mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(acc_full + s, 1);
      ptx::mbar_init(acc_empty + s, 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int sl = 0; sl < n_slabs; ++sl) {
        const int st = sl % G_STAGES, ph = (sl / G_STAGES) & 1;
        ptx::mbar_wait(empty + st, ph ^ 1);
        unsigned char* base = smem + st * STAGE_BYTES;
        const size_t a_off = ((size_t)sl * (size_t)(pb.m_pad >> 3) + (size_t)tl.mt * (BM / 8)) * 512;
        const size_t b_off = ((size_t)sl * (size_t)(pb.n_pad >> 3) + (size_t)tl.nt * (BN / 8)) * 512;
        ptx::mbar_arrive_expect_tx(full + st, STAGE_BYTES);
        ptx::bulk_g2s(base, pb.a_hi + a_off, A_BYTES, full + st);
        ptx::bulk_g2s(base + A_BYTES, pb.a_lo + a_off, A_BYTES, full + st);
        ptx::bulk_g2s(base + 2 * A_BYTES, pb.b_hi + b_off, B_BYTES, full + st);
        ptx::bulk_g2s(base + 2 * A_BYTES + B_BYTES, pb.b_lo + b_off, B_BYTES, full + st);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // operands are bf16 (fmt 1) or fp16 (fmt 0) hi + lo pairs; fp32 accumulate
      const uint32_t idesc = ptx::umma_idesc(pb.fp16 ? 0u : 1u, (uint32_t)BM, (uint32_t)BN);
      const uint32_t s0 = ptx::smem_u32(smem);
      int sl = 0;
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        ptx::mbar_wait(acc_empty + buf, ((ch >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        const int sl_end = sl + chunk < n_slabs ? sl + chunk : n_slabs;
        for (bool first = true; sl < sl_end; ++sl) {
          const int st = sl % G_STAGES, ph = (sl / G_STAGES) & 1;
          ptx::mbar_wait(full + st, ph);
          ptx::tc_fence_after();
          const uint32_t a_hi = s0 + st * STAGE_BYTES, a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t dah = ptx::umma_desc_kmajor_noswz(a_hi + ks * 256, 128, 1024);
            const uint64_t dal = ptx::umma_desc_kmajor_noswz(a_lo + ks * 256, 128, 1024);
            const uint64_t dbh = ptx::umma_desc_kmajor_noswz(b_hi + ks * 256, 128, 1024);
            const uint64_t dbl = ptx::umma_desc_kmajor_noswz(b_lo + ks * 256, 128, 1024);
            // small terms first
            ptx::mma_f16_ss(d_tmem, dal, dbh, idesc, first ? 0u : 1u);
            ptx::mma_f16_ss(d_tmem, dah, dbl, idesc, 1u);
            ptx::mma_f16_ss(d_tmem, dah, dbh, idesc, 1u);
            first = false;
          }
          ptx::mma_commit(empty + st);
        }
        ptx::mma_commit(acc_full + buf);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int m = tl.mt * BM + quarter * 32 + lane;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int buf = ch & 1;
      ptx::mbar_wait(acc_full + buf, (ch >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
        ptx::tmem_wait_ld();
        if (m < pb.M) {
          const int n0 = tl.nt * BN + c * 32;
          if (pb.out_fp64) {
            double* dst = reinterpret_cast<double*>(pb.C) + (size_t)n0 * pb.ldc + m;
            // all 32 partial sums of the previous chunks are fetched before the first store: the compiler cannot
            // prove that the 32 addresses (runtime ldc apart) are distinct and would otherwise serialise the round trips
            double old[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = (ch > 0 && n0 + j < pb.N) ? __ldcg(dst + (size_t)j * pb.ldc) : 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < pb.N) __stcg(dst + (size_t)j * pb.ldc, old[j] + (double)__uint_as_float(v[j]));
          } else {
            float* dst = reinterpret_cast<float*>(pb.C) + (size_t)n0 * pb.ldc + m;
            float old[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = (ch > 0 && n0 + j < pb.N) ? __ldcg(dst + (size_t)j * pb.ldc) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < pb.N) __stcg(dst + (size_t)j * pb.ldc, old[j] + __uint_as_float(v[j]));
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty + buf);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ operand builders
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// fp16 hi + lo pair in the same 16-bit containers (kind::f16 takes either format): 11 + 11 significant bits instead
// of 8 + 8, for operands whose range fits fp16 (the expression values of the Gram operand: 0 .. ~6)
__device__ __forceinline__ void split_fp16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(x - __half2float(h));
  hi = *reinterpret_cast<const __nv_bfloat16*>(&h);
  lo = *reinterpret_cast<const __nv_bfloat16*>(&l);
}

__global__ void tiled_from_csc_kernel(const int* __restrict__ csc_p, const int* __restrict__ csc_i,
                                      const double* __restrict__ csc_x, int n_genes, int rows_pad,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  // one warp per gene column
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= n_genes) return;
  const int lane = threadIdx.x & 31;
  for (int t = csc_p[g] + lane; t < csc_p[g + 1]; t += 32) {
    const size_t o = tiled_operand_offset(g, csc_i[t], rows_pad);
    __nv_bfloat16 h, l;
    split_bf16((float)csc_x[t], h, l);
    hi[o] = h;
    lo[o] = l;
  }
}

// Same operand from the CSR (by cell) form, built tile by tile in shared memory: a CTA owns 64 cells (one k-block) x
// 256 genes, whose image in the operand layout is 32 KB contiguous; the non-zeros of its 64 cell rows that fall into
// the gene range are dropped into a zeroed shared-memory image, which then leaves with coalesced 16-byte stores.
// Every element of the padded operand is written exactly once (no memset, no scattered 2-byte stores).
__global__ void __launch_bounds__(256) tiled_from_csr_kernel(const int* __restrict__ csr_p, const int* __restrict__ csr_i,
                                                             const double* __restrict__ csr_x, int n_cells, int rows_pad,
                                                             __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, int fp16) {
  extern __shared__ __align__(16) unsigned char csr_smem[];
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(csr_smem);
  __nv_bfloat16* sl = sh + 256 * 64;
  const int kb = blockIdx.x, g0 = blockIdx.y * 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int t = tid; t < 256 * 64 / 8; t += 256) {
    reinterpret_cast<uint4*>(sh)[t] = z;
    reinterpret_cast<uint4*>(sl)[t] = z;
  }
  __syncthreads();
  for (int cl = warp; cl < 64; cl += 8) {
    const int c = kb * 64 + cl;
    if (c >= n_cells) break;
    const int b = csr_p[c], e = csr_p[c + 1];
    int lo_i = b, hi_i = e;  // first entry with column >= g0 (columns ascending inside a row)
    while (lo_i < hi_i) {
      const int mid = (lo_i + hi_i) >> 1;
      if (csr_i[mid] < g0) lo_i = mid + 1; else hi_i = mid;
    }
    for (int t0 = lo_i; t0 < e; t0 += 32) {
      const int t = t0 + lane;
      const int col = t < e ? csr_i[t] : 0x7fffffff;
      if (col < g0 + 256) {
        const int gl = col - g0;
        __nv_bfloat16 h, l;
        if (fp16) split_fp16((float)csr_x[t], h, l); else split_bf16((float)csr_x[t], h, l);
        const int o = ((gl >> 3) * 8 + (cl >> 3)) * 64 + (gl & 7) * 8 + (cl & 7);
        sh[o] = h;
        sl[o] = l;
      }
      if (__any_sync(0xffffffffu, col >= g0 + 256)) break;
    }
  }
  __syncthreads();
  const size_t base = ((size_t)kb * (size_t)(rows_pad >> 3) + (size_t)(g0 >> 3)) * 512;
  for (int t = tid; t < 256 * 64 / 8; t += 256) {
    reinterpret_cast<uint4*>(hi + base)[t] = reinterpret_cast<const uint4*>(sh)[t];
    reinterpret_cast<uint4*>(lo + base)[t] = reinterpret_cast<const uint4*>(sl)[t];
  }
}

template <class T>
__global__ void tiled_from_colmajor_kernel(const T* __restrict__ src, int k, int rows, long long ld, int rows_pad,
                                           int k_pad, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  // one thread per padded element, k fastest (coalesced reads)
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)rows_pad * k_pad) return;
  const int kk = (int)(t % k_pad), r = (int)(t / k_pad);
  float x = 0.f;
  if (kk < k && r < rows) x = (float)src[(size_t)r * ld + kk];
  __nv_bfloat16 h, l;
  split_bf16(x, h, l);
  const size_t o = tiled_operand_offset(r, kk, rows_pad);
  hi[o] = h;
  lo[o] = l;
}

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace

void gemm_plan(Context& ctx, const GemmProblem* h_probs, int n_probs, GemmPlan& plan) {
  static_assert(sizeof(GemmTileRef) == sizeof(GemmTile), "tile table layout");
  plan.n_tiles = 0;
  if (n_probs == 0) return;
  std::vector<GemmTile> tiles;
  for (int p = 0; p < n_probs; ++p) {
    const GemmProblem& g = h_probs[p];
    CG_CHECK(g.m_pad % GEMM_ROW_PAD == 0 && g.n_pad % GEMM_ROW_PAD == 0 && g.k_pad % GEMM_K_PAD == 0 && g.k_pad > 0,
             "GEMM operands must be padded (rows to 256, k to 64)");
    const int mts = (g.M + BM - 1) / BM, nts = (g.N + BN - 1) / BN;
    for (int nt = 0; nt < nts; ++nt)
      for (int mt = 0; mt < mts; ++mt) {
        if (g.symmetric && mt * BM > nt * BN + BN - 1) continue;  // strictly below the diagonal
        tiles.push_back(GemmTile{p, mt, nt});
      }
  }
  if (tiles.empty()) return;
  plan.probs.upload(h_probs, (size_t)n_probs, ctx.stream);
  plan.tiles.upload(reinterpret_cast<const GemmTileRef*>(tiles.data()), tiles.size(), ctx.stream);
  plan.n_tiles = (unsigned)tiles.size();
  CG_CUDA(cudaFuncSetAttribute(gemm_tn_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
}

void gemm_run(Context& ctx, const GemmPlan& plan) {
  if (plan.n_tiles == 0) return;
  gemm_tn_bf16x3_kernel<<<plan.n_tiles, G_THREADS, G_SMEM, ctx.stream>>>(
      plan.probs.p, reinterpret_cast<const GemmTile*>(plan.tiles.p));
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

void gemm_tn_bf16x3(Context& ctx, const GemmProblem* h_probs, int n_probs) {
  GemmPlan plan;
  gemm_plan(ctx, h_probs, n_probs, plan);
  gemm_run(ctx, plan);   // the plan's device tables are released in stream order
}

void tiled_from_csc(Context& ctx, const int* csc_p, const int* csc_i, const double* csc_x, int n_cells, int n_genes,
                    TiledOperand& op) {
  op.rows = n_genes;
  op.k = n_cells;
  op.rows_pad = (int)round_up(n_genes, GEMM_ROW_PAD);
  op.k_pad = (int)round_up(n_cells, GEMM_K_PAD);
  const size_t n = (size_t)op.rows_pad * op.k_pad;
  CG_CUDA(cudaMemsetAsync(op.hi, 0, n * sizeof(__nv_bfloat16), ctx.stream));
  CG_CUDA(cudaMemsetAsync(op.lo, 0, n * sizeof(__nv_bfloat16), ctx.stream));
  tiled_from_csc_kernel<<<(n_genes + 7) / 8, 256, 0, ctx.stream>>>(csc_p, csc_i, csc_x, n_genes, op.rows_pad, op.hi,
                                                                  op.lo);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

void tiled_from_csr(Context& ctx, const int* csr_p, const int* csr_i, const double* csr_x, int n_cells, int n_genes,
                    TiledOperand& op, bool fp16) {
  op.rows = n_genes;
  op.k = n_cells;
  op.rows_pad = (int)round_up(n_genes, GEMM_ROW_PAD);
  op.k_pad = (int)round_up(n_cells, GEMM_K_PAD);
  static_assert(GEMM_ROW_PAD % 256 == 0 && GEMM_K_PAD == 64, "tile shape of tiled_from_csr_kernel");
  dim3 grid((unsigned)(op.k_pad / 64), (unsigned)(op.rows_pad / 256));
  constexpr int smem = 2 * 256 * 64 * (int)sizeof(__nv_bfloat16);
  CG_CUDA(cudaFuncSetAttribute(tiled_from_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tiled_from_csr_kernel<<<grid, 256, smem, ctx.stream>>>(csr_p, csr_i, csr_x, n_cells, op.rows_pad, op.hi, op.lo,
                                                         fp16 ? 1 : 0);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

void tiled_from_colmajor_f64(Context& ctx, const double* src, int k, int rows, long long ld, TiledOperand& op) {
  op.rows = rows;
  op.k = k;
  op.rows_pad = (int)round_up(rows, GEMM_ROW_PAD);
  op.k_pad = (int)round_up(k, GEMM_K_PAD);
  const size_t n = (size_t)op.rows_pad * op.k_pad;
  tiled_from_colmajor_kernel<double><<<nblk(n), 256, 0, ctx.stream>>>(src, k, rows, ld, op.rows_pad, op.k_pad, op.hi,
                                                                      op.lo);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

void tiled_from_colmajor_f32(Context& ctx, const float* src, int k, int rows, long long ld, TiledOperand& op) {
  op.rows = rows;
  op.k = k;
  op.rows_pad = (int)round_up(rows, GEMM_ROW_PAD);
  op.k_pad = (int)round_up(k, GEMM_K_PAD);
  const size_t n = (size_t)op.rows_pad * op.k_pad;
  tiled_from_colmajor_kernel<float><<<nblk(n), 256, 0, ctx.stream>>>(src, k, rows, ld, op.rows_pad, op.k_pad, op.hi,
                                                                     op.lo);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

}  // namespace cg
