// K2L: exact cross-kNN for LARGE k (32 < k <= KNN_TCL_MAXK) on the sm_100a tensor cores, angular metric.
//
// alignment.strength > 0 makes buildGraph ask for k1 = round(a^2 * max cells) neighbours (R/conclass.R:190-196):
// hundreds to thousands per cell -- 9 % of the references at configs[4].  Neither a sorted list per row nor
// k group maxima per row are affordable there.  Same tensor-core tiles as knn_tc2.cu (fp16 operands, threshold in
// the spare K dimension, sign-bit masks), but the threshold comes from COUNTING:
//
//   count passes   every pass multiplies all tiles and counts, per query row, the references whose approximate
//                  score reaches the row's current threshold (popc of the sign masks).  The thresholds are fp16
//                  numbers; a bisection over their ordered integer representation keeps, per row, the largest probed
//                  threshold lo with count(score >= lo) >= k.  A row stops when its count lands in [k, k + slack];
//                  the CTA stops counting when all of its 256 rows have (at most 15 passes).
//   dump pass      thr = lo - 2*margin: at least k references reach lo, so the exact k-th best dot product is
//                  >= lo - eps and every true neighbour has approximate score >= lo - 2 eps.  Survivors are appended
//                  to the row's candidate list in HBM.
//   finish         one CTA per query row: canonical fp32 re-score from the exact row-major panels, 64-bit
//                  (dist, index) bitonic sort in shared memory, first k written, same certificate as knn_tc2.cu.
//                  Rows that overflow their list (masses of ties) mark the problem for the exact SIMT kernel.
//
// Work per problem: (passes + 1) * 2*nq*nr*d_pad tensor flops instead of an O(nq*nr*d) SIMT scan with a
// k-element selection per row: 25k x 25k, k = 2250: 1.24 s -> tens of ms.
#include <cuda_fp16.h>

#include "knn.cuh"
#include "ptx.cuh"

namespace cg {
namespace {

constexpr int TL_THREADS = 576;  // warp 0 producer, warp 1 MMA, warps 2-17 epilogue
constexpr int TILE_N = 128;
constexpr int QBLOCK = 256;
constexpr int MAX_COUNT_PASSES = 15;
constexpr float MMA_MARGIN = 1.1e-3f;  // see knn_tc2.cu

template <int D_PAD>
struct TLCfg {
  static constexpr int KC = D_PAD / 8;
  static constexpr int SWZ = D_PAD * 2;
  static constexpr int ROW_BYTES = D_PAD * 2;
  static constexpr int A_BYTES = QBLOCK * ROW_BYTES;
  static constexpr int B_BYTES = TILE_N * ROW_BYTES;
  static constexpr int NSTAGE = D_PAD == 32 ? 10 : 6;
  static constexpr int KSTEPS = D_PAD / 16;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + A_BYTES;
  static constexpr int OFF_CNT = OFF_B + NSTAGE * B_BYTES;   // int cnt[2 halves][256 rows]
  static constexpr int OFF_CTL = OFF_CNT + 2 * 256 * 4;      // int active, mode
  static constexpr int OFF_BAR = OFF_CTL + 16;
  static constexpr int N_BAR = 2 + 2 * NSTAGE + 4 + 1;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16;
};

// fp16 <-> order-preserving 16-bit integers
__device__ __forceinline__ int h_ord(__half h) {
  const unsigned short b = __half_as_ushort(h);
  return (b & 0x8000u) ? (int)(unsigned short)~b : (int)(b | 0x8000u);
}
__device__ __forceinline__ __half ord_h(int o) {
  const unsigned short b = (unsigned short)o;
  return __ushort_as_half((b & 0x8000u) ? (unsigned short)(b & 0x7fffu) : (unsigned short)~b);
}

template <int D_PAD, bool FOLD>
__global__ void __launch_bounds__(TL_THREADS, 1)
knn_tcl_kernel(const KnnProblem* __restrict__ probs, const KnnTclAux* __restrict__ auxs,
               const KnnItem* __restrict__ items, int n_items, int k, int caph, int slack) {
  using C = TLCfg<D_PAD>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* smA = reinterpret_cast<__half*>(smem + C::OFF_A);
  unsigned char* smB = smem + C::OFF_B;
  int* s_cnt = reinterpret_cast<int*>(smem + C::OFF_CNT);
  volatile int* s_ctl = reinterpret_cast<volatile int*>(smem + C::OFF_CTL);  // [0] rows still searching, [1] next pass: 1 = dump
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;
  uint64_t* b_empty = b_full + C::NSTAGE;
  uint64_t* acc_full = b_empty + C::NSTAGE;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* thr_ready = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int s = 0; s < C::NSTAGE; ++s) {
      ptx::mbar_init(b_full + s, 1);
      ptx::mbar_init(b_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(acc_full + s, 1);
      ptx::mbar_init(acc_empty + s, 16);
    }
    ptx::mbar_init(thr_ready, 16);
    s_ctl[0] = 0;
    s_ctl[1] = 0;
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
    // ------------------------------------------------------------ producer: one sweep over the tiles per pass
    if (lane == 0) {
      uint32_t slot = 0, ph = 0, it = 0, tp = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const KnnItem wi = items[item];
        const KnnProblem pb = probs[wi.prob];
        ptx::mbar_wait(a_empty, (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(a_full, C::A_BYTES);
        const __half* q_mma = reinterpret_cast<const __half*>(pb.q_tiled + round_up((size_t)pb.nq, KNN_ROW_PAD) * D_PAD);
        const __half* r_mma = reinterpret_cast<const __half*>(pb.r_tiled + round_up((size_t)pb.nr, KNN_ROW_PAD) * D_PAD);
        ptx::bulk_g2s(smA, q_mma + (size_t)wi.qblock * QBLOCK * D_PAD, C::A_BYTES, a_full);
        const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
        bool last = false;
        while (!last) {
          ptx::mbar_wait(thr_ready, tp & 1);   // the pass type is known once the thresholds are published
          ++tp;
          last = s_ctl[1] != 0;
          for (int t = 0; t < n_tiles; ++t) {
            ptx::mbar_wait(b_empty + slot, ph ^ 1);
            ptx::mbar_arrive_expect_tx(b_full + slot, C::B_BYTES);
            ptx::bulk_g2s(smB + (size_t)slot * C::B_BYTES, r_mma + (size_t)t * TILE_N * D_PAD, C::B_BYTES,
                          b_full + slot);
            if (++slot == C::NSTAGE) { slot = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0u, 128u, (uint32_t)TILE_N);  // fp16 x fp16 -> fp32
      uint64_t adesc[2][C::KSTEPS];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks)
          adesc[mb][ks] = ptx::umma_desc_kmajor_swz(ptx::smem_u32(smA) + mb * 128 * C::ROW_BYTES + ks * 32, C::SWZ,
                                                   8 * C::ROW_BYTES);
      const uint64_t bdesc0 = ptx::umma_desc_kmajor_swz(ptx::smem_u32(smB), C::SWZ, 8 * C::ROW_BYTES);
      const uint32_t b_full_a = ptx::smem_u32(b_full), b_empty_a = ptx::smem_u32(b_empty);
      const uint32_t acc_full_a = ptx::smem_u32(acc_full), acc_empty_a = ptx::smem_u32(acc_empty);
      uint32_t slot = 0, ph = 0, stage = 0, aph = 0, it = 0, tp = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const KnnItem wi = items[item];
        const int nr = probs[wi.prob].nr;
        const int n_tiles = (nr + TILE_N - 1) / TILE_N;
        ptx::mbar_wait(a_full, it & 1);
        bool last = false;
        while (!last) {
          ptx::mbar_wait(thr_ready, tp & 1);   // thresholds of this pass are in the A tile (FOLD)
          ++tp;
          last = s_ctl[1] != 0;
          for (int t = 0; t < n_tiles; ++t) {
            ptx::mbar_wait_a(b_full_a + slot * 8, ph);
            ptx::mbar_wait_a(acc_empty_a + stage * 8, aph ^ 1);
            ptx::tc_fence_after();
            const uint64_t bd = bdesc0 + (uint64_t)((slot * (uint32_t)C::B_BYTES) >> 4);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint32_t d_tmem = tmem_base + (stage * 2 + mb) * TILE_N;
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks)
                ptx::mma_f16_ss(d_tmem, adesc[mb][ks], bd + (uint64_t)(ks * 2), idesc, ks > 0 ? 1u : 0u);
            }
            ptx::mma_commit_a(b_empty_a + slot * 8);
            ptx::mma_commit_a(acc_full_a + stage * 8);
            if (++slot == C::NSTAGE) { slot = 0; ph ^= 1; }
            aph ^= stage;
            stage ^= 1;
          }
        }
        ptx::mma_commit(a_empty);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue warps: thread = (query row, 64-column half)
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int sub = e >> 2;
    const int mb = sub & 1, half = sub >> 1;
    const int row_local = mb * 128 + quarter * 32 + lane;
    const uint32_t tm_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t acc_full_a = ptx::smem_u32(acc_full), acc_empty_a = ptx::smem_u32(acc_empty);
    __half* thr_slot = smA + knn_tiled16_offset(row_local, D_PAD - 1, D_PAD);
    const bool first_thread = warp == 2 && lane == 0;
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const KnnItem wi = items[item];
      const KnnProblem pb = probs[wi.prob];
      const KnnTclAux ax = auxs[wi.prob];
      const int qrow = wi.qblock * QBLOCK + row_local;
      const bool valid_row = qrow < pb.nq;
      const uint32_t row_mask = valid_row ? 0xffffffffu : 0u;
      const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
      // bisection state over the ordered fp16 thresholds: count(score >= lo) >= k always (lo starts below every score)
      int lo = h_ord(__float2half(-1.5f)), hi = h_ord(__float2half(1.5f));
      bool done = !valid_row;
      int probe = (lo + hi + 1) >> 1;
      ptx::mbar_wait(a_full, it & 1);  // the query block has landed: the threshold column may be written
      int pass = 0;
      bool dump = false;
      while (true) {
        // ---- publish the thresholds of this pass
        float thr;
        {
          __half th;
          if (dump) {
            const float t2 = __half2float(ord_h(lo)) - 2.0f * MMA_MARGIN;
            th = __float2half_rd(valid_row ? t2 : 60000.0f);
          } else {
            th = done ? ord_h(lo) : ord_h(probe);
            if (!valid_row) th = __float2half(60000.0f);
          }
          thr = __half2float(th);
          if (FOLD) {
            *thr_slot = th;  // both halves of a row write the same value
            ptx::fence_proxy_async();
          }
          if (dump && half == 0 && valid_row) ax.thr[(size_t)ax.row0 + qrow] = thr;
          if (first_thread) s_ctl[1] = dump ? 1 : 0;
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(thr_ready);
        // ---- one sweep over the tiles: sign masks of (score - thr)
        int cnt = 0;
        int* cbase = ax.cand + ((size_t)(ax.row0 + qrow) * 2 + half) * (size_t)caph;
        for (int t = 0; t < n_tiles; ++t, ++g) {
          const uint32_t stage = g & 1, aph = (g >> 1) & 1;
          ptx::mbar_wait_a(acc_full_a + stage * 8, aph);
          ptx::tc_fence_after();
          const uint32_t taddr = tm_lane + (stage * 2 + mb) * TILE_N + half * 64;
          uint32_t v0[16], v1[16], v2[16], v3[16];
          ptx::tmem_ld_32x32b_x16(taddr, v0);
          ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
          ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
          ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
          ptx::tmem_wait_ld();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_a(acc_empty_a + stage * 8);
          uint32_t ma = 0, mb2 = 0;
          if (FOLD) {
#pragma unroll
            for (int c = 15; c >= 0; --c) ma = __funnelshift_l(v1[c], ma, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) ma = __funnelshift_l(v0[c], ma, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) mb2 = __funnelshift_l(v3[c], mb2, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) mb2 = __funnelshift_l(v2[c], mb2, 1);
            ma = ~ma;
            mb2 = ~mb2;
          } else {
            // sign bit of (v - thr) is clear iff v >= thr
#pragma unroll
            for (int c = 15; c >= 0; --c) ma = __funnelshift_l(__float_as_uint(__uint_as_float(v1[c]) - thr), ma, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) ma = __funnelshift_l(__float_as_uint(__uint_as_float(v0[c]) - thr), ma, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) mb2 = __funnelshift_l(__float_as_uint(__uint_as_float(v3[c]) - thr), mb2, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) mb2 = __funnelshift_l(__float_as_uint(__uint_as_float(v2[c]) - thr), mb2, 1);
            ma = ~ma;
            mb2 = ~mb2;
          }
          ma &= row_mask;
          mb2 &= row_mask;
          const int jb = t * TILE_N + half * 64;
          const int nv = pb.nr - jb;
          if (nv < 64) {
            ma &= nv >= 32 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << nv) - 1u));
            mb2 &= nv >= 64 ? 0xffffffffu : (nv <= 32 ? 0u : ((1u << (nv - 32)) - 1u));
          }
          if (!dump) {
            cnt += __popc(ma) + __popc(mb2);
          } else if (ma | mb2) {
            while (ma) {
              const int c = __ffs(ma) - 1;
              ma &= ma - 1;
              if (cnt < caph) cbase[cnt] = jb + c;
              ++cnt;
            }
            while (mb2) {
              const int c = __ffs(mb2) - 1;
              mb2 &= mb2 - 1;
              if (cnt < caph) cbase[cnt] = jb + 32 + c;
              ++cnt;
            }
          }
        }
        if (dump) {
          if (valid_row) ax.cnt[((size_t)ax.row0 + qrow) * 2 + half] = cnt;
          break;
        }
        // ---- bisection step (both halves of a row reach the same decision from the same numbers)
        s_cnt[half * 256 + row_local] = cnt;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (first_thread) s_ctl[0] = 0;
        const int total = s_cnt[row_local] + s_cnt[256 + row_local];
        if (!done) {
          if (total >= k) {
            lo = probe;
            if (total <= k + slack) done = true;
          } else {
            hi = probe;
          }
          if (hi - lo <= 1) done = true;
          probe = (lo + hi + 1) >> 1;
          if (probe >= hi) probe = hi - 1 > lo ? hi - 1 : lo;
        }
        ++pass;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (!done && half == 0) atomicAdd((int*)&s_ctl[0], 1);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        dump = s_ctl[0] == 0 || pass >= MAX_COUNT_PASSES;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// --------------------------------------------------------------------------------------- finish kernel
__device__ __forceinline__ uint32_t f32_orderable(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

constexpr int FIN_THREADS = 256;

// one CTA per query row; keys in shared memory (dynamic: 8 * np bytes)
template <int D_PAD>
__global__ void __launch_bounds__(FIN_THREADS)
knn_tcl_finish_kernel(const KnnProblem* __restrict__ probs, const KnnTclAux* __restrict__ auxs, int prob, int k, int caph,
                      int fold, int* __restrict__ prob_fail) {
  extern __shared__ __align__(16) unsigned long long keys[];
  __shared__ float qs[D_PAD];
  const KnnProblem pb = probs[prob];
  const KnnTclAux ax = auxs[prob];
  const int qrow = blockIdx.x;
  const int tid = threadIdx.x;
  const size_t gr = (size_t)ax.row0 + qrow;
  const int c0 = ax.cnt[gr * 2], c1 = ax.cnt[gr * 2 + 1];
  const float thr = ax.thr[gr];
  if (c0 > caph || c1 > caph) {
    if (tid == 0) *prob_fail = 1;
    return;
  }
  const int c = c0 + c1;
  int np = 2;
  while (np < c) np <<= 1;
  const int* cand = ax.cand + gr * 2 * (size_t)caph;
  const float* r_rows = pb.r_tiled + round_up((size_t)pb.nr, KNN_ROW_PAD) * D_PAD * 3 / 2;
  const float* q_rows = pb.q_tiled + round_up((size_t)pb.nq, KNN_ROW_PAD) * D_PAD * 3 / 2;
  if (tid < D_PAD) qs[tid] = q_rows[(size_t)qrow * D_PAD + tid];
  __syncthreads();
  constexpr int KC = D_PAD / 4;
  for (int s = tid; s < np; s += FIN_THREADS) {
    unsigned long long key = ~0ull;
    if (s < c) {
      const int j = s < c0 ? cand[s] : cand[caph + s - c0];
      const float* rp = r_rows + (size_t)j * D_PAD;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < KC; ++c4) {
        const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + c4 * 4));
        a0 = fmaf(qs[c4 * 4 + 0], rv.x, a0);
        a1 = fmaf(qs[c4 * 4 + 1], rv.y, a1);
        a2 = fmaf(qs[c4 * 4 + 2], rv.z, a2);
        if (!(c4 == KC - 1 && fold)) a3 = fmaf(qs[c4 * 4 + 3], rv.w, a3);
      }
      const float dist = 1.0f - ((a0 + a1) + (a2 + a3));
      key = ((unsigned long long)f32_orderable(dist) << 32) | (unsigned)j;
    }
    keys[s] = key;
  }
  __syncthreads();
  for (int size = 2; size <= np; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (np >> 1); t += FIN_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  // certificate (see knn_tc2.cu): no filtered-out reference can beat or tie the k-th candidate
  const unsigned long long kk = (k - 1 < np) ? keys[k - 1] : ~0ull;
  const float dist_k = kk == ~0ull ? INFINITY : f32_from_orderable((uint32_t)(kk >> 32));
  const bool ok = c >= k && dist_k <= (1.0f - thr) - MMA_MARGIN;
  if (!ok && tid == 0) *prob_fail = 1;
  for (int p = tid; p < k; p += FIN_THREADS) {
    const unsigned long long ky = p < np ? keys[p] : ~0ull;
    const bool has = ky != ~0ull;
    pb.out_idx[(size_t)qrow * k + p] = has ? (int)(ky & 0xffffffffull) : -1;
    pb.out_dist[(size_t)qrow * k + p] = has ? f32_from_orderable((uint32_t)(ky >> 32)) : INFINITY;
  }
}

template <int D_PAD, bool FOLD>
void launch_main(Context& ctx, const KnnProblem* d_probs, const KnnTclAux* d_aux, const KnnItem* d_items, int n_items,
                 int k, int caph, int slack) {
  using C = TLCfg<D_PAD>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  CG_CUDA(cudaFuncSetAttribute(knn_tcl_kernel<D_PAD, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               C::SMEM_BYTES));
  const int grid = n_items < ctx.sm_count ? n_items : ctx.sm_count;
  knn_tcl_kernel<D_PAD, FOLD><<<grid, TL_THREADS, C::SMEM_BYTES, ctx.stream>>>(d_probs, d_aux, d_items, n_items, k,
                                                                              caph, slack);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

}  // namespace

bool knn_tcl_applies(int nr, int k) { return k > KNN_TC_MAXK && k <= KNN_TCL_MAXK && nr >= k + k / 4 && nr >= 3 * TILE_N; }

void knn_tcl_run(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, const std::vector<int>& plist,
                 int n_probs, int k, int d_pad, int d, std::vector<int>& redo_probs) {
  const bool fold = d < d_pad;
  int cap = 256;  // candidate slots per row: a power of two >= 1.25 k + 128
  while (cap < k + k / 4 + 128) cap <<= 1;
  const int caph = cap / 2;
  const int slack = k / 16 > 16 ? k / 16 : 16;
  std::vector<KnnTclAux> aux((size_t)n_probs);
  const size_t byte_budget = (size_t)8 << 30;
  size_t p0 = 0;
  while (p0 < plist.size()) {
    size_t rows = 0, p1 = p0;
    std::vector<KnnItem> items;
    while (p1 < plist.size()) {
      const int p = plist[p1];
      const size_t r = round_up((size_t)h_probs[p].nq, QBLOCK);
      if (p1 > p0 && (rows + r) * (size_t)cap * 4 > byte_budget) break;
      aux[p].row0 = (long long)rows;
      rows += r;
      const int nb = (h_probs[p].nq + QBLOCK - 1) / QBLOCK;
      for (int b = 0; b < nb; ++b) items.push_back(KnnItem{p, b});
      ++p1;
    }
    DevBuf<int> cand(rows * (size_t)cap), cnt(rows * 2), fail(p1 - p0);
    DevBuf<float> thr(rows);
    CG_CUDA(cudaMemsetAsync(fail.p, 0, (p1 - p0) * sizeof(int), ctx.stream));
    for (size_t q = p0; q < p1; ++q) {
      KnnTclAux& a = aux[plist[q]];
      a.cand = cand.p;
      a.cnt = cnt.p;
      a.thr = thr.p;
    }
    DevBuf<KnnTclAux> d_aux;
    d_aux.upload(aux.data(), (size_t)n_probs, ctx.stream);
    DevBuf<KnnItem> d_items;
    d_items.upload(items.data(), items.size(), ctx.stream);
    const int ni = (int)items.size();
    const size_t fin_smem = (size_t)cap * sizeof(unsigned long long);
    if (d_pad == 32) {
      if (fold) launch_main<32, true>(ctx, d_probs, d_aux.p, d_items.p, ni, k, caph, slack);
      else launch_main<32, false>(ctx, d_probs, d_aux.p, d_items.p, ni, k, caph, slack);
      CG_CUDA(cudaFuncSetAttribute(knn_tcl_finish_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    } else {
      if (fold) launch_main<64, true>(ctx, d_probs, d_aux.p, d_items.p, ni, k, caph, slack);
      else launch_main<64, false>(ctx, d_probs, d_aux.p, d_items.p, ni, k, caph, slack);
      CG_CUDA(cudaFuncSetAttribute(knn_tcl_finish_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    }
    for (size_t q = p0; q < p1; ++q) {
      const int p = plist[q];
      if (h_probs[p].nq <= 0) continue;
      if (d_pad == 32)
        knn_tcl_finish_kernel<32><<<h_probs[p].nq, FIN_THREADS, fin_smem, ctx.stream>>>(d_probs, d_aux.p, p, k, caph,
                                                                                       fold ? 1 : 0, fail.p + (q - p0));
      else
        knn_tcl_finish_kernel<64><<<h_probs[p].nq, FIN_THREADS, fin_smem, ctx.stream>>>(d_probs, d_aux.p, p, k, caph,
                                                                                       fold ? 1 : 0, fail.p + (q - p0));
      ctx.stats.kernels++;
    }
    CG_CUDA(cudaGetLastError());
    std::vector<int> h_fail(p1 - p0);
    fail.download(h_fail.data(), p1 - p0, ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    for (size_t q = p0; q < p1; ++q)
      if (h_fail[q - p0]) {
        redo_probs.push_back(plist[q]);
        ctx.knn_redo_blocks += (uint64_t)((h_probs[plist[q]].nq + QBLOCK - 1) / QBLOCK);
      }
    p0 = p1;
  }
}

}  // namespace cg
