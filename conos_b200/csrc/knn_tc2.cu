// K2 (two-phase) exact cross-kNN on the sm_100a tensor cores (angular metric, k <= 32).
//
// The streaming top-k of knn_tc.cu pays k*ln(n/k) list insertions per query and couples the tensor pipe to
// the slowest selection warp.  This kernel removes the list from the inner loop:
//
//   phase A  every (query row, reference group) gets the MAXIMUM approximate (TF32) dot product of the
//            group -- one FMNMX3 per two candidates.  A group is a set of reference columns (a 64-column
//            half of one or several 128-row tiles, or a 16/32-column piece of it when there are few tiles);
//            the groups partition the references.  tau = the k-th largest group maximum of a row: at least
//            k references (one per group) have approximate score >= tau, hence the exact k-th best dot
//            product is >= tau - eps, hence every true neighbour has approximate score >= tau - 2*eps.
//   phase B  the same tensor-core products again, now with thr = tau - 2*margin riding in the spare K
//            dimension (acc = dot - thr): ONE funnel shift per candidate collects the sign bits, and the few
//            survivors (~1.2 k per row) are appended to the row's candidate list in HBM.  No feedback.
//   finish   (separate kernel, one warp per query row) canonical fp32 re-score of the candidates from the
//            tiled panels, 64-bit (dist, index) bitonic sort, first k written.  A row whose candidate list
//            overflowed (ties / duplicates) or whose certificate  dist_k <= (1 - thr) - margin  fails marks
//            its 256-row block; knn_tc.cu's streaming kernel recomputes such blocks exactly.
//
// Roles per CTA (persistent, one per SM, 576 threads): warp 0 producer (1-D bulk async copies of the query
// block and of the reference tiles, streamed twice), warp 1 MMA issuer (tcgen05.mma kind::tf32, two
// M=128 x N=128 accumulators per tile, double buffered = 512 TMEM columns), warps 2-17 epilogue: thread =
// (query row, 64-column half), tcgen05.ld 32x32b.x16.
#include <algorithm>
#include <cuda_fp16.h>

#include "knn.cuh"
#include "ptx.cuh"

// Measurement builds (never shipped): -DTC2_MMA_ONLY drops the accumulators, -DTC2_LOAD_ONLY drains them without
// arithmetic; with TC2_MMA_ONLY, -DTC2_EXP=1 lets ONE epilogue warp do the hand-shake, 3 removes the reference-tile
// ring, 4 removes the ring and the consumer hand-shake (pure MMA issue).  Results: DESIGN.md K2.
#ifndef TC2_EXP
#define TC2_EXP 0
#endif

namespace cg {
namespace {

constexpr int T2_THREADS = 576;
constexpr int TILE_N = 128;
constexpr int QBLOCK = 256;
constexpr int NGT_MAX = 48;           // groups per (row, half) kept in shared memory
// Both operands are rounded to nearest fp16 (11 significant bits, the precision of TF32 at twice the tensor rate and
// half the bytes; components of unit vectors never overflow, and below 2^-14 the absolute error is < 2^-25): the
// products are exact in the tensor core and the accumulation is fp32, so for unit vectors
// |approximate dot - canonical fp32 dot| <= (2*2^-11 + 2^-22) * sum|q_i r_i| + d*2^-24 + accumulation < 9.9e-4.
constexpr float MMA_MARGIN = 1.1e-3f;

template <int D_PAD>
struct T2Cfg {
  // operands are the fp16 copies of the panels: rows of D_PAD halves, 16-byte chunks swizzled (knn_tiled16_offset)
  static constexpr int KC = D_PAD / 8;              // 16-byte chunks per row
  static constexpr int SWZ = D_PAD * 2;             // swizzle span = row bytes: 64 or 128
  static constexpr int ROW_BYTES = D_PAD * 2;
  static constexpr int A_BYTES = QBLOCK * ROW_BYTES;
  static constexpr int B_BYTES = TILE_N * ROW_BYTES;
  static constexpr int NSTAGE = D_PAD == 32 ? 10 : 6;  // reference tiles in flight: covers the L2 -> shared latency
  static constexpr int KSTEPS = D_PAD / 16;         // kind::f16 MMA K = 16
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + A_BYTES;
  static constexpr int OFF_G = OFF_B + NSTAGE * B_BYTES;       // group maxima [NGT_MAX][2 halves][256 rows]
  static constexpr int OFF_BAR = OFF_G + NGT_MAX * 512 * 4;
  static constexpr int N_BAR = 2 + 2 * NSTAGE + 4 + 1;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16;
};

__device__ __forceinline__ float max16(const uint32_t (&v)[16]) {
  float m = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
  for (int c = 2; c < 16; c += 2) m = fmaxf(fmaxf(m, __uint_as_float(v[c])), __uint_as_float(v[c + 1]));
  return m;
}

// 64 accumulator columns of this thread's lane: four x16 loads or (LD32) two x32 loads
template <bool LD32>
__device__ __forceinline__ void load_scores(uint32_t taddr, uint32_t (&v0)[16], uint32_t (&v1)[16],
                                            uint32_t (&v2)[16], uint32_t (&v3)[16]) {
  if constexpr (LD32) {
    uint32_t a[32], b[32];
    ptx::tmem_ld_32x32b_x32(taddr, a);
    ptx::tmem_ld_32x32b_x32(taddr + 32, b);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      v0[c] = a[c];
      v1[c] = a[16 + c];
      v2[c] = b[c];
      v3[c] = b[16 + c];
    }
  } else {
    ptx::tmem_ld_32x32b_x16(taddr, v0);
    ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
    ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
    ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
    ptx::tmem_wait_ld();
  }
}

template <int D_PAD, bool FOLD, bool LD32>
__global__ void __launch_bounds__(T2_THREADS, 1)
knn_tc2_kernel(const KnnProblem* __restrict__ probs, const KnnTc2Aux* __restrict__ auxs,
               const KnnItem* __restrict__ items, int n_items, int k, int cap, int stagger) {
  using C = T2Cfg<D_PAD>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* smA = reinterpret_cast<__half*>(smem + C::OFF_A);
  unsigned char* smB = smem + C::OFF_B;
  float* gmax_s = reinterpret_cast<float*>(smem + C::OFF_G);
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
      ptx::mbar_init(acc_empty + s, TC2_EXP == 1 ? 1 : 16);
    }
    ptx::mbar_init(thr_ready, 16);
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
    // ------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t slot = 0, ph = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const KnnItem wi = items[item];
        const KnnProblem pb = probs[wi.prob];
        ptx::mbar_wait(a_empty, (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(a_full, C::A_BYTES);
        // both operands come from the fp16 copies of the panels (stored right after the fp32 values)
        const __half* q_mma = reinterpret_cast<const __half*>(pb.q_tiled + round_up((size_t)pb.nq, KNN_ROW_PAD) * D_PAD);
        const __half* r_mma = reinterpret_cast<const __half*>(pb.r_tiled + round_up((size_t)pb.nr, KNN_ROW_PAD) * D_PAD);
        ptx::bulk_g2s(smA, q_mma + (size_t)wi.qblock * QBLOCK * D_PAD, C::A_BYTES, a_full);
        const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
        for (int ph2 = 0; ph2 < 2; ++ph2) {
          if (TC2_EXP >= 3) break;
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
      // descriptors are built once: the loop only adds the (16-byte granular) offset of the ring slot
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
      uint32_t slot = 0, ph = 0, stage = 0, aph = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const KnnItem wi = items[item];
        const int nr = probs[wi.prob].nr;
        const int n_tiles = (nr + TILE_N - 1) / TILE_N;
        ptx::mbar_wait(a_full, it & 1);
        for (int ph2 = 0; ph2 < 2; ++ph2) {
          if (ph2 == 1) ptx::mbar_wait(thr_ready, it & 1);  // the rows' thresholds are in the A tile (FOLD)
          for (int t = 0; t < n_tiles; ++t) {
            if (TC2_EXP < 3) ptx::mbar_wait_a(b_full_a + slot * 8, ph);
            if (TC2_EXP != 4) ptx::mbar_wait_a(acc_empty_a + stage * 8, aph ^ 1);
            ptx::tc_fence_after();
            const uint64_t bd = bdesc0 + (uint64_t)((slot * (uint32_t)C::B_BYTES) >> 4);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint32_t d_tmem = tmem_base + (stage * 2 + mb) * TILE_N;
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks) {
                ptx::mma_f16_ss(d_tmem, adesc[mb][ks], bd + (uint64_t)(ks * 2), idesc, ks > 0 ? 1u : 0u);
              }
            }
            if (TC2_EXP < 3) ptx::mma_commit_a(b_empty_a + slot * 8);
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
    // ------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int sub = e >> 2;
    const int mb = sub & 1, half = sub >> 1;
    const int row_local = mb * 128 + quarter * 32 + lane;
    const uint32_t tm_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t acc_full_a = ptx::smem_u32(acc_full), acc_empty_a = ptx::smem_u32(acc_empty);
    float* gm = gmax_s + half * 256 + row_local;  // + g * 512
    // this row's threshold slot in the A tile: last K dimension
    __half* thr_slot = smA + knn_tiled16_offset(row_local, D_PAD - 1, D_PAD);
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const KnnItem wi = items[item];
      const KnnProblem pb = probs[wi.prob];
      const KnnTc2Aux ax = auxs[wi.prob];
      const int qrow = wi.qblock * QBLOCK + row_local;
      const bool valid_row = qrow < pb.nq;
      const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
      // The 16 epilogue warps wait on the same barriers and would move in lock step: all of them drain tensor memory
      // (ALU idle), then all of them do arithmetic (TMEM port idle).  The warps of the second column half start each
      // phase `stagger` cycles late and stay behind (the accumulators are ready long before they are consumed), so
      // one half drains while the other computes.
      auto hold_back = [&]() {
        if (stagger > 0 && half == 1) {
          const long long t0 = clock64();
          while (clock64() - t0 < stagger) {
          }
        }
      };
      hold_back();
      // ------------------------------------------------ phase A: group maxima of the approximate scores
      {
        float run = -INFINITY;
        int tcount = 0;
        for (int t = 0; t < n_tiles; ++t, ++g) {
          if (TC2_EXP == 4 || (TC2_EXP == 1 && e != 0)) break;
          const uint32_t stage = g & 1, aph = (g >> 1) & 1;
          ptx::mbar_wait_a(acc_full_a + stage * 8, aph);
          ptx::tc_fence_after();
#ifdef TC2_MMA_ONLY  // measurement build: tensor pipe + pipeline only, accumulators are dropped
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_a(acc_empty_a + stage * 8);
          continue;
#endif
          const uint32_t taddr = tm_lane + (stage * 2 + mb) * TILE_N + half * 64;
          uint32_t v0[16], v1[16], v2[16], v3[16];
          load_scores<LD32>(taddr, v0, v1, v2, v3);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_a(acc_empty_a + stage * 8);
#ifdef TC2_LOAD_ONLY  // measurement build: tensor pipe + hand-shake + tensor-memory drain, no arithmetic
          if (v0[0] == 0x7fc12345u) run = __uint_as_float(v0[1] ^ v1[2] ^ v2[3] ^ v3[4]);  // keeps the loads alive
          continue;
#endif
          const int nv = pb.nr - t * TILE_N - half * 64;  // valid columns of this thread's 64
          if (nv < 64) {  // zero padding rows of the last reference tile never count
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if (c >= nv) v0[c] = 0xff800000u;
              if (16 + c >= nv) v1[c] = 0xff800000u;
              if (32 + c >= nv) v2[c] = 0xff800000u;
              if (48 + c >= nv) v3[c] = 0xff800000u;
            }
          }
          const float m0 = max16(v0), m1 = max16(v1), m2 = max16(v2), m3 = max16(v3);
          if (ax.sg == 16) {
            gm[(4 * t + 0) * 512] = m0;
            gm[(4 * t + 1) * 512] = m1;
            gm[(4 * t + 2) * 512] = m2;
            gm[(4 * t + 3) * 512] = m3;
          } else if (ax.sg == 32) {
            gm[(2 * t + 0) * 512] = fmaxf(m0, m1);
            gm[(2 * t + 1) * 512] = fmaxf(m2, m3);
          } else {
            run = fmaxf(fmaxf(run, fmaxf(m0, m1)), fmaxf(m2, m3));
            if (++tcount == ax.tg || t == n_tiles - 1) {
              gm[(t / ax.tg) * 512] = run;
              run = -INFINITY;
              tcount = 0;
            }
          }
        }
      }
      // every epilogue warp has written its maxima
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // ------------------------------------------------ threshold: tau with count(group max >= tau) >= k
      float thr;
      {
        const int ng2 = 2 * ax.ngt;
        const float* col = gmax_s + row_local;  // + g2 * 256, g2 = g * 2 + half
        float lo = INFINITY, hi = -INFINITY;
        for (int q = 0; q < ng2; ++q) {
          const float v = col[q * 256];
          hi = fmaxf(hi, v);
          lo = v > -INFINITY ? fminf(lo, v) : lo;
        }
        // bisection: count(>= lo) >= k always (the host guarantees >= k non-empty groups)
#pragma unroll 1
        for (int iter = 0; iter < 12; ++iter) {
          const float mid = 0.5f * (lo + hi);
          int cge = 0;
          for (int q = 0; q < ng2; ++q) cge += col[q * 256] >= mid ? 1 : 0;
          if (cge >= k) lo = mid; else hi = mid;
        }
        // phase A scores carry +1 when the spare dimension is in use (query -1 times reference -1)
        thr = (lo - (FOLD ? 1.0f : 0.0f)) - 2.0f * MMA_MARGIN;
        const __half thr_h = __float2half_rd(valid_row ? thr : 60000.0f);  // toward -inf to an fp16 value
        thr = __half2float(thr_h);
        if (FOLD) {
          // the threshold rides in the spare K dimension of the A tile: acc = dot - thr from now on
          *thr_slot = thr_h;  // both halves of a row write the same value
          ptx::fence_proxy_async();
        }
        if (half == 0 && valid_row) ax.thr[(size_t)ax.row0 + qrow] = thr;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(thr_ready);
      hold_back();
      // ------------------------------------------------ phase B: filter, survivors to the candidate lists
      {
        const uint32_t row_mask = valid_row ? 0xffffffffu : 0u;
        int* cbase = ax.cand + ((size_t)(ax.row0 + qrow) * 2 + half) * KNN_TC2_CAND_STRIDE;
        int cnt = 0;
        for (int t = 0; t < n_tiles; ++t, ++g) {
          if (TC2_EXP == 4 || (TC2_EXP == 1 && e != 0)) break;
          const uint32_t stage = g & 1, aph = (g >> 1) & 1;
          ptx::mbar_wait_a(acc_full_a + stage * 8, aph);
          ptx::tc_fence_after();
#ifdef TC2_MMA_ONLY
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_a(acc_empty_a + stage * 8);
          continue;
#endif
          const uint32_t taddr = tm_lane + (stage * 2 + mb) * TILE_N + half * 64;
          uint32_t v0[16], v1[16], v2[16], v3[16];
          load_scores<LD32>(taddr, v0, v1, v2, v3);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_a(acc_empty_a + stage * 8);
#ifdef TC2_LOAD_ONLY
          if (v0[0] == 0x7fc12345u) cnt += (int)(v0[1] ^ v1[2] ^ v2[3] ^ v3[4]);
          continue;
#endif
          uint32_t ma = 0, mb2 = 0;
          if (FOLD) {
            // accumulator = dot - thr: collect the sign bits (highest column first), survivors = NOT sign
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
            // sign bit of (thr - v) is set iff v > thr
#pragma unroll
            for (int c = 15; c >= 0; --c) ma = __funnelshift_l(__float_as_uint(thr - __uint_as_float(v1[c])), ma, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) ma = __funnelshift_l(__float_as_uint(thr - __uint_as_float(v0[c])), ma, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) mb2 = __funnelshift_l(__float_as_uint(thr - __uint_as_float(v3[c])), mb2, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) mb2 = __funnelshift_l(__float_as_uint(thr - __uint_as_float(v2[c])), mb2, 1);
          }
          ma &= row_mask;
          mb2 &= row_mask;
          const int jb = t * TILE_N + half * 64;
          const int nv = pb.nr - jb;
          if (nv < 64) {
            ma &= nv >= 32 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << nv) - 1u));
            mb2 &= nv >= 64 ? 0xffffffffu : (nv <= 32 ? 0u : ((1u << (nv - 32)) - 1u));
          }
          if (ma | mb2) {
            while (ma) {
              const int c = __ffs(ma) - 1;
              ma &= ma - 1;
              if (cnt < cap) cbase[cnt] = jb + c;
              ++cnt;
            }
            while (mb2) {
              const int c = __ffs(mb2) - 1;
              mb2 &= mb2 - 1;
              if (cnt < cap) cbase[cnt] = jb + 32 + c;
              ++cnt;
            }
          }
        }
        if (valid_row) ax.cnt[((size_t)ax.row0 + qrow) * 2 + half] = cnt;
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

// ascending bitonic sort of 32*U keys, position p = u*32 + lane
template <int U>
__device__ __forceinline__ void warp_sort_u64(unsigned long long (&key)[U], int lane) {
#pragma unroll
  for (int kk = 2; kk <= 32 * U; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int ju = j >> 5;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pu = u ^ ju;
          if (pu > u) {
            const bool asc = ((u * 32) & kk) == 0;
            const unsigned long long a = key[u], b = key[pu];
            const bool sw = (a > b) == asc;
            key[u] = sw ? b : a;
            key[pu] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[u], j);
          const bool asc = (((u * 32) | lane) & kk) == 0;
          const bool lower = (lane & j) == 0;
          const bool take_min = lower == asc;
          const bool other_less = other < key[u];
          key[u] = (take_min == other_less) ? other : key[u];
        }
      }
    }
  }
}

template <int D_PAD, int U>
__device__ __forceinline__ void finish_row(const KnnProblem& pb, const int* __restrict__ cand, int c0, int c1,
                                           int qrow, int k, bool fold, float thr, int lane, bool& ok) {
  constexpr int KC = D_PAD / 4;
  const int c = c0 + c1;
  int j[U];
  float a0[U], a1[U], a2[U], a3[U];
  const float* rp[U];
  // whole rows from the row-major exact copies of the panels
  const float* r_rows = pb.r_tiled + round_up((size_t)pb.nr, KNN_ROW_PAD) * D_PAD * 3 / 2;
  const float* q_rows = pb.q_tiled + round_up((size_t)pb.nq, KNN_ROW_PAD) * D_PAD * 3 / 2;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int s = u * 32 + lane;
    j[u] = s < c ? (s < c0 ? cand[s] : cand[KNN_TC2_CAND_STRIDE + s - c0]) : -1;
    const int jj = j[u] >= 0 ? j[u] : 0;
    rp[u] = r_rows + (size_t)jj * D_PAD;
    a0[u] = a1[u] = a2[u] = a3[u] = 0.f;
  }
  const float* qp = q_rows + (size_t)qrow * D_PAD;
#pragma unroll
  for (int c4 = 0; c4 < KC; ++c4) {
    const float4 qv = __ldg(reinterpret_cast<const float4*>(qp + c4 * 4));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(rp[u] + c4 * 4));
      a0[u] = fmaf(qv.x, rv.x, a0[u]);
      a1[u] = fmaf(qv.y, rv.y, a1[u]);
      a2[u] = fmaf(qv.z, rv.z, a2[u]);
      // columns d .. D_PAD-2 are zero on both sides; the last one is the threshold dimension when fold
      if (!(c4 == KC - 1 && fold)) a3[u] = fmaf(qv.w, rv.w, a3[u]);
    }
  }
  unsigned long long key[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const float dist = 1.0f - ((a0[u] + a1[u]) + (a2[u] + a3[u]));
    key[u] = j[u] >= 0 ? (((unsigned long long)f32_orderable(dist) << 32) | (unsigned)j[u]) : ~0ull;
  }
  warp_sort_u64<U>(key, lane);
  // certificate: every reference that was filtered out has exact dot < thr + eps, so it is strictly worse
  // than the k-th candidate whenever dist_k <= (1 - thr) - margin
  const unsigned long long kk = __shfl_sync(0xffffffffu, key[0], k - 1);
  const float dist_k = kk == ~0ull ? INFINITY : f32_from_orderable((uint32_t)(kk >> 32));
  ok = c >= k && dist_k <= (1.0f - thr) - MMA_MARGIN;
  if (lane < k) {
    const bool has = key[0] != ~0ull;
    pb.out_idx[(size_t)qrow * k + lane] = has ? (int)(key[0] & 0xffffffffull) : -1;
    pb.out_dist[(size_t)qrow * k + lane] = has ? f32_from_orderable((uint32_t)(key[0] >> 32)) : INFINITY;
  }
}

template <int D_PAD>
__global__ void __launch_bounds__(256)
knn_tc2_finish_kernel(const KnnProblem* __restrict__ probs, const KnnTc2Aux* __restrict__ auxs,
                      const KnnItem* __restrict__ items, int k, int cap, int fold, int* __restrict__ blk_fail) {
  const int item = blockIdx.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const KnnItem wi = items[item];
  const KnnProblem pb = probs[wi.prob];
  const KnnTc2Aux ax = auxs[wi.prob];
  const int qrow = wi.qblock * QBLOCK + (blockIdx.x & 31) * 8 + warp;
  if (qrow >= pb.nq) return;
  const size_t gr = (size_t)ax.row0 + qrow;
  const int c0 = ax.cnt[gr * 2], c1 = ax.cnt[gr * 2 + 1];
  const float thr = ax.thr[gr];
  const int* cand = ax.cand + gr * 2 * KNN_TC2_CAND_STRIDE;
  bool ok = false;
  if (c0 <= cap && c1 <= cap) {
    const int c = c0 + c1;
    if (c <= 32) finish_row<D_PAD, 1>(pb, cand, c0, c1, qrow, k, fold != 0, thr, lane, ok);
    else if (c <= 64) finish_row<D_PAD, 2>(pb, cand, c0, c1, qrow, k, fold != 0, thr, lane, ok);
    else finish_row<D_PAD, 4>(pb, cand, c0, c1, qrow, k, fold != 0, thr, lane, ok);
  }
  if (!ok && lane == 0) blk_fail[item] = 1;
}

template <int D_PAD, bool FOLD, bool LD32>
void launch_main_impl(Context& ctx, const KnnProblem* d_probs, const KnnTc2Aux* d_aux, const KnnItem* d_items,
                      int n_items, int k, int cap);

template <int D_PAD, bool FOLD>
void launch_main(Context& ctx, const KnnProblem* d_probs, const KnnTc2Aux* d_aux, const KnnItem* d_items,
                 int n_items, int k, int cap) {
  if (ctx.knn_ld32) launch_main_impl<D_PAD, FOLD, true>(ctx, d_probs, d_aux, d_items, n_items, k, cap);
  else launch_main_impl<D_PAD, FOLD, false>(ctx, d_probs, d_aux, d_items, n_items, k, cap);
}

template <int D_PAD, bool FOLD, bool LD32>
void launch_main_impl(Context& ctx, const KnnProblem* d_probs, const KnnTc2Aux* d_aux, const KnnItem* d_items,
                 int n_items, int k, int cap) {
  const int stagger = ctx.knn_stagger;
  using C = T2Cfg<D_PAD>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  CG_CUDA(cudaFuncSetAttribute(knn_tc2_kernel<D_PAD, FOLD, LD32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               C::SMEM_BYTES));
  const int grid = n_items < ctx.sm_count ? n_items : ctx.sm_count;
  knn_tc2_kernel<D_PAD, FOLD, LD32><<<grid, T2_THREADS, C::SMEM_BYTES, ctx.stream>>>(d_probs, d_aux, d_items, n_items, k,
                                                                              cap, stagger);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

}  // namespace

bool knn_tc2_plan(int nr, int k, KnnTc2Aux& ax) {
  const int n_tiles = (nr + TILE_N - 1) / TILE_N;
  int valid;
  if (n_tiles >= 24) {
    ax.sg = 64;
    ax.tg = (n_tiles + NGT_MAX - 1) / NGT_MAX;
    ax.ngt = (n_tiles + ax.tg - 1) / ax.tg;
    const int last_tiles = n_tiles - (ax.ngt - 1) * ax.tg;
    const int nv_last = nr - (n_tiles - 1) * TILE_N;
    valid = 2 * ax.ngt - ((last_tiles == 1 && nv_last <= 64) ? 1 : 0);
  } else if (n_tiles >= 12) {
    ax.sg = 32;
    ax.tg = 1;
    ax.ngt = 2 * n_tiles;
    valid = (nr + 31) / 32;
  } else {
    ax.sg = 16;
    ax.tg = 1;
    ax.ngt = 4 * n_tiles;
    valid = (nr + 15) / 16;
  }
  // at least 1.25 k non-empty groups keeps the candidate lists near 1.6 k and leaves >= 3 tiles per phase
  return n_tiles >= 3 && ax.ngt <= NGT_MAX && 4 * valid >= 5 * k;
}

void knn_tc2_run(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, const std::vector<int>& plist,
                 std::vector<KnnTc2Aux>& aux, int n_probs, int k, int d_pad, int d, std::vector<KnnItem>& redo) {
  const bool fold = d < d_pad;
  int cap = ctx.knn_cand_cap;
  if (cap < 1 || cap > KNN_TC2_CAND_STRIDE) cap = KNN_TC2_CAND_STRIDE;
  const size_t row_budget = (size_t)4 << 20;  // candidate lists: 512 B per query row
  size_t p0 = 0;
  while (p0 < plist.size()) {
    size_t rows = 0, p1 = p0;
    std::vector<KnnItem> items;
    while (p1 < plist.size()) {
      const int p = plist[p1];
      const size_t r = round_up((size_t)h_probs[p].nq, QBLOCK);
      if (p1 > p0 && rows + r > row_budget) break;
      aux[p].row0 = (long long)rows;
      rows += r;
      const int nb = (h_probs[p].nq + QBLOCK - 1) / QBLOCK;
      for (int b = 0; b < nb; ++b) items.push_back(KnnItem{p, b});
      ++p1;
    }
    DevBuf<int> cand, cnt, fail;
    DevBuf<float> thr;
    cand.alloc(rows * 2 * KNN_TC2_CAND_STRIDE);
    cnt.alloc(rows * 2);
    thr.alloc(rows);
    fail.alloc(items.size());
    CG_CUDA(cudaMemsetAsync(fail.p, 0, items.size() * sizeof(int), ctx.stream));
    for (size_t q = p0; q < p1; ++q) {
      KnnTc2Aux& a = aux[plist[q]];
      a.cand = cand.p;
      a.cnt = cnt.p;
      a.thr = thr.p;
    }
    DevBuf<KnnTc2Aux> d_aux;
    d_aux.upload(aux.data(), (size_t)n_probs, ctx.stream);
    DevBuf<KnnItem> d_items;
    d_items.upload(items.data(), items.size(), ctx.stream);
    const int ni = (int)items.size();
    if (d_pad == 32) {
      if (fold) launch_main<32, true>(ctx, d_probs, d_aux.p, d_items.p, ni, k, cap);
      else launch_main<32, false>(ctx, d_probs, d_aux.p, d_items.p, ni, k, cap);
      knn_tc2_finish_kernel<32><<<ni * 32, 256, 0, ctx.stream>>>(d_probs, d_aux.p, d_items.p, k, cap, fold ? 1 : 0,
                                                                 fail.p);
    } else {
      if (fold) launch_main<64, true>(ctx, d_probs, d_aux.p, d_items.p, ni, k, cap);
      else launch_main<64, false>(ctx, d_probs, d_aux.p, d_items.p, ni, k, cap);
      knn_tc2_finish_kernel<64><<<ni * 32, 256, 0, ctx.stream>>>(d_probs, d_aux.p, d_items.p, k, cap, fold ? 1 : 0,
                                                                 fail.p);
    }
    ctx.stats.kernels++;
    CG_CUDA(cudaGetLastError());
    std::vector<int> h_fail(items.size());
    fail.download(h_fail.data(), items.size(), ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
#if defined(TC2_MMA_ONLY) || defined(TC2_LOAD_ONLY)  // measurement builds produce no results: nothing to redo
    std::fill(h_fail.begin(), h_fail.end(), 0);
#endif
    for (size_t q = 0; q < items.size(); ++q)
      if (h_fail[q]) {
        redo.push_back(items[q]);
        ctx.knn_redo_blocks++;
      }
    p0 = p1;
  }
}

}  // namespace cg
