// K2s: streaming exact cross-kNN on the sm_100a tensor cores (angular metric, k <= 32).  The first design of the
// round; since knn_tc2.cu (two tensor-core phases, no list in the inner loop) it serves the small problems and the
// 256-row blocks whose candidate lists overflow there, and hosts the dispatch (knn_tc_launch, knn_launch).
//
// One persistent CTA per SM walks work items (problem, block of 256 query rows):
//   warp 0   producer : 1-D bulk async copies (TMA engine) of the query block (A, 256 rows) and of
//                       reference tiles (B, 128 rows) from the pre-tiled fp32 panels into shared memory
//   warp 1   MMA      : tcgen05.mma kind::tf32, 2 x (M=128, N=128) accumulators per reference tile,
//                       fp32 accumulate in TMEM, double-buffered (4 x 128 = 512 TMEM columns)
//   warps 2-9 filter  : one query row per thread. tcgen05.ld the row's 128 accumulators; the threshold
//                       (1 - kth_dist) - margin rides in a spare K dimension of the MMA (query rows hold thr,
//                       reference rows hold -1), so a survivor is a positive accumulator: ONE funnel shift
//                       per candidate collects the sign bits into masks; TMEM is released and the masks go
//                       to the paired re-rank warp through a small shared-memory ring.
//   warps 10-17 re-rank: compact the (row, column) survivors of their 32 rows into a queue, re-score them
//                       lane-parallel with the canonical fp32 fma chains from the SAME fp32 tiles (A and B)
//                       in shared memory, buffer them per row and merge them into the sorted top-k lists
//                       in lock step (register insertion network) on a schedule shared by the whole CTA;
//                       the row's new threshold is written back into the A tile.
// The TF32 product is only a filter: |tf32 dot - fp32 chain| < margin is a rigorous bound for unit
// vectors, so the returned (dist, index) lists are bit-identical to the exact fp32 search
// (oracle/conos_oracle.c) including tie order (dist, then lower reference index).
//
// Algorithmic work per launch: 2*nq*nr*d flops; HBM bytes 4*d*(nq+nr) read + 8*k*nq written.
#include "knn.cuh"
#include "ptx.cuh"

namespace cg {
namespace {

constexpr int TC_THREADS = 576;  // warp 0 producer, warp 1 MMA, warps 2-9 filter, warps 10-17 re-rank
constexpr int TILE_N = 128;
constexpr int QBLOCK = 256;
constexpr int LIST_STRIDE = KNN_TC_MAXK + 1;  // odd stride: thread-private rows hit distinct banks
// TF32 keeps 10 mantissa bits of each operand (worst case truncation): relative product error
// < 2*2^-10 + 2^-20, sum |q_i r_i| <= 1 for unit vectors, plus fp32 accumulation and the rounding of
// dist = 1 - dot. 2.2e-3 leaves > 10 % slack over the 1.96e-3 bound.
constexpr float TF32_MARGIN = 2.2e-3f;
constexpr int FRONT_FLAG = 0x40000000;  // marks a list element pushed from one 16-slot half into the next

template <int D_PAD>
struct TcCfg {
  static constexpr int KC = D_PAD / 4;             // 16-byte chunks per row
  static constexpr int ROW_BYTES = D_PAD * 4;
  static constexpr int A_BYTES = QBLOCK * ROW_BYTES;
  static constexpr int B_BYTES = TILE_N * ROW_BYTES;
  static constexpr int NSTAGE = D_PAD == 32 ? 4 : 2;
  static constexpr int KSTEPS = D_PAD / 8;          // tf32 MMA K = 8
  static constexpr int PEND = D_PAD == 32 ? 8 : 6;  // re-scored survivors a row buffers before it merges them
  static constexpr int PEND_STRIDE = PEND + 1;      // odd: thread-private rows hit distinct banks
  static constexpr int QCAP = D_PAD == 32 ? 512 : 256;  // survivor queue entries per re-rank warp
  static constexpr int RCOLS = QCAP / 32;               // columns per round when a tile overflows the queue
  static constexpr int MDEPTH = D_PAD == 32 ? 4 : 2;    // mask ring depth between a filter and its re-rank warp
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + A_BYTES;
  static constexpr int OFF_LD = OFF_B + NSTAGE * B_BYTES;
  static constexpr int OFF_LI = OFF_LD + QBLOCK * LIST_STRIDE * 4;
  static constexpr int OFF_Q = OFF_LI + QBLOCK * LIST_STRIDE * 4;    // per-warp survivor queues
  static constexpr int OFF_PD = OFF_Q + 8 * QCAP * 4;                // per-row pending survivors (dist)
  static constexpr int OFF_PI = OFF_PD + QBLOCK * PEND_STRIDE * 4;    // per-row pending survivors (index)
  static constexpr int OFF_MR = OFF_PI + QBLOCK * PEND_STRIDE * 4;    // mask ring: [8 pairs][MDEPTH][32 rows][4 words]
  static constexpr int OFF_QC = OFF_MR + 8 * MDEPTH * 32 * 16;      // column of every queued survivor
  static constexpr int OFF_BAR = OFF_QC + 8 * QCAP;
  static constexpr int N_BAR = 2 + 2 * NSTAGE + 4 + 2 * 8 * MDEPTH + 2;
  static constexpr int OFF_TMEM = OFF_BAR + N_BAR * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16;
};

// FOLD: the last (spare) K dimension carries the threshold -- query rows hold thr there, reference rows hold
// -1 (knn_prep writes it), so the accumulator is dot - thr and a survivor is simply a positive accumulator:
// the filter is one funnel shift per candidate.  Needs d < D_PAD.
template <int D_PAD, int KT, bool FOLD>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const KnnProblem* __restrict__ probs, const KnnItem* __restrict__ items, int n_items, int k, int d) {
  using C = TcCfg<D_PAD>;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* smA = reinterpret_cast<float*>(smem + C::OFF_A);
  float* smB = reinterpret_cast<float*>(smem + C::OFF_B);
  float* list_d = reinterpret_cast<float*>(smem + C::OFF_LD);
  int* list_i = reinterpret_cast<int*>(smem + C::OFF_LI);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;
  uint64_t* b_empty = bars + 2 + C::NSTAGE;
  uint64_t* acc_full = bars + 2 + 2 * C::NSTAGE;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* m_full = acc_empty + 2;                 // [8][MDEPTH]
  uint64_t* m_empty = m_full + 8 * C::MDEPTH;       // [8][MDEPTH]
  uint64_t* thr_ready = m_empty + 8 * C::MDEPTH;    // the 8 re-rank warps have merged tile 0 of the item
  uint64_t* lists_ready = thr_ready + 1;            // (!FOLD) the 8 re-rank warps have reset their lists for the item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1 + 8);  // MMA commit + 8 re-rank warps (they re-score from the A tile)
    for (int s = 0; s < C::NSTAGE; ++s) {
      ptx::mbar_init(b_full + s, 1);
      ptx::mbar_init(b_empty + s, 1 + 8);  // MMA commit + 8 re-rank warps (they re-score from the tile)
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(acc_full + s, 1);
      ptx::mbar_init(acc_empty + s, 8);    // 8 filter warps
    }
    for (int s = 0; s < 8 * C::MDEPTH; ++s) {
      ptx::mbar_init(m_full + s, 1);
      ptx::mbar_init(m_empty + s, 1);
    }
    ptx::mbar_init(thr_ready, 8);
    ptx::mbar_init(lists_ready, 8);
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
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const KnnItem wi = items[item];
        const KnnProblem pb = probs[wi.prob];
        while (!ptx::mbar_try_wait(a_empty, (it & 1) ^ 1)) __nanosleep(64);
        ptx::mbar_arrive_expect_tx(a_full, C::A_BYTES);
        ptx::bulk_g2s(smA, pb.q_tiled + (size_t)wi.qblock * QBLOCK * D_PAD, C::A_BYTES, a_full);
        const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
        for (int t = 0; t < n_tiles; ++t, ++g) {
          const uint32_t slot = g % C::NSTAGE, ph = (g / C::NSTAGE) & 1;
          while (!ptx::mbar_try_wait(b_empty + slot, ph ^ 1)) __nanosleep(64);  // not latency critical
          ptx::mbar_arrive_expect_tx(b_full + slot, C::B_BYTES);
          ptx::bulk_g2s(smB + (size_t)slot * (C::B_BYTES / 4), pb.r_tiled + (size_t)t * TILE_N * D_PAD, C::B_BYTES,
                        b_full + slot);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(2u, 128u, (uint32_t)TILE_N);
      const uint32_t a_addr = ptx::smem_u32(smA), b_addr0 = ptx::smem_u32(smB);
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const KnnItem wi = items[item];
        const int nr = probs[wi.prob].nr;
        const int n_tiles = (nr + TILE_N - 1) / TILE_N;
        ptx::mbar_wait(a_full, it & 1);
        for (int t = 0; t < n_tiles; ++t, ++g) {
          const uint32_t slot = g % C::NSTAGE, ph = (g / C::NSTAGE) & 1;
          const uint32_t stage = g & 1, aph = (g >> 1) & 1;
          while (!ptx::mbar_try_wait(b_full + slot, ph)) __nanosleep(32);
          while (!ptx::mbar_try_wait(acc_empty + stage, aph ^ 1)) __nanosleep(32);
          ptx::tc_fence_after();
          const uint32_t b_addr = b_addr0 + slot * C::B_BYTES;
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const uint32_t d_tmem = tmem_base + (stage * 2 + mb) * TILE_N;
#pragma unroll
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
              const uint64_t ad = ptx::umma_desc_kmajor_noswz(a_addr + mb * 128 * C::ROW_BYTES + ks * 256, 128,
                                                               C::KC * 128);
              const uint64_t bd = ptx::umma_desc_kmajor_noswz(b_addr + ks * 256, 128, C::KC * 128);
              ptx::mma_tf32_ss(d_tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
          }
          ptx::mma_commit(b_empty + slot);
          ptx::mma_commit(acc_full + stage);
          // the first tile of an item passes unfiltered; wait until its survivors are merged and the rows'
          // thresholds sit in the A tile, otherwise the pipeline depth worth of tiles passes unfiltered too
          if (FOLD && t == 0 && n_tiles > 1) ptx::mbar_wait(thr_ready, it & 1);
        }
        ptx::mma_commit(a_empty);
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // ------------------------------------------------------------ filter warps: TMEM -> survivor bit masks
    const int e = warp - 2;
    const int mb = e >> 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int row_local = mb * 128 + quarter * 32 + lane;
    uint4* mring = reinterpret_cast<uint4*>(smem + C::OFF_MR) + (size_t)e * C::MDEPTH * 32;
    uint32_t g = 0, mg = 0, fit = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++fit) {
      const KnnItem wi = items[item];
      const KnnProblem pb = probs[wi.prob];
      // without the folded threshold the filter reads the k-th distance from the row's list: it must not see
      // the previous item's list
      if (!FOLD) ptx::mbar_wait(lists_ready, fit & 1);
      const bool valid_row = wi.qblock * QBLOCK + row_local < pb.nq;  // zero padding rows tie with everything
      const uint32_t row_mask = valid_row ? 0xffffffffu : 0u;
      const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
      float thr = 0.f;
      for (int t = 0; t < n_tiles; ++t, ++g, ++mg) {
        const uint32_t stage = g & 1, aph = (g >> 1) & 1;
        const uint32_t ms = mg % C::MDEPTH, mph = (mg / C::MDEPTH) & 1;
        if (!FOLD) {
          // threshold published by the row's re-rank thread in its list slot KT (slot k-1 holds the k-th distance)
          const float kth = list_d[row_local * LIST_STRIDE + (k - 1)];
          thr = (1.0f - kth) - TF32_MARGIN;  // -inf while the list is not full
        }
        ptx::mbar_wait(acc_full + stage, aph);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (stage * 2 + mb) * TILE_N;
        uint32_t mask[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t va[16], vb[16];
          ptx::tmem_ld_32x32b_x16(taddr + ch * 32, va);
          ptx::tmem_ld_32x32b_x16(taddr + ch * 32 + 16, vb);
          ptx::tmem_wait_ld();
          uint32_t m = 0;
          if (FOLD) {
            // accumulator = dot - thr: collect the sign bits (column 31 first), survivors = NOT sign
#pragma unroll
            for (int c = 15; c >= 0; --c) m = __funnelshift_l(vb[c], m, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) m = __funnelshift_l(va[c], m, 1);
            m = ~m;
          } else {
            // sign bit of (thr - v) is set iff v > thr
#pragma unroll
            for (int c = 15; c >= 0; --c) m = __funnelshift_l(__float_as_uint(thr - __uint_as_float(vb[c])), m, 1);
#pragma unroll
            for (int c = 15; c >= 0; --c) m = __funnelshift_l(__float_as_uint(thr - __uint_as_float(va[c])), m, 1);
          }
          mask[ch] = m & row_mask;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_empty + stage);
        const int jbase = t * TILE_N;
        if (jbase + TILE_N > pb.nr) {  // zero padding rows of the last reference tile
          const int nv = pb.nr - jbase;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int lo = nv - w * 32;
            mask[w] &= lo >= 32 ? 0xffffffffu : (lo <= 0 ? 0u : ((1u << lo) - 1u));
          }
        }
        ptx::mbar_wait(m_empty + e * C::MDEPTH + ms, mph ^ 1);
        mring[ms * 32 + lane] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(m_full + e * C::MDEPTH + ms);
      }
    }
  } else {
    // ------------------------------------------------------------ re-rank warps: exact fp32 + top-k lists
    const int e = warp - 10;       // paired with filter warp e + 2
    const int mb = e >> 2;
    const int quarter = (e + 2) & 3;
    const int warp_row0 = mb * 128 + quarter * 32;
    const int row_local = warp_row0 + lane;
    float* my_d = list_d + row_local * LIST_STRIDE;
    int* my_i = list_i + row_local * LIST_STRIDE;
    uint32_t* qbuf = reinterpret_cast<uint32_t*>(smem + C::OFF_Q) + e * C::QCAP;
    unsigned char* qcol = smem + C::OFF_QC + e * C::QCAP;
    float* pend_d = reinterpret_cast<float*>(smem + C::OFF_PD) + row_local * C::PEND_STRIDE;
    int* pend_i = reinterpret_cast<int*>(smem + C::OFF_PI) + row_local * C::PEND_STRIDE;
    const uint4* mring = reinterpret_cast<const uint4*>(smem + C::OFF_MR) + (size_t)e * C::MDEPTH * 32;
    // this row's threshold slot in the A tile: last K dimension
    float* thr_slot = smA + (size_t)(row_local >> 3) * (8 * D_PAD) + (size_t)(row_local & 7) * 4 +
                      (size_t)((D_PAD - 1) >> 2) * 32 + ((D_PAD - 1) & 3);
    // survivors accepted per row between two scheduled merges ~ k*ln(t1/t0): keep it near PEND/2
    const float flush_growth = expf((float)(C::PEND / 2) / (float)k) - 1.0f;
    uint32_t g = 0, it = 0, mg = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const KnnItem wi = items[item];
      const KnnProblem pb = probs[wi.prob];
      const int qrow = wi.qblock * QBLOCK + row_local;
      const bool valid_row = qrow < pb.nq;
      const int n_tiles = (pb.nr + TILE_N - 1) / TILE_N;
      for (int p = 0; p < KT; ++p) { my_d[p] = INFINITY; my_i[p] = -1; }
      if (!FOLD) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(lists_ready);
      }
      float kth = INFINITY;
      int npend = 0;
      int next_flush = 0;
      // Merge the buffered survivors of all 32 rows into their sorted lists, in lock step: the list
      // goes to registers, every survivor runs through a branch-free insertion network
      //   new[p] = old[p-1] > x ? old[p-1] : (old[p] > x ? x : old[p])
      // (arrival = ascending reference index order, strict compares: equal distances keep the lower
      // index, survivors not below the k-th distance fall off the end), and the list goes back.
      auto flush = [&]() {
        const int maxp = __reduce_max_sync(0xffffffffu, npend);
        if (maxp == 0) return;
        const float old_kth = kth;
        // 16 list slots at a time in registers: a survivor is inserted into the lower half first, the element it
        // pushes out (or the survivor itself if it does not belong there) continues into the next half
#pragma unroll
        for (int h0 = 0; h0 < KT; h0 += 16) {
          float ld[16];
          int li[16];
#pragma unroll
          for (int p = 0; p < 16; ++p) { ld[p] = my_d[h0 + p]; li[p] = my_i[h0 + p]; }
          for (int u = 0; u < maxp; ++u) {
            const bool has = u < npend;
            const float x = has ? pend_d[u] : INFINITY;
            const int jr = has ? pend_i[u] : -1;
            // an element pushed out of the previous half is <= everything here and older than its equals:
            // it goes to the front of this half (compare as -inf); a fresh survivor goes after its equals
            const bool front = has && h0 > 0 && jr >= 0 && (jr & FRONT_FLAG) != 0;
            const int j = has ? (jr >= 0 ? (jr & ~FRONT_FLAG) : -1) : -1;
            const float xc = front ? -INFINITY : x;
            bool c_hi = ld[15] > xc;
            if (h0 + 16 < KT) {  // pass the evicted element on to the next half (same buffer slot)
              pend_d[u] = c_hi ? ld[15] : x;
              pend_i[u] = c_hi ? (li[15] >= 0 ? (li[15] | FRONT_FLAG) : -1) : j;  // empty slots (inf, -1) just drop
            }
#pragma unroll
            for (int p = 15; p >= 1; --p) {
              const bool c_lo = ld[p - 1] > xc;
              ld[p] = c_lo ? ld[p - 1] : (c_hi ? x : ld[p]);
              li[p] = c_lo ? li[p - 1] : (c_hi ? j : li[p]);
              c_hi = c_lo;
            }
            ld[0] = c_hi ? x : ld[0];
            li[0] = c_hi ? j : li[0];
          }
#pragma unroll
          for (int p = 0; p < 16; ++p) {
            my_d[h0 + p] = ld[p];
            my_i[h0 + p] = li[p];
            if (h0 + p == k - 1) kth = ld[p];
          }
        }
        npend = 0;
        if (FOLD && kth != old_kth) {
          // publish thr = (1 - kth) - margin in the A tile, rounded toward -inf to a TF32-exact value, so the
          // tensor core subtracts it from every later dot product of this row
          const float thr = (1.0f - kth) - TF32_MARGIN;
          uint32_t bits = __float_as_uint(thr);
          if ((bits & 0x1fffu) != 0u) bits = (bits & 0xffffe000u) + ((bits >> 31) ? 0x2000u : 0u);
          *thr_slot = __uint_as_float(bits);
          ptx::fence_proxy_async();
        }
      };
      ptx::mbar_wait(a_full, it & 1);  // the fp32 query block (re-rank operand)
      if (FOLD && !valid_row) {
        *thr_slot = 1e30f;  // padding rows never pass once this lands (the filter also masks them)
        ptx::fence_proxy_async();
      }
      for (int t = 0; t < n_tiles; ++t, ++g, ++mg) {
        const uint32_t slot = g % C::NSTAGE, ph = (g / C::NSTAGE) & 1;
        const uint32_t ms = mg % C::MDEPTH, mph = (mg / C::MDEPTH) & 1;
        ptx::mbar_wait(m_full + e * C::MDEPTH + ms, mph);
        const uint4 mk = mring[ms * 32 + lane];
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(m_empty + e * C::MDEPTH + ms);
        const uint32_t mask[4] = {mk.x, mk.y, mk.z, mk.w};
        const int jbase = t * TILE_N;
        const int cnt_all = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]);
        const int total = __reduce_add_sync(0xffffffffu, cnt_all);
        if (total > 0) {
          // the fp32 tile itself (async-proxy write, observed complete through its own barrier)
          ptx::mbar_wait(b_full + slot, ph);
          const float* tile = smB + (size_t)slot * (C::B_BYTES / 4);
          // one round over all 128 columns when the survivors fit the queue, else rounds of RCOLS columns
          const int n_rounds = total <= C::QCAP ? 1 : TILE_N / C::RCOLS;
#pragma unroll 1
          for (int r = 0; r < n_rounds; ++r) {
            uint32_t sm[4];
            if (n_rounds == 1) {
#pragma unroll
              for (int w = 0; w < 4; ++w) sm[w] = mask[w];
            } else {
              const int c0 = r * C::RCOLS;  // RCOLS divides 32
              const uint32_t sel = ((C::RCOLS == 32) ? 0xffffffffu : ((1u << C::RCOLS) - 1u)) << (c0 & 31);
#pragma unroll
              for (int w = 0; w < 4; ++w) sm[w] = (c0 >> 5) == w ? (mask[w] & sel) : 0u;
            }
            const int cnt = __popc(sm[0]) + __popc(sm[1]) + __popc(sm[2]) + __popc(sm[3]);
            int incl = cnt;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
              int tt = __shfl_up_sync(0xffffffffu, incl, dlt);
              if (lane >= dlt) incl += tt;
            }
            const int E = __shfl_sync(0xffffffffu, incl, 31);
            if (E == 0) continue;
            const int off = incl - cnt;
            const int maxc = __reduce_max_sync(0xffffffffu, cnt);
            // A: compaction -- each row appends its (row, column) survivors, ascending column
            {
              uint32_t s0 = sm[0], s1 = sm[1], s2 = sm[2], s3 = sm[3];
              for (int u = 0; u < maxc; ++u) {
                if (u < cnt) {
                  int col;
                  if (s0) { col = __ffs(s0) - 1; s0 &= s0 - 1; }
                  else if (s1) { col = 32 + __ffs(s1) - 1; s1 &= s1 - 1; }
                  else if (s2) { col = 64 + __ffs(s2) - 1; s2 &= s2 - 1; }
                  else { col = 96 + __ffs(s3) - 1; s3 &= s3 - 1; }
                  qbuf[off + u] = ((uint32_t)lane << 7) | (uint32_t)col;
                  qcol[off + u] = (unsigned char)col;
                }
              }
            }
            __syncwarp();
            // B: canonical fp32 re-score, one survivor per lane (only the d real components: the padding /
            // threshold dimensions are skipped)
            for (int q0 = 0; q0 < E; q0 += 32) {
              const int qi = q0 + lane;
              if (qi < E) {
                const uint32_t ent = qbuf[qi];
                const int rl = warp_row0 + (int)(ent >> 7);
                const int col = (int)(ent & 127u);
                const float* qp = smA + (size_t)(rl >> 3) * (8 * D_PAD) + (size_t)(rl & 7) * 4;
                const float* rp = tile + (size_t)(col >> 3) * (8 * D_PAD) + (size_t)(col & 7) * 4;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                // columns d .. D_PAD-2 are zero on both sides (fma(0,0,a) == a); the last column is the
                // threshold dimension when FOLD and is skipped
#pragma unroll
                for (int c4 = 0; c4 < C::KC; c4 += 4) {
                  float4 qv[4], rv[4];
#pragma unroll
                  for (int u4 = 0; u4 < 4; ++u4) {
                    qv[u4] = *reinterpret_cast<const float4*>(qp + (c4 + u4) * 32);
                    rv[u4] = *reinterpret_cast<const float4*>(rp + (c4 + u4) * 32);
                  }
#pragma unroll
                  for (int u4 = 0; u4 < 4; ++u4) {
                    a0 = fmaf(qv[u4].x, rv[u4].x, a0);
                    a1 = fmaf(qv[u4].y, rv[u4].y, a1);
                    a2 = fmaf(qv[u4].z, rv[u4].z, a2);
                    if (!(FOLD && c4 + u4 == C::KC - 1)) a3 = fmaf(qv[u4].w, rv[u4].w, a3);
                  }
                }
                const float dist = 1.0f - ((a0 + a1) + (a2 + a3));
                qbuf[qi] = __float_as_uint(dist);
              }
            }
            __syncwarp();
            // C: each row's owner buffers its survivors that beat the (possibly stale) k-th distance
            {
              for (int u = 0; u < maxc; ++u) {
                const bool has = u < cnt;
                float dist = INFINITY;
                int col = 0;
                if (has) {
                  dist = __uint_as_float(qbuf[off + u]);
                  col = qcol[off + u];
                }
                const bool take = has && dist < kth;
                if (__any_sync(0xffffffffu, take && npend == C::PEND)) flush();
                if (take && dist < kth) {
                  pend_d[npend] = dist;
                  pend_i[npend] = jbase + col;
                  ++npend;
                }
              }
            }
            __syncwarp();
          }
        }
        // scheduled merge: the same tiles for every warp of the CTA, geometrically spaced
        if (t >= next_flush) {
          flush();
          const int step = (int)((float)(t + 1) * flush_growth);
          next_flush = t + (step > 1 ? step : 1);
        }
        __syncwarp();
        if (lane == 0) {
          if (FOLD && t == 0 && n_tiles > 1) ptx::mbar_arrive(thr_ready);
          ptx::mbar_arrive(b_empty + slot);
        }
      }
      flush();
      if (valid_row) {
        for (int p = 0; p < k; ++p) {
          pb.out_idx[(size_t)qrow * k + p] = my_i[p];
          pb.out_dist[(size_t)qrow * k + p] = my_d[p];
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_empty);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

template <int D_PAD, int KT, bool FOLD>
void launch_tc(Context& ctx, const KnnProblem* d_probs, const KnnItem* d_items, int n_items, int k, int d) {
  using C = TcCfg<D_PAD>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  CG_CUDA(cudaFuncSetAttribute(knn_tc_kernel<D_PAD, KT, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               C::SMEM_BYTES));
  int grid = n_items < ctx.sm_count ? n_items : ctx.sm_count;
  knn_tc_kernel<D_PAD, KT, FOLD><<<grid, TC_THREADS, C::SMEM_BYTES, ctx.stream>>>(d_probs, d_items, n_items, k, d);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

}  // namespace

void knn_tc_launch(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, int n_probs, int k,
                   int d_pad, int d) {
  CG_CHECK(k >= 1 && k <= KNN_TC_MAXK, "tensor-core kNN serves 1 <= k <= 32");
  CG_CHECK(d_pad == 32 || d_pad == 64, "d_pad must be 32 or 64");
  // problems large enough for the two-phase kernel go there; the rest, and the 256-row blocks it hands back
  // (candidate overflow from ties / duplicates), run through the streaming kernel
  std::vector<KnnItem> items;
  std::vector<KnnTc2Aux> aux((size_t)n_probs);
  std::vector<int> plist;
  double cand = 0.0, scores = 0.0, bytes = 0.0;
  for (int p = 0; p < n_probs; ++p) {
    if (h_probs[p].nq <= 0) continue;
    // ONE distance matrix serves both directions of a pair -- a problem whose mirror image (operands
    // swapped) is part of the same launch counts half; a sample against itself counts in full
    bool mirrored = false;
    for (int o = 0; o < n_probs && !mirrored; ++o)
      mirrored = o != p && h_probs[o].q_tiled == h_probs[p].r_tiled && h_probs[o].r_tiled == h_probs[p].q_tiled &&
                 h_probs[p].q_tiled != h_probs[p].r_tiled;
    cand += (mirrored ? 0.5 : 1.0) * (double)h_probs[p].nq * (double)h_probs[p].nr;
    scores += (double)h_probs[p].nq * (double)h_probs[p].nr;
    bytes += 4.0 * d_pad * ((double)h_probs[p].nq + h_probs[p].nr) + 8.0 * k * (double)h_probs[p].nq;
    if (ctx.knn_engine == 0 && knn_tc2_plan(h_probs[p].nr, k, aux[p])) {
      plist.push_back(p);
      continue;
    }
    int nb = (h_probs[p].nq + QBLOCK - 1) / QBLOCK;
    for (int b = 0; b < nb; ++b) items.push_back(KnnItem{p, b});
  }
  if (items.empty() && plist.empty()) return;
  if (ctx.time_knn) {
    if (!ctx.knn_ev0) {
      CG_CUDA(cudaEventCreate(&ctx.knn_ev0));
      CG_CUDA(cudaEventCreate(&ctx.knn_ev1));
    }
    CG_CUDA(cudaEventRecord(ctx.knn_ev0, ctx.stream));
  }
  if (!plist.empty()) knn_tc2_run(ctx, d_probs, h_probs, plist, aux, n_probs, k, d_pad, d, items);
  if (!items.empty()) {
    DevBuf<KnnItem> d_items;
    d_items.upload(items.data(), items.size(), ctx.stream);
    const int ni = (int)items.size();
    const bool fold = d < d_pad;  // a spare K dimension exists: the threshold rides in the MMA
    if (d_pad == 32) {
      if (k <= 16) { if (fold) launch_tc<32, 16, true>(ctx, d_probs, d_items.p, ni, k, d); else launch_tc<32, 16, false>(ctx, d_probs, d_items.p, ni, k, d); }
      else { if (fold) launch_tc<32, 32, true>(ctx, d_probs, d_items.p, ni, k, d); else launch_tc<32, 32, false>(ctx, d_probs, d_items.p, ni, k, d); }
    } else {
      if (k <= 16) { if (fold) launch_tc<64, 16, true>(ctx, d_probs, d_items.p, ni, k, d); else launch_tc<64, 16, false>(ctx, d_probs, d_items.p, ni, k, d); }
      else { if (fold) launch_tc<64, 32, true>(ctx, d_probs, d_items.p, ni, k, d); else launch_tc<64, 32, false>(ctx, d_probs, d_items.p, ni, k, d); }
    }
  }
  if (ctx.time_knn) {
    CG_CUDA(cudaEventRecord(ctx.knn_ev1, ctx.stream));
    CG_CUDA(cudaEventSynchronize(ctx.knn_ev1));
    float ms = 0.f;
    CG_CUDA(cudaEventElapsedTime(&ms, ctx.knn_ev0, ctx.knn_ev1));
    ctx.knn_ms_accum += ms;
    ctx.knn_launches++;
    ctx.knn_candidates += cand;
    ctx.knn_scores += scores;
    ctx.knn_bytes += bytes;
  }
}

void knn_launch(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, int n_probs, int k, int d_pad,
                int d, Metric metric) {
  if (d_pad > 64) {  // wider than the tensor-core kernels' operand tiles: exact SIMT search
    knn_simt_launch(ctx, d_probs, h_probs, n_probs, k, d_pad, d, metric);
    return;
  }
  if (metric == METRIC_ANGULAR && k <= KNN_TC_MAXK && ctx.knn_engine != 2) {
    knn_tc_launch(ctx, d_probs, h_probs, n_probs, k, d_pad, d);
    return;
  }
  if (metric == METRIC_ANGULAR && k <= KNN_TCL_MAXK && ctx.knn_engine == 0) {
    // large k (alignment.strength > 0): counting passes on the tensor cores; whatever does not qualify, or
    // overflows its candidate lists, goes to the exact SIMT kernel
    std::vector<int> plist, rest;
    for (int p = 0; p < n_probs; ++p) {
      if (h_probs[p].nq <= 0) continue;
      (knn_tcl_applies(h_probs[p].nr, k) ? plist : rest).push_back(p);
    }
    if (!plist.empty()) knn_tcl_run(ctx, d_probs, h_probs, plist, n_probs, k, d_pad, d, rest);
    if (!rest.empty()) {
      std::vector<KnnProblem> sub;
      for (int p : rest) sub.push_back(h_probs[p]);
      DevBuf<KnnProblem> d_sub;
      d_sub.upload(sub.data(), sub.size(), ctx.stream);
      knn_simt_launch(ctx, d_sub.p, sub.data(), (int)sub.size(), k, d_pad, d, metric);
      CG_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    return;
  }
  knn_simt_launch(ctx, d_probs, h_probs, n_probs, k, d_pad, d, metric);
}

}  // namespace cg
