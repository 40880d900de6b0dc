// Issue-rate probe for tcgen05.mma.cta_group::1.kind::f16 with the tiny-K shapes of the kNN kernels (knn_tc2.cu):
// one CTA per SM, one issuing thread, operands resident in shared memory (SWIZZLE_64B rows of 32 halves), no
// producer and no epilogue.  Prints cycles per MMA instruction for several issue patterns, so that the kNN kernel's
// MMA stream can be compared with what the tensor pipe does when nothing else is in the way.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I conos_b200/csrc tests/cuda/mma_probe.cu -o mma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <vector>

#include "ptx.cuh"

using namespace cg;

__device__ __forceinline__ void mbar_spin_a(uint32_t bar, uint32_t parity) {  // non-blocking test_wait in a tight loop
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
#define WAIT_A(bar, par)                         \
  do {                                           \
    if (mode & 32) mbar_spin_a((bar), (par));    \
    else ptx::mbar_wait_a((bar), (par));         \
  } while (0)

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));         \
      exit(1);                                                                           \
    }                                                                                    \
  } while (0)

// pattern 0: the kNN tile pass -- 2 accumulators x 2 k-steps (overwrite, accumulate), stages alternate, commit per pass
// pattern 1: same without the per-pass commits
// pattern 2: one accumulator, long accumulate chain (GEMM-like K loop), commit every 4
// pattern 3: N = 256: one 128 x 256 accumulator per pass, 2 k-steps, stages alternate, commit per pass
// pattern 4: pattern 0 but the A operand of the second accumulator is the same as the first (A re-read from the same lines)
// pattern 5: four independent accumulators, 1 k-step each (K = 16)
__global__ void __launch_bounds__(128) mma_probe_kernel(int pattern, int passes, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_loop, bar_done;
  __shared__ uint32_t tmem_base;
  __half* sA = reinterpret_cast<__half*>(smem);              // 256 rows x 32 halves
  __half* sB = reinterpret_cast<__half*>(smem + 256 * 64);   // 256 rows x 32 halves
  for (int t = threadIdx.x; t < 2 * 256 * 32; t += blockDim.x)
    reinterpret_cast<__half*>(smem)[t] = __float2half(((t * 37) % 17 - 8) * 0.01f);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_loop, 1);
    ptx::mbar_init(&bar_done, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    uint64_t ad[2][2], bd[2];
    for (int mb = 0; mb < 2; ++mb)
      for (int ks = 0; ks < 2; ++ks) ad[mb][ks] = ptx::umma_desc_kmajor_swz(a0 + mb * 128 * 64 + ks * 32, 64, 512);
    for (int ks = 0; ks < 2; ++ks) bd[ks] = ptx::umma_desc_kmajor_swz(b0 + ks * 32, 64, 512);
    constexpr uint32_t id128 = ptx::umma_idesc(0u, 128u, 128u);
    constexpr uint32_t id256 = ptx::umma_idesc(0u, 128u, 256u);
    long long n_mma = 0;
    const long long t0 = clock64();
    for (int p = 0; p < passes; ++p) {
      const uint32_t st = (p & 1) * 256;
      switch (pattern) {
        case 0:
        case 1:
          ptx::mma_f16_ss(tm + st, ad[0][0], bd[0], id128, 0);
          ptx::mma_f16_ss(tm + st, ad[0][1], bd[1], id128, 1);
          ptx::mma_f16_ss(tm + st + 128, ad[1][0], bd[0], id128, 0);
          ptx::mma_f16_ss(tm + st + 128, ad[1][1], bd[1], id128, 1);
          n_mma += 4;
          if (pattern == 0) ptx::mma_commit(&bar_loop);
          break;
        case 2:
          ptx::mma_f16_ss(tm, ad[0][0], bd[0], id128, 1);
          ptx::mma_f16_ss(tm, ad[0][1], bd[1], id128, 1);
          ptx::mma_f16_ss(tm, ad[0][0], bd[0], id128, 1);
          ptx::mma_f16_ss(tm, ad[0][1], bd[1], id128, 1);
          n_mma += 4;
          ptx::mma_commit(&bar_loop);
          break;
        case 3:
          ptx::mma_f16_ss(tm + st, ad[p & 1][0], bd[0], id256, 0);
          ptx::mma_f16_ss(tm + st, ad[p & 1][1], bd[1], id256, 1);
          n_mma += 2;
          ptx::mma_commit(&bar_loop);
          break;
        case 4:
          ptx::mma_f16_ss(tm + st, ad[0][0], bd[0], id128, 0);
          ptx::mma_f16_ss(tm + st, ad[0][1], bd[1], id128, 1);
          ptx::mma_f16_ss(tm + st + 128, ad[0][0], bd[0], id128, 0);
          ptx::mma_f16_ss(tm + st + 128, ad[0][1], bd[1], id128, 1);
          n_mma += 4;
          ptx::mma_commit(&bar_loop);
          break;
        default:
          ptx::mma_f16_ss(tm, ad[0][0], bd[0], id128, 0);
          ptx::mma_f16_ss(tm + 128, ad[1][0], bd[0], id128, 0);
          ptx::mma_f16_ss(tm + 256, ad[0][1], bd[1], id128, 0);
          ptx::mma_f16_ss(tm + 384, ad[1][1], bd[1], id128, 0);
          n_mma += 4;
          ptx::mma_commit(&bar_loop);
          break;
      }
    }
    ptx::mma_commit(&bar_done);
    ptx::mbar_wait(&bar_done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = n_mma;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}


// Hand-shake probe: the accumulator round trip of knn_tc2.cu without the producer ring and without epilogue
// arithmetic.  Thread 32 issues a tile pass (4 MMAs) into one of two 256-column stages after waiting for the stage to
// be drained; 16 consumer warps wait for the stage, optionally read their 64 columns, and arrive.
//   mode bit 0: consumers tcgen05.ld their slice;  bit 1: only lane 0 of a consumer warp polls (then __syncwarp)
//   bit 2: no MMAs (barriers only);  bit 3: four stages of 128 columns (a pass = one accumulator, 2 MMAs)
__global__ void __launch_bounds__(576, 1) handshake_probe_kernel(int mode, int passes, long long* out,
                                                                 const unsigned char* gsrc, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t acc_full[4], acc_empty[4], b_full[10], b_empty[10];
  __shared__ uint32_t tmem_base;
  for (int t = threadIdx.x; t < 2 * 256 * 32; t += blockDim.x)
    reinterpret_cast<__half*>(smem)[t] = __float2half(((t * 37) % 17 - 8) * 0.01f);
  const bool ring = (mode & 16) != 0;   // bit 4: reference tiles arrive through a 10-deep ring of 8 KB bulk copies
  unsigned char* smB = smem + 2 * 256 * 64;  // 10 x 8 KB
  if (threadIdx.x == 0)
    for (int s2 = 0; s2 < 10; ++s2) {
      ptx::mbar_init(b_full + s2, 1);
      ptx::mbar_init(b_empty + s2, 1);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool four = (mode & 8) != 0;
  const int n_stage = four ? 4 : 2;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(acc_full + s, 1);
      ptx::mbar_init(acc_empty + s, four ? 8 : 16);
    }
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 0) {
    if (lane == 0 && ring) {
      uint32_t slot = 0, ph = 0;
      for (int p = 0; p < passes; ++p) {
        ptx::mbar_wait(b_empty + slot, ph ^ 1);
        ptx::mbar_arrive_expect_tx(b_full + slot, 8192);
        ptx::bulk_g2s(smB + slot * 8192, gsrc + (size_t)(p % n_tiles) * 8192, 8192, b_full + slot);
        if (++slot == 10) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 256 * 64;
      const uint32_t bfull_a = ptx::smem_u32(b_full), bempty_a = ptx::smem_u32(b_empty);
      uint32_t slot = 0, bph = 0;
      uint64_t ad[2][2], bd[2];
      for (int mb = 0; mb < 2; ++mb)
        for (int ks = 0; ks < 2; ++ks) ad[mb][ks] = ptx::umma_desc_kmajor_swz(a0 + mb * 128 * 64 + ks * 32, 64, 512);
      for (int ks = 0; ks < 2; ++ks) bd[ks] = ptx::umma_desc_kmajor_swz(b0 + ks * 32, 64, 512);
      constexpr uint32_t id128 = ptx::umma_idesc(0u, 128u, 128u);
      const uint32_t full_a = ptx::smem_u32(acc_full), empty_a = ptx::smem_u32(acc_empty);
      const long long t0 = clock64();
      for (int p = 0; p < passes; ++p) {
        const int stage = p % n_stage;
        const uint32_t par = (uint32_t)(p / n_stage) & 1u;
        uint64_t boff = 0;
        if (ring) {
          WAIT_A(bfull_a + slot * 8, bph);
          boff = (uint64_t)((256 * 64 - 256 * 64 + slot * 8192) >> 4);  // ring slots start where the static B tile is
        }
        WAIT_A(empty_a + stage * 8, par ^ 1);
        ptx::tc_fence_after();
        if (!(mode & 4)) {
          if (ring) {
            ptx::mma_f16_ss(tm + stage * 256, ad[0][0], bd[0] + boff, id128, 0);
            ptx::mma_f16_ss(tm + stage * 256, ad[0][1], bd[1] + boff, id128, 1);
            ptx::mma_f16_ss(tm + stage * 256 + 128, ad[1][0], bd[0] + boff, id128, 0);
            ptx::mma_f16_ss(tm + stage * 256 + 128, ad[1][1], bd[1] + boff, id128, 1);
          } else if (four) {
            ptx::mma_f16_ss(tm + stage * 128, ad[p & 1][0], bd[0], id128, 0);
            ptx::mma_f16_ss(tm + stage * 128, ad[p & 1][1], bd[1], id128, 1);
          } else {
            ptx::mma_f16_ss(tm + stage * 256, ad[0][0], bd[0], id128, 0);
            ptx::mma_f16_ss(tm + stage * 256, ad[0][1], bd[1], id128, 1);
            ptx::mma_f16_ss(tm + stage * 256 + 128, ad[1][0], bd[0], id128, 0);
            ptx::mma_f16_ss(tm + stage * 256 + 128, ad[1][1], bd[1], id128, 1);
          }
        }
        if (ring) {
          ptx::mma_commit_a(bempty_a + slot * 8);
          if (++slot == 10) { slot = 0; bph ^= 1; }
        }
        ptx::mma_commit_a(full_a + stage * 8);
      }
      // drain: wait until the consumers have seen the last pass of every stage
      for (int s = 0; s < n_stage; ++s) {
        const int last = ((passes - 1 - s) / n_stage) * n_stage + s;  // last pass that used stage s
        if (last >= 0 && last < passes) WAIT_A(empty_a + s * 8, (uint32_t)(last / n_stage) & 1u);
      }
      const long long t1 = clock64();
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = passes;
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int sub = e >> 2;  // 0..3: (accumulator, half) of the two-stage layout; stage-parity group of the four-stage one
    const uint32_t tm_lane = tm + ((uint32_t)(quarter * 32) << 16);
    const uint32_t full_a = ptx::smem_u32(acc_full), empty_a = ptx::smem_u32(acc_empty);
    uint32_t sink = 0;
    if (!four) {
      for (int p = 0; p < passes; ++p) {
        const int stage = p & 1;
        const uint32_t par = (uint32_t)(p >> 1) & 1u;
        if (mode & 2) {
          if (lane == 0) WAIT_A(full_a + stage * 8, par);
          __syncwarp();
        } else {
          WAIT_A(full_a + stage * 8, par);
        }
        ptx::tc_fence_after();
        if (mode & 1) {
          uint32_t v0[16], v1[16], v2[16], v3[16];
          const uint32_t taddr = tm_lane + stage * 256 + sub * 64;
          ptx::tmem_ld_32x32b_x16(taddr, v0);
          ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
          ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
          ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
          ptx::tmem_wait_ld();
          sink += v0[0] ^ v1[5] ^ v2[7] ^ v3[15];
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_a(empty_a + stage * 8);
      }
    } else {
      // four stages: warps 0-7 (sub 0,1) take the even passes, warps 8-15 the odd ones; a warp reads 64 columns
      const int grp = sub >> 1, half = sub & 1;
      for (int p = grp; p < passes; p += 2) {
        const int stage = p & 3;
        const uint32_t par = (uint32_t)(p >> 2) & 1u;
        if (mode & 2) {
          if (lane == 0) WAIT_A(full_a + stage * 8, par);
          __syncwarp();
        } else {
          WAIT_A(full_a + stage * 8, par);
        }
        ptx::tc_fence_after();
        if (mode & 1) {
          uint32_t v0[16], v1[16], v2[16], v3[16];
          const uint32_t taddr = tm_lane + stage * 128 + half * 64;
          ptx::tmem_ld_32x32b_x16(taddr, v0);
          ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
          ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
          ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
          ptx::tmem_wait_ld();
          sink += v0[0] ^ v1[5] ^ v2[7] ^ v3[15];
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_a(empty_a + stage * 8);
      }
    }
    if (sink == 0x12345678u) out[2] = sink;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

// Time-stamped hand-shake: the two-stage round trip of knn_tc2.cu (mode 1 of handshake_probe_kernel: MMAs, consumers
// read their slice) with %clock64 samples from the issuing thread and from one consumer warp, all on the same SM.
//   ts[p][0] issuer: `empty` seen      ts[p][1] issuer: MMAs + commit issued
//   ts[p][2] consumer (warp 2, lane 0): `full` seen      ts[p][3] consumer: slice read (wait::ld done)
//   ts[p][4] consumer: arrived on `empty`
__global__ void __launch_bounds__(576, 1) stamp_probe_kernel(int passes, long long* ts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base;
  for (int t = threadIdx.x; t < 2 * 256 * 32; t += blockDim.x)
    reinterpret_cast<__half*>(smem)[t] = __float2half(((t * 37) % 17 - 8) * 0.01f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(acc_full + s, 1);
      ptx::mbar_init(acc_empty + s, 16);
    }
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async();
  __syncthreads();
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  const bool rec = blockIdx.x == 0;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 256 * 64;
      uint64_t ad[2][2], bd[2];
      for (int mb = 0; mb < 2; ++mb)
        for (int ks = 0; ks < 2; ++ks) ad[mb][ks] = ptx::umma_desc_kmajor_swz(a0 + mb * 128 * 64 + ks * 32, 64, 512);
      for (int ks = 0; ks < 2; ++ks) bd[ks] = ptx::umma_desc_kmajor_swz(b0 + ks * 32, 64, 512);
      constexpr uint32_t id128 = ptx::umma_idesc(0u, 128u, 128u);
      const uint32_t full_a = ptx::smem_u32(acc_full), empty_a = ptx::smem_u32(acc_empty);
      for (int p = 0; p < passes; ++p) {
        const int stage = p & 1;
        const uint32_t par = (uint32_t)(p >> 1) & 1u;
        ptx::mbar_wait_a(empty_a + stage * 8, par ^ 1);
        ptx::tc_fence_after();
        if (rec) ts[(size_t)p * 8 + 0] = clock64();
        ptx::mma_f16_ss(tm + stage * 256, ad[0][0], bd[0], id128, 0);
        ptx::mma_f16_ss(tm + stage * 256, ad[0][1], bd[1], id128, 1);
        ptx::mma_f16_ss(tm + stage * 256 + 128, ad[1][0], bd[0], id128, 0);
        ptx::mma_f16_ss(tm + stage * 256 + 128, ad[1][1], bd[1], id128, 1);
        ptx::mma_commit_a(full_a + stage * 8);
        if (rec) ts[(size_t)p * 8 + 1] = clock64();
      }
      for (int s = 0; s < 2; ++s) {
        const int last = ((passes - 1 - s) / 2) * 2 + s;
        ptx::mbar_wait_a(empty_a + s * 8, (uint32_t)(last / 2) & 1u);
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int sub = e >> 2;
    const uint32_t tm_lane = tm + ((uint32_t)(quarter * 32) << 16);
    const uint32_t full_a = ptx::smem_u32(acc_full), empty_a = ptx::smem_u32(acc_empty);
    const bool stamp = rec && e == 0 && lane == 0;
    uint32_t sink = 0;
    for (int p = 0; p < passes; ++p) {
      const int stage = p & 1;
      const uint32_t par = (uint32_t)(p >> 1) & 1u;
      ptx::mbar_wait_a(full_a + stage * 8, par);
      ptx::tc_fence_after();
      if (stamp) ts[(size_t)p * 8 + 2] = clock64();
      uint32_t v0[16], v1[16], v2[16], v3[16];
      const uint32_t taddr = tm_lane + stage * 256 + sub * 64;
      ptx::tmem_ld_32x32b_x16(taddr, v0);
      ptx::tmem_ld_32x32b_x16(taddr + 16, v1);
      ptx::tmem_ld_32x32b_x16(taddr + 32, v2);
      ptx::tmem_ld_32x32b_x16(taddr + 48, v3);
      ptx::tmem_wait_ld();
      sink += v0[0] ^ v1[5] ^ v2[7] ^ v3[15];
      if (stamp) ts[(size_t)p * 8 + 3] = clock64();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(empty_a + stage * 8);
      if (stamp) ts[(size_t)p * 8 + 4] = clock64();
    }
    if (sink == 0x12345678u) ts[5] = sink;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int smem = 2 * 256 * 64 + 10 * 8192 + 1024;
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_out;
  CK(cudaMalloc(&d_out, 4 * sizeof(long long)));
  const char* names[6] = {"kNN pass: 2 acc x 2 k-steps, N=128, commit/pass", "same, no commits",
                          "one accumulator, accumulate chain, N=128", "N=256: 1 acc x 2 k-steps, commit/pass",
                          "kNN pass, both accumulators read the same A", "4 independent accumulators, K=16 each"};
  for (int rep = 0; rep < 2; ++rep)
    for (int pattern = 0; pattern < 6; ++pattern) {
      const int passes = 20000;
      mma_probe_kernel<<<prop.multiProcessorCount, 128, smem>>>(pattern, passes, d_out);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
      if (rep == 1)
        printf("pattern %d  %-52s  %8.1f cycles / MMA  (%lld MMAs, %d CTAs)\n", pattern, names[pattern],
               (double)h[0] / (double)h[1], h[1], prop.multiProcessorCount);
    }
  CK(cudaFuncSetAttribute(handshake_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int modes[] = {0, 1, 4, 16, 17, 32, 33, 35, 36, 48, 49, 51};
  const int n_tiles = 391;
  unsigned char* d_src;
  CK(cudaMalloc(&d_src, (size_t)n_tiles * 8192));
  CK(cudaMemset(d_src, 0, (size_t)n_tiles * 8192));
  for (int rep = 0; rep < 2; ++rep)
    for (int mode : modes) {
      const int passes = 20000;
      handshake_probe_kernel<<<prop.multiProcessorCount, 576, smem>>>(mode, passes, d_out, d_src, n_tiles);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
      if (rep == 1)
        printf("handshake mode %2d (%s%s%s%s)  %8.1f cycles / pass%s\n", mode, (mode & 1) ? "ld " : "",
               (mode & 2) ? "lane0-poll " : ((mode & 32) ? "spin " : ""), (mode & 4) ? "no-mma " : "", (mode & 8) ? "4x128 " : ((mode & 16) ? "2x256 +ring " : "2x256 "),
               (double)h[0] / (double)h[1], (mode & 8) ? "  (a pass is HALF a kNN tile pass)" : "");
    }
  {
    const int passes = 4096;
    long long* d_ts;
    CK(cudaMalloc(&d_ts, (size_t)passes * 8 * sizeof(long long)));
    CK(cudaMemset(d_ts, 0, (size_t)passes * 8 * sizeof(long long)));
    CK(cudaFuncSetAttribute(stamp_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    stamp_probe_kernel<<<prop.multiProcessorCount, 576, smem>>>(passes, d_ts);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h((size_t)passes * 8);
    CK(cudaMemcpy(h.data(), d_ts, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    // steady state: passes 1024 .. passes-3
    double d_issue = 0, d_commit_full = 0, d_read = 0, d_arrive = 0, d_empty_seen = 0, d_pass = 0;
    int n = 0;
    for (int p = 1024; p + 2 < passes; ++p, ++n) {
      const long long* t = &h[(size_t)p * 8];
      const long long* t2 = &h[(size_t)(p + 2) * 8];  // the next use of the same stage
      d_issue += (double)(t[1] - t[0]);        // empty seen -> MMAs + commit issued
      d_commit_full += (double)(t[2] - t[1]);  // commit issued -> consumer sees full (MMA execution + commit latency)
      d_read += (double)(t[3] - t[2]);         // tcgen05.ld x4 + wait::ld
      d_arrive += (double)(t[4] - t[3]);       // fence + syncwarp + arrive issued
      d_empty_seen += (double)(t2[0] - t[4]);  // this warp arrived -> issuer sees the stage empty (all 16 warps)
      d_pass += (double)(h[(size_t)(p + 1) * 8] - t[0]);
    }
    printf("stamped round trip (cycles, mean of %d passes): issue %.0f | commit->full seen %.0f | read %.0f | "
           "arrive %.0f | arrive->empty seen %.0f | pass %.0f\n",
           n, d_issue / n, d_commit_full / n, d_read / n, d_arrive / n, d_empty_seen / n, d_pass / n);
  }
  return 0;
}
