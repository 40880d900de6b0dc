// Probe for fp16 accumulators of tcgen05.mma.kind::f16 and the packed tensor-memory load (tcgen05.ld ... .pack::16b):
//   part 1: where does a 128 x 128 fp16 accumulator live in tensor memory, and how far is it from the fp32 one?
//   part 2: how fast do 4 / 8 warps drain 512 columns with plain x32 loads and with packed x32 loads (64 columns each)?
// Question behind it: the kNN tile pass is bound by draining 32 768 fp32 scores out of tensor memory
// (512 cycles) -- would 16-bit scores drain in half the time?
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I conos_b200/csrc tests/cuda/acc16_probe.cu -o build/acc16_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <vector>

#include "ptx.cuh"

using namespace cg;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

__device__ __forceinline__ void tmem_ld_pack_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- part 1
// out32: [128 lanes][128] fp32 accumulator; raw16: [128 lanes][128] raw words of the fp16 accumulator's columns;
// packed: [128 lanes][64] words of two packed loads over the same columns
__global__ void __launch_bounds__(128) layout_kernel(float* out32, uint32_t* raw16, uint32_t* packed, float* one32, uint32_t* one16, float thr) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base;
  __half* sA = reinterpret_cast<__half*>(smem);             // 128 rows x 32 halves (SWIZZLE_64B image)
  __half* sB = reinterpret_cast<__half*>(smem + 128 * 64);  // 128 rows x 32 halves
  // rows that look alike (a common direction of norm^2 ~ 0.70 plus noise), as neighbours do: with the kNN kernel's
  // trick -- last K column of A is -1, of B the threshold -- the accumulator holds score - thr, near zero for many
  for (int t = threadIdx.x; t < 2 * 128 * 32; t += blockDim.x) {
    const int row = (t >> 5) & 127, chunk_sw = (t >> 3) & 3, e = t & 7;
    const int k = ((chunk_sw ^ ((row >> 1) & 3)) << 3) | e;   // logical K index of this shared-memory slot
    uint32_t h = (uint32_t)t * 2654435761u;
    h ^= h >> 15;
    float v = ((k * 7) % 3 == 0 ? -0.15f : 0.15f) + ((int)(h % 2001u) - 1000) * 3e-5f;
    if (thr == 0.f) v = ((int)(h % 2001u) - 1000) * 1.8e-4f;
    if (k == 31) v = thr == 0.f ? v : (t < 128 * 32 ? -1.f : thr);
    reinterpret_cast<__half*>(smem)[t] = __float2half(v);
  }
  if (threadIdx.x == 0) {
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
  // clear the region the fp16 accumulator may touch, so that untouched halves read as 0xdead
  {
    uint32_t fill[16];
    for (int i = 0; i < 16; ++i) fill[i] = 0xdeaddeadu;
    const uint32_t lane_base = (threadIdx.x / 32) * 32;
    for (int c = 128; c < 384; c += 16) ptx::tmem_st_32x32b_x16(tm + (lane_base << 16) + c, fill);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a0 = ptx::smem_u32(sA), b0 = ptx::smem_u32(sB);
    constexpr uint32_t id32 = ptx::umma_idesc(0u, 128u, 128u);
    constexpr uint32_t id16 = id32 & ~(3u << 4);  // c_format 0 = F16
    for (int ks = 0; ks < 2; ++ks)
      ptx::mma_f16_ss(tm, ptx::umma_desc_kmajor_swz(a0 + ks * 32, 64, 512), ptx::umma_desc_kmajor_swz(b0 + ks * 32, 64, 512),
                      id32, ks);
    for (int ks = 0; ks < 2; ++ks)
      ptx::mma_f16_ss(tm + 128, ptx::umma_desc_kmajor_swz(a0 + ks * 32, 64, 512),
                      ptx::umma_desc_kmajor_swz(b0 + ks * 32, 64, 512), id16, ks);
    // one k-step only (K = 16): what is the rounding of the fp16 accumulator?
    ptx::mma_f16_ss(tm + 256, ptx::umma_desc_kmajor_swz(a0, 64, 512), ptx::umma_desc_kmajor_swz(b0, 64, 512), id32, 0);
    ptx::mma_f16_ss(tm + 384, ptx::umma_desc_kmajor_swz(a0, 64, 512), ptx::umma_desc_kmajor_swz(b0, 64, 512), id16, 0);
    ptx::mma_commit_a(ptx::smem_u32(&bar_done));
  }
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after();
  const uint32_t lane_base = (threadIdx.x / 32) * 32, row = threadIdx.x;
  uint32_t r[32];
  for (int c = 0; c < 128; c += 32) {
    ptx::tmem_ld_32x32b_x32(tm + (lane_base << 16) + 256 + c, r);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) one32[row * 128 + c + i] = __uint_as_float(r[i]);
    ptx::tmem_ld_32x32b_x32(tm + (lane_base << 16) + 384 + c, r);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) one16[row * 128 + c + i] = r[i];
  }
  for (int c = 0; c < 128; c += 32) {
    ptx::tmem_ld_32x32b_x32(tm + (lane_base << 16) + c, r);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) out32[row * 128 + c + i] = __uint_as_float(r[i]);
  }
  for (int c = 0; c < 128; c += 32) {
    ptx::tmem_ld_32x32b_x32(tm + (lane_base << 16) + 128 + c, r);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) raw16[row * 128 + c + i] = r[i];
  }
  for (int c = 0; c < 2; ++c) {
    tmem_ld_pack_x32(tm + (lane_base << 16) + 128 + c * 64, r);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) packed[row * 64 + c * 32 + i] = r[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------- part 2
// mode 0: plain x32 loads, 16 per 512-column sweep; mode 1: packed x32 loads, 8 per sweep; mode 2: plain x16 (32 per
// sweep); +4: one wait per two loads instead of per load.  ops: 0 = xor only, 1 = max (fp32 FMNMX / half2 HMNMX2)
__global__ void __launch_bounds__(256) drain_kernel(int mode, int ops, int sweeps, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base;
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  const int warp = threadIdx.x / 32, n_warps = blockDim.x / 32;
  const uint32_t lane_base = (warp % 4) * 32;
  // with 8 warps the two warps of a sub-partition take one column half each (as the kNN epilogue does)
  const int c0 = (n_warps == 8) ? (warp / 4) * 256 : 0, c1 = (n_warps == 8) ? c0 + 256 : 512;
  const bool two = (mode & 4) != 0;
  const int m = mode & 3;
  uint32_t acc = 0;
  float fm = -1e30f;
  __half2 hm = __float2half2_rn(-60000.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int s = 0; s < sweeps; ++s) {
    if (m == 0) {
      uint32_t r[32], q[32];
      for (int c = c0; c < c1; c += 64) {
        ptx::tmem_ld_32x32b_x32(tm + (lane_base << 16) + c, r);
        if (!two) tmem_wait_ld();
        ptx::tmem_ld_32x32b_x32(tm + (lane_base << 16) + c + 32, q);
        tmem_wait_ld();
        if (ops == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= r[i] ^ q[i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) fm = fmaxf(fm, fmaxf(__uint_as_float(r[i]), __uint_as_float(q[i])));
        }
      }
    } else if (m == 1) {
      uint32_t r[32], q[32];
      for (int c = c0; c < c1; c += 128) {
        tmem_ld_pack_x32(tm + (lane_base << 16) + c, r);
        if (!two) tmem_wait_ld();
        tmem_ld_pack_x32(tm + (lane_base << 16) + c + 64, q);
        tmem_wait_ld();
        if (ops == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= r[i] ^ q[i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            hm = __hmax2(hm, __hmax2(*reinterpret_cast<__half2*>(&r[i]), *reinterpret_cast<__half2*>(&q[i])));
        }
      }
    } else {
      uint32_t r[16], q[16];
      for (int c = c0; c < c1; c += 32) {
        ptx::tmem_ld_32x32b_x16(tm + (lane_base << 16) + c, r);
        if (!two) tmem_wait_ld();
        ptx::tmem_ld_32x32b_x16(tm + (lane_base << 16) + c + 16, q);
        tmem_wait_ld();
        if (ops == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc ^= r[i] ^ q[i];
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) fm = fmaxf(fm, fmaxf(__uint_as_float(r[i]), __uint_as_float(q[i])));
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(fm) ^ *reinterpret_cast<uint32_t*>(&hm);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------- part 3
// issue rate of the candidate epilogue instructions: 16 warps per SM, 8 independent chains per thread
template <int OP>
__global__ void __launch_bounds__(512) rate_kernel(int iters, long long* cycles, uint32_t* sink, uint32_t seed) {
  uint32_t x[8], y = seed * 3 + threadIdx.x, z = seed * 5 + 1;
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = __vimax3_s16x2(x[i], y, z);
      if (OP == 1) x[i] = __float_as_uint(fmaxf(fmaxf(__uint_as_float(x[i]), __uint_as_float(y)), __uint_as_float(z)));
      if (OP == 2) x[i] = __funnelshift_l(y, x[i], 1);
      if (OP == 3) {
        __half2 h = __hfma2_sat(*reinterpret_cast<__half2*>(&x[i]), *reinterpret_cast<__half2*>(&y), *reinterpret_cast<__half2*>(&z));
        x[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      if (OP == 4) x[i] = x[i] & y & z;          // LOP3
      if (OP == 5) x[i] = __byte_perm(x[i], y, 0x7531);
      if (OP == 6) {
        __half2 h = __hmax2(*reinterpret_cast<__half2*>(&x[i]), *reinterpret_cast<__half2*>(&y));
        x[i] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    y += 0x10001u;
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t a = 0;
  for (int i = 0; i < 8; ++i) a ^= x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

static float half_bits(uint16_t h) {
  __half_raw r;
  r.x = h;
  return __half2float(__half(r));
}

int main() {
  float* d32;
  uint32_t *draw, *dpk;
  CK(cudaMalloc(&d32, 128 * 128 * 4));
  CK(cudaMalloc(&draw, 128 * 128 * 4));
  CK(cudaMalloc(&dpk, 128 * 64 * 4));
  float* done32;
  uint32_t* done16;
  CK(cudaMalloc(&done32, 128 * 128 * 4));
  CK(cudaMalloc(&done16, 128 * 128 * 4));
  for (float thr : {0.f, 0.70f}) {
    layout_kernel<<<1, 128, 2 * 128 * 64 + 1024>>>(d32, draw, dpk, done32, done16, thr);
    CK(cudaDeviceSynchronize());
    std::vector<float> h32(128 * 128);
    std::vector<uint32_t> hraw(128 * 128), hpk(128 * 64);
    CK(cudaMemcpy(h32.data(), d32, h32.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hraw.data(), draw, hraw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hpk.data(), dpk, hpk.size() * 4, cudaMemcpyDeviceToHost));
    // hypothesis U: column c of the fp16 accumulator = low half of raw column c (one value per 32-bit column)
    // hypothesis P: raw column c holds values (2c, 2c+1) in (low, high)
    double errU = 0, errP = 0, errPk = 0, max32 = 0, errNear = 0;
    int untouched_hi = 0, untouched_cols = 0, sign_flip = 0, near = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 128; ++c) {
        const float f = h32[r * 128 + c];
        max32 = std::max(max32, (double)fabsf(f));
        const float u = half_bits((uint16_t)(hraw[r * 128 + c] & 0xffff));
        errU = std::max(errU, (double)fabsf(u - f));
        if ((hraw[r * 128 + c] >> 16) == 0xdead) ++untouched_hi;
        if (hraw[r * 128 + c] == 0xdeaddeadu) ++untouched_cols;
        if (c < 64) {
          const uint32_t w = hraw[r * 128 + c];
          errP = std::max(errP, (double)fabsf(half_bits((uint16_t)(w & 0xffff)) - h32[r * 128 + 2 * c]));
          errP = std::max(errP, (double)fabsf(half_bits((uint16_t)(w >> 16)) - h32[r * 128 + 2 * c + 1]));
        }
        // packed load: word j of the first load = columns (2j, 2j+1)?
        const uint32_t w = hpk[r * 64 + c / 2];
        const float pk = half_bits((uint16_t)((c & 1) ? (w >> 16) : (w & 0xffff)));
        errPk = std::max(errPk, (double)fabsf(pk - f));
        if (fabsf(f) < 2e-3f) {
          ++near;
          errNear = std::max(errNear, (double)fabsf(u - f));
          if ((u < 0) != (f < 0) && f != 0.f) ++sign_flip;
        }
      }
    {
      std::vector<float> o32(128 * 128);
      std::vector<uint32_t> o16(128 * 128);
      CK(cudaMemcpy(o32.data(), done32, o32.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(o16.data(), done16, o16.size() * 4, cudaMemcpyDeviceToHost));
      int rn = 0, rz = 0, rd = 0, ru = 0, other = 0, inexact = 0;
      double worst = 0;
      for (size_t i = 0; i < o32.size(); ++i) {
        const float f = o32[i];
        const uint16_t got = (uint16_t)(o16[i] & 0xffff);
        const uint16_t n = __half_as_ushort(__float2half_rn(f)), z = __half_as_ushort(__float2half_rz(f)),
                       dn = __half_as_ushort(__float2half_rd(f)), up = __half_as_ushort(__float2half_ru(f));
        if (dn == up) continue;  // exactly representable
        ++inexact;
        rn += got == n;
        rz += got == z;
        rd += got == dn;
        ru += got == up;
        other += !(got == dn || got == up);
        worst = std::max(worst, (double)fabsf(half_bits(got) - f));
      }
      printf("thr=%.2f  one k-step, %d inexact values: matches RN %d, RZ %d, RD %d, RU %d, neither neighbour %d, max err %.3e\n",
             thr, inexact, rn, rz, rd, ru, other, worst);
    }
    printf("thr=%.2f  max|fp32 acc|=%.4f\n", thr, max32);
    printf("  one value per column (low half):      max err %.3e   high halves untouched %d / 16384, columns untouched %d\n",
           errU, untouched_hi, untouched_cols);
    printf("  two values per column (packed in TMEM): max err %.3e\n", errP);
    printf("  packed load word j = columns (2j,2j+1): max err %.3e\n", errPk);
    printf("  scores within 2e-3 of zero: %d, max err there %.3e, sign flips %d\n", near, errNear, sign_flip);
  }
  // ---- part 2
  long long* dcy;
  uint32_t* dsink;
  CK(cudaMalloc(&dcy, 8));
  CK(cudaMalloc(&dsink, 148 * 512 * 4));
  const int sweeps = 2000;
  const char* names[] = {"plain x32 (32 cols/load)", "packed x32 (64 cols/load)", "plain x16 (16 cols/load)"};
  for (int warps : {4, 8})
    for (int ops : {0, 1})
      for (int two : {0, 4})
        for (int m = 0; m < 3; ++m) {
          drain_kernel<<<148, warps * 32>>>(m | two, ops, sweeps, dcy, dsink);
          CK(cudaDeviceSynchronize());
          long long cy;
          CK(cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost));
          printf("drain %d warps, %-26s wait per %s, %s: %7.1f cycles per 512-column sweep of 128 lanes\n", warps, names[m],
                 two ? "2 loads" : "load  ", ops ? "max" : "xor", (double)cy / sweeps);
        }
  {
    const char* nm[] = {"VIMNMX3.S16x2", "FMNMX3", "SHF.L (funnel)", "HFMA2.SAT", "LOP3", "PRMT", "HMNMX2"};
    const int iters = 4096;
    for (int op = 0; op < 7; ++op) {
      switch (op) {
        case 0: rate_kernel<0><<<148, 512>>>(iters, dcy, dsink, 7u); break;
        case 1: rate_kernel<1><<<148, 512>>>(iters, dcy, dsink, 7u); break;
        case 2: rate_kernel<2><<<148, 512>>>(iters, dcy, dsink, 7u); break;
        case 3: rate_kernel<3><<<148, 512>>>(iters, dcy, dsink, 7u); break;
        case 4: rate_kernel<4><<<148, 512>>>(iters, dcy, dsink, 7u); break;
        case 5: rate_kernel<5><<<148, 512>>>(iters, dcy, dsink, 7u); break;
        default: rate_kernel<6><<<148, 512>>>(iters, dcy, dsink, 7u); break;
      }
      CK(cudaDeviceSynchronize());
      long long cy;
      CK(cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost));
      // 16 warps x 8 ops per iteration = 128 warp instructions per SM = 32 per SMSP
      printf("rate %-16s %.2f cycles per warp instruction per SMSP\n", nm[op], (double)cy / iters / 32.0);
    }
  }
  return 0;
}
