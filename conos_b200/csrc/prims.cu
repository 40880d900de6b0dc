#include "prims.cuh"

namespace cg {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem_warp, T& block_total) {
  // inclusive scan inside the warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) smem_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < SCAN_THREADS / 32 ? smem_warp[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    if (lane < SCAN_THREADS / 32) smem_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == SCAN_THREADS / 32 - 1) smem_warp[SCAN_THREADS / 32] = wi;
  }
  __syncthreads();
  block_total = smem_warp[SCAN_THREADS / 32];
  return incl - v + smem_warp[warp];
}

template <class Tin, class Tout>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const Tin* __restrict__ in, size_t n,
                                                                   Tout* __restrict__ partial) {
  __shared__ Tout sw[SCAN_THREADS / 32 + 1];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  Tout s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += (Tout)in[base + i];
  Tout total;
  block_exclusive_scan<Tout>(s, sw, total);
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

template <class Tin, class Tout>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const Tin* __restrict__ in, size_t n,
                                                                 const Tout* __restrict__ partial_excl,
                                                                 Tout* __restrict__ out, Tout* __restrict__ total_out) {
  __shared__ Tout sw[SCAN_THREADS / 32 + 1];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  Tout v[SCAN_ITEMS];
  Tout s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? (Tout)in[base + i] : Tout(0);
    s += v[i];
  }
  Tout total;
  Tout off = block_exclusive_scan<Tout>(s, sw, total) + (partial_excl ? partial_excl[blockIdx.x] : Tout(0));
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = off;
}

template <class Tin, class Tout>
void scan_impl(Context& ctx, const Tin* d_in, Tout* d_out, size_t n, Tout* d_total) {
  if (n == 0) {
    if (d_total) CG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(Tout), ctx.stream));
    return;
  }
  size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb == 1) {
    scan_down_kernel<Tin, Tout><<<1, SCAN_THREADS, 0, ctx.stream>>>(d_in, n, nullptr, d_out, d_total);
    ctx.stats.kernels++;
    CG_CUDA(cudaGetLastError());
    return;
  }
  DevBuf<Tout> partial(nb), partial_excl(nb);
  scan_reduce_kernel<Tin, Tout><<<(unsigned)nb, SCAN_THREADS, 0, ctx.stream>>>(d_in, n, partial.p);
  ctx.stats.kernels++;
  scan_impl<Tout, Tout>(ctx, partial.p, partial_excl.p, nb, nullptr);
  scan_down_kernel<Tin, Tout><<<(unsigned)nb, SCAN_THREADS, 0, ctx.stream>>>(d_in, n, partial_excl.p, d_out, d_total);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
  // no host wait: the partial buffers are released in stream order (cudaFreeAsync on the stream they were used on)
}

// ------------------------------------------------------------------ segmented bitonic sort
typedef unsigned long long u64;

__device__ __forceinline__ void bitonic_smem(u64* s, int n_pow2, int tid, int nthreads, bool warp_only) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (n_pow2 >> 1); t += nthreads) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        u64 a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      if (warp_only) __syncwarp(); else __syncthreads();
    }
  }
}

constexpr int SEG_WARP_MAX = 128;
constexpr int SEG_CTA_MAX = 8192;

// one warp per segment, segments longer than SEG_WARP_MAX are left to the CTA kernel
__global__ void __launch_bounds__(256) segsort_warp_kernel(u64* __restrict__ keys, const long long* __restrict__ seg_off,
                                                           int n_seg) {
  __shared__ u64 sm[8][SEG_WARP_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.x * 8 + warp;
  if (seg >= n_seg) return;
  const long long b = seg_off[seg];
  const int len = (int)(seg_off[seg + 1] - b);
  if (len <= 1 || len > SEG_WARP_MAX) return;
  int np = 2;
  while (np < len) np <<= 1;
  u64* s = sm[warp];
  for (int i = lane; i < np; i += 32) s[i] = i < len ? keys[b + i] : ~0ull;
  __syncwarp();
  bitonic_smem(s, np, lane, 32, true);
  for (int i = lane; i < len; i += 32) keys[b + i] = s[i];
}

// one CTA per long segment (index list), shared memory up to SEG_CTA_MAX keys, global memory beyond
__global__ void __launch_bounds__(512) segsort_cta_kernel(u64* __restrict__ keys, const long long* __restrict__ seg_off,
                                                          const int* __restrict__ long_segs) {
  extern __shared__ __align__(16) unsigned char raw[];
  u64* s = reinterpret_cast<u64*>(raw);
  const int seg = long_segs[blockIdx.x];
  const long long b = seg_off[seg];
  const long long len = seg_off[seg + 1] - b;
  if (len <= SEG_CTA_MAX) {
    int np = 2;
    while (np < len) np <<= 1;
    for (int i = threadIdx.x; i < np; i += blockDim.x) s[i] = i < len ? keys[b + i] : ~0ull;
    __syncthreads();
    bitonic_smem(s, np, threadIdx.x, blockDim.x, false);
    for (int i = threadIdx.x; i < len; i += blockDim.x) keys[b + i] = s[i];
  } else {
    // global-memory bitonic network over a virtual power-of-two length. All compare-exchanges are
    // ascending (the first step of every merge pairs i with its mirror in the block), so the virtual
    // +inf padding beyond `len` never has to move and can simply be skipped.
    long long np = 2;
    while (np < len) np <<= 1;
    u64* k = keys + b;
    for (long long size = 2; size <= np; size <<= 1) {
      const long long half = size >> 1;
      for (long long t = threadIdx.x; t < (np >> 1); t += blockDim.x) {
        long long blk = t / half, off = t - blk * half;
        long long lo = blk * size + off, hi = blk * size + size - 1 - off;
        if (hi < len) {
          u64 x = k[lo], y = k[hi];
          if (x > y) { k[lo] = y; k[hi] = x; }
        }
      }
      __syncthreads();
      for (long long stride = half >> 1; stride > 0; stride >>= 1) {
        for (long long t = threadIdx.x; t < (np >> 1); t += blockDim.x) {
          long long lo = 2 * t - (t & (stride - 1));
          long long hi = lo + stride;
          if (hi < len) {
            u64 x = k[lo], y = k[hi];
            if (x > y) { k[lo] = y; k[hi] = x; }
          }
        }
        __syncthreads();
      }
    }
  }
}

__global__ void collect_long_segments_kernel(const long long* __restrict__ seg_off, int n_seg, int* __restrict__ list,
                                             int* __restrict__ count) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  if (seg_off[s + 1] - seg_off[s] > SEG_WARP_MAX) list[atomicAdd(count, 1)] = s;
}

}  // namespace

void exclusive_scan_i32(Context& ctx, const int* d_in, int* d_out, size_t n, int* d_total) {
  scan_impl<int, int>(ctx, d_in, d_out, n, d_total);
}
void exclusive_scan_i32_to_i64(Context& ctx, const int* d_in, long long* d_out, size_t n, long long* d_total) {
  scan_impl<int, long long>(ctx, d_in, d_out, n, d_total);
}

void segmented_sort_u64(Context& ctx, unsigned long long* d_keys, const long long* d_seg_off, int n_seg) {
  if (n_seg <= 0) return;
  segsort_warp_kernel<<<(n_seg + 7) / 8, 256, 0, ctx.stream>>>(d_keys, d_seg_off, n_seg);
  ctx.stats.kernels++;
  DevBuf<int> list(n_seg), count(1);
  count.zero(ctx.stream);
  collect_long_segments_kernel<<<(n_seg + 255) / 256, 256, 0, ctx.stream>>>(d_seg_off, n_seg, list.p, count.p);
  ctx.stats.kernels++;
  int h_count = 0;
  count.download(&h_count, 1, ctx.stream);
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
  if (h_count > 0) {
    size_t smem = (size_t)SEG_CTA_MAX * sizeof(u64);
    CG_CUDA(cudaFuncSetAttribute(segsort_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    segsort_cta_kernel<<<h_count, 512, smem, ctx.stream>>>(d_keys, d_seg_off, list.p);
    ctx.stats.kernels++;
    CG_CUDA(cudaGetLastError());   // `list` is released in stream order
  }
}

}  // namespace cg
