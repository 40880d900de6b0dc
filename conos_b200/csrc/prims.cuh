// Device-wide integer primitives the edge stages are built from (hand-written, no CUB):
//   exclusive_scan : 3-phase block scan (reduce / scan partials / downsweep), HBM-bound
//   segmented_sort : in-place ascending sort of 64-bit keys inside independent segments
//                    (warp bitonic for <= 128 keys, CTA bitonic in shared memory for <= 8192 keys,
//                     CTA bitonic in global memory beyond that)
#pragma once
#include "common.cuh"

namespace cg {

// out[i] = sum_{j<i} in[j]  (out may alias nothing; n may be 0). If d_total != nullptr it receives the sum.
void exclusive_scan_i32(Context& ctx, const int* d_in, int* d_out, size_t n, int* d_total);
void exclusive_scan_i32_to_i64(Context& ctx, const int* d_in, long long* d_out, size_t n, long long* d_total);

// keys[seg_off[s] .. seg_off[s+1]) sorted ascending, for s in [0, n_seg)
void segmented_sort_u64(Context& ctx, unsigned long long* d_keys, const long long* d_seg_off, int n_seg);

}  // namespace cg
