// Onesweep-style least-significant-digit radix sort of (u64 key, u32 value) pairs, stable,
// ascending on key bits [0, nbits), hand-written for sm_100a.
//
// Replaces, in the reference, std::sort in sort_morton_vector (include/chad/detail/morton.hpp:81-102)
// and the implicit per-voxel grouping that the pointer octree performs one leaf at a time
// (include/chad/detail/octree.hpp:31-78,153-164). Stability is what keeps the reference's update
// order (sorted-point rank, then ray step) inside each voxel segment.
//
// Structure (one upfront histogram kernel + one kernel per 8-bit digit):
//   * radix_histogram: every CTA accumulates 256-bin histograms for all active digits in shared
//     memory (warp-aggregated with match.any, because Morton-coherent inputs make whole warps hit
//     one bin) and adds them to the global histograms.
//   * radix_onesweep_pass: persistent CTAs take tiles of 2048 keys in order from an atomic ticket,
//     rank keys inside the tile with warp-level match.any multi-split (stable), publish the tile's
//     digit counts in a status word per (tile, digit) and resolve their global prefix by decoupled
//     look-back over earlier tiles, stage the tile in shared memory in sorted order and write
//     every digit run out coalesced.
// `n` and `nbits` are read from device memory (they are produced by earlier kernels of the same
// stream), so no host synchronisation is needed between enumeration and sort; passes beyond
// ceil(nbits/8) exit immediately and the result buffer parity is radix_result_in_alt(nbits).
#pragma once
#include "common.cuh"

namespace chadgpu {

constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 256;
constexpr int RS_THREADS = 256;  // == RS_RADIX: thread t owns digit t in scans and look-back
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys per tile
constexpr int RS_CTAS_PER_SM = 4;             // occupancy target: the ranking loop is latency bound (ncu r01: 20 % issue active at 2 CTAs/SM)
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 8;
constexpr u32 RS_FLAG_AGG = 1u << 30;
constexpr u32 RS_FLAG_PREFIX = 2u << 30;
constexpr u32 RS_FLAG_MASK = 3u << 30;
constexpr u32 RS_VALUE_MASK = ~RS_FLAG_MASK;
constexpr size_t RS_SMEM_BYTES = size_t(RS_TILE) * (8 + 4);

struct RadixWorkspace {
    u32* hist;          // [RS_MAX_PASSES][256]
    u32* tile_counter;  // [RS_MAX_PASSES]
    u32* lookback[2];   // [max_tiles][256] each, alternating by pass parity
    size_t max_tiles;
};
inline size_t radix_workspace_bytes(size_t max_n) {
    size_t tiles = (max_n + RS_TILE - 1) / RS_TILE + 1;
    return (size_t(RS_MAX_PASSES) * 256 + 64) * 4 + 2 * tiles * 256 * 4;
}
// carve a workspace out of `mem` (radix_workspace_bytes(max_n) bytes, 256-byte aligned)
inline RadixWorkspace radix_workspace_carve(void* mem, size_t max_n) {
    RadixWorkspace ws;
    size_t tiles = (max_n + RS_TILE - 1) / RS_TILE + 1;
    u32* p = static_cast<u32*>(mem);
    ws.hist = p;
    ws.tile_counter = p + RS_MAX_PASSES * 256;
    ws.lookback[0] = p + RS_MAX_PASSES * 256 + 64;
    ws.lookback[1] = ws.lookback[0] + tiles * 256;
    ws.max_tiles = tiles;
    return ws;
}
__host__ __device__ __forceinline__ u32 radix_num_passes(u32 nbits) { return (nbits + RS_RADIX_BITS - 1) / RS_RADIX_BITS; }
// after sorting `nbits` bits starting in the primary buffers, is the result in the alternate buffers?
__host__ __device__ __forceinline__ bool radix_result_in_alt(u32 nbits) { return (radix_num_passes(nbits) & 1u) != 0; }

// Queue the sort on `stream`. keys/vals: primary buffers (input); keys_alt/vals_alt: same size.
// d_n / d_nbits: device pointers. max_n: host upper bound of *d_n (sizes the grids / workspace).
// max_passes: host upper bound on ceil(nbits/8) (<= 8). Returns the number of kernels launched.
// hook/cls_base: optional instrumentation; class cls_base = histogram, cls_base + 1 + p = pass p.
int radix_sort_pairs(cudaStream_t stream, u64* keys, u32* vals, u64* keys_alt, u32* vals_alt, const u32* d_n,
                     const u32* d_nbits, size_t max_n, int max_passes, const RadixWorkspace& ws, int num_sms,
                     const LaunchHook* hook = nullptr, int cls_base = 0, const u32* d_shift0 = nullptr);
// d_shift0 (optional, device): when *d_shift0 != 0 the sorted bits are [*d_shift0, *d_shift0 + nbits) of the key and the bits
// below carry the payload (keys-only sort: the value arrays are not touched).
cudaError_t radix_sort_init();  // opt in to > 48 KB dynamic shared memory

}  // namespace chadgpu
