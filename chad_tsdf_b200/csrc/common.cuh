// Shared device helpers for the chad_tsdf B200 hot path: strict IEEE arithmetic wrappers,
// Morton bit interleave, compact sort keys, and the per-batch plan that lives in device memory.
//
// Float semantics: the reference is built with
// -ffp-contract=off and IEEE division / sqrt. Everything that feeds an observable result uses the
// explicit round-to-nearest intrinsics below so that nvcc can neither contract a*b+c into an FMA
// nor substitute an approximate reciprocal, whatever the compile flags are.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace chadgpu {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = unsigned long long;  // matches CUDA atomics
using i32 = int32_t;

// ---- strict fp32 / fp64 -------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
// glm::dot for 3-vectors, generic definition: (x*x' + y*y') + z*z'
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}
__device__ __forceinline__ double ddot3(double ax, double ay, double az, double bx, double by, double bz) {
    return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}
// std::clamp(v, lo, hi)
__device__ __forceinline__ float fclamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

// ---- Morton code (reference: include/chad/detail/morton.hpp:21-37; libmorton BMI2 pdep) ------
// Bit interleave of three 21-bit lanes; x -> bit 0, y -> bit 1, z -> bit 2; bias 2^20.
__host__ __device__ __forceinline__ u64 spread3(u64 v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__host__ __device__ __forceinline__ u32 compact3(u64 v) {
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
    v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
    v = (v | (v >> 8)) & 0x001F0000FF0000FFull;
    v = (v | (v >> 16)) & 0x001F00000000FFFFull;
    v = (v | (v >> 32)) & 0x1FFFFFull;
    return (u32)v;
}
__host__ __device__ __forceinline__ u64 morton_encode(i32 x, i32 y, i32 z) {
    u32 ux = (1u << 20) + (u32)x, uy = (1u << 20) + (u32)y, uz = (1u << 20) + (u32)z;
    return spread3(ux) | (spread3(uy) << 1) | (spread3(uz) << 2);
}
__host__ __device__ __forceinline__ void morton_decode(u64 key, i32& x, i32& y, i32& z) {
    x = (i32)(compact3(key) - (1u << 20));
    y = (i32)(compact3(key >> 1) - (1u << 20));
    z = (i32)(compact3(key >> 2) - (1u << 20));
}
// "range code" of a voxel coordinate: v fits in the signed range [-2^k, 2^k) iff rcode(v) < 2^k
__host__ __device__ __forceinline__ u32 rcode(i32 v) { return (u32)(v >= 0 ? v : ~v); }

// ---- compact sort keys --------------------------------------------------------------------
// For a batch whose voxel coordinates all lie in [-2^k, 2^k), bits k..19 of every biased axis value
// are the complement of bit 20, so the Morton triples k..19 carry no ordering information. The
// compact key keeps the low 3k bits and the top (bit-20) triple: 3k+3 bits, order-isomorphic to the
// full 63-bit key on that batch. This cuts the radix sort to ceil((3k+3)/8) passes (5 at k = 11,
// i.e. +-102 m at 5 cm voxels) without changing any observable order.
__host__ __device__ __forceinline__ u64 compact_key(u64 full, u32 k) {
    return (full & ((1ull << (3 * k)) - 1ull)) | ((full >> 60) << (3 * k));
}
__host__ __device__ __forceinline__ u64 expand_key(u64 ckey, u32 k) {
    const u64 low_mask = (1ull << (3 * k)) - 1ull;
    const u64 top = ckey >> (3 * k);  // bit-20 triple: x = bit 0, y = bit 1, z = bit 2
    // triples k..19 of one axis, all set
    const u64 axis_fill = 0x0249249249249249ull & ~low_mask;  // bits 3j, j < 20
    u64 full = (ckey & low_mask) | (top << 60);
    if (!(top & 1ull)) full |= axis_fill;
    if (!(top & 2ull)) full |= axis_fill << 1;
    if (!(top & 4ull)) full |= axis_fill << 2;
    return full;
}

// ---- per-batch plan (device resident; written by plan kernels, read by every later kernel) ----
enum : u32 {
    ERRF_RANGE = 1u,         // voxel coordinate outside the Morton range or outside the batch plan's k
    ERRF_KEY_BUDGET = 2u,    // 3k+3 + scan bits > 64
    ERRF_PAIR_CAPACITY = 4u, // more band voxels than the pair buffers hold
    ERRF_TABLE_FULL = 8u,    // resident chunk table over capacity
    ERRF_NUMERIC = 16u,      // NaN / Inf coordinate
    ERRF_DEDUP_FULL = 32u,   // DAG dedup table over capacity
    ERRF_BLOCKS_FULL = 64u,  // block table of the block-binned pair path over capacity (or > 8192 updates of ONE voxel in a batch)
    ERRF_EXCHANGE = 128u,    // Morton-range sharding: the runs for another rank did not fit the exchange buffer
};

struct BatchPlan {
    u32 rmax;          // max range code over the batch's point voxels (atomicMax)
    u32 k;             // compaction bits: every voxel touched by the batch lies in [-2^k, 2^k)
    u32 nbits_points;  // 3k + 3 + scan bits
    u32 nbits_pairs;   // 3k + 3
    u32 n_points;      // points in the batch
    u32 n_scans;
    u32 n_pairs;       // U of the batch (written by the offsets scan)
    u32 error;         // sticky ERRF_* flags
    u32 n_segments;    // distinct voxels in the batch (written by the segment count)
    u32 n_chunk_heads; // distinct leaf chunks in the batch (upper bound of the chunks the fold can insert)
    u32 n_new_chunks;  // chunks inserted by the fold
    u32 fold_ticket;   // dynamic work counter of the fold kernel
    u32 n_blocks;      // non-empty 8^3-voxel blocks of the batch (block-binned pair path)
    u32 sort_ticket;   // dynamic work counter of the per-block sort
    u32 n_runs;        // (tile, block) runs of the batch (tile-run pair path)
    u32 nbits_blocks;  // 3k - 6 + tile_bits: width of a run descriptor's sort key (compact 8^3-block id, tile)
    u32 tile_bits;     // bits of a 256-ray tile index inside the batch
    u32 n_big_blocks;  // tile-run path: blocks listed from the front of the work list (many updates) ...
    u32 n_small_blocks;//                ... and from its back
    u32 point_shift;   // 23 when the point sort carries the input index in the low bits of its key (keys-only sort), else 0
    u32 tsb;           // bits of a tile index inside ONE scan (tile_bits = scan bits + Morton-range rank bits + tsb)
    u32 n_batch;       // points of the whole batch (== n_points on a single GPU; a Morton-range shard sorts only its own n_points of them)
    u32 tail_lo, tail_hi;  // bit s set: this context holds the lowest-key point of scan s, which normals.hpp:100 never absorbs into a neighbourhood
    u32 n_runs_local;  // sharded: runs / records of this rank's own walk (before the runs received from the other ranks are appended)
    u32 n_pairs_local;
    u32 xfer_runs;     // sharded: runs / records this rank sent to other ranks
    u32 xfer_records;
};
constexpr u32 POINT_INDEX_BITS = 23;

constexpr int MAX_BATCH_SCANS = 64;
struct BatchScans {  // device resident
    u32 offset[MAX_BATCH_SCANS + 1];  // point offsets of the scans inside the batch
    float pose[MAX_BATCH_SCANS][3];
    u32 tile_prefix[MAX_BATCH_SCANS + 1];  // 256-ray tiles of the ray walk before scan s: tiles never straddle scans, so that
                                           // (scan, Morton-range rank, tile in scan) orders the runs of a block across GPUs
};
constexpr u32 RAY_TILE = 256;
// host or device: fill tile_prefix from offset
__host__ __device__ inline void batch_scans_tiles(BatchScans& sc, u32 n_scans) {
    u32 t = 0;
    for (u32 s = 0; s < n_scans; s++) { sc.tile_prefix[s] = t; t += (sc.offset[s + 1] - sc.offset[s] + RAY_TILE - 1) / RAY_TILE; }
    for (u32 s = n_scans; s <= MAX_BATCH_SCANS; s++) sc.tile_prefix[s] = t;
}

// scan id of batch point index i (n_scans <= 64: branch-free-ish binary search)
__device__ __forceinline__ u32 scan_of(const BatchScans* __restrict__ sc, u32 n_scans, u32 i) {
    u32 lo = 0, hi = n_scans;  // offset[lo] <= i < offset[hi]
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (sc->offset[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}


// optional instrumentation around kernel launches (bench.py's per-kernel CUDA-event timing)
struct LaunchHook {
    void* user;
    void (*begin)(void* user, int cls);
    void (*end)(void* user);
};

}  // namespace chadgpu
