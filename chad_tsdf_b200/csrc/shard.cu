// Morton-range sharding of ONE map across GPUs (north_star: "the map is partitioned by contiguous
// Morton ranges, so each GPU owns a spatial shard"): the point-stage side. Rank g owns the 8x8x8-voxel blocks whose id
// (Morton key >> 9) lies in [splitters[g], splitters[g + 1]). Every rank receives every scan, but sorts, estimates normals
// for and walks only the points whose voxel lies in its own range:
//   * a neighbourhood of estimate_normals (/root/reference/include/chad/detail/normals.hpp:94-108) never leaves a 4x4x4-voxel
//     block, so it never straddles two ranges; the one global exception -- the scan's lowest-key point is never absorbed
//     (:100) -- is kept by telling each rank whether it holds that point (BatchPlan::tail_*);
//   * the canonical point order (descending Morton, ties by input index, morton.hpp:85-89) restricted to a range is the order
//     of that range's points, so the filter below compacts the batch IN INPUT ORDER and the stable sort does the rest;
//   * the band voxels a ray adds outside its rank's range (a few per thousand) travel to their owner as whole tile runs
//     (runs.cu: runs_pack_kernel / runs_ingest_kernel); the order key (scan | range rank, descending | tile) of a run makes the
//     receiver's descriptor sort reproduce the reference's update order (octree.hpp:153-164).
// Kernels: splitters from a sorted sample of a submap's first scan; per-batch ownership count (fused with the plan's bounding
// box) -> one-block scan -> ordered scatter of the owned points' sort keys.
#include "kernels.cuh"
#include "points.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace chadgpu {

namespace {

constexpr int SH_SCAN_THREADS = 1024;

__device__ __forceinline__ u32 owner_of(u64 blk, const u64* __restrict__ splitters, u32 world) {
    u32 g = 0;
    for (u32 q = 1; q < world; q++) g += (splitters[q] <= blk) ? 1u : 0u;  // splitters ascend: the number of range starts at or below blk
    return g;
}

// Range starts of a new submap from its first scan, in ONE block: 4096 evenly spaced points of the scan -> block ids (full Morton key
// >> 9), bitonic sort in shared memory, quantiles. first_share_256: share of rank 0 relative to the others' 256 (rank 0 also builds
// the DAG of every closed submap, so it may be given fewer rays). Every rank runs this on the same points: identical ranges, no
// communication.
constexpr u32 SH_SAMPLES = 4096;
__global__ void __launch_bounds__(SH_SCAN_THREADS) shard_splitters_kernel(const float* __restrict__ xyz, u32 n_points, float recip, u32 world,
                                                                          u32 first_share_256, u64* __restrict__ splitters) {
    __shared__ u64 s_key[SH_SAMPLES];
    const u32 ns = min(n_points, SH_SAMPLES);
    for (u32 j = threadIdx.x; j < SH_SAMPLES; j += SH_SCAN_THREADS) {
        u64 key = ~0ull;  // padding sorts last
        if (j < ns) {
            const size_t i = (size_t)((u64)j * n_points / ns);
            i32 vx, vy, vz;
            voxel_of(__ldg(&xyz[i * 3]), __ldg(&xyz[i * 3 + 1]), __ldg(&xyz[i * 3 + 2]), recip, vx, vy, vz);  // (out of range -> origin; reported by the count kernel)
            key = morton_encode(vx, vy, vz) >> 9;
        }
        s_key[j] = key;
    }
    __syncthreads();
    for (u32 k = 2; k <= SH_SAMPLES; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = threadIdx.x; t < SH_SAMPLES; t += SH_SCAN_THREADS) {
                const u32 partner = t ^ j;
                if (partner > t) {
                    const u64 a = s_key[t], b = s_key[partner];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { s_key[t] = b; s_key[partner] = a; }
                }
            }
            __syncthreads();
        }
    }
    const u32 g = threadIdx.x;
    if (g > world) return;
    u64 v;
    if (g == 0) v = 0ull;
    else if (g == world || ns == 0) v = ~0ull;  // (nothing to go by: everything belongs to rank 0)
    else {
        const u64 total = u64(first_share_256) + 256ull * (world - 1);
        const u64 cum = u64(first_share_256) + 256ull * (g - 1);
        u64 idx = u64(ns) * cum / total;
        if (idx >= ns) idx = ns - 1;
        v = s_key[idx];
    }
    splitters[g] = v;
}

// Ownership count, fused with the plan's bounding box (points.cu: plan_bbox_kernel): per 256-point tile of the batch the number of
// points this rank owns; per scan the points it owns and the points owned by LOWER ranks (lower keys).
__global__ void __launch_bounds__(PT_THREADS) shard_count_kernel(const float* __restrict__ xyz, u32 n_points, const BatchScans* __restrict__ scans,
                                                                 u32 n_scans, float recip, BatchPlan* plan, const u64* __restrict__ splitters,
                                                                 u32 rank, u32 world, u32* __restrict__ tile_cnt, u32* __restrict__ scan_own,
                                                                 u32* __restrict__ scan_lower, u32* __restrict__ own_bits) {
    __shared__ __align__(16) float s_xyz[PT_THREADS * 3];
    __shared__ u32 s_r[PT_THREADS / 32], s_e[PT_THREADS / 32];
    __shared__ u32 s_cnt[4];  // own / lower of the tile's first scan, own / lower of the next one
    __shared__ u32 s_first;
    const u32 tile_base = blockIdx.x * PT_THREADS;
    __shared__ u64 s_split[SHARD_WORLD_MAX + 1];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_first = scan_of(scans, n_scans, tile_base);
    if (threadIdx.x >= 32 && threadIdx.x < 32 + SHARD_WORLD_MAX + 1) s_split[threadIdx.x - 32] = (threadIdx.x - 32 <= world) ? splitters[threadIdx.x - 32] : ~0ull;
    load_xyz_tile(xyz, tile_base, n_points, s_xyz);  // (ends with __syncthreads)
    const u32 i = tile_base + threadIdx.x, lane = threadIdx.x & 31;
    const u32 s_first_scan = s_first;
    u32 r = 0, err = 0;
    bool own = false, lower = false;
    u32 rel = 0;  // scan of the point relative to the tile's first scan
    const bool valid = i < n_points;
    if (valid) {
        const float px = s_xyz[threadIdx.x * 3], py = s_xyz[threadIdx.x * 3 + 1], pz = s_xyz[threadIdx.x * 3 + 2];
        i32 vx, vy, vz;
        if (!voxel_of(px, py, pz, recip, vx, vy, vz)) err = (isfinite(px) && isfinite(py) && isfinite(pz)) ? ERRF_RANGE : ERRF_NUMERIC;
        r = max(rcode(vx), max(rcode(vy), rcode(vz)));
        const u32 g = owner_of(morton_encode(vx, vy, vz) >> 9, s_split, world);
        own = g == rank;
        lower = g < rank;
        u32 s = s_first_scan;
        while (s + 1 < n_scans && scans->offset[s + 1] <= i) s++;  // a tile rarely straddles scans
        rel = s - s_first_scan;
    }
    r = __reduce_max_sync(0xffffffffu, r);
    err = __reduce_or_sync(0xffffffffu, err);
    if (lane == 0) { s_r[threadIdx.x >> 5] = r; s_e[threadIdx.x >> 5] = err; }
    // per-scan counts: the first two scans of the tile through shared memory, any further one (tiny scans) straight to global memory
#pragma unroll
    for (u32 q = 0; q < 2; q++) {
        const u32 bo = __ballot_sync(0xffffffffu, valid && rel == q && own), bl = __ballot_sync(0xffffffffu, valid && rel == q && lower);
        if (lane == 0) {
            if (bo) atomicAdd(&s_cnt[2 * q], (u32)__popc(bo));
            if (bl) atomicAdd(&s_cnt[2 * q + 1], (u32)__popc(bl));
        }
    }
    if (valid && rel >= 2) {
        if (own) atomicAdd(&scan_own[s_first_scan + rel], 1u);
        if (lower) atomicAdd(&scan_lower[s_first_scan + rel], 1u);
    }
    const u32 own_all = __ballot_sync(0xffffffffu, own);
    if (lane == 0) own_bits[i >> 5] = own_all;  // (i is a multiple of 32 here; warps beyond the batch write zero into the padding)
    __shared__ u32 s_own_total;
    if (threadIdx.x == 0) s_own_total = 0;
    __syncthreads();
    if (lane == 0 && own_all) atomicAdd(&s_own_total, (u32)__popc(own_all));
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 br = 0, be = 0;
#pragma unroll
        for (int w = 0; w < PT_THREADS / 32; w++) { br = max(br, s_r[w]); be |= s_e[w]; }
        if (br > *(volatile u32*)&plan->rmax) atomicMax(&plan->rmax, br);
        if (be) atomicOr(&plan->error, be);
        tile_cnt[blockIdx.x] = s_own_total;
        if (s_cnt[0]) atomicAdd(&scan_own[s_first_scan], s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&scan_lower[s_first_scan], s_cnt[1]);
        if (s_first_scan + 1 < n_scans) {
            if (s_cnt[2]) atomicAdd(&scan_own[s_first_scan + 1], s_cnt[2]);
            if (s_cnt[3]) atomicAdd(&scan_lower[s_first_scan + 1], s_cnt[3]);
        }
    }
}

// One block: the plan of the batch (k, key widths), the exclusive scan of the tile counts (-> where each tile's owned points go), and
// this rank's own scan table: offsets of its points per scan, the poses, the walk's tile table, and which scans' lowest point it holds.
// Clears the per-scan counters for the next batch.
__global__ void __launch_bounds__(SH_SCAN_THREADS) shard_plan_kernel(BatchPlan* plan, u32 margin, u32 tsb, u32 gbits, const BatchScans* __restrict__ scans,
                                                                     u32 n_scans, u32* __restrict__ tile_cnt, u32 n_tiles, u32* __restrict__ scan_own,
                                                                     u32* __restrict__ scan_lower, BatchScans* __restrict__ own_scans) {
    __shared__ u32 s_warp[SH_SCAN_THREADS / 32];
    u32 carry = 0;
    for (u32 base = 0; base < n_tiles; base += SH_SCAN_THREADS) {  // tile_cnt -> exclusive prefix, in place
        const u32 t = base + threadIdx.x;
        const u32 v = (t < n_tiles) ? tile_cnt[t] : 0u;
        u32 tot;
        const u32 ex = block_exclusive_scan<u32>(v, s_warp, tot);
        if (t < n_tiles) tile_cnt[t] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        u32 off = 0, tail_lo = 0, tail_hi = 0;
        for (u32 s = 0; s < n_scans; s++) {
            const u32 c = scan_own[s];
            own_scans->offset[s] = off;
            own_scans->pose[s][0] = scans->pose[s][0]; own_scans->pose[s][1] = scans->pose[s][1]; own_scans->pose[s][2] = scans->pose[s][2];
            if (c && scan_lower[s] == 0) { if (s < 32) tail_lo |= 1u << s; else tail_hi |= 1u << (s - 32); }  // no point of the scan lies below this range
            off += c;
            scan_own[s] = 0;
            scan_lower[s] = 0;
        }
        for (u32 s = n_scans; s <= MAX_BATCH_SCANS; s++) own_scans->offset[s] = off;
        batch_scans_tiles(*own_scans, n_scans);
        if (off != carry) atomicOr(&plan->error, ERRF_EXCHANGE);  // (cannot happen: both count the same points)
        plan->n_points = off;
        plan->tail_lo = tail_lo;
        plan->tail_hi = tail_hi;
        plan_finalize_body(plan, margin, tsb, gbits);
    }
}

// The owned points' sort keys, compacted in input order: key = (scan | ~compact Morton) with the BATCH index of the point below it
// (point_gather_kernel fetches the coordinates through it), or a separate index array when the key has no room. Only the owned
// points (the count kernel left one ownership ballot per warp) are read again: the pass costs 1 / world of the batch.
__global__ void __launch_bounds__(PT_THREADS) shard_scatter_kernel(const float* __restrict__ xyz, u32 n_points, const BatchScans* __restrict__ scans,
                                                                   float recip, const BatchPlan* __restrict__ plan, const u32* __restrict__ own_bits,
                                                                   const u32* __restrict__ tile_off, u64* __restrict__ sortkeys, u32* __restrict__ index) {
    __shared__ u32 s_wcnt[PT_THREADS / 32];
    const u32 i = blockIdx.x * PT_THREADS + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 b = own_bits[i >> 5];
    if (lane == 0) s_wcnt[warp] = (u32)__popc(b);
    __syncthreads();
    if (!((b >> lane) & 1u)) return;
    u32 before = 0;
#pragma unroll
    for (u32 w = 0; w < PT_THREADS / 32; w++) before += (w < warp) ? s_wcnt[w] : 0u;
    const u32 pos = tile_off[blockIdx.x] + before + (u32)__popc(b & ((1u << lane) - 1u));
    i32 vx, vy, vz;
    voxel_of(__ldg(&xyz[size_t(i) * 3]), __ldg(&xyz[size_t(i) * 3 + 1]), __ldg(&xyz[size_t(i) * 3 + 2]), recip, vx, vy, vz);
    const u64 sk = point_sort_key(morton_encode(vx, vy, vz), plan->k, scan_of(scans, plan->n_scans, i));
    if (plan->point_shift) sortkeys[pos] = (sk << POINT_INDEX_BITS) | (u64)i;
    else { sortkeys[pos] = sk; index[pos] = i; }
}

inline unsigned blocks_for(u32 n) { return (n + PT_THREADS - 1) / PT_THREADS; }

}  // namespace

size_t shard_filter_bytes(size_t max_points) {
    const size_t tiles = (max_points + PT_THREADS - 1) / PT_THREADS;
    return (tiles + tiles * (PT_THREADS / 32) + 2 * (MAX_BATCH_SCANS + 1) + 64) * sizeof(u32);
}

ShardFilter shard_filter_carve(void* mem, size_t max_points) {
    ShardFilter f;
    u32* p = static_cast<u32*>(mem);
    f.scan_own = p;
    f.scan_lower = p + (MAX_BATCH_SCANS + 1);
    f.tile_cnt = p + 2 * (MAX_BATCH_SCANS + 1) + 30;  // (keeps tile_cnt 16-byte aligned: 2 * 65 + 30 = 160 words)
    f.max_tiles = (u32)((max_points + PT_THREADS - 1) / PT_THREADS);
    f.own_bits = f.tile_cnt + f.max_tiles;  // one ballot per warp of the count kernel (whole tiles: PT_THREADS / 32 words each)
    return f;
}

// splitters of a new submap from its first scan (device pointer into the batch buffer)
int launch_shard_splitters(cudaStream_t s, const float* xyz_first_scan, u32 n_first_scan, const MapParams& mp, u32 world, u32 first_share_256,
                           u64* splitters) {
    shard_splitters_kernel<<<1, SH_SCAN_THREADS, 0, s>>>(xyz_first_scan, n_first_scan, mp.recip, world, first_share_256, splitters);
    return 1;
}

// plan + ownership filter of a batch: replaces launch_plan + launch_point_keys of the single-GPU point stage
int launch_shard_filter(cudaStream_t s, const float* xyz, u32 n_points, u32 n_scans, const BatchScans* scans, const MapParams& mp, BatchPlan* plan,
                        u32 tsb, u32 gbits, const u64* splitters, u32 rank, u32 world, const ShardFilter& f, BatchScans* own_scans, u64* sortkeys,
                        u32* index) {
    launch_plan_reset(s, plan, n_points, n_scans);
    int launches = 1;
    const u32 n_tiles = blocks_for(n_points);
    if (n_points) {
        shard_count_kernel<<<n_tiles, PT_THREADS, 0, s>>>(xyz, n_points, scans, n_scans, mp.recip, plan, splitters, rank, world, f.tile_cnt, f.scan_own,
                                                         f.scan_lower, f.own_bits);
        launches++;
    }
    shard_plan_kernel<<<1, SH_SCAN_THREADS, 0, s>>>(plan, mp.band_margin, tsb, gbits, scans, n_scans, f.tile_cnt, n_tiles, f.scan_own, f.scan_lower, own_scans);
    launches++;
    if (n_points) {
        shard_scatter_kernel<<<n_tiles, PT_THREADS, 0, s>>>(xyz, n_points, scans, mp.recip, plan, f.own_bits, f.tile_cnt, sortkeys, index);
        launches++;
    }
    return launches;
}

}  // namespace chadgpu
