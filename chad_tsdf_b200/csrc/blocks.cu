// Block-binned grouping of the band-voxel updates: the default replacement for the octree's per-voxel grouping
// (/root/reference/include/chad/detail/octree.hpp:31-78,153-164), instead of a global multi-pass radix sort.
//
// Every band voxel of a ray lies within (trunc/res + 1) voxels of the ray's point, so the updates of one 8x8x8-voxel
// block come from the few thousand rays that end in or next to it. The updates are therefore
//   1. counted per block (blocks_count_kernel: the DDA of octree.hpp:86-152, one hash find-or-insert + one atomic
//      per group of neighbouring rays that share a block),
//   2. given contiguous, exactly sized regions (prefix sum over the block table),
//   3. written as 8-byte records (local voxel << 23 | sorted-point rank, sd) into their block's region in arbitrary
//      order (blocks_emit_kernel, one atomic cursor bump per run),
//   4. sorted inside each block in shared memory by (voxel, rank) -- counting sort by voxel, then every record counts
//      the records of its voxel with a smaller rank -- and written out as the (compact key, sd) arrays the fold
//      consumes (blocks_sort_kernel).
// Per update this moves 8 B out + 8 B in + 12 B out instead of 8 + 24 x passes + 8; the reference's update order
// inside a voxel (sorted point rank, then ray step; a ray touches a voxel at most once) is restored exactly by the
// rank sort, so the result is bit-identical to the global-sort path and to the CPU reference.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "ray.cuh"
#include "scan.cuh"

namespace chadgpu {

namespace {

constexpr int BLK_THREADS = 256;
constexpr u32 BLK_SHIFT = 9;               // 8^3 voxels per block
constexpr u32 BLK_VOXELS = 512;
constexpr u32 BLK_RANK_BITS = 23;          // sorted-point rank inside the batch
constexpr u32 BLK_RMAX = 4096;             // records one CTA sorts in shared memory at once (32 KB -> 6 CTAs per SM)
constexpr int BLK_RAY_MAX = 32;            // voxels of one (ray, block) run buffered in the emit kernel
constexpr u64 BLK_EMPTY = ~0ull;

__device__ __forceinline__ u64 mix64(u64 h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}
__device__ __forceinline__ u32 block_insert(u64* __restrict__ keys, u32 capacity, u64 block) {
    const u32 mask = capacity - 1;
    u32 h = (u32)mix64(block) & mask;
    for (u32 probes = 0; probes < capacity; probes++) {
        const u64 cur = keys[h];
        if (cur == block) return h;
        if (cur == BLK_EMPTY) {
            const u64 old = atomicCAS(&keys[h], BLK_EMPTY, block);
            if (old == BLK_EMPTY || old == block) return h;
        }
        h = (h + 1) & mask;
    }
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ u32 block_find(const u64* __restrict__ keys, u32 capacity, u64 block) {
    const u32 mask = capacity - 1;
    u32 h = (u32)mix64(block) & mask;
    for (u32 probes = 0; probes < capacity; probes++) {
        const u64 cur = keys[h];
        if (cur == block) return h;
        if (cur == BLK_EMPTY) break;
        h = (h + 1) & mask;
    }
    return 0xFFFFFFFFu;
}

// ---- warp aggregation ---------------------------------------------------------------------------
// Neighbouring rays (adjacent lanes: the points are Morton sorted) mostly end in the same block. One lane per group
// of equal block ids issues the global atomic for the whole group (ncu/bench r01: one atomic per (ray, block) run made
// the count kernel 10x slower than the DDA it wraps). Returns the group's total in `total` and this lane's exclusive
// prefix (lanes ordered by index) in `prefix`; `leader` = lowest lane of the group. All 32 lanes must call.
__device__ __forceinline__ void group_sum(u64 key, u32 value, u32 lane, u32& total, u32& prefix, u32& leader) {
    const u32 m = __match_any_sync(0xffffffffu, key);
    leader = (u32)(__ffs(m) - 1);
    const u32 lt = m & ((1u << lane) - 1u);
    u32 t = 0, p = 0;
    // value <= BLK_RAY_MAX < 64: sum bit plane by bit plane with ballots (no loop over the group's lanes)
#pragma unroll
    for (int b = 0; b < 6; b++) {
        const u32 plane = __ballot_sync(0xffffffffu, (value >> b) & 1u);
        t += (u32)__popc(plane & m) << b;
        p += (u32)__popc(plane & lt) << b;
    }
    total = t;
    prefix = p;
}

constexpr int BLK_MAX_RUNS = 4;  // (ray, block) runs kept per lane for aggregation; further runs use their own atomics

// ---- 1. count -------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK_THREADS) blocks_count_kernel(const float* __restrict__ xyz_sorted, u32 n_points,
                                                                   const BatchScans* __restrict__ scans, float res, float trunc, float recip,
                                                                   u32 max_ray_voxels, BatchPlan* plan, u64* __restrict__ bkeys,
                                                                   u32* __restrict__ bcount, u32 capacity) {
    const u32 i = blockIdx.x * BLK_THREADS + threadIdx.x;
    const u32 lane = threadIdx.x & 31;
    const bool active = i < n_points;
    u64 rb[BLK_MAX_RUNS];
    u32 rl[BLK_MAX_RUNS];
    u32 nr = 0, err = 0;
    if (active) {
        const u32 s = scan_of(scans, plan->n_scans, i);
        const float pos[3] = {scans->pose[s][0], scans->pose[s][1], scans->pose[s][2]};
        Ray r;
        ray_setup(r, xyz_sorted[size_t(i) * 3], xyz_sorted[size_t(i) * 3 + 1], xyz_sorted[size_t(i) * 3 + 2], pos, res, trunc, recip);
        u64 full = morton_encode(r.cur[0], r.cur[1], r.cur[2]);
        u64 blk = full >> BLK_SHIFT;
        u32 run = 0, total = 0;
        while (true) {
            run++;
            total++;
            int axis = 0;
            bool more = ray_advance(r, axis);
            if (more && total >= max_ray_voxels) { err |= ERRF_PAIR_CAPACITY; more = false; }  // analytic per-ray bound exceeded: report it
            u64 next_full = full;
            if (more) next_full = morton_step(full, axis, axis == 0 ? r.step[0] : (axis == 1 ? r.step[1] : r.step[2]));
            if (!more || (next_full >> BLK_SHIFT) != blk) {  // the run ends
                if (nr < BLK_MAX_RUNS) {
#pragma unroll
                    for (int q = 0; q < BLK_MAX_RUNS; q++)
                        if ((u32)q == nr) { rb[q] = blk; rl[q] = run; }
                    nr++;
                } else {
                    const u32 slot = block_insert(bkeys, capacity, blk);
                    if (slot == 0xFFFFFFFFu) err |= ERRF_BLOCKS_FULL; else atomicAdd(&bcount[slot], run);
                }
                run = 0;
                blk = next_full >> BLK_SHIFT;
            }
            if (!more) break;
            full = next_full;
        }
    }
#pragma unroll
    for (int q = 0; q < BLK_MAX_RUNS; q++) {
        const bool has = (u32)q < nr;
        const u64 key = has ? rb[q] : (0xFFFFFFFF00000000ull | lane);  // unique per lane: never a real block id (54 bits)
        u32 total, prefix, leader;
        group_sum(key, has ? rl[q] : 0u, lane, total, prefix, leader);
        if (has && lane == leader) {
            const u32 slot = block_insert(bkeys, capacity, key);
            if (slot == 0xFFFFFFFFu) err |= ERRF_BLOCKS_FULL; else atomicAdd(&bcount[slot], total);
        }
    }
    if (err) atomicOr(&plan->error, err);
}

// ---- 2. list the non-empty blocks -----------------------------------------------------------------
__global__ void __launch_bounds__(BLK_THREADS) blocks_compact_kernel(const u32* __restrict__ bcount, u32 capacity, u32* __restrict__ list, BatchPlan* plan) {
    const u32 s = blockIdx.x * BLK_THREADS + threadIdx.x;
    const bool occ = (s < capacity) && bcount[s] != 0;
    const u32 ballot = __ballot_sync(0xffffffffu, occ);
    const u32 lane = threadIdx.x & 31;
    u32 base = 0;
    if (ballot && lane == 0) base = atomicAdd(&plan->n_blocks, (u32)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (occ) list[base + __popc(ballot & ((1u << lane) - 1u))] = s;
}

// ---- 3. emit ----------------------------------------------------------------------------------
// record = (sd bits << 32) | (local voxel << 23) | rank. `records` is the pair-key buffer the sort will NOT leave its
// result in (selected on the device like everywhere else). The ray is walked once; its records wait in a per-lane
// buffer until the warp has reserved their places (one atomic per group of lanes that share a block).
__global__ void __launch_bounds__(BLK_THREADS) blocks_emit_kernel(const float* __restrict__ xyz_sorted, const float* __restrict__ normals,
                                                                  u32 n_points, const BatchScans* __restrict__ scans, float res, float trunc,
                                                                  float recip, u32 max_ray_voxels, BatchPlan* plan,
                                                                  const u64* __restrict__ bkeys, const u32* __restrict__ boffset,
                                                                  u32* __restrict__ bcursor, u32 capacity, u64* keys_a, u64* keys_b,
                                                                  u32 pair_capacity) {
    const u32 i = blockIdx.x * BLK_THREADS + threadIdx.x;
    const u32 lane = threadIdx.x & 31;
    const bool active = (i < n_points) && !(plan->error & ERRF_BLOCKS_FULL);
    u64* __restrict__ records = radix_result_in_alt(plan->nbits_pairs) ? keys_a : keys_b;
    u64 buf[BLK_RAY_MAX];
    u64 rb[BLK_MAX_RUNS];
    u32 rl[BLK_MAX_RUNS];
    u32 nr = 0, err = 0, total = 0;
    if (active) {
        const u32 s = scan_of(scans, plan->n_scans, i);
        const float pos[3] = {scans->pose[s][0], scans->pose[s][1], scans->pose[s][2]};
        const float nx = normals[size_t(i) * 3], ny = normals[size_t(i) * 3 + 1], nz = normals[size_t(i) * 3 + 2];
        Ray r;
        ray_setup(r, xyz_sorted[size_t(i) * 3], xyz_sorted[size_t(i) * 3 + 1], xyz_sorted[size_t(i) * 3 + 2], pos, res, trunc, recip);
        u64 full = morton_encode(r.cur[0], r.cur[1], r.cur[2]);
        u64 blk = full >> BLK_SHIFT;
        u32 run = 0;
        const u32 cap = min(max_ray_voxels, (u32)BLK_RAY_MAX);
        while (true) {
            // octree.hpp:157-159: the voxel's LOWER CORNER projected on the normal, clamped to +-trunc (inside the Morton range
            // the stored code decodes back to the same voxel, so the reference's decode is the identity)
            float sd = dot3(nx, ny, nz, fsub(fmul((float)r.cur[0], res), r.px), fsub(fmul((float)r.cur[1], res), r.py),
                            fsub(fmul((float)r.cur[2], res), r.pz));
            sd = fclamp(sd, -trunc, trunc);
            if (max(rcode(r.cur[0]), max(rcode(r.cur[1]), rcode(r.cur[2]))) >= (1u << 20)) err |= ERRF_RANGE;
            const u32 v = (u32)full & (BLK_VOXELS - 1);
            buf[total] = (u64(__float_as_uint(sd)) << 32) | (u64)((v << BLK_RANK_BITS) | i);
            run++;
            total++;
            bool more = total < cap;
            int axis = 0;
            if (more) more = ray_advance(r, axis);
            u64 next_full = full;
            if (more) next_full = morton_step(full, axis, axis == 0 ? r.step[0] : (axis == 1 ? r.step[1] : r.step[2]));
            if (!more || (next_full >> BLK_SHIFT) != blk) {
                if (nr < BLK_MAX_RUNS) {
#pragma unroll
                    for (int q = 0; q < BLK_MAX_RUNS; q++)
                        if ((u32)q == nr) { rb[q] = blk; rl[q] = run; }
                    nr++;
                } else {  // rare: more block crossings than run slots -- reserve and write this run directly
                    const u32 slot = block_find(bkeys, capacity, blk);
                    if (slot == 0xFFFFFFFFu) err |= ERRF_BLOCKS_FULL;
                    else {
                        const u32 base = boffset[slot] + atomicAdd(&bcursor[slot], run);
                        if (base + run > pair_capacity) err |= ERRF_PAIR_CAPACITY;
                        else for (u32 q = 0; q < run; q++) records[base + q] = buf[total - run + q];
                    }
                }
                run = 0;
                blk = next_full >> BLK_SHIFT;
            }
            if (!more) break;
            full = next_full;
        }
    }
    u32 first = 0;  // index in buf of the current run's first record
#pragma unroll
    for (int q = 0; q < BLK_MAX_RUNS; q++) {
        const bool has = (u32)q < nr;
        const u64 key = has ? rb[q] : (0xFFFFFFFF00000000ull | lane);
        const u32 len = has ? rl[q] : 0u;
        u32 gtotal, prefix, leader;
        group_sum(key, len, lane, gtotal, prefix, leader);
        u32 gbase = 0;
        if (has && lane == leader) {
            const u32 slot = block_find(bkeys, capacity, key);
            if (slot == 0xFFFFFFFFu) { err |= ERRF_BLOCKS_FULL; gbase = 0xFFFFFFFFu; }
            else gbase = boffset[slot] + atomicAdd(&bcursor[slot], gtotal);
        }
        gbase = __shfl_sync(0xffffffffu, gbase, leader);
        if (has && gbase != 0xFFFFFFFFu) {
            const u32 base = gbase + prefix;
            if (base + len > pair_capacity) err |= ERRF_PAIR_CAPACITY;
            else for (u32 j = 0; j < len; j++) records[base + j] = buf[first + j];
        }
        first += len;
    }
    if (err) atomicOr(&plan->error, err);
}

// ---- 4. per-block sort --------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK_THREADS) blocks_sort_kernel(const u64* __restrict__ bkeys, const u32* __restrict__ bcount,
                                                                  const u32* __restrict__ boffset, const u32* __restrict__ list, BatchPlan* plan,
                                                                  u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    u64* s_rec = reinterpret_cast<u64*>(s_dyn);  // [BLK_RMAX]
    __shared__ u32 s_hist[BLK_VOXELS];
    __shared__ u32 s_start[BLK_VOXELS + 1];
    __shared__ u32 s_cursor[BLK_VOXELS];
    __shared__ u32 s_bounds[BLK_VOXELS + 1];  // voxel boundaries of the shared-memory passes
    __shared__ u32 s_npass;
    __shared__ u32 s_ticket;
    __shared__ u32 s_warp_tot[BLK_THREADS / 32];
    __shared__ u32 s_stats[2];
    const u32 tid = threadIdx.x;
    if (tid == 0) s_npass = plan->error & (ERRF_BLOCKS_FULL | ERRF_PAIR_CAPACITY);
    if (tid < 2) s_stats[tid] = 0;
    __syncthreads();
    if (s_npass) return;  // uniform: read once per CTA
    __syncthreads();
    const bool alt = radix_result_in_alt(plan->nbits_pairs);
    const u64* __restrict__ records = alt ? keys_a : keys_b;
    u64* __restrict__ keys_out = alt ? keys_b : keys_a;
    u32* __restrict__ sd_out = alt ? vals_b : vals_a;
    const u32 k = plan->k;
    const u32 n_blocks = plan->n_blocks;
    while (true) {
        if (tid == 0) s_ticket = atomicAdd(&plan->sort_ticket, 1u);
        __syncthreads();
        const u32 t = s_ticket;
        if (t >= n_blocks) break;
        const u32 slot = list[t];
        const u32 R = bcount[slot];
        const u32 off = boffset[slot];
        const u64 blk = bkeys[slot];
        // ---- histogram over the local voxel ----
        s_hist[tid] = 0;
        s_hist[tid + BLK_THREADS] = 0;
        __syncthreads();
        for (u32 j = tid; j < R; j += BLK_THREADS) atomicAdd(&s_hist[(u32)records[off + j] >> BLK_RANK_BITS], 1u);
        __syncthreads();
        // exclusive scan of the 512 bins (two per thread); distinct voxels / leaf chunks of the batch
        {
            const u32 a = s_hist[2 * tid], b = s_hist[2 * tid + 1];
            u32 tot;
            const u32 ex = block_exclusive_scan<u32>(a + b, s_warp_tot, tot);
            s_start[2 * tid] = ex;
            s_start[2 * tid + 1] = ex + a;
            if (tid == 0) s_start[BLK_VOXELS] = R;
            const u32 segs = (a ? 1u : 0u) + (b ? 1u : 0u);
            // a leaf chunk = 8 consecutive local voxels = 4 consecutive threads
            const u32 any4 = __ballot_sync(0xffffffffu, (a | b) != 0);
            u32 chunks = 0;
            if ((tid & 3) == 0) chunks = ((any4 >> (tid & 31)) & 0xFu) ? 1u : 0u;
            const u32 wsegs = __reduce_add_sync(0xffffffffu, segs), wchunks = __reduce_add_sync(0xffffffffu, chunks);
            if ((tid & 31) == 0) { atomicAdd(&s_stats[0], wsegs); atomicAdd(&s_stats[1], wchunks); }
        }
        __syncthreads();
        // ---- split the voxel range into passes of at most BLK_RMAX records (one pass unless the block is huge) ----
        if (R <= BLK_RMAX) {
            if (tid == 0) { s_bounds[0] = 0; s_bounds[1] = BLK_VOXELS; s_npass = 1; }
        } else if (tid == 0) {
            u32 np = 0, lo = 0;
            s_bounds[0] = 0;
            for (u32 v = 0; v < BLK_VOXELS; v++)
                if (s_start[v + 1] - s_start[lo] > BLK_RMAX && v > lo) { s_bounds[++np] = v; lo = v; }
            s_bounds[++np] = BLK_VOXELS;
            s_npass = np;
        }
        __syncthreads();
        const u32 npass = s_npass;
        for (u32 q = 0; q < npass; q++) {
            const u32 vlo = s_bounds[q], vhi = s_bounds[q + 1];
            const u32 base_q = s_start[vlo];
            const u32 n_q = s_start[vhi] - base_q;
            if (n_q > BLK_RMAX) {
                // degenerate: ONE voxel with more updates than fit in shared memory. Rank every record of the voxel by
                // counting the records of the same voxel with a smaller sort key (quadratic, global memory; correctness path)
                for (u32 j = tid; j < R; j += BLK_THREADS) {
                    const u64 rec = records[off + j];
                    const u32 v = (u32)rec >> BLK_RANK_BITS;
                    if (v != vlo) continue;
                    u32 pos = 0;
                    for (u32 c = 0; c < R; c++) {
                        const u32 key = (u32)records[off + c];
                        pos += ((key >> BLK_RANK_BITS) == v && key < (u32)rec) ? 1u : 0u;
                    }
                    keys_out[off + base_q + pos] = compact_key((blk << BLK_SHIFT) | (u64)v, k);
                    sd_out[off + base_q + pos] = (u32)(rec >> 32);
                }
                __syncthreads();
                continue;
            }
            for (u32 v = vlo + tid; v < vhi; v += BLK_THREADS) s_cursor[v] = s_start[v] - base_q;
            __syncthreads();
            // ---- counting sort by voxel into shared memory (order inside a voxel: arbitrary) ----
            for (u32 j = tid; j < R; j += BLK_THREADS) {
                const u64 rec = records[off + j];
                const u32 v = (u32)rec >> BLK_RANK_BITS;
                if (v >= vlo && v < vhi) s_rec[atomicAdd(&s_cursor[v], 1u)] = rec;
            }
            __syncthreads();
            // ---- position of every record inside its voxel = number of records of that voxel with a smaller rank. One
            //      thread per RECORD (neighbouring threads work on the same voxel: balanced, broadcast shared-memory reads),
            //      written straight to the output arrays (ncu r01c: one thread per voxel with an insertion sort left 4 of 32
            //      lanes active and 30 % of the samples waiting at the barrier behind the longest voxel) ----
            for (u32 j = tid; j < n_q; j += BLK_THREADS) {
                const u64 rec = s_rec[j];
                const u32 key = (u32)rec;
                const u32 v = key >> BLK_RANK_BITS;
                const u32 b0 = s_start[v] - base_q, b1 = s_start[v + 1] - base_q;
                u32 pos = b0;
                for (u32 c = b0; c < b1; c++) pos += ((u32)s_rec[c] < key) ? 1u : 0u;
                keys_out[off + base_q + pos] = compact_key((blk << BLK_SHIFT) | (u64)v, k);
                sd_out[off + base_q + pos] = (u32)(rec >> 32);
            }
            __syncthreads();
        }
    }
    if (tid == 0 && s_stats[0]) { atomicAdd(&plan->n_segments, s_stats[0]); atomicAdd(&plan->n_chunk_heads, s_stats[1]); }
}

inline unsigned blocks_for(u32 n) { return (n + BLK_THREADS - 1) / BLK_THREADS; }

}  // namespace

size_t blocks_table_bytes(u32 capacity) { return size_t(capacity) * (8 + 4 + 4 + 4 + 4); }

BlockTable blocks_table_carve(void* mem, u32 capacity) {
    BlockTable t;
    unsigned char* p = static_cast<unsigned char*>(mem);
    t.keys = reinterpret_cast<u64*>(p); p += size_t(capacity) * 8;
    t.count = reinterpret_cast<u32*>(p); p += size_t(capacity) * 4;
    t.cursor = reinterpret_cast<u32*>(p); p += size_t(capacity) * 4;
    t.offset = reinterpret_cast<u32*>(p); p += size_t(capacity) * 4;
    t.list = reinterpret_cast<u32*>(p);
    t.capacity = capacity;
    return t;
}

cudaError_t blocks_init() {
    return cudaFuncSetAttribute(blocks_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BLK_RMAX * 8));
}

u32 blocks_max_batch_points() { return 1u << BLK_RANK_BITS; }

// Everything between the point stage and the fold. Returns the number of kernels queued.
int launch_blocks_pairs(cudaStream_t s, const float* xyz_sorted, const float* normals, u32 n_points, const BatchScans* scans, const MapParams& mp,
                        BatchPlan* plan, const BlockTable& bt, void* scan_ws, u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, u32 pair_capacity,
                        int num_sms, const LaunchHook* hook, int cls_count, int cls_scan, int cls_emit, int cls_sort) {
    if (!n_points) return 0;
    int launches = 0;
    // keys = EMPTY; count, cursor = 0 (contiguous: keys | count | cursor)
    cudaMemsetAsync(bt.keys, 0xFF, size_t(bt.capacity) * 8, s);
    cudaMemsetAsync(bt.count, 0, size_t(bt.capacity) * (4 + 4), s);
    if (hook) hook->begin(hook->user, cls_count);
    blocks_count_kernel<<<blocks_for(n_points), BLK_THREADS, 0, s>>>(xyz_sorted, n_points, scans, mp.res, mp.trunc, mp.recip, mp.max_ray_voxels, plan,
                                                                     bt.keys, bt.count, bt.capacity);
    if (hook) hook->end(hook->user);
    launches++;
    if (hook) hook->begin(hook->user, cls_scan);
    launches += exclusive_scan<u32, u32>(s, bt.count, bt.offset, bt.capacity, scan_ws, (u32*)nullptr, &plan->n_pairs);
    blocks_compact_kernel<<<blocks_for(bt.capacity), BLK_THREADS, 0, s>>>(bt.count, bt.capacity, bt.list, plan);
    if (hook) hook->end(hook->user);
    launches++;
    if (hook) hook->begin(hook->user, cls_emit);
    blocks_emit_kernel<<<blocks_for(n_points), BLK_THREADS, 0, s>>>(xyz_sorted, normals, n_points, scans, mp.res, mp.trunc, mp.recip, mp.max_ray_voxels,
                                                                    plan, bt.keys, bt.offset, bt.cursor, bt.capacity, keys_a, keys_b, pair_capacity);
    if (hook) hook->end(hook->user);
    launches++;
    if (hook) hook->begin(hook->user, cls_sort);
    blocks_sort_kernel<<<num_sms * 6, BLK_THREADS, BLK_RMAX * 8, s>>>(bt.keys, bt.count, bt.offset, bt.list, plan, keys_a, keys_b, vals_a, vals_b);
    if (hook) hook->end(hook->user);
    launches++;
    return launches;
}


}  // namespace chadgpu
