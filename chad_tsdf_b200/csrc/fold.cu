// Ordered segmented fold of the sorted (voxel key, sd) stream into the resident leaf-chunk table.
//
// Replaces, per emitted voxel, Octree::insert(MortonCode) (hash probe + 21-level pointer walk,
// /root/reference/include/chad/detail/octree.hpp:31-78) and the running average of
// octree.hpp:161-163:   sd = (sd * float(w) + new) / float(++w)   -- three rounded fp32 operations
// per update, applied in the reference's order (sorted point, then ray step), which the stable
// radix sort has preserved inside every equal-key segment. One thread owns one segment, so the
// result is deterministic and bit-identical to the sequential CPU loop.
//
// The resident map is an open-addressing hash table of 2x2x2-voxel leaf chunks keyed by
// (Morton key >> 3): the octree's node layout and leaf addresses are not observable outside
// Submap::finalize's Morton-ordered walk, and a chunk is exactly the
// unit finalize packs into one LeafCluster (submap.hpp:74-100). weight == 0 marks an absent voxel
// (a touched voxel always has weight >= 1).
#include "kernels.cuh"
#include "radix_sort.cuh"

namespace chadgpu {

namespace {

constexpr int FOLD_THREADS = 256;

__device__ __forceinline__ u64 mix64(u64 h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

// find or insert `chunk`; returns the slot, or ~0 when the table is full
__device__ __forceinline__ u64 table_upsert(u64* __restrict__ keys, u64 capacity, u64 chunk, bool& inserted) {
    const u64 mask = capacity - 1;
    u64 h = mix64(chunk) & mask;
    inserted = false;
    for (u64 probes = 0; probes < capacity; probes++) {
        u64 cur = keys[h];
        if (cur == chunk) return h;
        if (cur == CHUNK_EMPTY) {
            const u64 old = atomicCAS(&keys[h], CHUNK_EMPTY, chunk);
            if (old == CHUNK_EMPTY) { inserted = true; return h; }
            if (old == chunk) return h;
        }
        h = (h + 1) & mask;
    }
    return ~0ull;
}

// Work distribution (ncu r01b: with one lane per segment the variance of segment lengths -- voxels near the sensor
// are hit by every scan of the batch -- left 6.8 of 32 lanes active). Every warp takes chunks of FOLD_CHUNK consecutive
// sorted updates and runs three phases:
//   0. stage keys + sd in shared memory (coalesced) and list the segment heads of the chunk;
//   1. lane-parallel over heads: probe / insert the leaf chunk, read the voxel's (sd, weight) seed;
//   2. the dependent (mul, add, div) chains with DYNAMIC lane scheduling: a lane that finishes its segment takes the
//      next unprocessed head of the chunk in the same iteration, so the lanes stay packed whatever the lengths are;
//   3. lane-parallel write-back of the folded cells.
// The chunk's last segment usually runs past the chunk end (ncu r01b: 1.7 M single-lane iterations, each behind a
// dependent global load, were the kernel's critical path): FOLD_LOOK further updates are staged so that it is folded
// in phase 2 like every other segment; only a segment longer than that is finished from global memory, by the whole
// warp with coalesced loads. Chunks are handed out dynamically (one ticket per warp) to even out the warps.
constexpr int FOLD_CHUNK = 256;            // updates per warp iteration
constexpr int FOLD_LOOK = 64;              // look-ahead past the chunk end
constexpr int FOLD_WARPS = FOLD_THREADS / 32;

__global__ void __launch_bounds__(FOLD_THREADS) fold_kernel(const u64* __restrict__ keys_a, const u64* __restrict__ keys_b,
                                                            const u32* __restrict__ sd_a, const u32* __restrict__ sd_b, BatchPlan* plan,
                                                            u64* __restrict__ tkeys, uint2* __restrict__ tcells, u64 capacity, u32* tcount,
                                                            u64* __restrict__ tlist) {
    __shared__ u64 s_keys[FOLD_WARPS][FOLD_CHUNK];   // phase 0/1: keys; afterwards slot h = (acc bits, weight) of head h
    __shared__ u64 s_cell[FOLD_WARPS][FOLD_CHUNK];   // cell index of head h in the table
    __shared__ u32 s_sd[FOLD_WARPS][FOLD_CHUNK + FOLD_LOOK];
    __shared__ unsigned short s_heads[FOLD_WARPS][FOLD_CHUNK + 1];
    const u32 n = plan->n_pairs;
    const u32 k = plan->k;
    const bool alt = radix_result_in_alt(plan->nbits_pairs);
    const u64* __restrict__ keys = alt ? keys_b : keys_a;
    const u32* __restrict__ sds = alt ? sd_b : sd_a;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lanemask_lt = (1u << lane) - 1u;
    const u32 num_chunks = (n + FOLD_CHUNK - 1) / FOLD_CHUNK;
    u32 new_chunks = 0, err = 0;
    while (true) {
        u32 chunk = 0;
        if (lane == 0) chunk = atomicAdd(&plan->fold_ticket, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= num_chunks) break;
        const u32 base = chunk * FOLD_CHUNK;
        const u32 cn = min((u32)FOLD_CHUNK, n - base);
        // ---- phase 0: stage the chunk and list its segment heads ----
        u64 prev_last = (base > 0) ? keys[base - 1] : 0ull;  // key just before the chunk (only used when base > 0)
        u32 n_heads = 0;
#pragma unroll
        for (int r = 0; r < FOLD_CHUNK / 32; r++) {
            const u32 e = r * 32 + lane;
            const bool valid = e < cn;
            const u64 key = valid ? keys[base + e] : 0ull;
            if (valid) { s_keys[warp][e] = key; s_sd[warp][e] = sds[base + e]; }
            u64 before = __shfl_up_sync(0xffffffffu, key, 1);
            if (lane == 0) before = prev_last;
            const bool head = valid && ((base + e == 0) || before != key);
            const u32 m = __ballot_sync(0xffffffffu, head);
            if (head) s_heads[warp][n_heads + __popc(m & lanemask_lt)] = (unsigned short)e;
            n_heads += __popc(m);
            prev_last = __shfl_sync(0xffffffffu, key, 31);
        }
        // look-ahead: how far does the chunk's last segment (key prev_last) extend past the chunk?
        u32 ext = 0;
        bool ext_open = (cn == FOLD_CHUNK);  // more updates may follow
#pragma unroll
        for (int r = 0; r < FOLD_LOOK / 32; r++) {
            const u32 idx = base + FOLD_CHUNK + r * 32 + lane;
            const bool valid = ext_open && idx < n;
            const u64 key = valid ? keys[idx] : ~prev_last;
            if (valid) s_sd[warp][FOLD_CHUNK + r * 32 + lane] = sds[idx];
            const u32 same = __ballot_sync(0xffffffffu, valid && key == prev_last);
            const u32 run = (same == 0xffffffffu) ? 32u : (u32)(__ffs(~same) - 1);
            if (ext_open) ext += run;
            if (run < 32) ext_open = false;
        }
        // ext_open still true: the segment is longer than the look-ahead; the rest is folded after phase 2
        if (lane == 0) s_heads[warp][n_heads] = (unsigned short)(cn + ext);  // sentinel: end of the chunk's last segment in shared memory
        __syncwarp();
        // ---- phase 1: seeds (rounds of 32 heads; round r only overwrites s_keys[32r .. 32r+31], whose keys belong to
        //      heads of rounds <= r because a head's position is >= its index) ----
        u64 last_key = 0ull;  // key of the chunk's last head (its segment may continue past the chunk)
        for (u32 h0 = 0; h0 < n_heads; h0 += 32) {
            const u32 h = h0 + lane;
            const bool has = h < n_heads;
            const u64 ckey = has ? s_keys[warp][s_heads[warp][h]] : 0ull;
            __syncwarp();
            bool inserted = false;
            u64 full = 0ull;
            if (has) {
                full = expand_key(ckey, k);
                const u64 slot = table_upsert(tkeys, capacity, full >> 3, inserted);
                if (slot == ~0ull) {
                    err |= ERRF_TABLE_FULL;
                    s_cell[warp][h] = ~0ull;
                    s_keys[warp][h] = 0ull;
                } else {
                    new_chunks += inserted ? 1u : 0u;
                    const u64 ci = slot * 8 + (full & 7ull);
                    const uint2 c = tcells[ci];  // (sd bits, weight); zero for a voxel touched for the first time (octree.hpp:68-75)
                    s_cell[warp][h] = ci;
                    s_keys[warp][h] = (u64(c.y) << 32) | c.x;
                }
                if (h + 1 == n_heads) last_key = ckey;
            }
            {   // new chunks join the table's chunk list (one reservation per warp)
                const u32 nb = __ballot_sync(0xffffffffu, inserted);
                if (nb) {
                    u32 lbase = 0;
                    if (lane == 0) lbase = atomicAdd(tcount, (u32)__popc(nb));
                    lbase = __shfl_sync(0xffffffffu, lbase, 0);
                    if (inserted) tlist[lbase + __popc(nb & lanemask_lt)] = full >> 3;
                }
            }
            __syncwarp();
        }
        last_key = __shfl_sync(0xffffffffu, last_key, (n_heads - 1) & 31);
        // every head but the last ends where the next one starts; fix the ends: head h ends at min(next head, cn), the
        // last one at the sentinel (cn + ext)
        // ---- phase 2: dependent chains, dynamically packed ----
        u32 next = 0;            // next unassigned head (warp-uniform)
        u32 my_h = 0xFFFFFFFFu;  // head this lane is folding
        u32 e = 0, e_end = 0, w = 0;
        float acc = 0.0f;
        while (true) {
            const bool need = my_h == 0xFFFFFFFFu;
            const u32 m = __ballot_sync(0xffffffffu, need);
            if (need) {
                const u32 cand = next + __popc(m & lanemask_lt);
                if (cand < n_heads) {
                    my_h = cand;
                    e = s_heads[warp][cand];
                    e_end = s_heads[warp][cand + 1];
                    const u64 seed = s_keys[warp][cand];
                    acc = __uint_as_float((u32)seed);
                    w = (u32)(seed >> 32);
                }
            }
            next += __popc(m);
            if (__ballot_sync(0xffffffffu, my_h != 0xFFFFFFFFu) == 0) break;
            if (my_h != 0xFFFFFFFFu) {
                acc = fadd(fmul(acc, __uint2float_rn(w)), __uint_as_float(s_sd[warp][e]));  // octree.hpp:161
                w++;                                                                          // octree.hpp:162
                acc = fdiv(acc, __uint2float_rn(w));                                          // octree.hpp:163
                e++;
                if (e == e_end) {
                    s_keys[warp][my_h] = (u64(w) << 32) | __float_as_uint(acc);
                    my_h = 0xFFFFFFFFu;
                }
            }
        }
        __syncwarp();
        if (ext_open && n_heads > 0) {
            // rare: the last segment is longer than the look-ahead. The whole warp walks it with coalesced loads; every lane
            // runs the same chain (uniform), lane 0 stores it.
            const u64 seed = s_keys[warp][n_heads - 1];
            float tacc = __uint_as_float((u32)seed);
            u32 tw = (u32)(seed >> 32);
            for (u32 j0 = base + FOLD_CHUNK + FOLD_LOOK; j0 < n; j0 += 32) {
                const u32 idx = j0 + lane;
                const bool valid = idx < n;
                const u64 key = valid ? keys[idx] : ~last_key;
                const u32 sdv = valid ? sds[idx] : 0u;
                const u32 same = __ballot_sync(0xffffffffu, valid && key == last_key);
                const u32 run = (same == 0xffffffffu) ? 32u : (u32)(__ffs(~same) - 1);
                for (u32 t = 0; t < run; t++) {
                    tacc = fadd(fmul(tacc, __uint2float_rn(tw)), __uint_as_float(__shfl_sync(0xffffffffu, sdv, t)));
                    tw++;
                    tacc = fdiv(tacc, __uint2float_rn(tw));
                }
                if (run < 32) break;
            }
            if (lane == 0) s_keys[warp][n_heads - 1] = (u64(tw) << 32) | __float_as_uint(tacc);
            __syncwarp();
        }
        // ---- phase 3: write the folded cells back ----
        for (u32 h = lane; h < n_heads; h += 32) {
            const u64 ci = s_cell[warp][h];
            if (ci != ~0ull) {
                const u64 v = s_keys[warp][h];
                tcells[ci] = make_uint2((u32)v, (u32)(v >> 32));
            }
        }
        __syncwarp();
    }
    // block-level reduction of the counters: one atomic per block
    __shared__ u32 s_red[2];
    if (threadIdx.x < 2) s_red[threadIdx.x] = 0;
    __syncthreads();
    new_chunks = __reduce_add_sync(0xffffffffu, new_chunks);
    err = __reduce_or_sync(0xffffffffu, err);
    if ((threadIdx.x & 31) == 0) {
        if (new_chunks) atomicAdd(&s_red[0], new_chunks);
        if (err) atomicOr(&s_red[1], err);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_red[0]) atomicAdd(&plan->n_new_chunks, s_red[0]);
        if (s_red[1]) atomicOr(&plan->error, s_red[1]);
    }
}

// distinct voxels (segments) and distinct leaf chunks of the sorted batch: the chunk count bounds
// what the fold can insert, so the host can size the table exactly before launching the fold
__global__ void __launch_bounds__(FOLD_THREADS) segment_count_kernel(const u64* __restrict__ keys_a, const u64* __restrict__ keys_b, BatchPlan* plan) {
    const u32 n = plan->n_pairs;
    const u64* __restrict__ keys = radix_result_in_alt(plan->nbits_pairs) ? keys_b : keys_a;
    u32 segments = 0, chunks = 0;
    for (u32 i = blockIdx.x * FOLD_THREADS + threadIdx.x; i < n; i += gridDim.x * FOLD_THREADS) {
        const u64 key = keys[i];
        const u64 prev = (i > 0) ? keys[i - 1] : ~key;
        segments += (prev != key) ? 1u : 0u;
        chunks += ((prev >> 3) != (key >> 3)) ? 1u : 0u;
    }
    __shared__ u32 s_red[2];
    if (threadIdx.x < 2) s_red[threadIdx.x] = 0;
    __syncthreads();
    segments = __reduce_add_sync(0xffffffffu, segments);
    chunks = __reduce_add_sync(0xffffffffu, chunks);
    if ((threadIdx.x & 31) == 0) {
        if (segments) atomicAdd(&s_red[0], segments);
        if (chunks) atomicAdd(&s_red[1], chunks);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_red[0]) atomicAdd(&plan->n_segments, s_red[0]);
        if (s_red[1]) atomicAdd(&plan->n_chunk_heads, s_red[1]);
    }
}

__global__ void __launch_bounds__(FOLD_THREADS) table_rehash_kernel(const u64* __restrict__ from_keys, const uint4* __restrict__ from_cells,
                                                                    u64 from_capacity, u64* __restrict__ to_keys, uint4* __restrict__ to_cells,
                                                                    u64 to_capacity, u32* to_count, u64* __restrict__ to_list) {
    for (u64 s = u64(blockIdx.x) * FOLD_THREADS + threadIdx.x; s < from_capacity; s += u64(gridDim.x) * FOLD_THREADS) {
        const u64 chunk = from_keys[s];
        if (chunk == CHUNK_EMPTY) continue;
        bool inserted;
        const u64 slot = table_upsert(to_keys, to_capacity, chunk, inserted);
        if (slot == ~0ull) continue;  // cannot happen: the target is larger
#pragma unroll
        for (int q = 0; q < 4; q++) to_cells[slot * 4 + q] = from_cells[s * 4 + q];
        to_list[atomicAdd(to_count, 1u)] = chunk;
    }
}

// the table's chunk list -> (full chunk key, index) + the max range code over all resident voxels
__global__ void __launch_bounds__(FOLD_THREADS) table_compact_kernel(const u64* __restrict__ tlist, const u32* __restrict__ tcount, u32 max_n,
                                                                     u64* __restrict__ out_keys, u32* __restrict__ out_vals, u32* d_count, u32* d_rmax) {
    const u32 n = min(*tcount, max_n);
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_count = n;
    u32 rmax = 0;
    for (u32 i = blockIdx.x * FOLD_THREADS + threadIdx.x; i < n; i += gridDim.x * FOLD_THREADS) {
        const u64 chunk = tlist[i];
        out_keys[i] = chunk;
        out_vals[i] = i;
        i32 x, y, z;
        morton_decode(chunk << 3, x, y, z);
        rmax = max(rmax, max(rcode(x), max(rcode(y), rcode(z))) | 1u);  // the chunk spans v and v+1 on every axis
    }
    rmax = __reduce_max_sync(0xffffffffu, rmax);
    if ((threadIdx.x & 31) == 0 && rmax) atomicMax(d_rmax, rmax);
}

__device__ __forceinline__ u32 chunk_k(u32 rmax) {
    u32 k = 32 - __clz(rmax);
    if (k < 3) k = 3;
    if (k > 20) k = 20;
    return k;
}

// chunk sort keys: compact(full voxel key) >> 3 -> 3k bits; writes k-derived nbits for the radix sort
__global__ void __launch_bounds__(FOLD_THREADS) chunk_sortkeys_kernel(u64* __restrict__ keys, const u32* __restrict__ d_count,
                                                                      const u32* __restrict__ d_rmax, u32* __restrict__ d_nbits) {
    const u32 n = *d_count;
    const u32 k = chunk_k(*d_rmax);
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_nbits = 3 * k;
    for (u32 i = blockIdx.x * FOLD_THREADS + threadIdx.x; i < n; i += gridDim.x * FOLD_THREADS) keys[i] = compact_key(keys[i] << 3, k) >> 3;
}

__device__ __forceinline__ u64 table_find(const u64* __restrict__ keys, u64 capacity, u64 chunk) {
    const u64 mask = capacity - 1;
    u64 h = mix64(chunk) & mask;
    for (u64 probes = 0; probes < capacity; probes++) {
        const u64 cur = keys[h];
        if (cur == chunk) return h;
        if (cur == CHUNK_EMPTY) return ~0ull;
        h = (h + 1) & mask;
    }
    return ~0ull;
}

// sorted chunk keys -> contiguous (full chunk key, 8 cells) for export / cluster building. The count and the buffer the
// radix sort left its result in are read from device memory, so the host needs only an upper bound of the count.
__global__ void __launch_bounds__(FOLD_THREADS) chunk_gather_kernel(const u64* __restrict__ tkeys, const uint4* __restrict__ tcells, u64 capacity,
                                                                    const u64* __restrict__ keys_a, const u64* __restrict__ keys_b,
                                                                    const u32* __restrict__ d_count, const u32* __restrict__ d_rmax,
                                                                    const u32* __restrict__ d_nbits, u64* __restrict__ out_keys,
                                                                    uint4* __restrict__ out_cells) {
    const u32 i = blockIdx.x * FOLD_THREADS + threadIdx.x;
    if (i >= *d_count) return;
    const u32 k = chunk_k(*d_rmax);
    const u64 ckey = radix_result_in_alt(*d_nbits) ? keys_b[i] : keys_a[i];
    const u64 chunk = expand_key(ckey << 3, k) >> 3;
    const u64 s = table_find(tkeys, capacity, chunk);
    out_keys[i] = chunk;
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int q = 0; q < 4; q++) out_cells[size_t(i) * 4 + q] = (s != ~0ull) ? tcells[size_t(s) * 4 + q] : z;
}

}  // namespace

int launch_table_clear(cudaStream_t s, const ChunkTable& t) {
    cudaMemsetAsync(t.keys, 0xFF, t.capacity * sizeof(u64), s);
    cudaMemsetAsync(t.cells, 0, t.capacity * 8 * sizeof(uint2), s);
    cudaMemsetAsync(t.count, 0, 4, s);
    return 0;
}

int launch_fold(cudaStream_t s, const u64* keys_a, const u64* keys_b, const u32* sd_a, const u32* sd_b, u32 max_pairs, BatchPlan* plan,
                const ChunkTable& t, int num_sms) {
    if (!max_pairs) return 0;
    u32 want = (max_pairs + FOLD_CHUNK * FOLD_WARPS - 1) / (FOLD_CHUNK * FOLD_WARPS);
    u32 cap = (u32)num_sms * 4;  // persistent: chunks are handed out by ticket
    fold_kernel<<<want < cap ? want : cap, FOLD_THREADS, 0, s>>>(keys_a, keys_b, sd_a, sd_b, plan, t.keys, t.cells, t.capacity, t.count, t.list);
    return 1;
}

int launch_segment_count(cudaStream_t s, const u64* keys_a, const u64* keys_b, u32 max_pairs, BatchPlan* plan, int num_sms) {
    if (!max_pairs) return 0;
    u32 want = (max_pairs + FOLD_THREADS - 1) / FOLD_THREADS;
    u32 cap = (u32)num_sms * 16;
    segment_count_kernel<<<want < cap ? want : cap, FOLD_THREADS, 0, s>>>(keys_a, keys_b, plan);
    return 1;
}

int launch_table_rehash(cudaStream_t s, const ChunkTable& from, const ChunkTable& to, int num_sms) {
    table_rehash_kernel<<<num_sms * 8, FOLD_THREADS, 0, s>>>(from.keys, reinterpret_cast<const uint4*>(from.cells), from.capacity, to.keys,
                                                             reinterpret_cast<uint4*>(to.cells), to.capacity, to.count, to.list);
    return 1;
}

int launch_table_compact(cudaStream_t s, const ChunkTable& t, u32 max_n, u64* out_keys, u32* out_vals, u32* d_count, u32* d_rmax, u32* d_nbits, int num_sms) {
    cudaMemsetAsync(d_rmax, 0, 4, s);
    table_compact_kernel<<<num_sms * 4, FOLD_THREADS, 0, s>>>(t.list, t.count, max_n, out_keys, out_vals, d_count, d_rmax);
    chunk_sortkeys_kernel<<<num_sms * 4, FOLD_THREADS, 0, s>>>(out_keys, d_count, d_rmax, d_nbits);
    return 2;
}

int launch_chunk_gather(cudaStream_t s, const ChunkTable& t, const u64* keys_a, const u64* keys_b, const u32* d_count, const u32* d_rmax,
                        const u32* d_nbits, u32 max_n, u64* out_keys, void* out_cells) {
    if (!max_n) return 0;
    chunk_gather_kernel<<<(max_n + FOLD_THREADS - 1) / FOLD_THREADS, FOLD_THREADS, 0, s>>>(t.keys, reinterpret_cast<const uint4*>(t.cells), t.capacity, keys_a,
                                                                                          keys_b, d_count, d_rmax, d_nbits, out_keys,
                                                                                          static_cast<uint4*>(out_cells));
    return 1;
}

}  // namespace chadgpu
