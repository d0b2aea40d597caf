// Submap::finalize on the device (/root/reference/include/chad/detail/submap.hpp:10-106): the
// resident leaf chunks of the active submap are folded bottom-up into the global, hash-deduplicated
// DAG (NodeLevels, /root/reference/include/chad/detail/levels.hpp:146-200).
//
// The reference runs one explicit-stack post-order DFS and calls LeafClusterLevel::add
// (levels.hpp:123-139) / NodeLevel::add (levels.hpp:57-88) once per record; an address is
// "number of distinct records added before, over the map's lifetime" (clusters) or "_occupied_n
// before the add" (nodes). Because the DFS visits children 0..7, every level's add sequence is:
// for each node of that level in ascending Morton order, TSDF record then weight record, and the
// sequences of different levels are independent. That makes each level a
// data-parallel first-occurrence dedup:
// (clusters: one small kernel per step below; node levels: the same five steps inside dag_levels_kernel)
//   probe   : find-or-insert every record of the sequence in the level's resident hash set; records
//             not yet resident keep the MINIMUM sequence index that carried them (atomicMin);
//   mark    : a record is new iff it is the first occurrence of a non-resident value;
//   scan    : exclusive prefix sum of the new records' sizes = the addresses the sequential
//             reference would have handed out;
//   commit  : write the new records to the level's raw array, make their set entries resident;
//   resolve : every sequence element reads its address.
#include "kernels.cuh"
#include "scan.cuh"

#include <cstdlib>

namespace chadgpu {

namespace {

constexpr int DAG_THREADS = 256;
constexpr u32 FIRST_IDLE = 0xFFFFFFFFu;
constexpr u64 WEIGHT_CLUSTER = ~0ull;  // every weight byte is 0xFF

__device__ __forceinline__ u64 mix64(u64 h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}
__device__ __forceinline__ u64 ld_entry(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_u32_volatile(const u32* p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void note_first(u32* first, u32 e) {
    if (ld_u32_volatile(first) > e) atomicMin(first, e);
}
inline unsigned blocks_for(u32 n) { return (n + DAG_THREADS - 1) / DAG_THREADS; }

// ------------------------------------------------------------------------------------------
// leaf clusters: cluster.hpp:13-32 (TSDFs::set / set_empty), submap.hpp:83-100
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DAG_THREADS) cluster_build_kernel(const uint2* __restrict__ cells, const u32* __restrict__ d_chunks, float trunc_recip,
                                                                    u64* __restrict__ tsdf_values) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= *d_chunks) return;
    u64 v = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const uint2 c = cells[size_t(i) * 8 + s];
        u64 byte = 0xFF;  // absent voxel (cluster.hpp:29-32)
        if (c.y != 0) {
            float sd = fmul(__uint_as_float(c.x), trunc_recip);  // cluster.hpp:19
            sd = fclamp(sd, -1.0f, 1.0f);
            sd = fadd(fmul(sd, 127.0f), 127.0f);                 // cluster.hpp:21
            byte = (u64)sd;                                      // cluster.hpp:26: truncation toward zero
        }
        v |= byte << (8 * s);
    }
    tsdf_values[i] = v;
}

// Sequence of the cluster level: e = 2i -> tsdf cluster of chunk i, e = 2i+1 -> the weight cluster.
// Only position p = 1 (e = 1) ever needs to add the weight cluster; later weight elements are dupes
// of it by construction. Work positions: p = 0 -> e 0, p = 1 -> e 1, p >= 2 -> e 2(p-1).
__device__ __forceinline__ u32 cluster_seq_of(u32 p) { return p < 2 ? p : 2 * (p - 1); }
__device__ __forceinline__ u64 cluster_value(const u64* __restrict__ tsdf_values, u32 e) { return (e & 1u) ? WEIGHT_CLUSTER : tsdf_values[e >> 1]; }

__device__ __forceinline__ u32 cluster_work(const u32* __restrict__ d_chunks) { const u32 c = *d_chunks; return c ? c + 1 : 0; }

__global__ void __launch_bounds__(DAG_THREADS) cluster_probe_kernel(u64* entries, u32* first, u64 capacity, const u64* __restrict__ tsdf_values,
                                                                    const u32* __restrict__ d_chunks, const u64* __restrict__ raw,
                                                                    u32* __restrict__ slot_of, u32* d_error) {
    const u32 p = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (p >= cluster_work(d_chunks)) return;
    const u32 e = cluster_seq_of(p);
    const u64 value = cluster_value(tsdf_values, e);
    const u32 tag = (u32)(mix64(value) >> 32);
    const u64 mask = capacity - 1;
    u64 slot = tag & mask;
    for (u64 probes = 0; probes < capacity; probes++) {
        u64 ent = ld_entry(&entries[slot]);
        if (ent == 0) {
            const u64 mine = (u64(tag) << 32) | (REF_PENDING | e);
            ent = atomicCAS(&entries[slot], 0ull, mine);
            if (ent == 0) { note_first(&first[slot], e); slot_of[p] = (u32)slot; return; }
        }
        if ((u32)(ent >> 32) == tag) {
            const u32 ref = (u32)ent;
            const u64 other = (ref & REF_PENDING) ? cluster_value(tsdf_values, ref & ~REF_PENDING) : raw[ref];
            if (other == value) {
                if (ref & REF_PENDING) note_first(&first[slot], e);
                slot_of[p] = (u32)slot;
                return;
            }
        }
        slot = (slot + 1) & mask;
    }
    atomicOr(d_error, ERRF_DEDUP_FULL);
    slot_of[p] = 0;
}

// is_new is written for all max_work positions (zero beyond the actual count): the prefix sum runs over max_work
__global__ void __launch_bounds__(DAG_THREADS) cluster_mark_kernel(const u64* __restrict__ entries, const u32* __restrict__ first,
                                                                   const u32* __restrict__ slot_of, const u32* __restrict__ d_chunks, u32 max_work,
                                                                   u32* __restrict__ is_new) {
    const u32 p = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (p >= max_work) return;
    if (p >= cluster_work(d_chunks)) { is_new[p] = 0; return; }
    const u32 slot = slot_of[p];
    const u32 ref = (u32)entries[slot];
    is_new[p] = ((ref & REF_PENDING) && first[slot] == cluster_seq_of(p)) ? 1u : 0u;
}

__global__ void __launch_bounds__(DAG_THREADS) cluster_commit_kernel(u64* entries, u32* first, const u32* __restrict__ slot_of,
                                                                     const u32* __restrict__ is_new, const u32* __restrict__ rank,
                                                                     const u32* __restrict__ d_chunks, const u64* __restrict__ tsdf_values,
                                                                     u64* __restrict__ raw, u32 uniques_before) {
    const u32 p = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (p >= cluster_work(d_chunks) || !is_new[p]) return;
    const u32 slot = slot_of[p];
    const u32 addr = uniques_before + 1 + rank[p];  // levels.hpp:125
    raw[addr] = cluster_value(tsdf_values, cluster_seq_of(p));
    entries[slot] = (entries[slot] & 0xFFFFFFFF00000000ull) | addr;
    first[slot] = FIRST_IDLE;
}

// addr_out[2i] = tsdf address, addr_out[2i+1] = weight address of chunk i
__global__ void __launch_bounds__(DAG_THREADS) cluster_resolve_kernel(const u64* __restrict__ entries, const u32* __restrict__ slot_of,
                                                                      const u32* __restrict__ d_chunks, u32* __restrict__ addr_out) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= *d_chunks) return;
    const u32 p = (i == 0) ? 0 : i + 1;
    addr_out[2 * i] = (u32)entries[slot_of[p]];
    addr_out[2 * i + 1] = (u32)entries[slot_of[1]];
}

// ------------------------------------------------------------------------------------------
// node levels: submap.hpp:31-61, levels.hpp:57-88
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 node_tag(const u32* __restrict__ rec) {
    u64 h = rec[0];
#pragma unroll
    for (int q = 1; q < 9; q++) h = mix64(h ^ (u64(rec[q]) << 8) ^ (u64(q) << 48));
    return (u32)(mix64(h) >> 32);
}

// ------------------------------------------------------------------------------------------
// All 20 node levels in ONE persistent kernel (submap.hpp:31-61 bottom-up). The per-level kernels above need the host
// between levels (grid sizes, exact table sizing) and ~12 launches per level; here the CTAs walk the levels together,
// separated by a software grid barrier, and read every count from device memory: Submap::finalize is queued in one go
// and nothing waits for the host. Once a level has few children left, CTA 0 finishes the remaining levels alone.
// ------------------------------------------------------------------------------------------
constexpr int LV_THREADS = 1024;
constexpr u32 LV_SOLO = 1024;  // children from which one CTA takes over (sweep r02h / r02i: 131 072 / 32 768 / 8 192 / 2 048 / 512 -> 6.4 / 6.4 / 4.8 / 4.2 / 3.9 ms of
                               // finalize per bench step: the one-block tail is the slow part, the grid barriers are cheap)

__device__ __forceinline__ void grid_barrier(u32* bar, u32& phase, u32 nctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        phase++;
        atomicAdd(bar, 1u);
        const u32 target = phase * nctas;
        while (ld_u32_volatile(bar) < target) __nanosleep(40);
        __threadfence();
    }
    __syncthreads();
}

// exclusive prefix of partial[0 .. nctas) (nctas <= LV_THREADS) into shared memory; returns the total
__device__ __forceinline__ u64 partial_prefix(const u64* partial, u32 nctas, u64* s_pref, u64* s_warp) {
    const u64 v = (threadIdx.x < nctas) ? ((const volatile u64*)partial)[threadIdx.x] : 0ull;
    u64 total;
    const u64 ex = block_exclusive_scan<u64>(v, s_warp, total);
    if (threadIdx.x < nctas) s_pref[threadIdx.x] = ex;
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(LV_THREADS, 1) dag_levels_kernel(LevelsArgs a) {
    __shared__ u64 s_warp[LV_THREADS / 32];
    __shared__ u64 s_pref[LV_THREADS];
    const u32 tid = threadIdx.x;
    const u32 NT = blockDim.x;  // <= LV_THREADS
    u32 cta = blockIdx.x, nctas = gridDim.x, phase = 0;
    const u32 C = *a.d_chunks;
    u32 n = C;   // children of the level being built
    int cur = 0;
    auto sync = [&]() { if (nctas == 1) __syncthreads(); else grid_barrier(a.bar, phase, gridDim.x); };
    for (int d = 19; d >= 0; d--) {
        if (nctas > 1 && n <= a.solo) {  // uniform over the grid: n comes from memory written before the last barrier
            if (cta != 0) return;
            nctas = 1;
        }
        const u64* __restrict__ ids = a.ids[cur];
        const u32* __restrict__ addr = a.addr[cur];
        u64* __restrict__ ids_next = a.ids[cur ^ 1];
        u32* __restrict__ addr_next = a.addr[cur ^ 1];
        const LevelDev L = a.lv[d];
        const u32 occ = a.counters[d].occupied;  // levels.hpp:77: addresses continue from _occupied_n
        const bool empty_root = (C == 0);        // empty octree: only the root is added, twice (submap.hpp:31-46)
        if (empty_root && d > 0) continue;
        // ---- 1. parents: heads of the groups of children with the same id >> 3, dense index inside this CTA's slice ----
        const u32 per = (n + nctas - 1) / nctas;
        const u32 c0 = min(n, cta * per), c1 = min(n, c0 + per);
        {
            u32 carry = 0;
            for (u32 t0 = c0; t0 < c1; t0 += NT) {
                const u32 i = t0 + tid;
                const bool head = (i < c1) && (i == 0 || (ids[i] >> 3) != (ids[i - 1] >> 3));
                u64 tot;
                const u64 ex = block_exclusive_scan<u64>(head ? 1ull : 0ull, s_warp, tot);
                if (i < c1) a.head_rank[i] = head ? (carry + (u32)ex) : 0xFFFFFFFFu;
                carry += (u32)tot;
            }
            if (tid == 0) a.partial[cta] = carry;
        }
        sync();
        const u32 P = empty_root ? 1u : (u32)partial_prefix(a.partial, nctas, s_pref, s_warp);
        const u32 pbase = empty_root ? 0u : (u32)s_pref[cta];
        const u32 R = 2 * P;  // records of this level: TSDF node then weight node per parent (submap.hpp:33-35)
        // ---- 2. candidate records: 9 words = child mask + children in ascending child index, zero padded (levels.hpp:63-74) ----
        if (empty_root) {
            if (cta == 0 && tid < 18) a.cand[tid] = 0;
        } else {
            for (u32 i = c0 + tid; i < c1; i += NT) {
                const u32 hr = a.head_rank[i];
                if (hr == 0xFFFFFFFFu) continue;
                const u32 p = pbase + hr;
                const u64 pid = ids[i] >> 3;
                u32 rt[9], rw[9];
#pragma unroll
                for (int q = 0; q < 9; q++) { rt[q] = 0; rw[q] = 0; }
                u32 c = 0;
                for (u32 j = i; j < n && (ids[j] >> 3) == pid; j++) {
                    const u32 bit = 1u << (u32)(ids[j] & 7ull);
                    rt[0] |= bit; rw[0] |= bit;
                    c++;
#pragma unroll
                    for (int q = 1; q < 9; q++)
                        if ((u32)q == c) { rt[q] = addr[2 * j]; rw[q] = addr[2 * j + 1]; }
                }
#pragma unroll
                for (int q = 0; q < 9; q++) {
                    a.cand[size_t(2 * p) * 9 + q] = rt[q];
                    a.cand[size_t(2 * p + 1) * 9 + q] = rw[q];
                }
                ids_next[p] = pid;
            }
        }
        sync();
        // ---- 3. probe: find-or-insert every record; a record not yet resident keeps the minimum sequence index that carried it ----
        for (u32 e = cta * NT + tid; e < R; e += nctas * NT) {
            u32 rec[9];
#pragma unroll
            for (int q = 0; q < 9; q++) rec[q] = a.cand[size_t(e) * 9 + q];
            const u32 tag = node_tag(rec);
            const u32 nchild = __popc(rec[0]);
            const u64 mask = L.capacity - 1;
            u64 slot = tag & mask;
            bool done = false;
            for (u64 probes = 0; probes < L.capacity && !done; probes++) {
                u64 ent = ld_entry(&L.entries[slot]);
                if (ent == 0) {
                    const u64 mine = (u64(tag) << 32) | (REF_PENDING | e);
                    ent = atomicCAS(&L.entries[slot], 0ull, mine);
                    if (ent == 0) { note_first(&L.first[slot], e); a.slot_of[e] = (u32)slot; done = true; break; }
                }
                if ((u32)(ent >> 32) == tag) {
                    const u32 ref = (u32)ent;
                    const u32* other = (ref & REF_PENDING) ? (a.cand + size_t(ref & ~REF_PENDING) * 9) : (L.raw + ref);
                    bool eq = (other[0] & 0xFFu) == rec[0];   // levels.hpp:27-44: same mask and the same children
                    for (u32 q = 1; eq && q <= nchild; q++) eq = other[q] == rec[q];
                    if (eq) {
                        if (ref & REF_PENDING) note_first(&L.first[slot], e);
                        a.slot_of[e] = (u32)slot;
                        done = true;
                        break;
                    }
                }
                slot = (slot + 1) & mask;
            }
            if (!done) { atomicOr(a.d_error, ERRF_DEDUP_FULL); a.slot_of[e] = 0; }
        }
        sync();
        // ---- 4. first occurrences of new records and their running size inside this CTA's slice of the sequence ----
        const u32 per2 = (R + nctas - 1) / nctas;
        const u32 e0 = min(R, cta * per2), e1 = min(R, e0 + per2);
        {
            u64 carry = 0;
            for (u32 t0 = e0; t0 < e1; t0 += NT) {
                const u32 e = t0 + tid;
                u64 v = 0;
                if (e < e1) {
                    const u32 slot = a.slot_of[e];
                    const u32 ref = (u32)ld_entry(&L.entries[slot]);
                    const bool fresh = (ref & REF_PENDING) && ld_u32_volatile(&L.first[slot]) == e;
                    if (fresh) v = (1ull << 32) | (1u + (u32)__popc(a.cand[size_t(e) * 9]));
                }
                u64 tot;
                const u64 ex = block_exclusive_scan<u64>(v, s_warp, tot);
                if (e < e1) a.rank[e] = (carry + ex) | (v ? (1ull << 63) : 0ull);  // bit 63: this element is a first occurrence
                carry += tot;
            }
            if (tid == 0) a.partial[LV_THREADS + cta] = carry;
        }
        sync();
        const u64 new_total = partial_prefix(a.partial + LV_THREADS, nctas, s_pref, s_warp);  // (new records << 32) | new words
        // ---- 5. commit the new records and resolve every element's address (levels.hpp:76-87) ----
        for (u32 e = e0 + tid; e < e1; e += NT) {
            const u32 slot = a.slot_of[e];
            const u64 rk = a.rank[e];
            u32 address;
            if (rk >> 63) {
                address = occ + (u32)(s_pref[cta] + rk);  // low 32 bits: words before this record
                const u32 words = 1u + (u32)__popc(a.cand[size_t(e) * 9]);
                for (u32 q = 0; q < words; q++) L.raw[address + q] = a.cand[size_t(e) * 9 + q];
                __threadfence();
                atomicExch(&L.entries[slot], (ld_entry(&L.entries[slot]) & 0xFFFFFFFF00000000ull) | address);
                __threadfence();
                atomicExch(&L.first[slot], FIRST_IDLE);
            } else {
                while (true) {
                    const u64 ent = ld_entry(&L.entries[slot]);
                    if (!((u32)ent & REF_PENDING)) { address = (u32)ent; break; }
                    const u32 f = ld_u32_volatile(&L.first[slot]);
                    if (f != FIRST_IDLE) {  // the first occurrence has not committed yet: its address is already determined
                        const u64 frk = ((const volatile u64*)a.rank)[f];
                        address = occ + (u32)(s_pref[f / per2] + frk);
                        break;
                    }
                }
            }
            addr_next[e] = address;
        }
        sync();
        if (cta == 0 && tid == 0) {
            const u32 fresh = (u32)(new_total >> 32), words = (u32)new_total;
            a.counters[d].occupied = occ + words;   // levels.hpp:79-81
            a.counters[d].uniques += fresh;
            a.counters[d].dupes += R - fresh;       // levels.hpp:83-86
            a.level_nodes[d] = P;
        }
        cur ^= 1;
        n = P;
    }
    if (tid == 0) { a.root_out[0] = a.addr[cur][0]; a.root_out[1] = a.addr[cur][1]; }
}

// ------------------------------------------------------------------------------------------
// DAG read path on the device: NodeLevels::get_child_addr (levels.hpp:147-161) down the 20 node levels, then
// try_get_lc (levels.hpp:177-192) and the voxel's byte of the leaf cluster (cluster.hpp:34-52). One thread per query.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DAG_THREADS) dag_query_kernel(DagReadArgs a, u32 root, const u64* __restrict__ keys, u32 n, u8* __restrict__ out) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= n) return;
    const u64 key = keys[i];
    u32 addr = root;
    u8 byte = 0xFF;  // absent
    for (u32 depth = 0; depth <= 20 && addr != 0; depth++) {
        if (depth == 20) { byte = (u8)(a.clusters[addr] >> (8 * (key & 7ull))); break; }
        const u32 child_i = (u32)(key >> (3 * (20 - depth))) & 7u;   // depth d picks Morton triple 20 - d
        const u32 mask = a.raw[depth][addr];
        const u32 bit = 1u << child_i;
        addr = (mask & bit) ? a.raw[depth][addr + 1 + __popc(mask & (bit - 1) & 0xFFu)] : 0u;
    }
    out[i] = byte;
}

// ------------------------------------------------------------------------------------------
// Restoring a saved map: the level arrays come back from a file, the dedup sets are rebuilt from them so that later
// Submap::finalize calls deduplicate against everything the saved map ever added (levels.hpp:90-93,141-143: the sets
// are never cleared). Every stored record is distinct, so insertion is a plain first-free probe.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DAG_THREADS) restore_clusters_kernel(u64* entries, u64 capacity, const u64* __restrict__ raw, u32 uniques) {
    const u32 a = blockIdx.x * DAG_THREADS + threadIdx.x + 1;  // address 0 is reserved (levels.hpp:119-120)
    if (a > uniques) return;
    const u32 tag = (u32)(mix64(raw[a]) >> 32);
    const u64 mask = capacity - 1;
    u64 slot = tag & mask;
    const u64 ent = (u64(tag) << 32) | a;
    for (u64 probes = 0; probes < capacity; probes++) {
        if (atomicCAS(&entries[slot], 0ull, ent) == 0ull) return;
        slot = (slot + 1) & mask;
    }
}
__global__ void __launch_bounds__(DAG_THREADS) restore_nodes_kernel(u64* entries, u64 capacity, const u32* __restrict__ raw, const u32* __restrict__ starts,
                                                                    u32 n_records) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= n_records) return;
    const u32 a = starts[i];
    u32 rec[9];
    rec[0] = raw[a] & 0xFFu;
    const u32 nchild = __popc(rec[0]);
#pragma unroll
    for (u32 q = 1; q < 9; q++) rec[q] = (q <= nchild) ? raw[a + q] : 0u;  // the probe's candidate records are zero padded
    const u32 tag = node_tag(rec);
    const u64 mask = capacity - 1;
    u64 slot = tag & mask;
    const u64 ent = (u64(tag) << 32) | a;
    for (u64 probes = 0; probes < capacity; probes++) {
        if (atomicCAS(&entries[slot], 0ull, ent) == 0ull) return;
        slot = (slot + 1) & mask;
    }
}

// ------------------------------------------------------------------------------------------
// Leaf iterator over a finalised submap's tree on the device (the reader the reference sketches but never finishes:
// tsdf.hpp:120-155, tsdf.cpp:88-159 walk root -> first leaf cluster with get_child_addr / try_get_lc). Data-parallel
// form: the tree is expanded one depth at a time; a node's children are appended in child order behind an exclusive
// prefix sum of the popcounts, so every frontier -- and finally the voxel list -- is in ascending Morton order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DAG_THREADS) iter_count_kernel(const u32* __restrict__ raw, const u32* __restrict__ addr, const u32* __restrict__ d_n,
                                                                 u32 max_n, u32* __restrict__ counts) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= max_n) return;
    counts[i] = (i < *d_n) ? (u32)__popc(raw[addr[i]] & 0xFFu) : 0u;
}
__global__ void __launch_bounds__(DAG_THREADS) iter_expand_kernel(const u32* __restrict__ raw, const u32* __restrict__ addr, const u64* __restrict__ prefix,
                                                                  const u32* __restrict__ d_n, const u32* __restrict__ offsets, u32 capacity,
                                                                  u32* __restrict__ addr_next, u64* __restrict__ prefix_next, u32* d_overflow) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= *d_n) return;
    const u32 a = addr[i];
    const u32 mask = raw[a] & 0xFFu;
    u32 out = offsets[i], c = 0;
    for (u32 child = 0; child < 8; child++) {
        if (!(mask & (1u << child))) continue;
        c++;
        if (out < capacity) { addr_next[out] = raw[a + c]; prefix_next[out] = (prefix[i] << 3) | child; }
        else atomicOr(d_overflow, 1u);
        out++;
    }
}
// frontier of leaf clusters (address, cluster id = key >> 3) -> per cluster the number of present voxels
__global__ void __launch_bounds__(DAG_THREADS) iter_leaf_count_kernel(const u64* __restrict__ clusters, const u32* __restrict__ addr, const u32* __restrict__ d_n,
                                                                      u32 max_n, u32* __restrict__ counts) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= max_n) return;
    u32 c = 0;
    if (i < *d_n) {
        const u64 v = clusters[addr[i]];
#pragma unroll
        for (int s = 0; s < 8; s++) c += (((v >> (8 * s)) & 0xFFull) != 0xFFull) ? 1u : 0u;
    }
    counts[i] = c;
}
__global__ void __launch_bounds__(DAG_THREADS) iter_leaf_emit_kernel(const u64* __restrict__ clusters, const u32* __restrict__ addr, const u64* __restrict__ prefix,
                                                                     const u32* __restrict__ d_n, const u32* __restrict__ offsets, u32 capacity,
                                                                     u64* __restrict__ keys, u8* __restrict__ bytes) {
    const u32 i = blockIdx.x * DAG_THREADS + threadIdx.x;
    if (i >= *d_n) return;
    const u64 v = clusters[addr[i]];
    u32 out = offsets[i];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const u8 b = (u8)(v >> (8 * s));
        if (b == 0xFF) continue;
        if (out < capacity) { keys[out] = (prefix[i] << 3) | (u64)s; bytes[out] = b; }
        out++;
    }
}

__global__ void __launch_bounds__(DAG_THREADS) dedup_rehash_kernel(const u64* __restrict__ from, u64 from_capacity, u64* to, u64 to_capacity) {
    const u64 mask = to_capacity - 1;
    for (u64 s = u64(blockIdx.x) * DAG_THREADS + threadIdx.x; s < from_capacity; s += u64(gridDim.x) * DAG_THREADS) {
        const u64 ent = from[s];
        if (ent == 0) continue;
        u64 slot = (ent >> 32) & mask;
        while (true) {  // all entries are distinct records: plain first-free insertion
            if (atomicCAS(&to[slot], 0ull, ent) == 0ull) break;
            slot = (slot + 1) & mask;
        }
    }
}

}  // namespace

int launch_dedup_clear(cudaStream_t s, const DedupTable& t) {
    cudaMemsetAsync(t.entries, 0, t.capacity * 8, s);
    cudaMemsetAsync(t.first, 0xFF, t.capacity * 4, s);
    return 0;
}

int launch_dedup_rehash(cudaStream_t s, const DedupTable& from, const DedupTable& to, int num_sms) {
    launch_dedup_clear(s, to);
    dedup_rehash_kernel<<<num_sms * 4, DAG_THREADS, 0, s>>>(from.entries, from.capacity, to.entries, to.capacity);
    return 1;
}

int launch_cluster_build(cudaStream_t s, const void* gathered_cells, const u32* d_chunks, u32 max_chunks, const MapParams& mp, u64* tsdf_values) {
    if (!max_chunks) return 0;
    cluster_build_kernel<<<blocks_for(max_chunks), DAG_THREADS, 0, s>>>(static_cast<const uint2*>(gathered_cells), d_chunks, mp.trunc_recip, tsdf_values);
    return 1;
}

int launch_cluster_dedup(cudaStream_t s, const DedupTable& t, const u64* tsdf_values, const u32* d_chunks, u32 max_chunks, u64* raw, u32 uniques_before,
                         u32* slot_of, u32* is_new, u32* rank, void* scan_ws, u32* addr_out, u32* d_new_count, u32* d_error) {
    if (!max_chunks) { cudaMemsetAsync(d_new_count, 0, 4, s); return 0; }
    const u32 max_work = max_chunks + 1;
    cluster_probe_kernel<<<blocks_for(max_work), DAG_THREADS, 0, s>>>(t.entries, t.first, t.capacity, tsdf_values, d_chunks, raw, slot_of, d_error);
    cluster_mark_kernel<<<blocks_for(max_work), DAG_THREADS, 0, s>>>(t.entries, t.first, slot_of, d_chunks, max_work, is_new);
    int launches = 2 + exclusive_scan<u32, u32>(s, is_new, rank, max_work, scan_ws, d_new_count);
    cluster_commit_kernel<<<blocks_for(max_work), DAG_THREADS, 0, s>>>(t.entries, t.first, slot_of, is_new, rank, d_chunks, tsdf_values, raw, uniques_before);
    cluster_resolve_kernel<<<blocks_for(max_chunks), DAG_THREADS, 0, s>>>(t.entries, slot_of, d_chunks, addr_out);
    return launches + 2;
}

int launch_dag_levels(cudaStream_t s, const LevelsArgs& args_in, int num_sms) {
    LevelsArgs args = args_in;
    static const u32 env_solo = [] { const char* e = std::getenv("CHAD_LEVELS_SOLO"); const long v = e ? std::atol(e) : 0; return (u32)(v > 0 ? v : LV_SOLO); }();
    args.solo = env_solo;  // children from which ONE block finishes the remaining levels alone (no grid barriers)
    cudaMemsetAsync(args.bar, 0, 4, s);
    static const int env_ctas = [] { const char* e = std::getenv("CHAD_LEVELS_CTAS"); return e ? std::atoi(e) : 0; }();
    // one CTA of 512 threads per SM: the CTAs spin at the grid barriers while the insert kernels of the next submap run beside them, so
    // they must leave registers and thread slots free (measured per bench step: 1024 threads x 148 / 74 / 37 CTAs -> 11.68 / 11.39 /
    // 11.36 ms; with the fold at 3 CTAs per SM: 1024 x 74 -> 10.36 ms, 512 x 148 -> 9.70 ms)
    static const int env_threads = [] { const char* e = std::getenv("CHAD_LEVELS_THREADS"); return e ? std::atoi(e) : 0; }();
    int threads = 512;
    if (env_threads >= 64 && env_threads <= LV_THREADS && env_threads % 32 == 0) threads = env_threads;
    // The software grid barrier needs every CTA resident at the same time. A cooperative launch makes the driver guarantee it (the grid
    // starts only once all of it fits, whatever else -- other contexts, MPS partitions, the persistent fold -- holds SM resources) and
    // the grid is bounded by what the occupancy calculator says one device can hold. If the device cannot launch cooperatively the
    // levels are built by ONE CTA (the kernel's solo path: __syncthreads only), which cannot deadlock.
    static const int coop_ok = [] {
        const char* e = std::getenv("CHAD_LEVELS_COOP");
        if (e && std::atoi(e) == 0) return 0;
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) { cudaGetLastError(); return -1; }
        return v ? 1 : -1;
    }();
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dag_levels_kernel, threads, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
    int grid = num_sms < LV_THREADS ? num_sms : LV_THREADS;
    if (env_ctas > 0 && env_ctas <= num_sms && env_ctas <= LV_THREADS) grid = env_ctas;
    if (grid > threads) grid = threads;  // partial_prefix scans one partial per thread
    if (per_sm < 1) grid = 1;
    else if (grid > per_sm * num_sms) grid = per_sm * num_sms;
    if (grid < 1) grid = 1;
    if (coop_ok == 1 && grid > 1) {
        LevelsArgs a = args;
        void* params[] = {&a};
        if (cudaLaunchCooperativeKernel((const void*)dag_levels_kernel, dim3(grid), dim3(threads), params, 0, s) == cudaSuccess) return 1;
        cudaGetLastError();
        grid = 1;  // rejected (e.g. a partitioned device): the one-CTA path
    } else if (coop_ok == -1) {
        grid = 1;
    }
    // coop_ok == 0 (CHAD_LEVELS_COOP=0): the round-1 plain launch, kept for A/B timing only
    dag_levels_kernel<<<grid, threads, 0, s>>>(args);
    return 1;
}

int launch_dedup_restore(cudaStream_t s, const DedupTable& t, bool cluster, const void* raw, const u32* starts, u32 n_records) {
    launch_dedup_clear(s, t);
    if (!n_records) return 0;
    if (cluster) restore_clusters_kernel<<<blocks_for(n_records), DAG_THREADS, 0, s>>>(t.entries, t.capacity, static_cast<const u64*>(raw), n_records);
    else restore_nodes_kernel<<<blocks_for(n_records), DAG_THREADS, 0, s>>>(t.entries, t.capacity, static_cast<const u32*>(raw), starts, n_records);
    return 1;
}

// One depth of the leaf iterator: frontier (addr, prefix)[*d_n] at node level `raw` -> the children, in order. counts / offsets: work
// arrays of `capacity` u32; *d_n_next = number of children. Returns the kernels queued.
int launch_iter_expand(cudaStream_t s, const u32* raw, const u32* addr, const u64* prefix, const u32* d_n, u32 capacity, u32* counts, u32* offsets,
                       void* scan_ws, u32* addr_next, u64* prefix_next, u32* d_n_next, u32* d_overflow) {
    iter_count_kernel<<<blocks_for(capacity), DAG_THREADS, 0, s>>>(raw, addr, d_n, capacity, counts);
    int launches = 1 + exclusive_scan<u32, u32>(s, counts, offsets, capacity, scan_ws, d_n_next);
    iter_expand_kernel<<<blocks_for(capacity), DAG_THREADS, 0, s>>>(raw, addr, prefix, d_n, offsets, capacity, addr_next, prefix_next, d_overflow);
    return launches + 1;
}
// the last depth: frontier of leaf clusters -> (Morton key, quantised byte) of every present voxel, ascending; *d_total = their number
int launch_iter_leaves(cudaStream_t s, const u64* clusters, const u32* addr, const u64* prefix, const u32* d_n, u32 capacity, u32* counts, u32* offsets,
                       void* scan_ws, u32 out_capacity, u64* keys, u8* bytes, u32* d_total) {
    iter_leaf_count_kernel<<<blocks_for(capacity), DAG_THREADS, 0, s>>>(clusters, addr, d_n, capacity, counts);
    int launches = 1 + exclusive_scan<u32, u32>(s, counts, offsets, capacity, scan_ws, d_total);
    if (keys) {
        iter_leaf_emit_kernel<<<blocks_for(capacity), DAG_THREADS, 0, s>>>(clusters, addr, prefix, d_n, offsets, out_capacity, keys, bytes);
        launches++;
    }
    return launches;
}

int launch_dag_query(cudaStream_t s, const DagReadArgs& args, u32 root, const u64* keys, u32 n, u8* out) {
    if (!n) return 0;
    dag_query_kernel<<<blocks_for(n), DAG_THREADS, 0, s>>>(args, root, keys, n, out);
    return 1;
}

}  // namespace chadgpu
