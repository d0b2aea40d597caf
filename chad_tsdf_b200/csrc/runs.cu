// Tile-run grouping of the band-voxel updates and the streaming fold: the default replacement for the octree's
// per-voxel grouping and running average (/root/reference/include/chad/detail/octree.hpp:31-78,86-164). Successor of
// blocks.cu (which stays as pair path 0): ncu r01b showed the count / emit / sort / fold chain issue bound (1360 warp
// instructions per warp in the count walk alone, DRAM below 17 % everywhere), so this path removes work:
//   * runs_emit_kernel walks every ray ONCE (no counting pass): a CTA takes a tile of 256 consecutive sorted rays, keeps
//     the tile's updates in shared memory, groups them by 8x8x8-voxel block with a shared-memory hash, reserves one
//     contiguous span of the record buffer with ONE global atomic and writes one "run" per (tile, block) -- its records
//     in (ray, step) = rank order -- plus a 16-byte run descriptor. No global hash table, no per-ray global atomics.
//   * The run descriptors (a few hundred thousand) are sorted by (block, tile) with the onesweep sort: the concatenation
//     of a block's runs then IS its update stream in the reference's order (sorted point rank, then ray step; a ray
//     touches a voxel at most once).
//   * runs_fold_kernel gives every block to one WARP that streams that sequence, 32 updates per iteration, into a
//     shared-memory copy of the block's 64 leaf chunks (lanes that share a voxel go in lane order) and writes the
//     touched chunks back: one table probe per 2x2x2 chunk instead of one per voxel, nothing is sorted, and the
//     (key, sd) stream never goes back to HBM.
// Results are bit-identical to paths 0 / 1 and to the CPU reference.
#include "kernels.cuh"
#include "points.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#include <algorithm>
#include <cstdlib>

namespace chadgpu {

namespace {

constexpr int RUN_THREADS = 256;
constexpr u32 RUN_HASH = 1024;             // shared-memory slots for the distinct blocks of a tile (>= 256 rays x 4 runs)
constexpr u32 RUN_BLK_SHIFT = 9;           // 8^3 voxels per block
constexpr u32 RUN_RANK_BITS = 23;          // sorted-point rank inside the batch
constexpr u64 RUN_EMPTY = ~0ull;
constexpr u32 RUN_BIG_BLOCK = 2048;        // updates from which a block is scheduled ahead of the others
constexpr u32 RUN_STASH_BITS = 50;         // a descriptor key carries its run's record count from this bit up when the radix sort does not look there
// The sort works in whole 8-bit digits: with nbits key bits it orders bits [0, 8 * ceil(nbits / 8)), so the count may only sit above
// THAT width (48 for nbits <= 48), not above nbits. Round 1 tested `nbits <= 50`: at 49 or 50 key bits -- voxel coordinates beyond
// +-8192 in batches of 5 to 16 scans -- the last digit then contained six bits of the count, the runs of a block were no longer
// adjacent after the sort, several warps folded the same block and chunks were inserted twice. Found by bench.py's hash check on the
// 1000-scan urban drive (profiles/probe_cfg3_r02.md); no shorter test reaches those widths.
__device__ __forceinline__ bool run_stash(u32 nbits_blocks) { return radix_num_passes(nbits_blocks) * RS_RADIX_BITS <= RUN_STASH_BITS; }
__device__ __forceinline__ u64 run_key_mask(u32 nbits_blocks) { return run_stash(nbits_blocks) ? ((1ull << RUN_STASH_BITS) - 1ull) : ~0ull; }

constexpr int RF_THREADS = 128;             // fold: four independent warps per CTA, one block per warp at a time
constexpr u32 RF_VOXELS = 512;
constexpr int RF_DEPTH = 4;                // 32-update chunks a fold warp keeps in flight

// ---- lean ray walk: octree.hpp:92-152 with the state in scalars and a branch-free step ------------------------
struct RayL {
    float px, py, pz;
    i32 cx, cy, cz;   // current voxel
    i32 ex, ey, ez;   // vf + step: the walk ends when the stepped axis reaches it (octree.hpp:131,138,145)
    i32 sx, sy, sz;
    float tx, ty, tz; // tMax
    float dx, dy, dz; // tDelta
};
__device__ __forceinline__ void axis_setup(float p, float d, float invl, float res, float trunc, float recip, i32& c, i32& e, i32& st, float& tmax,
                                           float& delta, u32& rmax) {
    const float dir = fmul(d, invl);
    const float dir_recip = __frcp_rn(dir);                   // :93  (1.0f / dir, correctly rounded)
    const float start = fsub(p, fmul(dir, trunc));            // :94
    const float fin = fadd(p, fmul(dir, trunc));              // :95
    const float sv = fmul(start, recip);
    const float fl = floorf(sv);
    const i32 vs = (i32)fl;                                   // :96
    const i32 vf = (i32)floorf(fmul(fin, recip));             // :97
    const i32 dv = vf - vs;
    st = (0 < dv) - (dv < 0);                                 // :100
    delta = fabsf(fmul(res, dir_recip));                      // :102
    float m;                                                  // :104-116
    if (st < 0) m = fmul(res, fl);
    else if (st > 0) m = fmul(res, ceilf(sv));
    else m = 3.402823466e+38f;
    m = fsub(m, start);                                       // :117
    tmax = fabsf(fmul(m, dir_recip));                         // :118
    c = vs;
    e = vf + st;
    rmax = max(rmax, max(rcode(vs), rcode(vf)));              // the walk is monotone: its extremes are vs and vf
}
__device__ __forceinline__ u32 rayl_setup(RayL& r, float px, float py, float pz, float ox, float oy, float oz, float res, float trunc, float recip) {
    r.px = px; r.py = py; r.pz = pz;
    const float dx = fsub(px, ox), dy = fsub(py, oy), dz = fsub(pz, oz);
    const float invl = __frcp_rn(fsqrt(dot3(dx, dy, dz, dx, dy, dz)));  // normalize, :92
    u32 rmax = 0;
    axis_setup(px, dx, invl, res, trunc, recip, r.cx, r.ex, r.sx, r.tx, r.dx, rmax);
    axis_setup(py, dy, invl, res, trunc, recip, r.cy, r.ey, r.sy, r.ty, r.dy, rmax);
    axis_setup(pz, dz, invl, res, trunc, recip, r.cz, r.ez, r.sz, r.tz, r.dz, rmax);
    return rmax;
}
// one iteration of octree.hpp:125-152; returns false on `break`; crossed = the step left the current 8^3 block
__device__ __forceinline__ bool rayl_advance(RayL& r, bool& crossed) {
    const bool xy = r.tx < r.ty, xz = r.tx < r.tz, yz = r.ty < r.tz;
    const bool ax = xy && xz, ay = !xy && yz;  // else z  (:126-150: x if tx<ty && tx<tz; y if !(tx<ty) && ty<tz; else z)
    const bool az = !ax && !ay;
    const i32 ox = r.cx, oy = r.cy, oz = r.cz;
    r.cx += ax ? r.sx : 0;
    r.cy += ay ? r.sy : 0;
    r.cz += az ? r.sz : 0;
    r.tx = ax ? fadd(r.tx, r.dx) : r.tx;
    r.ty = ay ? fadd(r.ty, r.dy) : r.ty;
    r.tz = az ? fadd(r.tz, r.dz) : r.tz;
    crossed = (((ox ^ r.cx) | (oy ^ r.cy) | (oz ^ r.cz)) >> 3) != 0;
    return ax ? (r.cx != r.ex) : (ay ? (r.cy != r.ey) : (r.cz != r.ez));
}
// low three bits of a coordinate spread to bits 0, 3, 6
__device__ __forceinline__ u32 spread_low3(i32 c) {
    const u32 l = (u32)c & 7u;
    return (l | (l << 2) | (l << 4)) & 0x49u;
}
__device__ __forceinline__ u32 local_voxel(i32 x, i32 y, i32 z) { return spread_low3(x) | (spread_low3(y) << 1) | (spread_low3(z) << 2); }

// shared-memory hash of the tile's distinct blocks, keyed by the packed block coordinates (cheap to form per crossing;
// the Morton block id is only computed once per distinct block, when the descriptors are written)
__device__ __forceinline__ u64 pack_block(i32 x, i32 y, i32 z) {
    return (u64)(u32)((x >> 3) + (1 << 17)) | ((u64)(u32)((y >> 3) + (1 << 17)) << 18) | ((u64)(u32)((z >> 3) + (1 << 17)) << 36);
}
__device__ __forceinline__ u64 packed_block_to_morton(u64 pb) {  // (Morton key of the block's first voxel) >> 9
    const u32 mask = (1u << 18) - 1u;
    return (spread3((u32)pb & mask) | (spread3((u32)(pb >> 18) & mask) << 1) | (spread3((u32)(pb >> 36) & mask) << 2));
}
// find-or-insert; the slots in use are also listed (s_list[0 .. *s_nslots)), so that the tile's bookkeeping later
// costs as much as it has distinct blocks (a few dozen), not as much as the table is wide
__device__ __forceinline__ u32 tile_hash_insert(u64* s_hkey, unsigned short* s_list, u32* s_nslots, u64 pb) {
    u32 h = ((u32)pb * 0x9E3779B1u + (u32)(pb >> 32) * 0x85EBCA77u) >> 22;  // RUN_HASH = 2^10
    for (u32 probes = 0; probes < RUN_HASH; probes++) {
        const u64 cur = s_hkey[h];
        if (cur == pb) return h;
        if (cur == RUN_EMPTY) {
            const u64 old = atomicCAS(&s_hkey[h], RUN_EMPTY, pb);
            if (old == RUN_EMPTY) { s_list[atomicAdd(s_nslots, 1u)] = (unsigned short)h; return h; }
            if (old == pb) return h;
        }
        h = (h + 1) & (RUN_HASH - 1);
    }
    return 0xFFFFFFFFu;  // cannot happen: a tile has at most 256 x max_ray_runs <= 1024 distinct blocks
}

// ---- 1. walk + emit ---------------------------------------------------------------------------------
// record = (sd bits << 32) | (local voxel << 23) | sorted-point rank. Inside a run the records are written in RANK
// order (ray by ray), and the descriptors sort by (block, tile): the concatenation of a block's runs is its update
// stream in the reference's order, which the fold then consumes sequentially without sorting anything.
constexpr u32 RUN_WARPS = RUN_THREADS / 32;
__global__ void __launch_bounds__(RUN_THREADS) runs_emit_kernel(const float* __restrict__ xyz_sorted, const float* __restrict__ normals, u32 n_points,
                                                                const BatchScans* __restrict__ scans, float res, float trunc, float recip, u32 mrv,
                                                                BatchPlan* plan, u64* __restrict__ records, u32 rec_capacity,
                                                                u64* __restrict__ desc_key, u32* __restrict__ desc_val, uint2* __restrict__ desc,
                                                                u32 desc_capacity, u32 order_rank) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint2* s_rec = reinterpret_cast<uint2*>(s_dyn);  // [RUN_THREADS][stride]: x = (slot << 9) | local voxel, y = sd bits
    __shared__ u64 s_hkey[RUN_HASH];
    __shared__ u32 s_wcnt[RUN_WARPS / 2][RUN_HASH];  // records per (warp, slot), two warps per word; later: per-warp cursors inside the slot
    __shared__ unsigned short s_hbase[RUN_HASH];     // first record of the slot's run inside the tile's span (a tile holds <= 8192 records)
    __shared__ u32 s_warp[RUN_THREADS / 32];
    __shared__ u32 s_gbase, s_dbase;
    __shared__ u32 s_scan[4];  // the tile's scan, first ray, end, order key
    __shared__ u32 s_prefix[RUN_WARPS][33];  // per warp: records of the lower lanes' rays (write-out)
    __shared__ unsigned short s_list[RUN_HASH];    // hash slots in use; their runs follow each other in this order
    __shared__ u32 s_total;
    __shared__ u32 s_nslots;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 stride = mrv | 1u;  // odd: the ray-major reads of the write-out and the step-major writes of the walk both spread over the banks
    {   // clear the hash (16-byte stores)
        uint4* hk = reinterpret_cast<uint4*>(s_hkey);
        uint4* wc4 = reinterpret_cast<uint4*>(&s_wcnt[0][0]);
        const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u), zero = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (u32 q = 0; q < RUN_HASH * 8 / 16 / RUN_THREADS; q++) hk[tid + q * RUN_THREADS] = ones;
#pragma unroll
        for (u32 q = 0; q < (RUN_WARPS / 2) * RUN_HASH * 4 / 16 / RUN_THREADS; q++) wc4[tid + q * RUN_THREADS] = zero;
        if (tid == 0) s_nslots = 0;
    }
    // Tiles never straddle scans: tile t of scan s holds the sorted rays offset[s] + 256 t .. of that scan, so that the order key
    // (scan | Morton-range rank, descending | tile in scan) sorts the runs of a block in the reference's update order, also when the
    // block's runs come from several GPUs
    if (tid == 0) {
        const u32 ns = plan->n_scans;
        u32 lo = 0, hi = ns;  // tile_prefix[lo] <= blockIdx.x < tile_prefix[hi] when the tile exists
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (scans->tile_prefix[mid] <= blockIdx.x) lo = mid; else hi = mid; }
        const bool exists = ns > 0 && blockIdx.x < scans->tile_prefix[ns];
        const u32 t = blockIdx.x - scans->tile_prefix[lo];
        const u32 first = scans->offset[lo] + t * RUN_THREADS;
        s_scan[0] = lo;
        s_scan[1] = exists ? first : 0xFFFFFFFFu;
        s_scan[2] = exists ? min(first + RUN_THREADS, scans->offset[lo + 1]) : 0u;
        s_scan[3] = (lo << (plan->tile_bits - bits_for(ns))) | order_rank | t;  // order_rank = (world - 1 - rank) << tsb
    }
    if (lane == 0) s_prefix[warp][32] = 0xFFFFFFFFu;  // sentinel of the write-out's binary search
    __syncthreads();
    const u32 tile0 = s_scan[1];
    if (tile0 == 0xFFFFFFFFu) return;  // (the grid is sized from an upper bound of the tile count)
    const u32 tile_end = min(min(s_scan[2], n_points), plan->n_points);
    const u32 okey = s_scan[3];
    const u32 i = tile0 + tid;
    u32 cnt = 0, err = 0;
    RayL r;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    bool alive = false;
    u32 slot = 0, run = 0;
    const u32 wsh = (warp & 1u) * 16u;
    if (i < tile_end) {
        const u32 s = s_scan[0];
        nx = normals[size_t(i) * 3]; ny = normals[size_t(i) * 3 + 1]; nz = normals[size_t(i) * 3 + 2];
        const u32 rmax = rayl_setup(r, xyz_sorted[size_t(i) * 3], xyz_sorted[size_t(i) * 3 + 1], xyz_sorted[size_t(i) * 3 + 2], scans->pose[s][0],
                                    scans->pose[s][1], scans->pose[s][2], res, trunc, recip);
        if (rmax >= (1u << 20)) err |= ERRF_RANGE;
        else {
            alive = true;
            slot = tile_hash_insert(s_hkey, s_list, &s_nslots, pack_block(r.cx, r.cy, r.cz));
        }
    }
    // warp-uniform walk: every lane iterates until the warp's longest ray is done, so the lanes stay converged
    while (__any_sync(0xffffffffu, alive)) {
        if (alive) {
            // octree.hpp:157-159: the voxel's LOWER CORNER projected on the normal, clamped to +-trunc
            float sd = dot3(nx, ny, nz, fsub(fmul((float)r.cx, res), r.px), fsub(fmul((float)r.cy, res), r.py), fsub(fmul((float)r.cz, res), r.pz));
            sd = fclamp(sd, -trunc, trunc);
            s_rec[tid * stride + cnt] = make_uint2((slot << 9) | local_voxel(r.cx, r.cy, r.cz), __float_as_uint(sd));
            cnt++;
            run++;
            bool crossed = false;
            bool more = rayl_advance(r, crossed);
            if (more && cnt >= mrv) { err |= ERRF_PAIR_CAPACITY; more = false; }  // the analytic per-ray bound was exceeded: report, never truncate silently
            if (!more || crossed) {
                if (slot != 0xFFFFFFFFu) atomicAdd(&s_wcnt[warp >> 1][slot], run << wsh); else err |= ERRF_BLOCKS_FULL;
                run = 0;
                if (more) slot = tile_hash_insert(s_hkey, s_list, &s_nslots, pack_block(r.cx, r.cy, r.cz));
                alive = more;
            }
        }
    }
    __syncthreads();
    // ---- one span of the record buffer per tile, one run per distinct block; inside a run the warps follow each other ----
    const u32 nslots = min(s_nslots, RUN_HASH);
    // pass 1 over the slots in use: per-warp cursors inside the run, records of the run
    u32 total = 0;
    for (u32 t0 = 0; t0 < nslots; t0 += RUN_THREADS) {
        const u32 t = t0 + tid;
        u32 sum = 0;
        if (t < nslots) {
            const u32 hs = s_list[t];
#pragma unroll
            for (u32 w = 0; w < RUN_WARPS / 2; w++) {
                const u32 word = s_wcnt[w][hs];
                const u32 lo = word & 0xFFFFu, hi = word >> 16;
                s_wcnt[w][hs] = sum | ((sum + lo) << 16);  // exclusive prefix over the warps = each warp's cursor inside the run
                sum += lo + hi;
            }
        }
        u32 tot;
        const u32 ex = block_exclusive_scan<u32>(sum, s_warp, tot);
        if (t < nslots) s_hbase[s_list[t]] = (unsigned short)(total + ex);
        total += tot;
    }
    if (tid == 0) {
        const u32 nrec = total, nd = nslots;
        u32 gb = 0, db = 0;
        if (nrec) { gb = atomicAdd(&plan->n_pairs, nrec); db = atomicAdd(&plan->n_runs, nd); }
        u32 e = 0;
        if (gb + nrec > rec_capacity || gb + nrec < gb) e |= ERRF_PAIR_CAPACITY;
        if (db + nd > desc_capacity) e |= ERRF_BLOCKS_FULL;
        if (plan->nbits_blocks > 64) e |= ERRF_KEY_BUDGET;
        if (e) { atomicOr(&plan->error, e); gb = 0xFFFFFFFFu; }
        s_gbase = gb;
        s_dbase = db;
        s_total = total;
    }
    __syncthreads();
    const u32 gbase = s_gbase;
    if (gbase != 0xFFFFFFFFu) {
        const u32 k = plan->k, tbits = plan->tile_bits;
        const bool stash = run_stash(plan->nbits_blocks);
        for (u32 t = tid; t < nslots; t += RUN_THREADS) {
            const u32 hs = s_list[t];
            const u32 d = s_dbase + t;
            const u64 cid = compact_key(packed_block_to_morton(s_hkey[hs]) << RUN_BLK_SHIFT, k) >> RUN_BLK_SHIFT;
            const u32 base = s_hbase[hs], next = (t + 1 < nslots) ? (u32)s_hbase[s_list[t + 1]] : s_total;
            desc_key[d] = (cid << tbits) | (u64)okey | (stash ? ((u64)(next - base) << RUN_STASH_BITS) : 0ull);
            desc_val[d] = d;
            desc[d] = make_uint2(gbase + base, next - base);
        }
    }
    __syncthreads();
    if (gbase != 0xFFFFFFFFu) {
        // ray-major write-out: the warp's records, ordered (ray, step), are taken 32 at a time (lane f of the flat sequence finds
        // its ray by binary search over the rays' prefix sums), so that the positions handed out inside a (warp, slot) follow
        // the rays' order; lanes of one iteration that share a slot are ordered by lane
        u32 P = cnt;  // inclusive scan over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, P, o); if (lane >= (u32)o) P += up; }
        const u32 T = __shfl_sync(0xffffffffu, P, 31);
        u32* sp = s_prefix[warp];
        sp[lane] = P - cnt;
        // owner[f] = the ray of flat record f (the block hash is no longer needed: the descriptors are written)
        unsigned char* owner = reinterpret_cast<unsigned char*>(s_hkey) + warp * (RUN_HASH * 8 / RUN_WARPS);
        for (u32 t = 0; t < cnt; t++) owner[P - cnt + t] = (unsigned char)lane;
        __syncwarp();
        u32* wc = s_wcnt[warp >> 1];
        for (u32 f = lane; f - lane < T; f += 32) {
            const bool has = f < T;
            const u32 ray = has ? (u32)owner[f] : 0u;
            uint2 rc = make_uint2(0xFFFFFFFFu, 0);
            if (has) rc = s_rec[(warp * 32 + ray) * stride + (f - sp[ray])];
            const u32 hs = (has && (rc.x >> 9) < RUN_HASH) ? (rc.x >> 9) : (0x80000000u | lane);  // unique for idle lanes
            const u32 m = __match_any_sync(0xffffffffu, hs);
            if (hs < RUN_HASH) {
                const u32 leader = (u32)(__ffs(m) - 1);
                u32 old = 0;
                if (lane == leader) old = atomicAdd(&wc[hs], (u32)__popc(m) << wsh);
                old = __shfl_sync(m, old, leader);
                const u32 pos = gbase + (u32)s_hbase[hs] + ((old >> wsh) & 0xFFFFu) + (u32)__popc(m & ((1u << lane) - 1u));
                records[pos] = (u64(rc.y) << 32) | (u64)(((rc.x & 511u) << RUN_RANK_BITS) | (tile0 + warp * 32 + ray));
            }
        }
    }
    if (err) atomicOr(&plan->error, err);
}

// ---- 2. after the descriptor sort: gather the descriptors in sorted order and list the blocks ------------------
// work[w] = (first sorted descriptor, descriptors) of block w (blocks in arbitrary order: every block is independent)
__global__ void __launch_bounds__(RUN_THREADS) runs_group_kernel(const u64* __restrict__ dkeys_a, const u64* __restrict__ dkeys_b,
                                                                 const u32* __restrict__ dvals_a, const u32* __restrict__ dvals_b,
                                                                 const uint2* __restrict__ desc, uint2* __restrict__ sdesc, BatchPlan* plan,
                                                                 uint2* __restrict__ work, u32 work_capacity) {
    const u32 n = plan->n_runs;
    if (plan->nbits_blocks > 64) return;
    const bool alt = radix_result_in_alt(plan->nbits_blocks);
    const u64* __restrict__ keys = alt ? dkeys_b : dkeys_a;
    const u32* __restrict__ vals = alt ? dvals_b : dvals_a;
    const u32 tbits = plan->tile_bits;
    const u64 kmask = run_key_mask(plan->nbits_blocks);
    const bool stash = run_stash(plan->nbits_blocks);
    const u32 lane = threadIdx.x & 31;
    for (u32 q0 = blockIdx.x * RUN_THREADS; q0 < n; q0 += gridDim.x * RUN_THREADS) {  // uniform per CTA
        const u32 p = q0 + threadIdx.x;
        u64 cid = 0;
        bool head = false;
        if (p < n) {
            cid = (keys[p] & kmask) >> tbits;
            sdesc[p] = desc[vals[p]];
            head = (p == 0) || ((keys[p - 1] & kmask) >> tbits) != cid;
        }
        u32 len = 0, recs = 0;
        if (head) {
            u32 q = p;
            while (q < n) {
                const u64 kq = keys[q];
                if (((kq & kmask) >> tbits) != cid) break;
                recs += stash ? (u32)(kq >> RUN_STASH_BITS) : desc[vals[q]].y;
                q++;
            }
            len = q - p;
        }
        // blocks with many updates go to the front of the list (a block is folded by ONE warp: start the long ones first),
        // the others are listed from the back
        const bool big = head && recs >= RUN_BIG_BLOCK;
        const bool small = head && !big && recs > 0;  // recs == 0: a block of another Morton range, its runs have been sent to their owner
        const u32 bb = __ballot_sync(0xffffffffu, big), bs = __ballot_sync(0xffffffffu, small);
        u32 base_b = 0, base_s = 0;
        if (lane == 0) {
            if (bb) base_b = atomicAdd(&plan->n_big_blocks, (u32)__popc(bb));
            if (bs) base_s = atomicAdd(&plan->n_small_blocks, (u32)__popc(bs));
            if (bb | bs) atomicAdd(&plan->n_blocks, (u32)__popc(bb | bs));
        }
        base_b = __shfl_sync(0xffffffffu, base_b, 0);
        base_s = __shfl_sync(0xffffffffu, base_s, 0);
        const u32 ltm = (1u << lane) - 1u;
        if (big) work[base_b + __popc(bb & ltm)] = make_uint2(p, len);
        if (small) work[work_capacity - 1 - (base_s + __popc(bs & ltm))] = make_uint2(p, len);
    }
}

// ---- 3. fold: one WARP per block streams the block's updates in order into the resident table ----------------
__device__ __forceinline__ u64 mix64(u64 h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
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
__device__ __forceinline__ u64 table_insert(u64* __restrict__ keys, u64 capacity, u64 chunk) {  // the chunk is not in the table
    const u64 mask = capacity - 1;
    u64 h = mix64(chunk) & mask;
    for (u64 probes = 0; probes < capacity; probes++) {
        if (keys[h] == CHUNK_EMPTY && atomicCAS(&keys[h], CHUNK_EMPTY, chunk) == CHUNK_EMPTY) return h;
        h = (h + 1) & mask;
    }
    return ~0ull;
}

struct FoldWarp {
    uint2 cell[RF_VOXELS];   // (sd bits, weight) of the block's 512 voxels: the table's cell layout, chunk c = cells 8c .. 8c+7
    u32 touched[RF_VOXELS / 32];
    u32 rprefix[33];         // the 32 runs being streamed: records before each run ([32] = sentinel) ...
    u32 rstart[32];          // ... and (first record - prefix)
    u32 pad[15];
};

__global__ void __launch_bounds__(RF_THREADS) runs_fold_kernel(const u64* __restrict__ records, const u64* __restrict__ dkeys_a,
                                                               const u64* __restrict__ dkeys_b, const uint2* __restrict__ sdesc,
                                                               const uint2* __restrict__ work, u32 work_capacity, BatchPlan* plan,
                                                               u64* __restrict__ tkeys, uint2* tcells, u64 capacity, u32* tcount,
                                                               u64* __restrict__ tlist) {
    __shared__ __align__(16) FoldWarp s_w[RF_THREADS / 32];
    __shared__ u32 s_red[4];
    const u32 tid = threadIdx.x, lane = tid & 31;
    FoldWarp& S = s_w[tid >> 5];
    if (tid < 4) s_red[tid] = 0;
    if (lane == 0) S.rprefix[32] = 0xFFFFFFFFu;
    __syncthreads();
    const u32 pe = plan->error;
    const u32 nbits = plan->nbits_blocks;
    const bool bad = (pe & (ERRF_BLOCKS_FULL | ERRF_PAIR_CAPACITY | ERRF_RANGE | ERRF_KEY_BUDGET)) != 0 || nbits > 64;
    const u64* __restrict__ dkeys = radix_result_in_alt(nbits) ? dkeys_b : dkeys_a;
    const u32 n_big = plan->n_big_blocks;
    const u32 n_blocks = bad ? 0u : (n_big + plan->n_small_blocks), k = plan->k, tbits = plan->tile_bits;
    auto work_item = [&](u32 t) { return work[t < n_big ? t : work_capacity - 1 - (t - n_big)]; };
    const u32 lt = (1u << lane) - 1u;
    const u64 kmask = run_key_mask(nbits);
    u32 st_segments = 0, st_chunks = 0, st_new = 0, err = 0;
    // ticket + work item of the first block
    u32 t = 0;
    uint2 wk = make_uint2(0, 0);
    if (lane == 0) { t = atomicAdd(&plan->fold_ticket, 1u); if (t < n_blocks) wk = work_item(t); }
    t = __shfl_sync(0xffffffffu, t, 0);
    wk.x = __shfl_sync(0xffffffffu, wk.x, 0);
    wk.y = __shfl_sync(0xffffffffu, wk.y, 0);
    while (t < n_blocks) {
        const u32 p0 = wk.x, nruns = wk.y;
        // lane 0 fetches the NEXT ticket and work item now; they are only waited for at the end of this block
        u32 t_next = 0;
        uint2 w_next = make_uint2(0, 0);
        if (lane == 0) { t_next = atomicAdd(&plan->fold_ticket, 1u); if (t_next < n_blocks) w_next = work_item(t_next); }
        const u64 blk = expand_key(((dkeys[p0] & kmask) >> tbits) << RUN_BLK_SHIFT, k) >> RUN_BLK_SHIFT;
        // ---- seeds: the block's 64 leaf chunks that are already resident (lane l owns chunks l and l + 32) ----
        u64 slot0 = table_find(tkeys, capacity, (blk << 6) | (u64)lane);
        u64 slot1 = table_find(tkeys, capacity, (blk << 6) | (u64)(lane + 32));
        {
            uint4* dst0 = reinterpret_cast<uint4*>(&S.cell[8 * lane]);
            uint4* dst1 = reinterpret_cast<uint4*>(&S.cell[8 * (lane + 32)]);
            const uint4 z = make_uint4(0, 0, 0, 0);  // a voxel touched for the first time starts from (0, 0) (octree.hpp:68-75)
            const uint4* src0 = reinterpret_cast<const uint4*>(tcells + slot0 * 8);
            const uint4* src1 = reinterpret_cast<const uint4*>(tcells + slot1 * 8);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                dst0[q] = (slot0 != ~0ull) ? src0[q] : z;
                dst1[q] = (slot1 != ~0ull) ? src1[q] : z;
            }
            if (lane < RF_VOXELS / 32) S.touched[lane] = 0;
        }
        __syncwarp();
        // ---- stream the runs in (tile, ray, step) order. The runs are taken 32 at a time (one descriptor per lane); their
        //      records form one flat sequence that is consumed 32 updates per iteration (lane f finds its run by binary search
        //      over the runs' prefix sums), with the next RF_DEPTH x 32 updates already in flight: one warp walks its block
        //      alone, so without the look-ahead every iteration would wait one DRAM round trip ----
        for (u32 r0 = 0; r0 < nruns; r0 += 32) {
            const uint2 dl = (r0 + lane < nruns) ? sdesc[p0 + r0 + lane] : make_uint2(0, 0);
            u32 P = dl.y;  // inclusive scan of the run lengths
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, P, o); if (lane >= (u32)o) P += up; }
            const u32 T = __shfl_sync(0xffffffffu, P, 31);
            __syncwarp();
            S.rprefix[lane] = P - dl.y;
            S.rstart[lane] = dl.x - (P - dl.y);  // record f of the flat sequence lives at rstart[run] + f
            __syncwarp();
            // the positions fetched by a lane only grow (by 32 per iteration): its run cursor walks forward, usually by 0 or 1
            u32 frun = 0;
            auto fetch = [&](u32 f) -> u64 {
                if (f >= T) return ~0ull;
                while (S.rprefix[frun + 1] <= f) frun++;  // rprefix[32] is a sentinel
                return records[S.rstart[frun] + f];
            };
            u64 buf[RF_DEPTH];
#pragma unroll
            for (int q = 0; q < RF_DEPTH; q++) buf[q] = fetch(q * 32 + lane);
            for (u32 f = lane; f - lane < T; f += 32) {
                const u64 cur = buf[0];
                const bool cur_valid = f < T;
#pragma unroll
                for (int q = 0; q + 1 < RF_DEPTH; q++) buf[q] = buf[q + 1];
                buf[RF_DEPTH - 1] = fetch(f + RF_DEPTH * 32);
                // ---- apply the 32 updates: lanes that share a voxel go one after the other in lane (= rank) order ----
                const u32 v = cur_valid ? ((u32)cur >> RUN_RANK_BITS) : (0x80000000u | lane);
                const u32 m = __match_any_sync(0xffffffffu, v);
                const u32 ord = (u32)__popc(m & lt);
                const u32 rounds = __reduce_max_sync(0xffffffffu, cur_valid ? (u32)__popc(m) : 0u);
                const float sd = __uint_as_float((u32)(cur >> 32));
                if (cur_valid && ord == 0) atomicOr(&S.touched[v >> 5], 1u << (v & 31u));
                for (u32 q = 0; q < rounds; q++) {
                    if (cur_valid && ord == q) {
                        const uint2 c = S.cell[v];
                        float acc = fadd(fmul(__uint_as_float(c.x), __uint2float_rn(c.y)), sd);  // octree.hpp:161
                        const u32 w = c.y + 1;                                                     // :162
                        acc = fdiv(acc, __uint2float_rn(w));                                       // :163
                        S.cell[v] = make_uint2(__float_as_uint(acc), w);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        // ---- write the touched chunks back (chunk c = byte c of the touched bitmap) ----
        {
            const u32 word0 = S.touched[lane >> 2], word1 = S.touched[(lane + 32) >> 2];
            const u32 b0 = (word0 >> ((lane & 3u) * 8u)) & 0xFFu, b1 = (word1 >> ((lane & 3u) * 8u)) & 0xFFu;
            st_segments += (u32)__popc(b0) + (u32)__popc(b1);
            // chunks touched for the first time in this submap are inserted now and join the table's chunk list
            const bool ins0 = b0 && slot0 == ~0ull, ins1 = b1 && slot1 == ~0ull;
            const u32 m0 = __ballot_sync(0xffffffffu, ins0), m1 = __ballot_sync(0xffffffffu, ins1);
            if (m0 | m1) {
                u32 lbase = 0;
                if (lane == 0) lbase = atomicAdd(tcount, (u32)(__popc(m0) + __popc(m1)));
                lbase = __shfl_sync(0xffffffffu, lbase, 0);
                if (ins0) {
                    slot0 = table_insert(tkeys, capacity, (blk << 6) | (u64)lane);
                    tlist[lbase + __popc(m0 & lt)] = (blk << 6) | (u64)lane;
                    st_new++;
                }
                if (ins1) {
                    slot1 = table_insert(tkeys, capacity, (blk << 6) | (u64)(lane + 32));
                    tlist[lbase + __popc(m0) + __popc(m1 & lt)] = (blk << 6) | (u64)(lane + 32);
                    st_new++;
                }
            }
            if (b0) {
                st_chunks++;
                if (slot0 == ~0ull) err |= ERRF_TABLE_FULL;
                else {
                    uint4* dst = reinterpret_cast<uint4*>(tcells + slot0 * 8);
                    const uint4* src = reinterpret_cast<const uint4*>(&S.cell[8 * lane]);
#pragma unroll
                    for (int q = 0; q < 4; q++) dst[q] = src[q];
                }
            }
            if (b1) {
                st_chunks++;
                if (slot1 == ~0ull) err |= ERRF_TABLE_FULL;
                else {
                    uint4* dst = reinterpret_cast<uint4*>(tcells + slot1 * 8);
                    const uint4* src = reinterpret_cast<const uint4*>(&S.cell[8 * (lane + 32)]);
#pragma unroll
                    for (int q = 0; q < 4; q++) dst[q] = src[q];
                }
            }
        }
        __syncwarp();
        t = __shfl_sync(0xffffffffu, t_next, 0);
        wk.x = __shfl_sync(0xffffffffu, w_next.x, 0);
        wk.y = __shfl_sync(0xffffffffu, w_next.y, 0);
    }
    // ---- counters: one atomic per CTA ----
    st_segments = __reduce_add_sync(0xffffffffu, st_segments);
    st_chunks = __reduce_add_sync(0xffffffffu, st_chunks);
    st_new = __reduce_add_sync(0xffffffffu, st_new);
    err = __reduce_or_sync(0xffffffffu, err);
    if (lane == 0) {
        if (st_segments) atomicAdd(&s_red[0], st_segments);
        if (st_chunks) atomicAdd(&s_red[1], st_chunks);
        if (st_new) atomicAdd(&s_red[2], st_new);
        if (err) atomicOr(&s_red[3], err);
    }
    __syncthreads();
    if (tid == 0) {
        if (s_red[0]) atomicAdd(&plan->n_segments, s_red[0]);
        if (s_red[1]) atomicAdd(&plan->n_chunk_heads, s_red[1]);
        if (s_red[2]) atomicAdd(&plan->n_new_chunks, s_red[2]);
        if (s_red[3]) atomicOr(&plan->error, s_red[3]);
    }
}

// ---- Morton-range sharding: runs of blocks another rank owns travel to that rank -------------
// A ray's band reaches at most a few voxels beyond its point's range, so a few per thousand of the runs are foreign. They
// are moved whole: the key (compact block id, order key) is valid on every rank (k and the order key layout follow from
// the whole batch), so the receiver appends records and descriptors to its own and the descriptor sort interleaves them
// with the local runs in the reference's update order.
__device__ __forceinline__ u32 run_owner(u64 blk, const u64* __restrict__ splitters, u32 world) {
    u32 g = 0;
    for (u32 q = 1; q < world; q++) g += (splitters[q] <= blk) ? 1u : 0u;
    return g;
}

__global__ void __launch_bounds__(RUN_THREADS) runs_pack_kernel(const u64* __restrict__ records, u64* __restrict__ desc_key, uint2* __restrict__ desc,
                                                                BatchPlan* plan, const u64* __restrict__ splitters, u32 rank, u32 world,
                                                                u64* __restrict__ out, u32 words) {
    const u32 n = plan->n_runs;
    if (blockIdx.x == 0 && threadIdx.x == 0) { plan->n_runs_local = n; plan->n_pairs_local = plan->n_pairs; }
    if (plan->nbits_blocks > 64) return;
    const u32 k = plan->k, tbits = plan->tile_bits, lane = threadIdx.x & 31;
    const u64 kmask = run_key_mask(plan->nbits_blocks);
    u32 sent_runs = 0, sent_records = 0;
    for (u32 d0 = blockIdx.x * RUN_THREADS; d0 < n; d0 += gridDim.x * RUN_THREADS) {  // uniform per CTA
        const u32 d = d0 + threadIdx.x;
        u64 key = 0;
        uint2 dd = make_uint2(0, 0);
        u32 owner = rank;
        if (d < n) {
            key = desc_key[d] & kmask;
            dd = desc[d];
            owner = run_owner(expand_key(((key >> tbits) << RUN_BLK_SHIFT), k) >> RUN_BLK_SHIFT, splitters, world);
        }
        const bool foreign = owner != rank;
        u32 todo = __ballot_sync(0xffffffffu, foreign);
        while (todo) {  // the warp moves one foreign run at a time
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const u32 first = __shfl_sync(0xffffffffu, dd.x, l), len = __shfl_sync(0xffffffffu, dd.y, l), ow = __shfl_sync(0xffffffffu, owner, l);
            const u64 rkey = __shfl_sync(0xffffffffu, key, l);
            u64* box = out + size_t(ow) * words;
            u64 old = 0;
            if (lane == 0) old = atomicAdd(&box[0], (u64(len) << 32) | 1ull);  // runs in the low half, records in the high half
            old = __shfl_sync(0xffffffffu, old, 0);
            const u32 dpos = (u32)old, rpos = (u32)(old >> 32);
            if (2ull + 2ull * (dpos + 1ull) + rpos + len <= (u64)words) {
                if (lane == 0) { box[2 + 2 * dpos] = rkey; box[3 + 2 * dpos] = (u64)rpos | (u64(len) << 32); }
                for (u32 q = lane; q < len; q += 32) box[words - 1 - (rpos + q)] = records[first + q];
            } else if (lane == 0) {
                atomicOr(&box[1], 1ull);
                atomicOr(&plan->error, ERRF_EXCHANGE);
            }
            if (lane == 0) { sent_runs++; sent_records += len; }
        }
        if (foreign) {  // the local copy of the run is dead: no records, so the block is not listed for the fold
            desc_key[d] = key;  // (stashed record count removed)
            desc[d] = make_uint2(dd.x, 0);
        }
    }
    if (sent_runs) { atomicAdd(&plan->xfer_runs, sent_runs); atomicAdd(&plan->xfer_records, sent_records); }
}

// grid (CTAs per source, world): the box received from rank blockIdx.y is appended to the local records / descriptors
__global__ void __launch_bounds__(RUN_THREADS) runs_ingest_kernel(u64* __restrict__ records, u32 rec_capacity, u64* __restrict__ desc_key,
                                                                  u32* __restrict__ desc_val, uint2* __restrict__ desc, u32 desc_capacity, BatchPlan* plan,
                                                                  u32 rank, u32 world, const u64* __restrict__ in, u32 words) {
    const u32 src = blockIdx.y;
    if (src == rank || plan->nbits_blocks > 64) return;
    // every CTA derives every source's place from the headers: sources are appended in rank order behind the local runs
    u64 total_r = plan->n_pairs_local, total_d = plan->n_runs_local;
    u64 my_r = 0, my_d = 0;
    u32 my_nd = 0, my_nr = 0;
    bool bad = false;
    for (u32 g = 0; g < world; g++) {
        if (g == rank) continue;
        const u64 hdr = in[size_t(g) * words];
        const u32 nd = (u32)hdr, nr = (u32)(hdr >> 32);
        bad |= (in[size_t(g) * words + 1] & 1ull) != 0 || (2ull + 2ull * nd + nr > (u64)words);
        if (g == src) { my_r = total_r; my_d = total_d; my_nd = nd; my_nr = nr; }
        total_r += nr;
        total_d += nd;
    }
    bad |= total_r > (u64)rec_capacity || total_d > (u64)desc_capacity;
    const u32 writer = rank == 0 ? 1u : 0u;  // one thread publishes the totals
    if (src == writer && blockIdx.x == 0 && threadIdx.x == 0) {
        if (bad) atomicOr(&plan->error, ERRF_EXCHANGE);
        else { plan->n_pairs = (u32)total_r; plan->n_runs = (u32)total_d; }
    }
    if (bad) return;
    const u64* __restrict__ box = in + size_t(src) * words;
    const bool stash = run_stash(plan->nbits_blocks);
    for (u32 j = blockIdx.x * RUN_THREADS + threadIdx.x; j < my_nd; j += gridDim.x * RUN_THREADS) {
        const u64 key = box[2 + 2 * j], fl = box[3 + 2 * j];
        const u32 first = (u32)fl, len = (u32)(fl >> 32);
        const u32 d = (u32)my_d + j;
        desc_key[d] = key | (stash ? ((u64)len << RUN_STASH_BITS) : 0ull);
        desc_val[d] = d;
        desc[d] = make_uint2((u32)my_r + first, len);
    }
    for (u32 q = blockIdx.x * RUN_THREADS + threadIdx.x; q < my_nr; q += gridDim.x * RUN_THREADS) records[(u32)my_r + q] = box[words - 1 - q];
}

inline unsigned blocks_for(u32 n) { return (n + RUN_THREADS - 1) / RUN_THREADS; }

}  // namespace

static int g_fold_ctas_per_sm = 4;

cudaError_t runs_init() {
    cudaError_t e = cudaFuncSetAttribute(runs_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RUN_THREADS * 33 * sizeof(uint2)));
    if (e != cudaSuccess) return e;
    int per_sm = 0;  // the fold is persistent (blocks are handed out by ticket): launch exactly what is resident
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, runs_fold_kernel, RF_THREADS, 0);
    if (e != cudaSuccess) return e;
    // ... but not more than 4 CTAs (16 warps) per SM: at 64 registers per thread the 7 that fit would take 57 k of the SM's 64 k registers
    // and the next batch's front, which runs concurrently on the main stream, could not place a single CTA beside them
    g_fold_ctas_per_sm = per_sm > 4 ? 4 : (per_sm > 0 ? per_sm : 1);  // (round 2 sweep, profiles/sweep_r02h.md: 2 / 3 / 4 / 5 CTAs -> 9.41 / 9.31 / 9.15 / 9.11 ms per step)
    if (const char* e = std::getenv("CHAD_FOLD_CTAS")) { const int v = std::atoi(e); if (v >= 1 && v <= per_sm) g_fold_ctas_per_sm = v; }
    return cudaSuccess;
}

u32 runs_max_batch_points() { return 1u << RUN_RANK_BITS; }
u32 runs_max_tiles(u32 n_points, u32 n_scans) { return (n_points + RUN_THREADS - 1) / RUN_THREADS + n_scans; }  // every scan may end in a partial tile
size_t runs_max_runs(u32 n_points, u32 n_scans) { return size_t(runs_max_tiles(n_points, n_scans)) * RUN_HASH; }
u32 runs_max_ray_voxels() { return 32; }
u32 runs_max_ray_runs() { return RUN_HASH / RUN_THREADS; }

size_t runs_desc_bytes(size_t capacity) { return capacity * (8 + 8 + 8 + 8 + 8 + 4 + 4) + 1024; }

RunBuffers runs_carve(void* mem, size_t capacity) {
    RunBuffers b;
    unsigned char* p = static_cast<unsigned char*>(mem);
    b.key_a = reinterpret_cast<u64*>(p); p += capacity * 8;
    b.key_b = reinterpret_cast<u64*>(p); p += capacity * 8;
    b.desc = reinterpret_cast<uint2*>(p); p += capacity * 8;
    b.sdesc = reinterpret_cast<uint2*>(p); p += capacity * 8;
    b.work = reinterpret_cast<uint2*>(p); p += capacity * 8;
    b.val_a = reinterpret_cast<u32*>(p); p += capacity * 4;
    b.val_b = reinterpret_cast<u32*>(p);
    b.capacity = (u32)capacity;
    return b;
}

// The ray walk of a batch (main stream). Returns the kernels queued. n_points / n_scans: host upper bounds (a Morton-range shard walks
// only its own rays: their count and the tile table are in device memory). order_rank = (world - 1 - rank) << tsb, 0 on one GPU.
int launch_runs_emit(cudaStream_t s, const float* xyz_sorted, const float* normals, u32 n_points, u32 n_scans, const BatchScans* scans,
                     const MapParams& mp, BatchPlan* plan, const RunBuffers& rb, u64* records, u32 rec_capacity, u32 order_rank,
                     const LaunchHook* hook, int cls_emit) {
    if (!n_points) return 0;
    if (hook) hook->begin(hook->user, cls_emit);
    runs_emit_kernel<<<runs_max_tiles(n_points, n_scans), RUN_THREADS, size_t(RUN_THREADS) * (mp.max_ray_voxels | 1u) * sizeof(uint2), s>>>(
        xyz_sorted, normals, n_points, scans, mp.res, mp.trunc, mp.recip, mp.max_ray_voxels, plan, records, rec_capacity, rb.key_a, rb.val_a, rb.desc,
        rb.capacity, order_rank);
    if (hook) hook->end(hook->user);
    return 1;
}

// Descriptor sort + block list of a batch: only the fold needs them, so they are queued on the fold's stream and the main stream goes
// straight on to the next batch's point stage. `rws` must not be the workspace of the point sort (it runs concurrently).
// max_runs: host upper bound of plan->n_runs.
int launch_runs_group(cudaStream_t s, size_t max_runs, BatchPlan* plan, const RunBuffers& rb, const RadixWorkspace& rws, int num_sms,
                      const LaunchHook* hook, int cls_sort) {
    if (!max_runs) return 0;
    int launches = 0;
    if (hook) hook->begin(hook->user, cls_sort);
    max_runs = std::min<size_t>(rb.capacity, max_runs);
    launches += radix_sort_pairs(s, rb.key_a, rb.val_a, rb.key_b, rb.val_b, &plan->n_runs, &plan->nbits_blocks, max_runs, RS_MAX_PASSES, rws, num_sms);
    runs_group_kernel<<<(unsigned)std::min<size_t>(blocks_for((u32)max_runs), size_t(num_sms) * 8), RUN_THREADS, 0, s>>>(rb.key_a, rb.key_b, rb.val_a, rb.val_b, rb.desc, rb.sdesc, plan, rb.work, rb.capacity);
    if (hook) hook->end(hook->user);
    launches++;
    return launches;
}

// sharded: between the walk and the descriptor sort (see runs_pack_kernel). The box headers are cleared here.
int launch_runs_pack(cudaStream_t s, size_t max_runs, BatchPlan* plan, const RunBuffers& rb, u64* records, const u64* splitters, u32 rank, u32 world,
                     const ShardBoxes& boxes, int num_sms) {
    cudaMemset2DAsync(boxes.out, size_t(boxes.words) * 8, 0, 16, world, s);
    max_runs = std::min<size_t>(rb.capacity, max_runs);
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(blocks_for((u32)max_runs), size_t(num_sms) * 8));
    runs_pack_kernel<<<grid, RUN_THREADS, 0, s>>>(records, rb.key_a, rb.desc, plan, splitters, rank, world, boxes.out, boxes.words);
    return 1;
}

int launch_runs_ingest(cudaStream_t s, BatchPlan* plan, const RunBuffers& rb, u64* records, u32 rec_capacity, u32 rank, u32 world, const ShardBoxes& boxes) {
    runs_ingest_kernel<<<dim3(16, world), RUN_THREADS, 0, s>>>(records, rec_capacity, rb.key_a, rb.val_a, rb.desc, rb.capacity, plan, rank, world, boxes.in,
                                                                boxes.words);
    return 1;
}

int launch_runs_fold(cudaStream_t s, const u64* records, const RunBuffers& rb, BatchPlan* plan, const ChunkTable& t, int num_sms) {
    runs_fold_kernel<<<num_sms * g_fold_ctas_per_sm, RF_THREADS, 0, s>>>(records, rb.key_a, rb.key_b, rb.sdesc, rb.work, rb.capacity, plan, t.keys, t.cells,
                                                                              t.capacity, t.count, t.list);
    return 1;
}

}  // namespace chadgpu
