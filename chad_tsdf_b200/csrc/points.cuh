// Device helpers shared by the point-stage kernels (points.cu) and the Morton-range sharding kernels (shard.cu):
// coalesced staging of AoS points, voxelisation (/root/reference/include/chad/detail/morton.hpp:71-73), the per-batch
// plan arithmetic and the point sort key.
#pragma once
#include "common.cuh"

namespace chadgpu {

constexpr int PT_THREADS = 256;

// Coalesced float4 staging of a tile of AoS xyz points (12 B each) into shared memory.
// `xyz` must be 16-byte aligned; tile_base (in points) must be a multiple of 4.
__device__ __forceinline__ void load_xyz_tile(const float* __restrict__ xyz, u32 tile_base, u32 n_points, float* s_xyz) {
    const u32 pts = min((u32)PT_THREADS, n_points - tile_base);
    const u32 nf = pts * 3;
    const float* src = xyz + size_t(tile_base) * 3;
    const u32 nvec = nf >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    float4* dst4 = reinterpret_cast<float4*>(s_xyz);
    for (u32 v = threadIdx.x; v < nvec; v += PT_THREADS) dst4[v] = __ldg(&src4[v]);
    for (u32 f = (nvec << 2) + threadIdx.x; f < nf; f += PT_THREADS) s_xyz[f] = __ldg(&src[f]);
    __syncthreads();
}

// morton.hpp:71-73: v = ivec3(floor(p * recip)), fp32
__device__ __forceinline__ bool voxel_of(float px, float py, float pz, float recip, i32& vx, i32& vy, i32& vz) {
    const float fx = floorf(fmul(px, recip)), fy = floorf(fmul(py, recip)), fz = floorf(fmul(pz, recip));
    const float lim = 1048576.0f;  // 2^20
    const bool ok = (fabsf(fx) < lim) && (fabsf(fy) < lim) && (fabsf(fz) < lim);  // false for NaN / Inf too
    vx = ok ? (i32)fx : 0;
    vy = ok ? (i32)fy : 0;
    vz = ok ? (i32)fz : 0;
    return ok;
}

__device__ __forceinline__ u32 bits_for(u32 count) { return (count > 1) ? (32 - __clz(count - 1)) : 0; }  // bits that hold 0 .. count - 1

// Everything of the plan that follows from rmax, the number of points THIS context sorts (plan->n_points) and the layout of the
// order key of a run descriptor (tile-run pair path): scan | Morton-range rank (descending) | 256-ray tile inside the scan.
//   tsb   = bits of a tile index inside one scan (host: from the largest scan of the whole batch -- identical on every rank)
//   gbits = bits of the Morton-range rank (0 on a single GPU)
__device__ __forceinline__ void plan_finalize_body(BatchPlan* plan, u32 margin, u32 tsb, u32 gbits) {
    const u32 reach = plan->rmax + margin;  // every band voxel of the batch has range code <= reach
    u32 k = 32 - __clz(reach);              // smallest k with reach < 2^k
    if (k < 3) k = 3;                       // keep at least the 4^3 neighbourhood bits below the top triple
    if (k > 20) { k = 20; atomicOr(&plan->error, ERRF_RANGE); }
    const u32 sbits = bits_for(plan->n_scans);
    u32 nb = 3 * k + 3 + sbits;
    if (nb > 64) { nb = 64; atomicOr(&plan->error, ERRF_KEY_BUDGET); }
    plan->k = k;
    plan->nbits_pairs = 3 * k + 3;
    {   // run descriptors of the tile-run pair path sort by (compact 8^3-block id, order key)
        const u32 tbits = sbits + gbits + tsb;
        plan->tile_bits = tbits;
        plan->tsb = tsb;
        plan->nbits_blocks = 3 * k - 6 + tbits;
        if (3 * k - 6 + tbits > 64) plan->nbits_blocks = 0xFFFFFFFFu;  // the tile-run path reports ERRF_KEY_BUDGET; paths 0 / 1 do not use it
    }
    plan->nbits_points = nb;
    // the input index (the sort's payload) fits under the key: one 8-byte array goes through the sort instead of 8 + 4 bytes
    plan->point_shift = (nb + POINT_INDEX_BITS <= 64 && plan->n_batch <= (1u << POINT_INDEX_BITS)) ? POINT_INDEX_BITS : 0u;
}

// sort key of a point: (scan << (3k+3)) | (~compact(morton) & mask): ascending sort == per scan descending Morton
// (morton.hpp:85-89); the stable LSD sort breaks ties by input index (canonical tie-break)
__device__ __forceinline__ u64 point_sort_key(u64 full, u32 k, u32 scan) {
    const u32 cbits = 3 * k + 3;
    const u64 cmask = (cbits >= 64) ? ~0ull : ((1ull << cbits) - 1ull);
    const u64 inv = ~compact_key(full, k) & cmask;
    return (cbits >= 64 ? 0ull : (u64(scan) << cbits)) | inv;
}

}  // namespace chadgpu
