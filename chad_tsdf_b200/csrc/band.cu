// Truncation-band voxel enumeration along every sensor ray and the signed distance of each band
// voxel: the loop body of Octree::insert(points, normals, position, res, trunc),
// /root/reference/include/chad/detail/octree.hpp:86-159 (Amanatides-Woo style DDA with the
// reference's exact operation order, strict-< comparisons and x / z / y / z axis preference).
// The reference walks the octree for every emitted voxel; here the voxels become a stream of
// (compact Morton key, sd) pairs in (sorted point, ray step) order -- the order the running
// average of octree.hpp:161-163 depends on -- that the stable radix sort groups per voxel.
#include "kernels.cuh"
#include "ray.cuh"

namespace chadgpu {

namespace {

constexpr int BAND_THREADS = 256;

__global__ void __launch_bounds__(BAND_THREADS) band_count_kernel(const float* __restrict__ xyz_sorted, u32 n_points,
                                                                  const BatchScans* __restrict__ scans, float res, float trunc, float recip,
                                                                  u32 max_ray_voxels, BatchPlan* __restrict__ plan, u32* __restrict__ counts) {
    const u32 i = blockIdx.x * BAND_THREADS + threadIdx.x;
    if (i >= n_points) return;
    const u32 s = scan_of(scans, plan->n_scans, i);
    const float pos[3] = {scans->pose[s][0], scans->pose[s][1], scans->pose[s][2]};
    Ray r;
    ray_setup(r, xyz_sorted[size_t(i) * 3], xyz_sorted[size_t(i) * 3 + 1], xyz_sorted[size_t(i) * 3 + 2], pos, res, trunc, recip);
    u32 c = 1;  // the start voxel is always emitted (:121-122)
    while (ray_advance(r)) {
        if (c >= max_ray_voxels) { atomicOr(&plan->error, ERRF_PAIR_CAPACITY); break; }  // the analytic per-ray bound was exceeded: report it
        c++;
    }
    counts[i] = c;
}

template <bool FULL_KEYS>
__global__ void __launch_bounds__(BAND_THREADS) band_emit_kernel(const float* __restrict__ xyz_sorted, const float* __restrict__ normals,
                                                                 u32 n_points, const BatchScans* __restrict__ scans, float res, float trunc,
                                                                 float recip, u32 max_ray_voxels, BatchPlan* plan,
                                                                 const u32* __restrict__ offsets, u64* __restrict__ pair_keys,
                                                                 u32* __restrict__ pair_sd, u32 pair_capacity) {
    const u32 i = blockIdx.x * BAND_THREADS + threadIdx.x;
    if (i >= n_points) return;
    const u32 s = scan_of(scans, plan->n_scans, i);
    const float pos[3] = {scans->pose[s][0], scans->pose[s][1], scans->pose[s][2]};
    const u32 k = plan->k;
    const float nx = normals[size_t(i) * 3], ny = normals[size_t(i) * 3 + 1], nz = normals[size_t(i) * 3 + 2];
    Ray r;
    ray_setup(r, xyz_sorted[size_t(i) * 3], xyz_sorted[size_t(i) * 3 + 1], xyz_sorted[size_t(i) * 3 + 2], pos, res, trunc, recip);
    u32 out = offsets[i];
    u32 c = 0;
    u32 err = 0;
    u64 full = morton_encode(r.cur[0], r.cur[1], r.cur[2]);
    while (true) {
        if (out >= pair_capacity) { err |= ERRF_PAIR_CAPACITY; break; }
        const i32 vx = r.cur[0], vy = r.cur[1], vz = r.cur[2];
        // inside the plan's range (hence inside the 21-bit Morton range) the stored code decodes back to (vx, vy, vz)
        // exactly, so the reference's decode (octree.hpp:157) is the identity and need not be recomputed
        if (max(rcode(vx), max(rcode(vy), rcode(vz))) >= (1u << k)) err |= ERRF_RANGE;
        // octree.hpp:157-159: the voxel's LOWER CORNER projected on the normal, clamped to +-trunc
        float sd = dot3(nx, ny, nz, fsub(fmul((float)vx, res), r.px), fsub(fmul((float)vy, res), r.py), fsub(fmul((float)vz, res), r.pz));
        sd = fclamp(sd, -trunc, trunc);
        pair_keys[out] = FULL_KEYS ? full : compact_key(full, k);
        pair_sd[out] = __float_as_uint(sd);
        out++;
        c++;
        if (c >= max_ray_voxels) break;
        int axis;
        const bool more = ray_advance(r, axis);
        if (!more) break;
        full = morton_step(full, axis, axis == 0 ? r.step[0] : (axis == 1 ? r.step[1] : r.step[2]));
    }
    if (err) atomicOr(&plan->error, err);
}

inline unsigned blocks_for(u32 n) { return (n + BAND_THREADS - 1) / BAND_THREADS; }

}  // namespace

int launch_band_count(cudaStream_t s, const float* xyz_sorted, u32 n_points, const BatchScans* scans, const MapParams& mp,
                      const BatchPlan* plan, u32* counts) {
    if (!n_points) return 0;
    band_count_kernel<<<blocks_for(n_points), BAND_THREADS, 0, s>>>(xyz_sorted, n_points, scans, mp.res, mp.trunc, mp.recip, mp.max_ray_voxels,
                                                                    const_cast<BatchPlan*>(plan), counts);
    return 1;
}

int launch_band_emit(cudaStream_t s, const float* xyz_sorted, const float* normals, u32 n_points, const BatchScans* scans,
                     const MapParams& mp, BatchPlan* plan, const u32* offsets, u64* pair_keys, u32* pair_sd, u32 pair_capacity,
                     bool full_keys) {
    if (!n_points) return 0;
    if (full_keys)
        band_emit_kernel<true><<<blocks_for(n_points), BAND_THREADS, 0, s>>>(xyz_sorted, normals, n_points, scans, mp.res, mp.trunc, mp.recip,
                                                                             mp.max_ray_voxels, plan, offsets, pair_keys, pair_sd, pair_capacity);
    else
        band_emit_kernel<false><<<blocks_for(n_points), BAND_THREADS, 0, s>>>(xyz_sorted, normals, n_points, scans, mp.res, mp.trunc, mp.recip,
                                                                              mp.max_ray_voxels, plan, offsets, pair_keys, pair_sd, pair_capacity);
    return 1;
}

}  // namespace chadgpu
