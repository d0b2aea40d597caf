// Point stage of TSDFMap::insert on the device:
//   * voxelise + Morton encode      -- include/chad/detail/morton.hpp:21-28,59-80
//   * sort keys (descending Morton per scan, ties by input index; the radix sort itself is in
//     radix_sort.cu)                -- include/chad/detail/morton.hpp:81-102
//   * greedy Morton neighbourhoods + FP64 plane fit + sensor-facing flip
//                                   -- include/chad/detail/normals.hpp:10-148
// All citations are into /root/reference.
#include "kernels.cuh"
#include "points.cuh"
#include "radix_sort.cuh"

namespace chadgpu {

namespace {

__global__ void __launch_bounds__(PT_THREADS) plan_reset_kernel(BatchPlan* plan, u32 n_points, u32 n_scans) {
    plan->rmax = 0; plan->k = 0; plan->nbits_points = 0; plan->nbits_pairs = 0;
    plan->n_points = n_points; plan->n_scans = n_scans; plan->n_pairs = 0;
    plan->n_segments = 0; plan->n_chunk_heads = 0; plan->n_new_chunks = 0; plan->fold_ticket = 0; plan->n_blocks = 0; plan->sort_ticket = 0; plan->n_runs = 0; plan->nbits_blocks = 0; plan->tile_bits = 0; plan->n_big_blocks = 0; plan->n_small_blocks = 0; plan->point_shift = 0;
    plan->tsb = 0; plan->n_batch = n_points; plan->tail_lo = 0xFFFFFFFFu; plan->tail_hi = 0xFFFFFFFFu;
    plan->n_runs_local = 0; plan->n_pairs_local = 0; plan->xfer_runs = 0; plan->xfer_records = 0;
    // plan->error is sticky: cleared by the host when it reports it
}

__global__ void __launch_bounds__(PT_THREADS) plan_bbox_kernel(const float* __restrict__ xyz, u32 n_points, float recip, BatchPlan* plan) {
    __shared__ __align__(16) float s_xyz[PT_THREADS * 3];
    const u32 tile_base = blockIdx.x * PT_THREADS;
    load_xyz_tile(xyz, tile_base, n_points, s_xyz);
    const u32 i = tile_base + threadIdx.x;
    u32 r = 0;
    u32 err = 0;
    if (i < n_points) {
        const float px = s_xyz[threadIdx.x * 3], py = s_xyz[threadIdx.x * 3 + 1], pz = s_xyz[threadIdx.x * 3 + 2];
        i32 vx, vy, vz;
        if (!voxel_of(px, py, pz, recip, vx, vy, vz)) err = (isfinite(px) && isfinite(py) && isfinite(pz)) ? ERRF_RANGE : ERRF_NUMERIC;
        r = max(rcode(vx), max(rcode(vy), rcode(vz)));
    }
    r = __reduce_max_sync(0xffffffffu, r);
    err = __reduce_or_sync(0xffffffffu, err);
    // one global atomic per block, and only when it would raise the maximum (the first blocks settle it)
    __shared__ u32 s_r[PT_THREADS / 32], s_e[PT_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { s_r[threadIdx.x >> 5] = r; s_e[threadIdx.x >> 5] = err; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 br = 0, be = 0;
#pragma unroll
        for (int w = 0; w < PT_THREADS / 32; w++) { br = max(br, s_r[w]); be |= s_e[w]; }
        if (br > *(volatile u32*)&plan->rmax) atomicMax(&plan->rmax, br);
        if (be) atomicOr(&plan->error, be);
    }
}

__global__ void plan_finalize_kernel(BatchPlan* plan, u32 margin, u32 tsb) { plan_finalize_body(plan, margin, tsb, 0u); }

// sort key of a point: (scan << (3k+3)) | (~compact(morton) & mask): ascending sort == per scan
// descending Morton (morton.hpp:85-89); the stable LSD sort breaks ties by input index (canonical).
__global__ void __launch_bounds__(PT_THREADS) point_keys_kernel(const float* __restrict__ xyz, u32 n_points, const BatchScans* __restrict__ scans,
                                                                float recip, const BatchPlan* __restrict__ plan, u64* __restrict__ sortkeys,
                                                                u32* __restrict__ index) {
    __shared__ __align__(16) float s_xyz[PT_THREADS * 3];
    const u32 tile_base = blockIdx.x * PT_THREADS;
    load_xyz_tile(xyz, tile_base, n_points, s_xyz);
    const u32 i = tile_base + threadIdx.x;
    if (i >= n_points) return;
    i32 vx, vy, vz;
    voxel_of(s_xyz[threadIdx.x * 3], s_xyz[threadIdx.x * 3 + 1], s_xyz[threadIdx.x * 3 + 2], recip, vx, vy, vz);
    const u64 sk = point_sort_key(morton_encode(vx, vy, vz), plan->k, scan_of(scans, plan->n_scans, i));
    if (plan->point_shift) sortkeys[i] = (sk << POINT_INDEX_BITS) | (u64)i;
    else { sortkeys[i] = sk; index[i] = i; }
}

__global__ void __launch_bounds__(PT_THREADS) point_gather_kernel(const float* __restrict__ xyz, u32 n_points, const BatchPlan* __restrict__ plan,
                                                                  const u64* __restrict__ keys_a, const u64* __restrict__ keys_b,
                                                                  const u32* __restrict__ idx_a, const u32* __restrict__ idx_b,
                                                                  u64* __restrict__ sorted_keys, u32* __restrict__ sorted_order,
                                                                  float* __restrict__ xyz_sorted) {
    const u32 i = blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n_points || i >= plan->n_points) return;  // (a Morton-range shard sorts only its own points of the batch)
    const bool alt = radix_result_in_alt(plan->nbits_points);
    u64 key = alt ? keys_b[i] : keys_a[i];
    u32 src;
    if (plan->point_shift) { src = (u32)key & ((1u << POINT_INDEX_BITS) - 1u); key >>= POINT_INDEX_BITS; }
    else src = alt ? idx_b[i] : idx_a[i];
    sorted_keys[i] = key;
    sorted_order[i] = src;
    const float x = __ldg(&xyz[size_t(src) * 3]), y = __ldg(&xyz[size_t(src) * 3 + 1]), z = __ldg(&xyz[size_t(src) * 3 + 2]);
    xyz_sorted[size_t(i) * 3] = x;
    xyz_sorted[size_t(i) * 3 + 1] = y;
    xyz_sorted[size_t(i) * 3 + 2] = z;
}

// Greedy neighbourhood segmentation (normals.hpp:86-108,137). A neighbourhood never crosses a
// 4^3-voxel block (the depth-2 mask) nor a scan, so every (scan, key >> 6) block is segmented
// independently by the thread sitting on its first point. seg_info[j] = neighbourhood size on the
// first point of a neighbourhood with >= 8 points, 0 on its other members, 1 on every point of a
// smaller neighbourhood (those take the per-point fallback normal, normals.hpp:127-134).
__global__ void __launch_bounds__(PT_THREADS) segment_kernel(const u64* __restrict__ sorted_keys, u32 n_points, const BatchScans* __restrict__ scans,
                                                             const BatchPlan* __restrict__ plan, u32* __restrict__ seg_info) {
    const u32 i = blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n_points || i >= plan->n_points) return;
    const u64 block_id = sorted_keys[i] >> 6;  // includes the scan bits
    if (i > 0 && (sorted_keys[i - 1] >> 6) == block_id) return;
    const u32 s = scan_of(scans, plan->n_scans, i);
    const u32 scan_end = scans->offset[s + 1];
    // normals.hpp:100: the scan's last point is never absorbed. Of a sharded map only the rank that holds
    // the scan's lowest key has that point (plan->tail_*); on the others no point is special
    const bool tail = ((s < 32 ? plan->tail_lo >> s : plan->tail_hi >> (s - 32)) & 1u) != 0;
    const u32 last = tail ? scan_end - 1 : 0xFFFFFFFFu;
    u32 it = i;
    while (it < scan_end && (sorted_keys[it] >> 6) == block_id) {
        const u64 key_it = sorted_keys[it];
        u32 end = it + 1;
#pragma unroll 1
        for (u32 depth = 0; depth < 3; depth++) {
            const u32 sh = depth * 3;
            while (end != last && end < scan_end && (sorted_keys[end] >> sh) == (key_it >> sh)) end++;
            if (end - it >= 8) break;
        }
        const u32 size = end - it;
        if (size >= 8) {
            seg_info[it] = size;
            for (u32 j = it + 1; j < end; j++) seg_info[j] = 0;
        } else {
            for (u32 j = it; j < end; j++) seg_info[j] = 1;
        }
        it = end;
    }
}

// normals.hpp:10-80 (plane fit, FP64, sequential accumulation order), :117-118 (flip), :127-134 (fallback)
__global__ void __launch_bounds__(PT_THREADS) normals_kernel(const float* __restrict__ xyz_sorted, u32 n_points, const BatchScans* __restrict__ scans,
                                                             const BatchPlan* __restrict__ plan, const u32* __restrict__ seg_info,
                                                             float* __restrict__ normals) {
    const u32 i = blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n_points || i >= plan->n_points) return;
    const u32 info = seg_info[i];
    if (info == 0) return;  // member of a neighbourhood whose first point writes the shared normal
    const u32 s = scan_of(scans, plan->n_scans, i);
    const float posx = scans->pose[s][0], posy = scans->pose[s][1], posz = scans->pose[s][2];
    const float px = xyz_sorted[size_t(i) * 3], py = xyz_sorted[size_t(i) * 3 + 1], pz = xyz_sorted[size_t(i) * 3 + 2];
    // normalize(position - point): v * (1 / sqrt(dot(v, v)))
    const float dx = fsub(posx, px), dy = fsub(posy, py), dz = fsub(posz, pz);
    const float invl = fdiv(1.0f, fsqrt(dot3(dx, dy, dz, dx, dy, dz)));
    const float tx = fmul(dx, invl), ty = fmul(dy, invl), tz = fmul(dz, invl);
    if (info < 8) {
        normals[size_t(i) * 3] = tx;
        normals[size_t(i) * 3 + 1] = ty;
        normals[size_t(i) * 3 + 2] = tz;
        return;
    }
    const u32 end = i + info;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (u32 j = i; j < end; j++) {
        cx = dadd(cx, (double)xyz_sorted[size_t(j) * 3]);
        cy = dadd(cy, (double)xyz_sorted[size_t(j) * 3 + 1]);
        cz = dadd(cz, (double)xyz_sorted[size_t(j) * 3 + 2]);
    }
    const double recip = ddiv(1.0, (double)info);
    cx = dmul(cx, recip); cy = dmul(cy, recip); cz = dmul(cz, recip);
    double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
    for (u32 j = i; j < end; j++) {
        const double rx = dsub((double)xyz_sorted[size_t(j) * 3], cx);
        const double ry = dsub((double)xyz_sorted[size_t(j) * 3 + 1], cy);
        const double rz = dsub((double)xyz_sorted[size_t(j) * 3 + 2], cz);
        xx = dadd(xx, dmul(rx, rx)); xy = dadd(xy, dmul(rx, ry)); xz = dadd(xz, dmul(rx, rz));
        yy = dadd(yy, dmul(ry, ry)); yz = dadd(yz, dmul(ry, rz)); zz = dadd(zz, dmul(rz, rz));
    }
    xx = dmul(xx, recip); xy = dmul(xy, recip); xz = dmul(xz, recip);
    yy = dmul(yy, recip); yz = dmul(yz, recip); zz = dmul(zz, recip);
    double wx = 0.0, wy = 0.0, wz = 0.0;
    {   // determinant x (normals.hpp:42-52)
        const double det = dsub(dmul(yy, zz), dmul(yz, yz));
        const double ax = det, ay = dsub(dmul(xz, yz), dmul(xy, zz)), az = dsub(dmul(xy, yz), dmul(xz, yy));
        double w = dmul(det, det);
        if (ddot3(wx, wy, wz, ax, ay, az) < 0.0) w = -w;
        wx = dadd(wx, dmul(ax, w)); wy = dadd(wy, dmul(ay, w)); wz = dadd(wz, dmul(az, w));
    }
    {   // determinant y (normals.hpp:54-64)
        const double det = dsub(dmul(xx, zz), dmul(xz, xz));
        const double ax = dsub(dmul(xz, yz), dmul(xy, zz)), ay = det, az = dsub(dmul(xy, xz), dmul(yz, xx));
        double w = dmul(det, det);
        if (ddot3(wx, wy, wz, ax, ay, az) < 0.0) w = -w;
        wx = dadd(wx, dmul(ax, w)); wy = dadd(wy, dmul(ay, w)); wz = dadd(wz, dmul(az, w));
    }
    {   // determinant z (normals.hpp:66-76)
        const double det = dsub(dmul(xx, yy), dmul(xy, xy));
        const double ax = dsub(dmul(xy, yz), dmul(xz, yy)), ay = dsub(dmul(xy, xz), dmul(yz, xx)), az = det;
        double w = dmul(det, det);
        if (ddot3(wx, wy, wz, ax, ay, az) < 0.0) w = -w;
        wx = dadd(wx, dmul(ax, w)); wy = dadd(wy, dmul(ay, w)); wz = dadd(wz, dmul(az, w));
    }
    const double inv = ddiv(1.0, dsqrt(ddot3(wx, wy, wz, wx, wy, wz)));
    float nx = (float)dmul(wx, inv), ny = (float)dmul(wy, inv), nz = (float)dmul(wz, inv);
    if (dot3(nx, ny, nz, tx, ty, tz) < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }  // normals.hpp:117-118
    for (u32 j = i; j < end; j++) {
        normals[size_t(j) * 3] = nx;
        normals[size_t(j) * 3 + 1] = ny;
        normals[size_t(j) * 3 + 2] = nz;
    }
}

__global__ void __launch_bounds__(PT_THREADS) point_full_keys_kernel(const u64* __restrict__ sorted_keys, u32 n_points,
                                                                     const BatchPlan* __restrict__ plan, u64* __restrict__ full_keys) {
    const u32 i = blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n_points || i >= plan->n_points) return;
    const u32 k = plan->k;
    const u32 cbits = 3 * k + 3;
    const u64 cmask = (cbits >= 64) ? ~0ull : ((1ull << cbits) - 1ull);
    full_keys[i] = expand_key(~sorted_keys[i] & cmask, k);
}

__global__ void __launch_bounds__(PT_THREADS) morton_encode_kernel(const i32* __restrict__ voxels, u32 n, u64* __restrict__ keys) {
    const u32 i = blockIdx.x * PT_THREADS + threadIdx.x;
    if (i >= n) return;
    keys[i] = morton_encode(voxels[size_t(i) * 3], voxels[size_t(i) * 3 + 1], voxels[size_t(i) * 3 + 2]);
}

inline unsigned blocks_for(u32 n) { return (n + PT_THREADS - 1) / PT_THREADS; }

}  // namespace

int launch_plan_reset(cudaStream_t s, BatchPlan* plan, u32 n_points, u32 n_scans) {
    plan_reset_kernel<<<1, 1, 0, s>>>(plan, n_points, n_scans);
    return 1;
}

int launch_plan(cudaStream_t s, const float* xyz, u32 n_points, u32 n_scans, const MapParams& mp, BatchPlan* plan, u32 tsb) {
    plan_reset_kernel<<<1, 1, 0, s>>>(plan, n_points, n_scans);
    int launches = 1;
    if (n_points) {
        plan_bbox_kernel<<<blocks_for(n_points), PT_THREADS, 0, s>>>(xyz, n_points, mp.recip, plan);
        launches++;
    }
    plan_finalize_kernel<<<1, 1, 0, s>>>(plan, mp.band_margin, tsb);
    return launches + 1;
}

int launch_point_keys(cudaStream_t s, const float* xyz, u32 n_points, const BatchScans* scans, const MapParams& mp, const BatchPlan* plan,
                      u64* sortkeys, u32* index) {
    if (!n_points) return 0;
    point_keys_kernel<<<blocks_for(n_points), PT_THREADS, 0, s>>>(xyz, n_points, scans, mp.recip, plan, sortkeys, index);
    return 1;
}

int launch_point_gather(cudaStream_t s, const float* xyz, u32 n_points, const BatchPlan* plan, const u64* keys_a, const u64* keys_b,
                        const u32* idx_a, const u32* idx_b, u64* sorted_keys, u32* sorted_order, float* xyz_sorted) {
    if (!n_points) return 0;
    point_gather_kernel<<<blocks_for(n_points), PT_THREADS, 0, s>>>(xyz, n_points, plan, keys_a, keys_b, idx_a, idx_b, sorted_keys, sorted_order,
                                                                    xyz_sorted);
    return 1;
}

int launch_normals(cudaStream_t s, const float* xyz_sorted, const u64* sorted_keys, u32 n_points, const BatchScans* scans,
                   const BatchPlan* plan, u32* seg_info, float* normals) {
    if (!n_points) return 0;
    segment_kernel<<<blocks_for(n_points), PT_THREADS, 0, s>>>(sorted_keys, n_points, scans, plan, seg_info);
    normals_kernel<<<blocks_for(n_points), PT_THREADS, 0, s>>>(xyz_sorted, n_points, scans, plan, seg_info, normals);
    return 2;
}

int launch_point_full_keys(cudaStream_t s, const u64* sorted_keys, u32 n_points, const BatchPlan* plan, u64* full_keys) {
    if (!n_points) return 0;
    point_full_keys_kernel<<<blocks_for(n_points), PT_THREADS, 0, s>>>(sorted_keys, n_points, plan, full_keys);
    return 1;
}

int launch_morton_encode(cudaStream_t s, const i32* voxels, u32 n, u64* keys) {
    if (!n) return 0;
    morton_encode_kernel<<<blocks_for(n), PT_THREADS, 0, s>>>(voxels, n, keys);
    return 1;
}

}  // namespace chadgpu
