// The reference's ray set-up and DDA step (/root/reference/include/chad/detail/octree.hpp:92-152), shared by the
// band kernels (band.cu: global-sort path) and the block-binning kernels (blocks.cu).
#pragma once
#include "common.cuh"

namespace chadgpu {

struct Ray {
    float px, py, pz;
    i32 cur[3], vf[3], step[3];
    float tmax[3], delta[3];
};

// octree.hpp:92-118
__device__ __forceinline__ void ray_setup(Ray& r, float px, float py, float pz, const float* pos, float res, float trunc, float recip) {
    r.px = px; r.py = py; r.pz = pz;
    const float p[3] = {px, py, pz};
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; a++) d[a] = fsub(p[a], pos[a]);
    const float invl = fdiv(1.0f, fsqrt(dot3(d[0], d[1], d[2], d[0], d[1], d[2])));  // normalize, :92
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float dir = fmul(d[a], invl);
        const float dir_recip = fdiv(1.0f, dir);                  // :93
        const float start = fsub(p[a], fmul(dir, trunc));         // :94
        const float fin = fadd(p[a], fmul(dir, trunc));           // :95
        const float sv = fmul(start, recip);
        const i32 vs = (i32)floorf(sv);                           // :96
        const i32 vf = (i32)floorf(fmul(fin, recip));             // :97
        const i32 dv = vf - vs;
        const i32 st = (0 < dv) - (dv < 0);                       // :100
        r.delta[a] = fabsf(fmul(res, dir_recip));                 // :102
        float m;                                                  // :104-116
        if (st < 0) m = fmul(res, floorf(sv));
        else if (st > 0) m = fmul(res, ceilf(sv));
        else m = 3.402823466e+38f;
        m = fsub(m, start);                                       // :117
        r.tmax[a] = fabsf(fmul(m, dir_recip));                    // :118
        r.cur[a] = vs; r.vf[a] = vf; r.step[a] = st;
    }
}
// one iteration of the while(true) loop of octree.hpp:125-152; returns false on `break`; `axis` = the axis stepped
__device__ __forceinline__ bool ray_advance(Ray& r, int& axis) {
    int a;
    if (r.tmax[0] < r.tmax[1]) a = (r.tmax[0] < r.tmax[2]) ? 0 : 2;
    else a = (r.tmax[1] < r.tmax[2]) ? 1 : 2;
    axis = a;
    // select without dynamic register indexing
    if (a == 0) { r.cur[0] += r.step[0]; r.tmax[0] = fadd(r.tmax[0], r.delta[0]); return r.cur[0] != r.vf[0] + r.step[0]; }
    if (a == 1) { r.cur[1] += r.step[1]; r.tmax[1] = fadd(r.tmax[1], r.delta[1]); return r.cur[1] != r.vf[1] + r.step[1]; }
    r.cur[2] += r.step[2]; r.tmax[2] = fadd(r.tmax[2], r.delta[2]); return r.cur[2] != r.vf[2] + r.step[2];
}
__device__ __forceinline__ bool ray_advance(Ray& r) { int a; return ray_advance(r, a); }
// Morton key of the neighbour one voxel along `axis` (dir = +1 / -1 / 0): add or subtract 1 inside the axis' bit lane
__device__ __forceinline__ u64 morton_step(u64 key, int axis, i32 dir) {
    const u64 lane = 0x1249249249249249ull << axis;
    if (dir > 0) return (((key | ~lane) + 1ull) & lane) | (key & ~lane);
    if (dir < 0) return (((key & lane) - 1ull) & lane) | (key & ~lane);
    return key;
}

}  // namespace chadgpu
