/* TEST INFRASTRUCTURE ONLY (oracle/). Never linked into, imported by, or called from the
 * product path (chad_tsdf_b200/, include/); used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the CHECKER.
 *
 * CPU restatement, in plain C, of the reference's integration hot path
 *   chad::TSDFMap::insert            /root/reference/src/chad/tsdf.cpp:39-75
 *   chad::detail::Submap::finalize   /root/reference/include/chad/detail/submap.hpp:10-106
 * in the DATA-PARALLEL formulation the CUDA path uses (sort by key + ordered segmented fold +
 * level-by-level first-occurrence dedup) rather than the reference's pointer octree + DFS. Each
 * function cites the reference lines it restates. Arithmetic is the canonical semantics of
 * strict IEEE-754 binary32/binary64 (build with -ffp-contract=off, no
 * -ffast-math), glm's generic scalar definitions (dot = (x+y)+z, normalize = v * (1/sqrt(dot))),
 * equal-key ties in the point sort broken by input index.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py checks this file bit-for-bit (voxels,
 * float bits, every DAG level word, counters, roots) against the reference's own sources compiled
 * in oracle/_ref (stable-sort variant), and tests/golden/ holds hashes generated from that build. */
#include "chad_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Morton code: morton.hpp:21-37 (libmorton BMI2 pdep/pext restated as magic-bit interleave)     */
/* ------------------------------------------------------------------------------------------ */
static uint64_t spread3(uint64_t v) { /* 21 bits -> every third bit */
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
static uint32_t compact3(uint64_t v) {
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
    v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
    v = (v | (v >> 8)) & 0x001F0000FF0000FFull;
    v = (v | (v >> 16)) & 0x001F00000000FFFFull;
    v = (v | (v >> 32)) & 0x1FFFFFull;
    return (uint32_t)v;
}
/* morton.hpp:21-28: bias 2^20 then interleave x -> bit 0, y -> bit 1, z -> bit 2. Valid for
 * |voxel| < 2^20; the 22nd x bit libmorton's mask would keep is 0 there. */
uint64_t oracle_morton_encode(int32_t x, int32_t y, int32_t z) {
    uint32_t ux = (1u << 20) + (uint32_t)x, uy = (1u << 20) + (uint32_t)y, uz = (1u << 20) + (uint32_t)z;
    return spread3(ux) | (spread3(uy) << 1) | (spread3(uz) << 2);
}
/* morton.hpp:29-37 */
void oracle_morton_decode(uint64_t key, int32_t* x, int32_t* y, int32_t* z) {
    *x = (int32_t)(compact3(key) - (1u << 20));
    *y = (int32_t)(compact3(key >> 1) - (1u << 20));
    *z = (int32_t)(compact3(key >> 2) - (1u << 20));
}

/* ------------------------------------------------------------------------------------------ */
/* helpers                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static void* xmalloc(size_t n) {
    void* p = malloc(n ? n : 1);
    if (!p) abort();
    return p;
}
static void* xrealloc(void* p, size_t n) {
    p = realloc(p, n ? n : 1);
    if (!p) abort();
    return p;
}
/* stable LSD radix sort of (key, payload) by ascending key */
static void radix_sort_u64(uint64_t* keys, uint32_t* pay, size_t n) {
    uint64_t* k2 = xmalloc(n * 8);
    uint32_t* p2 = xmalloc(n * 4);
    for (int pass = 0; pass < 8; pass++) {
        size_t hist[257] = {0};
        int shift = pass * 8;
        for (size_t i = 0; i < n; i++) hist[((keys[i] >> shift) & 0xFF) + 1]++;
        if (hist[((keys[0] >> shift) & 0xFF) + 1] == n) continue; /* digit constant */
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (size_t i = 0; i < n; i++) {
            size_t dst = hist[(keys[i] >> shift) & 0xFF]++;
            k2[dst] = keys[i];
            p2[dst] = pay[i];
        }
        memcpy(keys, k2, n * 8);
        memcpy(pay, p2, n * 4);
    }
    free(k2);
    free(p2);
}
static float vdot3(float ax, float ay, float az, float bx, float by, float bz) {
    float tx = ax * bx, ty = ay * by, tz = az * bz; /* glm compute_dot<vec3>: (x + y) + z */
    return (tx + ty) + tz;
}

/* ------------------------------------------------------------------------------------------ */
/* point stage                                                                                  */
/* ------------------------------------------------------------------------------------------ */
/* normals.hpp:10-80: plane normal of points [beg, end) of the sorted list, double precision,
 * sequential accumulation order. */
static void estimate_normal(const float* xyz, size_t beg, size_t end, float out[3]) {
    double cx = 0, cy = 0, cz = 0;
    for (size_t i = beg; i < end; i++) {
        cx += (double)xyz[3 * i];
        cy += (double)xyz[3 * i + 1];
        cz += (double)xyz[3 * i + 2];
    }
    double recip = 1.0 / (double)(end - beg);
    cx *= recip; cy *= recip; cz *= recip;
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    for (size_t i = beg; i < end; i++) {
        double rx = (double)xyz[3 * i] - cx, ry = (double)xyz[3 * i + 1] - cy, rz = (double)xyz[3 * i + 2] - cz;
        xx += rx * rx; xy += rx * ry; xz += rx * rz;
        yy += ry * ry; yz += ry * rz; zz += rz * rz;
    }
    xx *= recip; xy *= recip; xz *= recip; yy *= recip; yz *= recip; zz *= recip;
    double wx = 0, wy = 0, wz = 0;
    double ax[3][3];
    double det[3];
    det[0] = yy * zz - yz * yz;
    ax[0][0] = det[0]; ax[0][1] = xz * yz - xy * zz; ax[0][2] = xy * yz - xz * yy;
    det[1] = xx * zz - xz * xz;
    ax[1][0] = xz * yz - xy * zz; ax[1][1] = det[1]; ax[1][2] = xy * xz - yz * xx;
    det[2] = xx * yy - xy * xy;
    ax[2][0] = xy * yz - xz * yy; ax[2][1] = xy * xz - yz * xx; ax[2][2] = det[2];
    for (int a = 0; a < 3; a++) {
        double weight = det[a] * det[a];
        double tx = wx * ax[a][0], ty = wy * ax[a][1], tz = wz * ax[a][2];
        if ((tx + ty) + tz < 0.0) weight = -weight;
        wx += ax[a][0] * weight; wy += ax[a][1] * weight; wz += ax[a][2] * weight;
    }
    double tx = wx * wx, ty = wy * wy, tz = wz * wz;
    double inv = 1.0 / sqrt((tx + ty) + tz);
    out[0] = (float)(wx * inv); out[1] = (float)(wy * inv); out[2] = (float)(wz * inv);
}

/* morton.hpp:59-80 (voxelise + encode), :81-102 (descending sort; ties by input index =
 * canonical), normals.hpp:81-148 (greedy Morton neighbourhoods, min 8 points, depths 0..2;
 * the element at index n-1 can never be absorbed, and an out-of-bounds compare counts as a
 * mismatch: SURVEY.md section 9 Q3). `order[i]` = input index of sorted point i. */
void oracle_stage_points(const float* xyz, size_t n, const float* pos, float sdf_res, float* xyz_sorted,
                         uint64_t* keys, uint32_t* order, float* normals) {
    if (n == 0) return;
    const float recip = (float)(1.0 / (double)sdf_res); /* morton.hpp:63 */
    uint64_t* inv = xmalloc(n * 8);
    for (size_t i = 0; i < n; i++) {
        float vx = floorf(xyz[3 * i] * recip), vy = floorf(xyz[3 * i + 1] * recip), vz = floorf(xyz[3 * i + 2] * recip);
        uint64_t k = oracle_morton_encode((int32_t)vx, (int32_t)vy, (int32_t)vz);
        inv[i] = ~k; /* ascending ~key == descending key; LSD radix is stable => ties by input index */
        order[i] = (uint32_t)i;
    }
    radix_sort_u64(inv, order, n);
    for (size_t i = 0; i < n; i++) {
        keys[i] = ~inv[i];
        memcpy(&xyz_sorted[3 * i], &xyz[3 * (size_t)order[i]], 12);
    }
    free(inv);
    if (!normals) return;
    const size_t min_points = 8;
    for (size_t it = 0; it < n;) {
        size_t end = it + 1;
        for (int depth = 0; depth < 3; depth++) {
            uint64_t mask = ~(uint64_t)0 << (uint64_t)(depth * 3);
            uint64_t neigh = mask & keys[it];
            /* normals.hpp:100: `end != cend() - 1`; when it == n-1, end == n and the reference reads one
             * past the end -- canonical: mismatch */
            while (end != n - 1 && end < n) {
                if ((mask & keys[end]) == neigh) end++;
                else break;
            }
            if (end - it >= min_points) break;
        }
        size_t size = end - it;
        if (size >= min_points) {
            float nrm[3];
            estimate_normal(xyz_sorted, it, end, nrm);
            /* normals.hpp:117-118: flip toward the sensor, judged at the FIRST point of the neighbourhood */
            float dx = pos[0] - xyz_sorted[3 * it], dy = pos[1] - xyz_sorted[3 * it + 1], dz = pos[2] - xyz_sorted[3 * it + 2];
            float invl = 1.0f / sqrtf(vdot3(dx, dy, dz, dx, dy, dz));
            float d = vdot3(nrm[0], nrm[1], nrm[2], dx * invl, dy * invl, dz * invl);
            if (d < 0.0f) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
            for (size_t j = it; j < end; j++) memcpy(&normals[3 * j], nrm, 12);
        } else {
            for (size_t j = it; j < end; j++) { /* normals.hpp:127-134 */
                float dx = pos[0] - xyz_sorted[3 * j], dy = pos[1] - xyz_sorted[3 * j + 1], dz = pos[2] - xyz_sorted[3 * j + 2];
                float invl = 1.0f / sqrtf(vdot3(dx, dy, dz, dx, dy, dz));
                normals[3 * j] = dx * invl; normals[3 * j + 1] = dy * invl; normals[3 * j + 2] = dz * invl;
            }
        }
        it += size;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* truncation-band voxel enumeration + signed distance: octree.hpp:86-159                       */
/* ------------------------------------------------------------------------------------------ */
#define ORACLE_MAX_RAY_VOXELS 4096
static size_t ray_band(const float p[3], const float nrm[3], const float pos[3], float res, float trunc, float recip,
                       uint64_t* keys, float* sd) {
    float d[3], dir[3], dir_recip[3], start[3], fin[3], delta[3], tmax[3];
    int32_t vs[3], vf[3], step[3], cur[3];
    for (int a = 0; a < 3; a++) d[a] = p[a] - pos[a];
    float invl = 1.0f / sqrtf(vdot3(d[0], d[1], d[2], d[0], d[1], d[2])); /* :92 normalize */
    for (int a = 0; a < 3; a++) {
        dir[a] = d[a] * invl;
        dir_recip[a] = 1.0f / dir[a];                 /* :93 */
        start[a] = p[a] - dir[a] * trunc;             /* :94 */
        fin[a] = p[a] + dir[a] * trunc;               /* :95 */
        vs[a] = (int32_t)floorf(start[a] * recip);    /* :96 */
        vf[a] = (int32_t)floorf(fin[a] * recip);      /* :97 */
        int32_t dv = vf[a] - vs[a];
        step[a] = (0 < dv) - (dv < 0);                /* :100 */
        delta[a] = fabsf(res * dir_recip[a]);         /* :102 */
        float m;                                      /* :104-116 */
        if (step[a] < 0) m = res * floorf(start[a] * recip);
        else if (step[a] > 0) m = res * ceilf(start[a] * recip);
        else m = FLT_MAX;
        m = m - start[a];                             /* :117 */
        tmax[a] = fabsf(m * dir_recip[a]);            /* :118 */
        cur[a] = vs[a];
    }
    size_t cnt = 0;
    for (;;) {
        /* emit cur (first voxel always, :121-122; later voxels at the bottom of the loop body, :151) */
        if (cnt >= ORACLE_MAX_RAY_VOXELS) abort();
        if (keys) {
            uint64_t k = oracle_morton_encode(cur[0], cur[1], cur[2]);
            int32_t x, y, z;
            oracle_morton_decode(k, &x, &y, &z); /* :157 decodes the stored code */
            float s = vdot3(nrm[0], nrm[1], nrm[2], (float)x * res - p[0], (float)y * res - p[1], (float)z * res - p[2]);
            s = (s < -trunc) ? -trunc : ((trunc < s) ? trunc : s); /* std::clamp, :159 */
            keys[cnt] = k;
            sd[cnt] = s;
        }
        cnt++;
        int a; /* :126-150: strict <, axis preference x / z / y / z */
        if (tmax[0] < tmax[1]) a = (tmax[0] < tmax[2]) ? 0 : 2;
        else a = (tmax[1] < tmax[2]) ? 1 : 2;
        cur[a] += step[a];
        tmax[a] += delta[a];
        if (cur[a] == vf[a] + step[a]) break;
    }
    return cnt;
}

size_t oracle_stage_pairs(const float* xyz_sorted, const float* normals, size_t n, const float* pos, float sdf_res,
                          float sdf_trunc, uint64_t* keys, float* sd, uint32_t* counts) {
    const float recip = (float)(1.0 / (double)sdf_res); /* octree.hpp:82 */
    size_t total = 0;
    for (size_t i = 0; i < n; i++) {
        size_t c = ray_band(&xyz_sorted[3 * i], &normals[3 * i], pos, sdf_res, sdf_trunc, recip, keys ? keys + total : NULL,
                            keys ? sd + total : NULL);
        if (counts) counts[i] = (uint32_t)c;
        total += c;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* DAG levels: levels.hpp:57-88 (NodeLevel::add), :123-139 (LeafClusterLevel::add)              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t* raw; size_t raw_cap; uint32_t occupied; uint32_t uniques, dupes;
    uint32_t* table; size_t table_cap; size_t table_n; /* open addressing over addresses, 0 = empty */
} node_level;
typedef struct {
    uint64_t* raw; size_t raw_cap; uint32_t uniques, dupes;
    uint32_t* table; size_t table_cap; size_t table_n;
} cluster_level;

static uint64_t mix64(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33; return h; }
static uint64_t node_hash(const uint32_t* rec) {
    int n = __builtin_popcount(rec[0] & 0xFF);
    uint64_t h = rec[0] & 0xFF;
    for (int i = 1; i <= n; i++) h = mix64(h ^ ((uint64_t)rec[i] << 8));
    return mix64(h);
}
static int node_eq(const uint32_t* a, const uint32_t* b) { /* levels.hpp:27-44 */
    if ((a[0] & 0xFF) != (b[0] & 0xFF)) return 0;
    int n = __builtin_popcount(a[0] & 0xFF);
    return memcmp(a + 1, b + 1, (size_t)n * 4) == 0;
}
static void node_level_init(node_level* L) {
    memset(L, 0, sizeof *L);
    L->raw_cap = 64; L->raw = xmalloc(L->raw_cap * 4);
    L->raw[0] = 0; L->occupied = 1; /* levels.hpp:52-54: index 0 reserved */
    L->table_cap = 64; L->table = calloc(L->table_cap, 4);
}
static void node_table_insert(node_level* L, uint32_t addr) {
    size_t i = node_hash(L->raw + addr) & (L->table_cap - 1);
    while (L->table[i]) i = (i + 1) & (L->table_cap - 1);
    L->table[i] = addr;
}
static uint32_t node_level_add(node_level* L, const uint32_t children[8]) {
    if ((size_t)L->occupied + 9 >= L->raw_cap) { L->raw_cap *= 2; L->raw = xrealloc(L->raw, L->raw_cap * 4); }
    uint32_t* rec = L->raw + L->occupied;
    rec[0] = 0;
    int cn = 0;
    for (int i = 0; i < 8; i++) { /* levels.hpp:68-74: compact the non-zero children, mask bit i */
        if (children[i] == 0) continue;
        rec[cn + 1] = children[i];
        rec[0] |= 1u << i;
        cn++;
    }
    if ((L->table_n + 1) * 2 > L->table_cap) {
        uint32_t* old = L->table; size_t oc = L->table_cap;
        L->table_cap *= 2; L->table = calloc(L->table_cap, 4);
        for (size_t i = 0; i < oc; i++) if (old[i]) node_table_insert(L, old[i]);
        free(old);
    }
    size_t i = node_hash(rec) & (L->table_cap - 1);
    while (L->table[i]) {
        if (node_eq(L->raw + L->table[i], rec)) { L->dupes++; return L->table[i]; } /* :83-86 */
        i = (i + 1) & (L->table_cap - 1);
    }
    uint32_t addr = L->occupied; /* :76-82 */
    L->table[i] = addr; L->table_n++;
    L->uniques++;
    L->occupied += (uint32_t)cn + 1;
    return addr;
}
static void cluster_level_init(cluster_level* L) {
    memset(L, 0, sizeof *L);
    L->raw_cap = 64; L->raw = xmalloc(L->raw_cap * 8);
    L->raw[0] = 0; /* levels.hpp:119-120: index 0 reserved */
    L->table_cap = 64; L->table = calloc(L->table_cap, 4);
}
static uint32_t cluster_level_add(cluster_level* L, uint64_t value) {
    if ((size_t)L->uniques + 2 >= L->raw_cap) { L->raw_cap *= 2; L->raw = xrealloc(L->raw, L->raw_cap * 8); }
    if ((L->table_n + 1) * 2 > L->table_cap) {
        uint32_t* old = L->table; size_t oc = L->table_cap;
        L->table_cap *= 2; L->table = calloc(L->table_cap, 4);
        for (size_t j = 0; j < oc; j++) if (old[j]) {
            size_t i = mix64(L->raw[old[j]]) & (L->table_cap - 1);
            while (L->table[i]) i = (i + 1) & (L->table_cap - 1);
            L->table[i] = old[j];
        }
        free(old);
    }
    size_t i = mix64(value) & (L->table_cap - 1);
    while (L->table[i]) {
        if (L->raw[L->table[i]] == value) { L->dupes++; return L->table[i]; } /* :135-138 */
        i = (i + 1) & (L->table_cap - 1);
    }
    uint32_t addr = L->uniques + 1; /* :125, :130-133 */
    L->raw[addr] = value;
    L->table[i] = addr; L->table_n++;
    L->uniques++;
    return addr;
}

/* cluster.hpp:13-32 (TSDFs::set / set_empty): byte = uint64_t(clamp(sd * (1/trunc), -1, 1) * 127 + 127),
 * 0xFF for an absent voxel; slot = key & 7. submap.hpp:24: 1/trunc is a float division. */
uint64_t oracle_quantise_cluster(const float* sd, uint32_t present_mask, float sdf_trunc) {
    const float trunc_recip = 1.0f / sdf_trunc;
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) {
        if (!(present_mask & (1u << i))) { v |= (uint64_t)0xFF << (8 * i); continue; }
        float s = sd[i] * trunc_recip;
        s = (s < -1.0f) ? -1.0f : ((1.0f < s) ? 1.0f : s);
        s = s * 127.0f + 127.0f;
        v |= (uint64_t)s << (uint64_t)(8 * i);
    }
    return v;
}

/* ------------------------------------------------------------------------------------------ */
/* the map                                                                                      */
/* ------------------------------------------------------------------------------------------ */
struct oracle_map {
    float res, trunc;
    /* active submap working set ("octree leaves", octree.hpp:15), ascending key */
    uint64_t* vk; float* vsd; uint32_t* vw; size_t vn, vcap;
    /* active submap trajectory (submap.hpp:110): only first pose + count matter */
    int has_pose; float first_pose[3];
    node_level nodes[20];
    cluster_level clusters;
    uint32_t* roots; uint32_t nsub, subcap;
    uint64_t last_updates, last_distinct;
};

oracle_map* oracle_create(float sdf_res, float sdf_trunc) {
    oracle_map* m = calloc(1, sizeof *m);
    m->res = sdf_res; m->trunc = sdf_trunc;
    for (int d = 0; d < 20; d++) node_level_init(&m->nodes[d]);
    cluster_level_init(&m->clusters);
    return m;
}
void oracle_destroy(oracle_map* m) {
    if (!m) return;
    for (int d = 0; d < 20; d++) { free(m->nodes[d].raw); free(m->nodes[d].table); }
    free(m->clusters.raw); free(m->clusters.table);
    free(m->vk); free(m->vsd); free(m->vw); free(m->roots); free(m);
}

/* submap.hpp:10-106 restated bottom-up: the DFS visits children 0..7, i.e. ascending Morton, and at
 * every level emits "tsdf record, weight record" per node into that level's single dedup set; so
 * level d's add sequence is: for each level-d node in ascending Morton order, tsdf then weight
 * ("Level geometry"). Addresses of one level never depend on another level's
 * set, only on its children's addresses, so levels can be processed one after another. */
static void finalize_submap(oracle_map* m) {
    size_t n = m->vn;
    /* leaf clusters: group voxels by key >> 3 (submap.hpp:74-100) */
    uint64_t* ids = xmalloc((n + 1) * 8);
    uint32_t* at = xmalloc((n + 1) * 4);
    uint32_t* aw = xmalloc((n + 1) * 4);
    size_t cn = 0;
    for (size_t i = 0; i < n;) {
        uint64_t cid = m->vk[i] >> 3;
        float sd[8]; uint32_t mask = 0;
        size_t j = i;
        for (; j < n && (m->vk[j] >> 3) == cid; j++) { int s = (int)(m->vk[j] & 7); sd[s] = m->vsd[j]; mask |= 1u << s; }
        uint64_t tsdf = oracle_quantise_cluster(sd, mask, m->trunc);
        /* submap.hpp:92-95 + cluster.hpp:58-65: max<uint8_t>(w, 255) == 255, so every weight byte is 0xFF
         * whether the voxel exists or not */
        uint64_t weigh = ~(uint64_t)0;
        ids[cn] = cid;
        at[cn] = cluster_level_add(&m->clusters, tsdf);
        aw[cn] = cluster_level_add(&m->clusters, weigh);
        cn++;
        i = j;
    }
    /* node levels 19 .. 0 (submap.hpp:31-61) */
    for (int d = 19; d >= 0; d--) {
        size_t pn = 0;
        for (size_t i = 0; i < cn;) {
            uint64_t pid = ids[i] >> 3;
            uint32_t ct[8] = {0}, cw[8] = {0};
            size_t j = i;
            for (; j < cn && (ids[j] >> 3) == pid; j++) { int s = (int)(ids[j] & 7); ct[s] = at[j]; cw[s] = aw[j]; }
            uint32_t a = node_level_add(&m->nodes[d], ct);
            uint32_t b = node_level_add(&m->nodes[d], cw);
            ids[pn] = pid; at[pn] = a; aw[pn] = b; pn++; /* in place: pn <= i */
            i = j;
        }
        if (cn == 0 && d == 0) { /* empty octree: the root is still added twice (submap.hpp:31-46) */
            uint32_t z[8] = {0};
            at[0] = node_level_add(&m->nodes[0], z);
            aw[0] = node_level_add(&m->nodes[0], z);
            pn = 1;
        }
        cn = pn;
    }
    if (m->nsub == m->subcap) { m->subcap = m->subcap ? m->subcap * 2 : 8; m->roots = xrealloc(m->roots, (size_t)m->subcap * 8); }
    m->roots[2 * m->nsub] = at[0];
    m->roots[2 * m->nsub + 1] = aw[0];
    m->nsub++;
    free(ids); free(at); free(aw);
}

/* tsdf.cpp:39-75 */
uint32_t oracle_insert(oracle_map* m, const float* xyz, size_t n, const float* pos) {
    /* tsdf.cpp:46-61: submap switch when the pose is > 5 m from the submap's FIRST pose */
    if (!m->has_pose) { m->has_pose = 1; memcpy(m->first_pose, pos, 12); }
    else {
        float dx = m->first_pose[0] - pos[0], dy = m->first_pose[1] - pos[1], dz = m->first_pose[2] - pos[2]; /* distance(pose, start) = length(start - pose) */
        if (sqrtf(vdot3(dx, dy, dz, dx, dy, dz)) > 5.0f) {
            finalize_submap(m);
            m->vn = 0; /* octree.clear(), tsdf.cpp:57 */
            memcpy(m->first_pose, pos, 12);
        }
    }
    m->last_updates = m->last_distinct = 0;
    if (n == 0) return m->nsub;
    float* sorted = xmalloc(n * 12); float* nrm = xmalloc(n * 12);
    uint64_t* pk = xmalloc(n * 8); uint32_t* order = xmalloc(n * 4);
    oracle_stage_points(xyz, n, pos, m->res, sorted, pk, order, nrm);
    size_t U = oracle_stage_pairs(sorted, nrm, n, pos, m->res, m->trunc, NULL, NULL, NULL);
    uint64_t* uk = xmalloc(U * 8); float* usd = xmalloc(U * 4); uint32_t* uidx = xmalloc(U * 4);
    oracle_stage_pairs(sorted, nrm, n, pos, m->res, m->trunc, uk, usd, NULL);
    /* the octree applies updates in (sorted point, ray step) order (octree.hpp:86,153); per voxel that is
     * the emission order, which a STABLE sort by voxel key preserves */
    for (size_t i = 0; i < U; i++) uidx[i] = (uint32_t)i;
    radix_sort_u64(uk, uidx, U);
    /* merge the sorted update segments into the sorted resident set */
    size_t cap = m->vn + U;
    uint64_t* nk = xmalloc(cap * 8); float* nsd = xmalloc(cap * 4); uint32_t* nw = xmalloc(cap * 4);
    size_t o = 0, r = 0, distinct = 0;
    for (size_t i = 0; i < U;) {
        uint64_t k = uk[i];
        while (r < m->vn && m->vk[r] < k) { nk[o] = m->vk[r]; nsd[o] = m->vsd[r]; nw[o] = m->vw[r]; o++; r++; }
        float acc = 0.0f; uint32_t w = 0; /* new leaf: octree.hpp:68-75 */
        if (r < m->vn && m->vk[r] == k) { acc = m->vsd[r]; w = m->vw[r]; r++; }
        for (; i < U && uk[i] == k; i++) { /* octree.hpp:161-163: three rounded operations per update */
            acc = acc * (float)w + usd[uidx[i]];
            w++;
            acc = acc / (float)w;
        }
        nk[o] = k; nsd[o] = acc; nw[o] = w; o++;
        distinct++;
    }
    while (r < m->vn) { nk[o] = m->vk[r]; nsd[o] = m->vsd[r]; nw[o] = m->vw[r]; o++; r++; }
    free(m->vk); free(m->vsd); free(m->vw);
    m->vk = nk; m->vsd = nsd; m->vw = nw; m->vn = o; m->vcap = cap;
    m->last_updates = U; m->last_distinct = distinct;
    free(sorted); free(nrm); free(pk); free(order); free(uk); free(usd); free(uidx);
    return m->nsub;
}

/* tsdf.cpp:78-81 (the part of save() before meshing); afterwards a fresh active submap is started so
 * the map stays usable (the reference's save() is terminal) */
uint32_t oracle_finalize_active(oracle_map* m) {
    if (m->has_pose) {
        finalize_submap(m);
        m->vn = 0;
        m->has_pose = 0;
    }
    return m->nsub;
}
size_t oracle_voxel_count(const oracle_map* m) { return m->vn; }
size_t oracle_export_voxels(const oracle_map* m, uint64_t* keys, uint32_t* sd_bits, uint32_t* w) {
    memcpy(keys, m->vk, m->vn * 8); memcpy(sd_bits, m->vsd, m->vn * 4); memcpy(w, m->vw, m->vn * 4);
    return m->vn;
}
uint32_t oracle_submap_count(const oracle_map* m) { return m->nsub; }
void oracle_submap_roots(const oracle_map* m, uint32_t i, uint32_t* tsdf, uint32_t* weight) { *tsdf = m->roots[2 * i]; *weight = m->roots[2 * i + 1]; }
size_t oracle_level_words(const oracle_map* m, int level) { return level < 20 ? m->nodes[level].occupied : (size_t)m->clusters.uniques + 1; }
void oracle_level_counters(const oracle_map* m, int level, uint32_t* uniques, uint32_t* dupes) {
    if (level < 20) { *uniques = m->nodes[level].uniques; *dupes = m->nodes[level].dupes; }
    else { *uniques = m->clusters.uniques; *dupes = m->clusters.dupes; }
}
void oracle_export_level(const oracle_map* m, int level, void* dst) {
    if (level < 20) memcpy(dst, m->nodes[level].raw, (size_t)m->nodes[level].occupied * 4);
    else memcpy(dst, m->clusters.raw, ((size_t)m->clusters.uniques + 1) * 8);
}
void oracle_last_scan_stats(const oracle_map* m, uint64_t* updates, uint64_t* distinct_voxels) { *updates = m->last_updates; *distinct_voxels = m->last_distinct; }
