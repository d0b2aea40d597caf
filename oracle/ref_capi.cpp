// TEST INFRASTRUCTURE ONLY (oracle/). Never linked into, imported by or called from the
// product path (chad_tsdf_b200/, include/); only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
//
// C-ABI harness around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/): it drives chad::TSDFMap::insert
// (src/chad/tsdf.cpp:39-75) and Submap::finalize (include/chad/detail/submap.hpp:10-106)
// and dumps the state the parity tests compare against:
//   * the active octree's leaves as (Morton key, sd bits, u32 weight), ascending key
//     (include/chad/detail/octree.hpp:13-15,31-78);
//   * every NodeLevel's _raw_data[0.._occupied_n) and the LeafClusterLevel's
//     _raw_data[0.._uniques_n], with _uniques_n/_dupes_n (include/chad/detail/levels.hpp:90-93,141-143);
//   * per-submap roots (submap.hpp:108-109);
//   * the intermediate point stage (morton.hpp:59-102, normals.hpp:81-148).
// Built twice: "verbatim" (std::sort, morton.hpp:89) and "stable" (-DCHAD_REF_STABLE: the
// token `sort` is macro-renamed to `stable_sort` while the reference headers are parsed --
// the one-token canonicalisation of SURVEY.md section 8c, done without editing the reference).
#include <algorithm>
#include <array>
#include <bit>
#include <bitset>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>
#include <fmt/base.h>
#include <glm/glm.hpp>
#include <gtl/phmap.hpp>
#include <libmorton/morton.h>

#ifdef CHAD_REF_STABLE
#define sort stable_sort
#endif
#define private public
#include "chad/tsdf.hpp"
#undef private
#include "chad/detail/levels.hpp"
#include "chad/detail/morton.hpp"
#include "chad/detail/normals.hpp"
#include "chad/detail/octree.hpp"
#include "chad/detail/submap.hpp"
#include "chad/detail/lvr2.hpp"
#ifdef CHAD_REF_STABLE
#undef sort
#endif

// LVR2 meshing is out of scope; TSDFMap::save is never called here.
namespace chad::detail {
void reconstruct(const detail::Submap&, const NodeLevels&, float, float, std::string_view) {}
}  // namespace chad::detail

namespace {
struct Handle {
    chad::TSDFMap* map;
    std::vector<std::array<uint32_t, 2>> roots;  // per finalised submap: tsdf, weight
};

void walk_octree(const chad::detail::Octree& oct, uint64_t* keys, uint32_t* sd_bits, uint32_t* w, size_t& n, bool write) {
    // DFS, children 0..7 => ascending Morton (same order as submap.hpp:26-102)
    struct Frame { uint32_t node; uint32_t child; uint64_t prefix; };
    Frame stack[22];
    int depth = 0;
    stack[0] = {0, 0, 0};
    n = 0;
    while (depth >= 0) {
        Frame& f = stack[depth];
        if (f.child == 8) { depth--; continue; }
        uint32_t ci = f.child++;
        uint32_t addr = oct.get_child_addr(f.node, uint8_t(ci));
        if (addr == 0) continue;
        uint64_t prefix = (f.prefix << 3) | ci;
        if (depth == 20) {  // children of a depth-20 node are leaf addresses
            if (write) {
                const auto& leaf = oct.get_leaf(addr);
                keys[n] = prefix;
                std::memcpy(&sd_bits[n], &leaf._signed_distance, 4);
                w[n] = leaf._weight;
            }
            n++;
        } else {
            depth++;
            stack[depth] = {addr, 0, prefix};
        }
    }
}
}  // namespace

extern "C" {
const char* chadref_variant() {
#ifdef CHAD_REF_STABLE
    return "stable";
#else
    return "verbatim";
#endif
}
void* chadref_create(float res, float trunc) {
    auto* h = new Handle{new chad::TSDFMap(res, trunc), {}};
    return h;
}
void chadref_destroy(void* hp) {
    auto* h = static_cast<Handle*>(hp);
    // the reference leaks the active submap unless save() pushed it (tsdf.cpp:32-38)
    bool pushed = false;
    for (auto* s : h->map->_submaps) pushed |= (s == h->map->_active_submap_p);
    if (!pushed) delete h->map->_active_submap_p;
    delete h->map;
    delete h;
}
// returns the number of submaps finalised so far (a switch inside insert adds one)
uint32_t chadref_insert(void* hp, const float* xyz, size_t n, const float* pos) {
    auto* h = static_cast<Handle*>(hp);
    size_t before = h->map->_submaps.size();
    h->map->insert(xyz, n, pos);
    for (size_t i = before; i < h->map->_submaps.size(); i++)
        h->roots.push_back({h->map->_submaps[i]->root_addr_tsdf, h->map->_submaps[i]->root_addr_weight});
    return uint32_t(h->map->_submaps.size());
}
// what TSDFMap::save does before meshing (tsdf.cpp:78-81), then starts a fresh active submap so the
// map stays usable (the reference's save() is terminal)
uint32_t chadref_finalize_active(void* hp) {
    auto* h = static_cast<Handle*>(hp);
    auto* m = h->map;
    if (!m->_active_submap_p->positions.empty()) {
        m->_active_submap_p->finalize(*m->_active_octree_p, *m->_node_levels_p, m->_sdf_trunc);
        m->_submaps.push_back(m->_active_submap_p);
        h->roots.push_back({m->_active_submap_p->root_addr_tsdf, m->_active_submap_p->root_addr_weight});
        m->_active_submap_p = new chad::detail::Submap();
        m->_active_octree_p->clear();
    }
    return uint32_t(m->_submaps.size());
}
size_t chadref_voxel_count(void* hp) {
    auto* h = static_cast<Handle*>(hp);
    size_t n;
    walk_octree(*h->map->_active_octree_p, nullptr, nullptr, nullptr, n, false);
    return n;
}
size_t chadref_export_voxels(void* hp, uint64_t* keys, uint32_t* sd_bits, uint32_t* w) {
    auto* h = static_cast<Handle*>(hp);
    size_t n;
    walk_octree(*h->map->_active_octree_p, keys, sd_bits, w, n, true);
    return n;
}
uint32_t chadref_submap_count(void* hp) { return uint32_t(static_cast<Handle*>(hp)->roots.size()); }
void chadref_submap_roots(void* hp, uint32_t i, uint32_t* tsdf, uint32_t* weight) {
    auto* h = static_cast<Handle*>(hp);
    *tsdf = h->roots[i][0];
    *weight = h->roots[i][1];
}
// level 0..19 = NodeLevel (u32 words), level 20 = leaf clusters (u64 words, index 0 reserved)
size_t chadref_level_words(void* hp, int level) {
    auto* nl = static_cast<Handle*>(hp)->map->_node_levels_p;
    if (level < 20) return nl->_nodes[level]._occupied_n;
    return size_t(nl->_leaf_clusters._uniques_n) + 1;
}
void chadref_level_counters(void* hp, int level, uint32_t* uniques, uint32_t* dupes) {
    auto* nl = static_cast<Handle*>(hp)->map->_node_levels_p;
    if (level < 20) { *uniques = nl->_nodes[level]._uniques_n; *dupes = nl->_nodes[level]._dupes_n; }
    else { *uniques = nl->_leaf_clusters._uniques_n; *dupes = nl->_leaf_clusters._dupes_n; }
}
void chadref_export_level(void* hp, int level, void* dst) {
    auto* nl = static_cast<Handle*>(hp)->map->_node_levels_p;
    if (level < 20) std::memcpy(dst, nl->_nodes[level]._raw_data.data(), size_t(nl->_nodes[level]._occupied_n) * 4);
    else std::memcpy(dst, nl->_leaf_clusters._raw_data.data(), (size_t(nl->_leaf_clusters._uniques_n) + 1) * 8);
}
// intermediate point stage: sorted points, their Morton keys and normals (morton.hpp:59-102, normals.hpp:81-148)
void chadref_stage_points(const float* xyz, size_t n, const float* pos, float res, float* xyz_sorted, uint64_t* keys, float* normals) {
    using namespace chad::detail;
    std::vector<std::array<float, 3>> pts(n);
    std::memcpy(pts.data(), xyz, n * 12);
    MortonVector mv = calc_morton_vector(pts, res);
    std::vector<glm::vec3> sorted = sort_morton_vector(mv);
    std::vector<glm::vec3> nrm = estimate_normals(mv, glm::vec3{pos[0], pos[1], pos[2]});
    for (size_t i = 0; i < n; i++) {
        xyz_sorted[3 * i + 0] = sorted[i].x; xyz_sorted[3 * i + 1] = sorted[i].y; xyz_sorted[3 * i + 2] = sorted[i].z;
        keys[i] = mv[i].second._value;
        normals[3 * i + 0] = nrm[i].x; normals[3 * i + 1] = nrm[i].y; normals[3 * i + 2] = nrm[i].z;
    }
}
// phase timers scraped from the reference's own fmt::println lines
int chadref_phase_count() { return chad_ref_shim::phase_log().n; }
const char* chadref_phase_tag(int i) { return chad_ref_shim::phase_log().tag[i]; }
double chadref_phase_sum_ms(int i) { return chad_ref_shim::phase_log().sum[i]; }
void chadref_phase_reset() { chad_ref_shim::phase_log().n = 0; }
}
