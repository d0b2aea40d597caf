// chad::TSDFMap facade over the C ABI -- the host-side counterpart of
// /root/reference/src/chad/tsdf.cpp:27-86 (constructor, destructor, insert, save).
#include "chad/tsdf.hpp"

#include <bit>
#include <cstdio>
#include <stdexcept>

#include "chad_b200.h"

namespace chad {
namespace {
void check(chad_ctx* ctx, int rc, const char* what) {
    if (rc != CHAD_OK) throw std::runtime_error(std::string("chad::TSDFMap::") + what + ": " + chad_last_error(ctx));
}
}  // namespace

TSDFMap::TSDFMap(float sdf_res, float sdf_trunc): _sdf_res(sdf_res), _sdf_trunc(sdf_trunc), _ctx(nullptr) {
    int device = 0;
    if (const char* env = std::getenv("CHAD_DEVICE")) device = std::atoi(env);
    check(nullptr, chad_create(sdf_res, sdf_trunc, device, 0, &_ctx), "TSDFMap");
}
TSDFMap::~TSDFMap() { chad_destroy(_ctx); }

void TSDFMap::insert(const float* points_p, size_t points_count, const float* position_p) {
    check(_ctx, chad_insert(_ctx, points_p, points_count, position_p), "insert");
}
void TSDFMap::flush() { check(_ctx, chad_flush(_ctx), "flush"); }
size_t TSDFMap::submap_count() {
    uint32_t n = 0;
    check(_ctx, chad_submap_count(_ctx, &n), "submap_count");
    return n;
}
std::array<uint32_t, 2> TSDFMap::submap_roots(size_t i) {
    std::array<uint32_t, 2> r{};
    check(_ctx, chad_submap_roots(_ctx, (uint32_t)i, &r[0], &r[1]), "submap_roots");
    return r;
}
HostNodeLevels TSDFMap::node_levels() {
    HostNodeLevels out;
    for (int level = 0; level < CHAD_NUM_LEVELS; level++) {
        size_t words = 0;
        check(_ctx, chad_level_words(_ctx, level, &words), "node_levels");
        if (level == CHAD_LEVEL_CLUSTERS) {
            out.leaf_clusters.resize(words);
            check(_ctx, chad_export_level(_ctx, level, out.leaf_clusters.data(), words), "node_levels");
        } else {
            out.nodes[level].resize(words);
            check(_ctx, chad_export_level(_ctx, level, out.nodes[level].data(), words), "node_levels");
        }
    }
    return out;
}

// Flat little-endian dump: "CHADDAG1", f32 sdf_res, f32 sdf_trunc, u32 n_submaps, n_submaps x (u32 root_tsdf,
// u32 root_weight), then for level 0..19: u64 n_words + n_words x u32, then u64 n_clusters + n x u64.
void TSDFMap::save(const std::string& filename) {
    check(_ctx, chad_finalize_active(_ctx), "save");  // tsdf.cpp:78-81
    const HostNodeLevels levels = node_levels();
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("chad::TSDFMap::save: cannot open " + filename);
    auto put = [&](const void* p, size_t n) { if (std::fwrite(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("chad::TSDFMap::save: write failed"); } };
    put("CHADDAG1", 8);
    put(&_sdf_res, 4);
    put(&_sdf_trunc, 4);
    const uint32_t n_sub = (uint32_t)submap_count();
    put(&n_sub, 4);
    for (uint32_t i = 0; i < n_sub; i++) { auto r = submap_roots(i); put(r.data(), 8); }
    for (const auto& lv : levels.nodes) { const uint64_t n = lv.size(); put(&n, 8); put(lv.data(), n * 4); }
    const uint64_t n = levels.leaf_clusters.size();
    put(&n, 8);
    put(levels.leaf_clusters.data(), n * 8);
    std::fclose(f);
}

uint32_t HostNodeLevels::get_child_addr(uint32_t depth, uint32_t parent_addr, uint8_t child_i) const {
    const uint32_t child_mask = nodes[depth][parent_addr];
    const uint32_t child_bit = 1u << child_i;
    if (!(child_mask & child_bit)) return 0;
    const uint32_t before = (uint32_t)std::popcount(uint8_t(child_mask & (child_bit - 1)));
    return nodes[depth][parent_addr + before + 1];
}
bool HostNodeLevels::try_get_lc(uint32_t parent_addr, uint8_t child_i, uint64_t& cluster) const {
    const uint32_t addr = get_child_addr(MAX_DEPTH - 1, parent_addr, child_i);
    if (addr == 0) return false;
    cluster = leaf_clusters[addr];
    return true;
}
}  // namespace chad
