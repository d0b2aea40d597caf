// chad::TSDFMap facade over the C ABI -- the host-side counterpart of
// /root/reference/src/chad/tsdf.cpp:27-86 (constructor, destructor, insert, save).
#include "chad/tsdf.hpp"

#include <algorithm>
#include <bit>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

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

// The reference's ChadGrid constructor (lvr2.cpp:32-130) + ChadGrid::saveGrid (lvr2.cpp:170-200) on the host copy of the DAG.
void TSDFMap::save_grid(const std::string& filename, size_t submap) {
    check(_ctx, chad_finalize_active(_ctx), "save_grid");  // tsdf.cpp:78-81
    if (submap >= submap_count()) throw std::runtime_error("chad::TSDFMap::save_grid: no such submap");
    write_grid(node_levels(), submap_roots(submap)[0], _sdf_res, _sdf_trunc, filename);
}

void write_grid(const HostNodeLevels& levels, uint32_t root, float _sdf_res, float _sdf_trunc, const std::string& filename) {
    struct QueryPoint { float x, y, z, sd; };
    std::vector<QueryPoint> query_points;
    constexpr uint32_t INVALID = 0xFFFFFFFFu;  // lvr2 FastBox::INVALID_INDEX
    std::unordered_map<uint64_t, std::array<uint32_t, 8>> cells;
    static const int32_t cell_offsets[8][3] = { {0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {0, -1, 0}, {0, 0, -1}, {-1, 0, -1}, {-1, -1, -1}, {0, -1, -1} };  // lvr2.cpp:88-98
    std::array<uint8_t, HostNodeLevels::MAX_DEPTH> path_child{};
    std::array<uint32_t, HostNodeLevels::MAX_DEPTH> path_addr{};
    path_addr[0] = root;
    uint32_t depth = 0;
    while (true) {  // lvr2.cpp:33-113
        const uint8_t child_i = path_child[depth]++;
        if (child_i == 8) {
            if (depth > 0) depth--;
            else break;
        } else if (depth < HostNodeLevels::MAX_DEPTH - 1) {
            const uint32_t child_addr = levels.get_child_addr(depth, path_addr[depth], child_i);
            if (child_addr > 0) {
                depth++;
                path_child[depth] = 0;
                path_addr[depth] = child_addr;
            }
        } else {
            uint64_t cluster;
            if (!levels.try_get_lc(path_addr[depth], child_i, cluster)) continue;
            uint64_t code = 0;  // Morton code of the cluster from the path (lvr2.cpp:59-65)
            for (uint64_t k = 0; k < 63 / 3 - 1; k++) code |= uint64_t(path_child[k] - 1) << uint64_t(60 - k * 3);
            int32_t cx, cy, cz;
            chad_morton_decode(code, &cx, &cy, &cz);
            uint32_t leaf_i = 0;
            for (int32_t z = 0; z <= 1; z++) for (int32_t y = 0; y <= 1; y++) for (int32_t x = 0; x <= 1; x++, leaf_i++) {
                const uint64_t bits = (cluster >> (leaf_i * 8)) & 0xFF;  // TSDFs::try_get, cluster.hpp:34-52
                if (bits == 0xFF) continue;
                float sd = float(bits) - 127.0f;
                sd *= float(1.0 / 127.0f);
                sd *= _sdf_trunc;
                const int32_t lx = cx + x, ly = cy + y, lz = cz + z;
                const uint32_t qi = (uint32_t)query_points.size();
                query_points.push_back({ float(lx) * _sdf_res, float(ly) * _sdf_res, float(lz) * _sdf_res, sd });
                for (size_t i = 0; i < 8; i++) {
                    const uint64_t cell = chad_morton_encode(lx + cell_offsets[i][0], ly + cell_offsets[i][1], lz + cell_offsets[i][2]);
                    auto [it, fresh] = cells.try_emplace(cell);
                    if (fresh) it->second.fill(INVALID);
                    it->second[i] = qi;
                }
            }
        }
    }
    std::vector<uint64_t> complete;  // lvr2.cpp:115-129: cells with a missing corner are culled
    for (const auto& [code, verts] : cells)
        if (std::find(verts.begin(), verts.end(), INVALID) == verts.end()) complete.push_back(code);
    std::sort(complete.begin(), complete.end());
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("chad::write_grid: cannot open " + filename);
    auto put = [&](const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("chad::write_grid: write failed"); } };
    const float header = _sdf_trunc;  // the reference writes m_truncsize under the name voxel_res (lvr2.cpp:176-177)
    const size_t nq = query_points.size(), nc = complete.size();
    put(&header, sizeof(float));
    put(&nq, sizeof(size_t));
    put(&nc, sizeof(size_t));
    put(query_points.data(), nq * sizeof(QueryPoint));
    for (uint64_t code : complete) put(cells[code].data(), 8 * sizeof(uint32_t));
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
uint8_t HostNodeLevels::query(uint32_t root_addr, uint64_t morton_key) const {
    // depth d picks the child (key >> 3(20 - d)) & 7 (octree.hpp:44-56); the depth-19 node's children are leaf clusters whose
    // byte `key & 7` is the voxel (cluster.hpp:13-32)
    uint32_t addr = root_addr;
    if (addr == 0 || addr >= nodes[0].size()) return 0xFF;
    for (uint32_t depth = 0; depth + 1 < MAX_DEPTH; depth++) {
        addr = get_child_addr(depth, addr, uint8_t((morton_key >> (3 * (20 - depth))) & 7));
        if (addr == 0) return 0xFF;
    }
    uint64_t cluster;
    if (!try_get_lc(addr, uint8_t((morton_key >> 3) & 7), cluster)) return 0xFF;
    return uint8_t(cluster >> (8 * (morton_key & 7)));
}

SavedMap load_dag(const std::string& filename) {
    std::FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) throw std::runtime_error("chad::load_dag: cannot open " + filename);
    auto fail = [&](const char* what) { std::fclose(f); throw std::runtime_error("chad::load_dag: " + filename + ": " + what); };
    auto get = [&](void* p, size_t n) { if (n && std::fread(p, 1, n, f) != n) fail("truncated file"); };
    std::fseek(f, 0, SEEK_END);
    const uint64_t file_bytes = (uint64_t)std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    char magic[8];
    get(magic, 8);
    if (std::memcmp(magic, "CHADDAG1", 8) != 0) fail("not a CHADDAG1 file");
    SavedMap m;
    uint32_t n_sub = 0;
    get(&m.sdf_res, 4);
    get(&m.sdf_trunc, 4);
    get(&n_sub, 4);
    if (uint64_t(n_sub) * 8 > file_bytes) fail("submap count exceeds the file size");
    m.roots.resize(n_sub);
    for (auto& r : m.roots) get(r.data(), 8);
    for (auto& lv : m.levels.nodes) {
        uint64_t n = 0;
        get(&n, 8);
        if (n > file_bytes / 4) fail("level size exceeds the file size");
        lv.resize(n);
        get(lv.data(), n * 4);
    }
    uint64_t n = 0;
    get(&n, 8);
    if (n > file_bytes / 8) fail("cluster count exceeds the file size");
    m.levels.leaf_clusters.resize(n);
    get(m.levels.leaf_clusters.data(), n * 8);
    if ((uint64_t)std::ftell(f) != file_bytes) fail("trailing bytes");
    std::fclose(f);
    return m;
}
}  // namespace chad
