// chad::TSDFMap facade over the C ABI -- the host-side counterpart of
// /root/reference/src/chad/tsdf.cpp:27-86 (constructor, destructor, insert, save).
#include "chad/tsdf.hpp"

#include <algorithm>
#include <bit>
#include <condition_variable>
#include <cmath>
#include <functional>
#include <limits>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <unordered_map>

#include "chad_b200.h"

namespace chad {
namespace {
void check(chad_ctx* ctx, int rc, const char* what) {
    if (rc != CHAD_OK) throw std::runtime_error(std::string("chad::TSDFMap::") + what + ": " + chad_last_error(ctx));
}
}  // namespace

// ---- several GPUs behind the same class -------------------------------------------------------------------
// One worker thread per device, each owning the context of one Morton range (a context is single-threaded, like the reference's
// class). A collective call hands the same job to every worker and waits for all of them: every rank makes the same calls with the
// same arguments, which is what chad_create_sharded asks for.
struct TSDFMap::Group {
    std::vector<chad_ctx*> ctx;
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::function<int(chad_ctx*&, int)> job;
    uint64_t generation = 0;
    int pending = 0;
    bool quit = false;
    std::vector<int> rc;

    explicit Group(size_t n): ctx(n, nullptr), rc(n, CHAD_OK) {
        for (size_t r = 0; r < n; r++) threads.emplace_back([this, r] { work((int)r); });
    }
    ~Group() {
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cv_job.notify_all();
        for (auto& t : threads) t.join();
    }
    void work(int r) {
        uint64_t seen = 0;
        while (true) {
            std::function<int(chad_ctx*&, int)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
                f = job;
            }
            const int res = f(ctx[r], r);
            { std::lock_guard<std::mutex> lk(m); rc[r] = res; pending--; }
            cv_done.notify_all();
        }
    }
    // run f on every rank; returns the first failing rank (or -1)
    int run(std::function<int(chad_ctx*&, int)> f) {
        std::unique_lock<std::mutex> lk(m);
        job = std::move(f);
        pending = (int)threads.size();
        generation++;
        cv_job.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
        for (size_t r = 0; r < rc.size(); r++) if (rc[r] != CHAD_OK) return (int)r;
        return -1;
    }
};

namespace {
std::vector<int> devices_from_env() {
    std::vector<int> out;
    if (const char* env = std::getenv("CHAD_DEVICES")) {
        std::string tok;
        for (const char* p = env;; p++) {
            if (*p == ',' || *p == 0) { if (!tok.empty()) out.push_back(std::atoi(tok.c_str())); tok.clear(); if (!*p) break; }
            else tok.push_back(*p);
        }
    }
    return out;
}
}  // namespace

TSDFMap::TSDFMap(float sdf_res, float sdf_trunc): _sdf_res(sdf_res), _sdf_trunc(sdf_trunc), _ctx(nullptr) {
    const std::vector<int> devices = devices_from_env();
    if (devices.size() > 1) {
        alignas(8) unsigned char id[CHAD_SHARD_ID_BYTES];
        check(nullptr, chad_shard_unique_id(id), "TSDFMap");
        _group = new Group(devices.size());
        std::vector<std::string> errors(devices.size());
        const int world = (int)devices.size();
        const int bad = _group->run([&](chad_ctx*& c, int r) {
            const int rc = chad_create_sharded(sdf_res, sdf_trunc, devices[r], 0, r, world, id, &c);
            if (rc != CHAD_OK) errors[r] = chad_last_error(nullptr);  // (thread-local: must be read on the worker)
            return rc;
        });
        if (bad >= 0) {
            const std::string why = errors[bad];
            _group->run([](chad_ctx*& c, int) { chad_destroy(c); c = nullptr; return CHAD_OK; });
            delete _group;
            _group = nullptr;
            throw std::runtime_error("chad::TSDFMap::TSDFMap: rank " + std::to_string(bad) + ": " + why);
        }
        _ctx = _group->ctx[0];
        return;
    }
    int device = devices.size() == 1 ? devices[0] : 0;
    if (devices.empty()) if (const char* env = std::getenv("CHAD_DEVICE")) device = std::atoi(env);
    check(nullptr, chad_create(sdf_res, sdf_trunc, device, 0, &_ctx), "TSDFMap");
}
TSDFMap::~TSDFMap() {
    if (_group) {
        _group->run([](chad_ctx*& c, int) { chad_flush(c); return CHAD_OK; });  // every rank quiescent before any communicator goes away
        _group->run([](chad_ctx*& c, int) { chad_destroy(c); c = nullptr; return CHAD_OK; });
        delete _group;
    } else {
        chad_destroy(_ctx);
    }
}

// a call every rank must make (insert, flush, finalize): forwarded to all workers, or made directly on the single context
void TSDFMap::collective(const char* what, int (*call)(chad_ctx*, const void*), const void* arg) {
    if (!_group) { check(_ctx, call(_ctx, arg), what); return; }
    const int bad = _group->run([&](chad_ctx*& c, int) { return call(c, arg); });
    if (bad >= 0) check(_group->ctx[bad], CHAD_ERR_INVALID, what);
}

void TSDFMap::insert(const float* points_p, size_t points_count, const float* position_p) {
    struct Args { const float* p; size_t n; const float* pos; } a{points_p, points_count, position_p};
    collective("insert", [](chad_ctx* c, const void* v) { const Args* a = static_cast<const Args*>(v); return chad_insert(c, a->p, a->n, a->pos); }, &a);
}
void TSDFMap::flush() { collective("flush", [](chad_ctx* c, const void*) { return chad_flush(c); }, nullptr); }
void TSDFMap::finalize_active() { collective("save", [](chad_ctx* c, const void*) { return chad_finalize_active(c); }, nullptr); flush(); }
size_t TSDFMap::submap_count() {
    if (_group) flush();  // (a read settles the context it is made on; on a sharded map every rank has to)
    uint32_t n = 0;
    check(_ctx, chad_submap_count(_ctx, &n), "submap_count");
    return n;
}
std::array<uint32_t, 2> TSDFMap::submap_roots(size_t i) {
    if (_group) flush();
    std::array<uint32_t, 2> r{};
    check(_ctx, chad_submap_roots(_ctx, (uint32_t)i, &r[0], &r[1]), "submap_roots");
    return r;
}
HostNodeLevels TSDFMap::node_levels() {
    if (_group) flush();
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

// Flat little-endian dump: "CHADDAG2", f32 sdf_res, f32 sdf_trunc, u32 n_submaps, per submap
// (u32 root_tsdf, u32 root_weight, u32 n_poses, n_poses x 3 f32), then for level 0..19: u32 uniques, u32 dupes, u64 n_words,
// n_words x u32, then for the cluster level: u32 uniques, u32 dupes, u64 n_clusters, n x u64.
void save_dag(const SavedMap& m, const std::string& filename) {
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("chad::save_dag: cannot open " + filename);
    auto put = [&](const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("chad::save_dag: write failed"); } };
    put("CHADDAG2", 8);
    put(&m.sdf_res, 4);
    put(&m.sdf_trunc, 4);
    const uint32_t n_sub = (uint32_t)m.roots.size();
    put(&n_sub, 4);
    for (uint32_t i = 0; i < n_sub; i++) {
        put(m.roots[i].data(), 8);
        const uint32_t np = i < m.positions.size() ? (uint32_t)m.positions[i].size() : 0u;
        put(&np, 4);
        if (np) put(m.positions[i].data(), size_t(np) * 12);
    }
    for (size_t lv = 0; lv <= HostNodeLevels::MAX_DEPTH; lv++) {
        put(&m.uniques[lv], 4);
        put(&m.dupes[lv], 4);
        if (lv < HostNodeLevels::MAX_DEPTH) { const uint64_t n = m.levels.nodes[lv].size(); put(&n, 8); put(m.levels.nodes[lv].data(), n * 4); }
        else { const uint64_t n = m.levels.leaf_clusters.size(); put(&n, 8); put(m.levels.leaf_clusters.data(), n * 8); }
    }
    std::fclose(f);
}

void TSDFMap::save(const std::string& filename) {
    finalize_active();  // tsdf.cpp:78-81 (on every rank of a sharded map)
    SavedMap m;
    m.sdf_res = _sdf_res;
    m.sdf_trunc = _sdf_trunc;
    m.levels = node_levels();
    m.has_counters = true;
    for (int lv = 0; lv < CHAD_NUM_LEVELS; lv++) check(_ctx, chad_level_counters(_ctx, lv, &m.uniques[lv], &m.dupes[lv]), "save");
    const size_t n_sub = submap_count();
    for (size_t i = 0; i < n_sub; i++) {
        m.roots.push_back(submap_roots(i));
        m.positions.push_back(submap_positions(i));
    }
    save_dag(m, filename);
}

void TSDFMap::load(const std::string& filename) {
    if (_group) throw std::runtime_error("chad::TSDFMap::load: a saved map is restored into a single-GPU map (unset CHAD_DEVICES)");
    const SavedMap m = load_dag(filename);
    if (m.sdf_res != _sdf_res || m.sdf_trunc != _sdf_trunc) throw std::runtime_error("chad::TSDFMap::load: " + filename + " was saved with another voxel size / truncation");
    if (!m.has_counters) throw std::runtime_error("chad::TSDFMap::load: " + filename + " is a CHADDAG1 dump (no dedup counters): readable with load_dag, not restorable");
    chad_dag_image img{};
    for (size_t lv = 0; lv < HostNodeLevels::MAX_DEPTH; lv++) { img.node_words[lv] = m.levels.nodes[lv].data(); img.node_word_count[lv] = m.levels.nodes[lv].size(); }
    img.cluster_words = m.levels.leaf_clusters.data();
    img.cluster_word_count = m.levels.leaf_clusters.size();
    for (int lv = 0; lv < CHAD_NUM_LEVELS; lv++) { img.uniques[lv] = m.uniques[lv]; img.dupes[lv] = m.dupes[lv]; }
    std::vector<uint32_t> roots, counts;
    std::vector<float> poses;
    for (size_t i = 0; i < m.roots.size(); i++) {
        roots.push_back(m.roots[i][0]);
        roots.push_back(m.roots[i][1]);
        counts.push_back(i < m.positions.size() ? (uint32_t)m.positions[i].size() : 0u);
        if (i < m.positions.size()) for (const auto& p : m.positions[i]) poses.insert(poses.end(), p.begin(), p.end());
    }
    img.roots = roots.data();
    img.n_submaps = (uint32_t)m.roots.size();
    img.positions = poses.data();
    img.position_counts = counts.data();
    check(_ctx, chad_import_dag(_ctx, &img), "load");
}

std::vector<std::array<float, 3>> TSDFMap::submap_positions(size_t submap) {
    if (_group) flush();
    size_t n = 0;
    check(_ctx, chad_submap_positions(_ctx, (uint32_t)submap, nullptr, 0, &n), "submap_positions");
    std::vector<std::array<float, 3>> out(n);
    if (n) check(_ctx, chad_submap_positions(_ctx, (uint32_t)submap, out[0].data(), n, &n), "submap_positions");
    return out;
}

// ---- leaf iterator ------------------------------------------------------------------------------------
namespace {
Leaf make_leaf(uint64_t key, uint8_t byte, float sdf_res, float sdf_trunc) {
    int32_t x, y, z;
    chad_morton_decode(key, &x, &y, &z);
    float sd = float(byte) - 127.0f;  // TSDFs::try_get, cluster.hpp:46-50
    sd *= float(1.0 / 127.0f);
    sd *= sdf_trunc;
    return Leaf{float(x) * sdf_res, float(y) * sdf_res, float(z) * sdf_res, sd, byte, key};
}
}  // namespace

LeafCursor::LeafCursor(const HostNodeLevels& levels, uint32_t root_addr): _levels(&levels) {
    _addr[0] = root_addr;
    _leaf_i = 7;  // "behind the last voxel of a cluster": next() starts with the first cluster
    _cluster = ~0ull;
    if (root_addr == 0 || root_addr >= levels.nodes[0].size()) { _done = true; return; }
    next();
}
// depth-first, children in index order: the clusters come in ascending Morton order (submap.hpp:10-106 writes them in that order)
bool LeafCursor::next_cluster() {
    while (true) {
        if (_child[_depth] == 8) {  // this node is exhausted
            if (_depth == 0) return false;
            _depth--;
            continue;
        }
        const uint8_t c = _child[_depth]++;
        if (_depth + 1 < HostNodeLevels::MAX_DEPTH) {
            const uint32_t a = _levels->get_child_addr(_depth, _addr[_depth], c);
            if (a == 0) continue;
            _depth++;
            _addr[_depth] = a;
            _child[_depth] = 0;
        } else if (_levels->try_get_lc(_addr[_depth], c, _cluster)) {
            _cluster_key = 0;  // the path IS the key: child index d-th triple from the top (octree.hpp:44-56)
            for (uint32_t d = 0; d < HostNodeLevels::MAX_DEPTH; d++) _cluster_key = (_cluster_key << 3) | uint64_t(_child[d] - 1);
            return true;
        }
    }
}
void LeafCursor::next() {
    if (_done) return;
    while (true) {
        if (_leaf_i == 7) {
            if (!next_cluster()) { _done = true; return; }
            _leaf_i = 0;
        } else {
            _leaf_i++;
        }
        if (byte() != 0xFF) return;  // 0xFF = no voxel (cluster.hpp:29-32)
    }
}
Leaf LeafCursor::leaf(float sdf_res, float sdf_trunc) const { return make_leaf(key(), byte(), sdf_res, sdf_trunc); }

TSDFMap::LeafRange TSDFMap::leaves(size_t submap) {
    flush();
    if (submap >= submap_count()) throw std::runtime_error("chad::TSDFMap::leaves: no such submap");
    return LeafRange{node_levels(), submap_roots(submap)[0], _sdf_res, _sdf_trunc};
}
std::vector<Leaf> TSDFMap::collect_leaves(size_t submap) {
    if (_group) flush();
    size_t n = 0;
    check(_ctx, chad_iterate_leaves(_ctx, (uint32_t)submap, nullptr, nullptr, 0, &n), "collect_leaves");
    std::vector<uint64_t> keys(n);
    std::vector<uint8_t> bytes(n);
    if (n) check(_ctx, chad_iterate_leaves(_ctx, (uint32_t)submap, keys.data(), bytes.data(), n, &n), "collect_leaves");
    std::vector<Leaf> out;
    out.reserve(n);
    for (size_t i = 0; i < n; i++) out.push_back(make_leaf(keys[i], bytes[i], _sdf_res, _sdf_trunc));
    return out;
}

// ---- .grid file ------------------------------------------------------------------------------------------
void TSDFMap::save_grid(const std::string& filename, size_t submap) {
    finalize_active();  // tsdf.cpp:78-81
    if (submap >= submap_count()) throw std::runtime_error("chad::TSDFMap::save_grid: no such submap");
    write_grid(node_levels(), submap_roots(submap)[0], _sdf_res, _sdf_trunc, filename);
}

// File format of lvr2's ChadGrid::saveGrid (lvr2.cpp:170-200): the header float (the truncation distance, under the name voxel_res:
// SURVEY.md section 9 Q15), the counts, one (x, y, z, sd) query point per voxel and eight query-point indices per complete cell. The
// content follows from the leaf iterator: voxels in ascending Morton order are the query points; voxel v is corner j of the cell at
// v + corner_offset[j] (the corner numbering of lvr2.cpp:88-98); a cell with a missing corner is dropped (lvr2.cpp:115-129).
void write_grid(const HostNodeLevels& levels, uint32_t root, float sdf_res, float sdf_trunc, const std::string& filename) {
    struct QueryPoint { float x, y, z, sd; };
    constexpr uint32_t NO_POINT = 0xFFFFFFFFu;  // lvr2 FastBox::INVALID_INDEX
    static const int32_t corner_offset[8][3] = { {0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {0, -1, 0}, {0, 0, -1}, {-1, 0, -1}, {-1, -1, -1}, {0, -1, -1} };
    std::vector<QueryPoint> points;
    std::unordered_map<uint64_t, std::array<uint32_t, 8>> cells;
    for (LeafCursor it(levels, root); !it.done(); it.next()) {
        const Leaf v = it.leaf(sdf_res, sdf_trunc);
        int32_t vx, vy, vz;
        chad_morton_decode(v.morton, &vx, &vy, &vz);
        const uint32_t index = (uint32_t)points.size();
        points.push_back({v.x, v.y, v.z, v.signed_distance});
        for (size_t j = 0; j < 8; j++) {
            auto [cell, fresh] = cells.try_emplace(chad_morton_encode(vx + corner_offset[j][0], vy + corner_offset[j][1], vz + corner_offset[j][2]));
            if (fresh) cell->second.fill(NO_POINT);
            cell->second[j] = index;
        }
    }
    std::vector<uint64_t> complete;
    for (const auto& [code, corners] : cells)
        if (std::find(corners.begin(), corners.end(), NO_POINT) == corners.end()) complete.push_back(code);
    std::sort(complete.begin(), complete.end());
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("chad::write_grid: cannot open " + filename);
    auto put = [&](const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("chad::write_grid: write failed"); } };
    const size_t nq = points.size(), nc = complete.size();
    put(&sdf_trunc, sizeof(float));
    put(&nq, sizeof(size_t));
    put(&nc, sizeof(size_t));
    put(points.data(), nq * sizeof(QueryPoint));
    for (uint64_t code : complete) put(cells[code].data(), 8 * sizeof(uint32_t));
    std::fclose(f);
}

// ---- ray cast --------------------------------------------------------------------------------------------
namespace {
// Root-to-leaf walk for one voxel that also says how large the hole is where the voxel does not exist: `empty_shift` = s means that
// the whole aligned cube of 2^s voxels per axis around the voxel is absent (s = 20 - depth for a missing child of a depth-`depth`
// node, 1 for a missing leaf cluster, 0 for a missing voxel of an existing cluster).
uint8_t descend(const HostNodeLevels& levels, uint32_t root, uint64_t key, uint32_t& empty_shift) {
    uint32_t addr = root;
    for (uint32_t depth = 0; depth + 1 < HostNodeLevels::MAX_DEPTH; depth++) {
        addr = levels.get_child_addr(depth, addr, uint8_t((key >> (3 * (20 - depth))) & 7));
        if (addr == 0) { empty_shift = 20 - depth; return 0xFF; }
    }
    uint64_t cluster;
    if (!levels.try_get_lc(addr, uint8_t((key >> 3) & 7), cluster)) { empty_shift = 1; return 0xFF; }
    empty_shift = 0;
    return uint8_t(cluster >> (8 * (key & 7)));
}
}  // namespace

RayHit raycast(const HostNodeLevels& levels, uint32_t root, const std::array<float, 3>& origin, const std::array<float, 3>& direction,
               float max_distance, float sdf_res, float sdf_trunc, std::vector<Leaf>* along) {
    const double len = std::sqrt(double(direction[0]) * direction[0] + double(direction[1]) * direction[1] + double(direction[2]) * direction[2]);
    if (!(len > 0.0)) throw std::invalid_argument("chad::raycast: zero direction");
    if (!(sdf_res > 0.0f)) throw std::invalid_argument("chad::raycast: voxel size must be positive");
    RayHit out;
    if (root == 0 || root >= levels.nodes[0].size()) return out;
    constexpr double INF = std::numeric_limits<double>::infinity();
    constexpr int64_t BIAS = 1 << 20;  // morton.hpp:21-28: 21 bits per axis around 2^20
    const double res = double(sdf_res), limit = double(max_distance);
    double d[3], t_next[3], t_step[3];
    int64_t v[3];
    int step[3];
    for (int a = 0; a < 3; a++) {
        d[a] = double(direction[a]) / len;
        v[a] = (int64_t)std::floor(double(origin[a]) / res);
        step[a] = d[a] > 0.0 ? 1 : (d[a] < 0.0 ? -1 : 0);
        t_step[a] = step[a] ? res / std::fabs(d[a]) : INF;
        t_next[a] = step[a] ? (double(v[a] + (step[a] > 0 ? 1 : 0)) * res - double(origin[a])) / d[a] : INF;  // where the ray leaves the voxel along a
    }
    bool have_prev = false, in_hole = false;
    double t_prev = 0.0, sd_prev = 0.0, t_enter = 0.0;
    Leaf leaf_prev{};
    int64_t hole[3] = {0, 0, 0};
    uint32_t hole_shift = 0;
    while (t_enter < limit) {
        if (v[0] < -BIAS || v[0] >= BIAS || v[1] < -BIAS || v[1] >= BIAS || v[2] < -BIAS || v[2] >= BIAS) break;  // outside the key space
        const int axis = t_next[0] < t_next[1] ? (t_next[0] < t_next[2] ? 0 : 2) : (t_next[1] < t_next[2] ? 1 : 2);
        const double t_exit = t_next[axis];
        out.voxels_walked++;
        if (in_hole && (((v[0] + BIAS) >> hole_shift) != hole[0] || ((v[1] + BIAS) >> hole_shift) != hole[1] || ((v[2] + BIAS) >> hole_shift) != hole[2])) in_hole = false;
        if (!in_hole) {
            const uint64_t key = chad_morton_encode((int32_t)v[0], (int32_t)v[1], (int32_t)v[2]);
            uint32_t shift;
            const uint8_t byte = descend(levels, root, key, shift);
            out.tree_descents++;
            if (byte == 0xFF) {
                if (shift) {  // an absent octant: no descents until the ray has left it
                    in_hole = true;
                    hole_shift = shift;
                    for (int a = 0; a < 3; a++) hole[a] = (v[a] + BIAS) >> shift;
                }
            } else {
                out.voxels_found++;
                const Leaf leaf = make_leaf(key, byte, sdf_res, sdf_trunc);
                if (along) along->push_back(leaf);
                // the distance was measured at the voxel's lower corner (octree.hpp:157): the sample sits where that corner projects onto the ray
                const double t = (double(v[0]) * res - double(origin[0])) * d[0] + (double(v[1]) * res - double(origin[1])) * d[1] + (double(v[2]) * res - double(origin[2])) * d[2];
                const double sd = double(leaf.signed_distance);
                if (have_prev && sd_prev > 0.0 && sd <= 0.0) {
                    const double t_hit = std::max(0.0, t_prev + (t - t_prev) * (sd_prev / (sd_prev - sd)));
                    if (t_hit <= limit) {
                        out.hit = true;
                        out.distance = float(t_hit);
                        out.x = float(double(origin[0]) + t_hit * d[0]);
                        out.y = float(double(origin[1]) + t_hit * d[1]);
                        out.z = float(double(origin[2]) + t_hit * d[2]);
                        out.before = leaf_prev;
                        out.after = leaf;
                    }
                    return out;
                }
                have_prev = true;
                t_prev = t;
                sd_prev = sd;
                leaf_prev = leaf;
            }
        }
        t_enter = t_exit;
        t_next[axis] += t_step[axis];
        v[axis] += step[axis];
    }
    return out;
}

// ---- DAG readers ---------------------------------------------------------------------------------------
uint32_t HostNodeLevels::get_child_addr(uint32_t depth, uint32_t parent_addr, uint8_t child_i) const {
    const std::vector<uint32_t>& level = nodes[depth];
    if (parent_addr == 0 || parent_addr >= level.size()) return 0;
    const uint32_t child_mask = level[parent_addr];
    const uint32_t child_bit = 1u << child_i;
    if (!(child_mask & child_bit)) return 0;
    const size_t at = size_t(parent_addr) + 1 + (size_t)std::popcount(uint8_t(child_mask & (child_bit - 1)));
    return at < level.size() ? level[at] : 0;
}
bool HostNodeLevels::try_get_lc(uint32_t parent_addr, uint8_t child_i, uint64_t& cluster) const {
    const uint32_t addr = get_child_addr(MAX_DEPTH - 1, parent_addr, child_i);
    if (addr == 0 || addr >= leaf_clusters.size()) return false;
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
bool HostNodeLevels::consistent(std::string* why) const {
    auto no = [&](const std::string& w) { if (why) *why = w; return false; };
    if (leaf_clusters.empty()) return no("the cluster level lacks its reserved word 0");
    for (size_t d = 0; d < MAX_DEPTH; d++) {
        const std::vector<uint32_t>& level = nodes[d];
        if (level.empty()) return no("node level " + std::to_string(d) + " lacks its reserved word 0");
        const size_t below = d + 1 < MAX_DEPTH ? nodes[d + 1].size() : leaf_clusters.size();
        size_t a = 1;
        while (a < level.size()) {
            const size_t n = (size_t)std::popcount(uint8_t(level[a] & 0xFF));
            if (a + n >= level.size()) return no("a record of node level " + std::to_string(d) + " runs past the end of the level");
            for (size_t q = 1; q <= n; q++)
                if (level[a + q] == 0 || level[a + q] >= below) return no("a record of node level " + std::to_string(d) + " points outside the level below");
            a += 1 + n;
        }
        if (a != level.size()) return no("node level " + std::to_string(d) + " does not end on a record boundary");
    }
    return true;
}

SavedMap load_dag(const std::string& filename) {
    std::FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) throw std::runtime_error("chad::load_dag: cannot open " + filename);
    auto fail = [&](const std::string& what) { std::fclose(f); throw std::runtime_error("chad::load_dag: " + filename + ": " + what); };
    auto get = [&](void* p, size_t n) { if (n && std::fread(p, 1, n, f) != n) fail("truncated file"); };
    std::fseek(f, 0, SEEK_END);
    const uint64_t file_bytes = (uint64_t)std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    char magic[8];
    get(magic, 8);
    const bool v2 = std::memcmp(magic, "CHADDAG2", 8) == 0;
    if (!v2 && std::memcmp(magic, "CHADDAG1", 8) != 0) fail("not a CHADDAG file");
    SavedMap m;
    m.has_counters = v2;
    uint32_t n_sub = 0;
    get(&m.sdf_res, 4);
    get(&m.sdf_trunc, 4);
    get(&n_sub, 4);
    if (uint64_t(n_sub) * 8 > file_bytes) fail("submap count exceeds the file size");
    m.roots.resize(n_sub);
    m.positions.resize(n_sub);
    for (uint32_t i = 0; i < n_sub; i++) {
        get(m.roots[i].data(), 8);
        if (v2) {
            uint32_t np = 0;
            get(&np, 4);
            if (uint64_t(np) * 12 > file_bytes) fail("pose count exceeds the file size");
            m.positions[i].resize(np);
            if (np) get(m.positions[i].data(), size_t(np) * 12);
        }
    }
    for (size_t lv = 0; lv <= HostNodeLevels::MAX_DEPTH; lv++) {
        if (v2) { get(&m.uniques[lv], 4); get(&m.dupes[lv], 4); }
        uint64_t n = 0;
        get(&n, 8);
        if (lv < HostNodeLevels::MAX_DEPTH) {
            if (n > file_bytes / 4) fail("level size exceeds the file size");
            m.levels.nodes[lv].resize(n);
            get(m.levels.nodes[lv].data(), n * 4);
        } else {
            if (n > file_bytes / 8) fail("cluster count exceeds the file size");
            m.levels.leaf_clusters.resize(n);
            get(m.levels.leaf_clusters.data(), n * 8);
        }
    }
    if ((uint64_t)std::ftell(f) != file_bytes) fail("trailing bytes");
    std::string why;
    if (!m.levels.consistent(&why)) fail("inconsistent DAG: " + why);
    for (const auto& r : m.roots)
        for (uint32_t a : r)
            if (a == 0 || a >= m.levels.nodes[0].size()) fail("a submap root lies outside the root level");
    std::fclose(f);
    return m;
}

void* pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (chad_host_alloc(bytes, &p) != CHAD_OK || !p) throw std::bad_alloc();
    return p;
}
void pinned_free(void* p) noexcept { chad_host_free(p); }
}  // namespace chad
