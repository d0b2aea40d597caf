// chad::TSDFMap -- the reference's public class (/root/reference/include/chad/tsdf.hpp:21-171),
// re-implemented on top of the B200 C ABI (include/chad_b200.h). Same constructor arguments,
// insert() overloads (std::array / raw pointer / glm / Eigen), save(), public _sdf_res/_sdf_trunc,
// deleted copy/move. Link against libchad_b200.so. Differences, all documented in INTEGRATION.md:
//   * insert() queues device work and returns; results are complete after flush()/save();
//   * the construct-and-insert constructors work (the reference's dereference uninitialised
//     pointers);
//   * save() finalises the active submap and writes the DAG in the flat format of
//     INTEGRATION.md; LVR2 meshing stays a host-side consumer of that DAG (out of scope here);
//   * errors surface as std::runtime_error instead of undefined behaviour.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#if __has_include(<glm/vec3.hpp>)
#   include <glm/vec3.hpp>
#endif
#if __has_include(<Eigen/Eigen>)
#   include <Eigen/Eigen>
#endif

struct chad_ctx;

namespace chad {
    // Host copy of the DAG with the layout of detail::NodeLevels (levels.hpp:146-200): 20 node levels of u32
    // words [mask, children...] addressed by word offset, and one level of 64-bit leaf clusters.
    struct HostNodeLevels {
        static constexpr uint64_t MAX_DEPTH = 20;
        std::array<std::vector<uint32_t>, MAX_DEPTH> nodes;  // NodeLevel::_raw_data[0.._occupied_n)
        std::vector<uint64_t> leaf_clusters;                 // LeafClusterLevel::_raw_data[0.._uniques_n]
        // NodeLevels::get_child_addr (levels.hpp:147-161); 0 for an absent child AND for an address outside the level (a corrupt file
        // must not turn into an out-of-bounds read)
        uint32_t get_child_addr(uint32_t depth, uint32_t parent_addr, uint8_t child_i) const;
        // NodeLevels::try_get_lc (levels.hpp:177-192)
        bool try_get_lc(uint32_t parent_addr, uint8_t child_i, uint64_t& cluster) const;
        // Single-voxel read through the tree under `root_addr` (a submap's root_addr_tsdf or root_addr_weight): the root-to-leaf
        // walk the reference's readers make with get_child_addr / try_get_lc, for one Morton key (morton.hpp:21-28). Returns the
        // voxel's quantised byte (cluster.hpp:13-32; decode (byte - 127) / 127 * sdf_trunc), 0xFF where the voxel does not exist.
        // Host counterpart of chad_query_voxels.
        uint8_t query(uint32_t root_addr, uint64_t morton_key) const;
        // Every record of every level lies inside its level and points at addresses inside the level below; every node level ends on
        // a record boundary. load_dag checks this once, so the readers above never see an address they cannot follow.
        bool consistent(std::string* why = nullptr) const;
    };

    // One voxel as a reader sees it (the reference's iterator TODO, tsdf.hpp:120-123: "real floating point position per leaf",
    // "access to the data written there", "users should not be exposed to morton codes"): the voxel's lower corner in metres
    // (octree.hpp:157 measures the signed distance to that corner) and the decoded signed distance (cluster.hpp:46-50).
    struct Leaf {
        float x, y, z;
        float signed_distance;
        uint8_t quantised;      // the stored byte, 0 .. 254
        uint64_t morton;        // for callers that do want the key
    };

    // Forward cursor over the voxels of the tree under `root_addr`, in ascending Morton order -- the leaf iterator the reference
    // sketches (tsdf.hpp:125-155; tsdf.cpp:88-159 only reaches the first leaf cluster). The DAG must outlive the cursor.
    class LeafCursor {
    public:
        LeafCursor(const HostNodeLevels& levels, uint32_t root_addr);
        bool done() const { return _done; }
        void next();                       // to the next existing voxel
        uint64_t key() const { return (_cluster_key << 3) | _leaf_i; }
        uint8_t byte() const { return uint8_t(_cluster >> (8 * _leaf_i)); }
        Leaf leaf(float sdf_res, float sdf_trunc) const;
    private:
        bool next_cluster();
        const HostNodeLevels* _levels;
        std::array<uint8_t, HostNodeLevels::MAX_DEPTH> _child{};   // per depth: the next child index to try
        std::array<uint32_t, HostNodeLevels::MAX_DEPTH> _addr{};
        uint32_t _depth = 0;
        uint64_t _cluster = 0, _cluster_key = 0;
        uint32_t _leaf_i = 0;
        bool _done = false;
    };

    // A ray through the tree under `root_addr_tsdf` -- the second reader on the reference's list (tsdf.hpp:157-160: "raycast to retrieve
    // leaves along it + physics hit"). The ray is walked voxel by voxel (Amanatides-Woo, the walk octree.hpp:120-152 makes for the
    // truncation band, here in double precision from `origin` to `max_distance`); octants the tree does not hold are crossed without
    // descending into them. The hit is the first place where the signed distance goes from > 0 (in front of the surface) to <= 0, its
    // distance interpolated linearly between the two samples (a voxel's distance was measured at its lower corner, octree.hpp:157: the
    // sample sits where that corner projects onto the ray).
    struct RayHit {
        bool hit = false;
        float distance = 0.0f;             // metres along the ray
        float x = 0.0f, y = 0.0f, z = 0.0f;  // origin + distance * direction / |direction|
        Leaf before{}, after{};            // the voxels around the crossing (before: signed distance > 0, after: <= 0)
        size_t voxels_walked = 0;          // voxels the ray went through
        size_t voxels_found = 0;           // ... of which the tree holds this many
        size_t tree_descents = 0;          // root-to-leaf walks made (one per found voxel and one per empty octant entered)
    };
    // `along` (optional) receives every voxel of the tree the ray goes through, in ray order, up to and including `after`.
    // Throws std::invalid_argument for a zero direction or a non-positive voxel size.
    RayHit raycast(const HostNodeLevels& levels, uint32_t root_addr_tsdf, const std::array<float, 3>& origin, const std::array<float, 3>& direction,
                   float max_distance, float sdf_res, float sdf_trunc, std::vector<Leaf>* along = nullptr);

    // What TSDFMap::save writes (flat CHADDAG2 dump): map parameters, per finalised submap its roots
    // (submap.hpp:108-109) and poses (submap.hpp:110), the host copy of the DAG and the levels' dedup counters (levels.hpp:90-91,141).
    // Pure host code: no GPU needed to read a map back (CHADDAG1 files of round 1 -- no poses, no counters -- are still read).
    struct SavedMap {
        float sdf_res = 0.0f, sdf_trunc = 0.0f;
        std::vector<std::array<uint32_t, 2>> roots;  // per submap: root_addr_tsdf, root_addr_weight
        std::vector<std::vector<std::array<float, 3>>> positions;  // per submap: the poses of its scans
        std::array<uint32_t, 21> uniques{}, dupes{};
        bool has_counters = false;
        HostNodeLevels levels;
    };
    SavedMap load_dag(const std::string& filename);  // throws std::runtime_error on a malformed, truncated or inconsistent file
    void save_dag(const SavedMap& map, const std::string& filename);
    // The file LVR2's ChadGrid::saveGrid writes for the tree under `root_addr_tsdf` (format: lvr2.cpp:170-200; one query point per
    // voxel, one cell per voxel corner whose eight voxels all exist): what TSDFMap::save_grid writes, usable on a loaded map as well.
    void write_grid(const HostNodeLevels& levels, uint32_t root_addr_tsdf, float sdf_res, float sdf_trunc, const std::string& filename);

    // Page-locked host memory for point buffers: a std::vector<glm::vec3, chad::pinned_allocator<glm::vec3>> (or Eigen::Vector3f,
    // std::array<float, 3>, float) handed to TSDFMap::insert is DMA'd to the device straight from the caller's storage -- no staging
    // copy (the reference copies every overload into a std::vector first: tsdf.hpp:53,62,81-87,106-112).
    void* pinned_alloc(size_t bytes);
    void pinned_free(void* p) noexcept;
    template <typename T> struct pinned_allocator {
        using value_type = T;
        pinned_allocator() = default;
        template <typename U> pinned_allocator(const pinned_allocator<U>&) noexcept {}
        T* allocate(size_t n) { return static_cast<T*>(pinned_alloc(n * sizeof(T))); }
        void deallocate(T* p, size_t) noexcept { pinned_free(p); }
        template <typename U> bool operator==(const pinned_allocator<U>&) const noexcept { return true; }
        template <typename U> bool operator!=(const pinned_allocator<U>&) const noexcept { return false; }
    };

    class TSDFMap {
    public:
        TSDFMap(const TSDFMap&) = delete;
        TSDFMap(TSDFMap&&) = delete;
        TSDFMap& operator=(const TSDFMap&) = delete;
        TSDFMap& operator=(TSDFMap&&) = delete;

        // initialize a TSDF map with the given voxel size and truncation distance (tsdf.hpp:29)
        TSDFMap(float sdf_res = 0.05f, float sdf_trunc = 0.1f);
        template <typename A>
        TSDFMap(float sdf_res, float sdf_trunc, const std::vector<std::array<float, 3>, A>& points, const std::array<float, 3>& position)
            : TSDFMap(sdf_res, sdf_trunc) { insert(points, position); }
        TSDFMap(float sdf_res, float sdf_trunc, const float* points_p, size_t points_count, const float* position_p)
            : TSDFMap(sdf_res, sdf_trunc) { insert(points_p, points_count, position_p); }
        TSDFMap(float sdf_res, float sdf_trunc, const float* points_p, size_t points_count, float position_x, float position_y, float position_z)
            : TSDFMap(sdf_res, sdf_trunc) { insert(points_p, points_count, position_x, position_y, position_z); }
        ~TSDFMap();

        // insert pointcloud alongside scanner position (tsdf.hpp:48)
        // (any allocator: with chad::pinned_allocator the points are read by the GPU's copy engine where they lie)
        template <typename A>
        void insert(const std::vector<std::array<float, 3>, A>& points, const std::array<float, 3>& position) {
            insert(points.empty() ? nullptr : points[0].data(), points.size(), position.data());
        }
        // insert pointcloud as a raw array of repeating x,y,z coordinates (tsdf.hpp:50,59)
        void insert(const float* points_p, size_t points_count, const float* position_p);
        void insert(const float* points_p, size_t points_count, float position_x, float position_y, float position_z) {
            const float pos[3] = { position_x, position_y, position_z };
            insert(points_p, points_count, pos);
        }
        #if __has_include(<glm/vec3.hpp>)
            template <typename A>
            TSDFMap(float sdf_res, float sdf_trunc, const std::vector<glm::vec3, A>& points, const glm::vec3& position)
                : TSDFMap(sdf_res, sdf_trunc) { insert(points, position); }
            template <typename A>
            void insert(const std::vector<glm::vec3, A>& points, const glm::vec3& position) {  // tsdf.hpp:75-89
                if (sizeof(glm::vec3) == 12) insert(points.empty() ? nullptr : &points[0].x, points.size(), position.x, position.y, position.z);
                else {
                    std::vector<std::array<float, 3>> v;
                    v.reserve(points.size());
                    for (const auto& p: points) v.push_back({ p.x, p.y, p.z });
                    insert(v, { position.x, position.y, position.z });
                }
            }
        #endif
        #if __has_include(<Eigen/Eigen>)
            template <typename A>
            TSDFMap(float sdf_res, float sdf_trunc, const std::vector<Eigen::Vector3f, A>& points, const Eigen::Vector3f& position)
                : TSDFMap(sdf_res, sdf_trunc) { insert(points, position); }
            template <typename A>
            void insert(const std::vector<Eigen::Vector3f, A>& points, const Eigen::Vector3f& position) {  // tsdf.hpp:100-114
                if (sizeof(Eigen::Vector3f) == 12) insert(points.empty() ? nullptr : points[0].data(), points.size(), position.x(), position.y(), position.z());
                else {
                    std::vector<std::array<float, 3>> v;
                    v.reserve(points.size());
                    for (const auto& p: points) v.push_back({ p.x(), p.y(), p.z() });
                    insert(v, { position.x(), position.y(), position.z() });
                }
            }
        #endif

        // finalize the active submap and write the DAG to disk (tsdf.hpp:118; see the header comment)
        void save(const std::string& filename);

        // ---- additions (not in the reference) ----
        // The reference's save() hands the first submap to LVR2 as a "ChadGrid" (lvr2.cpp:32-113) whose saveGrid writes a
        // `.grid` file (lvr2.cpp:170-200). LVR2 is out of scope here, but the grid itself only needs the DAG: this writes the
        // same file for finalised submap `submap` (query points in the reference's traversal order; complete cells in
        // ascending Morton order of the cell -- the reference's order is its unordered_map's iteration order).
        void save_grid(const std::string& filename, size_t submap = 0);
        // Continue a saved map: read a file written by save() into this (empty) map; the dedup sets are rebuilt, so submaps inserted
        // afterwards land in the DAG exactly where an uninterrupted run would have put them.
        void load(const std::string& filename);
        // Leaf iterator (tsdf.hpp:120-161): the voxels of finalised submap `submap` in ascending Morton order. leaves() walks a host
        // copy of the DAG with a LeafCursor; collect_leaves() runs the data-parallel iterator on the device (chad_iterate_leaves).
        struct LeafRange {
            struct iterator {
                const LeafRange* range;
                LeafCursor cursor;
                Leaf operator*() const { return cursor.leaf(range->sdf_res, range->sdf_trunc); }
                iterator& operator++() { cursor.next(); return *this; }
                bool operator!=(const iterator& other) const { return cursor.done() != other.cursor.done() || (!cursor.done() && cursor.key() != other.cursor.key()); }
                bool operator==(const iterator& other) const { return !(*this != other); }
            };
            HostNodeLevels levels;
            uint32_t root = 0;
            float sdf_res = 0.0f, sdf_trunc = 0.0f;
            iterator begin() const { return iterator{this, LeafCursor(levels, root)}; }
            iterator end() const { return iterator{this, LeafCursor(levels, 0)}; }
        };
        LeafRange leaves(size_t submap);
        std::vector<Leaf> collect_leaves(size_t submap);
        std::vector<std::array<float, 3>> submap_positions(size_t submap);  // Submap::positions (submap.hpp:110)
        void flush();                                   // wait for queued inserts
        size_t submap_count();
        std::array<uint32_t, 2> submap_roots(size_t i); // root_addr_tsdf, root_addr_weight (submap.hpp:108-109)
        HostNodeLevels node_levels();                   // device -> host copy of the whole DAG
        chad_ctx* handle() { return _ctx; }

    public:
        const float _sdf_res;
        const float _sdf_trunc;

    private:
        // CHAD_DEVICES=0,1,... (environment, read by the constructor): ONE map cut into Morton ranges over those GPUs (chad_create_sharded,
        // one worker thread per GPU; insert / flush / save are forwarded to every rank, reads go to rank 0, which holds the DAG). Unset or
        // a single ordinal: the plain single-GPU map on CHAD_DEVICE (default 0).
        struct Group;
        Group* _group = nullptr;
        void collective(const char* what, int (*call)(chad_ctx*, const void*), const void* arg);
        void finalize_active();
        chad_ctx* _ctx;   // the map (rank 0's shard when sharded)
    };
}
