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
        // NodeLevels::get_child_addr (levels.hpp:147-161)
        uint32_t get_child_addr(uint32_t depth, uint32_t parent_addr, uint8_t child_i) const;
        // NodeLevels::try_get_lc (levels.hpp:177-192)
        bool try_get_lc(uint32_t parent_addr, uint8_t child_i, uint64_t& cluster) const;
        // Single-voxel read through the tree under `root_addr` (a submap's root_addr_tsdf or root_addr_weight): the root-to-leaf
        // walk the reference's readers make with get_child_addr / try_get_lc, for one Morton key (morton.hpp:21-28). Returns the
        // voxel's quantised byte (cluster.hpp:13-32; decode (byte - 127) / 127 * sdf_trunc), 0xFF where the voxel does not exist.
        // Host counterpart of chad_query_voxels.
        uint8_t query(uint32_t root_addr, uint64_t morton_key) const;
    };

    // What TSDFMap::save wrote (flat CHADDAG1 dump): map parameters, the roots of every finalised
    // submap (submap.hpp:108-109) and the host copy of the DAG. Pure host code: no GPU needed to read a map back.
    struct SavedMap {
        float sdf_res = 0.0f, sdf_trunc = 0.0f;
        std::vector<std::array<uint32_t, 2>> roots;  // per submap: root_addr_tsdf, root_addr_weight
        HostNodeLevels levels;
    };
    SavedMap load_dag(const std::string& filename);  // throws std::runtime_error on a malformed or truncated file
    // The reference's ChadGrid constructor + saveGrid (lvr2.cpp:32-130,170-200) for the tree under `root_addr_tsdf`: what
    // TSDFMap::save_grid writes, usable on a loaded map as well (pure host code).
    void write_grid(const HostNodeLevels& levels, uint32_t root_addr_tsdf, float sdf_res, float sdf_trunc, const std::string& filename);

    class TSDFMap {
    public:
        TSDFMap(const TSDFMap&) = delete;
        TSDFMap(TSDFMap&&) = delete;
        TSDFMap& operator=(const TSDFMap&) = delete;
        TSDFMap& operator=(TSDFMap&&) = delete;

        // initialize a TSDF map with the given voxel size and truncation distance (tsdf.hpp:29)
        TSDFMap(float sdf_res = 0.05f, float sdf_trunc = 0.1f);
        TSDFMap(float sdf_res, float sdf_trunc, const std::vector<std::array<float, 3>>& points, const std::array<float, 3>& position)
            : TSDFMap(sdf_res, sdf_trunc) { insert(points, position); }
        TSDFMap(float sdf_res, float sdf_trunc, const float* points_p, size_t points_count, const float* position_p)
            : TSDFMap(sdf_res, sdf_trunc) { insert(points_p, points_count, position_p); }
        TSDFMap(float sdf_res, float sdf_trunc, const float* points_p, size_t points_count, float position_x, float position_y, float position_z)
            : TSDFMap(sdf_res, sdf_trunc) { insert(points_p, points_count, position_x, position_y, position_z); }
        ~TSDFMap();

        // insert pointcloud alongside scanner position (tsdf.hpp:48)
        void insert(const std::vector<std::array<float, 3>>& points, const std::array<float, 3>& position) {
            insert(points.empty() ? nullptr : points[0].data(), points.size(), position.data());
        }
        // insert pointcloud as a raw array of repeating x,y,z coordinates (tsdf.hpp:50,59)
        void insert(const float* points_p, size_t points_count, const float* position_p);
        void insert(const float* points_p, size_t points_count, float position_x, float position_y, float position_z) {
            const float pos[3] = { position_x, position_y, position_z };
            insert(points_p, points_count, pos);
        }
        #if __has_include(<glm/vec3.hpp>)
            TSDFMap(float sdf_res, float sdf_trunc, const std::vector<glm::vec3>& points, const glm::vec3& position)
                : TSDFMap(sdf_res, sdf_trunc) { insert(points, position); }
            void insert(const std::vector<glm::vec3>& points, const glm::vec3& position) {  // tsdf.hpp:75-89
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
            TSDFMap(float sdf_res, float sdf_trunc, const std::vector<Eigen::Vector3f>& points, const Eigen::Vector3f& position)
                : TSDFMap(sdf_res, sdf_trunc) { insert(points, position); }
            void insert(const std::vector<Eigen::Vector3f>& points, const Eigen::Vector3f& position) {  // tsdf.hpp:100-114
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
        void flush();                                   // wait for queued inserts
        size_t submap_count();
        std::array<uint32_t, 2> submap_roots(size_t i); // root_addr_tsdf, root_addr_weight (submap.hpp:108-109)
        HostNodeLevels node_levels();                   // device -> host copy of the whole DAG
        chad_ctx* handle() { return _ctx; }

    public:
        const float _sdf_res;
        const float _sdf_trunc;

    private:
        chad_ctx* _ctx;
    };
}
