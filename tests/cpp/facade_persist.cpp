// SURVEY.md section 8f through the C++ class: save -> load -> continue (the reference has no persistence: its save() is terminal,
// /root/reference/src/chad/tsdf.cpp:76-86), the leaf iterator the reference sketches (include/chad/tsdf.hpp:120-161) on the host
// copy and on the device, and inserts from page-locked vectors (no staging copy; the reference copies every overload,
// tsdf.hpp:53,62,81-87,106-112). Needs a GPU. Prints one line per check; exit code 0 = all hold.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <random>

#include "chad/tsdf.hpp"

#if !__has_include(<glm/vec3.hpp>)
#error "compile with -I tests/cpp/shims (or real glm)"
#endif

static std::vector<char> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<char>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main() {
    const size_t n = 30000;
    std::mt19937 gen(11);
    std::uniform_real_distribution<double> dis(-1.0, 1.0);
    auto wall = [&](float cx) {  // a noisy wall patch 4 m in front of a sensor at (cx, 0, 0)
        std::vector<std::array<float, 3>> v(n);
        for (auto& p : v) p = { float(cx + 4.0 + 0.01 * dis(gen)), float(2.0 * dis(gen)), float(1.0 * dis(gen)) };
        return v;
    };
    const auto s1 = wall(0.0f), s2 = wall(7.0f), s3 = wall(14.0f);  // 7 m apart: every scan opens a new submap (tsdf.cpp:52)
    bool ok = true;
    {   // ---- save -> load -> continue == never stopped ----
        { chad::TSDFMap a; a.insert(s1, { 0.0f, 0.0f, 0.0f }); a.insert(s2, { 7.0f, 0.0f, 0.0f }); a.save("persist_a.chad"); }
        { chad::TSDFMap b; b.load("persist_a.chad"); b.insert(s3, { 14.0f, 0.0f, 0.0f }); b.save("persist_b.chad"); }
        { chad::TSDFMap c; c.insert(s1, { 0.0f, 0.0f, 0.0f }); c.insert(s2, { 7.0f, 0.0f, 0.0f }); c.save("persist_c0.chad");
          c.insert(s3, { 14.0f, 0.0f, 0.0f }); c.save("persist_c.chad"); }
        const auto fb = slurp("persist_b.chad"), fc = slurp("persist_c.chad");
        const bool same = !fb.empty() && fb == fc;
        const chad::SavedMap m = chad::load_dag("persist_b.chad");
        const bool shape = m.roots.size() == 3 && m.positions.size() == 3 && m.positions[2].size() == 1 && m.positions[2][0][0] == 14.0f && m.has_counters;
        std::printf("restore: continued map == uninterrupted map: %s (%zu bytes, %zu submaps)\n", same && shape ? "yes" : "NO", fb.size(), m.roots.size());
        ok = ok && same && shape;
    }
    {   // ---- leaf iterator: host cursor == device iterator; positions are voxel corners, distances inside the band ----
        chad::TSDFMap m;
        m.insert(s1, { 0.0f, 0.0f, 0.0f });
        m.save("persist_d.chad");
        const auto dev = m.collect_leaves(0);
        const auto range = m.leaves(0);
        size_t i = 0;
        bool same = true, sane = true;
        for (const chad::Leaf v : range) {
            if (i >= dev.size() || dev[i].morton != v.morton || dev[i].quantised != v.quantised || dev[i].x != v.x || dev[i].signed_distance != v.signed_distance) same = false;
            if (i > 0 && dev[i - 1].morton >= v.morton) same = false;  // ascending
            if (!(v.signed_distance >= -m._sdf_trunc && v.signed_distance <= m._sdf_trunc) || !(v.x > 3.5f && v.x < 4.5f)) sane = false;
            i++;
        }
        same = same && i == dev.size() && i > 1000;
        std::printf("leaves: host cursor == device iterator: %s (%zu voxels), plausible: %s\n", same ? "yes" : "NO", i, sane ? "yes" : "NO");
        ok = ok && same && sane;
    }
    {   // ---- page-locked vectors: same map, no staging copy ----
        std::vector<std::array<float, 3>, chad::pinned_allocator<std::array<float, 3>>> pa(s1.begin(), s1.end());
        std::vector<glm::vec3, chad::pinned_allocator<glm::vec3>> pg;
        for (const auto& p : s1) pg.push_back(glm::vec3(p[0], p[1], p[2]));
        chad::TSDFMap x, y, z;
        x.insert(s1, { 0.0f, 0.0f, 0.0f });
        y.insert(pa, { 0.0f, 0.0f, 0.0f });
        z.insert(pg, glm::vec3(0.0f, 0.0f, 0.0f));
        x.save("persist_x.chad"); y.save("persist_y.chad"); z.save("persist_z.chad");
        const auto fx = slurp("persist_x.chad");
        const bool same = !fx.empty() && fx == slurp("persist_y.chad") && fx == slurp("persist_z.chad");
        std::printf("pinned vectors: identical map: %s\n", same ? "yes" : "NO");
        ok = ok && same;
    }
    return ok ? 0 : 1;
}
