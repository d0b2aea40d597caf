// The reference's usage pattern (/root/reference/README.md:32-53, src/chad/main.cpp:7-38) against the drop-in
// chad::TSDFMap: points on a 5 m sphere, one insert at the origin, save(); then walks the DAG like the reference's
// mesh exporter does (src/chad/detail/lvr2.cpp:32-113) and checks decoded distances against the analytic sphere.
#include <cmath>
#include <cstdio>
#include <random>

#include "chad/tsdf.hpp"

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? std::stoul(argv[1]) : 200000;
    std::vector<std::array<float, 3>> points(n);
    std::mt19937 gen(420);
    std::uniform_real_distribution<double> dis(-1.0, 1.0);
    for (auto& p : points) {
        double x = dis(gen), y = dis(gen), z = dis(gen);
        const double inv = 5.0 / std::sqrt(x * x + y * y + z * z);
        p = { float(x * inv), float(y * inv), float(z * inv) };
    }
    chad::TSDFMap map{ 0.05f, 0.1f };
    map.insert(points, { 0.0f, 0.0f, 0.0f });
    map.save("facade_demo.chad");
    map.save_grid("facade_demo.grid");  // the reference's hashgrid.grid (lvr2.cpp:170-200) for the first submap
    const auto levels = map.node_levels();
    const auto roots = map.submap_roots(0);
    // depth-first walk of the TSDF tree
    size_t leaves = 0, bad = 0;
    struct Frame { uint32_t addr; uint32_t child; int32_t x, y, z; };
    std::vector<Frame> stack{ { roots[0], 0, 0, 0, 0 } };
    while (!stack.empty()) {
        const size_t depth = stack.size() - 1;
        Frame& f = stack.back();
        if (f.child == 8) { stack.pop_back(); continue; }
        const uint8_t ci = uint8_t(f.child++);
        const int32_t half = 1 << (20 - depth);  // voxel extent of a child at this depth
        const int32_t cx = f.x + ((ci >> 0) & 1) * half, cy = f.y + ((ci >> 1) & 1) * half, cz = f.z + ((ci >> 2) & 1) * half;
        if (depth < chad::HostNodeLevels::MAX_DEPTH - 1) {
            const uint32_t child = levels.get_child_addr((uint32_t)depth, f.addr, ci);
            if (child) stack.push_back({ child, 0, cx, cy, cz });
        } else {
            uint64_t lc;
            if (!levels.try_get_lc(f.addr, ci, lc)) continue;
            for (int s = 0; s < 8; s++) {
                const uint32_t byte = (lc >> (8 * s)) & 0xFF;
                if (byte == 0xFF) continue;
                const float sd = (float(byte) - 127.0f) * (1.0f / 127.0f) * 0.1f;  // cluster.hpp:46-50
                const float vx = float(cx + ((s >> 0) & 1) - (1 << 20)) * 0.05f, vy = float(cy + ((s >> 1) & 1) - (1 << 20)) * 0.05f,
                            vz = float(cz + ((s >> 2) & 1) - (1 << 20)) * 0.05f;
                const float analytic = 5.0f - std::sqrt(vx * vx + vy * vy + vz * vz);  // inward-facing normals: positive inside
                const float clamped = std::fmax(-0.1f, std::fmin(0.1f, analytic));
                leaves++;
                if (std::fabs(sd - clamped) > 0.03f) bad++;
            }
        }
    }
    std::printf("facade_demo: %zu points, %zu submaps, roots (%u, %u), %zu leaves, %zu off the analytic sphere by > 3 cm\n", n, map.submap_count(),
                roots[0], roots[1], leaves, bad);
    return (leaves > 0 && bad * 50 < leaves) ? 0 : 1;
}
