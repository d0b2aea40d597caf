// Every insert() overload and construct-and-insert constructor of the reference's public class
// (/root/reference/include/chad/tsdf.hpp:29-114, usage /root/reference/README.md:32-53) against the drop-in header:
// std::array, raw pointer + pose pointer, raw pointer + pose scalars, glm::vec3, Eigen::Vector3f. The glm / Eigen
// overloads are __has_include-gated like the reference's; tests/cpp/shims provides minimal stand-ins so that they are
// compiled here. Run on a GPU box, the program also checks that all overloads build the identical map.
#include <cstdio>
#include <random>

#include "chad/tsdf.hpp"

#if !__has_include(<glm/vec3.hpp>) || !__has_include(<Eigen/Eigen>)
#error "compile with -I tests/cpp/shims (or real glm / Eigen) so that the gated overloads are part of the test"
#endif

int main() {
    const size_t n = 20000;
    std::vector<std::array<float, 3>> arr(n);
    std::vector<glm::vec3> gv(n);
    std::vector<Eigen::Vector3f> ev(n);
    std::vector<float> raw(3 * n);
    std::mt19937 gen(7);
    std::uniform_real_distribution<double> dis(-1.0, 1.0);
    for (size_t i = 0; i < n; i++) {
        const float x = float(4.0 + 0.5 * dis(gen)), y = float(3.0 * dis(gen)), z = float(1.5 * dis(gen));
        arr[i] = { x, y, z };
        gv[i] = glm::vec3(x, y, z);
        ev[i] = Eigen::Vector3f(x, y, z);
        raw[3 * i] = x; raw[3 * i + 1] = y; raw[3 * i + 2] = z;
    }
    const float pose[3] = { 0.1f, -0.2f, 0.3f };
    auto roots_of = [](chad::TSDFMap& m) { m.save("facade_overloads.chad"); return m.submap_roots(0); };
    std::vector<std::array<uint32_t, 2>> roots;
    std::vector<size_t> words;
    auto record = [&](chad::TSDFMap& m) {
        roots.push_back(roots_of(m));
        const auto lv = m.node_levels();
        size_t w = lv.leaf_clusters.size();
        for (const auto& l : lv.nodes) w += l.size();
        words.push_back(w);
    };
    { chad::TSDFMap m; m.insert(arr, { pose[0], pose[1], pose[2] }); record(m); }                       // tsdf.hpp:48
    { chad::TSDFMap m; m.insert(raw.data(), n, pose); record(m); }                                        // tsdf.hpp:50
    { chad::TSDFMap m; m.insert(raw.data(), n, pose[0], pose[1], pose[2]); record(m); }                   // tsdf.hpp:59
    { chad::TSDFMap m; m.insert(gv, glm::vec3(pose[0], pose[1], pose[2])); record(m); }                   // tsdf.hpp:75
    { chad::TSDFMap m; m.insert(ev, Eigen::Vector3f(pose[0], pose[1], pose[2])); record(m); }             // tsdf.hpp:100
    { chad::TSDFMap m(0.05f, 0.1f, arr, { pose[0], pose[1], pose[2] }); record(m); }                      // tsdf.hpp:31
    { chad::TSDFMap m(0.05f, 0.1f, raw.data(), n, pose); record(m); }                                     // tsdf.hpp:33
    { chad::TSDFMap m(0.05f, 0.1f, raw.data(), n, pose[0], pose[1], pose[2]); record(m); }                // tsdf.hpp:36
    { chad::TSDFMap m(0.05f, 0.1f, gv, glm::vec3(pose[0], pose[1], pose[2])); record(m); }                // tsdf.hpp:70
    { chad::TSDFMap m(0.05f, 0.1f, ev, Eigen::Vector3f(pose[0], pose[1], pose[2])); record(m); }          // tsdf.hpp:95
    bool same = true;
    for (size_t i = 1; i < roots.size(); i++) same = same && roots[i] == roots[0] && words[i] == words[0];
    std::printf("facade_overloads: %zu maps, roots (%u, %u), %zu DAG words, identical: %s\n", roots.size(), roots[0][0], roots[0][1], words[0],
                same ? "yes" : "NO");
    return same && words[0] > 100 ? 0 : 1;
}
