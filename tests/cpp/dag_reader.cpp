// Reads a CHADDAG1 file back with chad::load_dag (pure host code, no GPU) and answers single-voxel queries through
// chad::HostNodeLevels::query -- the walk the reference's readers make (levels.hpp:147-192).
// usage: dag_reader <file.chad> <submap> <keys.u64> <out.u8>   -> prints "res trunc n_submaps root_tsdf root_weight n_keys"
//        dag_reader grid <file.chad> <submap> <out.grid>      -> the reference's hashgrid.grid (lvr2.cpp:170-200) of that submap
//        dag_reader leaves <file.chad> <submap> <keys.u64> <bytes.u8>  -> every voxel of the submap through chad::LeafCursor
//        dag_reader resave <in.chad> <out.chad>               -> load_dag + save_dag (CHADDAG2)
//        dag_reader ray <file.chad> <submap> <rays.f32> <out.f64> <along.u64>  -> chad::raycast per ray (7 floats: origin, direction, max distance);
//                   per ray 9 doubles: hit, distance, x, y, z, voxels walked, voxels found, tree descents, voxels in `along`
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>

#include "chad/tsdf.hpp"

int main(int argc, char** argv) {
    if (argc == 5 && std::string(argv[1]) == "grid") {  // dag_reader grid <file.chad> <submap> <out.grid>
        try {
            const chad::SavedMap m = chad::load_dag(argv[2]);
            const size_t submap = std::strtoul(argv[3], nullptr, 10);
            if (submap >= m.roots.size()) { std::fprintf(stderr, "no such submap\n"); return 3; }
            chad::write_grid(m.levels, m.roots[submap][0], m.sdf_res, m.sdf_trunc, argv[4]);
            return 0;
        } catch (const std::exception& e) {
            std::fprintf(stderr, "%s\n", e.what());
            return 1;
        }
    }
    if (argc == 6 && std::string(argv[1]) == "leaves") {  // dag_reader leaves <file.chad> <submap> <out_keys.u64> <out_bytes.u8>: the leaf iterator
        try {
            const chad::SavedMap m = chad::load_dag(argv[2]);
            const size_t submap = std::strtoul(argv[3], nullptr, 10);
            if (submap >= m.roots.size()) { std::fprintf(stderr, "no such submap\n"); return 3; }
            std::vector<uint64_t> keys;
            std::vector<uint8_t> bytes;
            double checksum = 0.0;  // also exercises Leaf: position and decoded distance
            for (chad::LeafCursor it(m.levels, m.roots[submap][0]); !it.done(); it.next()) {
                keys.push_back(it.key());
                bytes.push_back(it.byte());
                const chad::Leaf v = it.leaf(m.sdf_res, m.sdf_trunc);
                checksum += double(v.x) + double(v.y) + double(v.z) + double(v.signed_distance);
            }
            std::FILE* fk = std::fopen(argv[4], "wb");
            std::FILE* fb = std::fopen(argv[5], "wb");
            if (!fk || !fb) return 5;
            if (!keys.empty() && (std::fwrite(keys.data(), 8, keys.size(), fk) != keys.size() || std::fwrite(bytes.data(), 1, bytes.size(), fb) != bytes.size())) return 5;
            std::fclose(fk);
            std::fclose(fb);
            std::printf("%zu %.9g %zu\n", keys.size(), checksum, m.positions[submap].size());
            return 0;
        } catch (const std::exception& e) {
            std::fprintf(stderr, "%s\n", e.what());
            return 1;
        }
    }
    if (argc == 7 && std::string(argv[1]) == "ray") {
        try {
            const chad::SavedMap m = chad::load_dag(argv[2]);
            const size_t submap = std::strtoul(argv[3], nullptr, 10);
            if (submap >= m.roots.size()) { std::fprintf(stderr, "no such submap\n"); return 3; }
            std::FILE* fr = std::fopen(argv[4], "rb");
            if (!fr) return 4;
            std::fseek(fr, 0, SEEK_END);
            const size_t n = (size_t)std::ftell(fr) / 28;
            std::fseek(fr, 0, SEEK_SET);
            std::vector<float> rays(n * 7);
            if (n && std::fread(rays.data(), 28, n, fr) != n) return 4;
            std::fclose(fr);
            std::vector<double> out;
            std::vector<uint64_t> keys;
            for (size_t i = 0; i < n; i++) {
                const float* r = &rays[i * 7];
                std::vector<chad::Leaf> along;
                const chad::RayHit h = chad::raycast(m.levels, m.roots[submap][0], {r[0], r[1], r[2]}, {r[3], r[4], r[5]}, r[6], m.sdf_res, m.sdf_trunc, &along);
                for (const chad::Leaf& v : along) keys.push_back(v.morton);
                for (double x : {double(h.hit), double(h.distance), double(h.x), double(h.y), double(h.z), double(h.voxels_walked), double(h.voxels_found),
                                 double(h.tree_descents), double(along.size())}) out.push_back(x);
                if (h.hit && (along.empty() || along.back().morton != h.after.morton || !(h.before.signed_distance > 0.0f) || h.after.signed_distance > 0.0f)) return 6;
            }
            std::FILE* fo = std::fopen(argv[5], "wb");
            std::FILE* fk = std::fopen(argv[6], "wb");
            if (!fo || !fk) return 5;
            if (!out.empty() && std::fwrite(out.data(), 8, out.size(), fo) != out.size()) return 5;
            if (!keys.empty() && std::fwrite(keys.data(), 8, keys.size(), fk) != keys.size()) return 5;
            std::fclose(fo);
            std::fclose(fk);
            return 0;
        } catch (const std::exception& e) {
            std::fprintf(stderr, "%s\n", e.what());
            return 1;
        }
    }
    if (argc == 4 && std::string(argv[1]) == "resave") {  // dag_reader resave <in.chad> <out.chad>: load_dag -> save_dag
        try {
            chad::SavedMap m = chad::load_dag(argv[2]);
            chad::save_dag(m, argv[3]);
            std::printf("%d %zu\n", m.has_counters ? 1 : 0, m.roots.size());
            return 0;
        } catch (const std::exception& e) {
            std::fprintf(stderr, "%s\n", e.what());
            return 1;
        }
    }
    if (argc != 5) { std::fprintf(stderr, "usage: dag_reader <file.chad> <submap> <keys.u64> <out.u8>\n"); return 2; }
    try {
        const chad::SavedMap m = chad::load_dag(argv[1]);
        const size_t submap = std::strtoul(argv[2], nullptr, 10);
        if (submap >= m.roots.size()) { std::fprintf(stderr, "no such submap\n"); return 3; }
        std::FILE* fk = std::fopen(argv[3], "rb");
        if (!fk) return 4;
        std::fseek(fk, 0, SEEK_END);
        const size_t n = (size_t)std::ftell(fk) / 8;
        std::fseek(fk, 0, SEEK_SET);
        std::vector<uint64_t> keys(n);
        if (n && std::fread(keys.data(), 8, n, fk) != n) return 4;
        std::fclose(fk);
        std::vector<uint8_t> out(n);
        for (size_t i = 0; i < n; i++) out[i] = m.levels.query(m.roots[submap][0], keys[i]);
        std::FILE* fo = std::fopen(argv[4], "wb");
        if (!fo || (n && std::fwrite(out.data(), 1, n, fo) != n)) return 5;
        std::fclose(fo);
        std::printf("%.9g %.9g %zu %u %u %zu\n", m.sdf_res, m.sdf_trunc, m.roots.size(), m.roots[submap][0], m.roots[submap][1], n);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
