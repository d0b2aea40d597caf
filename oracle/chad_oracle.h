/* TEST INFRASTRUCTURE ONLY (oracle/). C-ABI of the CPU restatement of the reference's
 * TSDFMap::insert / Submap::finalize path; see chad_oracle.c for the file:line citations.
 * Parity status: PINNED -- validated bit-for-bit against the reference's own sources compiled
 * here (oracle/_ref, "stable" variant; tests/test_oracle_vs_reference.py) and against the golden
 * hashes in tests/golden/ that were generated from that reference build (tests/golden/make_golden.py).
 * Never linked into, imported by, or called from the product path. */
#ifndef CHAD_ORACLE_H
#define CHAD_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct oracle_map oracle_map;

oracle_map* oracle_create(float sdf_res, float sdf_trunc);
void oracle_destroy(oracle_map* m);
/* returns the number of submaps finalised so far */
uint32_t oracle_insert(oracle_map* m, const float* xyz, size_t n, const float* pos);
uint32_t oracle_finalize_active(oracle_map* m);
size_t oracle_voxel_count(const oracle_map* m);
size_t oracle_export_voxels(const oracle_map* m, uint64_t* keys, uint32_t* sd_bits, uint32_t* w);
uint32_t oracle_submap_count(const oracle_map* m);
void oracle_submap_roots(const oracle_map* m, uint32_t i, uint32_t* tsdf, uint32_t* weight);
size_t oracle_level_words(const oracle_map* m, int level);
void oracle_level_counters(const oracle_map* m, int level, uint32_t* uniques, uint32_t* dupes);
void oracle_export_level(const oracle_map* m, int level, void* dst);
/* workload properties of the LAST insert: U = emitted (voxel, sd) updates,
 * V_scan = distinct voxels touched by that scan */
void oracle_last_scan_stats(const oracle_map* m, uint64_t* updates, uint64_t* distinct_voxels);

/* stages, for kernel-by-kernel parity */
void oracle_stage_points(const float* xyz, size_t n, const float* pos, float sdf_res, float* xyz_sorted,
                         uint64_t* keys, uint32_t* order, float* normals);
/* upper bound-free two-call protocol: pass keys == NULL to get the count */
size_t oracle_stage_pairs(const float* xyz_sorted, const float* normals, size_t n, const float* pos, float sdf_res,
                          float sdf_trunc, uint64_t* keys, float* sd, uint32_t* counts);
uint64_t oracle_morton_encode(int32_t x, int32_t y, int32_t z);
void oracle_morton_decode(uint64_t key, int32_t* x, int32_t* y, int32_t* z);
uint64_t oracle_quantise_cluster(const float* sd, uint32_t present_mask, float sdf_trunc);
#ifdef __cplusplus
}
#endif
#endif
