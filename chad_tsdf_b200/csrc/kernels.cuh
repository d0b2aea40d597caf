// Host-callable launchers of the hot-path kernels. Each returns the number of kernels it queued.
#pragma once
#include "common.cuh"

namespace chadgpu {

struct MapParams {
    float res;          // sdf_res
    float trunc;        // sdf_trunc
    float recip;        // float(1.0 / double(res))   (morton.hpp:63, octree.hpp:82)
    float trunc_recip;  // 1.0f / trunc               (submap.hpp:24)
    u32 band_margin;    // voxels a band voxel can lie from its point's voxel (bounds the plan's k)
    u32 max_ray_voxels; // per-ray capacity bound used to size the pair buffers
    u32 max_ray_runs;   // 8^3-voxel blocks one ray can touch
};

// instrumentation classes (chad_profile_*): one per kernel (group); radix passes are numbered
enum ProfClass {
    PC_PLAN = 0, PC_POINT_KEYS, PC_POINT_SORT_HIST, PC_POINT_SORT_PASS0, PC_POINT_GATHER = PC_POINT_SORT_PASS0 + 8, PC_NORMALS, PC_BAND_COUNT,
    PC_BAND_SCAN, PC_BAND_EMIT, PC_PAIR_SORT_HIST, PC_PAIR_SORT_PASS0, PC_SEGMENT_COUNT = PC_PAIR_SORT_PASS0 + 8, PC_FOLD, PC_FINALIZE,
    PC_BLOCKS_COUNT, PC_BLOCKS_SCAN, PC_BLOCKS_EMIT, PC_BLOCKS_SORT, PC_RUNS_EMIT, PC_RUNS_SORT, PC_RUNS_FOLD, PC_SHARD_EXCHANGE, PC_COUNT
};

// ---- points.cu: voxelise + Morton (morton.hpp:59-80), sort keys, gather, normals (normals.hpp) ----
// tsb = bits of a 256-ray tile index inside the batch's largest scan (run descriptor order key, see points.cuh)
int launch_plan(cudaStream_t s, const float* xyz, u32 n_points, u32 n_scans, const MapParams& mp, BatchPlan* plan, u32 tsb);
int launch_point_keys(cudaStream_t s, const float* xyz, u32 n_points, const BatchScans* scans, const MapParams& mp, const BatchPlan* plan,
                      u64* sortkeys, u32* index);
// keys_a/keys_b, idx_a/idx_b: the two radix buffers; the sorted result is picked with plan->nbits_points
int launch_point_gather(cudaStream_t s, const float* xyz, u32 n_points, const BatchPlan* plan, const u64* keys_a, const u64* keys_b,
                        const u32* idx_a, const u32* idx_b, u64* sorted_keys, u32* sorted_order, float* xyz_sorted);
int launch_normals(cudaStream_t s, const float* xyz_sorted, const u64* sorted_keys, u32 n_points, const BatchScans* scans,
                   const BatchPlan* plan, u32* seg_info, float* normals);
// full Morton keys of the sorted points (for the stage API): expands the compact sort keys
int launch_point_full_keys(cudaStream_t s, const u64* sorted_keys, u32 n_points, const BatchPlan* plan, u64* full_keys);
int launch_morton_encode(cudaStream_t s, const i32* voxels, u32 n, u64* keys);

// ---- band.cu: truncation-band DDA (octree.hpp:86-159) ----
int launch_band_count(cudaStream_t s, const float* xyz_sorted, u32 n_points, const BatchScans* scans, const MapParams& mp,
                      const BatchPlan* plan, u32* counts);
int launch_band_emit(cudaStream_t s, const float* xyz_sorted, const float* normals, u32 n_points, const BatchScans* scans,
                     const MapParams& mp, BatchPlan* plan, const u32* offsets, u64* pair_keys, u32* pair_sd, u32 pair_capacity,
                     bool full_keys);

// ---- blocks.cu: block-binned grouping of the band-voxel updates (default pair path) ----
struct BlockTable {   // open addressing over 8^3-voxel block ids (Morton key >> 9), rebuilt per batch
    u64* keys;        // [capacity], ~0 = empty
    u32* count;       // updates per block
    u32* cursor;      // emit cursor
    u32* offset;      // exclusive prefix sum of count
    u32* list;        // slots of the non-empty blocks
    u32 capacity;     // power of two
};
size_t blocks_table_bytes(u32 capacity);
BlockTable blocks_table_carve(void* mem, u32 capacity);
cudaError_t blocks_init();
u32 blocks_max_batch_points();
int launch_blocks_pairs(cudaStream_t s, const float* xyz_sorted, const float* normals, u32 n_points, const BatchScans* scans, const MapParams& mp,
                        BatchPlan* plan, const BlockTable& bt, void* scan_ws, u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, u32 pair_capacity,
                        int num_sms, const LaunchHook* hook, int cls_count, int cls_scan, int cls_emit, int cls_sort);

// ---- shard.cu: Morton-range sharding of one map across GPUs, point-stage side ----
constexpr int SHARD_WORLD_MAX = 8;
struct RadixWorkspace;
struct ShardFilter {    // work arrays of the ownership filter
    u32* scan_own;      // [MAX_BATCH_SCANS + 1] points of scan s this rank owns (zero between batches)
    u32* scan_lower;    // [MAX_BATCH_SCANS + 1] points of scan s owned by lower ranks
    u32* tile_cnt;      // [max_tiles] owned points per 256-point tile of the batch -> their exclusive prefix
    u32* own_bits;      // [max_tiles * 8] per warp of 32 consecutive points: which of them this rank owns
    u32 max_tiles;
};
size_t shard_filter_bytes(size_t max_points);
ShardFilter shard_filter_carve(void* mem, size_t max_points);
int launch_plan_reset(cudaStream_t s, BatchPlan* plan, u32 n_points, u32 n_scans);
// splitters[0 .. world] (block ids = Morton key >> 9; [0] = 0, [world] = ~0) from a sorted sample of the submap's first scan
int launch_shard_splitters(cudaStream_t s, const float* xyz_first_scan, u32 n_first_scan, const MapParams& mp, u32 world, u32 first_share_256,
                           u64* splitters);
// plan + ownership filter of a batch (replaces launch_plan + launch_point_keys): plan->n_points = owned points, own_scans = their scan table
int launch_shard_filter(cudaStream_t s, const float* xyz, u32 n_points, u32 n_scans, const BatchScans* scans, const MapParams& mp, BatchPlan* plan,
                        u32 tsb, u32 gbits, const u64* splitters, u32 rank, u32 world, const ShardFilter& f, BatchScans* own_scans, u64* sortkeys,
                        u32* index);

// ---- runs.cu: tile-run grouping with the fold fused into the per-block sort (default pair path) ----
struct RunBuffers {   // run descriptors of a batch: sort ping-pong (block id, descriptor index), descriptor = (first record, records)
    u64 *key_a, *key_b;
    u32 *val_a, *val_b;
    uint2* desc;
    uint2* sdesc;     // descriptors in sorted (block id) order
    uint2* work;      // (first sorted descriptor, descriptors) of every block
    u32 capacity;
};
cudaError_t runs_init();
u32 runs_max_batch_points();
u32 runs_max_ray_voxels();
u32 runs_max_ray_runs();
size_t runs_desc_bytes(size_t capacity);
RunBuffers runs_carve(void* mem, size_t capacity);
struct ChunkTable;
u32 runs_max_tiles(u32 n_points, u32 n_scans);
size_t runs_max_runs(u32 n_points, u32 n_scans);
int launch_runs_emit(cudaStream_t s, const float* xyz_sorted, const float* normals, u32 n_points, u32 n_scans, const BatchScans* scans,
                     const MapParams& mp, BatchPlan* plan, const RunBuffers& rb, u64* records, u32 rec_capacity, u32 order_rank,
                     const LaunchHook* hook, int cls_emit);
int launch_runs_group(cudaStream_t s, size_t max_runs, BatchPlan* plan, const RunBuffers& rb, const RadixWorkspace& rws, int num_sms,
                      const LaunchHook* hook, int cls_sort);
int launch_runs_fold(cudaStream_t s, const u64* records, const RunBuffers& rb, BatchPlan* plan, const ChunkTable& t, int num_sms);
// Morton-range sharding, walk side: the runs of blocks another rank owns are moved into that rank's exchange box, the boxes received
// from the other ranks are appended to this rank's records / descriptors (between launch_runs_emit and launch_runs_group).
// A box is `words` u64: [0] = runs | records << 32, [1] = flags, then (key, first | length << 32) per run from the front and the
// records from the back.
struct ShardBoxes {
    u64* out;   // [world][words] one box per destination (the own one is unused)
    u64* in;    // [world][words] one box per source
    u32 words;
};
int launch_runs_pack(cudaStream_t s, size_t max_runs, BatchPlan* plan, const RunBuffers& rb, u64* records, const u64* splitters, u32 rank, u32 world,
                     const ShardBoxes& boxes, int num_sms);
int launch_runs_ingest(cudaStream_t s, BatchPlan* plan, const RunBuffers& rb, u64* records, u32 rec_capacity, u32 rank, u32 world, const ShardBoxes& boxes);

// ---- fold.cu: ordered segmented fold (octree.hpp:161-163) into the resident chunk table ----
struct ChunkTable {
    u64* keys;      // [capacity] chunk key = voxel key >> 3, EMPTY = ~0
    uint2* cells;   // [capacity][8] (sd bits, weight) per voxel slot, weight 0 = absent
    u64 capacity;   // power of two
    u32* count;     // device counter of occupied chunks
    u64* list;      // [capacity / 2] keys of the occupied chunks in insertion order (count entries): finalize reads the
                    // submap's chunks from here instead of scanning the table
};
constexpr u64 CHUNK_EMPTY = ~0ull;
int launch_table_clear(cudaStream_t s, const ChunkTable& t);
int launch_segment_count(cudaStream_t s, const u64* keys_a, const u64* keys_b, u32 max_pairs, BatchPlan* plan, int num_sms);
int launch_fold(cudaStream_t s, const u64* keys_a, const u64* keys_b, const u32* sd_a, const u32* sd_b, u32 max_pairs, BatchPlan* plan,
                const ChunkTable& t, int num_sms);
int launch_table_rehash(cudaStream_t s, const ChunkTable& from, const ChunkTable& to, int num_sms);
// the table's chunk list -> chunk sort keys on *d_nbits bits (arbitrary order); *d_count = number of chunks
int launch_table_compact(cudaStream_t s, const ChunkTable& t, u32 max_n, u64* out_keys, u32* out_vals, u32* d_count, u32* d_rmax, u32* d_nbits, int num_sms);
// sorted chunk keys -> contiguous (full chunk key, 8 x (sd bits, weight)); count and result buffer selected on the device
int launch_chunk_gather(cudaStream_t s, const ChunkTable& t, const u64* keys_a, const u64* keys_b, const u32* d_count, const u32* d_rmax,
                        const u32* d_nbits, u32 max_n, u64* out_keys, void* out_cells);

// ---- dag.cu: Submap::finalize (submap.hpp:10-106) level by level ----
struct DedupTable {   // open addressing; entry = (hash tag << 32) | ref, 0 = empty
    u64* entries;
    u32* first;       // per slot: min sequence index of a pending record, 0xFFFFFFFF when idle
    u64 capacity;     // power of two
};
constexpr u32 REF_PENDING = 0x80000000u;
// the 20 node levels in one persistent kernel (dag_levels_kernel)
struct LevelDev { u64* entries; u32* first; u64 capacity; u32* raw; };
struct LevelCounters { u32 occupied, uniques, dupes, pad; };   // device mirror of NodeLevel::_occupied_n / _uniques_n / _dupes_n
struct LevelsArgs {
    LevelDev lv[20];
    LevelCounters* counters;   // [21]
    const u32* d_chunks;       // leaf clusters of the submap
    u64* ids[2];               // node ids of the current children, ping-pong (ids[0] = sorted chunk ids on entry)
    u32* addr[2];              // (TSDF, weight) addresses of the current children (addr[0] = cluster addresses on entry)
    u32* head_rank; u32* cand; u32* slot_of; u64* rank;
    u64* partial;              // [2][1024]
    u32* bar;                  // grid barrier counter
    u32* d_error; u32* root_out; u32* level_nodes;
    u32 solo;                  // set by launch_dag_levels
};
int launch_dag_levels(cudaStream_t s, const LevelsArgs& args, int num_sms);
// device-side DAG reader: bytes of the queried voxels in the TSDF tree rooted at `root` (0xFF = absent)
struct DagReadArgs { const u32* raw[20]; const u64* clusters; };
int launch_dag_query(cudaStream_t s, const DagReadArgs& args, u32 root, const u64* keys, u32 n, u8* out);
int launch_dedup_clear(cudaStream_t s, const DedupTable& t);
// rebuild a level's dedup set from its stored records (restore of a saved map): cluster level: raw = u64 values, addresses 1 .. n_records;
// node level: raw = u32 words, starts[i] = address of record i
int launch_dedup_restore(cudaStream_t s, const DedupTable& t, bool cluster, const void* raw, const u32* starts, u32 n_records);
// leaf iterator over a finalised submap's tree, one depth per call (see dag.cu)
int launch_iter_expand(cudaStream_t s, const u32* raw, const u32* addr, const u64* prefix, const u32* d_n, u32 capacity, u32* counts, u32* offsets,
                       void* scan_ws, u32* addr_next, u64* prefix_next, u32* d_n_next, u32* d_overflow);
int launch_iter_leaves(cudaStream_t s, const u64* clusters, const u32* addr, const u64* prefix, const u32* d_n, u32 capacity, u32* counts, u32* offsets,
                       void* scan_ws, u32 out_capacity, u64* keys, u8* bytes, u32* d_total);
int launch_dedup_rehash(cudaStream_t s, const DedupTable& from, const DedupTable& to, int num_sms);
// the chunk count is read from device memory (*d_chunks <= max_chunks): finalize part 1 runs without a host round trip
int launch_cluster_build(cudaStream_t s, const void* gathered_cells, const u32* d_chunks, u32 max_chunks, const MapParams& mp, u64* tsdf_values);
// cluster level: sequence tsdf_0, W, tsdf_1, W, ...; addr_out[2i] / addr_out[2i+1] = tsdf / weight cluster address of chunk i
int launch_cluster_dedup(cudaStream_t s, const DedupTable& t, const u64* tsdf_values, const u32* d_chunks, u32 max_chunks, u64* raw, u32 uniques_before,
                         u32* slot_of, u32* is_new, u32* rank, void* scan_ws, u32* addr_out, u32* d_new_count, u32* d_error);

}  // namespace chadgpu
