/* chad_b200.h -- C ABI of the B200-native (sm_100a) implementation of the chad_tsdf integration
 * hot path. This is the drop-in boundary: plain pointers and sizes, no
 * C++/torch types. The reference has no FFI of its own; its only operator boundary is the C++
 * class chad::TSDFMap (/root/reference/include/chad/tsdf.hpp:21-171), whose implementation
 * (/root/reference/src/chad/tsdf.cpp:27-86) this library replaces. include/chad/tsdf.hpp in
 * this repo is the same class re-implemented on top of these entry points.
 *
 * Every function returns CHAD_OK (0) or a negative error code; chad_last_error() gives the
 * message. All host pointers are plain (pageable or pinned) memory owned by the caller unless
 * a parameter is explicitly called a device pointer. A context is bound to one CUDA device and
 * is not thread-safe (like the reference, SURVEY.md section 8b "Threading").
 *
 * There is NO CPU fallback: without a CUDA device chad_create fails with CHAD_ERR_CUDA. */
#ifndef CHAD_B200_H
#define CHAD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHAD_OK 0
#define CHAD_ERR_INVALID (-1)  /* bad argument */
#define CHAD_ERR_CUDA (-2)     /* CUDA runtime error (no device, launch failure, out of memory) */
#define CHAD_ERR_RANGE (-3)    /* a voxel coordinate left the 21-bit Morton range (morton.hpp:21-28) or the batch key budget */
#define CHAD_ERR_CAPACITY (-4) /* a device table or arena could not hold the data */
#define CHAD_ERR_NUMERIC (-5)  /* NaN/Inf input coordinates */

#define CHAD_LEVEL_CLUSTERS 20 /* level index of the leaf-cluster level (levels.hpp:146-200) */
#define CHAD_NUM_LEVELS 21

typedef struct chad_ctx chad_ctx;

/* Counters since creation (or the last chad_reset_stats). `updates` is U and `scan_voxels` is the
 * sum over insert calls of V_scan; scan_voxels is only exact while
 * max_batch_scans == 1 (with larger batches it counts distinct voxels per batch). */
typedef struct chad_stats {
    uint64_t scans;          /* insert calls */
    uint64_t points;         /* sum of N */
    uint64_t updates;        /* sum of emitted (voxel, sd) updates, U */
    uint64_t scan_voxels;    /* sum of distinct voxels per processed batch */
    uint64_t batches;        /* device batches processed */
    uint64_t submaps;        /* submaps finalised */
    uint64_t kernel_launches;/* CUDA kernels launched by this library */
    uint64_t h2d_bytes;      /* bytes copied host -> device for inserts */
    uint64_t d2h_bytes;      /* bytes copied device -> host on the insert path (batch plans, counters, finalize scalars) */
    uint64_t key_bits_points;/* sort-key width of the last batch's point sort (ceil(/8) = active radix passes) */
    uint64_t key_bits_pairs; /* sort-key width of the last batch's voxel-update sort */
    uint64_t resident_clusters; /* leaf chunks (2x2x2 voxels) in the active submap's table (as of last flush) */
} chad_stats;

/* Device memory behind the map and how it has grown. The reference reserves 352 GiB of virtual memory and lets the page faults do the
 * growing (virtual_array.hpp:12-38, levels.hpp:50,117; never freed: levels.hpp:92,142); here the DAG arenas, their dedup sets and the
 * chunk tables are device buffers that double when they fill up -- on the finalize stream, without synchronising the insert streams. */
typedef struct chad_memory {
    uint64_t dag_words_bytes;        /* bytes of the 21 levels actually in use (what save() writes) */
    uint64_t dag_arena_bytes;        /* bytes reserved for them */
    uint64_t dedup_bytes;            /* bytes of the 21 dedup sets */
    uint64_t dedup_records;          /* records in them (sum of the levels' unique counts) */
    uint64_t dedup_max_load_permille;/* fullest dedup set: records / slots x 1000 (kept <= 500) */
    uint64_t chunk_table_bytes;      /* the two resident chunk tables (active submap, submap being finalised) */
    uint64_t batch_buffer_bytes;     /* per-batch work buffers (points, update records) */
    uint64_t grow_events;            /* buffer doublings since creation (DAG arenas, dedup sets, chunk tables) */
    uint64_t grow_bytes;             /* bytes allocated by them */
    uint64_t grow_host_us;           /* host time spent in them (cudaMalloc + waits) */
    uint64_t device_used_bytes, device_total_bytes; /* cudaMemGetInfo */
} chad_memory;
int chad_memory_info(chad_ctx* ctx, chad_memory* out);

/* ---- lifetime ------------------------------------------------------------------------- */
/* Replaces TSDFMap::TSDFMap(float sdf_res, float sdf_trunc) (tsdf.cpp:27-31). `device` is the CUDA
 * ordinal. `max_batch_scans` >= 1: how many consecutive scans of one submap may be fused into one
 * device batch (results are identical for every value; 0 selects the default). */
int chad_create(float sdf_res, float sdf_trunc, int device, int max_batch_scans, chad_ctx** out);
/* Replaces TSDFMap::~TSDFMap (tsdf.cpp:32-38). */
void chad_destroy(chad_ctx* ctx);
/* Message of the last error on `ctx` (or of the last failed chad_create when ctx == NULL). */
const char* chad_last_error(const chad_ctx* ctx);

/* ---- the hot path --------------------------------------------------------------------- */
/* Replaces TSDFMap::insert(const std::vector<std::array<float,3>>&, const std::array<float,3>&)
 * (tsdf.cpp:39-75) and its raw-pointer overloads (tsdf.hpp:50-64). `xyz` = n x 3 floats (AoS) in
 * HOST memory; it is copied before the call returns, so the caller may reuse it (same ownership
 * rule as the reference, which copies at tsdf.hpp:53,62). The device work is queued and may still
 * be running when the call returns; the submap switch rule (> 5 m from the submap's first pose,
 * tsdf.cpp:51-58) is applied here. */
int chad_insert(chad_ctx* ctx, const float* xyz, size_t n, const float position[3]);
/* Same for PAGE-LOCKED host memory, without waiting for the copy: the transfers of consecutive scans then follow each other at
 * the link's full rate (a caller that recycles one buffer per scan needs chad_insert). `xyz` must stay valid and unchanged until
 * the next chad_flush (or any other synchronising call). Fails with CHAD_ERR_INVALID for pageable memory. */
int chad_insert_async(chad_ctx* ctx, const float* xyz, size_t n, const float position[3]);
/* `count` consecutive chad_insert (wait_for_copy != 0) or chad_insert_async (== 0) calls behind one entry point: xyz[i] = points of scan
 * i, n[i] their number, positions = count x 3 floats. For callers whose own loop is slow (Python); a C++ caller just loops. */
int chad_insert_many(chad_ctx* ctx, const float* const* xyz, const size_t* n, const float* positions, size_t count, int wait_for_copy);
/* Same, but `xyz_device` already lives in this context's device memory (no host copy). */
int chad_insert_device(chad_ctx* ctx, const float* xyz_device, size_t n, const float position[3]);
/* Wait until every queued insert has been applied; reports deferred device errors. */
int chad_flush(chad_ctx* ctx);
/* Replaces the part of TSDFMap::save before meshing (tsdf.cpp:78-81): Submap::finalize of the
 * active submap (submap.hpp:10-106) into the global DAG. A fresh active submap is started
 * afterwards (the reference's save() is terminal). No-op if the active
 * submap has no pose yet. */
int chad_finalize_active(chad_ctx* ctx);

/* ---- state export (parity tests, save()) ---------------------------------------------- */
int chad_submap_count(chad_ctx* ctx, uint32_t* count);
/* Submap::root_addr_tsdf / root_addr_weight (submap.hpp:108-109) of finalised submap i. */
int chad_submap_roots(chad_ctx* ctx, uint32_t i, uint32_t* root_tsdf, uint32_t* root_weight);
/* Octree leaves of the ACTIVE submap (octree.hpp:15) as parallel arrays in ascending Morton-key
 * order. Two-call protocol: chad_voxel_count, then chad_export_voxels with capacity >= count. */
int chad_voxel_count(chad_ctx* ctx, size_t* count);
int chad_export_voxels(chad_ctx* ctx, uint64_t* keys, uint32_t* sd_bits, uint32_t* weights, size_t capacity, size_t* count);
/* NodeLevel::_raw_data[0.._occupied_n) for level 0..19 (u32 words) and
 * LeafClusterLevel::_raw_data[0.._uniques_n] for level 20 (u64 words) (levels.hpp:90-93,141-143). */
int chad_level_words(chad_ctx* ctx, int level, size_t* words);
int chad_level_counters(chad_ctx* ctx, int level, uint32_t* uniques, uint32_t* dupes);
int chad_export_level(chad_ctx* ctx, int level, void* dst, size_t capacity_words);

/* DAG read path on the device (NodeLevels::get_child_addr / try_get_lc, levels.hpp:147-192, and the leaf byte of
 * cluster.hpp:34-52): for n Morton keys (host), the quantised TSDF byte of that voxel in finalised submap `submap`'s tree,
 * 0xFF where the voxel does not exist. Decode: (byte - 127) / 127 * sdf_trunc. */
int chad_query_voxels(chad_ctx* ctx, uint32_t submap, const uint64_t* keys, size_t n, uint8_t* bytes);

/* Leaf iterator over finalised submap `submap`'s TSDF tree on the device -- the reader the reference sketches but never finishes
 * (tsdf.hpp:120-155; tsdf.cpp:88-159 walks root -> first leaf cluster with get_child_addr / try_get_lc): every voxel of the submap in
 * ascending Morton order as (key, quantised byte). Two-call protocol: keys == NULL returns the count. */
int chad_iterate_leaves(chad_ctx* ctx, uint32_t submap, uint64_t* keys, uint8_t* bytes, size_t capacity, size_t* count);
/* Submap::positions (submap.hpp:110): the poses of the scans inserted into finalised submap `submap` (submap == chad_submap_count:
 * the active one), n x 3 floats. xyz == NULL returns the count. */
int chad_submap_positions(chad_ctx* ctx, uint32_t submap, float* xyz, size_t capacity, size_t* count);

/* Restoring a saved map (the reference has no persistence): the 21 level arrays exactly as chad_export_level
 * returned them, their counters (chad_level_counters), the submaps' roots and poses. The context must be empty; afterwards it behaves as
 * if it had built those submaps itself: the dedup sets are rebuilt from the arrays, so later submaps deduplicate against the restored
 * ones and receive the addresses an uninterrupted run would have given them. chad::TSDFMap::save / load wrap this with a file format. */
typedef struct chad_dag_image {
    const uint32_t* node_words[20];     /* NodeLevel::_raw_data[0.._occupied_n) per level */
    size_t node_word_count[20];
    const uint64_t* cluster_words;      /* LeafClusterLevel::_raw_data[0.._uniques_n] */
    size_t cluster_word_count;
    uint32_t uniques[CHAD_NUM_LEVELS], dupes[CHAD_NUM_LEVELS];
    const uint32_t* roots;              /* n_submaps x (root_addr_tsdf, root_addr_weight) */
    uint32_t n_submaps;
    const float* positions;             /* the submaps' poses, concatenated (x, y, z); may be NULL */
    const uint32_t* position_counts;    /* poses per submap; may be NULL */
} chad_dag_image;
int chad_import_dag(chad_ctx* ctx, const chad_dag_image* image);

/* How the band-voxel updates of a batch are grouped per voxel before the fold: 2 = tile runs + fused per-block sort and fold (default: runs.cu), 0 = block-binned (hashed
 * 8x8x8-voxel blocks + shared-memory sort), 1 = global onesweep radix sort. All give bit-identical results. */
int chad_set_pair_path(chad_ctx* ctx, int mode);

/* How the batches are pipelined on the device (context.cu header): *plan_slots = batches that can be between the start of their point
 * stage and the end of their fold at the same time (2 or 3), *walk_overlapped != 0 when the ray walk of a batch runs on its own stream
 * beside the next batch's point stage. Set at chad_create from CHAD_OVERLAP_WALK / CHAD_PLAN_SLOTS; never changes the results. */
int chad_pipeline_info(chad_ctx* ctx, int* plan_slots, int* walk_overlapped);

/* Forget everything (active submap, all DAG levels, submap roots, sticky errors) but keep the device
 * buffers: equivalent to destroying the map and constructing a new one with the same parameters. */
int chad_reset(chad_ctx* ctx);

int chad_get_stats(chad_ctx* ctx, chad_stats* out);
int chad_reset_stats(chad_ctx* ctx);

/* ---- stage entry points (kernel-by-kernel parity tests; host buffers) ------------------- */
/* calc_morton_vector + sort_morton_vector + estimate_normals (morton.hpp:59-102,
 * normals.hpp:81-148) of one scan: sorted points, their Morton keys, `order[i]` = input index of
 * sorted point i, normals. Any output pointer may be NULL. Does not touch the map. */
int chad_stage_points(chad_ctx* ctx, const float* xyz, size_t n, const float position[3], float* xyz_sorted,
                      uint64_t* keys, uint32_t* order, float* normals);
/* Band enumeration of octree.hpp:86-159 for already sorted points + normals: per-point voxel
 * counts and the (key, sd) stream in (point, ray step) order. `capacity` = room in keys/sd;
 * *total receives U. keys/sd may be NULL to query U. Does not touch the map. */
int chad_stage_pairs(chad_ctx* ctx, const float* xyz_sorted, const float* normals, size_t n, const float position[3],
                     uint32_t* counts, uint64_t* keys, float* sd, size_t capacity, size_t* total);
/* The onesweep radix sort on its own: stable ascending sort of (key, value) pairs on bits
 * [0, nbits) of the key (higher bits must be equal across keys or they are ignored). In place. */
int chad_stage_sort(chad_ctx* ctx, uint64_t* keys, uint32_t* values, size_t n, int nbits);
/* Morton encode (morton.hpp:21-28) on the device, n voxel coordinates (x,y,z int32 AoS). */
int chad_stage_morton(chad_ctx* ctx, const int32_t* voxels, size_t n, uint64_t* keys);

/* Input adaptors without copies. chad_insert DMAs page-locked memory straight from the caller's buffer (pageable
 * memory goes through a pinned staging ring): chad_host_alloc / chad_host_free hand out page-locked memory (chad::pinned_allocator wraps
 * them for std::vector<glm::vec3> / <Eigen::Vector3f> / <std::array<float,3>>), chad_host_register / chad_host_unregister page-lock
 * a buffer the caller already owns (e.g. a LiDAR driver's ring) for as long as it lives. */
int chad_host_alloc(size_t bytes, void** host_ptr);
int chad_host_free(void* host_ptr);
int chad_host_register(void* host_ptr, size_t bytes);
int chad_host_unregister(void* host_ptr);

/* Raw device pointer + CUDA stream used by the context, for callers that place inputs on the
 * device themselves (bench `value` leg). */
int chad_device_alloc(chad_ctx* ctx, size_t bytes, void** device_ptr);
int chad_device_free(chad_ctx* ctx, void* device_ptr);
int chad_upload(chad_ctx* ctx, void* device_dst, const void* host_src, size_t bytes);

/* Device timing helpers (CUDA events on the context's stream): begin/end a timed region and get
 * its duration; the region includes everything queued between the two calls. */
int chad_timer_begin(chad_ctx* ctx);
int chad_timer_end(chad_ctx* ctx, float* milliseconds);

/* ---- ONE map on several GPUs: Morton-range sharding (north_star) ---------------
 * The reference has a single map object (one octree + one NodeLevels, /root/reference/include/chad/tsdf.hpp:166-170); here the map
 * is cut into `world` contiguous Morton ranges of 8x8x8-voxel blocks and rank g -- one context, one GPU, one host thread or
 * process -- holds the voxels of range g. EVERY rank makes the SAME sequence of calls with the SAME scans (chad_insert* /
 * chad_flush / chad_finalize_active / chad_reset). A host scan crosses the host link once: each rank copies 1 / world of it and
 * one grouped all-gather per batch assembles the slices over NVLink (CHAD_SHARD_SLICE_H2D=0: every rank copies all). A rank sorts, estimates normals for and walks only the points of its own
 * range; the few band voxels a ray adds beyond the range travel to their owner once per batch (NCCL send / recv of fixed-size
 * boxes, counts inside -- no host round trip); at a submap's close the ranks' sorted leaf chunks are gathered on rank 0, which
 * runs Submap::finalize (submap.hpp:10-106) and holds the DAG (chad_level_* / chad_export_level / chad_query_voxels: rank 0
 * only; chad_submap_roots: every rank). chad_export_voxels returns the rank's own range (ascending; the concatenation in rank
 * order is the whole active submap). The union of the ranks' state is bit-identical to the single-GPU map.
 *   chad_shard_unique_id   rank 0: fill `id` (CHAD_SHARD_ID_BYTES) and hand it to every rank by any means (file, MPI, torch.distributed)
 *   chad_create_sharded    collective: every rank calls it with the same id; world == 1 is chad_create
 *   chad_shard_info        rank / world and what this rank has sent to the others so far (runs, 8-byte update records, exchanges) */
#define CHAD_SHARD_ID_BYTES 512
int chad_shard_unique_id(void* id);
int chad_create_sharded(float sdf_res, float sdf_trunc, int device, int max_batch_scans, int rank, int world, const void* id, chad_ctx** out);
int chad_shard_info(chad_ctx* ctx, int* rank, int* world, uint64_t* sent_runs, uint64_t* sent_records, uint64_t* exchanges);

/* ---- a submap integrated on another GPU (submap-parallel mode): chunk stream out, chunk stream in -----
 *   chad_shard_export_chunks  this context's leaf chunks, ascending: device pointers to n x u64 chunk keys and n x 64 B cells
 *   chad_shard_finalize_from  Submap::finalize from a device-resident, ascending chunk stream (a whole submap integrated by another
 *                           context), queued on the finalize stream; clear_local != 0 also clears the local table */
int chad_shard_export_chunks(chad_ctx* ctx, size_t* n_chunks, void** keys_device, void** cells_device);
int chad_shard_finalize_from(chad_ctx* ctx, const uint64_t* keys_device, const void* cells_device, size_t n_chunks, int clear_local);
/* Forget the active submap's voxels and first pose (octree.clear(), tsdf.cpp:57) without finalising anything here: the submap-parallel
 * mode calls it on a submap's owner after chad_shard_export_chunks (the finalize happens from the broadcast chunk stream on every rank). */
int chad_shard_clear(chad_ctx* ctx);

/* ---- host-only helpers (pure CPU bit arithmetic; usable without a GPU) ------------------------ */
/* MortonCode::encode / decode (morton.hpp:21-37): 21 bits per axis, bias 2^20, x -> bit 0. */
uint64_t chad_morton_encode(int32_t x, int32_t y, int32_t z);
void chad_morton_decode(uint64_t key, int32_t* x, int32_t* y, int32_t* z);
/* The order-preserving compact sort key used by the device radix sorts: for voxel coordinates in
 * [-2^k, 2^k) the 63-bit Morton key is reduced to its low 3k bits plus the top (sign) triple. */
uint64_t chad_key_compact(uint64_t key, unsigned k);
uint64_t chad_key_expand(uint64_t compact, unsigned k);

/* ---- instrumentation (bench.py roofline leg) ------------------------------------------------ */
/* When enabled, every kernel (group) launch on the insert path is bracketed by CUDA events on the
 * context's stream; the per-class totals are available after the next chad_flush. Enabling or
 * disabling flushes and zeroes the totals. Costs a few percent; leave it off in timed regions. */
int chad_profile_enable(chad_ctx* ctx, int on);
int chad_profile_classes(void);
int chad_profile_get(chad_ctx* ctx, int cls, const char** name, double* milliseconds, uint64_t* launches);
/* The instrumented launches of the last profiled flush in launch order: class, begin and end in milliseconds since the first one
 * (device time, all streams). Two-call protocol: NULL arrays return the count. */
int chad_profile_timeline(chad_ctx* ctx, int* classes, float* begin_ms, float* end_ms, size_t capacity, size_t* count);

#ifdef __cplusplus
}
#endif
#endif /* CHAD_B200_H */
