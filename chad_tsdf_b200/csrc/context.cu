// Host side of the C ABI (include/chad_b200.h): the context that owns all device state, the
// submap bookkeeping of TSDFMap::insert (/root/reference/src/chad/tsdf.cpp:39-75), batching of
// consecutive scans, and the level-by-level driver of Submap::finalize
// (/root/reference/include/chad/detail/submap.hpp:10-106).
//
// Scheduling. insert() copies the scan into the current batch (page-locked caller memory is DMA'd
// directly, pageable memory goes through a pinned staging ring) on the copy stream and returns. When a batch is full its
// "front" (plan, point sort, normals, ray walk) is queued on the main stream, the descriptor sort + block list on the
// group stream, and the batch joins a queue of at most two pending folds. A fold needs the batch's block count on the host
// (table sizing), so it is launched -- on the fold stream, beside the next batch's front -- by the first API call that
// finds the front's read-back there, at the latest when the batch's buffers are needed again two batches later: the
// host never waits inside the steady state. Submap::finalize runs on its own high-priority stream (finalize_begin).
// Optional (CHAD_OVERLAP_WALK=1, off by default -- measured no gain, profiles/ab_overlap_r01.md): the ray walk on its own
// stream beside the next batch's point stage, with three plan slots instead of two.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#include "chad_b200.h"
#include "kernels.cuh"
#include "nccl_dyn.h"
#include "radix_sort.cuh"
#include "scan.cuh"

using namespace chadgpu;

namespace {

thread_local std::string g_create_error;

constexpr int MAX_SLOTS = 3;  // plan slots = batches between the start of a point stage and the end of the batch's fold

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct Level {
    DevBuf raw;            // u32 words (node levels) or u64 words (cluster level)
    size_t raw_cap = 0;    // capacity in words
    DevBuf entries, first;
    DedupTable table{nullptr, nullptr, 0};
    u32 uniques = 0, dupes = 0, occupied = 0;
};

size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

struct chad_ctx {
    int device = 0, num_sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t xyz_free[2] = {nullptr, nullptr};      // the point stage has finished reading d_xyz[b]: the batch after next may fill it
    bool xyz_free_valid[2] = {false, false};
    cudaEvent_t scans_uploaded[2] = {nullptr, nullptr}; // h_scans_pinned[b] has been copied to the device: the host may rewrite it
    bool scans_uploaded_valid[2] = {false, false};
    MapParams mp{};
    int max_batch = 16;
    u32 first_batch = 4;  // scans of a batch that starts a burst (nothing in flight): see end_scan. CHAD_FIRST_BATCH, read at chad_create
    std::string error;
    int sticky_error = CHAD_OK;

    // submap state (tsdf.cpp:46-61, submap.hpp:108-110)
    bool has_pose = false;
    float first_pose[3] = {0, 0, 0};
    std::vector<std::array<u32, 2>> roots;
    std::vector<std::vector<std::array<float, 3>>> positions;  // Submap::positions (submap.hpp:110) of every closed submap, in closing order
    std::vector<std::array<float, 3>> active_positions;        // ... and of the active one

    // batch assembly
    size_t cap_points = 0, cap_pairs = 0;
    float* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_copied[2] = {nullptr, nullptr};  // last H2D out of h_stage[b] has completed
    bool stage_busy[2] = {false, false};
    DevBuf d_xyz[2];
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    int cur = 0;
    u32 batch_points = 0, batch_scans = 0;
    BatchScans h_scans{};
    BatchScans* h_scans_pinned[2] = {nullptr, nullptr};
    DevBuf d_scans, d_plan;       // d_plan = BatchPlan[MAX_SLOTS]: the batch being queued and the batches whose fold is pending / running
    BatchPlan* h_plan = nullptr;  // pinned BatchPlan[MAX_SLOTS]
    int plan_slot = 0;            // slot of the batch being assembled
    int n_slots = 2;              // slots in use: 2, or 3 with overlap_walk (a batch's walk, descriptor sort and fold then span two point stages)
    struct PendingFold { int slot; bool runs; u32 max_pairs; bool close; u64 close_index; };
    PendingFold pend[MAX_SLOTS];  // batches whose front is queued but whose fold is not (oldest first): the fold needs the batch's block
    int n_pend = 0;               // count on the host (table sizing), so it is launched as soon as the front's read-back has arrived --
                                  // by any later API call that finds it there, at the latest when the batch's plan slot is needed again
    u32* h_table_count = nullptr; // pinned: chunk count of `table` after its last fold
    u32* h_table_count2 = nullptr; // pinned: same for `table2`
    cudaEvent_t front_done[MAX_SLOTS] = {};  // per plan slot: the front's plan read-back has arrived

    // batch work buffers
    DevBuf bt_mem;                  // block table of the block-binned pair path
    BlockTable bt{};
    int pair_path = 2;              // 2 = tile runs + fused block sort/fold (default), 0 = block-binned, 1 = global radix sort
    DevBuf run_mem[MAX_SLOTS];      // run descriptors of the tile-run path, one set per plan slot: the fold of batch i runs on
    RunBuffers rb[MAX_SLOTS]{};     // fold_stream while the front of batch i + 1 runs on `stream` (records: keys_a / keys_b / keys_c by slot)
    cudaStream_t fold_stream = nullptr;
    cudaStream_t group_stream = nullptr;  // descriptor sort + block list of a batch: beside the next point stage AND the previous fold
    cudaStream_t walk_stream = nullptr;   // overlap_walk: the ray walk of batch i, beside the point stage of batch i + 1 on `stream`
    bool overlap_walk = false;            // (sorted points, normals and the scan table are then buffered per plan slot). CHAD_OVERLAP_WALK=1
                                          // turns it on; off by default: measured 9.24 vs 9.29 ms per bench step (profiles/ab_overlap_r01.md) --
                                          // the walk's 16 384 CTAs are dispatched before the next point stage's kernels get an SM either way
    cudaEvent_t points_done[MAX_SLOTS] = {};  // per plan slot: the point stage has written xyz_sorted / normals of the slot
    cudaEvent_t fold_done[MAX_SLOTS] = {};    // the fold that read slot b's records / descriptors / plan has finished
    bool fold_done_valid[MAX_SLOTS] = {};
    bool fold_in_flight = false;    // a fold has been queued on fold_stream since the last synchronisation
    cudaEvent_t submap_closed2 = nullptr, emit_done = nullptr;
    DevBuf radix_ws2;               // workspace of the descriptor sort (fold stream; the point sort's runs concurrently on the main stream)
    RadixWorkspace rws2{};
    u64 prev_fold_bound[MAX_SLOTS - 1] = {};  // chunk bounds of the last n_slots - 1 folds of the active submap (newest first): their exact
                                    // counts may not have arrived yet when the next fold's table is sized
    cudaStream_t prof_stream = nullptr;  // stream the instrumentation events are recorded on
    cudaStream_t last_fold_stream = nullptr;  // stream the most recent fold (and the copy of the table counter behind it) was queued on
    BatchPlan* h_plan_fold = nullptr;  // pinned BatchPlan[MAX_SLOTS]: the plan as the fused fold left it (distinct voxels, deferred errors)
    bool fold_stats_pending[MAX_SLOTS] = {};
    DevBuf pk_a, pk_b, pv_a, pv_b;  // point-sort ping-pong (N-sized): separate from the pair buffers so that the next batch's point
                                    // stage can be queued while the previous batch's pairs still wait for their fold
    DevBuf keys_a, keys_b, vals_a, vals_b, sorted_keys, sorted_order, xyz_sorted, normals, seg_info, counts, offsets, radix_ws, scan_ws;
    DevBuf xyz_sorted2[MAX_SLOTS - 1], normals2[MAX_SLOTS - 1], d_scans2[MAX_SLOTS - 1];  // plan slots 1 and 2 (slot 0 uses xyz_sorted / normals /
                                    // d_scans, like the stage and shard calls)
    DevBuf keys_c;                  // update records of plan slot 2 (slots 0 / 1: keys_a / keys_b)
    RadixWorkspace rws{};

    // resident chunk tables: `table` belongs to the active submap; the other one is being finalised / is spare
    DevBuf t_keys, t_cells, t_count, t_list;
    ChunkTable table{nullptr, nullptr, 0, nullptr, nullptr};
    DevBuf t2_keys, t2_cells, t2_count, t2_list;
    ChunkTable table2{nullptr, nullptr, 0, nullptr, nullptr};
    u64 table_count_known = 0;

    // asynchronous Submap::finalize (see finalize_begin): part 1 and part 2 run on fin_stream
    cudaStream_t fin_stream = nullptr;
    enum FinState { FIN_IDLE = 0, FIN_PART1 = 1 /* tables swapped, waiting for the closed submap's exact chunk count */,
                    FIN_PART2 = 2 /* everything queued on fin_stream */ };
    int fin_state = FIN_IDLE;
    u32 fin_max_chunks = 0;       // host upper bound of the chunk count of the submap being finalised
    u32 fin_chunks = 0;           // exact count (known after part 1)
    u32 fin_level_nodes[20] = {}; // exact node count per level (known after part 1)
    cudaEvent_t submap_closed = nullptr, fin_p1_done = nullptr, fin_done = nullptr;
    cudaEvent_t fin_t0 = nullptr, fin_t3 = nullptr;  // profiling: the whole finalize
    struct FinHost {              // pinned read-back area
        u32 scalars[16];
        u32 level_nodes[20];
        u32 root[2];
        LevelCounters counters[CHAD_NUM_LEVELS];
        u32 h2d[4];               // page-locked sources of small host-to-device copies (a pageable source would make the copy wait for the stream)
    }* h_fin = nullptr;
    DevBuf f_counters, f_partial;  // device LevelCounters[21]; partial sums of the persistent levels kernel
    bool fin_external = false;    // the finalize in flight consumes a caller-provided chunk stream (sharded mode): clear `table`, not `table2`

    // Morton-range sharding: this context is rank `rank` of `world` ranks that hold ONE map; world == 1 is the
    // plain single-GPU map. Every rank is fed the same scans; it sorts / walks the points of its own Morton range, the runs of blocks
    // beyond the range go to their owner (comm_x, group stream, once per batch) and at a submap's close the ranks' sorted leaf chunks
    // are gathered on rank 0 (comm_f, finalize stream), which holds the DAG.
    struct Shard {
        int rank = 0, world = 1;
        const NcclApi* nccl = nullptr;
        ncclComm_t comm_x = nullptr, comm_f = nullptr, comm_c = nullptr;
        bool slice_h2d = true;          // a scan comes over the host link once, not once per rank: every rank copies 1 / world of it and the
                                        // slices are all-gathered over NVLink (comm_c, copy stream). CHAD_SHARD_SLICE_H2D=0: every rank copies all
        cudaEvent_t h2d_done = nullptr;
        struct Slice { float* dst; size_t count; };
        std::vector<Slice> slices;      // scans of the batch being assembled that arrived as slices: {place in the batch buffer, floats per rank}
        DevBuf splitters;               // u64[2][SHARD_WORLD_MAX + 1]: range starts (block ids), two sets used alternately by the submaps
        int split_set = 0;              // set of the active submap
        int slot_split[MAX_SLOTS] = {}; // set a plan slot's batch was filtered with (its pack kernel runs later, on the group stream)
        bool need_splitters = true;     // the next batch is the first of a submap
        u32 first_share_256 = 256;      // rank 0's share of the rays relative to 256 for every other rank (CHAD_SHARD_RANK0_SHARE)
        DevBuf filter_mem;
        ShardFilter filter{};
        DevBuf batch_scans[MAX_SLOTS];  // scan table of the whole batch (slot_scans() holds the table of this rank's own points)
        DevBuf box_out, box_in;         // exchange boxes [world][box_words] u64
        u32 box_words = 0;
        DevBuf scalars;                 // u32: [0..1] sample sort (n, nbits) | [8] own chunk count | [16 .. 16 + world) all chunk counts
        u32* h_counts = nullptr;        // pinned [SHARD_WORLD_MAX]: chunk counts of the submap being closed
        cudaEvent_t counts_done = nullptr;
        // The gather of a closed submap's chunks is issued at a point of the BATCH sequence that is the same on every rank (in front of
        // the n_slots-th batch after the closing one, or at a flush -- whichever comes first), never from a poll: every rank has then
        // issued the same batch exchanges, so the two kinds of transfer meet in the same order everywhere. Reason: an
        // NCCL kernel that waits on the device for a peer whose HOST has not got there yet blocks this rank's other NCCL traffic, and the
        // peer's host may be waiting for exactly that traffic.
        std::deque<u64> gather_at;      // per closed, not yet gathered submap: the batch (sequence number) in front of which its gather is issued
        u64 front_seq = 0;              // batches queued so far
        u64 closes_marked = 0, closes_gathered = 0;
        // A close = table swap. It needs the spare table back, i.e. the gather of the submap closed before must have been issued. A
        // sharded rank never makes a poll wait for that: the closing batch's fold is launched at once, the close itself is deferred until
        // it can go through -- at the latest in front of the next fold.
        bool closed_waiting = false;    // the spare table holds a closed submap whose gather has not been issued yet. (On a sharded rank this is
                                        // NOT ctx->fin_state: the DAG stage of the submap closed before may still be running on rank 0 while
                                        // the next submap closes -- only the gather of submap j + 1 waits for the finalize of submap j)
        cudaEvent_t table_free = nullptr;  // the closed submap's chunks have left the spare table and it has been cleared (finalize stream)
        bool table_free_valid = false;
        bool close_deferred = false;
        u64 deferred_index = 0;
        cudaStream_t deferred_count_stream = nullptr;
        bool xfer_pending = false;      // counts_done marks the end of this rank's gather transfers: the next exchange waits for it
        u32* h_roots = nullptr;         // pinned staging of the root broadcast
        size_t roots_synced = 0;        // submaps whose roots this rank knows (rank 0: has broadcast)
        u64 sent_runs = 0, sent_records = 0, exchanges = 0;
    } sh;
    // growth of the never-freed DAG arenas and the chunk tables (VirtualArray, virtual_array.hpp:12-104, becomes explicit device buffers):
    // a buffer that has been outgrown is parked here instead of freed -- cudaFree synchronises the whole device, i.e. stalls every
    // stream of the pipeline -- and released at the next drain
    std::vector<void*> graveyard;
    u64 grow_events = 0, grow_bytes = 0;
    double grow_host_ms = 0.0;
    u32 burst_batches = 0;              // batches queued since the last drain (sharded: the batch boundaries must not depend on timing)

    // finalize work buffers
    size_t cap_chunks = 0;
    DevBuf f_sorted;  // full chunk keys in ascending order (gather output)
    DevBuf f_ids[2], f_slots[2], f_cells, f_tsdf, f_addr[2], f_head_rank, f_cand, f_slot_of, f_is_new, f_rank, f_radix_ws, f_scan_ws, f_scalars;
    RadixWorkspace f_rws{};

    Level levels[CHAD_NUM_LEVELS];
    chad_stats stats{};
    cudaEvent_t t0 = nullptr, t1 = nullptr;

    // optional per-kernel CUDA-event instrumentation (chad_profile_*)
    bool profiling = false;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    struct TimelineEntry { int cls; float t0, t1; };
    std::vector<TimelineEntry> timeline;  // instrumented launches of the last profiled flush, ms since its first launch
    double prof_ms[PC_COUNT] = {};
    u64 prof_launches[PC_COUNT] = {};
    LaunchHook hook{nullptr, nullptr, nullptr};
};

namespace {

// CHAD_TRACE=1: host-side milestones (rank, milliseconds since the first one) on stderr -- for diagnosing a stalled multi-rank run
void trace(const chad_ctx* ctx, const char* fmt, ...) {
    static const bool on = [] { const char* e = std::getenv("CHAD_TRACE"); return e && std::atoi(e) != 0; }();
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    std::fprintf(stderr, "[chad r%d %9.3f] %s\n", ctx ? ctx->sh.rank : -1, ms, buf);
}

int fail(chad_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->error = msg; else g_create_error = msg;
    return code;
}
#define CUDA_TRY(ctx, expr)                                                                                   \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return fail(ctx, CHAD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define TRY(expr)                   \
    do {                            \
        int _r = (expr);            \
        if (_r != CHAD_OK) return _r; \
    } while (0)
#define NCCL_TRY(ctx, expr)                                                                                      \
    do {                                                                                                         \
        ncclResult_t _n = (expr);                                                                                \
        if (_n != ncclSuccess) return fail(ctx, CHAD_ERR_CUDA, std::string(#expr) + ": " + (ctx)->sh.nccl->GetErrorString(_n)); \
    } while (0)

void prof_begin(void* user, int cls) {
    chad_ctx* ctx = static_cast<chad_ctx*>(user);
    if (!ctx->profiling) return;
    chad_ctx::Span sp{cls, nullptr, nullptr};
    for (cudaEvent_t* e : {&sp.a, &sp.b}) {
        if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(sp.a, ctx->prof_stream ? ctx->prof_stream : ctx->stream);
    ctx->spans.push_back(sp);
}
void prof_end(void* user) {
    chad_ctx* ctx = static_cast<chad_ctx*>(user);
    if (!ctx->profiling || ctx->spans.empty()) return;
    cudaEventRecord(ctx->spans.back().b, ctx->prof_stream ? ctx->prof_stream : ctx->stream);
}
// after a stream synchronisation: fold the recorded spans into the per-class totals
void prof_resolve(chad_ctx* ctx) {
    if (!ctx->spans.empty()) ctx->timeline.clear();
    for (auto& sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            ctx->prof_ms[sp.cls] += ms;
            ctx->prof_launches[sp.cls]++;
            float t0 = 0.f;
            if (cudaEventElapsedTime(&t0, ctx->spans.front().a, sp.a) == cudaSuccess) ctx->timeline.push_back({sp.cls, t0, t0 + ms});
            else cudaGetLastError();
        }
        else cudaGetLastError();
        ctx->event_pool.push_back(sp.a);
        ctx->event_pool.push_back(sp.b);
    }
    ctx->spans.clear();
}
#define PROF(ctx, cls, expr)         \
    do {                             \
        prof_begin(ctx, cls);        \
        launches += (expr);          \
        prof_end(ctx);               \
    } while (0)

int dev_ensure(chad_ctx* ctx, DevBuf& b, size_t bytes, bool preserve = false) {
    if (b.bytes >= bytes && b.p) return CHAD_OK;
    void* np = nullptr;
    size_t nb = bytes < 256 ? 256 : bytes;
    CUDA_TRY(ctx, cudaMalloc(&np, nb));
    if (b.p) {
        if (preserve) CUDA_TRY(ctx, cudaMemcpyAsync(np, b.p, b.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        // parked, not freed: cudaFree synchronises the whole DEVICE. Beside stalling every stream, on a sharded map that is a deadlock:
        // an NCCL kernel of this rank may be waiting for a peer whose host waits for an exchange this (blocked) host has yet to issue
        ctx->graveyard.push_back(b.p);
    }
    b.p = np;
    b.bytes = nb;
    return CHAD_OK;
}
void dev_free(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}
BatchPlan* plan_ptr(chad_ctx* ctx, int slot) { return ctx->d_plan.as<BatchPlan>() + slot; }
template <typename T> T* plan_field(chad_ctx* ctx, int slot, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(plan_ptr(ctx, slot)) + off); }
// what the point stage of a batch hands to its ray walk, per plan slot
float* slot_xyz_sorted(chad_ctx* ctx, int slot) { return (slot ? ctx->xyz_sorted2[slot - 1] : ctx->xyz_sorted).as<float>(); }
float* slot_normals(chad_ctx* ctx, int slot) { return (slot ? ctx->normals2[slot - 1] : ctx->normals).as<float>(); }
BatchScans* slot_scans(chad_ctx* ctx, int slot) { return (slot ? ctx->d_scans2[slot - 1] : ctx->d_scans).as<BatchScans>(); }
u64* slot_records(chad_ctx* ctx, int slot) { return (slot == 0 ? ctx->keys_a : slot == 1 ? ctx->keys_b : ctx->keys_c).as<u64>(); }
void fold_bounds_reset(chad_ctx* ctx) { for (u64& b : ctx->prev_fold_bound) b = 0; }

int error_from_flags(chad_ctx* ctx, u32 flags) {
    if (!flags) return CHAD_OK;
    int code = CHAD_ERR_RANGE;
    std::string msg = "device error:";
    if (flags & ERRF_NUMERIC) { msg += " NaN/Inf point coordinate;"; code = CHAD_ERR_NUMERIC; }
    if (flags & ERRF_RANGE) { msg += " voxel coordinate outside the 21-bit Morton range;"; code = CHAD_ERR_RANGE; }
    if (flags & ERRF_KEY_BUDGET) { msg += " batch key budget exceeded (reduce max_batch_scans);"; code = CHAD_ERR_RANGE; }
    if (flags & ERRF_PAIR_CAPACITY) { msg += " band voxel buffer overflow;"; code = CHAD_ERR_CAPACITY; }
    if (flags & ERRF_TABLE_FULL) { msg += " resident chunk table full;"; code = CHAD_ERR_CAPACITY; }
    if (flags & ERRF_DEDUP_FULL) { msg += " DAG dedup table full;"; code = CHAD_ERR_CAPACITY; }
    if (flags & ERRF_EXCHANGE) { msg += " sharded map: the runs for another rank did not fit the exchange box (raise CHAD_SHARD_BOX_MB);"; code = CHAD_ERR_CAPACITY; }
    if (flags & ERRF_BLOCKS_FULL) { msg += " block table of the block-binned pair path full (chad_set_pair_path(ctx, 1) selects the global sort);"; code = CHAD_ERR_CAPACITY; }
    ctx->sticky_error = code;
    return fail(ctx, code, msg);
}

// ---- resident chunk table ---------------------------------------------------------------
int table_alloc(chad_ctx* ctx, ChunkTable& t, DevBuf& keys, DevBuf& cells, DevBuf& count, DevBuf& list, u64 capacity) {
    TRY(dev_ensure(ctx, keys, capacity * 8));
    TRY(dev_ensure(ctx, cells, capacity * 64));
    TRY(dev_ensure(ctx, count, 256));
    TRY(dev_ensure(ctx, list, capacity * 4 + 256));
    t.list = list.as<u64>();
    t.keys = keys.as<u64>();
    t.cells = cells.as<uint2>();
    t.capacity = capacity;
    t.count = count.as<u32>();
    launch_table_clear(ctx->stream, t);
    return CHAD_OK;
}
int table_reserve(chad_ctx* ctx, u64 need_chunks) {
    if (need_chunks * 2 <= ctx->table.capacity) return CHAD_OK;
    const auto t_grow = std::chrono::steady_clock::now();
    trace(ctx, "chunk table grows: %llu chunks needed, capacity %llu", (unsigned long long)need_chunks, (unsigned long long)ctx->table.capacity);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->group_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fold_stream));  // a fold may still be writing the table
    ctx->fold_in_flight = false;
    const u64 new_cap = next_pow2(need_chunks * 4);
    DevBuf nk, nc, ncount, nlist;
    ChunkTable nt{nullptr, nullptr, 0, nullptr, nullptr};
    TRY(table_alloc(ctx, nt, nk, nc, ncount, nlist, new_cap));
    ctx->stats.kernel_launches += launch_table_rehash(ctx->stream, ctx->table, nt, ctx->num_sms);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (DevBuf* b : {&ctx->t_keys, &ctx->t_cells, &ctx->t_count, &ctx->t_list}) ctx->graveyard.push_back(b->p);  // (see dev_ensure)
    ctx->t_keys = nk; ctx->t_cells = nc; ctx->t_count = ncount; ctx->t_list = nlist;
    ctx->table = nt;
    ctx->grow_events++;
    ctx->grow_bytes += new_cap * 76;
    trace(ctx, "chunk table grown");
    ctx->grow_host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_grow).count();
    return CHAD_OK;
}

// ---- batch buffers ------------------------------------------------------------------------
int ensure_batch_capacity(chad_ctx* ctx, size_t points) {
    if (points <= ctx->cap_points) return CHAD_OK;
    // only called while no batch is in flight
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->walk_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->group_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fold_stream));
    const size_t np = points + points / 8 + 1024;
    const size_t pairs = np * ctx->mp.max_ray_voxels;
    if (pairs >= (1ull << 30)) return fail(ctx, CHAD_ERR_CAPACITY, "batch too large: more than 2^30 band voxels; lower max_batch_scans");
    for (int b = 0; b < 2; b++) {
        if (ctx->h_stage[b]) CUDA_TRY(ctx, cudaFreeHost(ctx->h_stage[b]));
        CUDA_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_stage[b]), np * 12));
        ctx->stage_busy[b] = false;
        TRY(dev_ensure(ctx, ctx->d_xyz[b], np * 12 + 1024, true));  // (+ slack: the sharded all-gather of a scan rounds the slices up)
    }
    TRY(dev_ensure(ctx, ctx->pk_a, np * 8));
    TRY(dev_ensure(ctx, ctx->pk_b, np * 8));
    TRY(dev_ensure(ctx, ctx->pv_a, np * 4));
    TRY(dev_ensure(ctx, ctx->pv_b, np * 4));
    TRY(dev_ensure(ctx, ctx->keys_a, pairs * 8));
    TRY(dev_ensure(ctx, ctx->keys_b, pairs * 8));
    TRY(dev_ensure(ctx, ctx->vals_a, pairs * 4));
    TRY(dev_ensure(ctx, ctx->vals_b, pairs * 4));
    TRY(dev_ensure(ctx, ctx->sorted_keys, np * 8));
    TRY(dev_ensure(ctx, ctx->sorted_order, np * 4));
    TRY(dev_ensure(ctx, ctx->xyz_sorted, np * 12));
    TRY(dev_ensure(ctx, ctx->normals, np * 12));
    for (int q = 0; q + 1 < ctx->n_slots; q++) {
        TRY(dev_ensure(ctx, ctx->xyz_sorted2[q], np * 12));
        TRY(dev_ensure(ctx, ctx->normals2[q], np * 12));
    }
    if (ctx->n_slots > 2) TRY(dev_ensure(ctx, ctx->keys_c, pairs * 8));
    TRY(dev_ensure(ctx, ctx->seg_info, np * 4));
    TRY(dev_ensure(ctx, ctx->counts, np * 4));
    TRY(dev_ensure(ctx, ctx->offsets, np * 4));
    TRY(dev_ensure(ctx, ctx->radix_ws, radix_workspace_bytes(pairs)));
    size_t bcap = next_pow2(np / 4);
    if (bcap < (1u << 16)) bcap = 1u << 16;
    if (bcap > (1u << 22)) bcap = 1u << 22;
    TRY(dev_ensure(ctx, ctx->bt_mem, blocks_table_bytes((u32)bcap)));
    ctx->bt = blocks_table_carve(ctx->bt_mem.p, (u32)bcap);
    TRY(dev_ensure(ctx, ctx->scan_ws, scan_workspace_bytes(np > bcap ? np : bcap)));
    if (ctx->mp.max_ray_runs <= runs_max_ray_runs() && ctx->mp.max_ray_voxels <= runs_max_ray_voxels()) {
        const size_t dcap = np * ctx->mp.max_ray_runs;
        for (int b = 0; b < ctx->n_slots; b++) {
            TRY(dev_ensure(ctx, ctx->run_mem[b], runs_desc_bytes(dcap)));
            ctx->rb[b] = runs_carve(ctx->run_mem[b].p, dcap);
        }
        TRY(dev_ensure(ctx, ctx->radix_ws2, radix_workspace_bytes(dcap)));
        ctx->rws2 = radix_workspace_carve(ctx->radix_ws2.p, dcap);
    }
    ctx->rws = radix_workspace_carve(ctx->radix_ws.p, pairs);
    if (ctx->sh.world > 1) {
        TRY(dev_ensure(ctx, ctx->sh.filter_mem, shard_filter_bytes(np)));
        ctx->sh.filter = shard_filter_carve(ctx->sh.filter_mem.p, np);
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->sh.filter_mem.p, 0, ctx->sh.filter_mem.bytes, ctx->stream));  // the per-scan counters start (and are left) at zero
    }
    ctx->cap_points = np;
    ctx->cap_pairs = pairs;
    return CHAD_OK;
}

int finalize_begin(chad_ctx* ctx, u32 max_chunks, bool external, cudaStream_t count_stream);
int shard_gather_now(chad_ctx* ctx);
int try_deferred_close(chad_ctx* ctx, bool block);
int finalize_poll(chad_ctx* ctx);
int finalize_wait(chad_ctx* ctx);

// device mirror of the node levels' counters: NodeLevel's constructor reserves index 0 (levels.hpp:52-54)
int level_counters_reset(chad_ctx* ctx) {
    LevelCounters init[CHAD_NUM_LEVELS];
    for (int d = 0; d < CHAD_NUM_LEVELS; d++) init[d] = LevelCounters{d == CHAD_LEVEL_CLUSTERS ? 0u : 1u, 0u, 0u, 0u};
    CUDA_TRY(ctx, cudaMemcpy(ctx->f_counters.p, init, sizeof(init), cudaMemcpyHostToDevice));
    return CHAD_OK;
}

// the fused fold of the tile-run path reports the batch's distinct voxels after the fact: collect what has arrived.
// Only called when the stream has passed the copies (after an event / stream synchronisation that follows them).
void account_fold_stats(chad_ctx* ctx, int only_slot = -1) {
    for (int slot = 0; slot < MAX_SLOTS; slot++) {
        if (!ctx->fold_stats_pending[slot] || (only_slot >= 0 && slot != only_slot)) continue;
        ctx->fold_stats_pending[slot] = false;
        ctx->stats.scan_voxels += ctx->h_plan_fold[slot].n_segments;
    }
}

// Launch the fold of the oldest batch whose front has been queued. block = false: only if its read-back has arrived already
// (*launched tells). The folds are launched in batch order.
int complete_one_fold(chad_ctx* ctx, bool block, bool* launched) {
    *launched = false;
    if (ctx->n_pend == 0) return CHAD_OK;
    const chad_ctx::PendingFold pf = ctx->pend[0];
    const int slot = pf.slot;
    if (!block) {
        const cudaError_t q = cudaEventQuery(ctx->front_done[slot]);
        if (q == cudaErrorNotReady) { cudaGetLastError(); return CHAD_OK; }
        CUDA_TRY(ctx, q);
    } else {
        trace(ctx, "wait: front of slot %d", slot);
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->front_done[slot]));
    }
    if (ctx->sh.close_deferred) {  // this fold goes into the NEXT submap's table: the swap must have happened
        TRY(try_deferred_close(ctx, block));
        if (ctx->sh.close_deferred) return CHAD_OK;  // (from a poll: not yet possible; the fold waits)
    }
    for (int q = 0; q + 1 < ctx->n_pend; q++) ctx->pend[q] = ctx->pend[q + 1];
    ctx->n_pend--;
    *launched = true;
    trace(ctx, "fold of slot %d (close %d)", slot, (int)pf.close);
    account_fold_stats(ctx, slot);  // the pair stage of this batch waited for the fold that used this slot before
    BatchPlan plan = ctx->h_plan[slot];
    const bool runs = pf.runs;
    ctx->table_count_known = *ctx->h_table_count;
    ctx->stats.updates += plan.n_pairs;
    ctx->sh.sent_runs += plan.xfer_runs;
    ctx->sh.sent_records += plan.xfer_records;
    ctx->stats.key_bits_points = plan.nbits_points;
    ctx->stats.key_bits_pairs = plan.nbits_pairs;
    if (runs) {
        // the fused fold counts the distinct voxels / chunks itself; before it runs only a bound is known:
        // a block holds 64 leaf chunks and every new chunk needs at least one update
        const u64 bound = std::min<u64>(u64(plan.n_blocks) * 64, plan.n_pairs);
        if (bound >= (1ull << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "batch touches more than 2^31 leaf chunks");
        plan.n_chunk_heads = (u32)bound;
    } else {
        ctx->stats.scan_voxels += plan.n_segments;
    }
    if (plan.error) {
        CUDA_TRY(ctx, cudaMemsetAsync(plan_field<u32>(ctx, slot, offsetof(BatchPlan, error)), 0, 4, ctx->stream));
        return error_from_flags(ctx, plan.error);
    }
    // *h_table_count is exact as of the last fold whose copy has arrived. This batch's point stage waited for the fold of n_slots batches
    // ago, so at most the n_slots - 1 folds before this one may still be queued or running on fold_stream: their bounds are added --
    // an upper bound of the chunk count after this fold either way
    u64 count_bound = ctx->table_count_known + plan.n_chunk_heads;
    for (int q = 0; q + 1 < ctx->n_slots; q++) count_bound += ctx->prev_fold_bound[q];
    TRY(table_reserve(ctx, count_bound));
    u64 launches = 0;
    cudaStream_t fold_on = ctx->stream;
    if (runs) {
        // the fused fold runs on its own stream, concurrently with the next batch's front (point stage, ray walk, descriptor sort)
        cudaStream_t fs = ctx->fold_stream;
        fold_on = fs;
        CUDA_TRY(ctx, cudaStreamWaitEvent(fs, ctx->front_done[slot], 0));
        ctx->prof_stream = fs;
        if (plan.n_pairs) PROF(ctx, PC_RUNS_FOLD, launch_runs_fold(fs, slot_records(ctx, slot), ctx->rb[slot], plan_ptr(ctx, slot), ctx->table, ctx->num_sms));
        ctx->prof_stream = nullptr;
        CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->h_plan_fold[slot], plan_ptr(ctx, slot), sizeof(BatchPlan), cudaMemcpyDeviceToHost, fs));
        ctx->stats.d2h_bytes += sizeof(BatchPlan);
        ctx->fold_stats_pending[slot] = true;
        for (int q = MAX_SLOTS - 2; q > 0; q--) ctx->prev_fold_bound[q] = ctx->prev_fold_bound[q - 1];
        ctx->prev_fold_bound[0] = plan.n_chunk_heads;
        ctx->fold_in_flight = true;
    } else {
        fold_bounds_reset(ctx);
        PROF(ctx, PC_FOLD, launch_fold(ctx->stream, ctx->keys_a.as<u64>(), ctx->keys_b.as<u64>(), ctx->vals_a.as<u32>(), ctx->vals_b.as<u32>(),
                                       pf.max_pairs, plan_ptr(ctx, slot), ctx->table, ctx->num_sms));
    }
    ctx->stats.kernel_launches += launches;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_table_count, ctx->table.count, 4, cudaMemcpyDeviceToHost, fold_on));
    ctx->stats.d2h_bytes += 4;
    ctx->last_fold_stream = fold_on;
    if (runs) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->fold_done[slot], ctx->fold_stream));
        ctx->fold_done_valid[slot] = true;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if (pf.close) {  // that was the submap's last batch: swap tables and start its asynchronous finalize
        if (count_bound >= (1ull << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "submap exceeds 2^31 leaf chunks");
        if (ctx->sh.world > 1) {
            ctx->sh.close_deferred = true;
            ctx->sh.deferred_index = pf.close_index;
            ctx->sh.deferred_count_stream = fold_on;
            TRY(try_deferred_close(ctx, false));
        } else {
            TRY(finalize_begin(ctx, 0, false, fold_on));
        }
        ctx->stats.resident_clusters = 0;
    }
    return CHAD_OK;
}

// sharded: perform the deferred close if it can go through (block = false), or make it go through (block = true: only from the fixed
// points of the batch sequence, where a gather may be issued)
int try_deferred_close(chad_ctx* ctx, bool block) {
    chad_ctx::Shard& sh = ctx->sh;
    if (!sh.close_deferred) return CHAD_OK;
    while (sh.closes_gathered + 1 < sh.deferred_index) {  // the submap closed before still sits in the spare table
        if (!block) return CHAD_OK;
        TRY(shard_gather_now(ctx));
    }
    sh.close_deferred = false;
    return finalize_begin(ctx, 0, false, sh.deferred_count_stream);
}

// every pending fold, waiting for the fronts
int complete_pending_fold(chad_ctx* ctx) {
    bool launched;
    while (ctx->n_pend) TRY(complete_one_fold(ctx, true, &launched));
    return CHAD_OK;
}
// the pending folds whose front has already reported
int poll_folds(chad_ctx* ctx) {
    bool launched = true;
    while (ctx->n_pend && launched) TRY(complete_one_fold(ctx, false, &launched));
    return CHAD_OK;
}

// bits of a 256-ray tile index inside the batch's largest scan (order key of a run descriptor: points.cuh). Follows from the scan
// table of the WHOLE batch, so every rank of a sharded map derives the same value.
u32 batch_tsb(const BatchScans& h, u32 ns) {
    u32 longest = 0;
    for (u32 i = 0; i < ns; i++) longest = std::max(longest, h.offset[i + 1] - h.offset[i]);
    const u32 tiles = (longest + RAY_TILE - 1) / RAY_TILE;
    u32 bits = 0;
    while (bits < 32 && (1ull << bits) < tiles) bits++;
    return bits;
}
u32 shard_gbits(const chad_ctx* ctx) { u32 b = 0; while ((1 << b) < ctx->sh.world) b++; return b; }
const u64* shard_splitters(const chad_ctx* ctx, int set) { return ctx->sh.splitters.as<u64>() + size_t(set) * (SHARD_WORLD_MAX + 1); }

// plan -> point sort keys -> sort -> gather -> normals of the batch in d_xyz[b] (n points, ns scans; the scan table `h` uploaded).
// Sharded: the plan + ownership filter replace the plan + key kernels and everything behind them works on this rank's points only.
void queue_point_stage(chad_ctx* ctx, int slot, int b, u32 n, u32 ns, const BatchScans& h) {
    cudaStream_t s = ctx->stream;
    BatchPlan* plan = plan_ptr(ctx, slot);
    const BatchScans* scans = slot_scans(ctx, slot);
    const float* xyz = ctx->d_xyz[b].as<float>();
    u64 launches = 0;
    const LaunchHook* hook = ctx->profiling ? &ctx->hook : nullptr;
    const u32 tsb = batch_tsb(h, ns);
    if (ctx->sh.world > 1) {
        if (ctx->sh.need_splitters) {
            // every rank samples and sorts the same points (the submap's first scan), so every rank derives the same ranges: no communication
            ctx->sh.split_set ^= 1;
            ctx->sh.need_splitters = false;
            PROF(ctx, PC_PLAN, launch_shard_splitters(s, xyz, h.offset[1], ctx->mp, (u32)ctx->sh.world, ctx->sh.first_share_256,
                                                      const_cast<u64*>(shard_splitters(ctx, ctx->sh.split_set))));
        }
        ctx->sh.slot_split[slot] = ctx->sh.split_set;
        PROF(ctx, PC_POINT_KEYS, launch_shard_filter(s, xyz, n, ns, ctx->sh.batch_scans[slot].as<BatchScans>(), ctx->mp, plan, tsb, shard_gbits(ctx),
                                                     shard_splitters(ctx, ctx->sh.split_set), (u32)ctx->sh.rank, (u32)ctx->sh.world, ctx->sh.filter,
                                                     slot_scans(ctx, slot), ctx->pk_a.as<u64>(), ctx->pv_a.as<u32>()));
    } else {
        PROF(ctx, PC_PLAN, launch_plan(s, xyz, n, ns, ctx->mp, plan, tsb));
        PROF(ctx, PC_POINT_KEYS, launch_point_keys(s, xyz, n, scans, ctx->mp, plan, ctx->pk_a.as<u64>(), ctx->pv_a.as<u32>()));
    }
    launches += radix_sort_pairs(s, ctx->pk_a.as<u64>(), ctx->pv_a.as<u32>(), ctx->pk_b.as<u64>(), ctx->pv_b.as<u32>(),
                                 plan_field<u32>(ctx, slot, offsetof(BatchPlan, n_points)), plan_field<u32>(ctx, slot, offsetof(BatchPlan, nbits_points)), n,
                                 RS_MAX_PASSES, ctx->rws, ctx->num_sms, hook, PC_POINT_SORT_HIST, plan_field<u32>(ctx, slot, offsetof(BatchPlan, point_shift)));
    PROF(ctx, PC_POINT_GATHER, launch_point_gather(s, xyz, n, plan, ctx->pk_a.as<u64>(), ctx->pk_b.as<u64>(), ctx->pv_a.as<u32>(),
                                                   ctx->pv_b.as<u32>(), ctx->sorted_keys.as<u64>(), ctx->sorted_order.as<u32>(),
                                                   slot_xyz_sorted(ctx, slot)));
    PROF(ctx, PC_NORMALS, launch_normals(s, slot_xyz_sorted(ctx, slot), ctx->sorted_keys.as<u64>(), n, scans, plan, ctx->seg_info.as<u32>(),
                                         slot_normals(ctx, slot)));
    ctx->stats.kernel_launches += launches;
}

// sharded: the runs of blocks outside this rank's range go to their owner, the runs received are appended (group stream, between the
// walk and the descriptor sort). One grouped send / receive of fixed-size boxes per batch: the counts travel inside the boxes, so the
// host never learns (or waits for) them.
int queue_shard_exchange(chad_ctx* ctx, int slot) {
    trace(ctx, "exchange %llu (slot %d)", (unsigned long long)ctx->sh.exchanges, slot);
    cudaStream_t gs = ctx->group_stream;
    chad_ctx::Shard& sh = ctx->sh;
    const ShardBoxes boxes{sh.box_out.as<u64>(), sh.box_in.as<u64>(), sh.box_words};
    if (sh.xfer_pending) {  // (see finalize_gather: no exchange overtakes the chunk gather)
        CUDA_TRY(ctx, cudaStreamWaitEvent(gs, sh.counts_done, 0));
        sh.xfer_pending = false;
    }
    u64 launches = 0;
    launches += launch_runs_pack(gs, ctx->rb[slot].capacity, plan_ptr(ctx, slot), ctx->rb[slot], slot_records(ctx, slot), shard_splitters(ctx, sh.slot_split[slot]),
                                 (u32)sh.rank, (u32)sh.world, boxes, ctx->num_sms);
    NCCL_TRY(ctx, sh.nccl->GroupStart());
    for (int g = 0; g < sh.world; g++) {
        if (g == sh.rank) continue;
        NCCL_TRY(ctx, sh.nccl->Send(boxes.out + size_t(g) * boxes.words, boxes.words, ncclUint64, g, sh.comm_x, gs));
        NCCL_TRY(ctx, sh.nccl->Recv(boxes.in + size_t(g) * boxes.words, boxes.words, ncclUint64, g, sh.comm_x, gs));
    }
    NCCL_TRY(ctx, sh.nccl->GroupEnd());
    launches += 1;
    launches += launch_runs_ingest(gs, plan_ptr(ctx, slot), ctx->rb[slot], slot_records(ctx, slot), (u32)ctx->cap_pairs, (u32)sh.rank, (u32)sh.world, boxes);
    ctx->stats.kernel_launches += launches;
    sh.exchanges++;
    return CHAD_OK;
}

// Queue everything of the assembled batch up to (not including) the fold. The point stage (which does not touch the
// pair buffers) is queued FIRST, before the host waits for the previous batch's counts and queues its fold: the
// device always has that much work in hand while the host synchronises and launches (the pipeline was launch-bound
// otherwise: ~50 launches per batch against ~2 ms of kernels).
int process_front(chad_ctx* ctx) {
    if (ctx->batch_scans == 0) return CHAD_OK;
    trace(ctx, "front: %u scans, %u points, slot %d", ctx->batch_scans, ctx->batch_points, ctx->plan_slot);
    const int b = ctx->cur;
    const int slot = ctx->plan_slot;
    const u32 n = ctx->batch_points, ns = ctx->batch_scans;
    cudaStream_t s = ctx->stream;
    ctx->h_scans.offset[ns] = n;
    if (!ctx->sh.slices.empty()) {  // scans that came as 1 / world slices over the host link (insert_host): assemble them, one launch
        chad_ctx::Shard& sh = ctx->sh;
        NCCL_TRY(ctx, sh.nccl->GroupStart());
        for (const auto& sl : sh.slices) NCCL_TRY(ctx, sh.nccl->AllGather(sl.dst + sl.count * (size_t)sh.rank, sl.dst, sl.count, ncclFloat, sh.comm_c, ctx->copy_stream));
        NCCL_TRY(ctx, sh.nccl->GroupEnd());
        sh.slices.clear();
        ctx->stats.kernel_launches += 1;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->copy_done[b], ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->copy_done[b], 0));
    if (ctx->stage_busy[b]) CUDA_TRY(ctx, cudaEventRecord(ctx->stage_copied[b], ctx->copy_stream));

    BatchPlan* plan = plan_ptr(ctx, slot);
    const BatchScans* scans = slot_scans(ctx, slot);
    u64 launches = 0;
    const LaunchHook* hook = ctx->profiling ? &ctx->hook : nullptr;
    // this slot's plan / scan table / sorted points / records / descriptors were last used by the batch of n_slots batches ago: its fold
    // (which follows its ray walk and its descriptor sort) must have been launched ...
    while (ctx->n_pend && ctx->pend[0].slot == slot) { bool launched; TRY(complete_one_fold(ctx, true, &launched)); }
    if (ctx->n_pend == ctx->n_slots) { bool launched; TRY(complete_one_fold(ctx, true, &launched)); }
    // the gathers due in front of this batch (the closing batch's fold has been launched by now: it used this slot, or an older one)
    while (!ctx->sh.gather_at.empty() && ctx->sh.gather_at.front() <= ctx->sh.front_seq) TRY(shard_gather_now(ctx));
    // ... and the device waits for it (on fold_stream) before touching the slot
    if (ctx->fold_done_valid[slot]) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->fold_done[slot], 0));
    if (ctx->scans_uploaded_valid[b]) CUDA_TRY(ctx, cudaEventSynchronize(ctx->scans_uploaded[b]));  // (two batches ago: long done)
    batch_scans_tiles(ctx->h_scans, ns);
    *ctx->h_scans_pinned[b] = ctx->h_scans;
    // (sharded: this is the table of the whole batch; the table of this rank's own points -- slot_scans -- is made by the filter)
    BatchScans* scans_dst = ctx->sh.world > 1 ? ctx->sh.batch_scans[slot].as<BatchScans>() : slot_scans(ctx, slot);
    CUDA_TRY(ctx, cudaMemcpyAsync(scans_dst, ctx->h_scans_pinned[b], sizeof(BatchScans), cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaEventRecord(ctx->scans_uploaded[b], s));
    ctx->scans_uploaded_valid[b] = true;
    queue_point_stage(ctx, slot, b, n, ns, ctx->h_scans);
    CUDA_TRY(ctx, cudaEventRecord(ctx->xyz_free[b], s));  // the point stage is the only reader of d_xyz[b]
    ctx->xyz_free_valid[b] = true;
    const bool use_runs = ctx->pair_path == 2 && n <= runs_max_batch_points() && ctx->rb[0].capacity != 0;
    const bool use_blocks = !use_runs && ctx->pair_path != 1 && n <= blocks_max_batch_points();
    if (ctx->sh.world > 1 && !use_runs)
        return fail(ctx, CHAD_ERR_INVALID, "a sharded map needs the tile-run pair path (truncation / voxel size <= 3.5, batches of at most 2^23 points)");
    if (!use_blocks && !use_runs) {
        PROF(ctx, PC_BAND_COUNT, launch_band_count(s, slot_xyz_sorted(ctx, slot), n, scans, ctx->mp, plan, ctx->counts.as<u32>()));
        PROF(ctx, PC_BAND_SCAN, (exclusive_scan<u32, u32>(s, ctx->counts.as<u32>(), ctx->offsets.as<u32>(), n, ctx->scan_ws.p, (u32*)nullptr,
                                                          plan_field<u32>(ctx, slot, offsetof(BatchPlan, n_pairs)))));
    }
    ctx->stats.kernel_launches += launches;
    launches = 0;
    // ---- the previous batch's fold. The tile-run path keeps its updates in per-slot buffers, so the fold is launched only if its
    //      front has reported already (the host does not wait here: it goes on to queue this batch and to copy the next scans);
    //      the other paths reuse the pair buffers and fold on this stream: everything before them must be queued first ----
    //      (so does a tile-run batch that follows a batch of another path, e.g. a scan beyond the tile-run path's 2^23 points: its
    //      sorted updates sit in the record buffers until its fold has been queued)
    bool other_path_pending = false;
    for (int q = 0; q < ctx->n_pend; q++) other_path_pending |= !ctx->pend[q].runs;
    if (use_runs && !other_path_pending) TRY(poll_folds(ctx)); else TRY(complete_pending_fold(ctx));
    // ---- pair stage ----
    const size_t max_pairs = size_t(n) * ctx->mp.max_ray_voxels;
    if (use_runs) {
        // overlap_walk: the walk only reads what the point stage left in this slot's buffers, so it runs on its own stream and the main
        // stream goes straight on to the next batch's point stage (which writes the other slot's buffers)
        cudaStream_t ws = ctx->overlap_walk ? ctx->walk_stream : s;
        if (ctx->overlap_walk) {
            CUDA_TRY(ctx, cudaEventRecord(ctx->points_done[slot], s));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ws, ctx->points_done[slot], 0));
            ctx->prof_stream = ws;
        }
        const u32 order_rank = ctx->sh.world > 1 ? u32(ctx->sh.world - 1 - ctx->sh.rank) << batch_tsb(ctx->h_scans, ns) : 0u;
        launches += launch_runs_emit(ws, slot_xyz_sorted(ctx, slot), slot_normals(ctx, slot), n, ns, scans, ctx->mp, plan, ctx->rb[slot],
                                     slot_records(ctx, slot), (u32)ctx->cap_pairs, order_rank, hook, PC_RUNS_EMIT);
        // the descriptor sort and the block list are only needed by the fold: they go on their own stream
        CUDA_TRY(ctx, cudaEventRecord(ctx->emit_done, ws));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->group_stream, ctx->emit_done, 0));
        ctx->prof_stream = ctx->group_stream;
        if (ctx->sh.world > 1) {
            prof_begin(ctx, PC_SHARD_EXCHANGE);
            TRY(queue_shard_exchange(ctx, slot));
            prof_end(ctx);
        }
        launches += launch_runs_group(ctx->group_stream, ctx->sh.world > 1 ? size_t(ctx->rb[slot].capacity) : runs_max_runs(n, ns), plan, ctx->rb[slot], ctx->rws2,
                                      ctx->num_sms, hook, PC_RUNS_SORT);
        ctx->prof_stream = nullptr;
        ctx->fold_in_flight = true;
    } else if (ctx->fold_in_flight) {
        // the other pair paths use both pair buffers and fold on this stream: order them after the folds still in flight
        for (int b = 0; b < MAX_SLOTS; b++)
            if (ctx->fold_done_valid[b]) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->fold_done[b], 0));
    }
    if (use_runs) {
        // (queued above)
    } else if (use_blocks) {
        launches += launch_blocks_pairs(s, slot_xyz_sorted(ctx, slot), slot_normals(ctx, slot), n, scans, ctx->mp, plan, ctx->bt, ctx->scan_ws.p,
                                        ctx->keys_a.as<u64>(), ctx->keys_b.as<u64>(), ctx->vals_a.as<u32>(), ctx->vals_b.as<u32>(), (u32)ctx->cap_pairs,
                                        ctx->num_sms, hook, PC_BLOCKS_COUNT, PC_BLOCKS_SCAN, PC_BLOCKS_EMIT, PC_BLOCKS_SORT);
    } else {
        PROF(ctx, PC_BAND_EMIT, launch_band_emit(s, slot_xyz_sorted(ctx, slot), slot_normals(ctx, slot), n, scans, ctx->mp, plan,
                                                 ctx->offsets.as<u32>(), ctx->keys_a.as<u64>(), ctx->vals_a.as<u32>(), (u32)ctx->cap_pairs, false));
        launches += radix_sort_pairs(s, ctx->keys_a.as<u64>(), ctx->vals_a.as<u32>(), ctx->keys_b.as<u64>(), ctx->vals_b.as<u32>(),
                                     plan_field<u32>(ctx, slot, offsetof(BatchPlan, n_pairs)), plan_field<u32>(ctx, slot, offsetof(BatchPlan, nbits_pairs)),
                                     max_pairs, RS_MAX_PASSES, ctx->rws, ctx->num_sms, hook, PC_PAIR_SORT_HIST);
        PROF(ctx, PC_SEGMENT_COUNT, launch_segment_count(s, ctx->keys_a.as<u64>(), ctx->keys_b.as<u64>(), (u32)max_pairs, plan, ctx->num_sms));
    }
    cudaStream_t plan_on = use_runs ? ctx->group_stream : s;
    CUDA_TRY(ctx, cudaMemcpyAsync(&ctx->h_plan[slot], plan, sizeof(BatchPlan), cudaMemcpyDeviceToHost, plan_on));
    ctx->stats.d2h_bytes += sizeof(BatchPlan);
    CUDA_TRY(ctx, cudaEventRecord(ctx->front_done[slot], plan_on));
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->stats.kernel_launches += launches;
    ctx->stats.batches++;
    ctx->burst_batches++;
    ctx->pend[ctx->n_pend++] = chad_ctx::PendingFold{slot, use_runs, (u32)max_pairs, false, 0};
    ctx->sh.front_seq++;
    ctx->plan_slot = (ctx->plan_slot + 1) % ctx->n_slots;
    ctx->cur ^= 1;
    ctx->batch_points = 0;
    ctx->batch_scans = 0;
    return CHAD_OK;
}

int finalize_part2(chad_ctx* ctx);

int drain(chad_ctx* ctx) {
    trace(ctx, "drain (%d folds pending, finalize state %d)", ctx->n_pend, ctx->fin_state);
    TRY(process_front(ctx));
    TRY(complete_pending_fold(ctx));
    while (!ctx->sh.gather_at.empty() || ctx->sh.close_deferred) {
        if (ctx->sh.close_deferred) TRY(try_deferred_close(ctx, true));
        if (!ctx->sh.gather_at.empty()) TRY(shard_gather_now(ctx));
    }
    if (ctx->fin_state == chad_ctx::FIN_PART1) {  // let part 2 of an in-flight finalize overlap the tail of the compute stream
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->fin_p1_done));
        TRY(finalize_part2(ctx));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->walk_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->group_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fold_stream));
    ctx->fold_in_flight = false;
    ctx->burst_batches = 0;
    if (!ctx->graveyard.empty() && ctx->fin_state == chad_ctx::FIN_IDLE) {  // (a finalize still in flight may be copying out of a parked buffer)
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fin_stream));
        for (void* q : ctx->graveyard) cudaFree(q);
        ctx->graveyard.clear();
    }
    fold_bounds_reset(ctx);
    account_fold_stats(ctx);
    prof_resolve(ctx);
    ctx->table_count_known = *ctx->h_table_count;
    ctx->stats.resident_clusters = ctx->table_count_known;
    // deferred flags raised by the fold (either plan slot)
    u32 flags = 0;
    for (int slot = 0; slot < MAX_SLOTS; slot++) {
        u32 f = 0;
        CUDA_TRY(ctx, cudaMemcpy(&f, plan_field<u32>(ctx, slot, offsetof(BatchPlan, error)), 4, cudaMemcpyDeviceToHost));
        if (f) CUDA_TRY(ctx, cudaMemset(plan_field<u32>(ctx, slot, offsetof(BatchPlan, error)), 0, 4));
        flags |= f;
    }
    if (flags) return error_from_flags(ctx, flags);
    return CHAD_OK;
}

int finalize_wait(chad_ctx* ctx);
int finalize_poll(chad_ctx* ctx);
// everything queued has been applied AND any asynchronous Submap::finalize has completed
int settle(chad_ctx* ctx) {
    TRY(drain(ctx));
    return finalize_wait(ctx);
}

// ---- finalize -----------------------------------------------------------------------------
// Submap::finalize runs asynchronously on its own stream, in two parts, so that neither the device nor the host
// waits for it and the next submap's inserts overlap it:
//   begin  (at the submap switch): the active chunk table is swapped with the spare one; fin_stream waits for the
//          compute stream's last fold into the old table.
//   part 1 (queued at begin): compact + sort + gather the old table, build and dedup the leaf clusters, count the
//          nodes every level will receive, copy those 20 counts to pinned host memory.
//   part 2 (queued by the next API call that finds part 1 complete -- every insert polls): with the exact counts the
//          host sizes every level (no worst-case allocation) and queues the 20 node levels back to back; per-level
//          results accumulate in device memory and come back in one copy; the old table is cleared.
//   finish (next API call that finds part 2 complete, or any call that needs the DAG): host mirrors are updated.
enum { SC_COUNT = 0, SC_RMAX = 1, SC_NBITS = 2, SC_PARENTS = 3, SC_NEW32 = 4, SC_ERR = 5, SC_BAR = 6, SC_ROOT = 8 /* u32[2] */, SC_LEVELS = 16 /* u32[20] */ };
u32* scalar32(chad_ctx* ctx, int i) { return ctx->f_scalars.as<u32>() + i; }

int ensure_finalize_capacity(chad_ctx* ctx, size_t chunks) {
    if (chunks <= ctx->cap_chunks && ctx->cap_chunks > 0) return CHAD_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fin_stream));
    const size_t nc = chunks + chunks / 8 + 1024;
    for (int i = 0; i < 2; i++) {
        TRY(dev_ensure(ctx, ctx->f_ids[i], nc * 8));
        if (i == 0) TRY(dev_ensure(ctx, ctx->f_sorted, nc * 8));
        TRY(dev_ensure(ctx, ctx->f_slots[i], nc * 4));
        TRY(dev_ensure(ctx, ctx->f_addr[i], nc * 2 * 4));
    }
    TRY(dev_ensure(ctx, ctx->f_cells, nc * 64));
    TRY(dev_ensure(ctx, ctx->f_tsdf, nc * 8));
    TRY(dev_ensure(ctx, ctx->f_head_rank, nc * 4));
    TRY(dev_ensure(ctx, ctx->f_cand, nc * 2 * 9 * 4));
    TRY(dev_ensure(ctx, ctx->f_slot_of, (nc * 2 + 2) * 4));
    TRY(dev_ensure(ctx, ctx->f_is_new, (nc * 2 + 2) * 8));
    TRY(dev_ensure(ctx, ctx->f_rank, (nc * 2 + 2) * 8));
    TRY(dev_ensure(ctx, ctx->f_radix_ws, radix_workspace_bytes(nc)));
    TRY(dev_ensure(ctx, ctx->f_scan_ws, scan_workspace_bytes(nc * 2 + 2)));
    ctx->f_rws = radix_workspace_carve(ctx->f_radix_ws.p, nc);
    ctx->cap_chunks = nc;
    return CHAD_OK;
}

// only called while the finalize stream is idle or between its kernels from the host's point of view: growth
// synchronises fin_stream, the only stream that touches the DAG levels
int level_reserve(chad_ctx* ctx, Level& L, bool cluster, size_t new_records) {
    cudaStream_t s = ctx->fin_stream;
    const size_t word = cluster ? 8 : 4;
    const size_t need_words = cluster ? (size_t(L.uniques) + new_records + 2) : (size_t(L.occupied) + 9 * new_records + 9);
    if (need_words >= (1ull << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "DAG level exceeds 2^31 words");
    if (need_words > L.raw_cap) {
        const size_t cap = next_pow2(need_words);  // (the callers' record counts are generous upper bounds already)
        const auto t_grow = std::chrono::steady_clock::now();
        void* np = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&np, cap * word));
        if (L.raw.p) {
            // stream order is enough: every later reader / writer of the level runs on fin_stream behind this copy
            CUDA_TRY(ctx, cudaMemcpyAsync(np, L.raw.p, L.raw.bytes, cudaMemcpyDeviceToDevice, s));
            ctx->graveyard.push_back(L.raw.p);
        }
        L.raw.p = np;
        L.raw.bytes = cap * word;
        L.raw_cap = cap;
        ctx->grow_events++;
        ctx->grow_bytes += cap * word;
        ctx->grow_host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_grow).count();
    }
    const size_t need_slots = (size_t(L.uniques) + new_records) * 2;
    if (need_slots > L.table.capacity) {
        const u64 cap = next_pow2(need_slots);
        void *ne = nullptr, *nf = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&ne, cap * 8));
        CUDA_TRY(ctx, cudaMalloc(&nf, cap * 4));
        DedupTable nt{static_cast<u64*>(ne), static_cast<u32*>(nf), cap};
        if (L.table.capacity) {
            ctx->stats.kernel_launches += launch_dedup_rehash(s, L.table, nt, ctx->num_sms);
            ctx->graveyard.push_back(L.entries.p);
            ctx->graveyard.push_back(L.first.p);
            L.entries = DevBuf{}; L.first = DevBuf{};
            ctx->grow_events++;
            ctx->grow_bytes += cap * 12;
        } else {
            launch_dedup_clear(s, nt);
        }
        dev_free(L.entries); dev_free(L.first);
        L.entries.p = ne; L.entries.bytes = cap * 8;
        L.first.p = nf; L.first.bytes = cap * 4;
        L.table = nt;
    }
    return CHAD_OK;
}

// compact + sort + gather the chunks of table `t` on stream `s`: f_ids[0] = full chunk keys ascending, f_cells = their
// cells; the exact count stays in device memory (SC_COUNT). max_chunks = host upper bound.
int queue_sorted_chunks(chad_ctx* ctx, cudaStream_t s, const ChunkTable& t, u32 max_chunks) {
    u64 launches = 0;
    launches += launch_table_compact(s, t, max_chunks, ctx->f_ids[0].as<u64>(), ctx->f_slots[0].as<u32>(), scalar32(ctx, SC_COUNT), scalar32(ctx, SC_RMAX),
                                     scalar32(ctx, SC_NBITS), ctx->num_sms);
    launches += radix_sort_pairs(s, ctx->f_ids[0].as<u64>(), ctx->f_slots[0].as<u32>(), ctx->f_ids[1].as<u64>(), ctx->f_slots[1].as<u32>(),
                                 scalar32(ctx, SC_COUNT), scalar32(ctx, SC_NBITS), max_chunks, RS_MAX_PASSES, ctx->f_rws, ctx->num_sms);
    // the gather writes f_ids[0]: when the sort left its result there, it reads it from a copy
    launches += launch_chunk_gather(s, t, ctx->f_ids[0].as<u64>(), ctx->f_ids[1].as<u64>(), scalar32(ctx, SC_COUNT), scalar32(ctx, SC_RMAX),
                                    scalar32(ctx, SC_NBITS), max_chunks, ctx->f_sorted.as<u64>(), ctx->f_cells.p);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->f_ids[0].p, ctx->f_sorted.p, size_t(max_chunks) * 8, cudaMemcpyDeviceToDevice, s));
    ctx->stats.kernel_launches += launches;
    CUDA_TRY(ctx, cudaGetLastError());
    return CHAD_OK;
}

int finalize_part2(chad_ctx* ctx);
int finalize_finish(chad_ctx* ctx);

int finalize_gather(chad_ctx* ctx);
// non-blocking progress of an in-flight finalize (called from every API entry)
int finalize_poll(chad_ctx* ctx) {
    if (ctx->sh.world == 1 && ctx->fin_state == chad_ctx::FIN_PART1 && cudaEventQuery(ctx->fin_p1_done) == cudaSuccess) TRY(finalize_part2(ctx));
    if (ctx->fin_state == chad_ctx::FIN_PART2 && cudaEventQuery(ctx->fin_done) == cudaSuccess) TRY(finalize_finish(ctx));
    cudaGetLastError();  // cudaErrorNotReady is not an error
    return CHAD_OK;
}
// blocking completion
int finalize_wait(chad_ctx* ctx) {
    if (ctx->fin_state == chad_ctx::FIN_PART1) {
        if (ctx->sh.world > 1) return fail(ctx, CHAD_ERR_INVALID, "internal: a sharded finalize was waited for before its gather point");
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->fin_p1_done));
        TRY(finalize_part2(ctx));
    }
    if (ctx->fin_state == chad_ctx::FIN_PART2) {
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->fin_done));
        TRY(finalize_finish(ctx));
    }
    return CHAD_OK;
}

// Close the active submap: everything of it must already be queued (process_front + complete_pending_fold).
// The tables are swapped right away, so the next submap's inserts go on; the finalize itself is queued as soon as the
// closed submap's exact chunk count has arrived on the host (count_stream: the stream its last fold and the copy of the
// table counter were queued on) -- by the next API call that finds it there, or by whoever needs the DAG.
// external = true (sharded mode): f_ids[0] / f_cells already hold `max_chunks` globally sorted chunks.
int finalize_begin(chad_ctx* ctx, u32 max_chunks, bool external, cudaStream_t count_stream) {
    trace(ctx, "finalize_begin (previous finalize in state %d)", ctx->fin_state);
    const bool sharded = ctx->sh.world > 1 && !external;
    if (sharded) {
        // the spare table is free as soon as the submap closed before has been gathered out of it (device: table_free); its DAG stage may
        // still be running -- it works on the gathered chunk stream
        if (ctx->sh.closed_waiting) return fail(ctx, CHAD_ERR_INVALID, "internal: close while the spare table still holds a closed submap");
        if (ctx->sh.table_free_valid) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->fold_stream, ctx->sh.table_free, 0));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->sh.table_free, 0));
        }
    } else {
        TRY(finalize_wait(ctx));  // one finalize in flight at a time
    }
    cudaStream_t fs = ctx->fin_stream;
    CUDA_TRY(ctx, cudaEventRecord(ctx->submap_closed, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(fs, ctx->submap_closed, 0));
    CUDA_TRY(ctx, cudaEventRecord(ctx->submap_closed2, ctx->fold_stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(fs, ctx->submap_closed2, 0));
    ctx->fin_external = external;
    ctx->fin_max_chunks = max_chunks;
    if (external) {
        ctx->fin_state = chad_ctx::FIN_PART1;
        return finalize_part2(ctx);
    }
    // swap tables: the spare one was cleared at the end of the previous finalize (complete: see finalize_wait above)
    std::swap(ctx->table, ctx->table2);
    std::swap(ctx->t_keys, ctx->t2_keys);
    std::swap(ctx->t_cells, ctx->t2_cells);
    std::swap(ctx->t_count, ctx->t2_count);
    std::swap(ctx->t_list, ctx->t2_list);
    std::swap(ctx->h_table_count, ctx->h_table_count2);  // the closed table's last count copy lands in h_table_count2
    *ctx->h_table_count = 0;
    ctx->table_count_known = 0;
    fold_bounds_reset(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->fin_p1_done, count_stream));
    if (sharded) { ctx->sh.closed_waiting = true; return CHAD_OK; }
    ctx->fin_state = chad_ctx::FIN_PART1;
    return CHAD_OK;
}

// The DAG stage of Submap::finalize on fin_stream: f_ids[0] / f_cells hold C ascending chunks (or are about to: `chunks` queues whatever
// produces them, after the buffers have been sized), *SC_COUNT = C.
template <typename QueueChunks>
int finalize_dag(chad_ctx* ctx, u32 C, bool count_from_host, QueueChunks chunks) {
    cudaStream_t fs = ctx->fin_stream;
    ctx->fin_chunks = C;
    ctx->fin_state = chad_ctx::FIN_IDLE;  // (until everything is queued: the reserves below may synchronise fin_stream)
    if (C >= (1u << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "submap exceeds 2^31 leaf chunks");
    TRY(ensure_finalize_capacity(ctx, C));
    Level& LC = ctx->levels[CHAD_LEVEL_CLUSTERS];
    TRY(level_reserve(ctx, LC, true, size_t(C) + 1));
    for (int d = 0; d < 20; d++) {
        // level d holds at most 8^d nodes and at most one node per leaf cluster; two records (TSDF, weight) per node
        const u64 nodes = std::max<u64>(1, (d < 10) ? std::min<u64>(C, 1ull << (3 * d)) : C);
        TRY(level_reserve(ctx, ctx->levels[d], false, size_t(2 * nodes)));
    }
    if (ctx->profiling) cudaEventRecord(ctx->fin_t0, fs);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->f_scalars.p, 0, 256, fs));
    u64 launches = 0;
    TRY(chunks());
    if (count_from_host) {
        ctx->h_fin->h2d[0] = C;
        CUDA_TRY(ctx, cudaMemcpyAsync(scalar32(ctx, SC_COUNT), &ctx->h_fin->h2d[0], 4, cudaMemcpyHostToDevice, fs));
    }
    if (C) {
        launches += launch_cluster_build(fs, ctx->f_cells.p, scalar32(ctx, SC_COUNT), C, ctx->mp, ctx->f_tsdf.as<u64>());
        launches += launch_cluster_dedup(fs, LC.table, ctx->f_tsdf.as<u64>(), scalar32(ctx, SC_COUNT), C, LC.raw.as<u64>(), LC.uniques,
                                         ctx->f_slot_of.as<u32>(), ctx->f_is_new.as<u32>(), ctx->f_rank.as<u32>(), ctx->f_scan_ws.p,
                                         ctx->f_addr[0].as<u32>(), scalar32(ctx, SC_NEW32), scalar32(ctx, SC_ERR));
    }
    LevelsArgs la{};
    for (int d = 0; d < 20; d++) {
        Level& L = ctx->levels[d];
        la.lv[d] = LevelDev{L.table.entries, L.table.first, L.table.capacity, L.raw.as<u32>()};
    }
    la.counters = ctx->f_counters.as<LevelCounters>();
    la.d_chunks = scalar32(ctx, SC_COUNT);
    la.ids[0] = ctx->f_ids[0].as<u64>(); la.ids[1] = ctx->f_ids[1].as<u64>();
    la.addr[0] = ctx->f_addr[0].as<u32>(); la.addr[1] = ctx->f_addr[1].as<u32>();
    la.head_rank = ctx->f_head_rank.as<u32>();
    la.cand = ctx->f_cand.as<u32>();
    la.slot_of = ctx->f_slot_of.as<u32>();
    la.rank = ctx->f_rank.as<u64>();
    la.partial = ctx->f_partial.as<u64>();
    la.bar = scalar32(ctx, SC_BAR);
    la.d_error = scalar32(ctx, SC_ERR);
    la.root_out = scalar32(ctx, SC_ROOT);
    la.level_nodes = scalar32(ctx, SC_LEVELS);
    launches += launch_dag_levels(fs, la, ctx->num_sms);
    ctx->stats.kernel_launches += launches;
    return CHAD_OK;
}

// the tail of a finalize on fin_stream: results to the host, the closed submap's table cleared
int finalize_tail(chad_ctx* ctx, bool clear_table2) {
    cudaStream_t fs = ctx->fin_stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_fin->scalars, ctx->f_scalars.p, 16 * 4 + 20 * 4, cudaMemcpyDeviceToHost, fs));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_fin->counters, ctx->f_counters.p, sizeof(LevelCounters) * 20, cudaMemcpyDeviceToHost, fs));
    ctx->stats.d2h_bytes += 16 * 4 + 20 * 4 + sizeof(LevelCounters) * 20;
    if (clear_table2) launch_table_clear(fs, ctx->table2);  // octree.clear(), tsdf.cpp:57 (external: done by the caller's stream)
    if (ctx->profiling) cudaEventRecord(ctx->fin_t3, fs);
    CUDA_TRY(ctx, cudaEventRecord(ctx->fin_done, fs));
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->fin_state = chad_ctx::FIN_PART2;
    return CHAD_OK;
}

// the exact chunk count is on the host: size everything and queue the whole finalize on fin_stream
int finalize_part2(chad_ctx* ctx) {
    cudaStream_t fs = ctx->fin_stream;
    const u32 C = ctx->fin_external ? ctx->fin_max_chunks : *ctx->h_table_count2;
    if (ctx->sh.world > 1 && !ctx->fin_external) return fail(ctx, CHAD_ERR_INVALID, "internal: sharded finalize outside its gather point");
    trace(ctx, "finalize: submap %zu has %u chunks", ctx->roots.size(), C);
    TRY(finalize_dag(ctx, C, ctx->fin_external, [&]() -> int {
        if (C && !ctx->fin_external) TRY(queue_sorted_chunks(ctx, fs, ctx->table2, C));
        return CHAD_OK;
    }));
    return finalize_tail(ctx, !ctx->fin_external);
}

// sharded, at the gather point of a closed submap (see Shard::gather_at). The ranges ascend with the rank, so the concatenation of the ranks' sorted chunks in
// rank order is the submap's chunk stream in ascending Morton order (submap.hpp:10-106 walks the octree in that order): every rank
// sorts its chunks and sends them to rank 0, which receives them behind its own and runs the DAG stage; the two root addresses
// (submap.hpp:108-109) reach the other ranks at the next flush.
int finalize_gather(chad_ctx* ctx) {
    cudaStream_t fs = ctx->fin_stream;
    chad_ctx::Shard& sh = ctx->sh;
    {   // every rank's chunk count (the sizes of the gather must be known on the host). Blocking, but every rank is at the same point of
        // its call sequence: nobody waits for more than the others' skew
        if (!sh.closed_waiting) return fail(ctx, CHAD_ERR_INVALID, "internal: gather without a closed submap");
        TRY(finalize_wait(ctx));  // the DAG stage of the submap before: its work buffers and the host mirrors of the level counters are needed now
        // An NCCL kernel spins on the device until its peer kernel runs, and CUDA maps streams onto a limited number of hardware queues
        // (CUDA_DEVICE_MAX_CONNECTIONS), so a spinning kernel can hold back kernels submitted LATER to other streams. With two
        // communicators that can close a cycle: A's receive waits for B's send, queued behind B's next batch exchange, which waits for A's
        // next exchange, queued behind A's receive. Work submitted EARLIER cannot be held back, and every rank submits the gather at the
        // same point of the exchange sequence, so the one thing to forbid is a LATER exchange overtaking the gather: the group stream
        // waits (on the device, see queue_shard_exchange) for this rank's part of the gather before the next exchange. The host waits
        // only for the counts.
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->fin_p1_done));
        const u32 C = *ctx->h_table_count2;
        trace(ctx, "gather point: own chunk count %u, all-gather of the counts", C);
        u32* d = sh.scalars.as<u32>();
        ctx->h_fin->h2d[1] = C;
        CUDA_TRY(ctx, cudaMemcpyAsync(d + 8, &ctx->h_fin->h2d[1], 4, cudaMemcpyHostToDevice, fs));
        NCCL_TRY(ctx, sh.nccl->AllGather(d + 8, d + 16, 1, ncclUint32, sh.comm_f, fs));
        CUDA_TRY(ctx, cudaMemcpyAsync(sh.h_counts, d + 16, size_t(sh.world) * 4, cudaMemcpyDeviceToHost, fs));
        CUDA_TRY(ctx, cudaStreamSynchronize(fs));
        ctx->stats.kernel_launches += 1;
    }
    sh.gather_at.pop_front();
    sh.closes_gathered++;
    sh.closed_waiting = false;
    auto release_table = [&]() -> int {  // octree.clear(), tsdf.cpp:57 -- as soon as the chunks are out, so that the next close can swap
        launch_table_clear(fs, ctx->table2);
        CUDA_TRY(ctx, cudaEventRecord(sh.table_free, fs));
        sh.table_free_valid = true;
        return CHAD_OK;
    };
    u64 total = 0;
    u64 offset[SHARD_WORLD_MAX + 1];
    for (int g = 0; g < sh.world; g++) { offset[g] = total; total += sh.h_counts[g]; }
    const u32 own = sh.h_counts[sh.rank];
    trace(ctx, "finalize: counts arrived (total %llu), gather on rank 0", (unsigned long long)total);
    if (total >= (1ull << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "submap exceeds 2^31 leaf chunks");
    if (sh.rank == 0) {
        TRY(finalize_dag(ctx, (u32)total, true, [&]() -> int {
            if (own) TRY(queue_sorted_chunks(ctx, fs, ctx->table2, own));
            TRY(release_table());
            bool any = false;
            for (int g = 1; g < sh.world; g++) any |= sh.h_counts[g] != 0;
            if (any) {
                NCCL_TRY(ctx, sh.nccl->GroupStart());
                for (int g = 1; g < sh.world; g++) {
                    const size_t c = sh.h_counts[g];
                    if (!c) continue;
                    NCCL_TRY(ctx, sh.nccl->Recv(ctx->f_ids[0].as<u64>() + offset[g], c, ncclUint64, g, sh.comm_f, fs));
                    NCCL_TRY(ctx, sh.nccl->Recv(static_cast<u64*>(ctx->f_cells.p) + offset[g] * 8, c * 8, ncclUint64, g, sh.comm_f, fs));
                }
                NCCL_TRY(ctx, sh.nccl->GroupEnd());
                ctx->stats.kernel_launches += 1;
            }
            CUDA_TRY(ctx, cudaEventRecord(sh.counts_done, fs));  // the transfers are complete here; the DAG stage behind them runs on
            sh.xfer_pending = true;
            return CHAD_OK;
        }));
    } else {
        ctx->fin_chunks = own;
        ctx->fin_state = chad_ctx::FIN_IDLE;
        TRY(ensure_finalize_capacity(ctx, own));
        if (ctx->profiling) cudaEventRecord(ctx->fin_t0, fs);
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->f_scalars.p, 0, 256, fs));
        if (own) TRY(queue_sorted_chunks(ctx, fs, ctx->table2, own));
        TRY(release_table());
        if (own) {
            NCCL_TRY(ctx, sh.nccl->GroupStart());
            NCCL_TRY(ctx, sh.nccl->Send(ctx->f_ids[0].p, own, ncclUint64, 0, sh.comm_f, fs));
            NCCL_TRY(ctx, sh.nccl->Send(ctx->f_cells.p, size_t(own) * 8, ncclUint64, 0, sh.comm_f, fs));
            NCCL_TRY(ctx, sh.nccl->GroupEnd());
            ctx->stats.kernel_launches += 1;
            CUDA_TRY(ctx, cudaEventRecord(sh.counts_done, fs));
            sh.xfer_pending = true;
        }
    }
    // (the roots reach the other ranks at the next flush: shard_sync_roots)
    trace(ctx, "finalize: gather + DAG stage queued");
    return finalize_tail(ctx, false);
}

// sharded, collective (chad_flush / chad_finalize_active): the roots of the submaps closed since the last flush, from rank 0 to all
int shard_sync_roots(chad_ctx* ctx) {
    chad_ctx::Shard& sh = ctx->sh;
    const size_t n = ctx->roots.size();  // the same on every rank: closes are counted by the common call sequence
    if (sh.world == 1 || n == sh.roots_synced) return CHAD_OK;
    const size_t fresh = n - sh.roots_synced;
    if (fresh > 4096) return fail(ctx, CHAD_ERR_CAPACITY, "more than 4096 submaps closed between two flushes of a sharded map");
    cudaStream_t fs = ctx->fin_stream;
    DevBuf d;
    TRY(dev_ensure(ctx, d, fresh * 8));
    if (sh.rank == 0) {
        for (size_t i = 0; i < fresh; i++) { sh.h_roots[2 * i] = ctx->roots[sh.roots_synced + i][0]; sh.h_roots[2 * i + 1] = ctx->roots[sh.roots_synced + i][1]; }
        CUDA_TRY(ctx, cudaMemcpyAsync(d.p, sh.h_roots, fresh * 8, cudaMemcpyHostToDevice, fs));
    }
    NCCL_TRY(ctx, sh.nccl->Broadcast(d.p, d.p, fresh * 2, ncclUint32, 0, sh.comm_f, fs));
    if (sh.rank != 0) CUDA_TRY(ctx, cudaMemcpyAsync(sh.h_roots, d.p, fresh * 8, cudaMemcpyDeviceToHost, fs));
    CUDA_TRY(ctx, cudaStreamSynchronize(fs));
    if (sh.rank != 0)
        for (size_t i = 0; i < fresh; i++) ctx->roots[sh.roots_synced + i] = {sh.h_roots[2 * i], sh.h_roots[2 * i + 1]};
    sh.roots_synced = n;
    ctx->graveyard.push_back(d.p);
    return CHAD_OK;
}

int finalize_finish(chad_ctx* ctx) {
    ctx->fin_state = chad_ctx::FIN_IDLE;
    trace(ctx, "finalize: done, submap %zu", ctx->roots.size());
    const u32* hs = ctx->h_fin->scalars;
    if (hs[SC_ERR]) return error_from_flags(ctx, hs[SC_ERR]);
    const u32 C = ctx->fin_chunks;
    Level& LC = ctx->levels[CHAD_LEVEL_CLUSTERS];
    const bool has_dag = ctx->sh.rank == 0 || ctx->fin_external;  // sharded: the levels live on rank 0, the others only learn the roots
    if (C && has_dag) {
        const u32 fresh = hs[SC_NEW32];
        LC.uniques += fresh;
        LC.dupes += 2 * C - fresh;  // levels.hpp:135-138
    }
    for (int d = 0; d < 20 && has_dag; d++) {
        Level& L = ctx->levels[d];
        L.uniques = ctx->h_fin->counters[d].uniques;   // levels.hpp:79-86, accumulated on the device
        L.dupes = ctx->h_fin->counters[d].dupes;
        L.occupied = ctx->h_fin->counters[d].occupied;
        ctx->fin_level_nodes[d] = ctx->h_fin->level_nodes[d];
    }
    ctx->roots.push_back({hs[SC_ROOT], hs[SC_ROOT + 1]});
    ctx->stats.submaps++;
    ctx->fin_external = false;
    if (ctx->profiling) {
        float a = 0.f;
        if (cudaEventElapsedTime(&a, ctx->fin_t0, ctx->fin_t3) == cudaSuccess) {
            ctx->prof_ms[PC_FINALIZE] += a;
            ctx->prof_launches[PC_FINALIZE]++;
        } else cudaGetLastError();
    }
    return CHAD_OK;
}

// Close the active submap. lazy = true (submap switch inside insert): if a batch of the submap is still in flight, only
// mark it; its fold and the finalize are queued by the next process_front / drain, so the host never waits for the
// device here. lazy = false (chad_finalize_active): queue everything now.
int finalize_submap(chad_ctx* ctx, bool lazy) {
    TRY(process_front(ctx));
    const u64 close_index = ++ctx->sh.closes_marked;
    if (ctx->sh.world > 1) ctx->sh.gather_at.push_back(ctx->sh.front_seq - 1 + (u64)ctx->n_slots);  // the closing batch is number front_seq - 1
    ctx->sh.need_splitters = true;  // (sharded) the next submap's ranges follow its own first scan
    ctx->positions.push_back(std::move(ctx->active_positions));  // closing order == the order the roots arrive in
    ctx->active_positions.clear();
    if (ctx->n_pend) {
        ctx->pend[ctx->n_pend - 1].close = true;  // the submap's last batch: the finalize begins right after its fold
        ctx->pend[ctx->n_pend - 1].close_index = close_index;
        if (lazy) return CHAD_OK;
        return complete_pending_fold(ctx);  // waits for the front of the last batch, queues its fold, begins the finalize
    }
    // every fold of the submap has been launched already: the copy of the table counter that followed the last one is in flight on
    // that fold's stream (or has arrived); the finalize is queued when it is there -- no host wait here either
    if (ctx->sh.close_deferred) TRY(try_deferred_close(ctx, true));
    while (ctx->sh.world > 1 && ctx->sh.closes_gathered + 1 < close_index) TRY(shard_gather_now(ctx));  // (the tables are about to be swapped)
    TRY(finalize_begin(ctx, 0, false, ctx->last_fold_stream ? ctx->last_fold_stream : ctx->stream));
    ctx->stats.resident_clusters = 0;
    return CHAD_OK;
}

// the gather of the oldest closed submap: its last fold has been launched (which began the finalize: tables swapped, count copy queued)
int shard_gather_now(chad_ctx* ctx) {
    if (ctx->sh.close_deferred && ctx->sh.deferred_index == ctx->sh.closes_gathered + 1) TRY(try_deferred_close(ctx, true));  // (it IS the oldest ungathered close)
    if (ctx->sh.gather_at.empty() || !ctx->sh.closed_waiting) return fail(ctx, CHAD_ERR_INVALID, "internal: no closed submap at a gather point");
    return finalize_gather(ctx);
}

int begin_scan(chad_ctx* ctx, size_t n, const float position[3], bool* skip) {
    *skip = false;
    if (ctx->sticky_error != CHAD_OK) return ctx->sticky_error;
    TRY(finalize_poll(ctx));
    TRY(poll_folds(ctx));
    // tsdf.cpp:46-61: a pose more than 5 m (strictly) from the submap's FIRST pose finalises the submap;
    // the triggering scan goes entirely into the new one
    const std::array<float, 3> pose{position[0], position[1], position[2]};
    if (!ctx->has_pose) {
        ctx->has_pose = true;
        std::memcpy(ctx->first_pose, position, 12);
    } else {
        const float dx = ctx->first_pose[0] - position[0], dy = ctx->first_pose[1] - position[1], dz = ctx->first_pose[2] - position[2];
        const float tx = dx * dx, ty = dy * dy, tz = dz * dz;
        volatile float sum = tx + ty;  // glm::distance = sqrt((x*x + y*y) + z*z), no contraction
        sum = sum + tz;
        if (std::sqrt((float)sum) > 5.0f) {
            TRY(finalize_submap(ctx, true));
            std::memcpy(ctx->first_pose, position, 12);
        }
    }
    ctx->active_positions.push_back(pose);  // tsdf.cpp:48,56,60: every scan's pose joins the (possibly new) active submap
    ctx->stats.scans++;
    ctx->stats.points += n;
    if (n == 0) { *skip = true; return CHAD_OK; }
    if (n >= (1ull << 31)) return fail(ctx, CHAD_ERR_INVALID, "scan too large");
    if (ctx->batch_scans > 0 && (ctx->batch_scans >= (u32)ctx->max_batch || ctx->batch_points + n > ctx->cap_points)) TRY(process_front(ctx));
    if (n > ctx->cap_points || ctx->cap_points == 0) {
        TRY(drain(ctx));
        size_t want = n > ctx->cap_points / 2 ? n * (size_t)ctx->max_batch : ctx->cap_points;
        // keep a batch inside the tile-run path's rank range (2^23 sorted points) unless a single scan is larger than that
        const size_t rank_cap = (size_t)runs_max_batch_points() - 4096;
        if (n * 9 / 8 + 1024 <= rank_cap && want * 9 / 8 + 1024 > rank_cap) want = (rank_cap - 1024) * 8 / 9;
        // ... and inside the 2^30 band-voxel limit of the pair buffers: a large scan gets a smaller batch (down to the scan alone)
        // instead of an error -- the reference accepts any scan size
        const size_t pair_cap = (((1ull << 30) - 1) / ctx->mp.max_ray_voxels - 1024) * 8 / 9 - 1;
        if (want > pair_cap && n <= pair_cap) want = pair_cap;
        TRY(ensure_batch_capacity(ctx, want));
    }
    return CHAD_OK;
}

// first scan of a batch into d_xyz[cur]: the transfers must not overtake the point stage of the batch that used the buffer before
// (the host no longer waits for that batch: it may be two batches ahead of the device)
int guard_batch_buffer(chad_ctx* ctx) {
    const int b = ctx->cur;
    if (ctx->batch_scans == 0 && ctx->xyz_free_valid[b]) {
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->xyz_free[b], 0));
    }
    return CHAD_OK;
}

int end_scan(chad_ctx* ctx, size_t n, const float position[3]) {
    ctx->h_scans.offset[ctx->batch_scans] = ctx->batch_points;
    std::memcpy(ctx->h_scans.pose[ctx->batch_scans], position, 12);
    ctx->batch_scans++;
    ctx->batch_points += (u32)n;
    // A burst starts with a short batch: while nothing is in flight the device would only wait for the host to copy a full batch
    // (24 scans = 1.4 ms over PCIe); once a batch is queued the following ones fill up behind it.
    // (sharded: every rank must cut the same batches, so the rule may not look at what happens to be in flight)
    const bool burst_start = ctx->sh.world > 1 ? ctx->burst_batches == 0 : ctx->n_pend == 0;
    const u32 target = burst_start ? std::min<u32>((u32)ctx->max_batch, ctx->first_batch) : (u32)ctx->max_batch;
    if (ctx->batch_scans >= target) TRY(process_front(ctx));
    return CHAD_OK;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

const char* chad_last_error(const chad_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

static int create_impl(float sdf_res, float sdf_trunc, int device, int max_batch_scans, int rank, int world, const void* shard_id, chad_ctx** out) {
    if (!out) return fail(nullptr, CHAD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > SHARD_WORLD_MAX || rank < 0 || rank >= world || (world > 1 && !shard_id))
        return fail(nullptr, CHAD_ERR_INVALID, "sharded map: 1 <= world <= 8, 0 <= rank < world, and the id of chad_shard_unique_id");
    const NcclApi* nccl = nullptr;
    if (world > 1) {
        const char* why = "";
        nccl = nccl_api(&why);
        if (!nccl) return fail(nullptr, CHAD_ERR_CUDA, std::string("sharded map: ") + why);
    }
    if (!(sdf_res > 0.0f) || !(sdf_trunc > 0.0f)) return fail(nullptr, CHAD_ERR_INVALID, "sdf_res and sdf_trunc must be positive");
    if (max_batch_scans < 0 || max_batch_scans > MAX_BATCH_SCANS) return fail(nullptr, CHAD_ERR_INVALID, "max_batch_scans must be in [0, 64]");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, CHAD_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return fail(nullptr, CHAD_ERR_INVALID, "device ordinal out of range");
    chad_ctx* ctx = new chad_ctx();
    ctx->device = device;
    ctx->hook = LaunchHook{ctx, prof_begin, prof_end};
    auto bail = [&](int code) { g_create_error = ctx->error; chad_destroy(ctx); return code; };
#define CREATE_TRY(expr)                                                                                               \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) { ctx->error = std::string(#expr) + ": " + cudaGetErrorString(_e); return bail(CHAD_ERR_CUDA); } \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        CREATE_TRY(cudaEventCreateWithFlags(&ctx->xyz_free[b], cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&ctx->scans_uploaded[b], cudaEventDisableTiming));
    }
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->fold_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->group_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ctx->walk_stream, cudaStreamNonBlocking));
    for (int b = 0; b < MAX_SLOTS; b++) CREATE_TRY(cudaEventCreateWithFlags(&ctx->points_done[b], cudaEventDisableTiming));
    for (int b = 0; b < MAX_SLOTS; b++) CREATE_TRY(cudaEventCreateWithFlags(&ctx->fold_done[b], cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&ctx->submap_closed2, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&ctx->emit_done, cudaEventDisableTiming));
    {   // the finalize stream gets the highest priority: its ~300 tiny dependent kernels then take the first SM slot that frees up
        // instead of queueing behind the thousands of CTAs of an insert kernel
        int prio_lo = 0, prio_hi = 0;
        CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CREATE_TRY(cudaStreamCreateWithPriority(&ctx->fin_stream, cudaStreamNonBlocking, prio_hi));
    }
    CREATE_TRY(cudaEventCreateWithFlags(&ctx->submap_closed, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&ctx->fin_p1_done, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&ctx->fin_done, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&ctx->fin_t0, &ctx->fin_t3}) CREATE_TRY(cudaEventCreate(e));
    CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_fin), sizeof(chad_ctx::FinHost)));
    std::memset(ctx->h_fin, 0, sizeof(chad_ctx::FinHost));
    for (int b = 0; b < 2; b++) {
        CREATE_TRY(cudaEventCreateWithFlags(&ctx->stage_copied[b], cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&ctx->copy_done[b], cudaEventDisableTiming));
        CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_scans_pinned[b]), sizeof(BatchScans)));
    }
    for (int b = 0; b < MAX_SLOTS; b++) CREATE_TRY(cudaEventCreateWithFlags(&ctx->front_done[b], cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreate(&ctx->t0));
    CREATE_TRY(cudaEventCreate(&ctx->t1));
    CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_plan), MAX_SLOTS * sizeof(BatchPlan)));
    CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_plan_fold), MAX_SLOTS * sizeof(BatchPlan)));
    CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_table_count), 64));
    CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_table_count2), 64));
    *ctx->h_table_count2 = 0;
    *ctx->h_table_count = 0;
    CREATE_TRY(radix_sort_init());
    CREATE_TRY(blocks_init());
    CREATE_TRY(runs_init());
    if (const char* env = std::getenv("CHAD_PAIR_PATH")) { const int m = std::atoi(env); if (m >= 0 && m <= 2) ctx->pair_path = m; }
    if (const char* env = std::getenv("CHAD_OVERLAP_WALK")) ctx->overlap_walk = std::atoi(env) != 0;
    if (const char* env = std::getenv("CHAD_FIRST_BATCH")) { const int v = std::atoi(env); ctx->first_batch = (u32)(v < 1 ? 1 : v); }
    ctx->n_slots = ctx->overlap_walk ? 3 : 2;
    if (const char* env = std::getenv("CHAD_PLAN_SLOTS")) { const int v = std::atoi(env); if (v >= 2 && v <= MAX_SLOTS) ctx->n_slots = v; }

    ctx->mp.res = sdf_res;
    ctx->mp.trunc = sdf_trunc;
    ctx->mp.recip = float(1.0 / double(sdf_res));
    ctx->mp.trunc_recip = 1.0f / sdf_trunc;
    const double ratio = double(sdf_trunc) / double(sdf_res);
    // a ray crosses at most 1 + sum_a(|dv_a|) voxels with |dv_a| <= 2*ratio*|dir_a| + 1 (octree.hpp:94-97,121-152)
    ctx->mp.max_ray_voxels = (u32)std::ceil(4.0 + 2.0 * std::sqrt(3.0) * ratio) + 2;
    ctx->mp.band_margin = (u32)std::ceil(ratio) + 3;
    {   // a ray takes at most L = ceil(2 ratio) + 1 steps along one axis, i.e. crosses at most ceil(L / 8) block faces per axis
        const u32 L = (u32)std::ceil(2.0 * ratio) + 1;
        ctx->mp.max_ray_runs = 1 + 3 * ((L + 7) / 8);
    }
    ctx->max_batch = max_batch_scans == 0 ? 24 : max_batch_scans;  // one batch per submap of ~21 scans on the bench trajectory

    int r = dev_ensure(ctx, ctx->d_scans, sizeof(BatchScans));
    for (int q = 0; q < MAX_SLOTS - 1; q++) if (r == CHAD_OK) r = dev_ensure(ctx, ctx->d_scans2[q], sizeof(BatchScans));
    if (r == CHAD_OK) r = dev_ensure(ctx, ctx->d_plan, MAX_SLOTS * sizeof(BatchPlan));
    if (r == CHAD_OK) r = dev_ensure(ctx, ctx->f_scalars, 256);
    if (r != CHAD_OK) return bail(r);
    CREATE_TRY(cudaMemsetAsync(ctx->d_plan.p, 0, MAX_SLOTS * sizeof(BatchPlan), ctx->stream));
    r = table_alloc(ctx, ctx->table, ctx->t_keys, ctx->t_cells, ctx->t_count, ctx->t_list, 1ull << 20);
    if (r == CHAD_OK) r = table_alloc(ctx, ctx->table2, ctx->t2_keys, ctx->t2_cells, ctx->t2_count, ctx->t2_list, 1ull << 20);
    if (r == CHAD_OK) r = dev_ensure(ctx, ctx->f_counters, sizeof(LevelCounters) * CHAD_NUM_LEVELS);
    if (r == CHAD_OK) r = dev_ensure(ctx, ctx->f_partial, 2 * 1024 * 8);
    if (r != CHAD_OK) return bail(r);
    // NodeLevel / LeafClusterLevel constructors reserve index 0 (levels.hpp:52-54,119-120)
    for (int d = 0; d < CHAD_NUM_LEVELS; d++) {
        Level& L = ctx->levels[d];
        const bool cluster = d == CHAD_LEVEL_CLUSTERS;
        L.occupied = cluster ? 0 : 1;
        r = level_reserve(ctx, L, cluster, 1024);
        if (r != CHAD_OK) return bail(r);
        CREATE_TRY(cudaMemsetAsync(L.raw.p, 0, 64, ctx->stream));
    }
    CREATE_TRY(cudaStreamSynchronize(ctx->stream));
    CREATE_TRY(cudaStreamSynchronize(ctx->fin_stream));
    r = level_counters_reset(ctx);
    if (r != CHAD_OK) return bail(r);
    if (world > 1) {
        chad_ctx::Shard& sh = ctx->sh;
        sh.rank = rank;
        sh.world = world;
        sh.nccl = nccl;
        r = dev_ensure(ctx, sh.splitters, 2 * (SHARD_WORLD_MAX + 1) * sizeof(u64));
        for (int q = 0; q < MAX_SLOTS; q++) if (r == CHAD_OK) r = dev_ensure(ctx, sh.batch_scans[q], sizeof(BatchScans));
        if (r == CHAD_OK) r = dev_ensure(ctx, sh.scalars, 256);
        size_t box_mb = world <= 2 ? 4 : (world <= 4 ? 2 : 1);  // the boundary a pair of ranks shares shrinks as the ranges do
        if (const char* env = std::getenv("CHAD_SHARD_BOX_MB")) { const long v = std::atol(env); if (v >= 1 && v <= 1024) box_mb = (size_t)v; }
        sh.box_words = (u32)(box_mb * (1u << 20) / 8);
        if (r == CHAD_OK) r = dev_ensure(ctx, sh.box_out, size_t(world) * sh.box_words * 8);
        if (r == CHAD_OK) r = dev_ensure(ctx, sh.box_in, size_t(world) * sh.box_words * 8);
        if (r != CHAD_OK) return bail(r);
        CREATE_TRY(cudaMemset(sh.scalars.p, 0, 256));
        CREATE_TRY(cudaMemset(sh.box_in.p, 0, size_t(world) * sh.box_words * 8));
        CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&sh.h_counts), SHARD_WORLD_MAX * sizeof(u32)));
        CREATE_TRY(cudaMallocHost(reinterpret_cast<void**>(&sh.h_roots), 4096 * 2 * sizeof(u32)));
        CREATE_TRY(cudaEventCreateWithFlags(&sh.counts_done, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&sh.table_free, cudaEventDisableTiming));
        if (const char* env = std::getenv("CHAD_SHARD_RANK0_SHARE")) { const int v = std::atoi(env); if (v >= 0 && v <= 256) sh.first_share_256 = (u32)v; }
        // two communicators: the per-batch exchange (group stream) and the per-submap gather (finalize stream) are queued from points of
        // the host code that are not ordered against each other, and NCCL wants one issue order per communicator
        const ncclUniqueId* ids = static_cast<const ncclUniqueId*>(shard_id);
        // NCCL sets up the connections an operation needs when the operation is first issued, and that handshake blocks the HOST until the
        // peer issues the matching call. Later the exchanges and gathers are queued from wherever a rank's poll finds their inputs ready
        // -- at different program points on different ranks -- so a blocking first use there can deadlock (rank A waits inside NCCL for
        // rank B, which waits for a batch exchange rank A has not issued yet). Hence: connect eagerly, and issue every kind of operation
        // once here, where all ranks are inside the same collective call.
        setenv("NCCL_RUNTIME_CONNECT", "0", 0);
        ncclResult_t n1 = nccl->CommInitRank(&sh.comm_x, world, ids[0], rank);
        ncclResult_t n2 = n1 == ncclSuccess ? nccl->CommInitRank(&sh.comm_f, world, ids[1], rank) : n1;
        if (n1 == ncclSuccess && n2 == ncclSuccess) n2 = nccl->CommInitRank(&sh.comm_c, world, ids[2], rank);
        if (n1 != ncclSuccess || n2 != ncclSuccess) {
            ctx->error = std::string("ncclCommInitRank: ") + nccl->GetErrorString(n1 != ncclSuccess ? n1 : n2);
            return bail(CHAD_ERR_CUDA);
        }
        CREATE_TRY(cudaEventCreateWithFlags(&sh.h2d_done, cudaEventDisableTiming));
        if (const char* env = std::getenv("CHAD_SHARD_SLICE_H2D")) sh.slice_h2d = std::atoi(env) != 0;
        {
            auto warm = [&]() -> ncclResult_t {
                ncclResult_t n;
                u64* out = sh.box_out.as<u64>();
                u64* in = sh.box_in.as<u64>();
                u32* d = sh.scalars.as<u32>();
                const size_t w = sh.box_words;
                cudaStream_t st = ctx->stream;
                if ((n = nccl->GroupStart()) != ncclSuccess) return n;  // the per-batch exchange: every pair, full boxes
                for (int g = 0; g < world; g++) {
                    if (g == rank) continue;
                    if ((n = nccl->Send(out + size_t(g) * w, w, ncclUint64, g, sh.comm_x, st)) != ncclSuccess) return n;
                    if ((n = nccl->Recv(in + size_t(g) * w, w, ncclUint64, g, sh.comm_x, st)) != ncclSuccess) return n;
                }
                if ((n = nccl->GroupEnd()) != ncclSuccess) return n;
                if ((n = nccl->AllGather(d + 8, d + 16, 1, ncclUint32, sh.comm_f, st)) != ncclSuccess) return n;      // chunk counts
                if ((n = nccl->AllGather(out, in, 1024, ncclFloat, sh.comm_c, st)) != ncclSuccess) return n;           // scan slices
                if ((n = nccl->Broadcast(d, d, 2, ncclUint32, 0, sh.comm_f, st)) != ncclSuccess) return n;           // roots
                if ((n = nccl->GroupStart()) != ncclSuccess) return n;                                               // chunk gather on rank 0
                for (int g = 1; g < world; g++) {
                    if (rank == 0) {
                        if ((n = nccl->Recv(in + size_t(g) * w, w, ncclUint64, g, sh.comm_f, st)) != ncclSuccess) return n;
                        if ((n = nccl->Recv(in + size_t(g) * w, 8, ncclUint64, g, sh.comm_f, st)) != ncclSuccess) return n;
                    } else if (g == rank) {
                        if ((n = nccl->Send(out, w, ncclUint64, 0, sh.comm_f, st)) != ncclSuccess) return n;
                        if ((n = nccl->Send(out, 8, ncclUint64, 0, sh.comm_f, st)) != ncclSuccess) return n;
                    }
                }
                return nccl->GroupEnd();
            };
            CREATE_TRY(cudaMemset(sh.box_out.p, 0, size_t(world) * sh.box_words * 8));
            const ncclResult_t n = warm();
            if (n != ncclSuccess) { ctx->error = std::string("NCCL warm-up: ") + nccl->GetErrorString(n); return bail(CHAD_ERR_CUDA); }
            CREATE_TRY(cudaStreamSynchronize(ctx->stream));
            CREATE_TRY(cudaMemset(sh.box_in.p, 0, size_t(world) * sh.box_words * 8));
            CREATE_TRY(cudaMemset(sh.scalars.p, 0, 256));
        }
    }
#undef CREATE_TRY
    *out = ctx;
    return CHAD_OK;
}

int chad_create(float sdf_res, float sdf_trunc, int device, int max_batch_scans, chad_ctx** out) {
    return create_impl(sdf_res, sdf_trunc, device, max_batch_scans, 0, 1, nullptr, out);
}

int chad_shard_unique_id(void* id) {
    if (!id) return fail(nullptr, CHAD_ERR_INVALID, "id is NULL");
    const char* why = "";
    const NcclApi* nccl = nccl_api(&why);
    if (!nccl) return fail(nullptr, CHAD_ERR_CUDA, std::string("chad_shard_unique_id: ") + why);
    ncclUniqueId* ids = static_cast<ncclUniqueId*>(id);
    for (int q = 0; q < 3; q++) {  // batch exchange, chunk gather, scan slices
        const ncclResult_t n = nccl->GetUniqueId(&ids[q]);
        if (n != ncclSuccess) return fail(nullptr, CHAD_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl->GetErrorString(n));
    }
    return CHAD_OK;
}

int chad_create_sharded(float sdf_res, float sdf_trunc, int device, int max_batch_scans, int rank, int world, const void* id, chad_ctx** out) {
    if (world == 1) return create_impl(sdf_res, sdf_trunc, device, max_batch_scans, 0, 1, nullptr, out);
    return create_impl(sdf_res, sdf_trunc, device, max_batch_scans, rank, world, id, out);
}

int chad_shard_info(chad_ctx* ctx, int* rank, int* world, uint64_t* sent_runs, uint64_t* sent_records, uint64_t* exchanges) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (rank) *rank = ctx->sh.rank;
    if (world) *world = ctx->sh.world;
    if (sent_runs) *sent_runs = ctx->sh.sent_runs;
    if (sent_records) *sent_records = ctx->sh.sent_records;
    if (exchanges) *exchanges = ctx->sh.exchanges;
    return CHAD_OK;
}

void chad_destroy(chad_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->sh.world > 1) {
        for (cudaStream_t st : {ctx->group_stream, ctx->fin_stream, ctx->copy_stream}) if (st) cudaStreamSynchronize(st);
        if (ctx->sh.comm_x) ctx->sh.nccl->CommDestroy(ctx->sh.comm_x);
        if (ctx->sh.comm_f) ctx->sh.nccl->CommDestroy(ctx->sh.comm_f);
        if (ctx->sh.comm_c) ctx->sh.nccl->CommDestroy(ctx->sh.comm_c);
        if (ctx->sh.h2d_done) cudaEventDestroy(ctx->sh.h2d_done);
        for (DevBuf* b : {&ctx->sh.splitters, &ctx->sh.filter_mem, &ctx->sh.batch_scans[0], &ctx->sh.batch_scans[1], &ctx->sh.batch_scans[2], &ctx->sh.box_out,
                          &ctx->sh.box_in, &ctx->sh.scalars})
            dev_free(*b);
        if (ctx->sh.h_counts) cudaFreeHost(ctx->sh.h_counts);
        if (ctx->sh.h_roots) cudaFreeHost(ctx->sh.h_roots);
        if (ctx->sh.counts_done) cudaEventDestroy(ctx->sh.counts_done);
        if (ctx->sh.table_free) cudaEventDestroy(ctx->sh.table_free);
    }
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->walk_stream) cudaStreamSynchronize(ctx->walk_stream);
    if (ctx->group_stream) cudaStreamSynchronize(ctx->group_stream);
    if (ctx->fold_stream) cudaStreamSynchronize(ctx->fold_stream);
    if (ctx->fin_stream) cudaStreamSynchronize(ctx->fin_stream);
    DevBuf* bufs[] = {&ctx->xyz_sorted2[0], &ctx->normals2[0], &ctx->d_scans2[0], &ctx->xyz_sorted2[1], &ctx->normals2[1], &ctx->d_scans2[1], &ctx->keys_c, &ctx->run_mem[2], &ctx->t2_keys, &ctx->t2_cells, &ctx->t2_count, &ctx->t2_list, &ctx->t_list, &ctx->f_counters, &ctx->f_partial, &ctx->d_xyz[0], &ctx->d_xyz[1], &ctx->d_scans, &ctx->d_plan, &ctx->bt_mem, &ctx->run_mem[0], &ctx->run_mem[1], &ctx->radix_ws2, &ctx->pk_a, &ctx->pk_b, &ctx->pv_a, &ctx->pv_b, &ctx->keys_a, &ctx->keys_b, &ctx->vals_a, &ctx->vals_b,
                      &ctx->sorted_keys, &ctx->sorted_order, &ctx->xyz_sorted, &ctx->normals, &ctx->seg_info, &ctx->counts, &ctx->offsets,
                      &ctx->radix_ws, &ctx->scan_ws, &ctx->t_keys, &ctx->t_cells, &ctx->t_count, &ctx->f_sorted, &ctx->f_ids[0], &ctx->f_ids[1], &ctx->f_slots[0],
                      &ctx->f_slots[1], &ctx->f_cells, &ctx->f_tsdf, &ctx->f_addr[0], &ctx->f_addr[1], &ctx->f_head_rank, &ctx->f_cand,
                      &ctx->f_slot_of, &ctx->f_is_new, &ctx->f_rank, &ctx->f_radix_ws, &ctx->f_scan_ws, &ctx->f_scalars};
    for (DevBuf* b : bufs) dev_free(*b);
    for (auto& L : ctx->levels) { dev_free(L.raw); dev_free(L.entries); dev_free(L.first); }
    for (void* q : ctx->graveyard) cudaFree(q);
    for (int b = 0; b < 2; b++) {
        if (ctx->h_stage[b]) cudaFreeHost(ctx->h_stage[b]);
        if (ctx->h_scans_pinned[b]) cudaFreeHost(ctx->h_scans_pinned[b]);
        if (ctx->stage_copied[b]) cudaEventDestroy(ctx->stage_copied[b]);
        if (ctx->copy_done[b]) cudaEventDestroy(ctx->copy_done[b]);
    }
    for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->h_plan) cudaFreeHost(ctx->h_plan);
    if (ctx->h_plan_fold) cudaFreeHost(ctx->h_plan_fold);
    if (ctx->h_table_count) cudaFreeHost(ctx->h_table_count);
    if (ctx->h_table_count2) cudaFreeHost(ctx->h_table_count2);
    if (ctx->h_fin) cudaFreeHost(ctx->h_fin);
    for (cudaEvent_t e : {ctx->submap_closed, ctx->fin_p1_done, ctx->fin_done, ctx->fin_t0, ctx->fin_t3}) if (e) cudaEventDestroy(e);
    if (ctx->fin_stream) cudaStreamDestroy(ctx->fin_stream);
    for (cudaEvent_t e : {ctx->fold_done[0], ctx->fold_done[1], ctx->fold_done[2], ctx->submap_closed2, ctx->emit_done}) if (e) cudaEventDestroy(e);
    if (ctx->fold_stream) cudaStreamDestroy(ctx->fold_stream);
    if (ctx->group_stream) cudaStreamDestroy(ctx->group_stream);
    if (ctx->walk_stream) cudaStreamDestroy(ctx->walk_stream);
    for (int b = 0; b < MAX_SLOTS; b++) if (ctx->points_done[b]) cudaEventDestroy(ctx->points_done[b]);
    for (int b = 0; b < MAX_SLOTS; b++) if (ctx->front_done[b]) cudaEventDestroy(ctx->front_done[b]);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int b = 0; b < 2; b++) {
        if (ctx->xyz_free[b]) cudaEventDestroy(ctx->xyz_free[b]);
        if (ctx->scans_uploaded[b]) cudaEventDestroy(ctx->scans_uploaded[b]);
    }
    delete ctx;
}

static int insert_host(chad_ctx* ctx, const float* xyz, size_t n, const float position[3], bool wait_for_copy) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (!position || (n && !xyz)) return fail(ctx, CHAD_ERR_INVALID, "NULL argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool pinned = n && is_pinned_host(xyz);
    if (!wait_for_copy && n && !pinned) return fail(ctx, CHAD_ERR_INVALID, "chad_insert_async needs page-locked (cudaHostAlloc / cudaHostRegister) memory");
    bool skip;
    TRY(begin_scan(ctx, n, position, &skip));
    if (skip) return CHAD_OK;
    TRY(guard_batch_buffer(ctx));
    const int b = ctx->cur;
    float* dst = ctx->d_xyz[b].as<float>() + size_t(ctx->batch_points) * 3;
    if (ctx->sh.world > 1 && ctx->sh.slice_h2d && n >= 4096) {
        // Sharded: every rank is handed the whole scan, but the host link is the scarce resource (N ranks x 3 MB per scan through one
        // host's memory) and NVLink is not: this rank copies slice `rank` of the scan (plus the few floats left over after the division
        // into `world` equal slices of whole float4s, which every rank copies); ONE grouped all-gather per batch, issued when the batch
        // is closed (process_front: the same point of the call sequence on every rank), assembles the scans on
        // every rank. (An all-gather per scan cost ~60 us of latency on the copy stream, 100 times per step: profiles/bench_r02_final_n2_sliced_h2d.json.)
        chad_ctx::Shard& sh = ctx->sh;
        const size_t total = n * 3;
        const size_t c = (total / (size_t)sh.world) & ~size_t(3);
        const size_t lo = c * (size_t)sh.rank, tail = c * (size_t)sh.world;
        const float* src = xyz;
        if (!pinned) {
            if (ctx->stage_busy[b] && ctx->batch_scans == 0) {
                CUDA_TRY(ctx, cudaEventSynchronize(ctx->stage_copied[b]));
                ctx->stage_busy[b] = false;
            }
            float* stage = ctx->h_stage[b] + size_t(ctx->batch_points) * 3;
            std::memcpy(stage + lo, xyz + lo, c * 4);
            if (total > tail) std::memcpy(stage + tail, xyz + tail, (total - tail) * 4);
            src = stage;
            ctx->stage_busy[b] = true;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(dst + lo, src + lo, c * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (total > tail) CUDA_TRY(ctx, cudaMemcpyAsync(dst + tail, src + tail, (total - tail) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        sh.slices.push_back({dst, c});
        if (pinned && wait_for_copy) {  // (the caller's buffer has been read; the gather follows at the batch's close)
            CUDA_TRY(ctx, cudaEventRecord(sh.h2d_done, ctx->copy_stream));
            CUDA_TRY(ctx, cudaEventSynchronize(sh.h2d_done));
        }
        ctx->stats.h2d_bytes += (c + total - tail) * 4;
        return end_scan(ctx, n, position);
    }
    if (pinned) {
        // page-locked caller memory: DMA straight from it. chad_insert waits for the copy, so the caller may reuse the buffer at once
        // (one 3 MB transfer at a time reaches 36-44 of the link's 55 GB/s on an idle GPU: profiles/h2d_probe.py); chad_insert_async
        // leaves it in flight. Beside the insert kernels either way moves 26-33 GB/s (profiles/timeline.py 24 host)
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, xyz, n * 12, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (wait_for_copy) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    } else {
        if (ctx->stage_busy[b] && ctx->batch_scans == 0) {  // first scan of a new batch: the staging buffer may still be draining
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->stage_copied[b]));
            ctx->stage_busy[b] = false;
        }
        float* stage = ctx->h_stage[b] + size_t(ctx->batch_points) * 3;
        std::memcpy(stage, xyz, n * 12);
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, stage, n * 12, cudaMemcpyHostToDevice, ctx->copy_stream));
        ctx->stage_busy[b] = true;
    }
    ctx->stats.h2d_bytes += n * 12;
    return end_scan(ctx, n, position);
}

int chad_insert(chad_ctx* ctx, const float* xyz, size_t n, const float position[3]) { return insert_host(ctx, xyz, n, position, true); }
int chad_insert_async(chad_ctx* ctx, const float* xyz, size_t n, const float position[3]) { return insert_host(ctx, xyz, n, position, false); }

// The host loop of a C++ caller (for (scan : trajectory) map.insert(scan.points, scan.pose);) behind one call, for callers whose own
// loop is slow (a ctypes call costs ~6 us; 100 scans per 9 ms step make that 6 % of the step)
int chad_insert_many(chad_ctx* ctx, const float* const* xyz, const size_t* n, const float* positions, size_t count, int wait_for_copy) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (count && (!xyz || !n || !positions)) return fail(ctx, CHAD_ERR_INVALID, "NULL argument");
    for (size_t i = 0; i < count; i++) TRY(insert_host(ctx, xyz[i], n[i], positions + 3 * i, wait_for_copy != 0));
    return CHAD_OK;
}

int chad_insert_device(chad_ctx* ctx, const float* xyz_device, size_t n, const float position[3]) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (!position || (n && !xyz_device)) return fail(ctx, CHAD_ERR_INVALID, "NULL argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    bool skip;
    TRY(begin_scan(ctx, n, position, &skip));
    if (skip) return CHAD_OK;
    TRY(guard_batch_buffer(ctx));
    float* dst = ctx->d_xyz[ctx->cur].as<float>() + size_t(ctx->batch_points) * 3;
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, xyz_device, n * 12, cudaMemcpyDeviceToDevice, ctx->copy_stream));
    return end_scan(ctx, n, position);
}

int chad_flush(chad_ctx* ctx) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (ctx->sticky_error != CHAD_OK) return ctx->sticky_error;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    return shard_sync_roots(ctx);
}

int chad_finalize_active(chad_ctx* ctx) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (ctx->sticky_error != CHAD_OK) return ctx->sticky_error;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->has_pose) { TRY(settle(ctx)); return shard_sync_roots(ctx); }
    TRY(finalize_submap(ctx, false));
    ctx->has_pose = false;
    if (ctx->sh.world == 1) return CHAD_OK;
    TRY(settle(ctx));  // sharded: collective anyway -- gather, DAG and roots are complete when the call returns
    return shard_sync_roots(ctx);
}

int chad_submap_count(chad_ctx* ctx, uint32_t* count) {
    if (!ctx || !count) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    *count = (uint32_t)ctx->roots.size();
    return CHAD_OK;
}

int chad_submap_roots(chad_ctx* ctx, uint32_t i, uint32_t* root_tsdf, uint32_t* root_weight) {
    if (!ctx || !root_tsdf || !root_weight) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    if (i >= ctx->roots.size()) return fail(ctx, CHAD_ERR_INVALID, "submap index out of range");
    *root_tsdf = ctx->roots[i][0];
    *root_weight = ctx->roots[i][1];
    return CHAD_OK;
}

// voxel export: sorted chunks are copied to the host and expanded there (parity / debugging path)
static int export_voxels_impl(chad_ctx* ctx, uint64_t* keys, uint32_t* sd_bits, uint32_t* weights, size_t capacity, size_t* count) {
    TRY(settle(ctx));
    const u32 C = (u32)ctx->table_count_known;
    TRY(ensure_finalize_capacity(ctx, C));
    if (C) TRY(queue_sorted_chunks(ctx, ctx->stream, ctx->table, C));
    std::vector<u64> ck(C);
    std::vector<uint2> cells(size_t(C) * 8);
    if (C) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ck.data(), ctx->f_ids[0].p, size_t(C) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(cells.data(), ctx->f_cells.p, size_t(C) * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    size_t n = 0;
    for (size_t c = 0; c < C; c++)
        for (int s = 0; s < 8; s++) {
            const uint2 cell = cells[c * 8 + s];
            if (cell.y == 0) continue;
            if (keys) {
                if (n >= capacity) return fail(ctx, CHAD_ERR_INVALID, "export capacity too small");
                keys[n] = (ck[c] << 3) | (u64)s;
                sd_bits[n] = cell.x;
                weights[n] = cell.y;
            }
            n++;
        }
    *count = n;
    return CHAD_OK;
}

int chad_voxel_count(chad_ctx* ctx, size_t* count) {
    if (!ctx || !count) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return export_voxels_impl(ctx, nullptr, nullptr, nullptr, 0, count);
}

int chad_export_voxels(chad_ctx* ctx, uint64_t* keys, uint32_t* sd_bits, uint32_t* weights, size_t capacity, size_t* count) {
    if (!ctx || !count || !keys || !sd_bits || !weights) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return export_voxels_impl(ctx, keys, sd_bits, weights, capacity, count);
}

int chad_level_words(chad_ctx* ctx, int level, size_t* words) {
    if (!ctx || !words || level < 0 || level >= CHAD_NUM_LEVELS) return CHAD_ERR_INVALID;
    if (ctx->sh.rank != 0) return fail(ctx, CHAD_ERR_INVALID, "sharded map: the DAG levels live on rank 0");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    const Level& L = ctx->levels[level];
    *words = (level == CHAD_LEVEL_CLUSTERS) ? size_t(L.uniques) + 1 : size_t(L.occupied);
    return CHAD_OK;
}

int chad_level_counters(chad_ctx* ctx, int level, uint32_t* uniques, uint32_t* dupes) {
    if (!ctx || !uniques || !dupes || level < 0 || level >= CHAD_NUM_LEVELS) return CHAD_ERR_INVALID;
    if (ctx->sh.rank != 0) return fail(ctx, CHAD_ERR_INVALID, "sharded map: the DAG levels live on rank 0");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    *uniques = ctx->levels[level].uniques;
    *dupes = ctx->levels[level].dupes;
    return CHAD_OK;
}

int chad_export_level(chad_ctx* ctx, int level, void* dst, size_t capacity_words) {
    if (!ctx || !dst || level < 0 || level >= CHAD_NUM_LEVELS) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t words = 0;
    TRY(chad_level_words(ctx, level, &words));
    if (capacity_words < words) return fail(ctx, CHAD_ERR_INVALID, "export capacity too small");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fin_stream));
    CUDA_TRY(ctx, cudaMemcpy(dst, ctx->levels[level].raw.p, words * (level == CHAD_LEVEL_CLUSTERS ? 8 : 4), cudaMemcpyDeviceToHost));
    return CHAD_OK;
}

int chad_query_voxels(chad_ctx* ctx, uint32_t submap, const uint64_t* keys, size_t n, uint8_t* bytes) {
    if (!ctx || (n && (!keys || !bytes))) return CHAD_ERR_INVALID;
    if (ctx->sh.rank != 0) return fail(ctx, CHAD_ERR_INVALID, "sharded map: the DAG levels live on rank 0");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    if (submap >= ctx->roots.size()) return fail(ctx, CHAD_ERR_INVALID, "submap index out of range");
    if (n >= (1ull << 31)) return fail(ctx, CHAD_ERR_INVALID, "too many queries");
    if (n == 0) return CHAD_OK;
    DevBuf dk, db;
    TRY(dev_ensure(ctx, dk, n * 8));
    int rc = dev_ensure(ctx, db, n);
    if (rc != CHAD_OK) { dev_free(dk); return rc; }
    DagReadArgs a{};
    for (int d = 0; d < 20; d++) a.raw[d] = ctx->levels[d].raw.as<u32>();
    a.clusters = ctx->levels[CHAD_LEVEL_CLUSTERS].raw.as<u64>();
    cudaError_t e = cudaMemcpyAsync(dk.p, keys, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) { ctx->stats.kernel_launches += launch_dag_query(ctx->stream, a, ctx->roots[submap][0], dk.as<u64>(), (u32)n, db.as<u8>()); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(bytes, db.p, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(dk);
    dev_free(db);
    if (e != cudaSuccess) return fail(ctx, CHAD_ERR_CUDA, std::string("chad_query_voxels: ") + cudaGetErrorString(e));
    return CHAD_OK;
}

int chad_reset(chad_ctx* ctx) {
    if (!ctx) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->sticky_error = CHAD_OK;
    ctx->batch_points = 0;
    ctx->batch_scans = 0;
    ctx->n_pend = 0;
    for (bool& f : ctx->fold_stats_pending) f = false;
    ctx->sh.need_splitters = true;
    ctx->sh.slices.clear();
    ctx->sh.gather_at.clear();
    ctx->sh.close_deferred = false;
    ctx->sh.closed_waiting = false;
    ctx->sh.table_free_valid = false;
    ctx->sh.front_seq = ctx->sh.closes_marked = ctx->sh.closes_gathered = 0;
    ctx->sh.xfer_pending = false;
    ctx->sh.roots_synced = 0;
    ctx->burst_batches = 0;
    ctx->sh.sent_runs = ctx->sh.sent_records = ctx->sh.exchanges = 0;
    if (ctx->sh.world > 1 && ctx->sh.filter_mem.p) CUDA_TRY(ctx, cudaMemsetAsync(ctx->sh.filter_mem.p, 0, ctx->sh.filter_mem.bytes, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->walk_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->group_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fold_stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->fin_stream));
    ctx->fold_in_flight = false;
    fold_bounds_reset(ctx);
    for (bool& f : ctx->fold_done_valid) f = false;
    ctx->fin_state = chad_ctx::FIN_IDLE;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_plan.p, 0, MAX_SLOTS * sizeof(BatchPlan), ctx->stream));
    launch_table_clear(ctx->stream, ctx->table);
    launch_table_clear(ctx->stream, ctx->table2);
    *ctx->h_table_count2 = 0;
    for (int d = 0; d < CHAD_NUM_LEVELS; d++) {
        Level& L = ctx->levels[d];
        L.uniques = 0; L.dupes = 0; L.occupied = (d == CHAD_LEVEL_CLUSTERS) ? 0 : 1;
        launch_dedup_clear(ctx->stream, L.table);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    TRY(level_counters_reset(ctx));
    *ctx->h_table_count = 0;
    ctx->table_count_known = 0;
    ctx->has_pose = false;
    ctx->roots.clear();
    ctx->positions.clear();
    ctx->active_positions.clear();
    ctx->stats.resident_clusters = 0;
    return CHAD_OK;
}

int chad_set_pair_path(chad_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    ctx->pair_path = mode;
    return CHAD_OK;
}

int chad_pipeline_info(chad_ctx* ctx, int* plan_slots, int* walk_overlapped) {
    if (!ctx || !plan_slots || !walk_overlapped) return CHAD_ERR_INVALID;
    *plan_slots = ctx->n_slots;
    *walk_overlapped = ctx->overlap_walk ? 1 : 0;
    return CHAD_OK;
}

int chad_profile_enable(chad_ctx* ctx, int on) {
    if (!ctx) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    ctx->profiling = on != 0;
    for (int c = 0; c < PC_COUNT; c++) { ctx->prof_ms[c] = 0.0; ctx->prof_launches[c] = 0; }
    return CHAD_OK;
}

int chad_profile_classes(void) { return PC_COUNT; }

int chad_profile_get(chad_ctx* ctx, int cls, const char** name, double* milliseconds, uint64_t* launches) {
    if (!ctx || cls < 0 || cls >= PC_COUNT) return CHAD_ERR_INVALID;
    static const char* base[] = {"plan_kernels", "point_keys_kernel"};
    static char names[PC_COUNT][48];
    const char* nm;
    if (cls < 2) nm = base[cls];
    else if (cls == PC_POINT_SORT_HIST) nm = "radix_histogram_kernel[points]";
    else if (cls >= PC_POINT_SORT_PASS0 && cls < PC_POINT_GATHER) { std::snprintf(names[cls], 48, "radix_onesweep_kernel[points,pass%d]", cls - PC_POINT_SORT_PASS0); nm = names[cls]; }
    else if (cls == PC_POINT_GATHER) nm = "point_gather_kernel";
    else if (cls == PC_NORMALS) nm = "segment_kernel+normals_kernel";
    else if (cls == PC_BAND_COUNT) nm = "band_count_kernel";
    else if (cls == PC_BAND_SCAN) nm = "scan_kernels[band offsets]";
    else if (cls == PC_BAND_EMIT) nm = "band_emit_kernel";
    else if (cls == PC_PAIR_SORT_HIST) nm = "radix_histogram_kernel[pairs]";
    else if (cls >= PC_PAIR_SORT_PASS0 && cls < PC_SEGMENT_COUNT) { std::snprintf(names[cls], 48, "radix_onesweep_kernel[pairs,pass%d]", cls - PC_PAIR_SORT_PASS0); nm = names[cls]; }
    else if (cls == PC_SEGMENT_COUNT) nm = "segment_count_kernel";
    else if (cls == PC_FOLD) nm = "fold_kernel";
    else if (cls == PC_BLOCKS_COUNT) nm = "blocks_count_kernel";
    else if (cls == PC_BLOCKS_SCAN) nm = "scan_kernels+blocks_compact_kernel";
    else if (cls == PC_BLOCKS_EMIT) nm = "blocks_emit_kernel";
    else if (cls == PC_BLOCKS_SORT) nm = "blocks_sort_kernel";
    else if (cls == PC_RUNS_EMIT) nm = "runs_emit_kernel";
    else if (cls == PC_RUNS_SORT) nm = "run descriptor sort + runs_group_kernel";
    else if (cls == PC_RUNS_FOLD) nm = "runs_fold_kernel";
    else if (cls == PC_SHARD_EXCHANGE) nm = "runs_pack_kernel + NCCL send/recv + runs_ingest_kernel";
    else nm = "finalize_submap[part 1 + part 2 on the finalize stream, overlapped with inserts]";
    if (name) *name = nm;
    if (milliseconds) *milliseconds = ctx->prof_ms[cls];
    if (launches) *launches = ctx->prof_launches[cls];
    return CHAD_OK;
}

int chad_profile_timeline(chad_ctx* ctx, int* classes, float* begin_ms, float* end_ms, size_t capacity, size_t* count) {
    if (!ctx || !count) return CHAD_ERR_INVALID;
    *count = ctx->timeline.size();
    if (!classes || !begin_ms || !end_ms) return CHAD_OK;
    for (size_t i = 0; i < ctx->timeline.size() && i < capacity; i++) {
        classes[i] = ctx->timeline[i].cls;
        begin_ms[i] = ctx->timeline[i].t0;
        end_ms[i] = ctx->timeline[i].t1;
    }
    return CHAD_OK;
}

int chad_memory_info(chad_ctx* ctx, chad_memory* out) {
    if (!ctx || !out) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    chad_memory m{};
    double worst = 0.0;
    for (int lv = 0; lv < CHAD_NUM_LEVELS; lv++) {
        const Level& L = ctx->levels[lv];
        const bool cluster = lv == CHAD_LEVEL_CLUSTERS;
        m.dag_words_bytes += cluster ? (u64(L.uniques) + 1) * 8 : u64(L.occupied) * 4;
        m.dag_arena_bytes += L.raw.bytes;
        m.dedup_bytes += L.entries.bytes + L.first.bytes;
        m.dedup_records += L.uniques;
        if (L.table.capacity) worst = std::max(worst, double(L.uniques) / double(L.table.capacity));
    }
    m.dedup_max_load_permille = (u64)(worst * 1000.0 + 0.5);
    m.chunk_table_bytes = (ctx->table.capacity + ctx->table2.capacity) * 76;
    m.batch_buffer_bytes = ctx->cap_pairs * 8 * 2 + ctx->cap_points * (12 * 2 + 8 * 3 + 4 * 6 + 12 * 2);
    m.grow_events = ctx->grow_events;
    m.grow_bytes = ctx->grow_bytes;
    m.grow_host_us = (u64)(ctx->grow_host_ms * 1000.0);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) { m.device_used_bytes = total_b - free_b; m.device_total_bytes = total_b; } else cudaGetLastError();
    *out = m;
    return CHAD_OK;
}

int chad_get_stats(chad_ctx* ctx, chad_stats* out) {
    if (!ctx || !out) return CHAD_ERR_INVALID;
    *out = ctx->stats;
    return CHAD_OK;
}

int chad_reset_stats(chad_ctx* ctx) {
    if (!ctx) return CHAD_ERR_INVALID;
    const uint64_t resident = ctx->stats.resident_clusters;
    ctx->stats = chad_stats{};
    ctx->stats.resident_clusters = resident;
    return CHAD_OK;
}

// ---- read path and persistence ---------------------------------------------
int chad_submap_positions(chad_ctx* ctx, uint32_t submap, float* xyz, size_t capacity, size_t* count) {
    if (!ctx || !count) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    if (submap > ctx->roots.size()) return fail(ctx, CHAD_ERR_INVALID, "submap index out of range");
    const std::vector<std::array<float, 3>> none;
    const auto& p = submap == ctx->roots.size() ? ctx->active_positions : (submap < ctx->positions.size() ? ctx->positions[submap] : none);
    *count = p.size();
    if (!xyz) return CHAD_OK;
    if (capacity < p.size()) return fail(ctx, CHAD_ERR_INVALID, "position capacity too small");
    if (!p.empty()) std::memcpy(xyz, p.data(), p.size() * 12);
    return CHAD_OK;
}

int chad_iterate_leaves(chad_ctx* ctx, uint32_t submap, uint64_t* keys, uint8_t* bytes, size_t capacity, size_t* count) {
    if (!ctx || !count || (keys && !bytes)) return CHAD_ERR_INVALID;
    if (ctx->sh.rank != 0) return fail(ctx, CHAD_ERR_INVALID, "sharded map: the DAG levels live on rank 0");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    if (submap >= ctx->roots.size()) return fail(ctx, CHAD_ERR_INVALID, "submap index out of range");
    cudaStream_t s = ctx->stream;
    const u32 root = ctx->roots[submap][0];
    int rc = CHAD_OK;
    // the widest frontier is the submap's number of leaf clusters, unknown before the walk: grow the work arrays until nothing overflows
    for (size_t cap = 1u << 20; rc == CHAD_OK; cap *= 2) {
        if (cap >= (1ull << 31)) { rc = fail(ctx, CHAD_ERR_CAPACITY, "chad_iterate_leaves: tree too wide"); break; }
        DevBuf addr[2], prefix[2], counts, offsets, ws, scal, dkeys, dbytes;
        auto release = [&]() { for (DevBuf* b : {&addr[0], &addr[1], &prefix[0], &prefix[1], &counts, &offsets, &ws, &scal, &dkeys, &dbytes}) dev_free(*b); };
        for (int q = 0; q < 2 && rc == CHAD_OK; q++) { rc = dev_ensure(ctx, addr[q], cap * 4); if (rc == CHAD_OK) rc = dev_ensure(ctx, prefix[q], cap * 8); }
        if (rc == CHAD_OK) rc = dev_ensure(ctx, counts, cap * 4);
        if (rc == CHAD_OK) rc = dev_ensure(ctx, offsets, cap * 4);
        if (rc == CHAD_OK) rc = dev_ensure(ctx, ws, scan_workspace_bytes(cap));
        if (rc == CHAD_OK) rc = dev_ensure(ctx, scal, 256);
        if (rc == CHAD_OK && keys) { rc = dev_ensure(ctx, dkeys, (capacity ? capacity : 1) * 8); if (rc == CHAD_OK) rc = dev_ensure(ctx, dbytes, capacity ? capacity : 1); }
        if (rc != CHAD_OK) { release(); break; }
        u32* d = scal.as<u32>();  // [0 .. 20] frontier sizes by depth, [21] voxels, [22] overflow
        const u32 init[2] = {1u, 0u};
        const u64 zero = 0;
        cudaMemsetAsync(d, 0, 256, s);
        cudaMemcpyAsync(d, &init[0], 4, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(addr[0].p, &root, 4, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(prefix[0].p, &zero, 8, cudaMemcpyHostToDevice, s);
        u64 launches = 0;
        int cur = 0;
        for (int depth = 0; depth < 20; depth++, cur ^= 1)
            launches += launch_iter_expand(s, ctx->levels[depth].raw.as<u32>(), addr[cur].as<u32>(), prefix[cur].as<u64>(), d + depth, (u32)cap, counts.as<u32>(),
                                           offsets.as<u32>(), ws.p, addr[cur ^ 1].as<u32>(), prefix[cur ^ 1].as<u64>(), d + depth + 1, d + 22);
        // a frontier wider than `cap` was cut: its size (d[depth]) still says so, and so does the overflow flag
        launches += launch_iter_leaves(s, ctx->levels[CHAD_LEVEL_CLUSTERS].raw.as<u64>(), addr[cur].as<u32>(), prefix[cur].as<u64>(), d + 20, (u32)cap,
                                       counts.as<u32>(), offsets.as<u32>(), ws.p, (u32)std::min<size_t>(capacity, 0xFFFFFFFFu), keys ? dkeys.as<u64>() : nullptr,
                                       keys ? dbytes.as<u8>() : nullptr, d + 21);
        ctx->stats.kernel_launches += launches;
        u32 h[24] = {0};
        cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { release(); rc = fail(ctx, CHAD_ERR_CUDA, std::string("chad_iterate_leaves: ") + cudaGetErrorString(e)); break; }
        bool cut = h[22] != 0;
        for (int depth = 0; depth <= 20; depth++) cut |= h[depth] > cap;
        if (cut) { release(); continue; }
        *count = h[21];
        if (keys) {
            if (capacity < h[21]) { release(); rc = fail(ctx, CHAD_ERR_INVALID, "leaf capacity too small"); break; }
            e = cudaMemcpy(keys, dkeys.p, size_t(h[21]) * 8, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(bytes, dbytes.p, h[21], cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = fail(ctx, CHAD_ERR_CUDA, std::string("chad_iterate_leaves: ") + cudaGetErrorString(e));
        }
        release();
        break;
    }
    return rc;
}

int chad_import_dag(chad_ctx* ctx, const chad_dag_image* img) {
    if (!ctx || !img) return CHAD_ERR_INVALID;
    if (ctx->sh.world > 1) return fail(ctx, CHAD_ERR_INVALID, "chad_import_dag: not on a sharded map");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    if (!ctx->roots.empty() || ctx->has_pose || ctx->levels[CHAD_LEVEL_CLUSTERS].uniques != 0)
        return fail(ctx, CHAD_ERR_INVALID, "chad_import_dag: the map must be empty (fresh, or after chad_reset)");
    if (img->n_submaps && (!img->roots)) return fail(ctx, CHAD_ERR_INVALID, "chad_import_dag: roots missing");
    cudaStream_t s = ctx->fin_stream;
    LevelCounters counters[CHAD_NUM_LEVELS];
    std::vector<u32> starts;
    for (int lv = 0; lv < CHAD_NUM_LEVELS; lv++) {
        const bool cluster = lv == CHAD_LEVEL_CLUSTERS;
        Level& L = ctx->levels[lv];
        const size_t words = cluster ? img->cluster_word_count : img->node_word_count[lv];
        const void* src = cluster ? static_cast<const void*>(img->cluster_words) : static_cast<const void*>(img->node_words[lv]);
        if (words == 0 || !src) return fail(ctx, CHAD_ERR_INVALID, "chad_import_dag: a level is missing (every level holds at least its reserved word 0)");
        if (words >= (1ull << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "chad_import_dag: level exceeds 2^31 words");
        u32 records = 0;
        starts.clear();
        if (cluster) {
            records = (u32)(words - 1);  // addresses 1 .. uniques
        } else {  // records follow each other from address 1: [mask, children ...] (levels.hpp:57-88)
            const u32* w = img->node_words[lv];
            size_t a = 1;
            while (a < words) {
                starts.push_back((u32)a);
                a += 1 + (size_t)__builtin_popcount(w[a] & 0xFFu);
            }
            if (a != words) return fail(ctx, CHAD_ERR_INVALID, "chad_import_dag: a node level does not end on a record boundary");
            records = (u32)starts.size();
        }
        if (records != img->uniques[lv]) return fail(ctx, CHAD_ERR_INVALID, "chad_import_dag: a level's unique count does not match its records");
        L.uniques = 0; L.dupes = 0; L.occupied = cluster ? 0 : 1;
        TRY(level_reserve(ctx, L, cluster, size_t(records) + 1024));
        if (size_t(words) > L.raw_cap) return fail(ctx, CHAD_ERR_CAPACITY, "chad_import_dag: level buffer too small");
        CUDA_TRY(ctx, cudaMemcpyAsync(L.raw.p, src, words * (cluster ? 8 : 4), cudaMemcpyHostToDevice, s));
        DevBuf dstarts;
        if (!cluster && records) {
            TRY(dev_ensure(ctx, dstarts, size_t(records) * 4));
            CUDA_TRY(ctx, cudaMemcpyAsync(dstarts.p, starts.data(), size_t(records) * 4, cudaMemcpyHostToDevice, s));
        }
        ctx->stats.kernel_launches += launch_dedup_restore(s, L.table, cluster, L.raw.p, dstarts.as<u32>(), records);
        CUDA_TRY(ctx, cudaStreamSynchronize(s));  // (`starts` and the staging buffer are reused by the next level)
        dev_free(dstarts);
        L.uniques = img->uniques[lv];
        L.dupes = img->dupes[lv];
        L.occupied = cluster ? 0 : (u32)words;
        counters[lv] = LevelCounters{cluster ? 0u : (u32)words, L.uniques, L.dupes, 0u};
    }
    CUDA_TRY(ctx, cudaMemcpy(ctx->f_counters.p, counters, sizeof(counters), cudaMemcpyHostToDevice));
    const float* pos = img->positions;
    for (u32 i = 0; i < img->n_submaps; i++) {
        ctx->roots.push_back({img->roots[2 * i], img->roots[2 * i + 1]});
        std::vector<std::array<float, 3>> p;
        const u32 np = (img->position_counts && pos) ? img->position_counts[i] : 0u;
        for (u32 q = 0; q < np; q++, pos += 3) p.push_back({pos[0], pos[1], pos[2]});
        ctx->positions.push_back(std::move(p));
    }
    ctx->stats.submaps = ctx->roots.size();
    return CHAD_OK;
}

// ---- stage entry points ---------------------------------------------------------------------
static int stage_prepare(chad_ctx* ctx, size_t n) {
    TRY(settle(ctx));
    if (n > ctx->cap_points || ctx->cap_points == 0) TRY(ensure_batch_capacity(ctx, n ? n : 1));
    return CHAD_OK;
}
static void stage_single_scan(chad_ctx* ctx, size_t n, const float position[3]) {
    ctx->h_scans.offset[0] = 0;
    ctx->h_scans.offset[1] = (u32)n;
    std::memcpy(ctx->h_scans.pose[0], position, 12);
    batch_scans_tiles(ctx->h_scans, 1);
    *ctx->h_scans_pinned[0] = ctx->h_scans;
    cudaMemcpyAsync(ctx->d_scans.p, ctx->h_scans_pinned[0], sizeof(BatchScans), cudaMemcpyHostToDevice, ctx->stream);
}
static int stage_check(chad_ctx* ctx) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    u32 flags = 0;
    CUDA_TRY(ctx, cudaMemcpy(&flags, plan_field<u32>(ctx, 0, offsetof(BatchPlan, error)), 4, cudaMemcpyDeviceToHost));
    if (flags) {
        CUDA_TRY(ctx, cudaMemset(plan_field<u32>(ctx, 0, offsetof(BatchPlan, error)), 0, 4));
        int code = error_from_flags(ctx, flags);
        ctx->sticky_error = CHAD_OK;  // stage calls do not poison the map
        return code;
    }
    return CHAD_OK;
}

int chad_stage_points(chad_ctx* ctx, const float* xyz, size_t n, const float position[3], float* xyz_sorted, uint64_t* keys, uint32_t* order,
                      float* normals) {
    if (!ctx || !position || (n && !xyz)) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(stage_prepare(ctx, n));
    if (n == 0) return CHAD_OK;
    cudaStream_t s = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_xyz[0].p, xyz, n * 12, cudaMemcpyHostToDevice, s));
    stage_single_scan(ctx, n, position);
    BatchPlan* plan = ctx->d_plan.as<BatchPlan>();
    const BatchScans* scans = ctx->d_scans.as<BatchScans>();
    const float* dxyz = ctx->d_xyz[0].as<float>();
    u64 launches = 0;
    launches += launch_plan(s, dxyz, (u32)n, 1, ctx->mp, plan, batch_tsb(ctx->h_scans, 1));
    launches += launch_point_keys(s, dxyz, (u32)n, scans, ctx->mp, plan, ctx->keys_a.as<u64>(), ctx->vals_a.as<u32>());
    launches += radix_sort_pairs(s, ctx->keys_a.as<u64>(), ctx->vals_a.as<u32>(), ctx->keys_b.as<u64>(), ctx->vals_b.as<u32>(),
                                 plan_field<u32>(ctx, 0, offsetof(BatchPlan, n_points)), plan_field<u32>(ctx, 0, offsetof(BatchPlan, nbits_points)), n,
                                 RS_MAX_PASSES, ctx->rws, ctx->num_sms, nullptr, 0, plan_field<u32>(ctx, 0, offsetof(BatchPlan, point_shift)));
    launches += launch_point_gather(s, dxyz, (u32)n, plan, ctx->keys_a.as<u64>(), ctx->keys_b.as<u64>(), ctx->vals_a.as<u32>(), ctx->vals_b.as<u32>(),
                                    ctx->sorted_keys.as<u64>(), ctx->sorted_order.as<u32>(), ctx->xyz_sorted.as<float>());
    launches += launch_normals(s, ctx->xyz_sorted.as<float>(), ctx->sorted_keys.as<u64>(), (u32)n, scans, plan, ctx->seg_info.as<u32>(),
                               ctx->normals.as<float>());
    launches += launch_point_full_keys(s, ctx->sorted_keys.as<u64>(), (u32)n, plan, ctx->keys_a.as<u64>());
    ctx->stats.kernel_launches += launches;
    TRY(stage_check(ctx));
    if (xyz_sorted) CUDA_TRY(ctx, cudaMemcpy(xyz_sorted, ctx->xyz_sorted.p, n * 12, cudaMemcpyDeviceToHost));
    if (keys) CUDA_TRY(ctx, cudaMemcpy(keys, ctx->keys_a.p, n * 8, cudaMemcpyDeviceToHost));
    if (order) CUDA_TRY(ctx, cudaMemcpy(order, ctx->sorted_order.p, n * 4, cudaMemcpyDeviceToHost));
    if (normals) CUDA_TRY(ctx, cudaMemcpy(normals, ctx->normals.p, n * 12, cudaMemcpyDeviceToHost));
    return CHAD_OK;
}

int chad_stage_pairs(chad_ctx* ctx, const float* xyz_sorted, const float* normals, size_t n, const float position[3], uint32_t* counts,
                     uint64_t* keys, float* sd, size_t capacity, size_t* total) {
    if (!ctx || !position || !total || (n && (!xyz_sorted || !normals))) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(stage_prepare(ctx, n));
    *total = 0;
    if (n == 0) return CHAD_OK;
    cudaStream_t s = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->xyz_sorted.p, xyz_sorted, n * 12, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->normals.p, normals, n * 12, cudaMemcpyHostToDevice, s));
    stage_single_scan(ctx, n, position);
    BatchPlan* plan = ctx->d_plan.as<BatchPlan>();
    const BatchScans* scans = ctx->d_scans.as<BatchScans>();
    u64 launches = 0;
    launches += launch_plan(s, ctx->xyz_sorted.as<float>(), (u32)n, 1, ctx->mp, plan, batch_tsb(ctx->h_scans, 1));
    launches += launch_band_count(s, ctx->xyz_sorted.as<float>(), (u32)n, scans, ctx->mp, plan, ctx->counts.as<u32>());
    launches += exclusive_scan<u32, u32>(s, ctx->counts.as<u32>(), ctx->offsets.as<u32>(), n, ctx->scan_ws.p, (u32*)nullptr,
                                         plan_field<u32>(ctx, 0, offsetof(BatchPlan, n_pairs)));
    launches += launch_band_emit(s, ctx->xyz_sorted.as<float>(), ctx->normals.as<float>(), (u32)n, scans, ctx->mp, plan, ctx->offsets.as<u32>(),
                                 ctx->keys_a.as<u64>(), ctx->vals_a.as<u32>(), (u32)ctx->cap_pairs, true);
    ctx->stats.kernel_launches += launches;
    TRY(stage_check(ctx));
    u32 U = 0;
    CUDA_TRY(ctx, cudaMemcpy(&U, plan_field<u32>(ctx, 0, offsetof(BatchPlan, n_pairs)), 4, cudaMemcpyDeviceToHost));
    *total = U;
    if (counts) CUDA_TRY(ctx, cudaMemcpy(counts, ctx->counts.p, n * 4, cudaMemcpyDeviceToHost));
    if (keys || sd) {
        if (capacity < U) return fail(ctx, CHAD_ERR_INVALID, "pair capacity too small");
        if (keys) CUDA_TRY(ctx, cudaMemcpy(keys, ctx->keys_a.p, size_t(U) * 8, cudaMemcpyDeviceToHost));
        if (sd) CUDA_TRY(ctx, cudaMemcpy(sd, ctx->vals_a.p, size_t(U) * 4, cudaMemcpyDeviceToHost));
    }
    return CHAD_OK;
}

int chad_stage_sort(chad_ctx* ctx, uint64_t* keys, uint32_t* values, size_t n, int nbits) {
    if (!ctx || (n && (!keys || !values)) || nbits < 0 || nbits > 64) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n >= (1ull << 30)) return fail(ctx, CHAD_ERR_INVALID, "n too large");
    TRY(stage_prepare(ctx, (n + ctx->mp.max_ray_voxels - 1) / ctx->mp.max_ray_voxels));
    if (n == 0) return CHAD_OK;
    cudaStream_t s = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->keys_a.p, keys, n * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->vals_a.p, values, n * 4, cudaMemcpyHostToDevice, s));
    const u32 scal[2] = {(u32)n, (u32)nbits};
    CUDA_TRY(ctx, cudaMemcpyAsync(scalar32(ctx, SC_COUNT), &scal[0], 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(scalar32(ctx, SC_NBITS), &scal[1], 4, cudaMemcpyHostToDevice, s));
    ctx->stats.kernel_launches += radix_sort_pairs(s, ctx->keys_a.as<u64>(), ctx->vals_a.as<u32>(), ctx->keys_b.as<u64>(), ctx->vals_b.as<u32>(),
                                                   scalar32(ctx, SC_COUNT), scalar32(ctx, SC_NBITS), n, RS_MAX_PASSES, ctx->rws, ctx->num_sms);
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    CUDA_TRY(ctx, cudaGetLastError());
    const bool alt = radix_result_in_alt((u32)nbits);
    CUDA_TRY(ctx, cudaMemcpy(keys, alt ? ctx->keys_b.p : ctx->keys_a.p, n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(ctx, cudaMemcpy(values, alt ? ctx->vals_b.p : ctx->vals_a.p, n * 4, cudaMemcpyDeviceToHost));
    return CHAD_OK;
}

int chad_stage_morton(chad_ctx* ctx, const int32_t* voxels, size_t n, uint64_t* keys) {
    if (!ctx || (n && (!voxels || !keys))) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(stage_prepare(ctx, n));
    if (n == 0) return CHAD_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->xyz_sorted.p, voxels, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.kernel_launches += launch_morton_encode(ctx->stream, ctx->xyz_sorted.as<i32>(), (u32)n, ctx->keys_a.as<u64>());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpy(keys, ctx->keys_a.p, n * 8, cudaMemcpyDeviceToHost));
    return CHAD_OK;
}

uint64_t chad_morton_encode(int32_t x, int32_t y, int32_t z) { return morton_encode(x, y, z); }
void chad_morton_decode(uint64_t key, int32_t* x, int32_t* y, int32_t* z) { morton_decode(key, *x, *y, *z); }
uint64_t chad_key_compact(uint64_t key, unsigned k) { return compact_key(key, k); }
uint64_t chad_key_expand(uint64_t compact, unsigned k) { return expand_key(compact, k); }

// ---- one submap integrated by another rank (submap-parallel mode of chad_tsdf_b200/sharded.py): chunk stream in, chunk stream out ----
int chad_shard_export_chunks(chad_ctx* ctx, size_t* n_chunks, void** keys_device, void** cells_device) {
    if (!ctx || !n_chunks || !keys_device || !cells_device) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(settle(ctx));
    const u32 C = (u32)ctx->table_count_known;
    TRY(ensure_finalize_capacity(ctx, C));
    if (C) TRY(queue_sorted_chunks(ctx, ctx->stream, ctx->table, C));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_chunks = C;
    *keys_device = ctx->f_ids[0].p;
    *cells_device = ctx->f_cells.p;
    return CHAD_OK;
}

int chad_shard_clear(chad_ctx* ctx) {
    if (!ctx) return CHAD_ERR_INVALID;
    if (ctx->sticky_error != CHAD_OK) return ctx->sticky_error;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(drain(ctx));
    launch_table_clear(ctx->stream, ctx->table);  // octree.clear(), tsdf.cpp:57
    *ctx->h_table_count = 0;
    ctx->table_count_known = 0;
    ctx->has_pose = false;
    ctx->stats.resident_clusters = 0;
    return CHAD_OK;
}

int chad_shard_finalize_from(chad_ctx* ctx, const uint64_t* keys_device, const void* cells_device, size_t n_chunks, int clear_local) {
    if (!ctx || (n_chunks && (!keys_device || !cells_device))) return CHAD_ERR_INVALID;
    if (ctx->sticky_error != CHAD_OK) return ctx->sticky_error;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n_chunks >= (1ull << 31)) return fail(ctx, CHAD_ERR_CAPACITY, "submap exceeds 2^31 leaf chunks");
    if (clear_local) TRY(drain(ctx));
    TRY(finalize_wait(ctx));  // the gathered stream goes into the finalize work buffers
    TRY(ensure_finalize_capacity(ctx, n_chunks));
    if (n_chunks) {
        if (keys_device != ctx->f_ids[0].p) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->f_ids[0].p, keys_device, n_chunks * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (cells_device != ctx->f_cells.p) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->f_cells.p, cells_device, n_chunks * 64, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (clear_local) {
        // the shard's chunks have been exported: clear it now (octree.clear(), tsdf.cpp:57) so that the next submap's folds can
        // start while the finalize runs on its own stream (it only reads the gathered chunk stream)
        launch_table_clear(ctx->stream, ctx->table);
        *ctx->h_table_count = 0;
        ctx->table_count_known = 0;
        ctx->has_pose = false;
        ctx->stats.resident_clusters = 0;
    }
    return finalize_begin(ctx, (u32)n_chunks, true, ctx->stream);
}

int chad_host_alloc(size_t bytes, void** host_ptr) {
    if (!host_ptr) return CHAD_ERR_INVALID;
    *host_ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(host_ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, CHAD_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    return CHAD_OK;
}
int chad_host_free(void* host_ptr) {
    if (!host_ptr) return CHAD_OK;
    return cudaFreeHost(host_ptr) == cudaSuccess ? CHAD_OK : CHAD_ERR_CUDA;
}
int chad_host_register(void* host_ptr, size_t bytes) {
    if (!host_ptr || !bytes) return CHAD_ERR_INVALID;
    const cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, CHAD_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
    return CHAD_OK;
}
int chad_host_unregister(void* host_ptr) {
    if (!host_ptr) return CHAD_ERR_INVALID;
    if (cudaHostUnregister(host_ptr) != cudaSuccess) { cudaGetLastError(); return CHAD_ERR_CUDA; }
    return CHAD_OK;
}

int chad_device_alloc(chad_ctx* ctx, size_t bytes, void** device_ptr) {
    if (!ctx || !device_ptr) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMalloc(device_ptr, bytes ? bytes : 1));
    return CHAD_OK;
}
int chad_device_free(chad_ctx* ctx, void* device_ptr) {
    if (!ctx) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaFree(device_ptr));
    return CHAD_OK;
}
int chad_upload(chad_ctx* ctx, void* device_dst, const void* host_src, size_t bytes) {
    if (!ctx || !device_dst || !host_src) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpy(device_dst, host_src, bytes, cudaMemcpyHostToDevice));
    return CHAD_OK;
}
int chad_timer_begin(chad_ctx* ctx) {
    if (!ctx) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    return CHAD_OK;
}
int chad_timer_end(chad_ctx* ctx, float* milliseconds) {
    if (!ctx || !milliseconds) return CHAD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaEventRecord(ctx->t1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->t1));
    CUDA_TRY(ctx, cudaEventElapsedTime(milliseconds, ctx->t0, ctx->t1));
    return CHAD_OK;
}

}  // extern "C"
