// Onesweep-style LSD radix sort, see radix_sort.cuh for the design and the reference lines replaced.
#include "radix_sort.cuh"

namespace chadgpu {

namespace {

__device__ __forceinline__ u32 ld_volatile_u32(const u32* p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(u32* p, u32 v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// exclusive scan of one u32 per thread across a 256-thread block; `total` = block sum
__device__ __forceinline__ u32 block_exclusive_scan_256(u32 v, u32* s_warp_totals, u32& total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (u32)off) incl += t;
    }
    if (lane == 31) s_warp_totals[warp] = incl;
    __syncthreads();
    u32 before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        u32 x = s_warp_totals[w];
        if ((u32)w < warp) before += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return before + incl - v;
}

__global__ void __launch_bounds__(RS_THREADS) radix_histogram_kernel(const u64* __restrict__ keys, const u32* __restrict__ d_n,
                                                                     const u32* __restrict__ d_nbits, const u32* __restrict__ d_shift0,
                                                                     u32* __restrict__ hist, u32* __restrict__ lookback0) {
    __shared__ u32 s_hist[RS_MAX_PASSES][RS_RADIX];
    const u32 n = *d_n;
    const u32 shift0 = d_shift0 ? *d_shift0 : 0u;  // the sorted bits are [shift0, shift0 + nbits); lower bits ride along (packed payload)
    u32 npasses = radix_num_passes(*d_nbits);
    if (npasses > RS_MAX_PASSES) npasses = RS_MAX_PASSES;
    const u32 tid = threadIdx.x, lane = tid & 31;
#pragma unroll
    for (int p = 0; p < RS_MAX_PASSES; p++) s_hist[p][tid] = 0;
    __syncthreads();
    const u32 num_tiles = (n + RS_TILE - 1) / RS_TILE;
    for (u32 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        lookback0[size_t(tile) * RS_RADIX + tid] = 0;  // status words of pass 0
        const u32 base = tile * RS_TILE;
#pragma unroll 4
        for (int j = 0; j < RS_ITEMS; j++) {
            const u32 idx = base + j * RS_THREADS + tid;
            const bool valid = idx < n;
            const u64 key = valid ? (keys[idx] >> shift0) : 0ull;
            // digit 0 is close to uniform: plain shared atomics. Digits >= 1 are Morton-coherent (whole warps share
            // them): ONE match.any on key >> 8 groups the lanes whose upper digits all agree, one lane per group adds.
            if (valid) atomicAdd(&s_hist[0][(u32)(key & (RS_RADIX - 1))], 1u);
            const u64 upper = valid ? (key >> RS_RADIX_BITS) : ~0ull;
            const u32 m = __match_any_sync(0xffffffffu, upper);
            if (valid && lane == (u32)(__ffs(m) - 1)) {
                const u32 cnt = (u32)__popc(m);
                for (u32 p = 1; p < npasses; p++) atomicAdd(&s_hist[p][(u32)((key >> (p * RS_RADIX_BITS)) & (RS_RADIX - 1))], cnt);
            }
        }
    }
    __syncthreads();
    for (u32 p = 0; p < npasses; p++) {
        const u32 c = s_hist[p][tid];
        if (c) atomicAdd(&hist[p * RS_RADIX + tid], c);
    }
}

__global__ void __launch_bounds__(RS_THREADS, RS_CTAS_PER_SM)
    radix_onesweep_kernel(const u64* __restrict__ keys_in, const u32* __restrict__ vals_in, u64* __restrict__ keys_out,
                          u32* __restrict__ vals_out, const u32* __restrict__ d_n, const u32* __restrict__ d_nbits,
                          const u32* __restrict__ d_shift0, u32 pass, const u32* __restrict__ hist, u32* __restrict__ tile_counter,
                          u32* lookback_cur, u32* __restrict__ lookback_next) {
    const u32 npasses = radix_num_passes(*d_nbits);
    if (pass >= npasses) return;
    const u32 shift0 = d_shift0 ? *d_shift0 : 0u;
    const bool keys_only = shift0 != 0;  // the payload is packed into the key bits below shift0: no value array to move
    const u32 n = *d_n;
    const u32 num_tiles = (n + RS_TILE - 1) / RS_TILE;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 shift = shift0 + pass * RS_RADIX_BITS;
    const u32 lanemask_lt = (1u << lane) - 1u;

    extern __shared__ __align__(16) unsigned char s_dyn[];
    u64* s_keys = reinterpret_cast<u64*>(s_dyn);
    u32* s_vals = reinterpret_cast<u32*>(s_dyn + size_t(RS_TILE) * 8);
    __shared__ u32 s_warp_hist[RS_WARPS][RS_RADIX];
    __shared__ u32 s_tile_start[RS_RADIX];
    __shared__ u32 s_gbase[RS_RADIX];
    __shared__ u32 s_warp_totals[RS_WARPS];
    __shared__ u32 s_tile;

    // global start of digit `tid` = exclusive scan of the global histogram of this pass
    u32 hist_total;
    const u32 digit_base = block_exclusive_scan_256(hist[pass * RS_RADIX + tid], s_warp_totals, hist_total);

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&tile_counter[pass], 1u);  // tickets in order => look-back cannot deadlock
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= num_tiles) break;
        lookback_next[size_t(tile) * RS_RADIX + tid] = 0;  // status words of the next pass
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) s_warp_hist[w][tid] = 0;
        __syncthreads();

        // ---- load (warp-striped: element index = tile*TILE + warp*512 + j*32 + lane) ----
        const u32 base = tile * RS_TILE + warp * (32 * RS_ITEMS) + lane;
        u64 key[RS_ITEMS];
        u32 val[RS_ITEMS];
        u32 rank[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const u32 idx = base + j * 32;
            const bool valid = idx < n;
            key[j] = valid ? keys_in[idx] : ~0ull;
            val[j] = (valid && !keys_only) ? vals_in[idx] : 0u;
        }
        // ---- stable rank inside (warp, digit): match.any multi-split ----
        // all matches first (independent -> pipelined; ncu r01: the match latency was 25 % of the stall samples when each
        // match sat inside the serial shared-memory chain), then the serial running count per (warp, digit), items in order
        u32 mask[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const bool valid = (base + j * 32) < n;
            const u32 d = valid ? (u32)((key[j] >> shift) & (RS_RADIX - 1)) : 0xFFFFFFFFu;
            mask[j] = __match_any_sync(0xffffffffu, d);
        }
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const bool valid = (base + j * 32) < n;
            const u32 d = (u32)((key[j] >> shift) & (RS_RADIX - 1));
            const u32 m = mask[j];
            const u32 leader = (u32)(__ffs(m) - 1);
            u32 pre = 0;
            if (valid && lane == leader) {
                pre = s_warp_hist[warp][d];
                s_warp_hist[warp][d] = pre + (u32)__popc(m);
            }
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rank[j] = pre + (u32)__popc(m & lanemask_lt);
            __syncwarp();
        }
        __syncthreads();

        // ---- per digit: exclusive offsets across warps, tile count; publish the aggregate EARLY ----
        u32 tile_count = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const u32 c = s_warp_hist[w][tid];
            s_warp_hist[w][tid] = tile_count;
            tile_count += c;
        }
        u32* status = lookback_cur + size_t(tile) * RS_RADIX + tid;
        st_volatile_u32(status, (tile == 0 ? RS_FLAG_PREFIX : RS_FLAG_AGG) | tile_count);
        u32 tile_total;
        const u32 tile_start = block_exclusive_scan_256(tile_count, s_warp_totals, tile_total);
        s_tile_start[tid] = tile_start;
        __syncthreads();

        // ---- stage the tile in shared memory in sorted order (needs only tile-local offsets) ----
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            if ((base + j * 32) < n) {
                const u32 d = (u32)((key[j] >> shift) & (RS_RADIX - 1));
                const u32 pos = s_tile_start[d] + s_warp_hist[warp][d] + rank[j];
                s_keys[pos] = key[j];
                if (!keys_only) s_vals[pos] = val[j];
            }
        }
        // ---- decoupled look-back, AFTER the staging so that the predecessors had time to publish their prefixes:
        // exclusive count of digit `tid` over all earlier tiles; predecessors are read four at a time ----
        u32 excl = 0;
        if (tile > 0) {
            i32 t = (i32)tile - 1;
            bool done = false;
            while (!done) {
                u32 s4[4];
#pragma unroll
                for (int q = 0; q < 4; q++) s4[q] = (t - q >= 0) ? ld_volatile_u32(lookback_cur + size_t(t - q) * RS_RADIX + tid) : RS_FLAG_PREFIX;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (done) break;
                    if ((s4[q] & RS_FLAG_MASK) == 0) break;  // not published yet: re-read from here
                    excl += s4[q] & RS_VALUE_MASK;
                    t--;
                    if (s4[q] & RS_FLAG_PREFIX) done = true;
                }
            }
            st_volatile_u32(status, RS_FLAG_PREFIX | (excl + tile_count));
        }
        s_gbase[tid] = digit_base + excl - tile_start;  // u32 wrap-around is intended
        __syncthreads();
        // ---- coalesced write-out of every digit run ----
        const u32 tile_n = min((u32)RS_TILE, n - tile * RS_TILE);
        for (u32 e = tid; e < tile_n; e += RS_THREADS) {
            const u64 k = s_keys[e];
            const u32 d = (u32)((k >> shift) & (RS_RADIX - 1));
            const u32 dst = s_gbase[d] + e;
            keys_out[dst] = k;
            if (!keys_only) vals_out[dst] = s_vals[e];
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t radix_sort_init() {
    return cudaFuncSetAttribute(radix_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM_BYTES);
}

int radix_sort_pairs(cudaStream_t stream, u64* keys, u32* vals, u64* keys_alt, u32* vals_alt, const u32* d_n, const u32* d_nbits,
                     size_t max_n, int max_passes, const RadixWorkspace& ws, int num_sms, const LaunchHook* hook, int cls_base,
                     const u32* d_shift0) {
    if (max_n == 0) return 0;
    if (max_passes > RS_MAX_PASSES) max_passes = RS_MAX_PASSES;
    size_t max_tiles = (max_n + RS_TILE - 1) / RS_TILE;
    int launches = 0;
    cudaMemsetAsync(ws.hist, 0, (size_t(RS_MAX_PASSES) * 256 + 64) * 4, stream);
    int hist_grid = (int)(max_tiles < size_t(num_sms) * 4 ? max_tiles : size_t(num_sms) * 4);
    if (hook) hook->begin(hook->user, cls_base);
    radix_histogram_kernel<<<hist_grid, RS_THREADS, 0, stream>>>(keys, d_n, d_nbits, d_shift0, ws.hist, ws.lookback[0]);
    if (hook) hook->end(hook->user);
    launches++;
    int grid = (int)(max_tiles < size_t(num_sms) * RS_CTAS_PER_SM ? max_tiles : size_t(num_sms) * RS_CTAS_PER_SM);
    u64* kin = keys; u32* vin = vals; u64* kout = keys_alt; u32* vout = vals_alt;
    for (int p = 0; p < max_passes; p++) {
        if (hook) hook->begin(hook->user, cls_base + 1 + p);
        radix_onesweep_kernel<<<grid, RS_THREADS, RS_SMEM_BYTES, stream>>>(kin, vin, kout, vout, d_n, d_nbits, d_shift0, (u32)p, ws.hist,
                                                                            ws.tile_counter, ws.lookback[p & 1], ws.lookback[(p + 1) & 1]);
        if (hook) hook->end(hook->user);
        launches++;
        u64* tk = kin; kin = kout; kout = tk;
        u32* tv = vin; vin = vout; vout = tv;
    }
    return launches;
}

}  // namespace chadgpu
