// Device-wide exclusive prefix sum (reduce / scan-of-block-sums / downsweep). Used for the
// per-ray voxel offsets of the band enumeration (the reference appends to a std::vector instead,
// include/chad/detail/octree.hpp:85,122,151) and for first-occurrence address assignment in the
// DAG dedup (the reference increments _occupied_n / _uniques_n one add at a time,
// include/chad/detail/levels.hpp:76-82,130-133).
#pragma once
#include "common.cuh"

namespace chadgpu {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 elements per block

inline size_t scan_num_blocks(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }
inline size_t scan_workspace_bytes(size_t max_n) { return (scan_num_blocks(max_n) + 1) * sizeof(u64); }

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp_totals, T& total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (u32)off) incl += t;
    }
    if (lane == 31) s_warp_totals[warp] = incl;
    __syncthreads();
    T before = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; w++) {
        T x = s_warp_totals[w];
        if ((u32)w < warp) before += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return before + incl - v;
}

template <typename T, typename In>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const In* __restrict__ in, size_t n, T* __restrict__ block_sums) {
    __shared__ T s_warp[SCAN_THREADS / 32];
    const size_t base = size_t(blockIdx.x) * SCAN_TILE;
    T sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        size_t i = base + size_t(j) * SCAN_THREADS + threadIdx.x;
        if (i < n) sum += (T)in[i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++) t += s_warp[w];
        block_sums[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_sums in place, total -> *total_out (and optionally *total_out32)
template <typename T>
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(T* __restrict__ block_sums, size_t num_blocks, T* __restrict__ total_out,
                                                               u32* __restrict__ total_out32) {
    __shared__ T s_warp[32];
    T carry = 0;
    for (size_t base = 0; base < num_blocks; base += 1024) {
        size_t i = base + threadIdx.x;
        T v = (i < num_blocks) ? block_sums[i] : T(0);
        T tot;
        T ex = block_exclusive_scan<T>(v, s_warp, tot);
        if (i < num_blocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = carry;
        if (total_out32) *total_out32 = (u32)carry;
    }
}

template <typename T, typename In>
__global__ void __launch_bounds__(SCAN_THREADS) scan_downsweep_kernel(const In* in, size_t n, const T* __restrict__ block_sums, T* out) {
    __shared__ T s_warp[SCAN_THREADS / 32];
    // blocked arrangement: thread t owns elements [t*ITEMS, (t+1)*ITEMS) of the tile
    const size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        size_t i = base + j;
        v[j] = (i < n) ? (T)in[i] : T(0);
        sum += v[j];
    }
    T tot;
    T ex = block_exclusive_scan<T>(sum, s_warp, tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        size_t i = base + j;
        if (i < n) out[i] = ex;
        ex += v[j];
    }
}

// out[i] = sum_{j<i} in[j]; *d_total = sum of all. `workspace` >= scan_workspace_bytes(n). Returns launches.
template <typename T, typename In>
inline int exclusive_scan(cudaStream_t stream, const In* in, T* out, size_t n, void* workspace, T* d_total, u32* d_total32 = nullptr) {
    if (n == 0) {
        if (d_total) cudaMemsetAsync(d_total, 0, sizeof(T), stream);
        if (d_total32) cudaMemsetAsync(d_total32, 0, 4, stream);
        return 0;
    }
    T* block_sums = static_cast<T*>(workspace);
    size_t nb = scan_num_blocks(n);
    scan_reduce_kernel<T, In><<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, block_sums);
    scan_block_sums_kernel<T><<<1, 1024, 0, stream>>>(block_sums, nb, d_total, d_total32);
    scan_downsweep_kernel<T, In><<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, block_sums, out);
    return 3;
}

}  // namespace chadgpu
