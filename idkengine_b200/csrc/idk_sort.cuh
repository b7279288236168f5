// Stable LSD radix sort of the alive list by the cached 21-bit hit key (3 passes x 7 bits).
// Replaces the reference's counting sort (PathTracer.RaySorting, PathTracer.cs:273-297 and
// Resource/Shaders/PathTracing/CountingSort/**): same result as a stable sort by `key & 0x1FFFFF`, without the
// fixed 2^21-bin histogram clear/scan per bounce and without the unordered atomics of Reorder/compute.glsl.
// The element count lives on the device (alive count of the bounce); nothing is read back.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IDK_SORT_BITS 7
#define IDK_SORT_BINS (1 << IDK_SORT_BITS)
#define IDK_SORT_THREADS 256
#define IDK_SORT_ROUNDS 8
#define IDK_SORT_TILE (IDK_SORT_THREADS * IDK_SORT_ROUNDS)

struct IdkSortScratch {
    uint32_t* keysA = nullptr;
    uint32_t* keysB = nullptr;
    uint32_t* valsA = nullptr;
    uint32_t* valsB = nullptr;
    uint32_t* hist = nullptr;      // [IDK_SORT_BINS][maxTiles]
    uint32_t capacity = 0;
    uint32_t maxTiles = 0;
};

static inline void idk_sort_release(IdkSortScratch& s) {
    cudaFree(s.keysA); cudaFree(s.keysB); cudaFree(s.valsA); cudaFree(s.valsB); cudaFree(s.hist);
    s = IdkSortScratch();
}

static inline int idk_sort_prepare(IdkSortScratch& s, uint32_t n) {
    if (n <= s.capacity && s.keysA) return 0;
    idk_sort_release(s);
    s.maxTiles = (n + IDK_SORT_TILE - 1) / IDK_SORT_TILE + 1;
    if (cudaMalloc(&s.keysA, (size_t)n * 4) != cudaSuccess) return -1;
    if (cudaMalloc(&s.keysB, (size_t)n * 4) != cudaSuccess) return -1;
    if (cudaMalloc(&s.valsA, (size_t)n * 4) != cudaSuccess) return -1;
    if (cudaMalloc(&s.valsB, (size_t)n * 4) != cudaSuccess) return -1;
    if (cudaMalloc(&s.hist, (size_t)IDK_SORT_BINS * s.maxTiles * 4) != cudaSuccess) return -1;
    s.capacity = n;
    return 0;
}

__global__ void __launch_bounds__(IDK_SORT_THREADS) k_sort_histogram(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ countPtr,
                                                                      uint32_t* __restrict__ hist, uint32_t maxTiles, int shift) {
    __shared__ uint32_t s_hist[IDK_SORT_BINS];
    const uint32_t count = *countPtr;
    const uint32_t numTiles = (count + IDK_SORT_TILE - 1) / IDK_SORT_TILE;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        if (threadIdx.x < IDK_SORT_BINS) s_hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t base = tile * IDK_SORT_TILE;
        for (int r = 0; r < IDK_SORT_ROUNDS; r++) {
            const uint32_t i = base + r * IDK_SORT_THREADS + threadIdx.x;
            if (i < count) atomicAdd(&s_hist[(keys[i] >> shift) & (IDK_SORT_BINS - 1)], 1u);
        }
        __syncthreads();
        if (threadIdx.x < IDK_SORT_BINS) hist[threadIdx.x * maxTiles + tile] = s_hist[threadIdx.x];
        __syncthreads();
    }
}

// Exclusive scan of hist in (bin-major, tile-minor) order over the valid tiles; one block.
__global__ void __launch_bounds__(1024) k_sort_scan(const uint32_t* __restrict__ countPtr, uint32_t* __restrict__ hist, uint32_t maxTiles) {
    __shared__ uint32_t s_part[1024];
    const uint32_t count = *countPtr;
    const uint32_t numTiles = (count + IDK_SORT_TILE - 1) / IDK_SORT_TILE;
    const uint32_t total = numTiles * IDK_SORT_BINS;
    const uint32_t per = (total + 1023) / 1024;
    const uint32_t b = threadIdx.x * per, e = min(total, b + per);
    uint32_t sum = 0;
    for (uint32_t k = b; k < e; k++) sum += hist[(k / numTiles) * maxTiles + (k % numTiles)];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= (uint32_t)off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;
    for (uint32_t k = b; k < e; k++) {
        const uint32_t idx = (k / numTiles) * maxTiles + (k % numTiles);
        const uint32_t v = hist[idx];
        hist[idx] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(IDK_SORT_THREADS) k_sort_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                    uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                                    const uint32_t* __restrict__ countPtr, const uint32_t* __restrict__ hist,
                                                                    uint32_t maxTiles, int shift) {
    __shared__ uint32_t s_running[IDK_SORT_BINS];
    __shared__ uint32_t s_warp[IDK_SORT_THREADS / 32][IDK_SORT_BINS];
    const uint32_t count = *countPtr;
    const uint32_t numTiles = (count + IDK_SORT_TILE - 1) / IDK_SORT_TILE;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < IDK_SORT_BINS) s_running[threadIdx.x] = hist[threadIdx.x * maxTiles + tile];
        const uint32_t base = tile * IDK_SORT_TILE;
        for (int r = 0; r < IDK_SORT_ROUNDS; r++) {
            for (uint32_t k = threadIdx.x; k < (IDK_SORT_THREADS / 32) * IDK_SORT_BINS; k += IDK_SORT_THREADS) (&s_warp[0][0])[k] = 0;
            __syncthreads();
            const uint32_t i = base + r * IDK_SORT_THREADS + threadIdx.x;
            const bool valid = i < count;
            uint32_t key = 0, val = 0, digit = 0xFFFFFFFFu;
            if (valid) {
                key = keysIn[i];
                val = valsIn ? valsIn[i] : i;
                digit = (key >> shift) & (IDK_SORT_BINS - 1);
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            const uint32_t rankInWarp = __popc(peers & ((1u << lane) - 1u));
            if (valid && rankInWarp == 0) s_warp[warp][digit] = __popc(peers);
            __syncthreads();
            if (valid) {
                uint32_t off = s_running[digit];
                for (uint32_t w = 0; w < warp; w++) off += s_warp[w][digit];
                const uint32_t dst = off + rankInWarp;
                keysOut[dst] = key;
                valsOut[dst] = val;
            }
            __syncthreads();
            if (threadIdx.x < IDK_SORT_BINS) {
                uint32_t add = 0;
                for (uint32_t w = 0; w < IDK_SORT_THREADS / 32; w++) add += s_warp[w][threadIdx.x];
                s_running[threadIdx.x] += add;
            }
            __syncthreads();
        }
    }
}

// Returns the number of kernels launched, or -1 on a launch error. perm[gid] = value (pixel) of the gid-th sorted ray.
// valsIn = the values that travel with the keys (the alive list); null = slot indices.
static inline int idk_sort_by_key(IdkSortScratch& s, const uint32_t* keys, const uint32_t* valsIn, uint32_t* perm, const uint32_t* countPtr,
                                  uint32_t capacity, int smCount, cudaStream_t stream) {
    (void)capacity;
    const int grid = smCount * 4;
    const uint32_t* kin[3] = {keys, s.keysB, s.keysA};
    const uint32_t* vin[3] = {valsIn, s.valsB, s.valsA};
    uint32_t* kout[3] = {s.keysB, s.keysA, s.keysB};
    uint32_t* vout[3] = {s.valsB, s.valsA, perm};
    for (int p = 0; p < 3; p++) {
        const int shift = p * IDK_SORT_BITS;
        k_sort_histogram<<<grid, IDK_SORT_THREADS, 0, stream>>>(kin[p], countPtr, s.hist, s.maxTiles, shift);
        k_sort_scan<<<1, 1024, 0, stream>>>(countPtr, s.hist, s.maxTiles);
        k_sort_scatter<<<grid, IDK_SORT_THREADS, 0, stream>>>(kin[p], vin[p], kout[p], vout[p], countPtr, s.hist, s.maxTiles, shift);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 9;
}
