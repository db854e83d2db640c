// binning.cu -- per-tile binning and depth sort.
//
// Replaces, with identical results, the reference's
//   cub::DeviceScan::InclusiveSum + D2H copy   (DGR/cuda_rasterizer/rasterizer_impl.cu:277-281)
//   duplicateWithKeys                          (rasterizer_impl.cu:70-111)
//   cub::DeviceRadixSort::SortPairs on 32+bit  (rasterizer_impl.cu:300-308)
//   identifyTileRanges                         (rasterizer_impl.cu:116-138)
//
// The reference sorts R 64-bit (tile | depth_bits) keys globally with a stable LSD radix sort, so
// within a tile instances are ordered by depth bits and, for equal depths, by emission order =
// ascending Gaussian index.  Here the per-tile instance counts are already known from preprocess,
// so the tile ranges come from one scan over T tiles, instances are scattered straight into their
// tile's segment, and each segment is sorted in shared memory by the 64-bit key
// (depth_bits << 32 | gaussian_index) -- a total order that equals the reference's stable order.
// Global traffic: 8 B written + 8 B read + 4 B written per instance instead of ~150 B/instance for
// a 6-pass global radix sort.
#include "common.cuh"

namespace fb200 {

namespace {

typedef unsigned long long u64;

// ---- tile scan: counts -> ranges, cursors, sort work lists, status ---------------------------------
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int T, const uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ cursor, uint32_t* __restrict__ list_small,
                 uint32_t* __restrict__ list_large, uint32_t* __restrict__ list_huge,
                 uint32_t* __restrict__ counters, long long capacity, int32_t* __restrict__ status) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t s_n[3];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x;
    if (tid < 3) s_n[tid] = 0;
    if (tid == 0) s_max = 0;
    const int per = (T + 1023) / 1024;
    const int begin = min(T, tid * per), end = min(T, begin + per);
    uint32_t local = 0, lmax = 0;
    for (int t = begin; t < end; ++t) {
        uint32_t c = tile_count[t];
        local += c;
        lmax = max(lmax, c);
    }
    // block exclusive scan of `local`
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += n;
    }
    if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        uint32_t w = warp_sums[tid];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= o) wi += n;
        }
        warp_sums[tid] = wi - w;   // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sums[tid >> 5] + incl - local;
    for (int t = begin; t < end; ++t) {
        uint32_t c = tile_count[t];
        // untouched tiles keep (0,0) exactly like the reference's memset + identifyTileRanges
        ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
        cursor[t] = run;
        if (c > 1) {
            if (c <= (uint32_t)kSortSmallMax) list_small[atomicAdd(&s_n[0], 1u)] = t;
            else if (c <= (uint32_t)kSortLargeMax) list_large[atomicAdd(&s_n[1], 1u)] = t;
            else list_huge[atomicAdd(&s_n[2], 1u)] = t;
        }
        run += c;
    }
    atomicMax(&s_max, lmax);
    __syncthreads();
    if (tid == 1023) {
        status[FB200_ST_NUM_RENDERED] = (int32_t)run;
        status[FB200_ST_OVERFLOW] = ((long long)run > capacity) ? 1 : 0;
        status[FB200_ST_MAX_TILE] = (int32_t)s_max;
    }
    if (tid < 3) counters[tid] = s_n[tid];
}

// ---- scatter: one (depth_bits<<32 | idx) key per (Gaussian, tile) into the tile's segment ----------
__global__ void __launch_bounds__(256)
scatter_kernel(int P, int tiles_x, const uint2* __restrict__ rect, const float* __restrict__ depth,
               uint32_t* __restrict__ cursor, u64* __restrict__ keys, const int32_t* __restrict__ status) {
    if (status[FB200_ST_OVERFLOW]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const uint2 r = rect[idx];
    const uint32_t minx = r.x & 0xffffu, miny = r.x >> 16, maxx = r.y & 0xffffu, maxy = r.y >> 16;
    if (maxx <= minx || maxy <= miny) return;
    const u64 key = ((u64)__float_as_uint(depth[idx]) << 32) | (uint32_t)idx;
    for (uint32_t ty = miny; ty < maxy; ++ty)
        for (uint32_t tx = minx; tx < maxx; ++tx) {
            const uint32_t slot = atomicAdd(cursor + ty * tiles_x + tx, 1u);
            keys[slot] = key;
        }
}

// ---- per-tile sort ----------------------------------------------------------------------------------
// Bitonic network in the "flip" formulation: every compare-exchange puts the smaller key at the
// lower index, so a virtual +inf padding beyond n is a no-op and arbitrary n needs no padding.
template <int kThreads>
__device__ __forceinline__ void bitonic_sort_shared(u64* __restrict__ s, int n) {
    int m = 2;
    while (m < n) m <<= 1;   // virtual power-of-two size
    for (int k = 2; k <= m; k <<= 1) {
        // flip step: partner = i ^ (k-1)
        for (int t = threadIdx.x; t < (m >> 1); t += kThreads) {
            const int blk = t / (k >> 1), off = t % (k >> 1);
            const int i = blk * k + off;
            const int l = blk * k + (k - 1 - off);
            if (l < n) {
                const u64 a = s[i], b = s[l];
                if (a > b) { s[i] = b; s[l] = a; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (m >> 1); t += kThreads) {
                const int i = ((t / j) * (j << 1)) + (t % j);
                const int l = i + j;
                if (l < n) {
                    const u64 a = s[i], b = s[l];
                    if (a > b) { s[i] = b; s[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}

template <int kThreads, int kMaxN, bool kDynamic>
__global__ void __launch_bounds__(kThreads)
tile_sort_shared_kernel(const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list,
                        const uint2* __restrict__ ranges, u64* __restrict__ keys,
                        uint32_t* __restrict__ point_list, const int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ u64 stat_smem[kDynamic ? 1 : kMaxN];
    u64* s = kDynamic ? reinterpret_cast<u64*>(dyn_smem) : stat_smem;
    if (status[FB200_ST_OVERFLOW]) return;
    const uint32_t count = *n_list;
    for (uint32_t w = blockIdx.x; w < count; w += gridDim.x) {
        const uint2 rg = ranges[list[w]];
        const int n = (int)(rg.y - rg.x);
        u64* g = keys + rg.x;
        for (int i = threadIdx.x; i < n; i += kThreads) s[i] = g[i];
        __syncthreads();
        bitonic_sort_shared<kThreads>(s, n);
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const u64 k = s[i];
            g[i] = k;
            point_list[rg.x + i] = (uint32_t)k;
        }
        __syncthreads();
    }
}

// Lists longer than the shared-memory classes: same network directly on global memory (L2).
__global__ void __launch_bounds__(1024)
tile_sort_global_kernel(const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list,
                        const uint2* __restrict__ ranges, u64* __restrict__ keys,
                        uint32_t* __restrict__ point_list, const int32_t* __restrict__ status) {
    if (status[FB200_ST_OVERFLOW]) return;
    const uint32_t count = *n_list;
    for (uint32_t w = blockIdx.x; w < count; w += gridDim.x) {
        const uint2 rg = ranges[list[w]];
        const int n = (int)(rg.y - rg.x);
        u64* g = keys + rg.x;
        bitonic_sort_shared<1024>(g, n);   // same code, global pointer; __syncthreads orders the stages
        for (int i = threadIdx.x; i < n; i += 1024) point_list[rg.x + i] = (uint32_t)g[i];
        __syncthreads();
    }
}

// tiles with exactly one instance need no sort, only the point_list entry
__global__ void __launch_bounds__(256)
single_instance_kernel(int T, const uint2* __restrict__ ranges, const u64* __restrict__ keys,
                       uint32_t* __restrict__ point_list, const int32_t* __restrict__ status) {
    if (status[FB200_ST_OVERFLOW]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const uint2 rg = ranges[t];
    if (rg.y - rg.x == 1u) point_list[rg.x] = (uint32_t)keys[rg.x];
}

}  // namespace

cudaError_t launch_binning(const FwdArgs& a, cudaStream_t s) {
    const int T = a.tiles_x * a.tiles_y;
    tile_scan_kernel<<<1, 1024, 0, s>>>(T, a.tile_count, a.ranges, a.cursor, a.list_small, a.list_large,
                                        a.list_huge, a.counters, a.capacity, a.status);
    if (a.prm.P > 0)
        scatter_kernel<<<(a.prm.P + 255) / 256, 256, 0, s>>>(a.prm.P, a.tiles_x, a.rect, a.depth, a.cursor,
                                                             a.keys, a.status);
    single_instance_kernel<<<(T + 255) / 256, 256, 0, s>>>(T, a.ranges, a.keys, a.point_list, a.status);
    // grid sizes: the work lists live on the device, so launch enough CTAs for the worst case and
    // let each CTA stride over the list.
    {
        const int grid = min(T, 148 * 8);
        tile_sort_shared_kernel<256, kSortSmallMax, false><<<grid, 256, 0, s>>>(
            a.list_small, a.counters + 0, a.ranges, a.keys, a.point_list, a.status);
    }
    {
        const int smem = kSortLargeMax * 8;
        cudaError_t e = cudaFuncSetAttribute(tile_sort_shared_kernel<1024, kSortLargeMax, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        const int grid = min(T, 148);
        tile_sort_shared_kernel<1024, kSortLargeMax, true><<<grid, 1024, smem, s>>>(
            a.list_large, a.counters + 1, a.ranges, a.keys, a.point_list, a.status);
    }
    {
        const int grid = min(T, 148);
        tile_sort_global_kernel<<<grid, 1024, 0, s>>>(a.list_huge, a.counters + 2, a.ranges, a.keys,
                                                      a.point_list, a.status);
    }
    return cudaGetLastError();
}

}  // namespace fb200
