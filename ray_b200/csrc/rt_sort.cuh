// rt_sort.cuh -- inter-bounce ray reordering (the role of Ref::SortRays_CPU, reference internal/CoreRef.cpp:1667-1710,
// and of the GPU reference's 8-pass LSD radix sort, internal/RendererGPU.h:756-780).
//
// The step is results-neutral: every pixel owns at most one ray per bounce and each ray carries its pixel in `xy`, so
// the ORDER of the ray list never reaches the image.  What matters is that rays that
// start close together and point the same way sit next to each other so a warp walks the same BVH nodes.  So instead of
// reproducing the reference's 32-bit hash + full radix sort (~240 B/ray), this is ONE counting-sort pass over a
// 15-bit key = direction octant (3 bits, major) | 12-bit Morton code of the origin in a 16^3 grid over the scene bounds:
//   k_sort_hist    read o,d (32 B) -> key (4 B) + global histogram (32768 bins, L2-resident atomics)
//   k_sort_scan    one block: exclusive scan of the histogram
//   k_sort_scatter read ray (72 B + key) -> slot = atomicAdd(bin) -> write ray (72 B)
// ~= 184 B/ray of HBM traffic, no multi-pass key shuffling.  Order inside a bin is arbitrary (and irrelevant).
#pragma once

#include "rt_kernels.cuh"

namespace rt {

constexpr int kSortKeyBits = 15;
constexpr int kSortBins = 1 << kSortKeyBits;

struct SortBufs {
    uint32_t *keys = nullptr;        // key per input ray
    uint32_t *keys_sorted = nullptr; // key per output ray (diagnostics / stage API)
    uint32_t *hist = nullptr;        // kSortBins
    float root_min[3] = {0, 0, 0};
    float inv_cell[3] = {1, 1, 1};
};

inline int alloc_sort_bufs(SortBufs &s, size_t n) {
    cudaFree(s.keys);
    cudaFree(s.keys_sorted);
    s.keys = s.keys_sorted = nullptr;
    if (!s.hist && cudaMalloc(&s.hist, kSortBins * sizeof(uint32_t)) != cudaSuccess) {
        return 1;
    }
    if (n == 0) {
        return 0;
    }
    if (cudaMalloc(&s.keys, n * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&s.keys_sorted, n * sizeof(uint32_t)) != cudaSuccess) {
        return 1;
    }
    return 0;
}

inline void free_sort_bufs(SortBufs &s) {
    cudaFree(s.keys);
    cudaFree(s.keys_sorted);
    cudaFree(s.hist);
    s = SortBufs{};
}

inline void set_sort_bounds(SortBufs &s, const float bmin[3], const float bmax[3]) {
    for (int i = 0; i < 3; ++i) {
        s.root_min[i] = bmin[i];
        const float ext = bmax[i] - bmin[i];
        s.inv_cell[i] = (ext > 0.0f) ? 16.0f / ext : 0.0f;
    }
}

struct SortGrid {
    float min_x, min_y, min_z, inv_x, inv_y, inv_z;
};

RT_DEV uint32_t spread4(uint32_t v) { // 4 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

RT_DEV uint32_t ray_sort_key(float4 o, float4 d, const SortGrid &g) {
    const int cx = min(max(int((o.x - g.min_x) * g.inv_x), 0), 15);
    const int cy = min(max(int((o.y - g.min_y) * g.inv_y), 0), 15);
    const int cz = min(max(int((o.z - g.min_z) * g.inv_z), 0), 15);
    const uint32_t morton = spread4(uint32_t(cx)) | (spread4(uint32_t(cy)) << 1) | (spread4(uint32_t(cz)) << 2);
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    return (oct << 12) | morton;
}

__global__ void __launch_bounds__(256) k_sort_hist(const uint32_t *counters, int bounce, RayBuf rays, SortGrid g,
                                                   uint32_t *keys, uint32_t *hist) {
    const uint32_t count = counters[CNT_RAYS + bounce];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t key = ray_sort_key(rays.o_cw[i], rays.d_cs[i], g);
        keys[i] = key;
        atomicAdd(&hist[key], 1u);
    }
}

// exclusive scan of kSortBins counters by one 1024-thread block (32 bins per thread)
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t *hist) {
    __shared__ uint32_t warp_sums[32];
    constexpr int per_thread = kSortBins / 1024;
    const int tid = threadIdx.x;
    uint32_t local[per_thread];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < per_thread; ++j) {
        local[j] = hist[tid * per_thread + j];
        sum += local[j];
    }
    // inclusive scan of `sum` across the block
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) {
            incl += v;
        }
    }
    if ((tid & 31) == 31) {
        warp_sums[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < 32) {
        uint32_t w = warp_sums[tid];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, off);
            if (tid >= off) {
                w += v;
            }
        }
        warp_sums[tid] = w;
    }
    __syncthreads();
    uint32_t base = incl - sum + ((tid >> 5) ? warp_sums[(tid >> 5) - 1] : 0u);
#pragma unroll
    for (int j = 0; j < per_thread; ++j) {
        hist[tid * per_thread + j] = base;
        base += local[j];
    }
}

__global__ void __launch_bounds__(256) k_sort_scatter(const uint32_t *counters, int bounce, RayBuf src, RayBuf dst,
                                                      const uint32_t *keys, uint32_t *offsets, uint32_t *keys_sorted) {
    const uint32_t count = counters[CNT_RAYS + bounce];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        const uint32_t slot = atomicAdd(&offsets[key], 1u);
        dst.o_cw[slot] = src.o_cw[i];
        dst.d_cs[slot] = src.d_cs[i];
        dst.c_pdf[slot] = src.c_pdf[i];
        dst.ior[slot] = src.ior[i];
        dst.xy_depth[slot] = src.xy_depth[i];
        keys_sorted[slot] = key;
    }
}

// Reorders list `bounce` from `src` into `dst` (caller swaps its notion of the current buffer).
inline void sort_rays(SortBufs &s, const KParams &p, const RayBuf &src, const RayBuf &dst, int bounce, int num_sms,
                      cudaStream_t stream) {
    SortGrid g{s.root_min[0], s.root_min[1], s.root_min[2], s.inv_cell[0], s.inv_cell[1], s.inv_cell[2]};
    cudaMemsetAsync(s.hist, 0, kSortBins * sizeof(uint32_t), stream);
    k_sort_hist<<<num_sms * 8, 256, 0, stream>>>(p.counters, bounce, src, g, s.keys, s.hist);
    k_sort_scan<<<1, 1024, 0, stream>>>(s.hist);
    k_sort_scatter<<<num_sms * 8, 256, 0, stream>>>(p.counters, bounce, src, dst, s.keys, s.hist, s.keys_sorted);
}

} // namespace rt
