// rt_lbvh.cuh -- device build of a binary BVH over primitive boxes: Morton codes of the
// box centroids -> radix sort -> Karras' fully parallel radix-tree construction -> bottom-up box fit.
//
// Role in the reference: the "fast" builder PreprocessPrims_HLBVH (internal/Core.cpp:574-720), which trades tree quality
// for build time; here the whole build runs on the GPU in O(n) work after the sort (T. Karras, "Maximizing Parallelism in
// the Construction of BVHs, Octrees, and k-d Trees", HPG 2012).  The result is the binary tree the host layer's
// SAH-optimal 8-wide collapse (BvhBuilder.h CollapseToWideSAH) consumes: internal nodes 0 .. n-2 (root = 0), leaves
// n-1 .. 2n-2 in Morton order, every node with its box, every internal node with its children.
// The sort is cub::DeviceRadixSort (scene-load path, not the per-sample hot path).
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "rt_math.cuh"

namespace rt {

struct LbvhNode { // == RayB200::BinaryNode (host/BvhBuilder.h)
    float mn[3], mx[3];
    uint32_t left, right, first, count;
};

RT_DEV uint32_t expand_bits10(uint32_t v) { // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_lbvh_bounds(const float *__restrict__ boxes, uint32_t n, float *bounds /* min xyz, max xyz as ordered ints */) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        for (int a = 0; a < 3; ++a) {
            const float c = 0.5f * (boxes[i * 6 + a] + boxes[i * 6 + 3 + a]);
            mn[a] = fminf(mn[a], c);
            mx[a] = fmaxf(mx[a], c);
        }
    }
    for (int a = 0; a < 3; ++a) {
        for (int off = 16; off > 0; off >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
        }
        if ((threadIdx.x & 31) == 0) {
            // floats compare like sign-magnitude ints: map to a monotone int key for atomicMin / atomicMax
            auto key = [](float f) {
                const int i = __float_as_int(f);
                return i >= 0 ? i : (i ^ 0x7fffffff);
            };
            atomicMin(reinterpret_cast<int *>(bounds) + a, key(mn[a]));
            atomicMax(reinterpret_cast<int *>(bounds) + 3 + a, key(mx[a]));
        }
    }
}

__global__ void k_lbvh_codes(const float *__restrict__ boxes, uint32_t n, const float *__restrict__ bounds, uint32_t *codes,
                             uint32_t *ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    auto unkey = [](int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); };
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        const float lo = unkey(reinterpret_cast<const int *>(bounds)[a]), hi = unkey(reinterpret_cast<const int *>(bounds)[3 + a]);
        const float c = 0.5f * (boxes[i * 6 + a] + boxes[i * 6 + 3 + a]);
        const float ext = hi - lo;
        const float t = ext > 0.0f ? (c - lo) / ext : 0.0f;
        q[a] = uint32_t(fminf(fmaxf(t * 1024.0f, 0.0f), 1023.0f));
    }
    codes[i] = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    ids[i] = i;
}

// length of the common prefix of the keys (code, position) of sorted leaves i and j; -1 when j is out of range
RT_DEV int lbvh_delta(const uint32_t *__restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) {
        return -1;
    }
    const uint32_t a = codes[i], b = codes[j];
    return a == b ? 32 + __clz(uint32_t(i) ^ uint32_t(j)) : __clz(a ^ b);
}

__global__ void k_lbvh_hierarchy(const uint32_t *__restrict__ codes, int n, LbvhNode *nodes, uint32_t *parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) {
        return;
    }
    const int d = (lbvh_delta(codes, n, i, i + 1) - lbvh_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(codes, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(codes, n, i, i + lmax * d) > dmin) {
        lmax *= 2;
    }
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) {
        if (lbvh_delta(codes, n, i, i + (l + t) * d) > dmin) {
            l += t;
        }
    }
    const int j = i + l * d;
    const int dnode = lbvh_delta(codes, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(codes, n, i, i + (s + t) * d) > dnode) {
            s += t;
        }
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t left = (lo == gamma) ? uint32_t(n - 1 + gamma) : uint32_t(gamma);
    const uint32_t right = (hi == gamma + 1) ? uint32_t(n - 1 + gamma + 1) : uint32_t(gamma + 1);
    nodes[i].left = left;
    nodes[i].right = right;
    nodes[i].first = uint32_t(lo);
    nodes[i].count = 0;
    parent[left] = uint32_t(i);
    parent[right] = uint32_t(i);
    if (i == 0) {
        parent[0] = 0xffffffffu;
    }
}

__global__ void k_lbvh_fit(const float *__restrict__ boxes, const uint32_t *__restrict__ ids, int n, LbvhNode *nodes,
                           const uint32_t *__restrict__ parent, uint32_t *visits) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) {
        return;
    }
    LbvhNode &leaf = nodes[n - 1 + k];
    const uint32_t prim = ids[k];
    for (int a = 0; a < 3; ++a) {
        leaf.mn[a] = boxes[prim * 6 + a];
        leaf.mx[a] = boxes[prim * 6 + 3 + a];
    }
    leaf.left = leaf.right = 0;
    leaf.first = uint32_t(k);
    leaf.count = 1;
    __threadfence();
    uint32_t p = parent[n - 1 + k];
    while (p != 0xffffffffu) {
        if (atomicAdd(&visits[p], 1u) == 0) {
            return; // the sibling subtree is not finished: its thread will carry on
        }
        __threadfence();
        LbvhNode &nd = nodes[p];
        const LbvhNode &a = nodes[nd.left], &b = nodes[nd.right];
        for (int c = 0; c < 3; ++c) {
            nd.mn[c] = fminf(a.mn[c], b.mn[c]);
            nd.mx[c] = fmaxf(a.mx[c], b.mx[c]);
        }
        __threadfence();
        p = parent[p];
    }
}

} // namespace rt
