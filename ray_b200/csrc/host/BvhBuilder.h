// BvhBuilder.h -- host-side acceleration-structure builders of the standalone host layer.
//
// Role in the reference: Ray::PreprocessPrims_SAH + SplitPrimitives_SAH (internal/Core.cpp:492-572,
// internal/BVHSplit.cpp:139) and FlattenBVH_r (internal/Core.cpp:722-857).  Inside the reference tree those are
// REUSED (north_star: "BVHSplit builder and SceneCPU storage reused"); this standalone library cannot link them, so it
// carries its own builders that emit the same node/triangle LAYOUTS (rt_types.h) the kernels consume.  The trees are
// not the reference's trees -- any valid BVH yields the same image up to exact-t tie-breaks.
#pragma once

#include <cstdint>
#include <vector>

#include "../rt_types.h"

struct rc_ctx;

namespace RayB200 {

struct Aabb {
    float mn[3], mx[3];
    void reset() {
        mn[0] = mn[1] = mn[2] = 3.402823466e+38F;
        mx[0] = mx[1] = mx[2] = -3.402823466e+38F;
    }
    void grow(const Aabb &b) {
        for (int i = 0; i < 3; ++i) {
            mn[i] = b.mn[i] < mn[i] ? b.mn[i] : mn[i];
            mx[i] = b.mx[i] > mx[i] ? b.mx[i] : mx[i];
        }
    }
    void grow(const float p[3]) {
        for (int i = 0; i < 3; ++i) {
            mn[i] = p[i] < mn[i] ? p[i] : mn[i];
            mx[i] = p[i] > mx[i] ? p[i] : mx[i];
        }
    }
    float half_area() const {
        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct BinaryNode {
    Aabb box;
    uint32_t left = 0, right = 0; // children (internal)
    uint32_t first = 0, count = 0; // range in the index array (leaf when count != 0)
};

// Binned-SAH top-down build (32 bins per axis).  Leaves hold at most `max_leaf` primitives.  `indices` receives a
// permutation of [0, prims.size()).  Returns the root index (always 0) -- `nodes` is cleared first.
void BuildBinaryBVH(const std::vector<Aabb> &prims, int max_leaf, std::vector<BinaryNode> &nodes,
                    std::vector<uint32_t> &indices);

// Fast build: the binary tree comes from the device (rc_build_lbvh: Morton order + Karras radix tree + bottom-up fit,
// ../rt_lbvh.cuh); this wrapper renumbers it breadth-first so that children follow their parent, the order the collapse
// functions below walk.  Leaves hold one primitive each.  Returns false when the device call failed.
bool BuildBinaryLBVH(rc_ctx *ctx, const std::vector<Aabb> &prims, std::vector<BinaryNode> &nodes,
                     std::vector<uint32_t> &indices);

// Collapse a binary BVH into 8-wide nodes (rt::WNode = wbvh_node_t).  Each binary leaf becomes a leaf WNode whose
// child[0] = LEAF | leaf_payload(leaf) and child[1] = count.  Node ids are relative to out.size() at entry + `base`.
// `leaf_payload` maps a binary leaf to the value stored in child[0] (triangle slot / instance index).
struct WideBuildResult {
    uint32_t root;
};
template <typename LeafFn>
uint32_t CollapseToWide(const std::vector<BinaryNode> &nodes, uint32_t node, std::vector<rt::WNode> &out, uint32_t base,
                        LeafFn &&leaf_payload);

// SAH-optimal collapse of a binary BVH (built down to single primitives) into 8-wide nodes: the dynamic programme of
// Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017), section 3.1.
//   C(n, 1) = min(C_leaf(n), C_internal(n)),   C(n, i) = min(C_distribute(n, i), C(n, i - 1))
//   C_leaf(n) = A(n) * c_leaf  if the subtree holds <= 8 primitives (ONE 8-wide triangle block, whatever its fill)
//   C_internal(n) = C_distribute(n, 8) + A(n) * c_node,   C_distribute(n, j) = min_k C(left, k) + C(right, j - k)
// `leaf_payload(first, count)` receives the primitive range [first, first + count) of the build's index array (subtree
// ranges are contiguous in a top-down partition build) and returns the value stored in child[0].
template <typename LeafFn>
uint32_t CollapseToWideSAH(const std::vector<BinaryNode> &nodes, std::vector<rt::WNode> &out, uint32_t base, float c_node,
                           float c_leaf, LeafFn &&leaf_payload);

} // namespace RayB200

// ---- template implementation ------------------------------------------------------------------------------------
#include <algorithm>
#include <cstring>

namespace RayB200 {

template <typename LeafFn>
uint32_t CollapseToWide(const std::vector<BinaryNode> &nodes, uint32_t node, std::vector<rt::WNode> &out, uint32_t base,
                        LeafFn &&leaf_payload) {
    const uint32_t my_index = uint32_t(out.size());
    out.emplace_back();
    std::memset(&out[my_index], 0, sizeof(rt::WNode));
    const BinaryNode &n = nodes[node];
    if (n.count != 0) {
        rt::WNode &w = out[my_index];
        for (int a = 0; a < 3; ++a) {
            w.bbox_min[a][0] = n.box.mn[a];
            w.bbox_max[a][0] = n.box.mx[a];
        }
        w.child[0] = rt::kLeafBit | leaf_payload(n);
        w.child[1] = n.count;
        return base + my_index;
    }
    // gather up to 8 children: repeatedly open the internal child with the largest surface area
    uint32_t kids[8];
    int nk = 0;
    kids[nk++] = n.left;
    kids[nk++] = n.right;
    while (nk < 8) {
        int best = -1;
        float best_area = -1.0f;
        for (int i = 0; i < nk; ++i) {
            const BinaryNode &c = nodes[kids[i]];
            if (c.count == 0) {
                const float a = c.box.half_area();
                if (a > best_area) {
                    best_area = a;
                    best = i;
                }
            }
        }
        if (best < 0) {
            break;
        }
        const BinaryNode &c = nodes[kids[best]];
        kids[best] = c.left;
        kids[nk++] = c.right;
    }
    uint32_t child_ids[8];
    for (int i = 0; i < 8; ++i) {
        child_ids[i] = (i < nk) ? CollapseToWide(nodes, kids[i], out, base, leaf_payload) : rt::kEmptyChild;
    }
    rt::WNode &w = out[my_index]; // (re-take: `out` may have been reallocated)
    for (int i = 0; i < 8; ++i) {
        w.child[i] = child_ids[i];
        if (i < nk) {
            const Aabb &b = nodes[kids[i]].box;
            for (int a = 0; a < 3; ++a) {
                w.bbox_min[a][i] = b.mn[a];
                w.bbox_max[a][i] = b.mx[a];
            }
        } // empty slots keep the zero box, like the reference (Core.cpp:849-853)
    }
    return base + my_index;
}

template <typename LeafFn>
uint32_t CollapseToWideSAH(const std::vector<BinaryNode> &nodes, std::vector<rt::WNode> &out, uint32_t base,
                           const float c_node, const float c_leaf, LeafFn &&leaf_payload) {
    const uint32_t n = uint32_t(nodes.size());
    struct Dp {
        float c[8];       // c[i] = C(node, i), i = 1..7 (c[0] unused)
        uint8_t split[9]; // split[j] = slots given to the left child by C_distribute(node, j), j = 2..8
        uint8_t internal; // C(node, 1) is realised by an internal wide node (else a leaf)
    };
    std::vector<Dp> dp(n);
    std::vector<uint32_t> sub_first(n), sub_count(n);
    const float kInf = 3.402823466e+38F;
    for (uint32_t i = n; i-- > 0;) { // children have larger indices than their parent: backwards = bottom-up
        const BinaryNode &nd = nodes[i];
        Dp &d = dp[i];
        const float area = nd.box.half_area();
        if (nd.count != 0) {
            sub_first[i] = nd.first;
            sub_count[i] = nd.count;
            for (int k = 1; k < 8; ++k) {
                d.c[k] = area * c_leaf;
            }
            d.internal = 0;
            continue;
        }
        sub_first[i] = std::min(sub_first[nd.left], sub_first[nd.right]);
        sub_count[i] = sub_count[nd.left] + sub_count[nd.right];
        const Dp &l = dp[nd.left], &r = dp[nd.right];
        float dist[9];
        for (int j = 2; j <= 8; ++j) {
            float best = kInf;
            int best_k = 1;
            for (int k = 1; k < j; ++k) {
                if (k > 7 || j - k > 7) {
                    continue;
                }
                const float c = l.c[k] + r.c[j - k];
                if (c < best) {
                    best = c;
                    best_k = k;
                }
            }
            dist[j] = best;
            d.split[j] = uint8_t(best_k);
        }
        const float c_internal = dist[8] + area * c_node;
        const float c_as_leaf = (sub_count[i] <= 8) ? area * c_leaf : kInf;
        d.internal = (c_internal < c_as_leaf) ? 1 : 0;
        d.c[1] = d.internal ? c_internal : c_as_leaf;
        for (int j = 2; j < 8; ++j) {
            d.c[j] = std::min(dist[j], d.c[j - 1]);
        }
    }
    struct Emit {
        const std::vector<BinaryNode> &nodes;
        const std::vector<Dp> &dp;
        const std::vector<uint32_t> &sub_first, &sub_count;
        std::vector<rt::WNode> &out;
        uint32_t base;
        LeafFn &leaf_payload;
        void collect(uint32_t node, int j, uint32_t *list, int &count) const {
            const BinaryNode &nd = nodes[node];
            if (j == 1 || nd.count != 0) {
                list[count++] = node;
                return;
            }
            const Dp &d = dp[node];
            // C(node, j) = min(C_distribute(node, j), C(node, j - 1)): take fewer slots while that is as good
            const float dist_j = dp[nd.left].c[d.split[j]] + dp[nd.right].c[j - d.split[j]];
            if (j < 8 && !(dist_j <= d.c[j - 1])) {
                collect(node, j - 1, list, count);
                return;
            }
            collect(nd.left, d.split[j], list, count);
            collect(nd.right, j - d.split[j], list, count);
        }
        uint32_t emit(uint32_t node) const {
            const uint32_t my_index = uint32_t(out.size());
            out.emplace_back();
            std::memset(&out[my_index], 0, sizeof(rt::WNode));
            const BinaryNode &nd = nodes[node];
            if (nd.count != 0 || !dp[node].internal) {
                rt::WNode &w = out[my_index];
                for (int a = 0; a < 3; ++a) {
                    w.bbox_min[a][0] = nd.box.mn[a];
                    w.bbox_max[a][0] = nd.box.mx[a];
                }
                const uint32_t cnt = sub_count[node];
                const uint32_t payload = leaf_payload(sub_first[node], cnt);
                out[my_index].child[0] = rt::kLeafBit | payload;
                out[my_index].child[1] = cnt;
                return base + my_index;
            }
            uint32_t kids[8];
            int nk = 0;
            collect(nd.left, dp[node].split[8], kids, nk);
            collect(nd.right, 8 - dp[node].split[8], kids, nk);
            uint32_t child_ids[8];
            for (int i = 0; i < 8; ++i) {
                child_ids[i] = (i < nk) ? emit(kids[i]) : rt::kEmptyChild;
            }
            rt::WNode &w = out[my_index];
            for (int i = 0; i < 8; ++i) {
                w.child[i] = child_ids[i];
                if (i < nk) {
                    const Aabb &b = nodes[kids[i]].box;
                    for (int a = 0; a < 3; ++a) {
                        w.bbox_min[a][i] = b.mn[a];
                        w.bbox_max[a][i] = b.mx[a];
                    }
                }
            }
            return base + my_index;
        }
    };
    const Emit e{nodes, dp, sub_first, sub_count, out, base, leaf_payload};
    return e.emit(0);
}

} // namespace RayB200
