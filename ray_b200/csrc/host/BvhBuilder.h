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

// Collapse a binary BVH into 8-wide nodes (rt::WNode = wbvh_node_t).  Each binary leaf becomes a leaf WNode whose
// child[0] = LEAF | leaf_payload(leaf) and child[1] = count.  Node ids are relative to out.size() at entry + `base`.
// `leaf_payload` maps a binary leaf to the value stored in child[0] (triangle slot / instance index).
struct WideBuildResult {
    uint32_t root;
};
template <typename LeafFn>
uint32_t CollapseToWide(const std::vector<BinaryNode> &nodes, uint32_t node, std::vector<rt::WNode> &out, uint32_t base,
                        LeafFn &&leaf_payload);

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

} // namespace RayB200
