// BvhBuilder.cpp -- binned SAH build (see BvhBuilder.h for the role this plays).
#include "BvhBuilder.h"

#include "ray_cuda.h"

#include <cmath>
#include <numeric>

namespace RayB200 {

namespace {
constexpr int kBins = 32;

struct Bin {
    Aabb box;
    uint32_t count;
};

struct BuildTask {
    uint32_t node, first, count;
};
} // namespace

void BuildBinaryBVH(const std::vector<Aabb> &prims, const int max_leaf, std::vector<BinaryNode> &nodes,
                    std::vector<uint32_t> &indices) {
    const uint32_t n = uint32_t(prims.size());
    nodes.clear();
    indices.resize(n);
    std::iota(indices.begin(), indices.end(), 0u);
    if (n == 0) {
        return;
    }
    std::vector<float> cx(n), cy(n), cz(n);
    for (uint32_t i = 0; i < n; ++i) {
        cx[i] = 0.5f * (prims[i].mn[0] + prims[i].mx[0]);
        cy[i] = 0.5f * (prims[i].mn[1] + prims[i].mx[1]);
        cz[i] = 0.5f * (prims[i].mn[2] + prims[i].mx[2]);
    }
    const float *cen[3] = {cx.data(), cy.data(), cz.data()};

    nodes.reserve(size_t(n) * 2 / std::max(max_leaf / 2, 1) + 16);
    nodes.emplace_back();
    std::vector<BuildTask> stack;
    stack.push_back({0u, 0u, n});

    while (!stack.empty()) {
        const BuildTask t = stack.back();
        stack.pop_back();

        Aabb box, cbox;
        box.reset();
        cbox.reset();
        for (uint32_t i = t.first; i < t.first + t.count; ++i) {
            const uint32_t p = indices[i];
            box.grow(prims[p]);
            const float c[3] = {cen[0][p], cen[1][p], cen[2][p]};
            cbox.grow(c);
        }
        nodes[t.node].box = box;

        if (int(t.count) <= max_leaf) {
            nodes[t.node].first = t.first;
            nodes[t.node].count = t.count;
            continue;
        }

        // choose the best of 3 axes x (kBins - 1) planes
        int best_axis = -1, best_split = -1;
        float best_cost = 3.402823466e+38F;
        for (int axis = 0; axis < 3; ++axis) {
            const float lo = cbox.mn[axis], ext = cbox.mx[axis] - cbox.mn[axis];
            if (!(ext > 0.0f)) {
                continue;
            }
            Bin bins[kBins];
            for (auto &b : bins) {
                b.box.reset();
                b.count = 0;
            }
            const float scale = float(kBins) / ext;
            for (uint32_t i = t.first; i < t.first + t.count; ++i) {
                const uint32_t p = indices[i];
                int b = int((cen[axis][p] - lo) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                bins[b].box.grow(prims[p]);
                bins[b].count++;
            }
            float right_area[kBins];
            uint32_t right_count[kBins];
            Aabb acc;
            acc.reset();
            uint32_t cnt = 0;
            for (int b = kBins - 1; b > 0; --b) {
                if (bins[b].count) {
                    acc.grow(bins[b].box);
                }
                cnt += bins[b].count;
                right_area[b] = cnt ? acc.half_area() : 0.0f;
                right_count[b] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                if (bins[b].count) {
                    acc.grow(bins[b].box);
                }
                cnt += bins[b].count;
                if (cnt == 0 || right_count[b + 1] == 0) {
                    continue;
                }
                const float cost = acc.half_area() * float(cnt) + right_area[b + 1] * float(right_count[b + 1]);
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = axis;
                    best_split = b;
                }
            }
        }

        uint32_t mid;
        if (best_axis < 0) {
            // all centroids coincide: split the range in the middle
            mid = t.first + t.count / 2;
        } else {
            const float lo = cbox.mn[best_axis], ext = cbox.mx[best_axis] - cbox.mn[best_axis];
            const float scale = float(kBins) / ext;
            const float *c = cen[best_axis];
            auto it = std::partition(indices.begin() + t.first, indices.begin() + t.first + t.count, [&](uint32_t p) {
                int b = int((c[p] - lo) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= best_split;
            });
            mid = uint32_t(it - indices.begin());
            if (mid == t.first || mid == t.first + t.count) {
                mid = t.first + t.count / 2;
            }
        }
        const uint32_t l = uint32_t(nodes.size());
        nodes.emplace_back();
        nodes.emplace_back();
        nodes[t.node].left = l;
        nodes[t.node].right = l + 1;
        nodes[t.node].count = 0;
        stack.push_back({l + 1, mid, t.first + t.count - mid});
        stack.push_back({l, t.first, mid - t.first});
    }
}

bool BuildBinaryLBVH(rc_ctx *ctx, const std::vector<Aabb> &prims, std::vector<BinaryNode> &nodes,
                     std::vector<uint32_t> &indices) {
    static_assert(sizeof(BinaryNode) == sizeof(rc_lbvh_node) && sizeof(Aabb) == 6 * sizeof(float), "layout");
    const uint32_t n = uint32_t(prims.size());
    std::vector<rc_lbvh_node> raw(size_t(2) * n - 1);
    indices.resize(n);
    if (rc_build_lbvh(ctx, &prims[0].mn[0], n, raw.data(), indices.data()) != 0) {
        return false;
    }
    // breadth-first renumbering from the root: parents before children
    nodes.clear();
    nodes.resize(raw.size());
    std::vector<uint32_t> queue(raw.size());
    uint32_t head = 0, tail = 0;
    queue[tail++] = 0;
    while (head < tail) {
        const uint32_t dst = head;
        const rc_lbvh_node &src = raw[queue[head++]];
        BinaryNode &nd = nodes[dst];
        for (int a = 0; a < 3; ++a) {
            nd.box.mn[a] = src.mn[a];
            nd.box.mx[a] = src.mx[a];
        }
        nd.first = src.first;
        nd.count = src.count;
        if (src.count == 0) {
            if (tail + 2 > queue.size()) {
                return false; // not a tree
            }
            nd.left = tail;
            queue[tail++] = src.left;
            nd.right = tail;
            queue[tail++] = src.right;
        }
    }
    return tail == uint32_t(raw.size());
}

} // namespace RayB200
