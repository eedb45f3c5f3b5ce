// rt_traverse.cuh -- shared pieces of the BVH8 walks: the reference's stack discipline (sort_top3/4/N), the 8-wide slab
// test used by the light-tree walks (rt_lights.cuh), the hit record and the scene geometry view.  The scene traversal
// itself (TLAS -> BLAS, triangle test) lives in rt_trace.cuh.
//
// Behavioural spec: reference internal/CoreRef.cpp
//   IntersectTri(mtri_accel_t)                 :54-119    bbox_test_oct(wbvh_node_t)       :323-350
//   TraversalStack sort_top3/4/N               :493-588   Traverse_TLAS_*_ClosestHit(BVH8) :2027-2133
//   Traverse_BLAS_*_ClosestHit(BVH8)           :2495-2578 Traverse_{TLAS,BLAS}_*_AnyHit    :2282-2396, :2695-2787
//   IntersectTris_ClosestHit / _AnyHit (mtris) :1819-1838, :1865-1888
//
// The visit ORDER is part of the contract: a triangle whose t equals the current closest t replaces it (the sign test
// accepts det*t - dett == +0), so two coplanar/duplicated triangles resolve to whichever is tested last, and the
// any-hit query returns at the first SOLID hit it meets.  Every walk keeps the reference's exact stack
// discipline (ordered push of the hit children, sort of the newly pushed group, pop nearest), so the sequence of
// visited nodes -- and therefore every tie -- is the reference's.  TLAS and BLAS levels share one stack array; the
// BLAS level runs on top of the TLAS entries and is drained before the TLAS continues, which is the same order as the
// reference's nested call.
#pragma once

#include "rt_math.cuh"

namespace rt {

struct StackEntry {
    uint32_t index;
    float dist;
};

struct LightStackEntry {
    uint32_t index;
    float dist;
    float factor;
};

// Ordered-push helpers over a caller-owned array; `sp` is the stack size.  All comparisons are the reference's
// (strict `>`/`<` exactly as written there) so equal distances keep the reference's relative order.
template <typename E> RT_DEV void swap_e(E &a, E &b) {
    const E t = a;
    a = b;
    b = t;
}

template <typename E> RT_DEV void sort_top3(E *st, int sp) {
    const int i = sp - 3;
    if (st[i].dist > st[i + 1].dist) {
        if (st[i + 1].dist > st[i + 2].dist) {
            return;
        } else if (st[i].dist > st[i + 2].dist) {
            swap_e(st[i + 1], st[i + 2]);
        } else {
            const E tmp = st[i];
            st[i] = st[i + 2];
            st[i + 2] = st[i + 1];
            st[i + 1] = tmp;
        }
    } else {
        if (st[i].dist > st[i + 2].dist) {
            swap_e(st[i], st[i + 1]);
        } else if (st[i + 2].dist > st[i + 1].dist) {
            swap_e(st[i], st[i + 2]);
        } else {
            const E tmp = st[i];
            st[i] = st[i + 1];
            st[i + 1] = st[i + 2];
            st[i + 2] = tmp;
        }
    }
}

template <typename E> RT_DEV void sort_top4(E *st, int sp) {
    const int i = sp - 4;
    if (st[i + 0].dist < st[i + 1].dist) {
        swap_e(st[i + 0], st[i + 1]);
    }
    if (st[i + 2].dist < st[i + 3].dist) {
        swap_e(st[i + 2], st[i + 3]);
    }
    if (st[i + 0].dist < st[i + 2].dist) {
        swap_e(st[i + 0], st[i + 2]);
    }
    if (st[i + 1].dist < st[i + 3].dist) {
        swap_e(st[i + 1], st[i + 3]);
    }
    if (st[i + 1].dist < st[i + 2].dist) {
        swap_e(st[i + 1], st[i + 2]);
    }
}

template <typename E> RT_DEV void sort_topN(E *st, int sp, int count) {
    const int start = sp - count;
    for (int i = start + 1; i < sp; ++i) {
        const E key = st[i];
        int j = i - 1;
        while (j >= start && st[j].dist < key.dist) {
            st[j + 1] = st[j];
            j--;
        }
        st[j + 1] = key;
    }
}

// 8 slab tests against the children of a wide node.  Returns the hit mask (bit i = child i) and tmin per child.
// min/max of the slab test.  The reference's are SSE min_ps/max_ps, i.e. (a < b ? a : b) / (a > b ? a : b); FMNMX differs
// from that only (1) when an operand is NaN -- impossible here: inv_d is finite by safe_invert (|1/d| <= 1e7) and box
// coordinates are finite (+-MAX_DIST for empty light-tree slots), so no inf - inf or 0 * inf can form -- and (2) in the
// sign of a zero result, which no comparison downstream (tmin <= tmax, tmin <= t, tmax > 0, the distance ordering of
// the stack) can see.  One FMNMX instead of FSETP + FSEL removes ~100 of the ~350 instructions of a node visit.
RT_DEV float box_min(float a, float b) { return fminf(a, b); }
RT_DEV float box_max(float a, float b) { return fmaxf(a, b); }

RT_DEV uint32_t box8(const float *__restrict__ bmin, const float *__restrict__ bmax, v3 o, v3 inv_d, float t,
                     float dist[8]) {
    // bmin/bmax: [3][8] as in wbvh_node_t
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo = inv_d.x * (bmin[0 * 8 + i] - o.x);
        float hi = inv_d.x * (bmax[0 * 8 + i] - o.x);
        float tmin = box_min(lo, hi);
        float tmax = box_max(lo, hi);
        lo = inv_d.y * (bmin[1 * 8 + i] - o.y);
        hi = inv_d.y * (bmax[1 * 8 + i] - o.y);
        tmin = box_max(tmin, box_min(lo, hi));
        tmax = box_min(tmax, box_max(lo, hi));
        lo = inv_d.z * (bmin[2 * 8 + i] - o.z);
        hi = inv_d.z * (bmax[2 * 8 + i] - o.z);
        tmin = box_max(tmin, box_min(lo, hi));
        tmax = box_min(tmax, box_max(lo, hi));
        tmax *= 1.00000024f;
        dist[i] = tmin;
        if ((tmin <= tmax) & (tmin <= t) & (tmax > 0.0f)) {
            mask |= (1u << i);
        }
    }
    return mask;
}

struct Hit {
    int obj, prim;
    float t, u, v;
};

struct SceneGeo {
    const WNode *__restrict__ nodes;  // as uploaded (wbvh_node_t)
    const WNode *__restrict__ dnodes; // device-built copy the trace kernels walk (rt_trace.cuh: k_build_dnodes)
    const uint32_t *__restrict__ blas_roots; // per mesh instance: node word of its BLAS root
    const void *__restrict__ dmtris;         // triangle blocks re-laid out for the 4-lane leaf test (k_build_dmtris)
    uint32_t tlas_root_word;
    const MTri *__restrict__ mtris;
    const uint32_t *__restrict__ tri_indices;
    const TriMat *__restrict__ tri_materials;
    const MeshInstance *__restrict__ instances;
    uint32_t tlas_root;
};

struct TraverseCounters {
    uint32_t nodes, leaves;
};

} // namespace rt
