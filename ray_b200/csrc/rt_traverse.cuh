// rt_traverse.cuh -- BVH8 TLAS->BLAS traversal and the 8-wide plane-form triangle test.
//
// Behavioural spec: reference internal/CoreRef.cpp
//   IntersectTri(mtri_accel_t)                 :54-119    bbox_test_oct(wbvh_node_t)       :323-350
//   TraversalStack sort_top3/4/N               :493-588   Traverse_TLAS_*_ClosestHit(BVH8) :2027-2133
//   Traverse_BLAS_*_ClosestHit(BVH8)           :2495-2578 Traverse_{TLAS,BLAS}_*_AnyHit    :2282-2396, :2695-2787
//   IntersectTris_ClosestHit / _AnyHit (mtris) :1819-1838, :1865-1888
//
// The visit ORDER is part of the contract: a triangle whose t equals the current closest t replaces it (the sign test
// accepts det*t - dett == +0), so two coplanar/duplicated triangles resolve to whichever is tested last, and the
// any-hit query returns at the first SOLID hit it meets.  One thread walks one ray with the reference's exact stack
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

// 8-wide plane-form triangle test with the reference's lane semantics: lane k tests triangle k and then triangle k+4
// against the lane's running t; the winner is the lowest lane holding the minimum t.
RT_DEV bool intersect_mtri(const MTri *__restrict__ tri, v3 ro, v3 rd, int prim_base, Hit &inter) {
    float lt[4], lu[4], lv[4];
    int lp[4];
    bool any_lane[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lt[k] = inter.t;
        lu[k] = 0.0f;
        lv[k] = 0.0f;
        lp[k] = 0;
        any_lane[k] = false;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4 *np = reinterpret_cast<const float4 *>(&tri->n_plane[0][0]) + half;
        const float4 *up = reinterpret_cast<const float4 *>(&tri->u_plane[0][0]) + half;
        const float4 *vp = reinterpret_cast<const float4 *>(&tri->v_plane[0][0]) + half;
        // rows are 8 floats = 2 float4; row r of this half is at index r*2
        const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 2), n2 = __ldg(np + 4), n3 = __ldg(np + 6);
        const float nx[4] = {n0.x, n0.y, n0.z, n0.w}, ny[4] = {n1.x, n1.y, n1.z, n1.w},
                    nz[4] = {n2.x, n2.y, n2.z, n2.w}, nw[4] = {n3.x, n3.y, n3.z, n3.w};
        float det[4], dett[4];
        bool act[4];
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            det[k] = rd.x * nx[k] + rd.y * ny[k] + rd.z * nz[k];
            dett[k] = nw[k] - ro.x * nx[k] - ro.y * ny[k] - ro.z * nz[k];
            act[k] = (__float_as_int(dett[k]) ^ __float_as_int(det[k] * lt[k] - dett[k])) >= 0;
            any |= act[k];
        }
        if (!any) {
            continue;
        }
        const float4 u0 = __ldg(up + 0), u1 = __ldg(up + 2), u2 = __ldg(up + 4), u3 = __ldg(up + 6);
        const float ux[4] = {u0.x, u0.y, u0.z, u0.w}, uy[4] = {u1.x, u1.y, u1.z, u1.w},
                    uz[4] = {u2.x, u2.y, u2.z, u2.w}, uw[4] = {u3.x, u3.y, u3.z, u3.w};
        float px[4], py[4], pz[4], detu[4];
        any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            px[k] = det[k] * ro.x + dett[k] * rd.x;
            py[k] = det[k] * ro.y + dett[k] * rd.y;
            pz[k] = det[k] * ro.z + dett[k] * rd.z;
            detu[k] = px[k] * ux[k] + py[k] * uy[k] + pz[k] * uz[k] + det[k] * uw[k];
            act[k] = act[k] && ((__float_as_int(detu[k]) ^ __float_as_int(det[k] - detu[k])) >= 0);
            any |= act[k];
        }
        if (!any) {
            continue;
        }
        const float4 w0 = __ldg(vp + 0), w1 = __ldg(vp + 2), w2 = __ldg(vp + 4), w3 = __ldg(vp + 6);
        const float vx[4] = {w0.x, w0.y, w0.z, w0.w}, vy[4] = {w1.x, w1.y, w1.z, w1.w},
                    vz[4] = {w2.x, w2.y, w2.z, w2.w}, vw[4] = {w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float detv = px[k] * vx[k] + py[k] * vy[k] + pz[k] * vz[k] + det[k] * vw[k];
            const bool a = act[k] && ((__float_as_int(detv) ^ __float_as_int(det[k] - detu[k] - detv)) >= 0);
            if (a) {
                const float rdet = 1.0f / det[k];
                const int idx = prim_base + half * 4 + k;
                lp[k] = (det[k] < 0.0f) ? idx : (-idx - 1);
                lt[k] = dett[k] * rdet;
                lu[k] = detu[k] * rdet;
                lv[k] = detv * rdet;
                any_lane[k] = true;
            }
        }
    }
    const float min_t = fminf(lt[0], fminf(lt[1], fminf(lt[2], lt[3])));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (any_lane[k] && lt[k] == min_t) {
            inter.prim = lp[k];
            inter.t = lt[k];
            inter.u = lu[k];
            inter.v = lv[k];
            return true;
        }
    }
    return false;
}

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

// Closest-hit (ANY_HIT=false) or first-solid-hit (ANY_HIT=true) query.  `ray_mask` is (1 << ray type).
// Returns: closest -> whether any triangle was hit; any-hit -> whether a SOLID hit was found.
// On return (closest, or any-hit without a solid hit) inter.prim has been resolved through tri_indices exactly as the
// reference does, including for misses (CoreRef.cpp:2125-2130).
template <bool ANY_HIT>
RT_DEV bool traverse_scene(const SceneGeo &sc, v3 ro, v3 rd, uint32_t ray_mask, Hit &inter, StackEntry *st,
                           TraverseCounters &cnt) {
    bool res = false;
    const v3 w_inv_d = safe_invert(rd);

    v3 o = ro, d = rd, inv_d = w_inv_d;
    int sp = 0, base = 0;
    bool in_blas = false;
    int obj_index = -1;

    st[sp++] = StackEntry{sc.tlas_root, 0.0f};

    while (true) {
        if (sp == base) {
            if (!in_blas) {
                break;
            }
            // BLAS drained: back to the TLAS level with the world-space ray
            in_blas = false;
            base = 0;
            o = ro;
            d = rd;
            inv_d = w_inv_d;
            continue;
        }
        StackEntry cur = st[--sp];
        if (cur.dist > inter.t) {
            continue;
        }
        while (true) { // TRAVERSE
            const WNode *__restrict__ n = sc.nodes + cur.index;
            const uint32_t c0 = __ldg(&n->child[0]);
            if ((c0 & kLeafBit) == 0) {
                ++cnt.nodes;
                float dist[8];
                uint32_t mask = box8(&n->bbox_min[0][0], &n->bbox_max[0][0], o, inv_d, inter.t, dist);
                if (mask != 0) {
                    // Empty slots carry a zero-size box at the origin (Core.cpp:849-853) that a ray aimed exactly at
                    // the world origin can "hit"; the reference would then index node 0x7fffffff.  Mask them by id.
                    const uint4 ca = __ldg(reinterpret_cast<const uint4 *>(&n->child[0]));
                    const uint4 cb = __ldg(reinterpret_cast<const uint4 *>(&n->child[4]));
                    const uint32_t valid = uint32_t(ca.x != kEmptyChild) | (uint32_t(ca.y != kEmptyChild) << 1) |
                                           (uint32_t(ca.z != kEmptyChild) << 2) | (uint32_t(ca.w != kEmptyChild) << 3) |
                                           (uint32_t(cb.x != kEmptyChild) << 4) | (uint32_t(cb.y != kEmptyChild) << 5) |
                                           (uint32_t(cb.z != kEmptyChild) << 6) | (uint32_t(cb.w != kEmptyChild) << 7);
                    mask &= valid;
                }
                if (mask == 0) {
                    break;
                }
                int i = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    cur.index = __ldg(&n->child[i]);
                    continue;
                }
                const int i2 = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    if (dist[i] < dist[i2]) {
                        st[sp++] = StackEntry{__ldg(&n->child[i2]), dist[i2]};
                        cur.index = __ldg(&n->child[i]);
                    } else {
                        st[sp++] = StackEntry{__ldg(&n->child[i]), dist[i]};
                        cur.index = __ldg(&n->child[i2]);
                    }
                    continue;
                }
                st[sp++] = StackEntry{__ldg(&n->child[i]), dist[i]};
                st[sp++] = StackEntry{__ldg(&n->child[i2]), dist[i2]};
                i = __ffs(mask) - 1;
                mask &= mask - 1;
                st[sp++] = StackEntry{__ldg(&n->child[i]), dist[i]};
                if (mask == 0) {
                    sort_top3(st, sp);
                    cur.index = st[--sp].index;
                    continue;
                }
                i = __ffs(mask) - 1;
                mask &= mask - 1;
                st[sp++] = StackEntry{__ldg(&n->child[i]), dist[i]};
                if (mask == 0) {
                    sort_top4(st, sp);
                    cur.index = st[--sp].index;
                    continue;
                }
                const int size_before = sp;
                do {
                    i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    st[sp++] = StackEntry{__ldg(&n->child[i]), dist[i]};
                } while (mask != 0);
                sort_topN(st, sp, sp - size_before + 4);
                cur.index = st[--sp].index;
                continue;
            }
            // leaf
            if (in_blas) {
                ++cnt.leaves;
                const int tri_start = int(c0 & kPrimIndexBits);
                const int tri_end = tri_start + int(__ldg(&n->child[1]));
                Hit local;
                local.obj = obj_index;
                local.prim = 0;
                local.t = inter.t;
                local.u = 0.0f;
                local.v = -1.0f;
                for (int b = tri_start / 8; b < (tri_end + 7) / 8; ++b) {
                    intersect_mtri(sc.mtris + b, o, d, b * 8, local);
                    // (the reference's in-leaf early break of the any-hit variant, CoreRef.cpp:1873-1879, is moot:
                    //  a leaf is a single 8-triangle block)
                }
                inter.t = local.t;
                if (local.v >= 0.0f) {
                    inter.obj = local.obj;
                    inter.prim = local.prim;
                    inter.u = local.u;
                    inter.v = local.v;
                    res = true;
                    if (ANY_HIT) {
                        const bool backfacing = inter.prim < 0;
                        const uint32_t slot = backfacing ? uint32_t(-inter.prim - 1) : uint32_t(inter.prim);
                        const TriMat tm = sc.tri_materials[__ldg(&sc.tri_indices[slot])];
                        if ((!backfacing && (tm.front_mi & kMatSolidBit)) || (backfacing && (tm.back_mi & kMatSolidBit))) {
                            return true;
                        }
                    }
                }
                break;
            } else {
                const uint32_t mi_index = c0 & kPrimIndexBits;
                const MeshInstance *__restrict__ mi = sc.instances + mi_index;
                if ((__ldg(&mi->ray_visibility) & ray_mask) == 0) {
                    break;
                }
                const float *__restrict__ m = mi->inv_xform;
                // TransformRay, CoreRef.cpp:2789-2798
                o = v3{m[0] * ro.x + m[4] * ro.y + m[8] * ro.z + m[12], m[1] * ro.x + m[5] * ro.y + m[9] * ro.z + m[13],
                       m[2] * ro.x + m[6] * ro.y + m[10] * ro.z + m[14]};
                d = v3{m[0] * rd.x + m[4] * rd.y + m[8] * rd.z, m[1] * rd.x + m[5] * rd.y + m[9] * rd.z,
                       m[2] * rd.x + m[6] * rd.y + m[10] * rd.z};
                inv_d = safe_invert(d);
                in_blas = true;
                obj_index = int(mi_index);
                base = sp;
                cur.index = __ldg(&mi->node_index);
                cur.dist = 0.0f;
                continue; // the reference pushes the BLAS root with dist 0 and pops it straight away
            }
        }
    }

    // resolve primitive index indirection (also for misses, like the reference)
    if (inter.prim < 0) {
        inter.prim = -int(__ldg(&sc.tri_indices[-inter.prim - 1])) - 1;
    } else {
        inter.prim = int(__ldg(&sc.tri_indices[inter.prim]));
    }
    if (ANY_HIT) {
        return false;
    }
    return res;
}

} // namespace rt
