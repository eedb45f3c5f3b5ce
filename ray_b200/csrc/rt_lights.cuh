// rt_lights.cuh -- light tree (quantised 8-wide), light sampling for NEE, analytic-light intersection.
//
// Behavioural spec: reference internal/CoreRef.cpp
//   bbox_test_oct(cwbvh_node_t)        :243-281, :393-479     calc_lnode_importance(cwbvh)      :1004-1066
//   decode_oct_dir / decode_cosines    :914-956               map_to_cone                        :691-714
//   SampleSphericalRectangle           :1288-1350             SampleSphericalTriangle            :1354-1416
//   slerp / orthogonalize / angle_between :1103-1126, :1278-1284
//   SampleLightSource                  :3264-3614             IntersectAreaLights(rays, cwbvh)   :3616-3860
//   IntersectAreaLights(shadow, cwbvh) :4451-4592             EvalTriLightFactor(cwbvh)          :4692-4736
#pragma once

#include "rt_env.cuh"
#include "rt_traverse.cuh"

namespace rt {

// light_t bit-field word (Core.h:194-201)
RT_DEV int l_type(const Light &l) { return int(l.bits & 7u); }
RT_DEV bool l_doublesided(const Light &l) { return (l.bits >> 3) & 1u; }
RT_DEV bool l_cast_shadow(const Light &l) { return (l.bits >> 4) & 1u; }
RT_DEV bool l_visible(const Light &l) { return (l.bits >> 5) & 1u; }
RT_DEV bool l_sky_portal(const Light &l) { return (l.bits >> 6) & 1u; }
RT_DEV uint32_t l_ray_visibility(const Light &l) { return (l.bits >> 7) & 0xffu; }

struct LightSample { // light_sample_t, CoreRef.h:122-130 (col/L/lp start uninitialised in the reference too)
    v3 col, L, lp;
    float area, dist_mul, pdf;
    bool cast_shadow, from_env;
    uint32_t ray_flags;
};

struct SceneLights {
    const Light *__restrict__ lights;
    const LightCWNode *__restrict__ nodes;
    uint32_t nodes_count;
    uint32_t visible_lights_count, blocker_lights_count;
    uint32_t env_light_index;
    float env_col[3], back_col[3];
    SceneEnv env;
};

// Unpack the 8 quantised child boxes of a light-tree node into [3][8] arrays.
RT_FN void unpack_cw_bounds(const LightCWNode &n, float bmin[24], float bmax[24]) {
    const float ext0 = (n.bbox_max[0] - n.bbox_min[0]) / 255.0f, ext1 = (n.bbox_max[1] - n.bbox_min[1]) / 255.0f,
                ext2 = (n.bbox_max[2] - n.bbox_min[2]) / 255.0f;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        bmin[0 * 8 + i] = bmin[1 * 8 + i] = bmin[2 * 8 + i] = -kMaxDist;
        bmax[0 * 8 + i] = bmax[1 * 8 + i] = bmax[2 * 8 + i] = kMaxDist;
        if (n.ch_bbox_min[0][i] != 0xff || n.ch_bbox_max[0][i] != 0) {
            bmin[0 * 8 + i] = n.bbox_min[0] + float(n.ch_bbox_min[0][i]) * ext0;
            bmin[1 * 8 + i] = n.bbox_min[1] + float(n.ch_bbox_min[1][i]) * ext1;
            bmin[2 * 8 + i] = n.bbox_min[2] + float(n.ch_bbox_min[2][i]) * ext2;
            bmax[0 * 8 + i] = n.bbox_min[0] + float(n.ch_bbox_max[0][i]) * ext0;
            bmax[1 * 8 + i] = n.bbox_min[1] + float(n.ch_bbox_max[1][i]) * ext1;
            bmax[2 * 8 + i] = n.bbox_min[2] + float(n.ch_bbox_max[2][i]) * ext2;
        }
    }
}

RT_DEV float sse_abs(float v) { return sse_max(v, -v); } // abs(fvec4) = max(v, -v), simd.h:562-565

// calc_lnode_importance for the quantised node, one child lane at a time (the reference works on 4 lanes at once; all
// operations are lane-wise, so the per-lane result is identical).
RT_FN void lnode_importance(const LightCWNode &n, const float bmin[24], const float bmax[24], v3 P, float imp[8]) {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        float v = n.flux[i];
        // A zero-flux lane (empty slot, Core.cpp:1176-1183) can only produce +-0 below (flux * mul with a finite or
        // zeroed mul), and +-0 behaves identically in every sum, quotient and comparison downstream: skip the math.
        if (v != 0.0f && bmin[0 * 8 + i] > -kMaxDist) {
            // decode_oct_dir (vector form)
            const uint32_t oct = n.axis[i];
            float a0 = -1.0f + 2.0f * float((oct >> 16) & 0xffffu) / 65535.0f;
            float a1 = -1.0f + 2.0f * float(oct & 0xffffu) / 65535.0f;
            float a2 = 1.0f - sse_abs(a0) - sse_abs(a1);
            if (a2 < 0.0f) {
                const float temp = a0;
                a0 = (1.0f - sse_abs(a1)) * copysignf(1.0f, temp);
                a1 = (1.0f - sse_abs(temp)) * copysignf(1.0f, a1);
            }
            const float al = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
            a0 = a0 / al;
            a1 = a1 / al;
            a2 = a2 / al;

            const float e0 = bmax[0 * 8 + i] - bmin[0 * 8 + i], e1 = bmax[1 * 8 + i] - bmin[1 * 8 + i],
                        e2 = bmax[2 * 8 + i] - bmin[2 * 8 + i];
            const float extent = 0.5f * sqrtf(e0 * e0 + e1 * e1 + e2 * e2);

            const float pc0 = 0.5f * (bmin[0 * 8 + i] + bmax[0 * 8 + i]), pc1 = 0.5f * (bmin[1 * 8 + i] + bmax[1 * 8 + i]),
                        pc2 = 0.5f * (bmin[2 * 8 + i] + bmax[2 * 8 + i]);
            float w0 = P.x - pc0, w1 = P.y - pc1, w2 = P.z - pc2;
            const float dist2 = w0 * w0 + w1 * w1 + w2 * w2;
            const float dist = sqrtf(dist2);
            w0 /= dist;
            w1 /= dist;
            w2 /= dist;

            const float v_len2 = sse_max(dist2, extent);

            const float cos_omega_w = a0 * w0 + a1 * w1 + a2 * w2;
            const float sin_omega_w = sqrtf(sse_max(1.0f - cos_omega_w * cos_omega_w, 0.0f));

            float cos_omega_b = sqrtf(sse_max(1.0f - (extent * extent) / dist2, 0.0f));
            if (dist2 < extent * extent) {
                cos_omega_b = -1.0f;
            }
            const float sin_omega_b = sqrtf(1.0f - cos_omega_b * cos_omega_b);

            const uint32_t cv = n.cos_omega_ne[i];
            const float cos_omega_n = 2.0f * (float((cv >> 16) & 0xffffu) / 65534.0f) - 1.0f;
            const float cos_omega_e = 2.0f * (float(cv & 0xffffu) / 65534.0f) - 1.0f;
            const float sin_omega_n = sqrtf(1.0f - cos_omega_n * cos_omega_n);

            float cos_omega_x = cos_omega_w * cos_omega_n + sin_omega_w * sin_omega_n;
            float sin_omega_x = sin_omega_w * cos_omega_n - cos_omega_w * sin_omega_n;
            if (cos_omega_w > cos_omega_n) {
                cos_omega_x = 1.0f;
                sin_omega_x = 0.0f;
            }
            float cos_omega = cos_omega_x * cos_omega_b + sin_omega_x * sin_omega_b;
            if (cos_omega_x > cos_omega_b) {
                cos_omega = 1.0f;
            }
            float mul = 0.0f;
            if (cos_omega > cos_omega_e) {
                mul = cos_omega / v_len2;
            }
            v = v * mul;
        }
        imp[i] = v;
    }
}

// calc_lnode_importance straight from the quantised node: the child's box is decoded where it is used, so the 48
// decoded bounds never exist as (local-memory) arrays.  Same operations per child as unpack_cw_bounds +
// lnode_importance above.
RT_FN void lnode_importance_q(const LightCWNode &n, v3 P, float imp[8]) {
    const float bm0 = n.bbox_min[0], bm1 = n.bbox_min[1], bm2 = n.bbox_min[2];
    const float ext0 = (n.bbox_max[0] - bm0) / 255.0f, ext1 = (n.bbox_max[1] - bm1) / 255.0f,
                ext2 = (n.bbox_max[2] - bm2) / 255.0f;
    // 6 rows of 8 quantised bytes
    const uint2 *q = reinterpret_cast<const uint2 *>(&n.ch_bbox_min[0][0]);
    const uint2 qmin0 = q[0], qmin1 = q[1], qmin2 = q[2], qmax0 = q[3], qmax1 = q[4], qmax2 = q[5];
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        float v = n.flux[i];
        const int sh = (i & 3) * 8;
        const bool hi = i >= 4;
        const uint32_t cmin0 = ((hi ? qmin0.y : qmin0.x) >> sh) & 0xffu, cmax0 = ((hi ? qmax0.y : qmax0.x) >> sh) & 0xffu;
        // (an "empty" child -- min 0xff / max 0 on axis 0 -- decodes to a box at -/+MAX_DIST and is skipped below)
        if (v != 0.0f && (cmin0 != 0xffu || cmax0 != 0u)) {
            const uint32_t cmin1 = ((hi ? qmin1.y : qmin1.x) >> sh) & 0xffu, cmin2 = ((hi ? qmin2.y : qmin2.x) >> sh) & 0xffu,
                           cmax1 = ((hi ? qmax1.y : qmax1.x) >> sh) & 0xffu, cmax2 = ((hi ? qmax2.y : qmax2.x) >> sh) & 0xffu;
            const float bmin0 = bm0 + float(cmin0) * ext0, bmin1 = bm1 + float(cmin1) * ext1, bmin2 = bm2 + float(cmin2) * ext2;
            const float bmax0 = bm0 + float(cmax0) * ext0, bmax1 = bm1 + float(cmax1) * ext1, bmax2 = bm2 + float(cmax2) * ext2;
            if (bmin0 > -kMaxDist) {
                // decode_oct_dir (vector form)
                const uint32_t oct = n.axis[i];
                float a0 = -1.0f + 2.0f * float((oct >> 16) & 0xffffu) / 65535.0f;
                float a1 = -1.0f + 2.0f * float(oct & 0xffffu) / 65535.0f;
                float a2 = 1.0f - sse_abs(a0) - sse_abs(a1);
                if (a2 < 0.0f) {
                    const float temp = a0;
                    a0 = (1.0f - sse_abs(a1)) * copysignf(1.0f, temp);
                    a1 = (1.0f - sse_abs(temp)) * copysignf(1.0f, a1);
                }
                const float al = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
                a0 = a0 / al;
                a1 = a1 / al;
                a2 = a2 / al;

                const float e0 = bmax0 - bmin0, e1 = bmax1 - bmin1, e2 = bmax2 - bmin2;
                const float extent = 0.5f * sqrtf(e0 * e0 + e1 * e1 + e2 * e2);

                const float pc0 = 0.5f * (bmin0 + bmax0), pc1 = 0.5f * (bmin1 + bmax1), pc2 = 0.5f * (bmin2 + bmax2);
                float w0 = P.x - pc0, w1 = P.y - pc1, w2 = P.z - pc2;
                const float dist2 = w0 * w0 + w1 * w1 + w2 * w2;
                const float dist = sqrtf(dist2);
                w0 /= dist;
                w1 /= dist;
                w2 /= dist;

                const float v_len2 = sse_max(dist2, extent);

                const float cos_omega_w = a0 * w0 + a1 * w1 + a2 * w2;
                const float sin_omega_w = sqrtf(sse_max(1.0f - cos_omega_w * cos_omega_w, 0.0f));

                float cos_omega_b = sqrtf(sse_max(1.0f - (extent * extent) / dist2, 0.0f));
                if (dist2 < extent * extent) {
                    cos_omega_b = -1.0f;
                }
                const float sin_omega_b = sqrtf(1.0f - cos_omega_b * cos_omega_b);

                const uint32_t cv = n.cos_omega_ne[i];
                const float cos_omega_n = 2.0f * (float((cv >> 16) & 0xffffu) / 65534.0f) - 1.0f;
                const float cos_omega_e = 2.0f * (float(cv & 0xffffu) / 65534.0f) - 1.0f;
                const float sin_omega_n = sqrtf(1.0f - cos_omega_n * cos_omega_n);

                float cos_omega_x = cos_omega_w * cos_omega_n + sin_omega_w * sin_omega_n;
                float sin_omega_x = sin_omega_w * cos_omega_n - cos_omega_w * sin_omega_n;
                if (cos_omega_w > cos_omega_n) {
                    cos_omega_x = 1.0f;
                    sin_omega_x = 0.0f;
                }
                float cos_omega = cos_omega_x * cos_omega_b + sin_omega_x * sin_omega_b;
                if (cos_omega_x > cos_omega_b) {
                    cos_omega = 1.0f;
                }
                float mul = 0.0f;
                if (cos_omega > cos_omega_e) {
                    mul = cos_omega / v_len2;
                }
                v = v * mul;
            }
        }
        imp[i] = v;
    }
}

// hsum(imp[0..3] + imp[4..7]) with the SSE2 association
RT_DEV float sum_importance(const float imp[8]) {
    return (imp[0] + imp[4]) + (imp[1] + imp[5]) + (imp[2] + imp[6]) + (imp[3] + imp[7]);
}

RT_FN v3 map_to_cone(float r1, float r2, v3 N, float radius) {
    const float ox = 2.0f * r1 - 1.0f, oy = 2.0f * r2 - 1.0f;
    if (ox == 0.0f && oy == 0.0f) {
        return N;
    }
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) {
        r = ox;
        theta = 0.25f * kPi * (oy / ox);
    } else {
        r = oy;
        theta = 0.5f * kPi * (1.0f - 0.5f * (ox / oy));
    }
    const v2 sc = portable_sincos(theta);
    const float ux = radius * r * sc.y, uy = radius * r * sc.x;
    v3 LT, LB;
    create_tbn(normalize(N), LT, LB);
    return N + ux * LT + uy * LB;
}

RT_DEV float sphere_intersection(v3 center, float radius, v3 ro, v3 rd) {
    const v3 oc = ro - center;
    const float a = dot(rd, rd);
    const float b = 2 * dot(oc, rd);
    const float c = dot(oc, oc) - radius * radius;
    const float discriminant = b * b - 4 * a * c;
    return (-b - sqrtf(fmaxf(discriminant, 0.0f))) / (2 * a);
}

RT_DEV bool quadratic(float a, float b, float c, float &t0, float &t1) {
    const float d = b * b - 4.0f * a * c;
    if (d < 0.0f) {
        return false;
    }
    const float sqrt_d = sqrtf(d);
    float q;
    if (b < 0.0f) {
        q = -0.5f * (b - sqrt_d);
    } else {
        q = -0.5f * (b + sqrt_d);
    }
    t0 = q / a;
    t1 = c / q;
    return true;
}

RT_FN v3 orthogonalize(v3 a, v3 b) { return normalize(b - dot(a, b) * a); }

RT_FN v3 slerp(v3 start, v3 end, float percent) {
    float cos_theta = dot(start, end);
    cos_theta = clampf(cos_theta, -1.0f, 1.0f);
    const float theta = libm_acosf(cos_theta) * percent;
    const v3 relative_vec = safe_normalize(end - start * cos_theta);
    const v2 sc = portable_sincos(theta);
    return start * sc.y + relative_vec * sc.x;
}

RT_FN float angle_between(v3 v1, v3 v2) {
    if (dot(v1, v2) < 0) {
        return kPi - 2 * portable_asinf(length(v1 + v2) / 2);
    } else {
        return 2 * portable_asinf(length(v2 - v1) / 2);
    }
}

// Returns pdf (1/solid angle) or 0; writes the sampled point when out_p != nullptr.
RT_FN float sample_spherical_rectangle(v3 P, v3 light_pos, v3 axis_u, v3 axis_v, v2 Xi, v3 *out_p) {
    const v3 corner = light_pos - 0.5f * axis_u - 0.5f * axis_v;
    float axisu_len, axisv_len;
    const v3 x = normalize_len(axis_u, axisu_len), y = normalize_len(axis_v, axisv_len);
    v3 z = cross(x, y);
    const v3 dir = corner - P;
    float z0 = dot(dir, z);
    if (z0 > 0.0f) {
        z = -z;
        z0 = -z0;
    }
    const float x0 = dot(dir, x);
    const float y0 = dot(dir, y);
    const float x1 = x0 + axisu_len;
    const float y1 = y0 + axisv_len;
    // lanes: diff = {x0,y1,x1,y0} - {x1,y0,x0,y1}; nz = {y0,x1,y1,x0} * diff
    const float df[4] = {x0 - x1, y1 - y0, x1 - x0, y0 - y1};
    const float nm[4] = {y0, x1, y1, x0};
    float nz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float n = nm[k] * df[k];
        nz[k] = n / sqrtf(z0 * z0 * df[k] * df[k] + n * n);
    }
    const float g0 = portable_acosf(clampf(-nz[0] * nz[1], -1.0f, 1.0f));
    const float g1 = portable_acosf(clampf(-nz[1] * nz[2], -1.0f, 1.0f));
    const float g2 = portable_acosf(clampf(-nz[2] * nz[3], -1.0f, 1.0f));
    const float g3 = portable_acosf(clampf(-nz[3] * nz[0], -1.0f, 1.0f));
    const float b0 = nz[0];
    const float b1 = nz[2];
    const float b0sq = b0 * b0;
    const float k = 2 * kPi - g2 - g3;
    const float area = g0 + g1 - k;
    if (area <= kSphericalAreaThreshold) {
        return 0.0f;
    }
    if (out_p) {
        const float au = Xi.x * area + k;
        const v2 sc = portable_sincos(au);
        const float fu = safe_div((sc.y * b0 - b1), sc.x);
        float cu = 1.0f / sqrtf(fu * fu + b0sq) * (fu > 0.0f ? 1.0f : -1.0f);
        cu = clampf(cu, -1.0f, 1.0f);
        float xu = -(cu * z0) / fmaxf(sqrtf(1.0f - cu * cu), 1e-7f);
        xu = clampf(xu, x0, x1);
        const float z0sq = z0 * z0;
        const float y0sq = y0 * y0;
        const float y1sq = y1 * y1;
        const float d = sqrtf(xu * xu + z0sq);
        const float h0 = y0 / sqrtf(d * d + y0sq);
        const float h1 = y1 / sqrtf(d * d + y1sq);
        const float hv = h0 + Xi.y * (h1 - h0), hv2 = hv * hv;
        const float yv = (hv2 < 1.0f - 1e-6f) ? (hv * d) / sqrtf(1.0f - hv2) : y1;
        (*out_p) = P + xu * x + yv * y + z0 * z;
    }
    return (1.0f / area);
}

RT_FN float sample_spherical_triangle(v3 P, v3 p1, v3 p2, v3 p3, v2 Xi, v3 *out_dir) {
    const v3 A = normalize(p1 - P), B = normalize(p2 - P), C = normalize(p3 - P);
    const v3 BA = orthogonalize(A, B - A);
    const v3 CA = orthogonalize(A, C - A);
    const v3 AB = orthogonalize(B, A - B);
    const v3 CB = orthogonalize(B, C - B);
    const v3 BC = orthogonalize(C, B - C);
    const v3 AC = orthogonalize(C, A - C);
    const float alpha = angle_between(BA, CA);
    const float beta = angle_between(AB, CB);
    const float gamma = angle_between(BC, AC);
    const float area = alpha + beta + gamma - kPi;
    if (area <= kSphericalAreaThreshold) {
        return 0.0f;
    }
    if (out_dir) {
        const float b = portable_acosf(clampf(dot(C, A), -1.0f, 1.0f));
        const float c = portable_acosf(clampf(dot(A, B), -1.0f, 1.0f));
        const float area_S = Xi.x * area;
        const v2 sc_area = portable_sincos(area_S - alpha);
        const float p = sc_area.x;
        const float q = sc_area.y;
        const v2 sc_alpha = portable_sincos(alpha);
        const float u = q - sc_alpha.y;
        const float v = p + sc_alpha.x * portable_cos(c);
        const float denom = ((v * p + u * q) * sc_alpha.x);
        const float s = safe_div(1.0f, b) *
                        portable_acosf(clampf(safe_div(((v * q - u * p) * sc_alpha.y - v), denom), -1.0f, 1.0f));
        const v3 C_s = slerp(A, C, s);
        const float denom2 = portable_acosf(clampf(dot(C_s, B), -1.0f, 1.0f));
        const float t = safe_div(portable_acosf(clampf(1.0f - Xi.y * (1.0f - dot(C_s, B)), -1.0f, 1.0f)), denom2);
        (*out_dir) = slerp(B, C_s, t);
    }
    return (1.0f / area);
}

struct SceneSurf { // the arrays shading needs besides SceneGeo
    const Vertex *__restrict__ vertices;
    const uint32_t *__restrict__ vtx_indices;
    const Material *__restrict__ materials;
};

// SampleLightSource with hierarchical NEE (USE_HIERARCHICAL_NEE, USE_SPHERICAL_AREA_LIGHT_SAMPLING = true).
// Textured lights / env maps are not supported by this backend (rc_upload_scene rejects them).
RT_FN void sample_light_source(const bool tex_on, v3 P, v3 T, v3 B, v3 N, const SceneLights &sl, const SceneGeo &sg, const SceneSurf &ss,
                                const SceneTex &tx, float rand_pick_light, v2 rand_light_uv, v2 rand_tex_uv,
                                LightSample &ls) {
    float u1 = rand_pick_light;
    float factor = 1.0f;
    uint32_t i = 0;
    while ((i & kLeafBit) == 0) {
        const LightCWNode &n = sl.nodes[i];
        float importance[8];
        lnode_importance_q(n, P, importance);
        const float total_importance = sum_importance(importance);
        if (total_importance == 0.0f) {
            return; // no light can be sampled from here
        }
        float factors[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            factors[j] = importance[j] / total_importance;
        }
        float cdf[9];
        cdf[0] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cdf[j + 1] = cdf[j] + factors[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (cdf[j + 1] == cdf[8]) {
                cdf[j + 1] = 1.01f;
            }
        }
        int next = 0;
#pragma unroll
        for (int j = 1; j <= 8; ++j) {
            next += (cdf[j] <= u1) ? 1 : 0;
        }
        float f_next = factors[0], c_next = cdf[0];
        uint32_t ch_next = n.child[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            if (next == j) {
                f_next = factors[j];
                c_next = cdf[j];
                ch_next = n.child[j];
            }
        }
        u1 = fractf((u1 - c_next) / f_next);
        i = ch_next;
        factor *= f_next;
    }
    const uint32_t light_index = (i & kPrimIndexBits);
    factor = 1.0f / factor;

    const Light &l = sl.lights[light_index];
    const int type = l_type(l);
    ls.col = mk3(l.col);
    ls.cast_shadow = l_cast_shadow(l);
    ls.from_env = false;

    if (type == LIGHT_SPHERE) {
        const float r1 = rand_light_uv.x, r2 = rand_light_uv.y;
        const v3 center = mk3(&l.p[0]);
        const float radius = l.p[7];
        float d;
        const v3 light_normal = normalize_len(center - P, d);
        if (d > radius) {
            const float temp = sqrtf(d * d - radius * radius);
            const float disk_radius = (temp * radius) / d;
            float disk_dist = radius > 0.0f ? ((temp * disk_radius) / radius) : d;
            const v3 sampled_dir = normalize_len(map_to_cone(r1, r2, disk_dist * light_normal, disk_radius), disk_dist);
            if (radius > 0.0f) {
                const float ls_dist = sphere_intersection(center, radius, P, sampled_dir);
                const v3 light_surf_pos = P + sampled_dir * ls_dist;
                const v3 light_forward = normalize(light_surf_pos - center);
                const float sampled_area = kPi * disk_radius * disk_radius;
                const float cos_theta = dot(sampled_dir, light_normal);
                ls.lp = offset_ray(light_surf_pos, light_forward);
                ls.pdf = (disk_dist * disk_dist) / (sampled_area * cos_theta);
            } else {
                ls.lp = center;
                ls.pdf = (disk_dist * disk_dist) / kPi;
            }
            ls.L = sampled_dir;
            ls.area = kPi * disk_radius * disk_radius;
            ls.ray_flags = l_ray_visibility(l);
            if (!l_visible(l)) {
                ls.area = 0.0f;
            }
            const float spot = l.p[8], blend = l.p[9];
            if (spot > 0.0f) {
                const float _dot = -dot(ls.L, mk3(&l.p[4]));
                if (_dot > 0.0f) {
                    const float _angle = libm_acosf(saturatef(_dot));
                    ls.col *= saturatef((spot - _angle) / blend);
                } else {
                    ls.col *= 0.0f;
                }
            }
        }
    } else if (type == LIGHT_DIR) {
        const v3 ldir = mk3(&l.p[0]);
        const float tan_angle = l.p[4];
        ls.L = ldir;
        ls.area = 0.0f;
        ls.pdf = 1.0f;
        if (tan_angle != 0.0f) {
            const float radius = tan_angle;
            ls.L = normalize(map_to_cone(rand_light_uv.x, rand_light_uv.y, ls.L, radius));
            ls.area = kPi * radius * radius;
            const float cos_theta = dot(ls.L, ldir);
            ls.pdf = 1.0f / (ls.area * cos_theta);
        }
        ls.lp = P + ls.L;
        ls.dist_mul = kMaxDist;
        ls.ray_flags = l_ray_visibility(l);
        if (!l_visible(l)) {
            ls.area = 0.0f;
        }
    } else if (type == LIGHT_RECT) {
        const v3 light_pos = mk3(&l.p[0]);
        const v3 light_u = mk3(&l.p[4]), light_v = mk3(&l.p[8]);
        const float rect_area = l.p[3];
        const v3 light_forward = normalize(cross(light_u, light_v));
        v3 lp;
        float pdf = sample_spherical_rectangle(P, light_pos, light_u, light_v, rand_light_uv, &lp);
        if (pdf <= 0.0f) {
            const float r1 = rand_light_uv.x - 0.5f, r2 = rand_light_uv.y - 0.5f;
            lp = light_pos + light_u * r1 + light_v * r2;
        }
        float ls_dist;
        ls.L = normalize_len(lp - P, ls_dist);
        ls.ray_flags = l_ray_visibility(l);
        const float cos_theta = dot(-ls.L, light_forward);
        if (cos_theta > 0.0f) {
            ls.lp = offset_ray(lp, light_forward);
            ls.pdf = (pdf > 0.0f) ? pdf : (ls_dist * ls_dist) / (rect_area * cos_theta);
            ls.area = l_visible(l) ? rect_area : 0.0f;
            if (l_sky_portal(l)) {
                v3 env_col = mk3(sl.env_col);
                if (tex_on && sl.env.env_map != kTexInvalid) {
                    env_col *= sample_latlong_rgbe(tx, sl.env.env_map, ls.L, sl.env.env_map_rotation, rand_tex_uv);
                }
                ls.col *= env_col;
                ls.from_env = true;
            }
        }
    } else if (type == LIGHT_DISK) {
        const v3 light_pos = mk3(&l.p[0]);
        const v3 light_u = mk3(&l.p[4]), light_v = mk3(&l.p[8]);
        float ox = 2.0f * rand_light_uv.x - 1.0f, oy = 2.0f * rand_light_uv.y - 1.0f;
        if (ox != 0.0f && oy != 0.0f) {
            float theta, r;
            if (fabsf(ox) > fabsf(oy)) {
                r = ox;
                theta = 0.25f * kPi * (oy / ox);
            } else {
                r = oy;
                theta = 0.5f * kPi - 0.25f * kPi * (ox / oy);
            }
            const v2 sc = portable_sincos(theta);
            ox = 0.5f * r * sc.y;
            oy = 0.5f * r * sc.x;
        }
        const v3 lp = light_pos + light_u * ox + light_v * oy;
        const v3 light_forward = normalize(cross(light_u, light_v));
        ls.lp = offset_ray(lp, light_forward);
        float ls_dist;
        ls.L = normalize_len(lp - P, ls_dist);
        ls.area = l.p[3];
        ls.ray_flags = l_ray_visibility(l);
        const float cos_theta = dot(-ls.L, light_forward);
        if (cos_theta > 0.0f) {
            ls.pdf = (ls_dist * ls_dist) / (ls.area * cos_theta);
        }
        if (!l_visible(l)) {
            ls.area = 0.0f;
        }
        if (l_sky_portal(l)) {
            v3 env_col = mk3(sl.env_col);
            if (tex_on && sl.env.env_map != kTexInvalid) {
                env_col *= sample_latlong_rgbe(tx, sl.env.env_map, ls.L, sl.env.env_map_rotation, rand_tex_uv);
            }
            ls.col *= env_col;
            ls.from_env = true;
        }
    } else if (type == LIGHT_LINE) {
        const v3 light_pos = mk3(&l.p[0]);
        const v3 light_dir = mk3(&l.p[8]);
        const float radius = l.p[7], height = l.p[11];
        const float r1 = rand_light_uv.x, r2 = rand_light_uv.y;
        const v3 center_to_surface = P - light_pos;
        const v3 light_u = normalize(cross(center_to_surface, light_dir));
        const v3 light_v = cross(light_u, light_dir);
        const float phi = kPi * r1;
        const v2 sc = portable_sincos(phi);
        const v3 normal = sc.y * light_u + sc.x * light_v;
        const v3 lp = light_pos + normal * radius + (r2 - 0.5f) * light_dir * height;
        ls.lp = lp;
        float ls_dist;
        ls.L = normalize_len(lp - P, ls_dist);
        ls.area = l.p[3];
        ls.ray_flags = l_ray_visibility(l);
        const float cos_theta = 1.0f - fabsf(dot(ls.L, light_dir));
        if (cos_theta != 0.0f) {
            ls.pdf = (ls_dist * ls_dist) / (ls.area * cos_theta);
        }
        if (!l_visible(l)) {
            ls.area = 0.0f;
        }
    } else if (type == LIGHT_TRI) {
        const uint32_t ltri_index = __float_as_uint(l.p[0]);
        const MeshInstance &lmi = sg.instances[__float_as_uint(l.p[1])];
        const Vertex &v1 = ss.vertices[ss.vtx_indices[ltri_index * 3 + 0]],
                     &v2_ = ss.vertices[ss.vtx_indices[ltri_index * 3 + 1]],
                     &v3_ = ss.vertices[ss.vtx_indices[ltri_index * 3 + 2]];
        const v3 p1 = transform_point(mk3(v1.p), lmi.xform), p2 = transform_point(mk3(v2_.p), lmi.xform),
                 p3 = transform_point(mk3(v3_.p), lmi.xform);
        const v3 e1 = p2 - p1, e2 = p3 - p1;
        float light_fwd_len;
        const v3 light_forward = normalize_len(cross(e1, e2), light_fwd_len);
        ls.area = 0.5f * light_fwd_len;
        ls.ray_flags = l_ray_visibility(l);
        v3 lp;
        v2 luvs;
        float pdf = sample_spherical_triangle(P, p1, p2, p3, rand_light_uv, &ls.L);
        if (pdf > 0.0f) {
            const v3 pvec = cross(ls.L, e2);
            const v3 tvec = P - p1, qvec = cross(tvec, e1);
            const float inv_det = 1.0f / dot(e1, pvec);
            const float tri_u = dot(tvec, pvec) * inv_det, tri_v = dot(ls.L, qvec) * inv_det;
            lp = (1.0f - tri_u - tri_v) * p1 + tri_u * p2 + tri_v * p3;
            const float w0 = 1.0f - tri_u - tri_v;
            luvs = v2{w0 * v1.t[0] + tri_u * v2_.t[0] + tri_v * v3_.t[0], w0 * v1.t[1] + tri_u * v2_.t[1] + tri_v * v3_.t[1]};
        } else {
            const float r1 = sqrtf(rand_light_uv.x), r2 = rand_light_uv.y;
            luvs = v2{v1.t[0] * (1.0f - r1) + r1 * (v2_.t[0] * (1.0f - r2) + v3_.t[0] * r2),
                      v1.t[1] * (1.0f - r1) + r1 * (v2_.t[1] * (1.0f - r2) + v3_.t[1] * r2)};
            lp = p1 * (1.0f - r1) + r1 * (p2 * (1.0f - r2) + p3 * r2);
            float ls_dist;
            ls.L = normalize_len(lp - P, ls_dist);
            const float cos_theta = -dot(ls.L, light_forward);
            pdf = safe_div_pos(ls_dist * ls_dist, ls.area * cos_theta);
        }
        float cos_theta = -dot(ls.L, light_forward);
        ls.lp = offset_ray(lp, cos_theta >= 0.0f ? light_forward : -light_forward);
        if (l_doublesided(l)) {
            cos_theta = fabsf(cos_theta);
        }
        if (cos_theta > 0.0f) {
            ls.pdf = pdf;
            const uint32_t tex_index = __float_as_uint(l.p[2]); // light_t::tri.tex_index
            if (tex_on && tex_index != kTexInvalid) {
                const c4 tex_color = tex_sample_color(tx, tex_index, luvs, 0, rand_tex_uv, true);
                ls.col.x *= tex_color.x;
                ls.col.y *= tex_color.y;
                ls.col.z *= tex_color.z;
            }
        }
    } else if (type == LIGHT_ENV) {
        const float rx = rand_light_uv.x, ry = rand_light_uv.y;
        float env_pdf;
        if (tex_on && sl.env.qtree_levels != 0) {
            // importance-sample the environment map through its quad-tree (CoreRef.cpp:3579-3584)
            ls.L = sample_env_qtree(sl.env, sl.env.env_map_rotation, u1, rx, ry, &env_pdf);
        } else {
            // no quad-tree (no env map, SceneCPU.cpp:905-908): sample the hemisphere around N
            const float phi = 2 * kPi * ry;
            const v2 sc = portable_sincos(phi);
            const float cos_phi = sc.y, sin_phi = sc.x;
            const float dir = sqrtf(1.0f - rx * rx);
            const v3 V = v3{dir * cos_phi, dir * sin_phi, rx};
            ls.L = world_from_tangent(T, B, N, V);
            env_pdf = 0.5f / kPi;
        }
        ls.col *= mk3(sl.env_col);
        if (tex_on && sl.env.env_map != kTexInvalid) {
            ls.col *= sample_latlong_rgbe(tx, sl.env.env_map, ls.L, sl.env.env_map_rotation, rand_tex_uv);
        }
        ls.area = 1.0f;
        ls.lp = P + ls.L;
        ls.dist_mul = kMaxDist;
        ls.pdf = env_pdf;
        ls.from_env = true;
        ls.ray_flags = l_ray_visibility(l);
    }
    ls.pdf /= factor;
}

// IntersectAreaLights for one ray (secondary rays): analytic lights through the light tree, tracking the pdf factor.
RT_FN void intersect_area_lights(const SceneLights &sl, v3 ro, v3 rd, uint32_t ray_flags, Hit &inter,
                                  LightStackEntry *st) {
    const v3 inv_d = safe_invert(rd);
    int sp = 0;
    st[sp++] = LightStackEntry{0u, 0.0f, 1.0f};
    while (sp) {
        LightStackEntry cur = st[--sp];
        if (cur.dist > inter.t || cur.factor == 0.0f) {
            continue;
        }
        while (true) { // TRAVERSE
            if (cur.index == kEmptyChild) {
                break; // empty slot (degenerate point box, zero flux): the reference would index out of bounds here
            }
            if ((cur.index & kLeafBit) == 0) {
                const LightCWNode &n = sl.nodes[cur.index];
                float bmin[24], bmax[24], dist[8];
                unpack_cw_bounds(n, bmin, bmax);
                uint32_t mask = box8(bmin, bmax, ro, inv_d, inter.t, dist);
                if (mask == 0) {
                    break;
                }
                float factors[8];
                lnode_importance(n, bmin, bmax, ro, factors);
                const float total_importance = sum_importance(factors);
                if (total_importance == 0.0f) {
                    break;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    factors[j] /= total_importance;
                }
                int i = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    cur.index = n.child[i];
                    cur.factor *= factors[i];
                    continue;
                }
                const int i2 = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    if (dist[i] < dist[i2]) {
                        st[sp++] = LightStackEntry{n.child[i2], dist[i2], cur.factor * factors[i2]};
                        cur.index = n.child[i];
                        cur.factor *= factors[i];
                    } else {
                        st[sp++] = LightStackEntry{n.child[i], dist[i], cur.factor * factors[i]};
                        cur.index = n.child[i2];
                        cur.factor *= factors[i2];
                    }
                    continue;
                }
                st[sp++] = LightStackEntry{n.child[i], dist[i], cur.factor * factors[i]};
                st[sp++] = LightStackEntry{n.child[i2], dist[i2], cur.factor * factors[i2]};
                i = __ffs(mask) - 1;
                mask &= mask - 1;
                st[sp++] = LightStackEntry{n.child[i], dist[i], cur.factor * factors[i]};
                if (mask == 0) {
                    sort_top3(st, sp);
                    cur = st[--sp];
                    continue;
                }
                i = __ffs(mask) - 1;
                mask &= mask - 1;
                st[sp++] = LightStackEntry{n.child[i], dist[i], cur.factor * factors[i]};
                if (mask == 0) {
                    sort_top4(st, sp);
                    cur = st[--sp];
                    continue;
                }
                const int size_before = sp;
                do {
                    i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    st[sp++] = LightStackEntry{n.child[i], dist[i], cur.factor * factors[i]};
                } while (mask != 0);
                sort_topN(st, sp, sp - size_before + 4);
                cur = st[--sp];
                continue;
            }
            // leaf = one light
            const int light_index = int(cur.index & kPrimIndexBits);
            const Light &l = sl.lights[light_index];
            if (!l_visible(l) || (l_ray_visibility(l) & ray_flags) == 0) {
                break;
            }
            if (l_sky_portal(l) && inter.v >= 0.0f) {
                break;
            }
            const bool no_shadow = !l_cast_shadow(l);
            const int type = l_type(l);
            if (type == LIGHT_SPHERE) {
                const v3 light_pos = mk3(&l.p[0]);
                const float radius = l.p[7];
                const v3 op = light_pos - ro;
                const float b = dot(op, rd);
                float det = b * b - dot(op, op) + radius * radius;
                if (det >= 0.0f) {
                    det = sqrtf(det);
                    const float t1 = b - det, t2 = b + det;
                    if (t1 > kHitEps && (t1 < inter.t || no_shadow)) {
                        bool accept = true;
                        const float spot = l.p[8];
                        if (spot > 0.0f) {
                            const float _dot = -dot(rd, mk3(&l.p[4]));
                            if (_dot > 0.0f) {
                                const float _angle = libm_acosf(saturatef(_dot));
                                accept &= (_angle <= spot);
                            } else {
                                accept = false;
                            }
                        }
                        if (accept) {
                            inter.v = 0.0f;
                            inter.obj = -light_index - 1;
                            inter.t = t1;
                            inter.u = cur.factor;
                        }
                    } else if (t2 > kHitEps && (t2 < inter.t || no_shadow)) {
                        inter.v = 0.0f;
                        inter.obj = -light_index - 1;
                        inter.t = t2;
                        inter.u = cur.factor;
                    }
                }
            } else if (type == LIGHT_DIR) {
                const v3 light_dir = mk3(&l.p[0]);
                const float cos_theta = dot(rd, light_dir);
                if ((inter.v < 0.0f || no_shadow) && cos_theta > l.p[3]) {
                    inter.v = 0.0f;
                    inter.obj = -light_index - 1;
                    inter.t = 1.0f / cos_theta;
                    inter.u = cur.factor;
                }
            } else if (type == LIGHT_RECT) {
                const v3 light_pos = mk3(&l.p[0]);
                v3 light_u = mk3(&l.p[4]), light_v = mk3(&l.p[8]);
                const v3 light_forward = normalize(cross(light_u, light_v));
                const float plane_dist = dot(light_forward, light_pos);
                const float cos_theta = dot(rd, light_forward);
                const float t = (plane_dist - dot(light_forward, ro)) / fminf(cos_theta, -kFltEps);
                if (cos_theta < 0.0f && t > kHitEps && (t < inter.t || no_shadow)) {
                    light_u /= dot(light_u, light_u);
                    light_v /= dot(light_v, light_v);
                    const v3 p = ro + rd * t;
                    const v3 vi = p - light_pos;
                    const float a1 = dot(light_u, vi);
                    if (a1 >= -0.5f && a1 <= 0.5f) {
                        const float a2 = dot(light_v, vi);
                        if (a2 >= -0.5f && a2 <= 0.5f) {
                            inter.v = 0.0f;
                            inter.obj = -light_index - 1;
                            inter.t = t;
                            inter.u = cur.factor;
                        }
                    }
                }
            } else if (type == LIGHT_DISK) {
                const v3 light_pos = mk3(&l.p[0]);
                v3 light_u = mk3(&l.p[4]), light_v = mk3(&l.p[8]);
                const v3 light_forward = normalize(cross(light_u, light_v));
                const float plane_dist = dot(light_forward, light_pos);
                const float cos_theta = dot(rd, light_forward);
                const float t = safe_div_neg(plane_dist - dot(light_forward, ro), cos_theta);
                if (cos_theta < 0.0f && t > kHitEps && (t < inter.t || no_shadow)) {
                    light_u /= dot(light_u, light_u);
                    light_v /= dot(light_v, light_v);
                    const v3 p = ro + rd * t;
                    const v3 vi = p - light_pos;
                    const float a1 = dot(light_u, vi);
                    const float a2 = dot(light_v, vi);
                    if (sqrtf(a1 * a1 + a2 * a2) <= 0.5f) {
                        inter.v = 0.0f;
                        inter.obj = -light_index - 1;
                        inter.t = t;
                        inter.u = cur.factor;
                    }
                }
            } else if (type == LIGHT_LINE) {
                const v3 light_pos = mk3(&l.p[0]);
                const v3 light_u = mk3(&l.p[4]), light_dir = mk3(&l.p[8]);
                const float radius = l.p[7], height = l.p[11];
                const v3 light_v = cross(light_u, light_dir);
                v3 _ro = ro - light_pos;
                _ro = v3{dot(_ro, light_dir), dot(_ro, light_u), dot(_ro, light_v)};
                const v3 _rd = v3{dot(rd, light_dir), dot(rd, light_u), dot(rd, light_v)};
                const float A = _rd.z * _rd.z + _rd.y * _rd.y;
                const float Bq = 2.0f * (_rd.z * _ro.z + _rd.y * _ro.y);
                const float C = sqr(_ro.z) + sqr(_ro.y) - sqr(radius);
                float t0, t1;
                if (quadratic(A, Bq, C, t0, t1) && t0 > kHitEps && t1 > kHitEps) {
                    const float t = fminf(t0, t1);
                    const v3 p = _ro + t * _rd;
                    if (fabsf(p.x) < 0.5f * height && (t < inter.t || no_shadow)) {
                        inter.v = 0.0f;
                        inter.obj = -light_index - 1;
                        inter.t = t;
                        inter.u = cur.factor;
                    }
                }
            } else if (type == LIGHT_ENV && inter.v < 0.0f) {
                inter.obj = -light_index - 1;
                inter.u = cur.factor;
            }
            break;
        }
    }
}

// Blocker lights for shadow rays: returns 0 when a rect/disk light blocks the ray, 1 otherwise.
RT_FN float intersect_area_lights_shadow(const SceneLights &sl, v3 ro, v3 rd, float ray_dist, StackEntry *st) {
    const float rdist = fabsf(ray_dist);
    const v3 inv_d = safe_invert(rd);
    int sp = 0;
    st[sp++] = StackEntry{0u, 0.0f};
    while (sp) {
        StackEntry cur = st[--sp];
        if (cur.dist > rdist) {
            continue;
        }
        while (true) {
            if (cur.index == kEmptyChild) {
                break;
            }
            if ((cur.index & kLeafBit) == 0) {
                const LightCWNode &n = sl.nodes[cur.index];
                float bmin[24], bmax[24], dist[8];
                unpack_cw_bounds(n, bmin, bmax);
                uint32_t mask = box8(bmin, bmax, ro, inv_d, rdist, dist);
                if (mask == 0) {
                    break;
                }
                int i = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    cur.index = n.child[i];
                    continue;
                }
                const int i2 = __ffs(mask) - 1;
                mask &= mask - 1;
                if (mask == 0) {
                    if (dist[i] < dist[i2]) {
                        st[sp++] = StackEntry{n.child[i2], dist[i2]};
                        cur.index = n.child[i];
                    } else {
                        st[sp++] = StackEntry{n.child[i], dist[i]};
                        cur.index = n.child[i2];
                    }
                    continue;
                }
                st[sp++] = StackEntry{n.child[i], dist[i]};
                st[sp++] = StackEntry{n.child[i2], dist[i2]};
                i = __ffs(mask) - 1;
                mask &= mask - 1;
                st[sp++] = StackEntry{n.child[i], dist[i]};
                if (mask == 0) {
                    sort_top3(st, sp);
                    cur.index = st[--sp].index;
                    continue;
                }
                i = __ffs(mask) - 1;
                mask &= mask - 1;
                st[sp++] = StackEntry{n.child[i], dist[i]};
                if (mask == 0) {
                    sort_top4(st, sp);
                    cur.index = st[--sp].index;
                    continue;
                }
                const int size_before = sp;
                do {
                    i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    st[sp++] = StackEntry{n.child[i], dist[i]};
                } while (mask != 0);
                sort_topN(st, sp, sp - size_before + 4);
                cur.index = st[--sp].index;
                continue;
            }
            const int light_index = int(cur.index & kPrimIndexBits);
            const Light &l = sl.lights[light_index];
            if ((l_ray_visibility(l) & (1u << RAY_SHADOW)) == 0) {
                break;
            }
            if (l_sky_portal(l) && ray_dist >= 0.0f) {
                break;
            }
            const int type = l_type(l);
            if (type == LIGHT_RECT || type == LIGHT_DISK) {
                const v3 light_pos = mk3(&l.p[0]);
                v3 light_u = mk3(&l.p[4]), light_v = mk3(&l.p[8]);
                const v3 light_forward = normalize(cross(light_u, light_v));
                const float plane_dist = dot(light_forward, light_pos);
                const float cos_theta = dot(rd, light_forward);
                const float t = (type == LIGHT_RECT)
                                    ? (plane_dist - dot(light_forward, ro)) / fminf(cos_theta, -kFltEps)
                                    : safe_div_neg(plane_dist - dot(light_forward, ro), cos_theta);
                if (cos_theta < 0.0f && t > kHitEps && t < rdist) {
                    light_u /= dot(light_u, light_u);
                    light_v /= dot(light_v, light_v);
                    const v3 p = ro + rd * t;
                    const v3 vi = p - light_pos;
                    const float a1 = dot(light_u, vi);
                    if (type == LIGHT_RECT) {
                        if (a1 >= -0.5f && a1 <= 0.5f) {
                            const float a2 = dot(light_v, vi);
                            if (a2 >= -0.5f && a2 <= 0.5f) {
                                return 0.0f;
                            }
                        }
                    } else {
                        const float a2 = dot(light_v, vi);
                        if (sqrtf(a1 * a1 + a2 * a2) <= 0.5f) {
                            return 0.0f;
                        }
                    }
                }
            }
            break;
        }
    }
    return 1.0f;
}

// point-in-box mask for the quantised node (bbox_test_oct(p, cwbvh), CoreRef.cpp:243-281)
RT_DEV uint32_t point_in_box8(const float bmin[24], const float bmax[24], v3 p) {
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if ((bmin[0 * 8 + i] <= p.x) & (bmin[1 * 8 + i] <= p.y) & (bmin[2 * 8 + i] <= p.z) & (bmax[0 * 8 + i] >= p.x) &
            (bmax[1 * 8 + i] >= p.y) & (bmax[2 * 8 + i] >= p.z)) {
            mask |= (1u << i);
        }
    }
    return mask;
}

// EvalTriLightFactor: probability factor with which NEE would have picked emissive triangle `tri_index` from `ro`.
RT_FN float eval_tri_light_factor(const SceneLights &sl, v3 P, v3 ro, uint32_t tri_index, uint32_t *stack,
                                   float *stack_factors) {
    int sp = 0;
    stack_factors[sp] = 1.0f;
    stack[sp++] = 0;
    while (sp) {
        const uint32_t cur = stack[--sp];
        const float cur_factor = stack_factors[sp];
        if ((cur & kLeafBit) == 0) {
            const LightCWNode &n = sl.nodes[cur];
            float bmin[24], bmax[24];
            unpack_cw_bounds(n, bmin, bmax);
            uint32_t mask = point_in_box8(bmin, bmax, P);
            if (mask) {
                float importance[8];
                lnode_importance(n, bmin, bmax, ro, importance);
                const float total_importance = sum_importance(importance);
                if (total_importance == 0.0f) {
                    continue;
                }
                do {
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (importance[i] > 0.0f) {
                        stack_factors[sp] = cur_factor * importance[i] / total_importance;
                        stack[sp++] = n.child[i];
                    }
                } while (mask != 0);
            }
        } else {
            const int light_index = int(cur & kPrimIndexBits);
            const Light &l = sl.lights[light_index];
            if (l_type(l) == LIGHT_TRI && __float_as_uint(l.p[0]) == tri_index) {
                return 1.0f / cur_factor;
            }
        }
    }
    return 1.0f;
}

} // namespace rt
