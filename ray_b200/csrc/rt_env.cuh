// rt_env.cuh -- environment map lighting (without the procedural sky).
//
// Behavioural spec: reference
//   SampleLatlong_RGBE (stochastic filtering)     internal/CoreRef.cpp:2995-3021      rgbe_to_rgb  CoreRef.h:234-237
//   CanonicalToDir / DirToCanonical               internal/Core.cpp:110-143           to_norm_float Core.h:410-417
//   Evaluate_EnvQTree / Sample_EnvQTree           internal/CoreRef.cpp:4738-4839
// The env map is an ordinary RGBA8 texture of the scene's texel pool holding RGBE; the quad-tree levels built by
// Cpu::Scene::PrepareEnvMapQTree_nolock (SceneCPU.cpp:1058-1211) arrive through rc_scene_view and sit concatenated in
// one float4 array.  acosf / atan2f / sinf / cosf are the restated host libm functions of rt_math.cuh.
#pragma once

#include "rt_tex.cuh"

namespace rt {

constexpr int kMaxQTreeLevels = 16;

struct SceneEnv {
    uint32_t env_map, back_map; // kTexInvalid or dense texture id (flag bits stripped)
    float env_map_rotation, back_map_rotation;
    int qtree_levels;
    const float4 *__restrict__ qtree;      // all levels, level i at qtree_offset[i]
    uint32_t qtree_offset[kMaxQTreeLevels]; // in float4 units
};

RT_DEV float to_norm_float(uint32_t v) { // Core.h:410-417
    const uint32_t val = 0x3f800000u + v * 0x8080u + (v + 1u) / 2u;
    return __uint_as_float(val) - 1.0f;
}

RT_FN v3 sample_latlong_rgbe(const SceneTex &t, uint32_t tex, v3 dir, float y_rotation, v2 rand) {
    const float theta = libm_acosf(clampf(dir.y, -1.0f, 1.0f)) / kPi;
    float phi = libm_atan2f(dir.z, dir.x) + y_rotation;
    if (phi < 0) {
        phi += 2 * kPi;
    }
    if (phi > 2 * kPi) {
        phi -= 2 * kPi;
    }
    const float u = fractf(0.5f * phi / kPi);
    const uint32_t id = tex & kTexIdBits;
    const TexDesc &d = t.descs[id];
    float ux = u * float(d.w[0]), uy = theta * float(d.h[0]);
    ux += rand.x;
    uy += rand.y;
    const uint32_t px = tex_fetch(t, id, int(ux), int(uy), 0);
    const float f = rgbe_scale(px >> 24);
    return v3{to_norm_float(px & 0xffu) * f, to_norm_float((px >> 8) & 0xffu) * f, to_norm_float((px >> 16) & 0xffu) * f};
}

RT_FN v3 canonical_to_dir(v2 p, float y_rotation) {
    const float cos_theta = 2 * p.x - 1;
    float phi = 2 * kPi * p.y + y_rotation;
    if (phi < 0) {
        phi += 2 * kPi;
    }
    if (phi > 2 * kPi) {
        phi -= 2 * kPi;
    }
    const float sin_theta = sqrtf(1 - cos_theta * cos_theta);
    const float sin_phi = libm_sinf(phi);
    const float cos_phi = libm_cosf(phi);
    return v3{sin_theta * cos_phi, cos_theta, -sin_theta * sin_phi};
}

RT_FN v2 dir_to_canonical(v3 d, float y_rotation) {
    const float cos_theta = fminf(fmaxf(d.y, -1.0f), 1.0f);
    float phi = -libm_atan2f(d.z, d.x) + y_rotation;
    if (phi < 0) {
        phi += 2 * kPi;
    }
    if (phi > 2 * kPi) {
        phi -= 2 * kPi;
    }
    return v2{(cos_theta + 1.0f) / 2.0f, phi / (2.0f * kPi)};
}

RT_DEV float quad_lane(float4 q, int i) { return i == 0 ? q.x : (i == 1 ? q.y : (i == 2 ? q.z : q.w)); }

RT_FN float evaluate_env_qtree(const SceneEnv &e, float y_rotation, v3 L) {
    int res = 2;
    int lod = e.qtree_levels - 1;
    const v2 p = dir_to_canonical(L, -y_rotation);
    float factor = 1.0f;
    while (lod >= 0) {
        const int x = min(max(int(p.x * float(res)), 0), res - 1);
        const int y = min(max(int(p.y * float(res)), 0), res - 1);
        const int index = (x & 1) | ((y & 1) << 1);
        const int qx = x / 2, qy = y / 2;
        const float4 quad = __ldg(&e.qtree[e.qtree_offset[lod] + uint32_t(qy * res / 2 + qx)]);
        const float total = quad.x + quad.y + quad.z + quad.w;
        if (total <= 0.0f) {
            break;
        }
        factor *= 4.0f * quad_lane(quad, index) / total;
        --lod;
        res *= 2;
    }
    return factor / (4.0f * kPi);
}

// returns the direction; pdf through *out_pdf
RT_FN v3 sample_env_qtree(const SceneEnv &e, float y_rotation, float rand, float rx, float ry, float *out_pdf) {
    int res = 2;
    float step = 1.0f / float(res);
    float sample = rand;
    int lod = e.qtree_levels - 1;
    v2 origin = v2{0.0f, 0.0f};
    float factor = 1.0f;
    while (lod >= 0) {
        const int qx = int(origin.x * float(res)) / 2;
        const int qy = int(origin.y * float(res)) / 2;
        const float4 quad = __ldg(&e.qtree[e.qtree_offset[lod] + uint32_t(qy * res / 2 + qx)]);
        const float top_left = quad.x;
        const float top_right = quad.y;
        float partial = top_left + quad.z;
        const float total = partial + top_right + quad.w;
        if (total <= 0.0f) {
            break;
        }
        float boundary = partial / total;
        int index = 0;
        if (sample < boundary) {
            sample /= boundary;
            boundary = top_left / partial;
        } else {
            partial = total - partial;
            origin.x = origin.x + step;
            sample = (sample - boundary) / (1.0f - boundary);
            boundary = top_right / partial;
            index |= (1 << 0);
        }
        if (sample < boundary) {
            sample /= boundary;
        } else {
            origin.y = origin.y + step;
            sample = (sample - boundary) / (1.0f - boundary);
            index |= (1 << 1);
        }
        factor *= 4.0f * quad_lane(quad, index) / total;
        --lod;
        res *= 2;
        step *= 0.5f;
    }
    origin.x += 2 * step * rx;
    origin.y += 2 * step * ry;
    *out_pdf = factor / (4.0f * kPi);
    return canonical_to_dir(origin, y_rotation);
}

} // namespace rt
