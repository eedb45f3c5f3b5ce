// rt_tex.cuh -- device texture lookups.
//
// Behavioural spec: reference internal/CoreRef.cpp
//   get_texture_lod(lambda) :2838-2850     SampleBilinear (USE_STOCH_TEXTURE_FILTERING = 1, CoreSIMD.h:31) :2859-2892
//   TexStorageSwizzled::Fetch(index, int x, int y, lod)  internal/TextureStorageCPU.h:259-307
//   srgb_to_linear / YCoCg_to_RGB  internal/CoreRef.h:208-251
//
// Texels live DECODED on the device: one RGBA8 word per texel, row-major per mip level, with the channel expansion of
// TexStorage*::Fetch already applied (channels the storage lacks repeat the last stored one), so a lookup is one 32-bit
// load whatever storage (RGBA/RGB/RG/R/BCn) the reference keeps the texture in.  Handles in the device copies of
// material_t / light_t are rewritten at upload to  flags(bits 24..27, Core.h:159-161) | dense texture id (bits 0..23).
//
// srgb_to_linear is powf(., 2.4f) of the HOST libm in the reference; a texel channel has 256 possible values, so the
// host builds the 256-entry table with that very powf and the device indexes it by the stored byte: bit-identical.
#pragma once

#include "rt_math.cuh"

namespace rt {

constexpr uint32_t kTexInvalid = 0xffffffffu;
constexpr uint32_t kTexSrgbBit = 1u << 24, kTexReconstructZBit = 2u << 24, kTexYCoCgBit = 4u << 24;
constexpr uint32_t kTexIdBits = 0x00ffffffu;
constexpr int kTexMipLevels = 12, kMaxMipLevel = 11;

struct TexDesc {
    uint32_t offset[kTexMipLevels]; // in texels, into SceneTex::texels
    uint16_t w[kTexMipLevels], h[kTexMipLevels];
};

struct SceneTex {
    const TexDesc *__restrict__ descs;
    const uint32_t *__restrict__ texels; // RGBA8, r in the low byte
    const float *__restrict__ srgb_lut;  // 256: srgb_to_linear(i / 255.0f)
};

// TexStorage*::Fetch(index, int x, int y, lod) -> the stored bytes
RT_DEV uint32_t tex_fetch(const SceneTex &t, uint32_t id, int x, int y, int lod) {
    const TexDesc &d = t.descs[id];
    const int w = d.w[lod], h = d.h[lod];
    x %= w;
    y %= h;
    return __ldg(&t.texels[d.offset[lod] + uint32_t(y * w + x)]);
}

// SampleBilinear with stochastic filtering: ONE jittered nearest lookup.  Returns the stored bytes.
RT_DEV uint32_t tex_sample_bytes(const SceneTex &t, uint32_t handle, v2 uvs, int lod, v2 rand) {
    const uint32_t id = handle & kTexIdBits;
    const TexDesc &d = t.descs[id];
    const float sx = float(d.w[lod]), sy = float(d.h[lod]);
    float u = fractf(uvs.x), v = fractf(uvs.y);
    u = u * sx - 0.5f;
    v = v * sy - 0.5f;
    u += rand.x;
    v += rand.y;
    return tex_fetch(t, id, int(u), int(v), lod);
}

RT_DEV c4 tex_unpack(uint32_t px) { // color_rgba_t of Fetch(): byte / 255.0f
    return c4{float(px & 0xffu) / 255.0f, float((px >> 8) & 0xffu) / 255.0f, float((px >> 16) & 0xffu) / 255.0f,
              float(px >> 24) / 255.0f};
}

// YCoCg_to_RGB (CoreRef.h:239-251): the scaled YCoCg the reference stores base-colour maps in when texture compression is
// on (decoded BC3 texels cross the C-ABI: {Co, Cg, scale, Y})
RT_DEV c4 ycocg_to_rgb(c4 col) {
    const float scale = (col.z * (255.0f / 8.0f)) + 1.0f;
    const float Y = col.w;
    const float Co = (col.x - (0.5f * 256.0f / 255.0f)) / scale;
    const float Cg = (col.y - (0.5f * 256.0f / 255.0f)) / scale;
    c4 rgb;
    rgb.x = sse_max(0.0f, sse_min(Y + Co - Cg, 1.0f));
    rgb.y = sse_max(0.0f, sse_min(Y + Cg, 1.0f));
    rgb.z = sse_max(0.0f, sse_min(Y - Co - Cg, 1.0f));
    rgb.w = 1.0f;
    return rgb;
}

RT_DEV float srgb_to_linear_f(float c) { // CoreRef.h:208-220 on a float (after YCoCg the value is no longer a byte / 255)
    return (c > 0.04045f) ? libm_powf((c + 0.055f) / 1.055f, 2.4f) : (c / 12.92f);
}

// fvec4 colour of a texture sample after the handle's colour-space flags.  `ycocg`: the call site is one of those where
// the reference converts YCoCg-coded texels (base colour, texture-driven Mix, textured emissive triangles:
// ShadeRef.cpp:1308,1411, CoreRef.cpp:3108,3232,3567); the other sites (roughness, specular ...) read the raw channels.
RT_DEV c4 tex_sample_color(const SceneTex &t, uint32_t handle, v2 uvs, int lod, v2 rand, bool ycocg = false) {
    const uint32_t px = tex_sample_bytes(t, handle, uvs, lod, rand);
    if (ycocg && (handle & kTexYCoCgBit)) {
        c4 col = ycocg_to_rgb(tex_unpack(px));
        if (handle & kTexSrgbBit) {
            col.x = srgb_to_linear_f(col.x);
            col.y = srgb_to_linear_f(col.y);
            col.z = srgb_to_linear_f(col.z);
        }
        return col;
    }
    if (handle & kTexSrgbBit) {
        return c4{__ldg(&t.srgb_lut[px & 0xffu]), __ldg(&t.srgb_lut[(px >> 8) & 0xffu]),
                  __ldg(&t.srgb_lut[(px >> 16) & 0xffu]), float(px >> 24) / 255.0f};
    }
    return tex_unpack(px);
}

// get_texture_lod(textures, index, lambda)
RT_DEV int tex_lod(const SceneTex &t, uint32_t handle, float lambda) {
    const TexDesc &d = t.descs[handle & kTexIdBits];
    float lod = lambda + 0.5f * fast_log2(float(d.w[0]) * float(d.h[0]));
    lod = clampf(lod - 1.0f, 0.0f, float(kMaxMipLevel));
    return int(lod);
}

} // namespace rt
