// rt_types.h -- plain-old-data layouts shared by host code and sm_100a kernels.
//
// Scene arrays cross the C-ABI (include/ray_cuda.h, rc_upload_scene) in the byte layouts the reference keeps in
// Cpu::Scene (reference internal/Core.h); the device keeps them as they arrive, so a maintainer wiring this backend
// into the reference hands over SparseStorage::data() pointers with no conversion.  Each struct below names the
// reference struct whose layout it must match and static_asserts the size the reference asserts.
#pragma once

#include <stdint.h>

namespace rt {

// ---- constants (reference internal/Constants.inl) -----------------------------------------------------------------
constexpr int kMaxStack = 48;                  // MAX_STACK_SIZE, Constants.inl:4
constexpr float kHitBias = 0.00001f;           // HIT_BIAS
constexpr float kHitEps = 0.000001f;           // HIT_EPS
constexpr float kFltEps = 0.0000001f;          // FLT_EPS
constexpr float kMaxDist = 3.402823466e+30F;   // MAX_DIST (deliberately not FLT_MAX)
constexpr float kFltMax = 3.402823466e+38F;    // FLT_MAX
constexpr float kFltMin = 1.175494351e-38F;    // FLT_MIN
constexpr float kSphericalAreaThreshold = 0.00005f;
constexpr float kPi = 3.141592653589793238463f;
constexpr uint32_t kLeafBit = 1u << 31;        // LEAF_NODE_BIT
constexpr uint32_t kPrimIndexBits = ~kLeafBit; // PRIM_INDEX_BITS
constexpr uint32_t kEmptyChild = 0x7fffffffu;  // empty BVH8 slot (Core.cpp:849-853)

constexpr int kRandDimFilter = 0, kRandDimLens = 1, kRandDimBase = 2;
constexpr int kRandDimBsdfPick = 0, kRandDimBsdf = 1, kRandDimLightPick = 2, kRandDimLight = 3, kRandDimTex = 4;
constexpr int kRandDimBounce = 8;
constexpr int kRandSamples = 4096, kRandDims = 32; // __pmj02_sample_count / __pmj02_dims_count

enum : int { LIGHT_SPHERE = 0, LIGHT_DIR, LIGHT_LINE, LIGHT_RECT, LIGHT_DISK, LIGHT_TRI, LIGHT_ENV };
enum : int { RAY_CAMERA = 0, RAY_DIFFUSE, RAY_SPECULAR, RAY_REFR, RAY_SHADOW };
enum : uint32_t { NODE_DIFFUSE = 0, NODE_GLOSSY, NODE_REFRACTIVE, NODE_EMISSIVE, NODE_MIX, NODE_TRANSPARENT,
                  NODE_PRINCIPLED };

constexpr int kTexNormals = 0, kTexBase = 1, kTexRough = 2, kTexMetallic = 3, kTexSpecular = 4;
constexpr int kMixMat1 = 3, kMixMat2 = 4;
constexpr uint32_t kTexSrgbBitHost = 1u << 24, kTexReconstructZBitHost = 2u << 24; // TEX_*_BIT, Core.h:159-160
constexpr uint32_t kMatSolidBit = 32768, kMatIndexBits = 16383;
constexpr uint32_t kMatFlagImpSample = 1u, kMatFlagMixAdd = 2u;
constexpr float kMaxConeSpreadInc = 0.05f;
constexpr int kFilterTableSize = 1024;

// ---- scene PODs ----------------------------------------------------------------------------------------------------
struct alignas(32) MTri { // mtri_accel_t, Core.h:79-84: 8 triangles, SoA
    float n_plane[4][8];
    float u_plane[4][8];
    float v_plane[4][8];
};
static_assert(sizeof(MTri) == 384, "mtri_accel_t");

struct alignas(32) WNode { // wbvh_node_t, Core.h:118-123
    float bbox_min[3][8];
    float bbox_max[3][8];
    uint32_t child[8];
};
static_assert(sizeof(WNode) == 224, "wbvh_node_t");

struct alignas(16) LightCWNode { // light_cwbvh_node_t, Core.h:132-148
    float bbox_min[3];
    float _unused0;
    float bbox_max[3];
    float _unused1;
    uint8_t ch_bbox_min[3][8];
    uint8_t ch_bbox_max[3][8];
    uint32_t child[8];
    float flux[8];
    uint32_t axis[8];
    uint32_t cos_omega_ne[8];
};
static_assert(sizeof(LightCWNode) == 208, "light_cwbvh_node_t");

struct TriMat { // tri_mat_data_t, Core.h:163-165
    uint16_t front_mi, back_mi;
};
static_assert(sizeof(TriMat) == 4, "tri_mat_data_t");

struct Material { // material_t, Core.h:167-192
    uint32_t textures[5];
    float base_color[3];
    uint32_t flags;
    uint32_t type;
    float tangent_rotation_or_strength;
    uint16_t roughness_unorm;
    uint16_t anisotropic_unorm;
    float ior;
    uint16_t sheen_unorm;
    uint16_t sheen_tint_unorm;
    uint16_t tint_unorm;
    uint16_t metallic_unorm;
    uint16_t transmission_unorm;
    uint16_t transmission_roughness_unorm;
    uint16_t specular_unorm;
    uint16_t specular_tint_unorm;
    uint16_t clearcoat_unorm;
    uint16_t clearcoat_roughness_unorm;
    uint16_t normal_map_strength_unorm;
    uint16_t _pad;
};
static_assert(sizeof(Material) == 76, "material_t");

struct Light { // light_t, Core.h:194-237. First word: type:3 doublesided:1 cast_shadow:1 visible:1 sky_portal:1
               // ray_visibility:8 (gcc LSB-first bit-field order).
    uint32_t bits;
    float col[3];
    float p[12]; // union payload, see accessors below
};
static_assert(sizeof(Light) == 64, "light_t");

struct Vertex { // vertex_t, Core.h:370-373
    float p[3], n[3], b[3], t[2];
};
static_assert(sizeof(Vertex) == 44, "vertex_t");

struct MeshInstance { // mesh_instance_t, Core.h:384-391
    uint32_t mesh_index;
    uint32_t node_index;
    uint32_t lights_index;
    uint32_t ray_visibility; // upper 24 bits: lights block
    float xform[16], inv_xform[16];
};
static_assert(sizeof(MeshInstance) == 144, "mesh_instance_t");

// ---- stream records in the reference's own (AoS) layouts: used at the C-ABI stage/debug boundary -------------------
struct RayAoS { // Ref::ray_data_t, CoreRef.h:57-71
    float o[3], d[3], pdf;
    float c[3];
    float ior[4];
    float cone_width, cone_spread;
    uint32_t xy;
    uint32_t depth;
};
static_assert(sizeof(RayAoS) == 72, "ray_data_t");

struct ShadowRayAoS { // Ref::shadow_ray_t, CoreRef.h:74-86
    float o[3];
    uint32_t depth;
    float d[3], dist;
    float c[3];
    uint32_t xy;
};
static_assert(sizeof(ShadowRayAoS) == 48, "shadow_ray_t");

struct HitAoS { // Ref::hit_data_t, CoreRef.h:89-105
    int obj_index;
    int prim_index;
    float t, u, v;
};
static_assert(sizeof(HitAoS) == 20, "hit_data_t");

} // namespace rt
