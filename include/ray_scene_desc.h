/* ray_scene_desc.h -- flat C descriptors for scene construction.
 *
 * These are plain-C (no spans, no string_views, no handles-as-structs) restatements of the descriptor structs a
 * caller of the reference fills in before calling SceneBase::Add*:
 *
 *   rs_shading_node_desc    <- Ray::shading_node_desc_t      (reference SceneBase.h:44-64)
 *   rs_principled_mat_desc  <- Ray::principled_mat_desc_t    (SceneBase.h:67-94)
 *   rs_mat_group_desc       <- Ray::mat_group_desc_t         (SceneBase.h:97-110)
 *   rs_mesh_desc            <- Ray::mesh_desc_t              (SceneBase.h:119-131)
 *   rs_mesh_instance_desc   <- Ray::mesh_instance_desc_t     (SceneBase.h:134-142)
 *   rs_*_light_desc         <- Ray::{directional,sphere,spot,rect,disk,line}_light_desc_t (SceneBase.h:200-268)
 *   rs_camera_desc          <- Ray::camera_desc_t            (SceneBase.h:271-311)
 *   rs_environment_desc     <- Ray::environment_desc_t       (SceneBase.h:347-357) without the procedural sky
 *
 * Field names, meaning and defaults are the reference's. Handles are the 32-bit `_index` of the reference's
 * {_index,_block} handle pairs; RS_INVALID (0xffffffff) is "no handle".
 *
 * The same descriptors drive BOTH the product's host layer (ray_host.h, rh_* functions) and the test oracle
 * (oracle/ref_harness.cpp, ro_* functions), so a parity test describes a scene once and feeds it to both.
 */
#ifndef RAY_SCENE_DESC_H
#define RAY_SCENE_DESC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_INVALID 0xffffffffu

/* eShadingNode (SceneBase.h:41) */
enum { RS_NODE_DIFFUSE = 0, RS_NODE_GLOSSY, RS_NODE_REFRACTIVE, RS_NODE_EMISSIVE, RS_NODE_MIX, RS_NODE_TRANSPARENT,
       RS_NODE_PRINCIPLED };

typedef struct rs_shading_node_desc {
    uint32_t type;
    float base_color[3];
    uint32_t base_texture;
    uint32_t normal_map;
    float normal_map_intensity;
    uint32_t mix_materials[2];
    float roughness;
    uint32_t roughness_texture;
    float anisotropic;
    float anisotropic_rotation;
    float sheen;
    float specular;
    float strength;
    float fresnel;
    float ior;
    float tint;
    uint32_t metallic_texture;
    uint32_t importance_sample; /* bool */
    uint32_t mix_add;           /* bool */
} rs_shading_node_desc;

typedef struct rs_principled_mat_desc {
    float base_color[3];
    uint32_t base_texture;
    float metallic;
    uint32_t metallic_texture;
    float specular;
    uint32_t specular_texture;
    float specular_tint;
    float roughness;
    uint32_t roughness_texture;
    float anisotropic;
    float anisotropic_rotation;
    float sheen;
    float sheen_tint;
    float clearcoat;
    float clearcoat_roughness;
    float ior;
    float transmission;
    float transmission_roughness;
    float emission_color[3];
    uint32_t emission_texture;
    float emission_strength;
    float alpha;
    uint32_t alpha_texture;
    uint32_t normal_map;
    float normal_map_intensity;
    uint32_t importance_sample; /* bool */
} rs_principled_mat_desc;

/* tex_desc_t (SceneBase.h:177-192), uncompressed formats.  The returned handle is the reference's TextureHandle::_index:
 * (storage << 28) | flag bits 24..27 | index (SceneCPU.cpp:191-204). */
enum { RS_TEX_RGBA8888 = 1, RS_TEX_RGB888 = 2, RS_TEX_RG88 = 3, RS_TEX_R8 = 4 }; /* eTextureFormat, SceneBase.h:145 */
typedef struct rs_tex_desc {
    uint32_t format;      /* RS_TEX_* */
    uint32_t convention;  /* eTextureConvention: 0 = OGL, 1 = DX (inverts y of normal maps) */
    const uint8_t *data;  /* w * h texels, row-major */
    int32_t w, h;
    uint32_t is_srgb;          /* bool, default 1 */
    uint32_t is_normalmap;     /* bool */
    uint32_t generate_mipmaps; /* bool (the uncompressed storages ignore it, TextureStorageCPU.cpp:232) */
    uint32_t reconstruct_z;    /* bool */
} rs_tex_desc;

typedef struct rs_mat_group_desc {
    uint32_t front_mat;
    uint32_t back_mat;
    uint64_t vtx_start; /* first INDEX of the group (the reference calls it vtx_start) */
    uint64_t vtx_count; /* number of indices in the group */
} rs_mat_group_desc;

/* One interleaved float array per attribute: {data, count_of_floats, offset, stride} as vtx_attribute_t. */
typedef struct rs_vtx_attribute {
    const float *data;
    uint64_t count; /* number of floats reachable through data */
    int32_t offset; /* in floats */
    int32_t stride; /* in floats */
} rs_vtx_attribute;

typedef struct rs_mesh_desc {
    rs_vtx_attribute vtx_positions; /* 3 floats */
    rs_vtx_attribute vtx_normals;   /* 3 floats */
    rs_vtx_attribute vtx_binormals; /* 3 floats, optional (data == NULL) */
    rs_vtx_attribute vtx_uvs;       /* 2 floats */
    const uint32_t *vtx_indices;
    uint64_t vtx_indices_count;
    int32_t base_vertex;
    const rs_mat_group_desc *groups;
    uint32_t groups_count;
    uint32_t allow_spatial_splits; /* bool */
    uint32_t use_fast_bvh_build;   /* bool */
} rs_mesh_desc;

typedef struct rs_mesh_instance_desc {
    float xform[16]; /* column-major 4x4 */
    uint32_t mesh;
    uint32_t camera_visibility, diffuse_visibility, specular_visibility, refraction_visibility, shadow_visibility;
} rs_mesh_instance_desc;

typedef struct rs_light_common {
    float color[3];
    uint32_t multiple_importance, cast_shadow, diffuse_visibility, specular_visibility, refraction_visibility;
} rs_light_common;

typedef struct rs_directional_light_desc {
    rs_light_common c;
    float direction[3], angle;
} rs_directional_light_desc;

typedef struct rs_sphere_light_desc {
    rs_light_common c;
    float position[3], radius;
} rs_sphere_light_desc;

typedef struct rs_spot_light_desc {
    rs_light_common c;
    float position[3], direction[3];
    float spot_size, spot_blend, radius;
} rs_spot_light_desc;

typedef struct rs_rect_light_desc {
    rs_light_common c;
    float width, height;
    uint32_t doublesided, sky_portal;
    float xform[16];
} rs_rect_light_desc;

typedef struct rs_disk_light_desc {
    rs_light_common c;
    float size_x, size_y;
    uint32_t doublesided, sky_portal;
    float xform[16];
} rs_disk_light_desc;

typedef struct rs_line_light_desc {
    rs_light_common c;
    float radius, height;
    uint32_t sky_portal;
    float xform[16];
} rs_line_light_desc;

/* eCamType / ePixelFilter / eViewTransform / eLensUnits (Types.h:62-84) */
enum { RS_CAM_PERSP = 0, RS_CAM_ORTHO, RS_CAM_GEO };
enum { RS_FILTER_BOX = 0, RS_FILTER_GAUSSIAN, RS_FILTER_BLACKMAN_HARRIS };
enum { RS_VIEW_STANDARD = 0 };
enum { RS_LENS_FOV = 0, RS_LENS_FLENGTH };

typedef struct rs_camera_desc {
    uint32_t type, filter, view_transform, ltype;
    float filter_width;
    float origin[3], fwd[3], up[3], shift[2];
    float exposure, fov, gamma, sensor_height, focus_distance, focal_length, fstop, lens_rotation, lens_ratio;
    int32_t lens_blades;
    float clip_start, clip_end;
    uint32_t mi_index, uv_index;
    uint32_t lighting_only, skip_direct_lighting, skip_indirect_lighting, no_background, output_sh;
    uint32_t max_diff_depth, max_spec_depth, max_refr_depth, max_transp_depth, max_total_depth;
    uint32_t min_total_depth, min_transp_depth;
    float clamp_direct, clamp_indirect;
    int32_t min_samples;
    float variance_threshold, regularize_alpha;
} rs_camera_desc;

typedef struct rs_environment_desc {
    float env_col[3];
    float back_col[3];
    uint32_t importance_sample; /* bool */
    uint32_t env_map;           /* RS_INVALID or the handle of an RGBA8888 texture holding RGBE (lat-long) */
    uint32_t back_map;          /* likewise, seen by camera rays */
    float env_map_rotation;     /* radians */
    float back_map_rotation;
} rs_environment_desc;

/* Fill a descriptor with the reference's defaults (SceneBase.h initialisers). */
static inline void rs_shading_node_defaults(rs_shading_node_desc *d) {
    d->type = RS_NODE_DIFFUSE;
    d->base_color[0] = d->base_color[1] = d->base_color[2] = 1.0f;
    d->base_texture = d->normal_map = RS_INVALID;
    d->normal_map_intensity = 1.0f;
    d->mix_materials[0] = d->mix_materials[1] = RS_INVALID;
    d->roughness = 0.0f;
    d->roughness_texture = RS_INVALID;
    d->anisotropic = d->anisotropic_rotation = d->sheen = d->specular = 0.0f;
    d->strength = 1.0f;
    d->fresnel = 1.0f;
    d->ior = 1.0f;
    d->tint = 0.0f;
    d->metallic_texture = RS_INVALID;
    d->importance_sample = 0;
    d->mix_add = 0;
}

static inline void rs_principled_defaults(rs_principled_mat_desc *d) {
    d->base_color[0] = d->base_color[1] = d->base_color[2] = 1.0f;
    d->base_texture = RS_INVALID;
    d->metallic = 0.0f;
    d->metallic_texture = RS_INVALID;
    d->specular = 0.5f;
    d->specular_texture = RS_INVALID;
    d->specular_tint = 0.0f;
    d->roughness = 0.5f;
    d->roughness_texture = RS_INVALID;
    d->anisotropic = d->anisotropic_rotation = 0.0f;
    d->sheen = 0.0f;
    d->sheen_tint = 0.5f;
    d->clearcoat = d->clearcoat_roughness = 0.0f;
    d->ior = 1.45f;
    d->transmission = d->transmission_roughness = 0.0f;
    d->emission_color[0] = d->emission_color[1] = d->emission_color[2] = 0.0f;
    d->emission_texture = RS_INVALID;
    d->emission_strength = 1.0f;
    d->alpha = 1.0f;
    d->alpha_texture = RS_INVALID;
    d->normal_map = RS_INVALID;
    d->normal_map_intensity = 1.0f;
    d->importance_sample = 0;
}

static inline void rs_light_common_defaults(rs_light_common *c) {
    c->color[0] = c->color[1] = c->color[2] = 1.0f;
    c->multiple_importance = c->cast_shadow = 1;
    c->diffuse_visibility = c->specular_visibility = c->refraction_visibility = 1;
}

static inline void rs_camera_defaults(rs_camera_desc *c) {
    c->type = RS_CAM_PERSP;
    c->filter = RS_FILTER_BLACKMAN_HARRIS;
    c->view_transform = RS_VIEW_STANDARD;
    c->ltype = RS_LENS_FOV;
    c->filter_width = 1.5f;
    for (int i = 0; i < 3; ++i) c->origin[i] = c->fwd[i] = c->up[i] = 0.0f;
    c->shift[0] = c->shift[1] = 0.0f;
    c->exposure = 0.0f;
    c->fov = 45.0f;
    c->gamma = 1.0f;
    c->sensor_height = 0.036f;
    c->focus_distance = 1.0f;
    c->focal_length = 0.0f;
    c->fstop = 0.0f;
    c->lens_rotation = 0.0f;
    c->lens_ratio = 1.0f;
    c->lens_blades = 0;
    c->clip_start = 0.0f;
    c->clip_end = 3.402823466e+30F;
    c->mi_index = 0xffffffffu;
    c->uv_index = 0;
    c->lighting_only = c->skip_direct_lighting = c->skip_indirect_lighting = c->no_background = c->output_sh = 0;
    c->max_diff_depth = 4;
    c->max_spec_depth = 8;
    c->max_refr_depth = 8;
    c->max_transp_depth = 8;
    c->max_total_depth = 8;
    c->min_total_depth = 2;
    c->min_transp_depth = 2;
    c->clamp_direct = c->clamp_indirect = 0.0f;
    c->min_samples = 128;
    c->variance_threshold = 0.0f;
    c->regularize_alpha = 0.03f;
}

#ifdef __cplusplus
}
#endif
#endif /* RAY_SCENE_DESC_H */
