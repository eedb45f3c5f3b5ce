// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE: a flat C API over the UNMODIFIED reference (sergcpp/Ray).
//
// This file is OURS; it includes the reference's headers from $(REF) (see oracle/Makefile) and is linked against the
// reference compiled as-is into oracle/_ref/libray_ref.a.  It exists so tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg can (a) build a scene with the reference's own Cpu::Scene (SAH build, BVH8
// flatten, material lowering, light tree) and hand the resulting arrays to the CUDA backend byte-for-byte ("oracle
// mode 1b" of SURVEY.md section 8(c)), (b) run the reference's own renderers and stage functions on the same inputs, and
// (c) time the reference's CPU backends.  Nothing under ray_b200/ may link or call this library.
//
// Reference entry points used (file:line in the reference tree):
//   Cpu::Scene                        internal/SceneCPU.h:40-169   (subclassed only to reach its protected arrays)
//   scene_data_t assembly             internal/RendererCPU.h:390-413 (restated in make_scene_data below)
//   Ref::GeneratePrimaryRays          internal/CoreRef.cpp:1429     Ref::TraceRays        internal/CoreRef.cpp:4841
//   Ref::ShadePrimary/ShadeSecondary  internal/ShadeRef.cpp:1654,1702
//   Ref::TraceShadowRays              internal/CoreRef.cpp:4856     Ref::SortRays_CPU     internal/CoreRef.cpp:1667
//   {Ref,Sse41,Avx,Avx2,Avx512}::CreateRenderer  internal/Renderer*.h
#include <atomic>
#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <list>
#include <vector>

#include "Ray.h"
#include "internal/CDFUtils.h"
#include "internal/CoreRef.h"
#include "internal/RendererAVX.h"
#include "internal/RendererAVX2.h"
#include "internal/RendererAVX512.h"
#include "internal/RendererRef.h"
#include "internal/RendererSSE41.h"
#include "internal/SceneCPU.h"
#include "internal/ShadeRef.h"
#include "internal/simd/detect.h"

#include "../include/ray_cuda.h"
#include "../include/ray_scene_desc.h"

namespace Ray {
extern const uint32_t *transform_luts[]; // TonemapRef.cpp:15
}
using namespace Ray;

// the OIDN weight set the reference's UNet filter uses (Apache-2.0 data file of the reference tree, included from where
// it lies): handed to the CUDA backend through rc_unet_set_weights by the parity tests
namespace oidn_hdr_alb_nrm {
#include "internal/precomputed/__oidn_weights_hdr_alb_nrm.inl"
}

namespace {

class CountingLog final : public ILog {
  public:
    std::atomic<int> errors{0};
    bool verbose = false;
    void Info(const char *fmt, ...) override {
        if (verbose) {
            va_list vl;
            va_start(vl, fmt);
            vprintf(fmt, vl);
            va_end(vl);
            putc('\n', stdout);
        }
    }
    void Warning(const char *fmt, ...) override {
        if (verbose) {
            va_list vl;
            va_start(vl, fmt);
            vprintf(fmt, vl);
            va_end(vl);
            putc('\n', stdout);
        }
    }
    void Error(const char *fmt, ...) override {
        ++errors;
        va_list vl;
        va_start(vl, fmt);
        vfprintf(stderr, fmt, vl);
        va_end(vl);
        fputc('\n', stderr);
    }
};

CountingLog g_log;

// Cpu::Scene keeps its arrays protected; a subclass is the documented way in (SURVEY.md section 8(c) mode 1b).
class OracleScene final : public Cpu::Scene {
  public:
    explicit OracleScene(bool wide, bool tex_compression = false)
        : Cpu::Scene(&g_log, wide, tex_compression, false /* spatial cache */) {}

    bool wide() const { return use_wide_bvh_; }

    void fill_view(rc_scene_view &v) const {
        memset(&v, 0, sizeof(v));
        // only the live prefix of each SparseStorage is handed over (nothing is ever removed through this harness)
        v.wnodes = {wnodes_.data(), wnodes_.size(), sizeof(wbvh_node_t)};
        v.mtris = {mtris_.data(), mtris_.size(), sizeof(mtri_accel_t)};
        v.tri_indices = {tri_indices_.data(), tri_indices_.size(), sizeof(uint32_t)};
        v.tri_materials = {tri_materials_.data(), tri_materials_.size(), sizeof(tri_mat_data_t)};
        v.materials = {materials_.data(), materials_.size(), sizeof(material_t)};
        v.mesh_instances = {mesh_instances_.data(), mesh_instances_.size(), sizeof(mesh_instance_t)};
        v.vertices = {vertices_.data(), vertices_.size(), sizeof(vertex_t)};
        v.vtx_indices = {vtx_indices_.data(), vtx_indices_.size(), sizeof(uint32_t)};
        v.lights = {lights_.data(), lights_.size(), sizeof(light_t)};
        v.li_indices = {li_indices_.data(), uint32_t(li_indices_.size()), sizeof(uint32_t)};
        v.light_cwnodes = {light_cwnodes_.data(), uint32_t(light_cwnodes_.size()), sizeof(light_cwbvh_node_t)};
        v.tlas_root = tlas_root_;
        v.visible_lights_count = visible_lights_count_;
        v.blocker_lights_count = blocker_lights_count_;
        memcpy(v.env_col, env_.env_col, sizeof(v.env_col));
        v.env_map = env_.env_map;
        memcpy(v.back_col, env_.back_col, sizeof(v.back_col));
        v.back_map = env_.back_map;
        v.env_light_index = env_.light_index;
        v.sky_map_spread_angle = env_.sky_map_spread_angle;
        GetBounds(v.bounds_min, v.bounds_max);
        export_textures(v);
        v.env_map_rotation = env_.env_map_rotation;
        v.back_map_rotation = env_.back_map_rotation;
        v.qtree_levels = env_.qtree_levels;
        for (int i = 0; i < 16; ++i) {
            v.qtree_mips[i] = (i < env_.qtree_levels) ? env_.qtree_mips[i] : nullptr;
        }
    }

    const camera_t &cam() const { return cams_[current_cam_._index]; }

    // Textures cross the C-ABI decoded (include/ray_cuda.h rc_texture): walk the storage's own Fetch(), which applies the
    // swizzle / block decode and the channel expansion, and turn byte / 255.0f back into the byte.
    void add_texture_handle(uint32_t h) { tex_handles_.push_back(h); }
    void export_textures(rc_scene_view &v) const {
        tex_pixels_.clear();
        tex_views_.clear();
        for (const uint32_t h : tex_handles_) {
            const Cpu::TexStorageBase *st = tex_storages_[h >> 28];
            const int index = int(h & 0x00ffffff);
            rc_texture t = {};
            t.handle = h & 0xf0ffffffu;
            t.channels = 4;
            for (int lod = 0; lod < NUM_MIP_LEVELS; ++lod) {
                int res[2];
                st->GetIRes(index, lod, res);
                t.res[lod][0] = uint16_t(res[0]);
                t.res[lod][1] = uint16_t(res[1]);
                if (lod > 0 && res[0] == t.res[lod - 1][0] && res[1] == t.res[lod - 1][1]) {
                    t.pixels[lod] = t.pixels[lod - 1]; // level absent: aliases the previous one (Allocate, :244-249)
                    continue;
                }
                tex_pixels_.emplace_back(size_t(res[0]) * res[1] * 4);
                uint8_t *dst = tex_pixels_.back().data();
                for (int y = 0; y < res[1]; ++y) {
                    for (int x = 0; x < res[0]; ++x) {
                        const color_rgba_t c = st->Fetch(index, x, y, lod);
                        for (int k = 0; k < 4; ++k) {
                            dst[(size_t(y) * res[0] + x) * 4 + k] = uint8_t(lrintf(c.v[k] * 255.0f));
                        }
                    }
                }
                t.pixels[lod] = dst;
            }
            tex_views_.push_back(t);
        }
        v.textures = tex_views_.empty() ? nullptr : tex_views_.data();
        v.texture_count = uint32_t(tex_views_.size());
    }

    // scene_data_t exactly as Cpu::Renderer<P>::RenderScene assembles it
    scene_data_t make_scene_data(const cache_grid_params_t &cache_grid) const {
        return scene_data_t{env_,
                            mesh_instances_.empty() ? nullptr : &mesh_instances_[0],
                            meshes_.empty() ? nullptr : &meshes_[0],
                            vtx_indices_.empty() ? nullptr : &vtx_indices_[0],
                            vertices_.empty() ? nullptr : &vertices_[0],
                            nodes_.empty() ? nullptr : &nodes_[0],
                            wnodes_.empty() ? nullptr : &wnodes_[0],
                            tris_.empty() ? nullptr : &tris_[0],
                            tri_indices_.empty() ? nullptr : &tri_indices_[0],
                            mtris_.data(),
                            tri_materials_.empty() ? nullptr : &tri_materials_[0],
                            materials_.empty() ? nullptr : &materials_[0],
                            {lights_.data(), lights_.capacity()},
                            {li_indices_},
                            {dir_lights_},
                            visible_lights_count_,
                            blocker_lights_count_,
                            {light_nodes_},
                            {light_cwnodes_},
                            {sky_transmittance_lut_},
                            {sky_multiscatter_lut_},
                            cache_grid,
                            {spatial_cache_entries_},
                            {spatial_cache_voxels_prev_}};
    }

    const Cpu::TexStorageBase *const *textures() const { return tex_storages_; }

  private:
    std::vector<uint32_t> tex_handles_;
    mutable std::list<std::vector<uint8_t>> tex_pixels_;
    mutable std::vector<rc_texture> tex_views_;

  public:
    uint32_t tlas_root() const { return tlas_root_; }
    uint32_t counts(int which) const {
        switch (which) {
        case 0: return wnodes_.size();
        case 1: return mtris_.size() * 8;
        case 2: return uint32_t(tri_materials_.size());
        case 3: return lights_.size();
        case 4: return uint32_t(light_cwnodes_.size());
        case 5: return mesh_instances_.size();
        case 6: return vertices_.size();
        default: return 0;
        }
    }
};

MaterialHandle mh(uint32_t i) { return MaterialHandle{i, 0}; }
TextureHandle th(uint32_t i) { return TextureHandle{i, 0}; }

vtx_attribute_t attr(const rs_vtx_attribute &a) {
    vtx_attribute_t r;
    r.data = Span<const float>{a.data, a.data ? size_t(a.count) : size_t(0)};
    r.offset = a.offset;
    r.stride = a.stride;
    return r;
}

template <typename D> void light_common(D &d, const rs_light_common &c) {
    memcpy(d.color, c.color, sizeof(d.color));
    d.multiple_importance = c.multiple_importance != 0;
    d.cast_shadow = c.cast_shadow != 0;
    d.diffuse_visibility = c.diffuse_visibility != 0;
    d.specular_visibility = c.specular_visibility != 0;
    d.refraction_visibility = c.refraction_visibility != 0;
}

struct OracleRenderer {
    std::unique_ptr<RendererBase> r;
    std::vector<std::unique_ptr<RegionContext>> regions;
};

std::vector<float> make_filter_table(ePixelFilter filter, float filter_width) {
    // restates Cpu::Renderer<P>::UpdateFilterTable (internal/RendererCPU.h:1234-1258) on top of the reference's CDFInverted
    float (*filter_func)(float v, float width) = filter_box;
    switch (filter) {
    case ePixelFilter::Box:
        filter_func = filter_box;
        filter_width = 1.0f;
        break;
    case ePixelFilter::Gaussian:
        filter_func = filter_gaussian;
        filter_width *= 3.0f;
        break;
    case ePixelFilter::BlackmanHarris:
        filter_func = filter_blackman_harris;
        filter_width *= 2.0f;
        break;
    default:
        break;
    }
    return Ray::CDFInverted(FILTER_TABLE_SIZE, 0.0f, filter_width * 0.5f,
                            std::bind(filter_func, std::placeholders::_1, filter_width), true);
}

} // namespace

extern "C" {

typedef struct ro_scene ro_scene;       // OracleScene
typedef struct ro_renderer ro_renderer; // OracleRenderer

const uint32_t *ro_pmj_table(int *dims, int *samples) {
    if (dims) {
        *dims = RAND_DIMS_COUNT;
    }
    if (samples) {
        *samples = RAND_SAMPLES_COUNT;
    }
    return __pmj02_samples;
}

int ro_error_count(void) { return g_log.errors.load(); }
void ro_set_verbose(int v) { g_log.verbose = v != 0; }

int ro_cpu_features(void) {
    const CpuFeatures f = GetCpuFeatures();
    return int(f.sse41_supported) | (int(f.avx_supported) << 1) | (int(f.avx2_supported) << 2) |
           (int(f.avx512_supported) << 3);
}

ro_scene *ro_scene_create(int use_wide_bvh) { return reinterpret_cast<ro_scene *>(new OracleScene(use_wide_bvh != 0)); }
// use_tex_compression: settings_t::use_tex_compression of the reference (BCn blocks, YCoCg-coded base colour maps)
ro_scene *ro_scene_create_ex(int use_wide_bvh, int use_tex_compression) {
    return reinterpret_cast<ro_scene *>(new OracleScene(use_wide_bvh != 0, use_tex_compression != 0));
}
void ro_scene_destroy(ro_scene *s) { delete reinterpret_cast<OracleScene *>(s); }

uint32_t ro_add_texture(ro_scene *s, const rs_tex_desc *d) {
    tex_desc_t t;
    t.format = eTextureFormat(d->format);
    t.convention = eTextureConvention(d->convention);
    const int channels = TexFormatChannelCount[d->format];
    t.data = Span<const uint8_t>{d->data, size_t(d->w) * d->h * channels};
    t.w = d->w;
    t.h = d->h;
    t.is_srgb = d->is_srgb != 0;
    t.is_normalmap = d->is_normalmap != 0;
    t.generate_mipmaps = d->generate_mipmaps != 0;
    t.reconstruct_z = d->reconstruct_z != 0;
    t.force_no_compression = false; // the scene's own use_tex_compression decides (off unless ro_scene_create_ex asked)
    const TextureHandle h = reinterpret_cast<OracleScene *>(s)->AddTexture(t);
    if (h._index != 0xffffffffu) {
        reinterpret_cast<OracleScene *>(s)->add_texture_handle(h._index);
    }
    return h._index;
}

uint32_t ro_add_material_node(ro_scene *s, const rs_shading_node_desc *d) {
    shading_node_desc_t m;
    m.type = eShadingNode(d->type);
    memcpy(m.base_color, d->base_color, sizeof(m.base_color));
    m.base_texture = th(d->base_texture);
    m.normal_map = th(d->normal_map);
    m.normal_map_intensity = d->normal_map_intensity;
    m.mix_materials[0] = mh(d->mix_materials[0]);
    m.mix_materials[1] = mh(d->mix_materials[1]);
    m.roughness = d->roughness;
    m.roughness_texture = th(d->roughness_texture);
    m.anisotropic = d->anisotropic;
    m.anisotropic_rotation = d->anisotropic_rotation;
    m.sheen = d->sheen;
    m.specular = d->specular;
    m.strength = d->strength;
    m.fresnel = d->fresnel;
    m.ior = d->ior;
    m.tint = d->tint;
    m.metallic_texture = th(d->metallic_texture);
    m.importance_sample = d->importance_sample != 0;
    m.mix_add = d->mix_add != 0;
    return reinterpret_cast<OracleScene *>(s)->AddMaterial(m)._index;
}

uint32_t ro_add_material_principled(ro_scene *s, const rs_principled_mat_desc *d) {
    principled_mat_desc_t m;
    memcpy(m.base_color, d->base_color, sizeof(m.base_color));
    m.base_texture = th(d->base_texture);
    m.metallic = d->metallic;
    m.metallic_texture = th(d->metallic_texture);
    m.specular = d->specular;
    m.specular_texture = th(d->specular_texture);
    m.specular_tint = d->specular_tint;
    m.roughness = d->roughness;
    m.roughness_texture = th(d->roughness_texture);
    m.anisotropic = d->anisotropic;
    m.anisotropic_rotation = d->anisotropic_rotation;
    m.sheen = d->sheen;
    m.sheen_tint = d->sheen_tint;
    m.clearcoat = d->clearcoat;
    m.clearcoat_roughness = d->clearcoat_roughness;
    m.ior = d->ior;
    m.transmission = d->transmission;
    m.transmission_roughness = d->transmission_roughness;
    memcpy(m.emission_color, d->emission_color, sizeof(m.emission_color));
    m.emission_texture = th(d->emission_texture);
    m.emission_strength = d->emission_strength;
    m.alpha = d->alpha;
    m.alpha_texture = th(d->alpha_texture);
    m.normal_map = th(d->normal_map);
    m.normal_map_intensity = d->normal_map_intensity;
    m.importance_sample = d->importance_sample != 0;
    return reinterpret_cast<OracleScene *>(s)->AddMaterial(m)._index;
}

uint32_t ro_add_mesh(ro_scene *s, const rs_mesh_desc *d) {
    mesh_desc_t m;
    m.name = "mesh";
    m.prim_type = ePrimType::TriangleList;
    m.vtx_positions = attr(d->vtx_positions);
    m.vtx_normals = attr(d->vtx_normals);
    m.vtx_binormals = attr(d->vtx_binormals);
    m.vtx_uvs = attr(d->vtx_uvs);
    m.vtx_indices = Span<const uint32_t>{d->vtx_indices, size_t(d->vtx_indices_count)};
    m.base_vertex = d->base_vertex;
    std::vector<mat_group_desc_t> groups;
    for (uint32_t i = 0; i < d->groups_count; ++i) {
        groups.emplace_back(mh(d->groups[i].front_mat), mh(d->groups[i].back_mat), size_t(d->groups[i].vtx_start),
                            size_t(d->groups[i].vtx_count));
    }
    m.groups = groups;
    m.allow_spatial_splits = d->allow_spatial_splits != 0;
    m.use_fast_bvh_build = d->use_fast_bvh_build != 0;
    return reinterpret_cast<OracleScene *>(s)->AddMesh(m)._index;
}

uint32_t ro_add_mesh_instance(ro_scene *s, const rs_mesh_instance_desc *d) {
    mesh_instance_desc_t mi;
    mi.xform = d->xform;
    mi.mesh = MeshHandle{d->mesh, 0};
    mi.camera_visibility = d->camera_visibility != 0;
    mi.diffuse_visibility = d->diffuse_visibility != 0;
    mi.specular_visibility = d->specular_visibility != 0;
    mi.refraction_visibility = d->refraction_visibility != 0;
    mi.shadow_visibility = d->shadow_visibility != 0;
    return reinterpret_cast<OracleScene *>(s)->AddMeshInstance(mi)._index;
}

uint32_t ro_add_light_directional(ro_scene *s, const rs_directional_light_desc *d) {
    directional_light_desc_t l;
    light_common(l, d->c);
    memcpy(l.direction, d->direction, sizeof(l.direction));
    l.angle = d->angle;
    return reinterpret_cast<OracleScene *>(s)->AddLight(l)._index;
}
uint32_t ro_add_light_sphere(ro_scene *s, const rs_sphere_light_desc *d) {
    sphere_light_desc_t l;
    light_common(l, d->c);
    memcpy(l.position, d->position, sizeof(l.position));
    l.radius = d->radius;
    return reinterpret_cast<OracleScene *>(s)->AddLight(l)._index;
}
uint32_t ro_add_light_spot(ro_scene *s, const rs_spot_light_desc *d) {
    spot_light_desc_t l;
    light_common(l, d->c);
    memcpy(l.position, d->position, sizeof(l.position));
    memcpy(l.direction, d->direction, sizeof(l.direction));
    l.spot_size = d->spot_size;
    l.spot_blend = d->spot_blend;
    l.radius = d->radius;
    return reinterpret_cast<OracleScene *>(s)->AddLight(l)._index;
}
uint32_t ro_add_light_rect(ro_scene *s, const rs_rect_light_desc *d) {
    rect_light_desc_t l;
    light_common(l, d->c);
    l.width = d->width;
    l.height = d->height;
    l.doublesided = d->doublesided != 0;
    l.sky_portal = d->sky_portal != 0;
    return reinterpret_cast<OracleScene *>(s)->AddLight(l, d->xform)._index;
}
uint32_t ro_add_light_disk(ro_scene *s, const rs_disk_light_desc *d) {
    disk_light_desc_t l;
    light_common(l, d->c);
    l.size_x = d->size_x;
    l.size_y = d->size_y;
    l.doublesided = d->doublesided != 0;
    l.sky_portal = d->sky_portal != 0;
    return reinterpret_cast<OracleScene *>(s)->AddLight(l, d->xform)._index;
}
uint32_t ro_add_light_line(ro_scene *s, const rs_line_light_desc *d) {
    line_light_desc_t l;
    light_common(l, d->c);
    l.radius = d->radius;
    l.height = d->height;
    l.sky_portal = d->sky_portal != 0;
    return reinterpret_cast<OracleScene *>(s)->AddLight(l, d->xform)._index;
}

void ro_set_environment(ro_scene *s, const rs_environment_desc *d) {
    environment_desc_t e;
    memcpy(e.env_col, d->env_col, sizeof(e.env_col));
    memcpy(e.back_col, d->back_col, sizeof(e.back_col));
    e.importance_sample = d->importance_sample != 0;
    e.env_map = th(d->env_map);
    e.back_map = th(d->back_map);
    e.env_map_rotation = d->env_map_rotation;
    e.back_map_rotation = d->back_map_rotation;
    reinterpret_cast<OracleScene *>(s)->SetEnvironment(e);
}

uint32_t ro_add_camera(ro_scene *s, const rs_camera_desc *d) {
    camera_desc_t c;
    c.type = eCamType(d->type);
    c.filter = ePixelFilter(d->filter);
    c.view_transform = eViewTransform(d->view_transform);
    c.ltype = eLensUnits(d->ltype);
    c.filter_width = d->filter_width;
    memcpy(c.origin, d->origin, sizeof(c.origin));
    memcpy(c.fwd, d->fwd, sizeof(c.fwd));
    memcpy(c.up, d->up, sizeof(c.up));
    memcpy(c.shift, d->shift, sizeof(c.shift));
    c.exposure = d->exposure;
    c.fov = d->fov;
    c.gamma = d->gamma;
    c.sensor_height = d->sensor_height;
    c.focus_distance = d->focus_distance;
    c.focal_length = d->focal_length;
    c.fstop = d->fstop;
    c.lens_rotation = d->lens_rotation;
    c.lens_ratio = d->lens_ratio;
    c.lens_blades = d->lens_blades;
    c.clip_start = d->clip_start;
    c.clip_end = d->clip_end;
    c.mi_index = d->mi_index;
    c.uv_index = d->uv_index;
    c.lighting_only = d->lighting_only != 0;
    c.skip_direct_lighting = d->skip_direct_lighting != 0;
    c.skip_indirect_lighting = d->skip_indirect_lighting != 0;
    c.no_background = d->no_background != 0;
    c.output_sh = d->output_sh != 0;
    c.max_diff_depth = uint8_t(d->max_diff_depth);
    c.max_spec_depth = uint8_t(d->max_spec_depth);
    c.max_refr_depth = uint8_t(d->max_refr_depth);
    c.max_transp_depth = uint8_t(d->max_transp_depth);
    c.max_total_depth = uint8_t(d->max_total_depth);
    c.min_total_depth = uint8_t(d->min_total_depth);
    c.min_transp_depth = uint8_t(d->min_transp_depth);
    c.clamp_direct = d->clamp_direct;
    c.clamp_indirect = d->clamp_indirect;
    c.min_samples = d->min_samples;
    c.variance_threshold = d->variance_threshold;
    c.regularize_alpha = d->regularize_alpha;
    auto *sc = reinterpret_cast<OracleScene *>(s);
    const CameraHandle h = sc->AddCamera(c);
    sc->set_current_cam(h);
    return h._index;
}

void ro_finalize(ro_scene *s) { reinterpret_cast<OracleScene *>(s)->Finalize(parallel_for_serial); }

void ro_scene_view(ro_scene *s, rc_scene_view *out) { reinterpret_cast<OracleScene *>(s)->fill_view(*out); }

uint32_t ro_scene_count(ro_scene *s, int which) { return reinterpret_cast<OracleScene *>(s)->counts(which); }

// camera_t -> rc_camera (the conversion Cuda::Renderer::RenderScene does before rc_render)
void ro_get_camera(ro_scene *s, rc_camera *out) {
    const camera_t &c = reinterpret_cast<OracleScene *>(s)->cam();
    memset(out, 0, sizeof(*out));
    out->type = uint32_t(c.type);
    out->filter = uint32_t(c.filter);
    out->view_transform = uint32_t(c.view_transform);
    out->fov = c.fov;
    out->exposure = c.exposure;
    out->gamma = c.gamma;
    out->sensor_height = c.sensor_height;
    out->focus_distance = c.focus_distance;
    out->focal_length = c.focal_length;
    out->fstop = c.fstop;
    out->lens_rotation = c.lens_rotation;
    out->lens_ratio = c.lens_ratio;
    out->lens_blades = c.lens_blades;
    out->clip_start = c.clip_start;
    out->clip_end = c.clip_end;
    memcpy(out->origin, c.origin, sizeof(out->origin));
    memcpy(out->fwd, c.fwd, sizeof(out->fwd));
    memcpy(out->side, c.side, sizeof(out->side));
    memcpy(out->up, c.up, sizeof(out->up));
    memcpy(out->shift, c.shift, sizeof(out->shift));
    out->max_diff_depth = c.pass_settings.max_diff_depth;
    out->max_spec_depth = c.pass_settings.max_spec_depth;
    out->max_refr_depth = c.pass_settings.max_refr_depth;
    out->max_transp_depth = c.pass_settings.max_transp_depth;
    out->max_total_depth = c.pass_settings.max_total_depth;
    out->min_total_depth = c.pass_settings.min_total_depth;
    out->min_transp_depth = c.pass_settings.min_transp_depth;
    out->clamp_direct = c.pass_settings.clamp_direct;
    out->clamp_indirect = c.pass_settings.clamp_indirect;
    out->min_samples = c.pass_settings.min_samples;
    out->variance_threshold = c.pass_settings.variance_threshold;
    out->regularize_alpha = c.pass_settings.regularize_alpha;
}

// 1024-entry inverse-CDF table of the current camera's pixel filter
void ro_get_filter_table(ro_scene *s, float out[1024]) {
    const camera_t &c = reinterpret_cast<OracleScene *>(s)->cam();
    const std::vector<float> t = make_filter_table(c.filter, c.filter_width);
    memcpy(out, t.data(), sizeof(float) * FILTER_TABLE_SIZE);
}

// ---- whole renderers ----------------------------------------------------------------------------------------------
// type: eRendererType value (0 Reference, 1 SSE41, 2 AVX, 3 AVX2, 4 AVX512); returns null if the CPU lacks the ISA
ro_renderer *ro_renderer_create(int type, int w, int h) {
    settings_t st;
    st.w = w;
    st.h = h;
    st.use_tex_compression = false;
    st.use_spatial_cache = false;
    const CpuFeatures f = GetCpuFeatures();
    RendererBase *r = nullptr;
    switch (eRendererType(type)) {
    case eRendererType::Reference: r = Ref::CreateRenderer(st, &g_log); break;
    case eRendererType::SIMD_SSE41: r = f.sse41_supported ? Sse41::CreateRenderer(st, &g_log) : nullptr; break;
    case eRendererType::SIMD_AVX: r = f.avx_supported ? Avx::CreateRenderer(st, &g_log) : nullptr; break;
    case eRendererType::SIMD_AVX2: r = f.avx2_supported ? Avx2::CreateRenderer(st, &g_log) : nullptr; break;
    case eRendererType::SIMD_AVX512: r = f.avx512_supported ? Avx512::CreateRenderer(st, &g_log) : nullptr; break;
    default: break;
    }
    if (!r) {
        return nullptr;
    }
    auto *o = new OracleRenderer();
    o->r.reset(r);
    return reinterpret_cast<ro_renderer *>(o);
}

void ro_renderer_destroy(ro_renderer *r) { delete reinterpret_cast<OracleRenderer *>(r); }

void ro_renderer_clear(ro_renderer *r, const float rgba[4]) {
    reinterpret_cast<OracleRenderer *>(r)->r->Clear(color_rgba_t{rgba[0], rgba[1], rgba[2], rgba[3]});
}

// one RenderScene call over `rect`; *iteration is RegionContext::iteration before (in) and after (out) the call
void ro_render(ro_renderer *r, ro_scene *s, const rc_rect *rect, int *iteration) {
    RegionContext region(rect_t{rect->x, rect->y, rect->w, rect->h});
    region.iteration = *iteration;
    reinterpret_cast<OracleRenderer *>(r)->r->RenderScene(*reinterpret_cast<OracleScene *>(s), region);
    *iteration = region.iteration;
}

// RendererBase::DenoiseImage(const RegionContext &) (the NLM overload) over `rect` at RegionContext::iteration = iteration
void ro_denoise(ro_renderer *r, const rc_rect *rect, int iteration) {
    RegionContext region(rect_t{rect->x, rect->y, rect->w, rect->h});
    region.iteration = iteration;
    reinterpret_cast<OracleRenderer *>(r)->r->DenoiseImage(region);
}

// which: 0 final (tonemapped), 1 raw, 2 base colour, 3 depth-normals
const float *ro_get_pixels(ro_renderer *r, int which, int *pitch) {
    RendererBase *rb = reinterpret_cast<OracleRenderer *>(r)->r.get();
    color_data_rgba_t d{};
    switch (which) {
    case 0: d = rb->get_pixels_ref(); break;
    case 1: d = rb->get_raw_pixels_ref(); break;
    case 2: d = rb->get_aux_pixels_ref(eAUXBuffer::BaseColor); break;
    case 3: d = rb->get_aux_pixels_ref(eAUXBuffer::DepthNormals); break;
    default: break;
    }
    if (pitch) {
        *pitch = d.pitch;
    }
    return d.ptr ? d.ptr->v : nullptr;
}

void ro_get_stats(ro_renderer *r, uint64_t us[11]) {
    RendererBase::stats_t st = {};
    reinterpret_cast<OracleRenderer *>(r)->r->GetStats(st);
    const unsigned long long v[11] = {st.time_primary_ray_gen_us, st.time_primary_trace_us, st.time_primary_shade_us,
                                      st.time_primary_shadow_us,  st.time_secondary_sort_us, st.time_secondary_trace_us,
                                      st.time_secondary_shade_us, st.time_secondary_shadow_us, st.time_denoise_us,
                                      st.time_cache_update_us,    st.time_cache_resolve_us};
    for (int i = 0; i < 11; ++i) {
        us[i] = v[i];
    }
}

// Multi-threaded render of `spp` samples over the whole frame: `tile`x`tile` regions, one RegionContext per tile,
// dynamic queue over `threads` std::threads (the pattern of the reference's samples/02_multithreading and
// tests/test_scene.cpp:1027-1084).  Returns wall-clock seconds of the RenderScene calls.
double ro_render_mt(ro_renderer *r, ro_scene *s, int w, int h, int spp, int threads, int tile) {
    RendererBase *rb = reinterpret_cast<OracleRenderer *>(r)->r.get();
    OracleScene *sc = reinterpret_cast<OracleScene *>(s);
    std::vector<RegionContext> regions;
    for (int y = 0; y < h; y += tile) {
        for (int x = 0; x < w; x += tile) {
            regions.emplace_back(rect_t{x, y, std::min(tile, w - x), std::min(tile, h - y)});
        }
    }
    if (!RendererSupportsMultithreading(rb->type())) {
        threads = 1;
    }
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < spp; ++i) {
        std::atomic<int> next{0};
        auto worker = [&]() {
            for (;;) {
                const int k = next.fetch_add(1);
                if (k >= int(regions.size())) {
                    break;
                }
                rb->RenderScene(*sc, regions[k]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) {
            pool.emplace_back(worker);
        }
        worker();
        for (auto &t : pool) {
            t.join();
        }
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- stage functions (Ref::*) on caller-owned AoS buffers ----------------------------------------------------------
static uint32_t rand_seed_for(int iteration) { return Ref::hash(uint32_t((iteration - 1) / RAND_SAMPLES_COUNT)); }

int ro_stage_generate_primary_rays(ro_scene *s, int w, int h, const rc_rect *rect, int iteration, void *rays_out,
                                   void *hits_out) {
    OracleScene *sc = reinterpret_cast<OracleScene *>(s);
    const camera_t &cam = sc->cam();
    const std::vector<float> table = make_filter_table(cam.filter, cam.filter_width);
    aligned_vector<Ref::ray_data_t> rays;
    aligned_vector<Ref::hit_data_t> hits;
    Ref::GeneratePrimaryRays(cam, rect_t{rect->x, rect->y, rect->w, rect->h}, w, h, __pmj02_samples,
                             rand_seed_for(iteration), table.data(), iteration, nullptr, rays, hits);
    memcpy(rays_out, rays.data(), rays.size() * sizeof(Ref::ray_data_t));
    memcpy(hits_out, hits.data(), hits.size() * sizeof(Ref::hit_data_t));
    return int(rays.size());
}

void ro_stage_trace_rays(ro_scene *s, int iteration, void *rays, void *hits, int count, int trace_lights) {
    OracleScene *sc = reinterpret_cast<OracleScene *>(s);
    const camera_t &cam = sc->cam();
    cache_grid_params_t cg;
    const scene_data_t sd = sc->make_scene_data(cg);
    if (sc->tlas_root() == 0xffffffff) {
        return;
    }
    Ref::TraceRays(Span<Ref::ray_data_t>{static_cast<Ref::ray_data_t *>(rays), count}, cam.pass_settings.min_transp_depth,
                   cam.pass_settings.max_transp_depth, sd, sc->tlas_root(), trace_lights != 0, sc->textures(),
                   __pmj02_samples, rand_seed_for(iteration), iteration,
                   Span<Ref::hit_data_t>{static_cast<Ref::hit_data_t *>(hits), count});
}

// temp/base_color/depth_normals: w*h RGBA float images (in/out).  Returns nothing; counts through the out params.
void ro_stage_shade(ro_scene *s, int w, int h, int iteration, int primary, int bounce, const void *rays,
                    const void *hits, int count, void *secondary_out, int *secondary_count, void *shadow_out,
                    int *shadow_count, float *temp, float *base_color, float *depth_normals) {
    OracleScene *sc = reinterpret_cast<OracleScene *>(s);
    const camera_t &cam = sc->cam();
    cache_grid_params_t cg;
    const scene_data_t sd = sc->make_scene_data(cg);
    std::vector<uint32_t> def_sky(size_t(count) + 1);
    int def_sky_count = 0;
    *secondary_count = *shadow_count = 0;
    const Span<const Ref::hit_data_t> hs{static_cast<const Ref::hit_data_t *>(hits), count};
    const Span<const Ref::ray_data_t> rs{static_cast<const Ref::ray_data_t *>(rays), count};
    if (primary) {
        Ref::ShadePrimary(cam.pass_settings, hs, rs, __pmj02_samples, rand_seed_for(iteration), iteration,
                          eSpatialCacheMode::None, sd, sc->textures(), static_cast<Ref::ray_data_t *>(secondary_out),
                          secondary_count, static_cast<Ref::shadow_ray_t *>(shadow_out), shadow_count, def_sky.data(),
                          &def_sky_count, w, 1.0f / float(iteration), reinterpret_cast<color_rgba_t *>(temp),
                          reinterpret_cast<color_rgba_t *>(base_color), reinterpret_cast<color_rgba_t *>(depth_normals));
    } else {
        const float clamp_direct = (bounce == 1) ? cam.pass_settings.clamp_direct : cam.pass_settings.clamp_indirect;
        Ref::ShadeSecondary(cam.pass_settings, clamp_direct, hs, rs, __pmj02_samples, rand_seed_for(iteration),
                            iteration, eSpatialCacheMode::None, sd, sc->textures(),
                            static_cast<Ref::ray_data_t *>(secondary_out), secondary_count,
                            static_cast<Ref::shadow_ray_t *>(shadow_out), shadow_count, def_sky.data(), &def_sky_count,
                            w, reinterpret_cast<color_rgba_t *>(temp), nullptr, nullptr);
    }
    (void)h;
}

void ro_stage_trace_shadow_rays(ro_scene *s, int w, int iteration, const void *shadow_rays, int count, float clamp_val,
                                float *temp) {
    OracleScene *sc = reinterpret_cast<OracleScene *>(s);
    const camera_t &cam = sc->cam();
    cache_grid_params_t cg;
    const scene_data_t sd = sc->make_scene_data(cg);
    if (sc->tlas_root() == 0xffffffff) {
        return;
    }
    Ref::TraceShadowRays(Span<const Ref::shadow_ray_t>{static_cast<const Ref::shadow_ray_t *>(shadow_rays), count},
                         cam.pass_settings.max_transp_depth, clamp_val, sd, sc->tlas_root(), __pmj02_samples,
                         rand_seed_for(iteration), iteration, sc->textures(), w, reinterpret_cast<color_rgba_t *>(temp));
}

// ---- the same Ref:: stage functions over CALLER-PROVIDED scene arrays (an rc_scene_view) ---------------------------
// Lets a CPU test run the reference's own traversal / shading code over the arrays produced by the product's host layer
// (ray_b200/csrc/host: its own BVH8, triangle blocks, light tree), i.e. validate those builders without a GPU.
namespace {
struct ViewScene {
    environment_t env = {};
    cache_grid_params_t cache_grid;
    const Cpu::TexStorageBase *textures[8] = {};
    std::vector<uint32_t> dir_lights;

    scene_data_t make(const rc_scene_view &v) {
        memcpy(env.env_col, v.env_col, sizeof(env.env_col));
        env.env_map = v.env_map;
        memcpy(env.back_col, v.back_col, sizeof(env.back_col));
        env.back_map = v.back_map;
        env.qtree_levels = 0;
        env.importance_sample = true;
        env.light_index = v.env_light_index;
        env.sky_map_spread_angle = v.sky_map_spread_angle;
        return scene_data_t{env,
                            static_cast<const mesh_instance_t *>(v.mesh_instances.ptr),
                            nullptr,
                            static_cast<const uint32_t *>(v.vtx_indices.ptr),
                            static_cast<const vertex_t *>(v.vertices.ptr),
                            nullptr,
                            static_cast<const wbvh_node_t *>(v.wnodes.ptr),
                            nullptr,
                            static_cast<const uint32_t *>(v.tri_indices.ptr),
                            static_cast<const mtri_accel_t *>(v.mtris.ptr),
                            static_cast<const tri_mat_data_t *>(v.tri_materials.ptr),
                            static_cast<const material_t *>(v.materials.ptr),
                            {static_cast<const light_t *>(v.lights.ptr), v.lights.count},
                            {static_cast<const uint32_t *>(v.li_indices.ptr), v.li_indices.count},
                            {dir_lights},
                            v.visible_lights_count,
                            v.blocker_lights_count,
                            {},
                            {static_cast<const light_cwbvh_node_t *>(v.light_cwnodes.ptr), v.light_cwnodes.count},
                            {},
                            {},
                            cache_grid,
                            {},
                            {}};
    }
};

pass_settings_t pass_from(const rc_camera &c) {
    pass_settings_t ps = {};
    ps.max_diff_depth = uint8_t(c.max_diff_depth);
    ps.max_spec_depth = uint8_t(c.max_spec_depth);
    ps.max_refr_depth = uint8_t(c.max_refr_depth);
    ps.max_transp_depth = uint8_t(c.max_transp_depth);
    ps.max_total_depth = uint8_t(c.max_total_depth);
    ps.min_total_depth = uint8_t(c.min_total_depth);
    ps.min_transp_depth = uint8_t(c.min_transp_depth);
    ps.clamp_direct = c.clamp_direct;
    ps.clamp_indirect = c.clamp_indirect;
    ps.min_samples = c.min_samples;
    ps.variance_threshold = c.variance_threshold;
    ps.regularize_alpha = c.regularize_alpha;
    return ps;
}
} // namespace

void ro_view_trace_rays(const rc_scene_view *v, const rc_camera *cam, int iteration, void *rays, void *hits, int count,
                        int trace_lights) {
    ViewScene vs;
    const scene_data_t sd = vs.make(*v);
    if (v->tlas_root == 0xffffffff) {
        return;
    }
    Ref::TraceRays(Span<Ref::ray_data_t>{static_cast<Ref::ray_data_t *>(rays), count}, int(cam->min_transp_depth),
                   int(cam->max_transp_depth), sd, v->tlas_root, trace_lights != 0, vs.textures, __pmj02_samples,
                   rand_seed_for(iteration), iteration, Span<Ref::hit_data_t>{static_cast<Ref::hit_data_t *>(hits), count});
}

void ro_view_shade(const rc_scene_view *v, const rc_camera *cam, int w, int h, int iteration, int primary, int bounce,
                   const void *rays, const void *hits, int count, void *secondary_out, int *secondary_count,
                   void *shadow_out, int *shadow_count, float *temp, float *base_color, float *depth_normals) {
    ViewScene vs;
    const scene_data_t sd = vs.make(*v);
    const pass_settings_t ps = pass_from(*cam);
    std::vector<uint32_t> def_sky(size_t(count) + 1);
    int def_sky_count = 0;
    *secondary_count = *shadow_count = 0;
    const Span<const Ref::hit_data_t> hs{static_cast<const Ref::hit_data_t *>(hits), count};
    const Span<const Ref::ray_data_t> rs{static_cast<const Ref::ray_data_t *>(rays), count};
    if (primary) {
        Ref::ShadePrimary(ps, hs, rs, __pmj02_samples, rand_seed_for(iteration), iteration, eSpatialCacheMode::None, sd,
                          vs.textures, static_cast<Ref::ray_data_t *>(secondary_out), secondary_count,
                          static_cast<Ref::shadow_ray_t *>(shadow_out), shadow_count, def_sky.data(), &def_sky_count, w,
                          1.0f / float(iteration), reinterpret_cast<color_rgba_t *>(temp),
                          reinterpret_cast<color_rgba_t *>(base_color), reinterpret_cast<color_rgba_t *>(depth_normals));
    } else {
        const float clamp_direct = (bounce == 1) ? ps.clamp_direct : ps.clamp_indirect;
        Ref::ShadeSecondary(ps, clamp_direct, hs, rs, __pmj02_samples, rand_seed_for(iteration), iteration,
                            eSpatialCacheMode::None, sd, vs.textures, static_cast<Ref::ray_data_t *>(secondary_out),
                            secondary_count, static_cast<Ref::shadow_ray_t *>(shadow_out), shadow_count, def_sky.data(),
                            &def_sky_count, w, reinterpret_cast<color_rgba_t *>(temp), nullptr, nullptr);
    }
    (void)h;
}

void ro_view_trace_shadow_rays(const rc_scene_view *v, const rc_camera *cam, int w, int iteration,
                               const void *shadow_rays, int count, float clamp_val, float *temp) {
    ViewScene vs;
    const scene_data_t sd = vs.make(*v);
    if (v->tlas_root == 0xffffffff) {
        return;
    }
    Ref::TraceShadowRays(Span<const Ref::shadow_ray_t>{static_cast<const Ref::shadow_ray_t *>(shadow_rays), count},
                         int(cam->max_transp_depth), clamp_val, sd, v->tlas_root, __pmj02_samples,
                         rand_seed_for(iteration), iteration, vs.textures, w, reinterpret_cast<color_rgba_t *>(temp));
}

// One sample of the whole RenderScene sequence (the loop body of Cpu::Renderer<P>::RenderScene, RendererCPU.h:374-606)
// through the reference's own Ref:: stage functions over CALLER-PROVIDED scene arrays, `threads` threads over row
// strips (every pixel is independent: SURVEY.md section 8(e)).  `temp` (w*h RGBA, zeroed by the caller) receives the
// radiance of this sample; rays[0] / rays[1] += closest-hit / shadow rays traced.  Camera (primary rays) comes from
// `cam_scene` (any scene holding the same camera).  Lets a GPU test compare the CUDA backend BITWISE with the reference
// code on the arrays built by the product's own host layer, at full frame size, in seconds.
void ro_view_render_sample(const rc_scene_view *v, const rc_camera *cam, ro_scene *cam_scene, int w, int h, int iteration,
                           int threads, float *temp, unsigned long long rays_out[2]) {
    OracleScene *cs = reinterpret_cast<OracleScene *>(cam_scene);
    const camera_t &camera = cs->cam();
    const std::vector<float> table = make_filter_table(camera.filter, camera.filter_width);
    const pass_settings_t ps = pass_from(*cam);
    const uint32_t seed = rand_seed_for(iteration);
    threads = std::max(threads, 1);
    const int rows = (h + threads - 1) / threads;
    std::atomic<unsigned long long> n_rays{0}, n_shadow{0};
    auto worker = [&](int y0) {
        const int y1 = std::min(h, y0 + rows);
        if (y0 >= y1) {
            return;
        }
        ViewScene vs;
        const scene_data_t sd = vs.make(*v);
        aligned_vector<Ref::ray_data_t> rays, sec;
        aligned_vector<Ref::hit_data_t> hits;
        aligned_vector<Ref::shadow_ray_t> sh;
        Ref::GeneratePrimaryRays(camera, rect_t{0, y0, w, y1 - y0}, w, h, __pmj02_samples, seed, table.data(), iteration,
                                 nullptr, rays, hits);
        color_rgba_t *out = reinterpret_cast<color_rgba_t *>(temp);
        std::vector<uint32_t> def_sky;
        int count = int(rays.size());
        for (int bounce = 0; bounce <= int(ps.max_total_depth) && count != 0; ++bounce) {
            if (v->tlas_root != 0xffffffff) {
                Ref::TraceRays(Span<Ref::ray_data_t>{rays.data(), count}, int(ps.min_transp_depth), int(ps.max_transp_depth),
                               sd, v->tlas_root, bounce != 0, vs.textures, __pmj02_samples, seed, iteration,
                               Span<Ref::hit_data_t>{hits.data(), count});
            }
            n_rays += uint64_t(count);
            sec.resize(size_t(count) + 1);
            sh.resize(size_t(count) + 1);
            def_sky.resize(size_t(count) + 1);
            int n_sec = 0, n_sh = 0, n_sky = 0;
            const Span<const Ref::hit_data_t> hs{hits.data(), count};
            const Span<const Ref::ray_data_t> rs{rays.data(), count};
            if (bounce == 0) {
                Ref::ShadePrimary(ps, hs, rs, __pmj02_samples, seed, iteration, eSpatialCacheMode::None, sd, vs.textures,
                                  sec.data(), &n_sec, sh.data(), &n_sh, def_sky.data(), &n_sky, w, 1.0f / float(iteration),
                                  out, nullptr, nullptr); // AOV planes are not compared here
            } else {
                const float clamp_direct = (bounce == 1) ? ps.clamp_direct : ps.clamp_indirect;
                Ref::ShadeSecondary(ps, clamp_direct, hs, rs, __pmj02_samples, seed, iteration, eSpatialCacheMode::None, sd,
                                    vs.textures, sec.data(), &n_sec, sh.data(), &n_sh, def_sky.data(), &n_sky, w, out,
                                    nullptr, nullptr);
            }
            if (v->tlas_root != 0xffffffff && n_sh != 0) {
                Ref::TraceShadowRays(Span<const Ref::shadow_ray_t>{sh.data(), n_sh}, int(ps.max_transp_depth),
                                     bounce == 0 ? ps.clamp_direct : ps.clamp_indirect, sd, v->tlas_root, __pmj02_samples,
                                     seed, iteration, vs.textures, w, out);
            }
            n_shadow += uint64_t(n_sh);
            rays.assign(sec.begin(), sec.begin() + n_sec);
            hits.assign(size_t(n_sec), Ref::hit_data_t{});
            count = n_sec;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        pool.emplace_back(worker, t * rows);
    }
    worker(0);
    for (auto &t : pool) {
        t.join();
    }
    if (rays_out) {
        rays_out[0] += n_rays.load();
        rays_out[1] += n_shadow.load();
    }
}

// RendererBase::InitUNetFilter + DenoiseImage(pass, region) for every pass over `rect` (the UNet overload)
int ro_denoise_unet(ro_renderer *r, const rc_rect *rect, int iteration) {
    RendererBase *rb = reinterpret_cast<OracleRenderer *>(r)->r.get();
    const unet_filter_properties_t props = rb->InitUNetFilter(false, parallel_for_serial);
    RegionContext region(rect_t{rect->x, rect->y, rect->w, rect->h});
    region.iteration = iteration;
    for (int pass = 0; pass < props.pass_count; ++pass) {
        rb->DenoiseImage(pass, region);
    }
    return props.pass_count;
}

// the 48^3 packed table of AgX / Filmic view transform `vt` (TonemapRef.cpp:5-27), for rc_set_view_lut in the parity tests
const uint32_t *ro_view_lut(int vt) { return (vt > 0 && vt < int(Ray::eViewTransform::_Count)) ? Ray::transform_luts[vt] : nullptr; }

// layer i (pass order) of the reference's UNet weight set: fp16 OIHW weights + fp16 biases
void ro_unet_layer(int i, const uint16_t **weights, int *weights_count, const uint16_t **bias, int *bias_count) {
    using namespace oidn_hdr_alb_nrm;
#define RO_L(n) {n##_weight, int(sizeof(n##_weight) / 2), n##_bias, int(sizeof(n##_bias) / 2)}
    static const struct {
        const uint16_t *w;
        int wn;
        const uint16_t *b;
        int bn;
    } L[16] = {RO_L(enc_conv0),  RO_L(enc_conv1),  RO_L(enc_conv2),  RO_L(enc_conv3),  RO_L(enc_conv4), RO_L(enc_conv5a),
               RO_L(enc_conv5b), RO_L(dec_conv4a), RO_L(dec_conv4b), RO_L(dec_conv3a), RO_L(dec_conv3b), RO_L(dec_conv2a),
               RO_L(dec_conv2b), RO_L(dec_conv1a), RO_L(dec_conv1b), RO_L(dec_conv0)};
#undef RO_L
    *weights = L[i].w;
    *weights_count = L[i].wn;
    *bias = L[i].b;
    *bias_count = L[i].bn;
}

} // extern "C"
