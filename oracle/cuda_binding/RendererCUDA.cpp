// RendererCUDA.cpp -- the binding INTEGRATION.md describes, compiled for real against the reference's own headers:
//   Ray::Cuda::Scene    : Ray::Cpu::Scene   (the reference's SAH builder, storage and light tree are reused as they are;
//                                            the subclass only hands the arrays to rc_upload_scene)
//   Ray::Cuda::Renderer : Ray::RendererBase (every call goes through the C-ABI of libray_cuda.so, include/ray_cuda.h)
// TEST INFRASTRUCTURE (lives under oracle/, built by oracle/Makefile `ref_tests`): it exists so that the reference's
// own regression suite (tests/test_shading.cpp, tests/test_aux_channels.cpp) can be run with `--arch CUDA` against the
// CUDA backend and its committed ref.tga gates.  Inside the reference tree a maintainer would drop this file in as
// internal/RendererCUDA.cpp.  Nothing here is linked into the product libraries.
#include <cstring>
#include <list>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "Ray.h"
#include "internal/CDFUtils.h"
#include "internal/Core.h"
#include "internal/SceneCPU.h"

#include "../../include/ray_cuda.h"

namespace Ray {
namespace oidn_hdr_alb_nrm { // the weight set of the reference's UNet filter (UNetFilter.cpp:13-15), from where it lies
#include "internal/precomputed/__oidn_weights_hdr_alb_nrm.inl"
}
extern const uint32_t __pmj02_samples[]; // internal/precomputed/__pmj02_samples.inl through Core.cpp
namespace Cuda {
// the enumerator a maintainer appends to eRendererType (RendererBase.h:22-34); the unmodified header ends at DirectX12 = 7
static const eRendererType kTypeCUDA = eRendererType(8);
} // namespace Cuda
extern const uint32_t *transform_luts[]; // TonemapRef.cpp:15
namespace Cuda {

class Scene final : public Cpu::Scene {
    friend class Renderer;
    uint64_t revision_ = 1;
    std::vector<uint32_t> tex_handles_;
    mutable std::list<std::vector<uint8_t>> tex_pixels_;
    mutable std::vector<rc_texture> tex_views_;
    mutable std::vector<material_t> live_materials_;
    mutable std::vector<light_t> live_lights_;

  public:
    // texture compression as the settings ask, like the CPU backends (RendererCPU.h: CreateScene): BCn blocks and YCoCg
    // base-colour maps are decoded through TexStorage::Fetch / YCoCg_to_RGB on the device
    Scene(ILog *log, bool use_tex_compression)
        : Cpu::Scene(log, true /* wide BVH */, use_tex_compression, false /* spatial cache */) {}

    TextureHandle AddTexture(const tex_desc_t &t) override {
        const TextureHandle h = Cpu::Scene::AddTexture(t);
        tex_handles_.push_back(h._index);
        return h;
    }

    void Finalize(const std::function<void(int, int, ParallelForFunction &&)> &parallel_for) override {
        Cpu::Scene::Finalize(parallel_for);
        ++revision_;
    }

    const camera_t &current_camera() const { return cams_[current_cam_._index]; }

    // SparseStorage::data() handed over as it is (all indices in the arrays are absolute offsets)
    void FillView(rc_scene_view &v) const {
        memset(&v, 0, sizeof(v));
        v.wnodes = {wnodes_.data(), wnodes_.capacity(), sizeof(wbvh_node_t)};
        v.mtris = {mtris_.data(), mtris_.capacity(), sizeof(mtri_accel_t)};
        v.tri_indices = {tri_indices_.data(), tri_indices_.capacity(), sizeof(uint32_t)};
        v.tri_materials = {tri_materials_.data(), tri_materials_.capacity(), sizeof(tri_mat_data_t)};
        // the live range of a SparseStorage is not a prefix and freed slots keep stale bytes: hand over a copy in which every
        // slot that is not live carries an invalid node type, which rc_upload_scene skips when it resolves texture handles
        live_materials_.assign(materials_.capacity(), material_t{});
        for (material_t &m : live_materials_) {
            memset(&m, 0xff, sizeof(m));
        }
        for (auto it = materials_.cbegin(); it != materials_.cend(); ++it) {
            live_materials_[it.index()] = *it;
        }
        v.materials = {live_materials_.data(), uint32_t(live_materials_.size()), sizeof(material_t)};
        v.mesh_instances = {mesh_instances_.data(), mesh_instances_.capacity(), sizeof(mesh_instance_t)};
        v.vertices = {vertices_.data(), vertices_.capacity(), sizeof(vertex_t)};
        v.vtx_indices = {vtx_indices_.data(), vtx_indices_.capacity(), sizeof(uint32_t)};
        live_lights_.assign(lights_.capacity(), light_t{});
        for (light_t &l : live_lights_) {
            memset(&l, 0, sizeof(l));
        }
        for (auto it = lights_.cbegin(); it != lights_.cend(); ++it) {
            live_lights_[it.index()] = *it;
        }
        v.lights = {live_lights_.data(), uint32_t(live_lights_.size()), sizeof(light_t)};
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
        v.env_map_rotation = env_.env_map_rotation;
        v.back_map_rotation = env_.back_map_rotation;
        v.qtree_levels = env_.qtree_levels;
        for (int i = 0; i < 16; ++i) {
            v.qtree_mips[i] = (i < env_.qtree_levels) ? env_.qtree_mips[i] : nullptr;
        }
        const_cast<Scene *>(this)->GetBounds(v.bounds_min, v.bounds_max);
        // textures cross the C-ABI decoded: walk the storage's own Fetch() (swizzle / block decode / channel expansion)
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
                    t.pixels[lod] = t.pixels[lod - 1];
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
};

static void flatten(const camera_t &c, rc_camera &out) {
    memset(&out, 0, sizeof(out));
    out.type = uint32_t(c.type);
    out.filter = uint32_t(c.filter);
    out.view_transform = uint32_t(c.view_transform);
    out.fov = c.fov;
    out.exposure = c.exposure;
    out.gamma = c.gamma;
    out.sensor_height = c.sensor_height;
    out.focus_distance = c.focus_distance;
    out.focal_length = c.focal_length;
    out.fstop = c.fstop;
    out.lens_rotation = c.lens_rotation;
    out.lens_ratio = c.lens_ratio;
    out.lens_blades = c.lens_blades;
    out.clip_start = c.clip_start;
    out.clip_end = c.clip_end;
    memcpy(out.origin, c.origin, sizeof(out.origin));
    memcpy(out.fwd, c.fwd, sizeof(out.fwd));
    memcpy(out.side, c.side, sizeof(out.side));
    memcpy(out.up, c.up, sizeof(out.up));
    memcpy(out.shift, c.shift, sizeof(out.shift));
    out.max_diff_depth = c.pass_settings.max_diff_depth;
    out.max_spec_depth = c.pass_settings.max_spec_depth;
    out.max_refr_depth = c.pass_settings.max_refr_depth;
    out.max_transp_depth = c.pass_settings.max_transp_depth;
    out.max_total_depth = c.pass_settings.max_total_depth;
    out.min_total_depth = c.pass_settings.min_total_depth;
    out.min_transp_depth = c.pass_settings.min_transp_depth;
    out.clamp_direct = c.pass_settings.clamp_direct;
    out.clamp_indirect = c.pass_settings.clamp_indirect;
    out.min_samples = c.pass_settings.min_samples;
    out.variance_threshold = c.pass_settings.variance_threshold;
    out.regularize_alpha = c.pass_settings.regularize_alpha;
}

// Cpu::Renderer<P>::UpdateFilterTable (internal/RendererCPU.h:1234-1258) on the reference's own CDFInverted
static std::vector<float> make_filter_table(ePixelFilter filter, float filter_width) {
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
    default: break;
    }
    return CDFInverted(FILTER_TABLE_SIZE, 0.0f, filter_width * 0.5f,
                       std::bind(filter_func, std::placeholders::_1, filter_width), true);
}

class Renderer final : public RendererBase {
    ILog *log_;
    rc_ctx *ctx_ = nullptr;
    int w_ = 0, h_ = 0;
    bool use_tex_compression_ = true;
    std::string device_name_;
    mutable std::vector<color_rgba_t> final_, raw_, base_, dn_;
    mutable bool final_dirty_ = true, raw_dirty_ = true, base_dirty_ = true, dn_dirty_ = true;
    const Scene *uploaded_ = nullptr;
    uint64_t uploaded_revision_ = 0;
    ePixelFilter table_filter_ = ePixelFilter(-1);
    float table_width_ = 0.0f;
    std::vector<float> filter_table_;

    void Readback(int which, std::vector<color_rgba_t> &dst) const {
        dst.resize(size_t(w_) * h_);
        if (w_ == 0 || h_ == 0) {
            return;
        }
        const rc_rect r{0, 0, w_, h_};
        if (rc_readback(ctx_, which, &r, &dst[0].v[0], w_) != 0) {
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
        }
    }

  public:
    Renderer(const settings_t &s, ILog *log) : log_(log) {
        if (rc_create(0, &ctx_) != 0 || !ctx_) {
            throw std::runtime_error("no usable sm_100 CUDA device");
        }
        device_name_ = rc_device_name(ctx_);
        // AgX / Filmic view transforms: the tree's own tables (TonemapRef.cpp:5-27)
        for (int vt = 1; vt < int(eViewTransform::_Count); ++vt) {
            if (rc_set_view_lut(ctx_, uint32_t(vt), transform_luts[vt], 48) != 0) {
                throw std::runtime_error(rc_last_error(ctx_));
            }
        }
        use_tex_compression_ = s.use_tex_compression;
        Resize(s.w, s.h);
    }
    ~Renderer() override { rc_destroy(ctx_); }

    eRendererType type() const override { return kTypeCUDA; }
    ILog *log() const override { return log_; }
    std::string_view device_name() const override { return device_name_; }
    std::pair<int, int> size() const override { return {w_, h_}; }

    color_data_rgba_t get_pixels_ref() const override {
        if (final_dirty_) {
            Readback(RC_BUF_FINAL, final_);
            final_dirty_ = false;
        }
        return {final_.data(), w_};
    }
    color_data_rgba_t get_raw_pixels_ref() const override {
        if (raw_dirty_) {
            Readback(RC_BUF_RAW, raw_);
            raw_dirty_ = false;
        }
        return {raw_.data(), w_};
    }
    color_data_rgba_t get_aux_pixels_ref(const eAUXBuffer buf) const override {
        if (buf == eAUXBuffer::BaseColor) {
            if (base_dirty_) {
                Readback(RC_BUF_BASE_COLOR, base_);
                base_dirty_ = false;
            }
            return {base_.data(), w_};
        } else if (buf == eAUXBuffer::DepthNormals) {
            if (dn_dirty_) {
                Readback(RC_BUF_DEPTH_NORMALS, dn_);
                dn_dirty_ = false;
            }
            return {dn_.data(), w_};
        }
        return {nullptr, 0};
    }
    const shl1_data_t *get_sh_data_ref() const override { return nullptr; }

    void Resize(const int w, const int h) override {
        if (rc_resize(ctx_, w, h) != 0) {
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
            return;
        }
        w_ = w;
        h_ = h;
        final_dirty_ = raw_dirty_ = base_dirty_ = dn_dirty_ = true;
    }
    void Clear(const color_rgba_t &c) override {
        if (rc_clear(ctx_, c.v) != 0) {
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
        }
    }
    SceneBase *CreateScene() override { return new Scene(log_, use_tex_compression_); }

    void RenderScene(const SceneBase &scene, RegionContext &region) override {
        const auto *sp = dynamic_cast<const Scene *>(&scene); // pattern of RendererCPU.h:377
        if (!sp) {
            log_->Error("Ray(CUDA): RenderScene needs a scene created by this renderer");
            return;
        }
        const Scene &s = *sp;
        std::shared_lock<std::shared_timed_mutex> lock(s.mtx_); // RendererCPU.h:379
        const camera_t &cam = s.current_camera();
        ++region.iteration; // RendererCPU.h:384
        if (cam.filter != table_filter_ || cam.filter_width != table_width_) {
            filter_table_ = make_filter_table(cam.filter, cam.filter_width);
            table_filter_ = cam.filter;
            table_width_ = cam.filter_width;
            if (rc_upload_tables(ctx_, __pmj02_samples, RAND_DIMS_COUNT, RAND_SAMPLES_COUNT, filter_table_.data(),
                                 int(filter_table_.size())) != 0) {
                log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
                return;
            }
        }
        if (uploaded_ != &s || uploaded_revision_ != s.revision_) {
            rc_scene_view v;
            s.FillView(v);
            if (rc_upload_scene(ctx_, &v) != 0) {
                log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
                return;
            }
            uploaded_ = &s;
            uploaded_revision_ = s.revision_;
        }
        rc_pass_desc p;
        memset(&p, 0, sizeof(p));
        flatten(cam, p.cam);
        p.rect = rc_rect{region.rect().x, region.rect().y, region.rect().w, region.rect().h};
        p.iteration = region.iteration;
        if (rc_render(ctx_, &p) != 0) { // blocking, like the CPU backends
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
        }
        final_dirty_ = raw_dirty_ = base_dirty_ = dn_dirty_ = true;
    }

    void DenoiseImage(const RegionContext &region) override { // NLM overload, RendererCPU.h:661-787
        const rc_rect r = {region.rect().x, region.rect().y, region.rect().w, region.rect().h};
        if (rc_denoise_nlm(ctx_, &r, region.iteration) != 0) {
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
        }
        final_dirty_ = raw_dirty_ = true;
    }
    void DenoiseImage(const int pass, const RegionContext &region) override { // UNet overload, RendererCPU.h:790-1007
        const rc_rect r = {region.rect().x, region.rect().y, region.rect().w, region.rect().h};
        if (rc_denoise_unet(ctx_, pass, &r, RC_UNET_TENSOR_CORES) != 0) {
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
        }
        final_dirty_ = raw_dirty_ = true;
    }
    void UpdateSpatialCache(const SceneBase &, RegionContext &) override { log_->Error("Ray(CUDA): no spatial cache"); }
    void ResolveSpatialCache(const SceneBase &, const std::function<void(int, int, ParallelForFunction &&)> &) override {
        log_->Error("Ray(CUDA): no spatial cache");
    }
    void ResetSpatialCache(const SceneBase &, const std::function<void(int, int, ParallelForFunction &&)> &) override {}
    void GetStats(stats_t &st) override {
        uint64_t us[11] = {};
        rc_get_stats(ctx_, us);
        st.time_primary_ray_gen_us = us[0];
        st.time_primary_trace_us = us[1];
        st.time_primary_shade_us = us[2];
        st.time_primary_shadow_us = us[3];
        st.time_secondary_sort_us = us[4];
        st.time_secondary_trace_us = us[5];
        st.time_secondary_shade_us = us[6];
        st.time_secondary_shadow_us = us[7];
        st.time_denoise_us = us[8];
        st.time_cache_update_us = us[9];
        st.time_cache_resolve_us = us[10];
    }
    void ResetStats() override { rc_reset_stats(ctx_); }
    unet_filter_properties_t InitUNetFilter(bool, const std::function<void(int, int, ParallelForFunction &&)> &) override {
        using namespace oidn_hdr_alb_nrm;
#define RC_L(n, ci, co) rc_unet_layer{n##_weight, n##_bias, ci, co}
        const rc_unet_layer layers[16] = {
            RC_L(enc_conv0, 9, 32),     RC_L(enc_conv1, 32, 32),   RC_L(enc_conv2, 32, 48),    RC_L(enc_conv3, 48, 64),
            RC_L(enc_conv4, 64, 80),    RC_L(enc_conv5a, 80, 96),  RC_L(enc_conv5b, 96, 96),   RC_L(dec_conv4a, 160, 112),
            RC_L(dec_conv4b, 112, 112), RC_L(dec_conv3a, 160, 96), RC_L(dec_conv3b, 96, 96),   RC_L(dec_conv2a, 128, 64),
            RC_L(dec_conv2b, 64, 64),   RC_L(dec_conv1a, 73, 64),  RC_L(dec_conv1b, 64, 32),   RC_L(dec_conv0, 32, 3)};
#undef RC_L
        unet_filter_properties_t props = {};
        if (rc_unet_set_weights(ctx_, layers) != 0) {
            log_->Error("Ray(CUDA): %s", rc_last_error(ctx_));
            return props;
        }
        props.pass_count = 16; // UNetFilterPasses; tensors are not aliased on the device: no inter-pass dependencies to report
        for (int i = 0; i < 16; ++i) {
            for (int j = 0; j < 4; ++j) {
                props.alias_dependencies[i][j] = -1;
            }
        }
        return props;
    }
};

RendererBase *CreateRenderer(const settings_t &s, ILog *log) { return new Renderer(s, log); }
} // namespace Cuda
} // namespace Ray
