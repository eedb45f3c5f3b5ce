// RayB200.h -- public C++ API of the standalone host layer (libray_host.so).
//
// Mirrors, name for name, the part of the reference's public API that a user of the wavefront path touches
// (reference Ray.h, RendererBase.h, SceneBase.h, Types.h, Log.h) with ONE addition: eRendererType::CUDA.
// It lives in namespace RayB200 so a process may hold this library and the reference side by side.
//
//   Ray::CreateRenderer (Ray.cpp:53-133)          -> RayB200::CreateRenderer   (CUDA first, no CPU fallback)
//   Ray::RendererBase   (RendererBase.h:133-253)  -> RayB200::RendererBase     (same virtuals, same semantics)
//   Ray::SceneBase      (SceneBase.h:371-516)     -> RayB200::SceneBase
//   Ray::RegionContext  (RendererBase.h:78-92)    -> RayB200::RegionContext
//   descriptor structs  (SceneBase.h:44-357)      -> the flat C descriptors of include/ray_scene_desc.h, used directly
//
// In the reference tree the same Cuda::Renderer is added as a backend behind Ray's own headers; see INTEGRATION.md.
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>

#include "../../../include/ray_scene_desc.h"

namespace RayB200 {

enum class eRendererType : uint32_t {
    Reference, SIMD_SSE41, SIMD_AVX, SIMD_AVX2, SIMD_AVX512, SIMD_NEON, Vulkan, DirectX12,
    CUDA // new: sm_100a backend
};
std::string_view RendererTypeName(eRendererType rt);
eRendererType RendererTypeFromName(std::string_view name);
inline bool RendererSupportsMultithreading(eRendererType rt) { return rt != eRendererType::CUDA && rt <= eRendererType::SIMD_NEON; }
inline bool RendererSupportsHWRT(eRendererType) { return false; }

class ILog { // reference Log.h:17-24
  public:
    virtual ~ILog() = default;
    virtual void Info(const char *fmt, ...) = 0;
    virtual void Warning(const char *fmt, ...) = 0;
    virtual void Error(const char *fmt, ...) = 0;
};
class LogNull final : public ILog {
  public:
    void Info(const char *, ...) override {}
    void Warning(const char *, ...) override {}
    void Error(const char *, ...) override {}
};
class LogStdout final : public ILog {
  public:
    void Info(const char *fmt, ...) override;
    void Warning(const char *fmt, ...) override;
    void Error(const char *fmt, ...) override;
};
extern LogNull g_null_log;
extern LogStdout g_stdout_log;

struct settings_t { // reference RendererBase.h:52-63 (Vulkan-only members dropped)
    int w = 0, h = 0;
    std::string_view preferred_device; // "" or a decimal CUDA device index
    bool use_tex_compression = true;
    bool use_hwrt = true;
    bool use_bindless = true;
    bool use_spatial_cache = false;
    int validation_level = 0;
};

struct rect_t {
    int x, y, w, h;
};
struct color_rgba_t {
    float v[4];
};
struct color_data_rgba_t {
    const color_rgba_t *ptr;
    int pitch;
};
enum class eAUXBuffer : uint32_t { SHL1 = 0, BaseColor = 1, DepthNormals = 2 };
struct shl1_data_t {
    float coeff_r[4], coeff_g[4], coeff_b[4];
};
struct unet_filter_properties_t {
    int pass_count = 0;
    int alias_dependencies[16][4] = {};
};

class RegionContext { // reference RendererBase.h:78-92
    rect_t rect_;

  public:
    int iteration = 0;
    int cache_iteration = 0;
    explicit RegionContext(const rect_t &rect) : rect_(rect) {}
    const rect_t &rect() const { return rect_; }
    void Clear() { iteration = 0; }
};

#define RB_HANDLE(name)                                                                                                \
    struct name {                                                                                                      \
        uint32_t _index = 0xffffffffu;                                                                                 \
        uint32_t _block = 0;                                                                                           \
    };                                                                                                                 \
    inline bool operator==(name a, name b) { return a._index == b._index; }                                            \
    inline bool operator!=(name a, name b) { return a._index != b._index; }
RB_HANDLE(CameraHandle)
RB_HANDLE(LightHandle)
RB_HANDLE(MaterialHandle)
RB_HANDLE(MeshHandle)
RB_HANDLE(MeshInstanceHandle)
RB_HANDLE(TextureHandle)
#undef RB_HANDLE

using ParallelForFunction = std::function<void(int)>;
inline void parallel_for_serial(int from, int to, ParallelForFunction &&f) {
    for (int i = from; i < to; ++i) {
        f(i);
    }
}
using ParallelFor = std::function<void(int, int, ParallelForFunction &&)>;

// Descriptors: the flat C structs ARE the descriptors here (field names and defaults are the reference's).
using shading_node_desc_t = rs_shading_node_desc;
using principled_mat_desc_t = rs_principled_mat_desc;
using mesh_desc_t = rs_mesh_desc;
using mesh_instance_desc_t = rs_mesh_instance_desc;
using directional_light_desc_t = rs_directional_light_desc;
using sphere_light_desc_t = rs_sphere_light_desc;
using spot_light_desc_t = rs_spot_light_desc;
using rect_light_desc_t = rs_rect_light_desc;
using disk_light_desc_t = rs_disk_light_desc;
using line_light_desc_t = rs_line_light_desc;
using camera_desc_t = rs_camera_desc;
using environment_desc_t = rs_environment_desc;
using tex_desc_t = rs_tex_desc; // tex_desc_t, SceneBase.h:177-192 (uncompressed formats)

class SceneBase { // reference SceneBase.h:371-516
  protected:
    ILog *log_ = nullptr;

  public:
    virtual ~SceneBase() = default;
    ILog *log() const { return log_; }
    virtual void GetEnvironment(environment_desc_t &env) = 0;
    virtual void SetEnvironment(const environment_desc_t &env) = 0;
    virtual TextureHandle AddTexture(const tex_desc_t &t) = 0;
    virtual void RemoveTexture(TextureHandle t) = 0;
    virtual MaterialHandle AddMaterial(const shading_node_desc_t &m) = 0;
    virtual MaterialHandle AddMaterial(const principled_mat_desc_t &m) = 0;
    virtual void RemoveMaterial(MaterialHandle m) = 0;
    virtual MeshHandle AddMesh(const mesh_desc_t &m) = 0;
    virtual void RemoveMesh(MeshHandle m) = 0;
    virtual LightHandle AddLight(const directional_light_desc_t &l) = 0;
    virtual LightHandle AddLight(const sphere_light_desc_t &l) = 0;
    virtual LightHandle AddLight(const spot_light_desc_t &l) = 0;
    virtual LightHandle AddLight(const rect_light_desc_t &l) = 0; // xform travels inside the flat descriptor
    virtual LightHandle AddLight(const disk_light_desc_t &l) = 0;
    virtual LightHandle AddLight(const line_light_desc_t &l) = 0;
    virtual void RemoveLight(LightHandle l) = 0;
    virtual MeshInstanceHandle AddMeshInstance(const mesh_instance_desc_t &mi) = 0;
    virtual void SetMeshInstanceTransform(MeshInstanceHandle mi, const float *xform) = 0;
    virtual void RemoveMeshInstance(MeshInstanceHandle mi) = 0;
    virtual void Finalize(const ParallelFor &parallel_for = parallel_for_serial) = 0;
    virtual CameraHandle AddCamera(const camera_desc_t &c) = 0;
    virtual void GetCamera(CameraHandle i, camera_desc_t &c) const = 0;
    virtual void SetCamera(CameraHandle i, const camera_desc_t &c) = 0;
    virtual void RemoveCamera(CameraHandle i) = 0;
    virtual CameraHandle current_cam() const = 0;
    virtual void set_current_cam(CameraHandle i) = 0;
    virtual uint32_t triangle_count() const = 0;
    virtual uint32_t node_count() const = 0;
};

class RendererBase { // reference RendererBase.h:133-253
  public:
    virtual ~RendererBase() = default;
    virtual eRendererType type() const = 0;
    virtual ILog *log() const = 0;
    virtual std::string_view device_name() const = 0;
    virtual bool is_hwrt() const { return false; }
    virtual bool is_spatial_caching_enabled() const { return false; }
    virtual std::pair<int, int> size() const = 0;
    virtual color_data_rgba_t get_pixels_ref() const = 0;
    virtual color_data_rgba_t get_raw_pixels_ref() const = 0;
    virtual color_data_rgba_t get_aux_pixels_ref(eAUXBuffer buf) const = 0;
    virtual const shl1_data_t *get_sh_data_ref() const = 0;
    virtual void Resize(int w, int h) = 0;
    virtual void Clear(const color_rgba_t &c) = 0;
    virtual SceneBase *CreateScene() = 0;
    virtual void RenderScene(const SceneBase &scene, RegionContext &region) = 0;
    virtual void DenoiseImage(const RegionContext &region) = 0;
    virtual void DenoiseImage(int pass, const RegionContext &region) = 0;
    virtual void UpdateSpatialCache(const SceneBase &scene, RegionContext &region) = 0;
    virtual void ResolveSpatialCache(const SceneBase &scene, const ParallelFor &parallel_for = parallel_for_serial) = 0;
    virtual void ResetSpatialCache(const SceneBase &scene, const ParallelFor &parallel_for = parallel_for_serial) = 0;
    struct stats_t {
        unsigned long long time_primary_ray_gen_us, time_primary_trace_us, time_primary_shade_us, time_primary_shadow_us,
            time_secondary_sort_us, time_secondary_trace_us, time_secondary_shade_us, time_secondary_shadow_us,
            time_denoise_us, time_cache_update_us, time_cache_resolve_us;
    };
    virtual void GetStats(stats_t &st) = 0;
    virtual void ResetStats() = 0;
    virtual unet_filter_properties_t InitUNetFilter(bool alias_memory,
                                                    const ParallelFor &parallel_for = parallel_for_serial) = 0;
};

constexpr uint32_t DefaultEnabledRenderTypes = 1u << uint32_t(eRendererType::CUDA);

/// Creates a renderer. Only eRendererType::CUDA exists in this library: if it is not enabled, or no sm_100 device can
/// be opened, the call logs the reason and returns nullptr (there is deliberately no CPU fallback here; inside the
/// reference tree the factory falls through to Ray's own CPU backends instead).
RendererBase *CreateRenderer(const settings_t &s, ILog *log = &g_null_log,
                             const ParallelFor &parallel_for = parallel_for_serial,
                             uint32_t enabled_types = DefaultEnabledRenderTypes);
const char *Version();

} // namespace RayB200
