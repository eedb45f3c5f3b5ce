// SceneCuda.h -- RayB200::Cuda::Scene: host-side scene storage + builders for the CUDA backend.
//
// Role in the reference: Cpu::Scene (internal/SceneCPU.{h,cpp}) which Ray::Cuda::Scene SUBCLASSES inside the reference
// tree.  This standalone version re-implements the subset of Cpu::Scene the hot path needs -- material
// lowering, mesh preprocessing into BVH8 + 8-wide plane-form triangle blocks, instances, analytic + emissive-triangle
// lights, TLAS, the quantised 8-wide light tree -- emitting the SAME array layouts (rt_types.h == reference Core.h),
// so Cuda::Renderer feeds rc_upload_scene exactly as it would from a real Cpu::Scene.
// Not supported (logged as errors, like any backend missing a feature): textures, env maps, physical sky.
#pragma once

#include <mutex>
#include <shared_mutex>
#include <vector>

#include "../../../include/ray_cuda.h"
#include "../rt_types.h"
#include "BvhBuilder.h"
#include "RayB200.h"

namespace RayB200 {
namespace Cuda {

struct camera_t { // reference Types.h:102-115
    rc_camera rc;          // everything the device needs, already flattened
    rs_camera_desc desc;   // what the user set (GetCamera)
};

class Scene final : public SceneBase {
    friend class Renderer;
    mutable std::shared_timed_mutex mtx_;

    struct TexImage { // one texture of one of the four uncompressed storages (SceneCPU.h:65-68), no mips
        uint32_t handle;   // (storage << 28) | index
        uint32_t channels; // 4, 3, 2, 1 for storages 0..3
        int w, h;
        std::vector<uint8_t> pixels;
    };
    std::vector<TexImage> textures_;
    uint32_t tex_storage_counts_[4] = {0, 0, 0, 0};
    std::vector<rc_texture> tex_views_; // rebuilt by AddTexture / Finalize (unique lock)
    // Page-locked copies of the big geometry arrays, refreshed by Finalize: FillView hands these to rc_upload_scene so a
    // (re-)upload is a straight DMA at PCIe speed instead of the driver staging pageable vectors through a bounce buffer.
    struct PinnedMirror {
        void *ptr = nullptr;
        size_t capacity = 0, bytes = 0;
    };
    enum { PM_WNODES = 0, PM_MTRIS, PM_VERTICES, PM_VTX_INDICES, PM_TRI_INDICES, PM_TRI_MATERIALS, PM_COUNT };
    PinnedMirror pinned_[PM_COUNT];
    uint64_t pinned_revision_ = 0;
    void RefreshPinnedMirrors_nolock();
    // importance-sampling quad-tree of the environment map (reference env_map_qtree_, SceneCPU.h / SceneCPU.cpp:1058-1211)
    std::vector<std::vector<float>> env_qtree_mips_; // level i: 4^(levels-1-i) quads x 4 floats
    void PrepareEnvMapQTree_nolock();
    const TexImage *FindTexture(uint32_t handle) const;
    std::vector<rt::Material> materials_;
    std::vector<rt::Vertex> vertices_;
    std::vector<uint32_t> vtx_indices_;
    std::vector<rt::TriMat> tri_materials_;
    std::vector<uint32_t> tri_indices_;
    std::vector<rt::MTri> mtris_;
    std::vector<rt::WNode> wnodes_;
    uint32_t blas_nodes_end_ = 0; // wnodes_[0, blas_nodes_end_) are BLAS nodes; the TLAS is appended by Finalize

    struct MeshRec {
        Aabb box;
        uint32_t node_index;
        uint32_t tri_first, tri_count; // global triangle ids
        bool alive;
    };
    std::vector<MeshRec> meshes_;
    std::vector<rt::MeshInstance> mesh_instances_;
    std::vector<uint8_t> instance_alive_;

    std::vector<rt::Light> lights_;
    std::vector<uint8_t> light_alive_;
    std::vector<uint32_t> li_indices_;
    std::vector<rt::LightCWNode> light_cwnodes_;
    uint32_t visible_lights_count_ = 0, blocker_lights_count_ = 0;
    uint32_t env_light_index_ = 0xffffffffu;

    std::vector<camera_t> cams_;
    CameraHandle current_cam_;
    environment_desc_t env_{};
    uint32_t tlas_root_ = 0xffffffffu;
    float bounds_min_[3] = {0, 0, 0}, bounds_max_[3] = {0, 0, 0};
    // process-wide unique, so (scene address, revision) never repeats for a new scene at a recycled address
    static uint64_t NextRevision();
    mutable uint64_t revision_ = NextRevision(); // renewed by Finalize: tells the renderer to re-upload
    // revision at which geometry / materials / textures / environment / the instance set last changed: a Finalize that
    // only follows SetMeshInstanceTransform or light edits keeps it, and the renderer refreshes the top level only
    uint64_t structure_revision_ = 0;
    bool structure_dirty_ = true;

    rc_ctx *build_ctx_ = nullptr;

    void RebuildTLAS_nolock();
    void RebuildLightTree_nolock();
    MaterialHandle AddMaterial_nolock(const shading_node_desc_t &m);
    uint32_t AddLight_nolock(const rt::Light &l);
    void RemoveMeshInstance_nolock(uint32_t index);
    void RebuildTexViews_nolock();

  public:
    // build_ctx: device context the fast (LBVH) mesh build runs on (mesh_desc_t::use_fast_bvh_build)
    explicit Scene(ILog *log, rc_ctx *build_ctx = nullptr);
    ~Scene() override;

    void GetEnvironment(environment_desc_t &env) override;
    void SetEnvironment(const environment_desc_t &env) override;
    TextureHandle AddTexture(const tex_desc_t &t) override;
    void RemoveTexture(TextureHandle) override {}
    MaterialHandle AddMaterial(const shading_node_desc_t &m) override;
    MaterialHandle AddMaterial(const principled_mat_desc_t &m) override;
    void RemoveMaterial(MaterialHandle) override {}
    MeshHandle AddMesh(const mesh_desc_t &m) override;
    void RemoveMesh(MeshHandle m) override;
    LightHandle AddLight(const directional_light_desc_t &l) override;
    LightHandle AddLight(const sphere_light_desc_t &l) override;
    LightHandle AddLight(const spot_light_desc_t &l) override;
    LightHandle AddLight(const rect_light_desc_t &l) override;
    LightHandle AddLight(const disk_light_desc_t &l) override;
    LightHandle AddLight(const line_light_desc_t &l) override;
    void RemoveLight(LightHandle l) override;
    MeshInstanceHandle AddMeshInstance(const mesh_instance_desc_t &mi) override;
    void SetMeshInstanceTransform(MeshInstanceHandle mi, const float *xform) override;
    void RemoveMeshInstance(MeshInstanceHandle mi) override;
    void Finalize(const ParallelFor &parallel_for = parallel_for_serial) override;
    CameraHandle AddCamera(const camera_desc_t &c) override;
    void GetCamera(CameraHandle i, camera_desc_t &c) const override;
    void SetCamera(CameraHandle i, const camera_desc_t &c) override;
    void RemoveCamera(CameraHandle) override {}
    CameraHandle current_cam() const override { return current_cam_; }
    void set_current_cam(CameraHandle i) override { current_cam_ = i; }
    uint32_t triangle_count() const override { return uint32_t(tri_materials_.size()); }
    uint32_t node_count() const override { return uint32_t(wnodes_.size()); }

    // what Cuda::Renderer hands to rc_upload_scene
    void FillView(rc_scene_view &v) const;
    uint64_t revision() const { return revision_; }
    uint64_t structure_revision() const { return structure_revision_; }
    uint32_t first_tlas_node() const { return blas_nodes_end_; }
    bool GetDeviceCamera(rc_camera &out) const; // the flattened camera_t RenderScene passes to rc_render
    void GetBounds(float bbox_min[3], float bbox_max[3]) const;
};

void InverseMatrix4(const float m[16], float out[16]);

} // namespace Cuda
} // namespace RayB200
