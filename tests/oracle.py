"""TEST INFRASTRUCTURE: ctypes wrapper over oracle/_ref/libray_oracle.so (the unmodified reference + oracle/ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
It implements the same scene-building verbs as the product's host layer so ray_b200.scenes.build() can drive either.
"""
import ctypes as C
import os

import numpy as np

from ray_b200 import capi
from ray_b200.cuda import HIT_DTYPE, RAY_DTYPE, SHADOW_DTYPE

P_u16 = C.POINTER(C.c_uint16)
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "oracle", "_ref", "libray_oracle.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C oracle` (needs /root/reference)")
    lib = C.CDLL(LIB_PATH)
    vp, P = C.c_void_p, C.POINTER
    sig = {
        "ro_pmj_table": (P(C.c_uint32), [P(C.c_int), P(C.c_int)]),
        "ro_error_count": (C.c_int, []),
        "ro_set_verbose": (None, [C.c_int]),
        "ro_cpu_features": (C.c_int, []),
        "ro_scene_create": (vp, [C.c_int]),
        "ro_scene_create_ex": (vp, [C.c_int, C.c_int]),
        "ro_scene_destroy": (None, [vp]),
        "ro_add_texture": (C.c_uint32, [vp, P(capi.rs_tex_desc)]),
        "ro_add_material_node": (C.c_uint32, [vp, P(capi.rs_shading_node_desc)]),
        "ro_add_material_principled": (C.c_uint32, [vp, P(capi.rs_principled_mat_desc)]),
        "ro_add_mesh": (C.c_uint32, [vp, P(capi.rs_mesh_desc)]),
        "ro_add_mesh_instance": (C.c_uint32, [vp, P(capi.rs_mesh_instance_desc)]),
        "ro_add_light_directional": (C.c_uint32, [vp, P(capi.rs_directional_light_desc)]),
        "ro_add_light_sphere": (C.c_uint32, [vp, P(capi.rs_sphere_light_desc)]),
        "ro_add_light_spot": (C.c_uint32, [vp, P(capi.rs_spot_light_desc)]),
        "ro_add_light_rect": (C.c_uint32, [vp, P(capi.rs_rect_light_desc)]),
        "ro_add_light_disk": (C.c_uint32, [vp, P(capi.rs_disk_light_desc)]),
        "ro_add_light_line": (C.c_uint32, [vp, P(capi.rs_line_light_desc)]),
        "ro_set_environment": (None, [vp, P(capi.rs_environment_desc)]),
        "ro_add_camera": (C.c_uint32, [vp, P(capi.rs_camera_desc)]),
        "ro_finalize": (None, [vp]),
        "ro_scene_view": (None, [vp, P(capi.rc_scene_view)]),
        "ro_scene_count": (C.c_uint32, [vp, C.c_int]),
        "ro_get_camera": (None, [vp, P(capi.rc_camera)]),
        "ro_get_filter_table": (None, [vp, vp]),
        "ro_renderer_create": (vp, [C.c_int, C.c_int, C.c_int]),
        "ro_renderer_destroy": (None, [vp]),
        "ro_renderer_clear": (None, [vp, P(C.c_float)]),
        "ro_render": (None, [vp, vp, P(capi.rc_rect), P(C.c_int)]),
        "ro_denoise": (None, [vp, P(capi.rc_rect), C.c_int]),
        "ro_denoise_unet": (C.c_int, [vp, P(capi.rc_rect), C.c_int]),
        "ro_view_lut": (P(C.c_uint32), [C.c_int]),
        "ro_unet_layer": (None, [C.c_int, P(P(C.c_uint16)), P(C.c_int), P(P(C.c_uint16)), P(C.c_int)]),
        "ro_get_pixels": (P(C.c_float), [vp, C.c_int, P(C.c_int)]),
        "ro_get_stats": (None, [vp, P(C.c_uint64)]),
        "ro_render_mt": (C.c_double, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "ro_stage_generate_primary_rays": (C.c_int, [vp, C.c_int, C.c_int, P(capi.rc_rect), C.c_int, vp, vp]),
        "ro_stage_trace_rays": (None, [vp, C.c_int, vp, vp, C.c_int, C.c_int]),
        "ro_stage_shade": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, P(C.c_int), vp,
                                  P(C.c_int), vp, vp, vp]),
        "ro_stage_trace_shadow_rays": (None, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_float, vp]),
        "ro_view_trace_rays": (None, [P(capi.rc_scene_view), P(capi.rc_camera), C.c_int, vp, vp, C.c_int, C.c_int]),
        "ro_view_shade": (None, [P(capi.rc_scene_view), P(capi.rc_camera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 vp, vp, C.c_int, vp, P(C.c_int), vp, P(C.c_int), vp, vp, vp]),
        "ro_view_trace_shadow_rays": (None, [P(capi.rc_scene_view), P(capi.rc_camera), C.c_int, C.c_int, vp, C.c_int,
                                             C.c_float, vp]),
        "ro_view_render_sample": (None, [P(capi.rc_scene_view), P(capi.rc_camera), vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                         vp, P(C.c_ulonglong)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def view_lut(view_transform):
    """The reference's 48^3 packed table of an AgX / Filmic view transform (for rc_set_view_lut)."""
    ptr = load().ro_view_lut(int(view_transform))
    assert ptr, view_transform
    return np.ctypeslib.as_array(ptr, shape=(48 ** 3,)).copy()


def unet_layers():
    """The reference's UNet weight set as 16 x (weights fp16 [cout, cin, 3, 3], bias fp16 [cout]) in pass order."""
    lib = load()
    out = []
    for i in range(16):
        w, b = P_u16(), P_u16()
        wn, bn = C.c_int(), C.c_int()
        lib.ro_unet_layer(i, C.byref(w), C.byref(wn), C.byref(b), C.byref(bn))
        bias = np.ctypeslib.as_array(b, shape=(bn.value,)).view(np.float16).copy()
        wts = np.ctypeslib.as_array(w, shape=(wn.value,)).view(np.float16).copy()
        cout = bn.value
        out.append((wts.reshape(cout, wn.value // (9 * cout), 3, 3), bias))
    return out


def pmj_table():
    lib = load()
    d, s = C.c_int(), C.c_int()
    p = lib.ro_pmj_table(C.byref(d), C.byref(s))
    return np.ctypeslib.as_array(p, shape=(d.value * s.value * 2,)).copy()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """The reference's Cpu::Scene(use_wide_bvh) behind the scene-building verbs of ray_b200.scenes.build()."""

    def __init__(self, wide=True, tex_compression=False):
        self.lib = load()
        self.h = self.lib.ro_scene_create_ex(1 if wide else 0, 1 if tex_compression else 0)
        self._keep = []

    def close(self):
        if self.h:
            self.lib.ro_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_environment(self, env_col, back_col, importance_sample=True, env_map=capi.RS_INVALID,
                        back_map=capi.RS_INVALID, env_map_rotation=0.0, back_map_rotation=0.0):
        d = capi.rs_environment_desc(env_col=tuple(env_col), back_col=tuple(back_col),
                                     importance_sample=1 if importance_sample else 0, env_map=env_map,
                                     back_map=back_map, env_map_rotation=env_map_rotation,
                                     back_map_rotation=back_map_rotation)
        self.lib.ro_set_environment(self.h, C.byref(d))

    def add_texture(self, pixels, is_srgb=True, is_normalmap=False, generate_mipmaps=False, reconstruct_z=False,
                    convention=0):
        d, keep = capi.make_tex_desc(pixels, is_srgb, is_normalmap, generate_mipmaps, reconstruct_z, convention)
        return self.lib.ro_add_texture(self.h, C.byref(d))

    def add_material_node(self, d):
        return self.lib.ro_add_material_node(self.h, C.byref(d))

    def add_material_principled(self, d):
        return self.lib.ro_add_material_principled(self.h, C.byref(d))

    def add_mesh(self, attrs, indices, groups, allow_spatial_splits=False, use_fast_bvh_build=False):
        attrs = np.ascontiguousarray(attrs, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        fp = attrs.ctypes.data_as(C.POINTER(C.c_float))
        m = capi.rs_mesh_desc()
        m.vtx_positions = capi.rs_vtx_attribute(fp, attrs.size, 0, 8)
        m.vtx_normals = capi.rs_vtx_attribute(fp, attrs.size, 3, 8)
        m.vtx_binormals = capi.rs_vtx_attribute(None, 0, 0, 0)
        m.vtx_uvs = capi.rs_vtx_attribute(fp, attrs.size, 6, 8)
        m.vtx_indices = indices.ctypes.data_as(C.POINTER(C.c_uint32))
        m.vtx_indices_count = len(indices)
        m.base_vertex = 0
        garr = (capi.rs_mat_group_desc * len(groups))(*[capi.rs_mat_group_desc(*g) for g in groups])
        m.groups = garr
        m.groups_count = len(groups)
        m.allow_spatial_splits = 1 if allow_spatial_splits else 0
        m.use_fast_bvh_build = 1 if use_fast_bvh_build else 0
        return self.lib.ro_add_mesh(self.h, C.byref(m))

    def add_mesh_instance(self, mesh, xform, camera_visibility=True, diffuse_visibility=True, specular_visibility=True,
                          refraction_visibility=True, shadow_visibility=True):
        d = capi.rs_mesh_instance_desc(xform=tuple(float(x) for x in xform), mesh=mesh,
                                       camera_visibility=int(camera_visibility),
                                       diffuse_visibility=int(diffuse_visibility),
                                       specular_visibility=int(specular_visibility),
                                       refraction_visibility=int(refraction_visibility),
                                       shadow_visibility=int(shadow_visibility))
        return self.lib.ro_add_mesh_instance(self.h, C.byref(d))

    def add_light(self, kind, d):
        return getattr(self.lib, f"ro_add_light_{kind}")(self.h, C.byref(d))

    def add_camera(self, d):
        return self.lib.ro_add_camera(self.h, C.byref(d))

    def finalize(self):
        self.lib.ro_finalize(self.h)

    # ---- views for the CUDA backend ----
    def view(self):
        v = capi.rc_scene_view()
        self.lib.ro_scene_view(self.h, C.byref(v))
        return v

    def camera(self):
        c = capi.rc_camera()
        self.lib.ro_get_camera(self.h, C.byref(c))
        return c

    def filter_table(self):
        t = np.zeros(1024, dtype=np.float32)
        self.lib.ro_get_filter_table(self.h, _ptr(t))
        return t

    def count(self, which):
        return self.lib.ro_scene_count(self.h, which)

    # ---- Ref:: stage functions ----
    def generate_primary_rays(self, w, h, rect, iteration):
        r = capi.rc_rect(*rect)
        rays = np.zeros(r.w * r.h, dtype=RAY_DTYPE)
        hits = np.zeros(r.w * r.h, dtype=HIT_DTYPE)
        n = self.lib.ro_stage_generate_primary_rays(self.h, w, h, C.byref(r), iteration, _ptr(rays), _ptr(hits))
        return rays[:n], hits[:n]

    def trace_rays(self, iteration, rays, hits, trace_lights):
        rays = np.ascontiguousarray(rays.copy())
        hits = np.ascontiguousarray(hits.copy())
        self.lib.ro_stage_trace_rays(self.h, iteration, _ptr(rays), _ptr(hits), len(rays), 1 if trace_lights else 0)
        return rays, hits

    def shade(self, w, h, iteration, primary, bounce, rays, hits, temp, base_color=None, depth_normals=None):
        rays = np.ascontiguousarray(rays)
        hits = np.ascontiguousarray(hits)
        n = len(rays)
        sec = np.zeros(n + 1, dtype=RAY_DTYPE)
        sh = np.zeros(n + 1, dtype=SHADOW_DTYPE)
        ns, nh = C.c_int(0), C.c_int(0)
        base_color = np.zeros((h, w, 4), np.float32) if base_color is None else base_color
        depth_normals = np.zeros((h, w, 4), np.float32) if depth_normals is None else depth_normals
        self.lib.ro_stage_shade(self.h, w, h, iteration, 1 if primary else 0, bounce, _ptr(rays), _ptr(hits), n,
                                _ptr(sec), C.byref(ns), _ptr(sh), C.byref(nh), _ptr(temp), _ptr(base_color),
                                _ptr(depth_normals))
        return sec[:ns.value].copy(), sh[:nh.value].copy(), base_color, depth_normals

    def trace_shadow_rays(self, w, iteration, shadow_rays, clamp_val, temp):
        shadow_rays = np.ascontiguousarray(shadow_rays)
        self.lib.ro_stage_trace_shadow_rays(self.h, w, iteration, _ptr(shadow_rays), len(shadow_rays),
                                            float(clamp_val), _ptr(temp))


class ViewScene:
    """The reference's Ref:: stage functions run over CALLER-PROVIDED scene arrays (an rc_scene_view + rc_camera), e.g.
    the arrays built by the product's host layer: validates those builders with the reference's own code, on CPU."""

    def __init__(self, view, cam):
        self.lib = load()
        self.view, self.cam = view, cam

    def trace_rays(self, iteration, rays, hits, trace_lights):
        rays = np.ascontiguousarray(rays.copy())
        hits = np.ascontiguousarray(hits.copy())
        self.lib.ro_view_trace_rays(C.byref(self.view), C.byref(self.cam), iteration, _ptr(rays), _ptr(hits), len(rays),
                                    1 if trace_lights else 0)
        return rays, hits

    def shade(self, w, h, iteration, primary, bounce, rays, hits, temp, base_color=None, depth_normals=None):
        rays = np.ascontiguousarray(rays)
        hits = np.ascontiguousarray(hits)
        n = len(rays)
        sec = np.zeros(n + 1, dtype=RAY_DTYPE)
        sh = np.zeros(n + 1, dtype=SHADOW_DTYPE)
        ns, nh = C.c_int(0), C.c_int(0)
        base_color = np.zeros((h, w, 4), np.float32) if base_color is None else base_color
        depth_normals = np.zeros((h, w, 4), np.float32) if depth_normals is None else depth_normals
        self.lib.ro_view_shade(C.byref(self.view), C.byref(self.cam), w, h, iteration, 1 if primary else 0, bounce,
                               _ptr(rays), _ptr(hits), n, _ptr(sec), C.byref(ns), _ptr(sh), C.byref(nh), _ptr(temp),
                               _ptr(base_color), _ptr(depth_normals))
        return sec[:ns.value].copy(), sh[:nh.value].copy(), base_color, depth_normals

    def trace_shadow_rays(self, w, iteration, shadow_rays, clamp_val, temp):
        shadow_rays = np.ascontiguousarray(shadow_rays)
        self.lib.ro_view_trace_shadow_rays(C.byref(self.view), C.byref(self.cam), w, iteration, _ptr(shadow_rays),
                                           len(shadow_rays), float(clamp_val), _ptr(temp))


def host_threads():
    """Threads this process may actually use: the affinity mask, capped by the cgroup cpu quota when there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def view_render(view, cam, cam_scene, w, h, spp, threads=None, first_iteration=1):
    """`spp` samples of the whole RenderScene sequence through the reference's Ref:: stage functions over caller-provided
    scene arrays (multi-threaded over row strips), accumulated like RendererCPU.h:607-633 (exposure 0: full += (temp -
    full) / iteration, in float32).  Returns (full image (h, w, 4) float32, closest-hit rays traced, shadow rays traced)."""
    lib = load()
    threads = threads or host_threads()
    full = np.zeros((h, w, 4), np.float32)
    counts = (C.c_ulonglong * 2)(0, 0)
    for it in range(first_iteration, first_iteration + spp):
        temp = np.zeros((h, w, 4), np.float32)
        lib.ro_view_render_sample(C.byref(view), C.byref(cam), cam_scene.h, w, h, it, threads, _ptr(temp), counts)
        full += (temp - full) * (np.float32(1.0) / np.float32(it))
    return full, int(counts[0]), int(counts[1])


def render_with_stages(sc, gen_scene, w, h, spp, max_bounces=8):
    """One full RenderScene sequence per sample (primary + bounces) through stage functions of `sc` (a Scene or a
    ViewScene); primary rays come from `gen_scene` (a Scene: ray generation only needs the camera). Returns the mean
    radiance image (h, w, 4) and the number of closest-hit rays traced."""
    acc = np.zeros((h, w, 4), np.float64)
    n_rays = 0
    cam = gen_scene.camera()
    for it in range(1, spp + 1):
        temp = np.zeros((h, w, 4), np.float32)
        rays, hits = gen_scene.generate_primary_rays(w, h, (0, 0, w, h), it)
        rays, hits = sc.trace_rays(it, rays, hits, False)
        n_rays += len(rays)
        sec, sh, _, _ = sc.shade(w, h, it, True, 0, rays, hits, temp)
        sc.trace_shadow_rays(w, it, sh, cam.clamp_direct, temp)
        for bounce in range(1, min(max_bounces, cam.max_total_depth) + 1):
            if len(sec) == 0:
                break
            hits0 = np.zeros(len(sec), dtype=HIT_DTYPE)
            hits0["obj_index"] = -1
            hits0["prim_index"] = -1
            hits0["t"] = np.float32(3.402823466e+30)
            hits0["v"] = -1.0
            n_rays += len(sec)
            r2, h2 = sc.trace_rays(it, sec, hits0, True)
            sec, sh, _, _ = sc.shade(w, h, it, False, bounce, r2, h2, temp)
            sc.trace_shadow_rays(w, it, sh, cam.clamp_indirect, temp)
        acc += temp
    return (acc / spp).astype(np.float32), n_rays


class Renderer:
    """One of the reference's CPU renderers (Reference, SSE41, AVX, AVX2, AVX512)."""

    def __init__(self, rtype, w, h):
        self.lib = load()
        self.h = self.lib.ro_renderer_create(rtype, w, h)
        if not self.h:
            raise RuntimeError(f"reference renderer type {rtype} is not available on this CPU")
        self.w, self.hh = w, h

    def close(self):
        if self.h:
            self.lib.ro_renderer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self, rgba=(0, 0, 0, 0)):
        self.lib.ro_renderer_clear(self.h, (C.c_float * 4)(*rgba))

    def render(self, scene, rect, iteration):
        """One RenderScene call; `iteration` is RegionContext::iteration before the call; returns it after."""
        r = capi.rc_rect(*rect)
        it = C.c_int(iteration)
        self.lib.ro_render(self.h, scene.h, C.byref(r), C.byref(it))
        return it.value

    def denoise_unet(self, rect, iteration):
        """RendererBase::InitUNetFilter + DenoiseImage(pass, region) for all 16 passes (the UNet overload)."""
        r = capi.rc_rect(*rect)
        return self.lib.ro_denoise_unet(self.h, C.byref(r), int(iteration))

    def denoise(self, rect, iteration):
        """RendererBase::DenoiseImage(region) (NLM)."""
        r = capi.rc_rect(*rect)
        self.lib.ro_denoise(self.h, C.byref(r), int(iteration))

    def pixels(self, which=1):
        pitch = C.c_int(0)
        p = self.lib.ro_get_pixels(self.h, which, C.byref(pitch))
        return np.ctypeslib.as_array(p, shape=(self.hh, pitch.value, 4))[:, :self.w, :].copy()

    def stats_us(self):
        a = (C.c_uint64 * 11)()
        self.lib.ro_get_stats(self.h, a)
        return list(a)

    def render_mt(self, scene, spp, threads, tile=64):
        return self.lib.ro_render_mt(self.h, scene.h, self.w, self.hh, spp, threads, tile)
