"""CPU: the C-ABI libraries load, export every symbol the headers declare, and agree with the ctypes struct mirrors."""
import ctypes as C
import os
import re
import subprocess

import pytest

from ray_b200 import capi, cuda, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(r[ch]_[a-z0-9_]+)\s*\(", txt)))


def _exported(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


def test_ray_cuda_exports_every_declared_symbol():
    decl = _declared("ray_cuda.h")
    exp = _exported(cuda.LIB_PATH)
    missing = [s for s in decl if s not in exp]
    assert not missing, f"libray_cuda.so does not export {missing}"
    assert set(cuda.EXPORTED_SYMBOLS) == set(decl), set(cuda.EXPORTED_SYMBOLS) ^ set(decl)
    cuda.load_library()


def test_ray_host_exports_every_declared_symbol():
    decl = _declared("ray_host.h")
    exp = _exported(host.LIB_PATH)
    missing = [s for s in decl if s not in exp]
    assert not missing, f"libray_host.so does not export {missing}"
    assert set(host.EXPORTED_SYMBOLS) == set(decl), set(host.EXPORTED_SYMBOLS) ^ set(decl)
    host.load_library()


def test_struct_sizes_match_ctypes_mirrors():
    lc, lh = cuda.load_library(), host.load_library()
    for i, t in enumerate([capi.rc_array, capi.rc_scene_view, capi.rc_camera, capi.rc_rect, capi.rc_pass_desc,
                           capi.rc_counters, capi.rc_texture]):
        assert lc.rc_abi_sizeof(i) == C.sizeof(t), t.__name__
    for i, t in enumerate([capi.rs_shading_node_desc, capi.rs_principled_mat_desc, capi.rs_mat_group_desc,
                           capi.rs_vtx_attribute, capi.rs_mesh_desc, capi.rs_mesh_instance_desc, capi.rs_light_common,
                           capi.rs_directional_light_desc, capi.rs_sphere_light_desc, capi.rs_spot_light_desc,
                           capi.rs_rect_light_desc, capi.rs_disk_light_desc, capi.rs_line_light_desc,
                           capi.rs_camera_desc, capi.rs_environment_desc]):
        assert lh.rh_abi_sizeof(i) == C.sizeof(t), t.__name__


def test_no_cpu_fallback_without_a_device():
    """On a box without an sm_100 GPU the product must fail loudly, not fall back to anything."""
    lib = cuda.load_library()
    if lib.rc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cuda.CudaError):
        cuda.Context(0)
    with pytest.raises(host.HostError):
        host.Renderer(16, 16)


def test_product_does_not_link_or_mention_the_oracle():
    for lib in (cuda.LIB_PATH, host.LIB_PATH):
        deps = subprocess.check_output(["ldd", lib], text=True)
        assert "oracle" not in deps and "libray_ref" not in deps
    for base, _, files in os.walk(os.path.join(ROOT, "ray_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in txt and "libray_oracle" not in txt and "ref_harness" not in txt, f
