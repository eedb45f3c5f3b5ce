"""CPU: the synthetic workloads of BASELINE.json are deterministic and have the advertised shape."""
import numpy as np

from ray_b200 import capi, scenes


def test_hall_250k_shape():
    d = scenes.hall("diffuse")
    assert d.width == 1920 and d.height == 1080
    assert d.triangle_count() == 253964  # 131,072 + 122,880 + 12 tris
    m = d.meshes[0]
    assert m.attrs.dtype == np.float32 and m.attrs.shape[1] == 8 and m.indices.dtype == np.uint32
    assert int(m.indices.max()) < len(m.attrs)
    covered = sum(c for _, _, _, c in m.groups)
    assert covered == len(m.indices)
    n = np.linalg.norm(m.attrs[:, 3:6], axis=1)
    assert np.abs(n - 1).max() < 1e-3
    cam = d.camera
    assert cam.filter == capi.FILTER_BOX and cam.max_total_depth == 8 and cam.max_diff_depth == 8


def test_generators_are_deterministic():
    for make in (lambda: scenes.hall("principled", 64, 36, floor_res=16, n_columns=4, col_seg=8, col_rings=4),
                 lambda: scenes.instanced(9, 200, 32, 32), lambda: scenes.material_zoo(32, 24), scenes.cornell_box):
        a, b = make(), make()
        assert len(a.meshes) == len(b.meshes)
        for ma, mb in zip(a.meshes, b.meshes):
            assert ma.attrs.tobytes() == mb.attrs.tobytes() and ma.indices.tobytes() == mb.indices.tobytes()
        assert bytes(a.camera) == bytes(b.camera)


def test_cornell_box_is_the_samples_00_basic_scene():
    d = scenes.cornell_box()
    assert d.triangle_count() == 32 and len(d.meshes[0].attrs) == 64
    assert d.meshes[0].groups[1][2:] == (19, 6)  # the sample's own (odd) group offsets
    assert abs(d.camera.fov - 39.1463) < 1e-4 and d.camera.filter == capi.FILTER_BLACKMAN_HARRIS


def test_principled_variant_has_a_deep_light_tree_workload():
    d = scenes.hall("principled", 64, 36, floor_res=8, n_columns=2, col_seg=6, col_rings=4)
    emissive = [m for k, m in d.materials if k == "node" and m.type == capi.NODE_EMISSIVE]
    assert len(emissive) == 2
    assert sum(1 for g in d.meshes[0].groups if g[1] == capi.RS_INVALID) >= 2
