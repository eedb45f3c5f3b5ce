"""Backend-neutral scene descriptions and the synthetic workloads of BASELINE.json.

A SceneDesc is a plain list of the calls a user of the reference would make on a SceneBase
(AddMaterial / AddMesh / AddMeshInstance / AddLight / SetEnvironment / AddCamera, reference SceneBase.h:371-516),
expressed with the flat descriptors of include/ray_scene_desc.h.  `build(desc, backend)` replays them on any object
exposing the same verbs -- the product's host layer (ray_b200.host.Scene) or, in tests only, the oracle wrapper.

Workloads:
  cornell_box()      C1: the scene of the reference's samples/00_basic (256x256, default camera depths)
  hall(...)          C2/C3: "hall-250k" -- sinusoidal height-field floor + lathe columns + walls + emissive ceiling quad
  instanced(...)     C5 (scaled): many instances of one BLAS under a TLAS, non-uniform transforms
  material_zoo()     small scene touching every shading node and analytic light type (parity tests)
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from . import capi

IDENTITY = np.eye(4, dtype=np.float32)


@dataclass
class MeshDesc:
    attrs: np.ndarray  # (n_verts, 8) float32: position(3) normal(3) uv(2)
    indices: np.ndarray  # uint32
    groups: List[Tuple[int, int, int, int]]  # (front_mat, back_mat, first_index, index_count)
    allow_spatial_splits: bool = False
    use_fast_bvh_build: bool = False


@dataclass
class SceneDesc:
    name: str = "scene"
    width: int = 256
    height: int = 256
    materials: list = field(default_factory=list)  # ("node", rs_shading_node_desc) | ("principled", rs_principled_mat_desc)
    meshes: List[MeshDesc] = field(default_factory=list)
    instances: list = field(default_factory=list)  # (mesh_index, xform(4x4 column-major as 16 floats), vis dict)
    lights: list = field(default_factory=list)  # (kind, desc)
    env_col: tuple = (0.0, 0.0, 0.0)
    back_col: tuple = (0.0, 0.0, 0.0)
    env_importance_sample: bool = True
    env_map: int = capi.RS_INVALID  # index into `textures` of an RGBE lat-long map (RGBA8888), or RS_INVALID
    back_map: int = capi.RS_INVALID
    env_map_rotation: float = 0.0
    back_map_rotation: float = 0.0
    camera: capi.rs_camera_desc = None
    textures: list = field(default_factory=list)  # (uint8 array (h, w, c), flags dict); material descs refer by index

    def add_texture(self, pixels, **flags):
        self.textures.append((np.ascontiguousarray(pixels, dtype=np.uint8), flags))
        return len(self.textures) - 1

    def add_node(self, **kw):
        self.materials.append(("node", capi.rs_shading_node_desc.default(**kw)))
        return len(self.materials) - 1

    def add_principled(self, **kw):
        self.materials.append(("principled", capi.rs_principled_mat_desc.default(**kw)))
        return len(self.materials) - 1

    def triangle_count(self):
        return sum(len(m.indices) // 3 for m in self.meshes)


def build(desc: SceneDesc, backend):
    """Replay `desc` on `backend`; returns the backend (finalized)."""
    tex_ids = [backend.add_texture(px, **flags) for px, flags in desc.textures]
    backend.set_environment(desc.env_col, desc.back_col, desc.env_importance_sample,
                            env_map=tex_ids[desc.env_map] if desc.env_map != capi.RS_INVALID else capi.RS_INVALID,
                            back_map=tex_ids[desc.back_map] if desc.back_map != capi.RS_INVALID else capi.RS_INVALID,
                            env_map_rotation=desc.env_map_rotation, back_map_rotation=desc.back_map_rotation)

    def with_textures(d, fields):
        """texture fields hold indices into desc.textures -> translate to backend handles (on a copy)"""
        d2 = type(d).from_buffer_copy(d)
        for f in fields:
            v = getattr(d, f)
            if v != capi.RS_INVALID:
                setattr(d2, f, tex_ids[v])
        return d2

    mat_ids = []
    for kind, d in desc.materials:
        if kind == "node":
            d = with_textures(d, ("base_texture", "normal_map", "roughness_texture", "metallic_texture"))
            # mix children refer to indices in desc.materials -> translate to backend handles
            if d.type == capi.NODE_MIX:
                d.mix_materials[0] = mat_ids[d.mix_materials[0]]
                d.mix_materials[1] = mat_ids[d.mix_materials[1]]
            mat_ids.append(backend.add_material_node(d))
        else:
            d = with_textures(d, ("base_texture", "metallic_texture", "specular_texture", "roughness_texture",
                                  "emission_texture", "alpha_texture", "normal_map"))
            mat_ids.append(backend.add_material_principled(d))
    mesh_ids = []
    for m in desc.meshes:
        groups = [(mat_ids[f], mat_ids[b] if b != capi.RS_INVALID else capi.RS_INVALID, s, c) for f, b, s, c in m.groups]
        mesh_ids.append(backend.add_mesh(m.attrs, m.indices, groups, m.allow_spatial_splits, m.use_fast_bvh_build))
    for mesh, xform, vis in desc.instances:
        backend.add_mesh_instance(mesh_ids[mesh], np.asarray(xform, dtype=np.float32).reshape(16), **vis)
    for kind, d in desc.lights:
        backend.add_light(kind, d)
    backend.add_camera(desc.camera)
    backend.finalize()
    return backend


# ---------------------------------------------------------------------------------------------------------------------
def _quad(p, n, idx, uv=None):
    """4 vertices (position list p, shared normal n) -> (4,8) rows."""
    uv = uv or [(0.0, 0.0)] * 4
    return [list(p[i]) + list(n) + list(uv[i]) for i in range(4)], idx


def cornell_box(width=256, height=256) -> SceneDesc:
    """The Cornell box of the reference's samples/00_basic/main.cpp:26-190: same vertices, winding, material groups
    (including that sample's group offsets 19/25/31/37, which land on triangles 6,7 / 8,9 / 10,11 / 12.. by integer
    division), camera and default depth limits."""
    s = SceneDesc(name="cornell", width=width, height=height)
    grey = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.5, 0.5, 0.5))
    red = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.5, 0.0, 0.0))
    green = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.0, 0.5, 0.0))
    light = s.add_node(type=capi.NODE_EMISSIVE, strength=100.0, importance_sample=1)

    A, B, C_, D = [0, 2, 1, 0, 3, 2], [0, 1, 2, 0, 2, 3], [0, 1, 2, 1, 3, 2], [0, 1, 2, 2, 1, 3]
    quads = [
        _quad([(0.0, 0.0, -0.5592), (0.0, 0.0, 0.0), (-0.5528, 0.0, 0.0), (-0.5496, 0.0, -0.5592)], (0, 1, 0), A,
              [(1.0, 1.0), (1.0, 0.0), (0.0, 0.0), (0.0, 1.0)]),  # floor
        _quad([(0.0, 0.0, -0.5592), (-0.5496, 0.0, -0.5592), (-0.556, 0.5488, -0.5592), (0.0, 0.5488, -0.5592)],
              (0, 0, 1), A),  # back wall
        _quad([(-0.556, 0.5488, -0.5592), (0.0, 0.5488, -0.5592), (0.0, 0.5488, 0.0), (-0.556, 0.5488, 0.0)],
              (0, -1, 0), B),  # ceiling
        _quad([(-0.5528, 0.0, 0.0), (-0.5496, 0.0, -0.5592), (-0.556, 0.5488, 0.0), (-0.556, 0.5488, -0.5592)],
              (1, 0, 0), C_),  # left wall
        _quad([(0.0, 0.0, -0.5592), (0.0, 0.0, 0.0), (0.0, 0.5488, -0.5592), (0.0, 0.5488, 0.0)], (-1, 0, 0), D),  # right
        _quad([(-0.213, 0.5478, -0.227), (-0.343, 0.5478, -0.227), (-0.343, 0.5478, -0.332), (-0.213, 0.5478, -0.332)],
              (0, -1, 0), B),  # light
    ]
    # short block: corners (x,z) a,b,c,d and the 4 side normals, top at y=0.165
    sa, sb, sc_, sd = (-0.240464, -0.271646), (-0.082354, -0.224464), (-0.129536, -0.066354), (-0.287646, -0.113536)
    n1, n2 = 0.285951942, 0.958243966
    hs = 0.165

    def side(p, q, n, idx, h):
        return _quad([(p[0], 0.0, p[1]), (p[0], h, p[1]), (q[0], h, q[1]), (q[0], 0.0, q[1])], n, idx)

    quads += [
        side(sa, sb, (n1, 0.0, -n2), B, hs),
        side(sa, sd, (-n2, 0.0, -n1), A, hs),
        side(sb, sc_, (n2, 0.0, n1), B, hs),
        side(sd, sc_, (-n1, 0.0, n2), A, hs),
        _quad([(sa[0], hs, sa[1]), (sb[0], hs, sb[1]), (sc_[0], hs, sc_[1]), (sd[0], hs, sd[1])], (0, 1, 0), A),
    ]
    ta, tb, tc, td = (-0.471239, -0.405353), (-0.313647, -0.454239), (-0.264761, -0.296647), (-0.422353, -0.247761)
    m1, m2 = 0.296278358, 0.955101609
    ht = 0.33
    quads += [
        side(ta, tb, (-m1, 0.0, -m2), B, ht),
        side(tc, tb, (m2, 0.0, -m1), A, ht),
        side(ta, td, (-m2, 0.0, m1), A, ht),
        side(td, tc, (m1, 0.0, m2), A, ht),
        _quad([(ta[0], ht, ta[1]), (tb[0], ht, tb[1]), (tc[0], ht, tc[1]), (td[0], ht, td[1])], (0, 1, 0), A),
    ]
    rows, indices = [], []
    for q, idx in quads:
        base = len(rows)
        rows += q
        indices += [base + i for i in idx]
    attrs = np.asarray(rows, dtype=np.float32)
    inv = capi.RS_INVALID
    groups = [(grey, grey, 0, 18), (red, red, 19, 6), (green, green, 25, 6), (light, inv, 31, 6), (grey, grey, 37, 60)]
    s.meshes.append(MeshDesc(attrs, np.asarray(indices, dtype=np.uint32), groups))
    s.instances.append((0, IDENTITY.T.reshape(16), {}))
    s.camera = capi.rs_camera_desc.default(origin=(-0.278, 0.273, 0.8), fwd=(0.0, 0.0, -1.0), fov=39.1463)
    return s


# ---------------------------------------------------------------------------------------------------------------------
def _grid_mesh(nx, nz, pos_fn, flip=False):
    """(nx x nz) quads over parameters u,v in [0,1]; pos_fn(u, v) -> (x,y,z) arrays. Returns attrs (with smooth normals
    from the parametric derivative via finite differences) and indices."""
    u = np.linspace(0.0, 1.0, nx + 1, dtype=np.float64)
    v = np.linspace(0.0, 1.0, nz + 1, dtype=np.float64)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    p = np.stack(pos_fn(uu, vv), axis=-1)
    eps = 1e-4
    du = (np.stack(pos_fn(uu + eps, vv), axis=-1) - np.stack(pos_fn(uu - eps, vv), axis=-1))
    dv = (np.stack(pos_fn(uu, vv + eps), axis=-1) - np.stack(pos_fn(uu, vv - eps), axis=-1))
    n = np.cross(dv, du)
    n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-20)
    if flip:
        n = -n
    attrs = np.concatenate([p, n, np.stack([uu, vv], axis=-1)], axis=-1).reshape(-1, 8).astype(np.float32)
    i0 = (np.arange(nz)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    a, b, c, d = i0, i0 + 1, i0 + nx + 2, i0 + nx + 1
    if flip:
        tri = np.stack([a, b, c, a, c, d], axis=-1)
    else:
        tri = np.stack([a, c, b, a, d, c], axis=-1)
    return attrs, tri.reshape(-1).astype(np.uint32)


def _flat_quad(p0, p1, p2, p3, n):
    rows = [list(p) + list(n) + [u, v] for p, (u, v) in zip((p0, p1, p2, p3), ((0, 0), (1, 0), (1, 1), (0, 1)))]
    return np.asarray(rows, dtype=np.float32), np.asarray([0, 1, 2, 0, 2, 3], dtype=np.uint32)


def hall(variant="diffuse", width=1920, height=1080, floor_res=256, n_columns=60, col_seg=32, col_rings=32,
         extra_lights=None, seed=1337) -> SceneDesc:
    """"hall-250k" (SURVEY.md section 8(d) C2/C3): 20 m x 8 m hall, 4.5 m high.

    floor   : sinusoidal height field, floor_res^2 quads                       (256 -> 131,072 tris)
    columns : n_columns lathe profiles, col_seg x col_rings quads each         (60x32x32 -> 122,880 tris)
    shell   : 4 walls + ceiling (5 quads) and one 12 m x 2 m emissive ceiling panel, strength 30, importance sampled
    variant : "diffuse"    grey / red Diffuse nodes                            (config #2)
              "principled" principled dielectric floor/walls + metallic red columns + `extra_lights` (default 64)
                           small emissive quads so the light tree is >= 2 levels deep (config #3)
    One mesh, one identity instance, black environment, Box filter, all depth limits 8, min_total_depth 2.
    """
    s = SceneDesc(name=f"hall-{variant}", width=width, height=height)
    rng = np.random.RandomState(seed)
    LX, LZ, H = 20.0, 8.0, 4.5
    if variant == "diffuse":
        m_floor = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.6, 0.6, 0.6))
        m_wall = m_floor
        m_col = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.6, 0.1, 0.1))
        if extra_lights is None:
            extra_lights = 0
    elif variant == "principled":
        m_floor = s.add_principled(base_color=(0.6, 0.6, 0.6), roughness=0.4, specular=0.5)
        m_wall = s.add_principled(base_color=(0.55, 0.55, 0.6), roughness=0.6, specular=0.5)
        m_col = s.add_principled(base_color=(0.7, 0.12, 0.1), metallic=1.0, roughness=0.2)
        if extra_lights is None:
            extra_lights = 64
    else:
        raise ValueError(variant)
    m_light = s.add_node(type=capi.NODE_EMISSIVE, strength=30.0, base_color=(1.0, 0.95, 0.9), importance_sample=1)
    m_small = s.add_node(type=capi.NODE_EMISSIVE, strength=60.0, base_color=(0.9, 0.95, 1.0), importance_sample=1)

    parts, groups = [], []

    def add_part(attrs, idx, mat, back=None):
        parts.append((attrs, idx))
        groups.append((mat, mat if back is None else back, len(idx)))

    # floor: y = 0.06 sin(1.7 x) cos(2.3 z) + 0.03 sin(5.1 x + 1.3 z)
    def floor_fn(u, v):
        x = (u - 0.5) * LX
        z = (v - 0.5) * LZ
        y = 0.06 * np.sin(1.7 * x) * np.cos(2.3 * z) + 0.03 * np.sin(5.1 * x + 1.3 * z)
        return x, y, z

    fa, fi = _grid_mesh(floor_res, floor_res, floor_fn)
    add_part(fa, fi, m_floor)

    # columns: two rows along x, lathe radius r(t) with a base, a waist and a capital
    n_per_row = max(n_columns // 2, 1)
    for c in range(n_columns):
        row, k = divmod(c, n_per_row)
        cx = (k + 0.5) / n_per_row * (LX - 2.0) - (LX - 2.0) * 0.5
        cz = -2.2 if row == 0 else 2.2
        phase = rng.uniform(0.0, 2.0 * np.pi)

        def col_fn(u, v, cx=cx, cz=cz, phase=phase):
            t = v
            r = 0.16 + 0.05 * np.exp(-((t - 0.03) / 0.04) ** 2) + 0.06 * np.exp(-((t - 0.97) / 0.05) ** 2) \
                + 0.012 * np.sin(18.0 * t + phase)
            ang = 2.0 * np.pi * u
            flute = 1.0 + 0.04 * np.cos(12.0 * ang)
            return cx + r * flute * np.cos(ang), t * (H - 0.02) + 0.01, cz + r * flute * np.sin(ang)

        ca, ci = _grid_mesh(col_seg, col_rings, col_fn, flip=True)
        add_part(ca, ci, m_col)

    hx, hz = LX * 0.5, LZ * 0.5
    shell = [
        ((-hx, 0, -hz), (hx, 0, -hz), (hx, H, -hz), (-hx, H, -hz), (0, 0, 1)),  # back (z = -hz)
        ((hx, 0, hz), (-hx, 0, hz), (-hx, H, hz), (hx, H, hz), (0, 0, -1)),  # front
        ((-hx, 0, hz), (-hx, 0, -hz), (-hx, H, -hz), (-hx, H, hz), (1, 0, 0)),  # left
        ((hx, 0, -hz), (hx, 0, hz), (hx, H, hz), (hx, H, -hz), (-1, 0, 0)),  # right
        ((-hx, H, -hz), (hx, H, -hz), (hx, H, hz), (-hx, H, hz), (0, -1, 0)),  # ceiling
    ]
    for p0, p1, p2, p3, n in shell:
        qa, qi = _flat_quad(p0, p1, p2, p3, n)
        add_part(qa, qi, m_wall)
    # emissive ceiling panel 12 m x 2 m, 2 cm below the ceiling, facing down
    y = H - 0.02
    qa, qi = _flat_quad((-6, y, -1), (6, y, -1), (6, y, 1), (-6, y, 1), (0, -1, 0))
    add_part(qa, qi, m_light, back=capi.RS_INVALID)
    # optional small emissive quads on the walls (light-tree depth)
    lrng = np.random.RandomState(7)
    for i in range(extra_lights):
        side = i % 2
        x = lrng.uniform(-hx + 0.5, hx - 0.5)
        yy = lrng.uniform(2.0, 3.8)
        z = (-hz + 0.02) if side == 0 else (hz - 0.02)
        n = (0, 0, 1) if side == 0 else (0, 0, -1)
        w, h = 0.15, 0.1
        if side == 0:
            qa, qi = _flat_quad((x - w, yy - h, z), (x + w, yy - h, z), (x + w, yy + h, z), (x - w, yy + h, z), n)
        else:
            qa, qi = _flat_quad((x + w, yy - h, z), (x - w, yy - h, z), (x - w, yy + h, z), (x + w, yy + h, z), n)
        add_part(qa, qi, m_small, back=capi.RS_INVALID)

    attrs = np.concatenate([a for a, _ in parts], axis=0)
    idx_parts, mesh_groups, voff, ioff = [], [], 0, 0
    for (a, i), (front, back, n_idx) in zip(parts, groups):
        idx_parts.append(i + np.uint32(voff))
        mesh_groups.append((front, back, ioff, n_idx))
        voff += len(a)
        ioff += n_idx
    # merge consecutive groups with the same materials (keeps the group list short)
    merged = []
    for g in mesh_groups:
        if merged and merged[-1][0] == g[0] and merged[-1][1] == g[1] and merged[-1][2] + merged[-1][3] == g[2]:
            merged[-1] = (g[0], g[1], merged[-1][2], merged[-1][3] + g[3])
        else:
            merged.append(g)
    s.meshes.append(MeshDesc(np.ascontiguousarray(attrs), np.concatenate(idx_parts).astype(np.uint32), merged))
    s.instances.append((0, IDENTITY.T.reshape(16), {}))
    fwd = np.array([0.97, -0.05, -0.24])
    fwd /= np.linalg.norm(fwd)
    s.camera = capi.rs_camera_desc.default(origin=(-9.0, 1.7, 2.5), fwd=tuple(fwd.astype(np.float32)), fov=60.0,
                                           filter=capi.FILTER_BOX, max_diff_depth=8, max_spec_depth=8,
                                           max_refr_depth=8, max_transp_depth=8, max_total_depth=8,
                                           min_total_depth=2)
    return s


def instanced(n_instances=64, tris_per_blas=2000, width=512, height=512, seed=99) -> SceneDesc:
    """Config #5 in miniature: `n_instances` transformed copies (rotation + non-uniform scale + translation) of one
    bumpy sphere BLAS over a ground quad, one emissive panel; exercises TLAS traversal and instance transforms."""
    s = SceneDesc(name="instanced", width=width, height=height)
    rng = np.random.RandomState(seed)
    m_a = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.7, 0.7, 0.7))
    m_g = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.4, 0.45, 0.5))
    m_l = s.add_node(type=capi.NODE_EMISSIVE, strength=25.0, importance_sample=1)
    res = max(int(np.sqrt(tris_per_blas / 2)), 4)

    def sph(u, v):
        th = np.pi * (0.02 + 0.96 * v)
        ph = 2 * np.pi * u
        r = 0.5 * (1.0 + 0.08 * np.sin(6 * ph) * np.sin(5 * th))
        return r * np.sin(th) * np.cos(ph), r * np.cos(th), r * np.sin(th) * np.sin(ph)

    sa, si = _grid_mesh(res, res, sph, flip=True)
    s.meshes.append(MeshDesc(sa, si, [(m_a, m_a, 0, len(si))]))
    side = int(np.ceil(np.sqrt(n_instances)))
    g = max(12.0, 0.8 * side + 2.0)  # ground half-size: covers the instance grid (12 m for the small test scenes)
    ls = 3.0 * g / 12.0
    ga, gi = _flat_quad((-g, 0, -g), (-g, 0, g), (g, 0, g), (g, 0, -g), (0, 1, 0))
    la, li = _flat_quad((-ls, 7, -ls), (ls, 7, -ls), (ls, 7, ls), (-ls, 7, ls), (0, -1, 0))
    s.meshes.append(MeshDesc(np.concatenate([ga, la]), np.concatenate([gi, li + np.uint32(4)]),
                             [(m_g, m_g, 0, 6), (m_l, capi.RS_INVALID, 6, 6)]))
    for i in range(n_instances):
        gx, gz = i % side, i // side
        ang = rng.uniform(0, 2 * np.pi)
        sc = rng.uniform(0.5, 1.3, size=3)
        c, sn = np.cos(ang), np.sin(ang)
        R = np.array([[c, 0, sn], [0, 1, 0], [-sn, 0, c]])
        M = np.eye(4)
        M[:3, :3] = R @ np.diag(sc)
        M[:3, 3] = [(gx - side / 2 + 0.5) * 1.6, 0.5 * sc[1] + 0.02, (gz - side / 2 + 0.5) * 1.6]
        s.instances.append((0, M.T.astype(np.float32).reshape(16), {}))  # column-major
    s.instances.append((1, IDENTITY.T.reshape(16), {}))
    fwd = np.array([0.0, -0.45, -1.0])
    fwd /= np.linalg.norm(fwd)
    s.camera = capi.rs_camera_desc.default(origin=(0.0, 6.0 * g / 12.0, g), fwd=tuple(fwd.astype(np.float32)), fov=50.0,
                                           filter=capi.FILTER_BOX, max_diff_depth=8, max_total_depth=8)
    return s


def c5_instanced(width=4096, height=4096) -> SceneDesc:
    """BASELINE.json config #5: 10 M instanced triangles = a TLAS over 1,000 instances (rotation + non-uniform scale,
    seed 99) of one 10,000-triangle BLAS, diffuse, 4096 x 4096 -- the traversal-divergence stress."""
    s = instanced(1000, 10000, width, height, seed=99)
    s.name = "c5-instanced"
    return s


def material_zoo(width=160, height=120, lights=("rect", "sphere", "dir", "spot", "disk", "line"), env=(0.0, 0.0, 0.0),
                 filter=capi.FILTER_BOX, fstop=0.0, transparent=True) -> SceneDesc:
    """A row of bumpy spheres over a ground plane, one per shading node family (diffuse, glossy, refractive, principled
    dielectric / metal / clearcoat+sheen / transmissive, mix, additive mix with emission, alpha-blended), lit by the
    requested analytic lights plus one emissive quad.  Small enough for the scalar oracle to finish in seconds."""
    s = SceneDesc(name="zoo", width=width, height=height)
    mats = [
        s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.7, 0.3, 0.2), roughness=0.5),
        s.add_node(type=capi.NODE_GLOSSY, base_color=(0.9, 0.8, 0.5), roughness=0.25),
        s.add_node(type=capi.NODE_REFRACTIVE, base_color=(0.9, 0.95, 1.0), roughness=0.05, ior=1.45),
        s.add_principled(base_color=(0.2, 0.5, 0.8), roughness=0.35, specular=0.5),
        s.add_principled(base_color=(0.9, 0.6, 0.2), metallic=1.0, roughness=0.15, anisotropic=0.6),
        s.add_principled(base_color=(0.5, 0.1, 0.1), roughness=0.5, clearcoat=1.0, clearcoat_roughness=0.1, sheen=0.8,
                         specular_tint=0.5),
        s.add_principled(base_color=(0.8, 0.9, 0.8), roughness=0.1, transmission=1.0, ior=1.5,
                         transmission_roughness=0.1),
        s.add_principled(base_color=(0.3, 0.3, 0.3), roughness=0.6, emission_color=(0.2, 0.8, 0.3),
                         emission_strength=2.0),
    ]
    d0 = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.1, 0.6, 0.1))
    g0 = s.add_node(type=capi.NODE_GLOSSY, base_color=(0.8, 0.8, 0.8), roughness=0.1)
    mats.append(s.add_node(type=capi.NODE_MIX, mix_materials=(d0, g0), strength=0.5, ior=1.5))
    if transparent:
        mats.append(s.add_principled(base_color=(0.8, 0.7, 0.1), roughness=0.4, alpha=0.5))
    m_ground = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.5, 0.5, 0.5))
    m_light = s.add_node(type=capi.NODE_EMISSIVE, strength=12.0, importance_sample=1)

    def sph(u, v):
        th = np.pi * (0.01 + 0.98 * v)
        ph = 2 * np.pi * u
        r = 0.45 * (1.0 + 0.05 * np.sin(4 * ph) * np.sin(3 * th))
        return r * np.sin(th) * np.cos(ph), r * np.cos(th), r * np.sin(th) * np.sin(ph)

    sa, si = _grid_mesh(14, 10, sph, flip=True)
    for m in mats:
        s.meshes.append(MeshDesc(sa, si, [(m, m, 0, len(si))]))
    ga, gi = _flat_quad((-8, 0, -6), (-8, 0, 6), (8, 0, 6), (8, 0, -6), (0, 1, 0))
    la, li = _flat_quad((-1.0, 3.2, -1.5), (1.0, 3.2, -1.5), (1.0, 3.2, -0.5), (-1.0, 3.2, -0.5), (0, -1, 0))
    s.meshes.append(MeshDesc(np.concatenate([ga, la]), np.concatenate([gi, li + np.uint32(4)]),
                             [(m_ground, m_ground, 0, 6), (m_light, capi.RS_INVALID, 6, 6)]))
    n = len(mats)
    for i in range(n):
        M = np.eye(4)
        M[:3, 3] = [(i - (n - 1) / 2) * 1.1, 0.5, 0.3 * np.sin(i * 1.3)]
        s.instances.append((i, M.T.astype(np.float32).reshape(16), {}))
    s.instances.append((n, IDENTITY.T.reshape(16), {}))

    def xf(pos, rot_x=0.0):
        c, sn = np.cos(rot_x), np.sin(rot_x)
        M = np.eye(4)
        M[:3, :3] = np.array([[1, 0, 0], [0, c, -sn], [0, sn, c]])
        M[:3, 3] = pos
        return tuple(M.T.astype(np.float32).reshape(16))

    LC = capi.rs_light_common.default
    if "rect" in lights:
        s.lights.append(("rect", capi.rs_rect_light_desc(c=LC(color=(8.0, 7.0, 6.0)), width=1.2, height=0.8,
                                                         doublesided=0, sky_portal=0, xform=xf((-3.0, 3.0, 1.0)))))
    if "sphere" in lights:
        s.lights.append(("sphere", capi.rs_sphere_light_desc(c=LC(color=(20.0, 20.0, 25.0)), position=(3.0, 2.5, 1.5),
                                                             radius=0.2)))
    if "dir" in lights:
        d = np.array([0.3, -1.0, -0.2])
        d /= np.linalg.norm(d)
        s.lights.append(("directional", capi.rs_directional_light_desc(c=LC(color=(0.6, 0.6, 0.5)),
                                                                       direction=tuple(d.astype(np.float32)),
                                                                       angle=2.0)))
    if "spot" in lights:
        d = np.array([0.2, -1.0, -0.3])
        d /= np.linalg.norm(d)
        s.lights.append(("spot", capi.rs_spot_light_desc(c=LC(color=(30.0, 25.0, 20.0)), position=(-1.0, 3.5, 2.0),
                                                         direction=tuple(d.astype(np.float32)), spot_size=50.0,
                                                         spot_blend=0.2, radius=0.1)))
    if "disk" in lights:
        s.lights.append(("disk", capi.rs_disk_light_desc(c=LC(color=(6.0, 8.0, 6.0)), size_x=1.0, size_y=0.7,
                                                         doublesided=0, sky_portal=0, xform=xf((1.5, 3.0, 2.0)))))
    if "line" in lights:
        s.lights.append(("line", capi.rs_line_light_desc(c=LC(color=(5.0, 5.0, 9.0)), radius=0.05, height=2.0,
                                                         sky_portal=0, xform=xf((0.0, 2.8, -2.5)))))
    s.env_col = env
    s.back_col = env
    fwd = np.array([0.0, -0.32, -1.0])
    fwd /= np.linalg.norm(fwd)
    s.camera = capi.rs_camera_desc.default(origin=(0.0, 2.6, 7.5), fwd=tuple(fwd.astype(np.float32)), fov=42.0,
                                           filter=filter, max_diff_depth=4, max_total_depth=8, fstop=fstop,
                                           focus_distance=7.5, focal_length=0.05 if fstop > 0 else 0.0)
    return s


def _checker(n, cells, a, b, channels=3):
    y, x = np.mgrid[0:n, 0:n]
    m = (((x * cells) // n + (y * cells) // n) & 1).astype(bool)
    img = np.where(m[..., None], np.asarray(a, np.uint8), np.asarray(b, np.uint8)).astype(np.uint8)
    return img if channels > 1 else img[..., 0]


def textured(width=96, height=72) -> SceneDesc:
    """Every place the reference fetches a texture on the path: sRGB base colour (RGB and
    RGBA storages), linear roughness / metallic / specular maps (R8), a tangent-space normal map (RG storage with
    reconstructed z, intensity < 1), an alpha-textured principled material (Mix with a Transparent node: the transparency
    loops of the closest-hit and shadow traces), a Mix node driven by a texture, and a textured emissive quad sampled
    through the light tree."""
    s = SceneDesc(name="textured", width=width, height=height)
    rng = np.random.RandomState(11)
    n = 64
    t_checker = s.add_texture(_checker(n, 8, (230, 220, 200), (40, 60, 120)), is_srgb=True)
    noise = rng.randint(0, 256, size=(n, n, 4)).astype(np.uint8)
    noise[..., 3] = 255
    t_rgba = s.add_texture(noise, is_srgb=True)
    yy, xx = np.mgrid[0:n, 0:n]
    t_rough = s.add_texture((40 + 180 * ((xx // 8) % 2)).astype(np.uint8), is_srgb=False)
    t_metal = s.add_texture((255 * ((yy // 16) % 2)).astype(np.uint8), is_srgb=False)
    t_spec = s.add_texture(rng.randint(64, 256, size=(n, n)).astype(np.uint8), is_srgb=True)
    nx = 0.35 * np.sin(xx * (2 * np.pi / 16.0))
    ny = 0.35 * np.cos(yy * (2 * np.pi / 12.0))
    nz = np.sqrt(np.maximum(1.0 - nx * nx - ny * ny, 0.0))
    nm = np.stack([(nx * 0.5 + 0.5) * 255, (ny * 0.5 + 0.5) * 255, (nz * 0.5 + 0.5) * 255], axis=-1)
    t_normal = s.add_texture(np.round(nm).astype(np.uint8), is_srgb=False, is_normalmap=True)
    disk = (((xx - n / 2) ** 2 + (yy - n / 2) ** 2) < (n * 0.38) ** 2)
    t_alpha = s.add_texture((255 * disk).astype(np.uint8), is_srgb=False)
    t_mix = s.add_texture(_checker(n, 4, (255,), (0,), channels=1), is_srgb=False)
    emit = np.zeros((n, n, 3), np.uint8)
    emit[..., 0] = 255 * ((xx // 16) % 2)
    emit[..., 1] = 200
    emit[..., 2] = 255 * ((yy // 16) % 2)
    t_emit = s.add_texture(emit, is_srgb=True)

    m_floor = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.9, 0.9, 0.9), base_texture=t_checker)
    mats = [
        s.add_principled(base_color=(1.0, 1.0, 1.0), base_texture=t_rgba, roughness=0.8, roughness_texture=t_rough),
        s.add_principled(base_color=(0.9, 0.7, 0.3), metallic=1.0, metallic_texture=t_metal, roughness=0.3,
                         specular=0.8, specular_texture=t_spec),
        s.add_principled(base_color=(0.3, 0.6, 0.3), roughness=0.4, normal_map=t_normal, normal_map_intensity=0.7),
        s.add_node(type=capi.NODE_GLOSSY, base_color=(0.9, 0.9, 0.9), roughness=0.5, roughness_texture=t_rough,
                   normal_map=t_normal),
    ]
    d0 = s.add_node(type=capi.NODE_DIFFUSE, base_color=(0.8, 0.2, 0.2))
    g0 = s.add_node(type=capi.NODE_GLOSSY, base_color=(0.9, 0.9, 0.9), roughness=0.05)
    mats.append(s.add_node(type=capi.NODE_MIX, mix_materials=(d0, g0), strength=1.0, base_texture=t_mix))
    m_cutout = s.add_principled(base_color=(0.9, 0.8, 0.2), base_texture=t_checker, roughness=0.5, alpha=1.0,
                                alpha_texture=t_alpha)
    m_light = s.add_node(type=capi.NODE_EMISSIVE, strength=14.0, base_color=(1.0, 1.0, 1.0), base_texture=t_emit,
                         importance_sample=1)

    def sph(u, v):
        th = np.pi * (0.01 + 0.98 * v)
        ph = 2 * np.pi * u
        r = 0.45
        return r * np.sin(th) * np.cos(ph), r * np.cos(th), r * np.sin(th) * np.sin(ph)

    sa, si = _grid_mesh(16, 12, sph, flip=True)
    sa = sa.copy()
    sa[:, 6] *= 3.0  # tile the textures 3x around the sphere
    sa[:, 7] *= 2.0
    for m in mats:
        s.meshes.append(MeshDesc(sa, si, [(m, m, 0, len(si))]))
    ga, gi = _flat_quad((-5, 0, -4), (-5, 0, 4), (5, 0, 4), (5, 0, -4), (0, 1, 0))
    ga = ga.copy()
    ga[:, 6:8] *= 2.5
    la, li = _flat_quad((-1.0, 3.0, -1.0), (1.0, 3.0, -1.0), (1.0, 3.0, 0.0), (-1.0, 3.0, 0.0), (0, -1, 0))
    ca, ci = _flat_quad((-1.2, 0.2, 1.6), (1.2, 0.2, 1.6), (1.2, 1.8, 1.6), (-1.2, 1.8, 1.6), (0, 0, 1))
    s.meshes.append(MeshDesc(np.concatenate([ga, la, ca]), np.concatenate([gi, li + np.uint32(4), ci + np.uint32(8)]),
                             [(m_floor, m_floor, 0, 6), (m_light, capi.RS_INVALID, 6, 6), (m_cutout, m_cutout, 12, 6)]))
    k = len(mats)
    for i in range(k):
        M = np.eye(4)
        M[:3, 3] = [(i - (k - 1) / 2) * 1.15, 0.5, 0.2 * np.cos(i * 1.7)]
        s.instances.append((i, M.T.astype(np.float32).reshape(16), {}))
    s.instances.append((k, IDENTITY.T.reshape(16), {}))
    s.lights.append(("sphere", capi.rs_sphere_light_desc(c=capi.rs_light_common.default(color=(15.0, 15.0, 18.0)),
                                                         position=(2.5, 2.5, 2.0), radius=0.15)))
    fwd = np.array([0.0, -0.30, -1.0])
    fwd /= np.linalg.norm(fwd)
    s.camera = capi.rs_camera_desc.default(origin=(0.0, 2.2, 6.5), fwd=tuple(fwd.astype(np.float32)), fov=40.0,
                                           filter=capi.FILTER_BOX, max_diff_depth=4, max_total_depth=8)
    return s


def rgbe_image(rgb):
    """float RGB (h, w, 3) -> RGBE bytes (h, w, 4), Ward's encoding (what the reference's rgbe_to_rgb decodes)."""
    rgb = np.asarray(rgb, dtype=np.float64)
    m = rgb.max(axis=-1)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, -128)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0.0)
    out = np.zeros(rgb.shape[:2] + (4,), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.clip(e + 128, 0, 255).astype(np.uint8)
    return out


def envmap_zoo(width=128, height=96, importance_sample=True, rotation=0.7) -> SceneDesc:
    """material_zoo lit by a lat-long RGBE environment map: a sky gradient with a small,
    very bright sun, so the importance-sampling quad-tree has several levels; the same map (rotated differently) is the
    background camera rays see; one rect light is a sky portal."""
    s = material_zoo(width, height, lights=("sphere",), env=(1.0, 1.0, 1.0), filter=capi.FILTER_BOX)
    s.name = "envmap_zoo"
    w, h = 128, 64
    v, u = np.mgrid[0:h, 0:w]
    theta = (v + 0.5) / h * np.pi
    phi = (u + 0.5) / w * 2 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], axis=-1)
    up = np.clip(d[..., 1], 0.0, 1.0)
    sky = np.stack([0.25 + 0.25 * up, 0.35 + 0.35 * up, 0.5 + 0.6 * up], axis=-1) * 0.6
    sky = np.where(d[..., 1:2] < 0.0, np.asarray([0.12, 0.10, 0.08]), sky)
    sun_dir = np.asarray([0.5, 0.65, 0.57])
    sun_dir = sun_dir / np.linalg.norm(sun_dir)
    sun = np.clip((d @ sun_dir - 0.985) / 0.015, 0.0, 1.0) ** 2
    img = sky + sun[..., None] * np.asarray([90.0, 80.0, 60.0])
    t_env = s.add_texture(rgbe_image(img), is_srgb=False)
    s.env_map = t_env
    s.back_map = t_env
    s.env_map_rotation = rotation
    s.back_map_rotation = rotation * 0.5
    s.env_importance_sample = importance_sample
    M = np.eye(4)
    M[:3, 3] = (-2.0, 3.0, 1.0)
    s.lights.append(("rect", capi.rs_rect_light_desc(c=capi.rs_light_common.default(color=(1.0, 1.0, 1.0)), width=1.5,
                                                     height=1.0, doublesided=0, sky_portal=1,
                                                     xform=tuple(M.T.astype(np.float32).reshape(16)))))
    return s
