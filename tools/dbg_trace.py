"""dev: primary-trace of the cornell golden fixture, print the records that differ from the fixture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ray_b200 import capi, cuda
from ray_b200.cuda import RAY_DTYPE, HIT_DTYPE
import test_golden as tg
path = [p for p in tg.GOLDEN if "cornell" in p][0]
g = np.load(path); keep = []
v = tg._view_from_golden(g, keep)
w, h = [int(x) for x in g["wh"]]
ctx = cuda.Context(0); ctx.resize(w, h)
ctx.upload_tables(np.zeros(32 * 4096 * 2, np.uint32), g["filter_table"]); ctx.upload_scene(v)
cam = capi.rc_camera.from_buffer_copy(g["cam"].tobytes())
p = ctx.make_pass(cam, (0, 0, w, h), int(g["iteration"]))
rays = g["primary_rays"].view(RAY_DTYPE)
_, hits = ctx.stage_trace_rays(p, rays, g["primary_hits_in"].view(HIT_DTYPE), False)
ref = g["primary_hits_out"].view(HIT_DTYPE)
bad = [i for i in range(len(ref)) if hits[i].tobytes() != ref[i].tobytes()]
print("records", len(ref), "differ", len(bad))
for i in bad[:12]:
    print(i, "got", hits[i], "ref", ref[i], "t bits", hex(hits[i]["t"].view(np.uint32)), hex(ref[i]["t"].view(np.uint32)), "ray o", rays[i]["o"], "d", rays[i]["d"])
nodes = np.frombuffer(keep[0].tobytes(), np.uint8) if False else None
print("wnodes", v.wnodes.count, "tlas_root", v.tlas_root, "instances", v.mesh_instances.count)
