"""Dev tool (not the product path): time the kernel families on the hall-250k workload using an oracle-built scene."""
import argparse
import json
import sys
import time
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ray_b200 import capi, scenes, cuda as _cuda
if os.environ.get("RC_DEV_CUDA_LIB"):  # A/B runs of an alternative build of the kernels (dev tool only)
    _cuda.LIB_PATH = os.path.abspath(os.environ["RC_DEV_CUDA_LIB"])
import hashlib
import oracle
from common import Pair

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="diffuse")
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--no-sort", action="store_true")
ap.add_argument("--scene", default="hall")
ap.add_argument("--max-total-depth", type=int, default=0)
args = ap.parse_args()

t0 = time.time()
if args.scene == "hall":
    desc = scenes.hall(args.variant, args.w, args.h)
elif args.scene == "cornell":
    desc = scenes.cornell_box(args.w, args.h)
else:
    desc = scenes.instanced(1000, 10000, args.w, args.h)
if args.max_total_depth:
    desc.camera.max_total_depth = args.max_total_depth
pair = Pair(oracle, desc)
print(f"scene {desc.name}: {desc.triangle_count()} tris, wnodes {pair.view.wnodes.count}, build {time.time()-t0:.1f}s", flush=True)
flags = capi.RC_RENDER_ASYNC | (capi.RC_RENDER_NO_SORT if args.no_sort else 0)
it = 0
for i in range(args.warmup):
    it += 1
    pair.ctx.render(pair.make_pass(it, flags=flags))
pair.ctx.sync()
pair.ctx.reset_stats()
t0 = time.time()
for i in range(args.spp):
    it += 1
    pair.ctx.render(pair.make_pass(it, flags=flags))
pair.ctx.sync()
dt = time.time() - t0
c = pair.ctx.counters()
k = pair.ctx.kernel_ms()
rays = c["primary_rays"] + c["secondary_rays"]
raw = pair.ctx.readback(capi.RC_BUF_RAW)
print(json.dumps({"lib": os.path.basename(_cuda.LIB_PATH), "fin_min": os.environ.get("RC_TRACE_FIN_MIN"),
                  "raw_sha1": hashlib.sha1(raw.tobytes()).hexdigest()[:16], "wall_s": dt, "ms_per_sample": dt / args.spp * 1e3, "Mrays_s": rays / dt / 1e6,
                  "Mshadow_s": c["shadow_rays"] / dt / 1e6, "counters": c, "kernel_ms": k,
                  "rays_per_sample": rays / args.spp, "nodes_per_ray": c["nodes_visited"] / max(rays + c["shadow_rays"], 1),
                  "leaves_per_ray": c["leaves_tested"] / max(rays + c["shadow_rays"], 1)}))
