"""Dev tool: where does an end-to-end step (scene re-upload + one sample + read-back) spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_b200 import host, scenes

w, h = 1920, 1080
desc = scenes.hall("diffuse", w, h)
r = host.Renderer(w, h)
s = scenes.build(desc, r.create_scene())
it = 0
for _ in range(3):
    it = r.render(s, (0, 0, w, h), it, 1)
t = {"invalidate+render": 0.0, "render_only": 0.0, "pixels": 0.0}
n = 6
for _ in range(n):
    t0 = time.perf_counter()
    it = r.render(s, (0, 0, w, h), it, 1)
    t1 = time.perf_counter()
    r.invalidate_scene()
    it = r.render(s, (0, 0, w, h), it, 1)
    t2 = time.perf_counter()
    img = r.pixels(host.RAW)
    t3 = time.perf_counter()
    t["render_only"] += t1 - t0
    t["invalidate+render"] += t2 - t1
    t["pixels"] += t3 - t2
print({k: round(v / n * 1e3, 2) for k, v in t.items()}, "ms per call; pinned =", "off" if os.environ.get("RAY_HOST_NO_PINNED") else "on")
