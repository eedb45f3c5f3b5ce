"""Dev tool: time rc_denoise_nlm (RendererBase::DenoiseImage(region)) on a 1920x1080 hall render."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_b200 import host, scenes

w, h = 1920, 1080
r = host.Renderer(w, h)
s = scenes.build(scenes.hall("diffuse", w, h), r.create_scene())
it = r.render(s, (0, 0, w, h), 0, 8)
r.denoise((0, 0, w, h), it)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    r.denoise((0, 0, w, h), it)
dt = (time.perf_counter() - t0) / n
px = w * h
print(f"NLM 7x7/3x3 joint filter {w}x{h}: {dt * 1e3:.2f} ms per call, {px / dt / 1e6:.0f} Mpx/s, "
      f"{px * 49 * 9 * 4 * 10 / dt / 1e12:.2f} TFLOP/s-ish (49 x 9 x 4 lanes x ~10 flop)")
