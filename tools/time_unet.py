"""Dev tool: time the UNet denoiser (RendererBase::DenoiseImage(pass, region), 16 passes) at 1920x1080 on both arithmetic
paths and print the tensor-core path's TFLOP/s against MEASURED_PEAKS.json.  Needs the oracle only for the weight set."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
from ray_b200 import capi, host, scenes

w, h = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
r = host.Renderer(w, h)
s = scenes.build(scenes.hall("diffuse", w, h), r.create_scene())
it = r.render(s, (0, 0, w, h), 0, 8)
layers = oracle.unet_layers()
wr, hr = (w + 15) // 16 * 16, (h + 15) // 16 * 16
shape = [(9, 32, 0), (32, 32, 0), (32, 48, 1), (48, 64, 2), (64, 80, 3), (80, 96, 4), (96, 96, 4), (160, 112, 3), (112, 112, 3),
         (160, 96, 2), (96, 96, 2), (128, 64, 1), (64, 64, 1), (73, 64, 0), (64, 32, 0), (32, 3, 0)]
flops = sum(2.0 * 9 * ci * co * (wr >> lv) * (hr >> lv) for ci, co, lv in shape)
out = {}
paths = (("fp32", capi.RC_UNET_FP32), ("tensor_cores", capi.RC_UNET_TENSOR_CORES))
if os.environ.get("TC_ONLY"):
    paths = paths[1:]
for name, flag in paths:
    r.set_unet_weights(layers, flag)
    r.denoise_unet((0, 0, w, h), it)  # warm-up (allocations, tensor maps)
    n = 5
    t0 = time.perf_counter()
    for _ in range(n):
        r.denoise_unet((0, 0, w, h), it)
    dt = (time.perf_counter() - t0) / n
    out[name] = {"ms": dt * 1e3, "tflops": flops / dt / 1e12}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = peaks.get("bf16_tflops_sustained")
out["network_gflop"] = flops / 1e9
out["frac_of_measured_bf16_sustained"] = (out["tensor_cores"]["tflops"] / peak) if peak else None
print(json.dumps(out))
