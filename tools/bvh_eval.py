"""Dev tool: rank BVH builders by node / leaf visits on the bench workload (primary, diffuse-bounce and shadow rays)."""
import ctypes as C, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ray_b200 import capi, host, scenes

so = "/tmp/bvh_eval.so"
subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "bvh_eval.cpp"), "-o", so])
lib = C.CDLL(so)
lib.bvh_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

def run(view, o, d, tmax=None, any_hit=0):
    o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
    n = len(o); nn = C.c_uint64(0); nl = C.c_uint64(0); t = np.zeros(n, np.float32)
    tm = np.ascontiguousarray(tmax, np.float32) if tmax is not None else None
    lib.bvh_eval(C.byref(view), o.ctypes.data, d.ctypes.data, tm.ctypes.data if tm is not None else None, n, any_hit,
                 C.byref(nn), C.byref(nl), t.ctypes.data)
    return nn.value / n, nl.value / n, t

def rays_for(desc, w=240, h=135):
    cam = desc.camera
    org = np.array(list(cam.origin), np.float64); fwd = np.array(list(cam.fwd), np.float64); fwd /= np.linalg.norm(fwd)
    up = np.array([0, 1, 0.0]); side = np.cross(fwd, up); side /= np.linalg.norm(side); upv = np.cross(side, fwd)
    k = np.tan(np.radians(cam.fov) / 2)
    ys, xs = np.mgrid[0:h, 0:w]
    px = ((xs + 0.5) / w * 2 - 1) * k * (w / h); py = (1 - (ys + 0.5) / h * 2) * k
    d = fwd[None, None] + px[..., None] * side + py[..., None] * upv
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return np.tile(org, (w * h, 1)), d.reshape(-1, 3)

def evaluate(name, view, desc):
    o, d = rays_for(desc)
    pn, pl, t = run(view, o, d)
    hitm = t > 0
    P = o[hitm] + d[hitm] * t[hitm, None] * 0.999
    rng = np.random.RandomState(1)
    # diffuse bounce: uniform sphere directions flipped away from the incoming ray (cheap stand-in for a cosine lobe)
    v = rng.normal(size=P.shape); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v = np.where((np.sum(v * d[hitm], axis=-1) > 0)[:, None], -v, v)
    sn, sl, t2 = run(view, P, v)
    # shadow rays towards the ceiling panel (12 m x 2 m at y = 4.48)
    L = np.stack([rng.uniform(-6, 6, len(P)), np.full(len(P), 4.47), rng.uniform(-1, 1, len(P))], axis=-1)
    sd = L - P; dist = np.linalg.norm(sd, axis=-1); sd /= dist[:, None]
    hn, hl, _ = run(view, P, sd, dist, any_hit=1)
    cost = lambda a, b: a + 1.6 * b
    print(f"{name:28s} primary {pn:5.2f}/{pl:4.2f}  bounce {sn:5.2f}/{sl:4.2f}  shadow {hn:5.2f}/{hl:4.2f}   "
          f"cost(n + 1.6 l): {cost(pn, pl):5.2f} {cost(sn, sl):5.2f} {cost(hn, hl):5.2f}   nodes {view.wnodes.count} blocks {view.mtris.count}")

if __name__ == "__main__":
    desc = scenes.hall("diffuse", 1920, 1080)
    t0 = time.time(); hs = scenes.build(desc, host.Scene(None)); th = time.time() - t0
    evaluate(f"host builder ({th:.2f}s)", hs.view(), desc)
    if "--oracle" in sys.argv:
        import oracle
        t0 = time.time(); osc = scenes.build(desc, oracle.Scene(wide=True)); to = time.time() - t0
        evaluate(f"reference builder ({to:.2f}s)", osc.view(), desc)
