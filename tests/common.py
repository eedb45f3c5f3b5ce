"""Shared helpers of the parity tests: build one scene description into both the oracle and the CUDA context."""
import numpy as np

from ray_b200 import capi, cuda, scenes


class Pair:
    """Oracle scene (reference Cpu::Scene, wide BVH) + CUDA context holding byte-identical arrays (oracle mode 1b)."""

    def __init__(self, oracle, desc, device=0, tex_compression=False):
        self.oracle = oracle
        self.desc = desc
        self.w, self.h = desc.width, desc.height
        self.osc = scenes.build(desc, oracle.Scene(wide=True, tex_compression=tex_compression))
        self.cam = self.osc.camera()
        self.ctx = cuda.Context(device)
        self.ctx.resize(self.w, self.h)
        ft = self.osc.filter_table() if self.cam.filter != capi.FILTER_BOX else None
        self.ctx.upload_tables(oracle.pmj_table(), ft)
        self.view = self.osc.view()
        self.ctx.upload_scene(self.view)

    def make_pass(self, iteration, rect=None, flags=0):
        return self.ctx.make_pass(self.cam, rect or (0, 0, self.w, self.h), iteration, flags)

    def close(self):
        self.ctx.close()
        self.osc.close()


def by_xy(a):
    """Sort a ray / shadow-ray record array by its pixel key (one record per pixel per stage)."""
    order = np.argsort(a["xy"], kind="stable")
    return a[order]


def bits_equal(a, b):
    """Exact (bitwise) equality of two structured / float arrays, NaNs included."""
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def field_mismatch(a, b):
    """Per-field count of records that differ bitwise (diagnostics)."""
    out = {}
    for name in a.dtype.names:
        x = np.ascontiguousarray(a[name]).view(np.uint8).reshape(len(a), -1)
        y = np.ascontiguousarray(b[name]).view(np.uint8).reshape(len(b), -1)
        n = int((x != y).any(axis=1).sum())
        if n:
            out[name] = n
    return out
