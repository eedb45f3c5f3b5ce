"""Multi-GPU sharding of the render: image strips per rank + ONE framebuffer gather per sample batch.

The reference has no multi-device story; the path shards trivially because a pixel's estimate
depends only on (x, y, iteration, scene) (section 8(e)): the RNG is keyed by the absolute pixel coordinate
(reference internal/CoreRef.cpp:1477-1480), jitter is sampled inside the pixel, and every ray carries its `xy`.  So each
rank renders a horizontal strip of the frame through the ordinary RegionContext{rect} API into its own replica of the
scene, with no communication between bounces, and the only exchange is a gather of the accumulated strips to rank 0.
This module is backend-agnostic torch.distributed code (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import List, Tuple


def strip_rect(rank: int, world: int, w: int, h: int) -> Tuple[int, int, int, int]:
    """Rect (x, y, w, h) of the rows rank `rank` owns: contiguous strips, heights differ by at most one row."""
    base, rem = divmod(h, world)
    y0 = rank * base + min(rank, rem)
    rows = base + (1 if rank < rem else 0)
    return (0, y0, w, rows)


def all_rects(world: int, w: int, h: int) -> List[Tuple[int, int, int, int]]:
    return [strip_rect(r, world, w, h) for r in range(world)]


def gather_strips(local_strip, w: int, h: int, dst: int = 0):
    """Gather every rank's (rows_r, w, 4) float32 strip to `dst` and assemble the (h, w, 4) frame there.

    Strip heights may differ by one row, so strips are padded to the maximum height for the collective (NCCL gather
    needs equal shapes) and cropped on arrival.  Returns the assembled frame on `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    rects = all_rects(world, w, h)
    max_rows = max(r[3] for r in rects)
    rows = rects[rank][3]
    assert tuple(local_strip.shape) == (rows, w, 4), (tuple(local_strip.shape), (rows, w, 4))
    if rows < max_rows:
        pad = torch.zeros((max_rows - rows, w, 4), dtype=local_strip.dtype, device=local_strip.device)
        send = torch.cat([local_strip, pad], dim=0).contiguous()
    else:
        send = local_strip.contiguous()
    if rank == dst:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, gather_list=parts, dst=dst)
        frame = torch.empty((h, w, 4), dtype=send.dtype, device=send.device)
        for r, (x, y, ww, hh) in enumerate(rects):
            frame[y:y + hh] = parts[r][:hh]
        return frame
    dist.gather(send, gather_list=None, dst=dst)
    return None


class DeviceImage:
    """Zero-copy torch view of one of the renderer's device frame buffers (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, h: int, w: int):
        self.__cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def device_frame_tensor(renderer, which, device_index: int):
    """torch.Tensor aliasing frame buffer `which` (ray_b200.capi.RC_BUF_*) of a ray_b200.host.Renderer."""
    import torch

    from . import cuda

    lib = cuda.load_library()
    ptr = lib.rc_device_ptr(renderer.native_context(), which)
    if not ptr:
        raise RuntimeError("rc_device_ptr returned NULL")
    return torch.as_tensor(DeviceImage(ptr, renderer.hh, renderer.w), device=torch.device("cuda", device_index))
